// oracle/cvshim stub: forwards to the single shim header (test infrastructure, see cvshim.hpp)
#include "../../cvshim.hpp"
