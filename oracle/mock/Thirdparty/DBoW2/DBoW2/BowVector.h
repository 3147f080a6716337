// oracle/mock stub (test infrastructure), see FeatureVector.h
#pragma once
#include <map>
namespace DBoW2 {
typedef unsigned int WordId;
typedef double WordValue;
class BowVector : public std::map<WordId, WordValue> {};
}
