// oracle/mock stub (test infrastructure): DBoW2 is not vendored in the reference; only the
// container shape is needed to compile ORBmatcher.cc / Frame.cc (SearchByBoW is out of scope).
#pragma once
#include <map>
#include <vector>
namespace DBoW2 {
typedef unsigned int NodeId;
class FeatureVector : public std::map<NodeId, std::vector<unsigned int> > {};
}
