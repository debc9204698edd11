// DBoW3 stand-in: only the types the reference's headers name (bag of words is off the extractor / ExtendMapMatches path).
#pragma once
#include <map>
#include <string>
#include <vector>
namespace DBoW3 {
typedef unsigned int WordId;
typedef double WordValue;
typedef unsigned int NodeId;
class BowVector : public std::map<WordId, WordValue> {};
class FeatureVector : public std::map<NodeId, std::vector<unsigned int>> {};
class Vocabulary {
   public:
    template <typename M>
    void transform(const M&, BowVector&, FeatureVector&, int) const {}
    double score(const BowVector&, const BowVector&) const { return 0; }
    bool empty() const { return true; }
};
}  // namespace DBoW3
