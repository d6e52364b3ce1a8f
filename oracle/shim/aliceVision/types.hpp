// Oracle shim for <aliceVision/types.hpp> (src/aliceVision/types.hpp:19-24 without Eigen).
#pragma once
#include <cstdint>
#include <limits>
#include <map>
#include <set>
#include <vector>
namespace aliceVision {
typedef uint32_t IndexT;
static const IndexT UndefinedIndexT = std::numeric_limits<IndexT>::max();
typedef std::pair<IndexT, IndexT> Pair;
typedef std::set<Pair> PairSet;
typedef std::vector<Pair> PairVec;
}  // namespace aliceVision
