// Oracle/adaptor-test shim: the real pairBuilder.hpp pulls sfmData (Eigen, Boost).  IImageCollectionMatcher.hpp:13
// only needs PairSet from types.hpp.
#pragma once
#include <aliceVision/types.hpp>
