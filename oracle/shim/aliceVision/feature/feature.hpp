// Oracle shim for <aliceVision/feature/feature.hpp>: the real umbrella header drags in
// ImageDescriber.hpp -> image/OIIO. The matching headers only need the data model.
#pragma once
#include <aliceVision/feature/Descriptor.hpp>
#include <aliceVision/feature/PointFeature.hpp>
#include <aliceVision/feature/Regions.hpp>
#include <aliceVision/feature/regionsFactory.hpp>
