// Oracle shim: RegionsPerView.hpp:12 includes this for nothing the matcher path uses.
#pragma once
