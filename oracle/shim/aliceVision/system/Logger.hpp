// Oracle shim for <aliceVision/system/Logger.hpp>: log macros become no-ops / stderr.
#pragma once
#include <iostream>
#define ALICEVISION_LOG_TRACE(a) do {} while (0)
#define ALICEVISION_LOG_DEBUG(a) do {} while (0)
#define ALICEVISION_LOG_INFO(a) do {} while (0)
#define ALICEVISION_LOG_WARNING(a) do { std::cerr << a << std::endl; } while (0)
#define ALICEVISION_LOG_ERROR(a) do { std::cerr << a << std::endl; } while (0)
#define ALICEVISION_COUT(a) do {} while (0)
#define ALICEVISION_CERR(a) do { std::cerr << a << std::endl; } while (0)
