// Oracle shim for the generated <aliceVision/config.hpp> (mirrors src/cmake/config.hpp.in).
// SSE and OpenMP ON, as in the reference's release builds (src/CMakeLists.txt:194-208).
#pragma once
#define ALICEVISION_IS_DEFINED(F) F() == 1
#define ALICEVISION_HAVE_OPENMP() 1
#define ALICEVISION_HAVE_SSE() 1
#define ALICEVISION_HAVE_OPENCV() 0
#define ALICEVISION_HAVE_OCVSIFT() 0
#define ALICEVISION_HAVE_CCTAG() 0
#define ALICEVISION_HAVE_APRILTAG() 0
#define ALICEVISION_HAVE_POPSIFT() 0
#define ALICEVISION_HAVE_CUDA() 0
#define ALICEVISION_HAVE_OPENGV() 0
#define ALICEVISION_HAVE_ALEMBIC() 0
#define ALICEVISION_HAVE_ONNX() 0
