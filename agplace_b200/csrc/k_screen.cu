// k_screen.cu -- one instantiation of the single-pass certified screen + its finish (compile with -DAGP_E=<2|4|8|16>).
#include "knn_screen.cuh"
namespace agp {
template cudaError_t launch_knn_screen<AGP_E>(const CUtensorMap&, const CUtensorMap&, const ScreenParams&, int, size_t, cudaStream_t);
template cudaError_t launch_screen_finalize<AGP_E>(const uint64_t*, const int*, int, int64_t, int, int, int, const float*, const float*, int, int,
                                                   const float*, const float*, const uint32_t*, const int*, int*, int*, int64_t, float*,
                                                   int64_t*, int, cudaStream_t);
}
