// k_tc.cu -- one instantiation of the fused tcgen05 distance + top-k kernel (compile with -DAGP_E=<2|4|8|16>).
#include "knn_tc.cuh"
namespace agp {
template cudaError_t launch_knn_tc<AGP_E>(const CUtensorMap&, const CUtensorMap&, const CUtensorMap&, const CUtensorMap&,
                                          const TcParams&, int, cudaStream_t);
}
