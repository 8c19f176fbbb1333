// k_select.cu -- row-select and k-way merge kernels for one register budget (compile with -DAGP_E=<2..32>).
#include "merge.cuh"
#include "select.cuh"
namespace agp {
template cudaError_t launch_select_rows<AGP_E>(const float*, int64_t, int64_t, int, int, int, uint64_t*, int, cudaStream_t);
template cudaError_t launch_merge_keys<AGP_E>(const uint64_t*, int64_t, int, int, int64_t, float*, int64_t*, int, cudaStream_t);
template cudaError_t launch_merge_lists<AGP_E>(const float*, int64_t, const int64_t*, int64_t, bool, int64_t, int, int, float*, int64_t*,
                                               int, cudaStream_t);
template cudaError_t launch_merge_ragged<AGP_E>(const uint64_t*, const int*, int, int64_t, int, int, int64_t, float*, int64_t*, cudaStream_t);
template cudaError_t launch_subset_topk<AGP_E>(const float*, const float*, int, const int64_t*, const int64_t*, int64_t, int64_t, int, uint64_t*,
                                               float*, int64_t*, cudaStream_t);
template cudaError_t launch_rerank<AGP_E>(const float*, const float*, int, const int64_t*, int, int64_t, int, int64_t, float*, int64_t*,
                                          cudaStream_t);
}
