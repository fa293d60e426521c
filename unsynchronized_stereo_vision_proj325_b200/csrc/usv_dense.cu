// usv_dense.cu — dense stride-1 sweep kernels (sliding-window formulation).
#include "usv_common.cuh"

namespace usv {

cudaError_t launch_dense(const DevJob& J, int n_pairs, cudaStream_t st, const char** kernel_name, int* n_launches) {
  (void)J; (void)n_pairs; (void)st; (void)kernel_name; (void)n_launches;
  return cudaErrorNotSupported;
}

}  // namespace usv
