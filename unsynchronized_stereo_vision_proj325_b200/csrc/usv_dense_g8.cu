// usv_dense_g8.cu — the one-plane, 8-disparities-per-thread variant of the sliding-window SAD kernel (usv_dense.cu).
#include "usv_dense_kernel.cuh"

namespace usv {

USV_DENSE_DEFINE_VARIANT(1, 8)

}  // namespace usv
