// component-major kernels of the superquadric environment kind (explicit instantiation, see cplb_kernels_cm.cuh)
#include "cplb_kernels_cm.cuh"

namespace cplb {
template cudaError_t launch_cm_env<CPLB_ENV_SUPERQUADRIC_K>(const CplbParams&, const CplbIo&, unsigned, const CplbInstParams*, int, cudaStream_t);
}  // namespace cplb
