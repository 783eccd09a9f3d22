// instance-major kernels of the superquadric environment kind (explicit instantiation, see cplb_kernels_im.cuh)
#include "cplb_kernels_im.cuh"

namespace cplb {
template cudaError_t launch_im_env<CPLB_ENV_SUPERQUADRIC_K>(const CplbParams&, const CplbIo&, unsigned, const CplbInstParams*, int, cudaStream_t);
}  // namespace cplb
