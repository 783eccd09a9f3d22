// cplb_kernels.cu -- the two evaluation kernels (sm_100a, fp64, no tensor cores: the path is
// ~0.1 flop/byte, HBM-bound) and their launchers.
//
//  eval_component_major_split : struct-of-arrays buffers buf[e*ld + i].  One thread per (instance,
//      contact): a warp is one contact of 32 consecutive instances, so every load and store of a
//      warp is one contiguous 256-byte segment; the six statics sums cross warps through shared memory.
//  eval_instance_major  : per-instance contiguous slices buf[i*len + e] (what an IPOPT thread
//      consumes).  A warp owns a tile of 32/LPI consecutive instances, LPI lanes per instance
//      (one lane per contact).  The tile's x slice is fetched with ONE bulk async copy
//      (cp.async.bulk, TMA unit) into shared memory, the lanes scatter their results into shared
//      memory tiles laid out exactly like the output slices, and the tiles leave with bulk async
//      stores -- HBM only ever sees full, contiguous, 16-byte aligned bursts.
//
// The kernels live in cplb_kernels_cm.cuh / cplb_kernels_im.cuh and are instantiated per environment kind in
// cplb_kernels_{cm,im}_{none,ground,superquadric}.cu; this file only dispatches on the kind.
#include "cplb_kernels.h"

#include "cplb_device.cuh"

namespace cplb {

template <int ENV>
cudaError_t launch_cm_env(const CplbParams& P, const CplbIo& io, unsigned flags, const CplbInstParams* Q, int cm_kernel, cudaStream_t st);
template <int ENV>
cudaError_t launch_im_env(const CplbParams& P, const CplbIo& io, unsigned flags, const CplbInstParams* Q, int im_kernel, cudaStream_t st);
extern template cudaError_t launch_cm_env<CPLB_ENV_NONE_K>(const CplbParams&, const CplbIo&, unsigned, const CplbInstParams*, int, cudaStream_t);
extern template cudaError_t launch_cm_env<CPLB_ENV_GROUND_K>(const CplbParams&, const CplbIo&, unsigned, const CplbInstParams*, int, cudaStream_t);
extern template cudaError_t launch_cm_env<CPLB_ENV_SUPERQUADRIC_K>(const CplbParams&, const CplbIo&, unsigned, const CplbInstParams*, int, cudaStream_t);
extern template cudaError_t launch_im_env<CPLB_ENV_NONE_K>(const CplbParams&, const CplbIo&, unsigned, const CplbInstParams*, int, cudaStream_t);
extern template cudaError_t launch_im_env<CPLB_ENV_GROUND_K>(const CplbParams&, const CplbIo&, unsigned, const CplbInstParams*, int, cudaStream_t);
extern template cudaError_t launch_im_env<CPLB_ENV_SUPERQUADRIC_K>(const CplbParams&, const CplbIo&, unsigned, const CplbInstParams*, int, cudaStream_t);

cudaError_t launch_component_major(const CplbParams& P, const CplbIo& io, unsigned flags, const CplbInstParams* Q, int cm_kernel, cudaStream_t st)
{
    if (io.N <= 0) return cudaSuccess;
    if (io.ld >= (1LL << 29)) return cudaErrorInvalidValue;  // row pitch must fit 32 bits of bytes
    switch (P.env) {
    case CPLB_ENV_NONE_K: return launch_cm_env<CPLB_ENV_NONE_K>(P, io, flags, Q, cm_kernel, st);
    case CPLB_ENV_GROUND_K: return launch_cm_env<CPLB_ENV_GROUND_K>(P, io, flags, Q, cm_kernel, st);
    default: return launch_cm_env<CPLB_ENV_SUPERQUADRIC_K>(P, io, flags, Q, cm_kernel, st);
    }
}

cudaError_t launch_instance_major(const CplbParams& P, const CplbIo& io, unsigned flags, const CplbInstParams* Q, int im_kernel, cudaStream_t st)
{
    if (io.N <= 0) return cudaSuccess;
    switch (P.env) {
    case CPLB_ENV_NONE_K: return launch_im_env<CPLB_ENV_NONE_K>(P, io, flags, Q, im_kernel, st);
    case CPLB_ENV_GROUND_K: return launch_im_env<CPLB_ENV_GROUND_K>(P, io, flags, Q, im_kernel, st);
    default: return launch_im_env<CPLB_ENV_SUPERQUADRIC_K>(P, io, flags, Q, im_kernel, st);
    }
}

}  // namespace cplb
