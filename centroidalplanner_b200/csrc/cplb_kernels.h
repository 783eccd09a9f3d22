// cplb_kernels.h -- launchers of the evaluation kernels (cplb_kernels.cu).
#ifndef CPLB_KERNELS_H
#define CPLB_KERNELS_H

#include <cuda_runtime.h>

#include "cplb_params.h"

namespace cplb {

// per_instance: optional per-instance parameter arrays (device pointers), nullptr = the shared parameter block.
// buf[e*ld + i]: fully coalesced; cm_kernel picks the kernel: auto (by shape and batch size), one thread per
// (instance, contact) ["split"], or one thread per instance ["whole"; 4- and 8-contact shared-parameter batches only].
#define CPLB_CM_AUTO 0
#define CPLB_CM_SPLIT 1
#define CPLB_CM_WHOLE 2
cudaError_t launch_component_major(const CplbParams& P, const CplbIo& io, unsigned flags, const CplbInstParams* per_instance, int cm_kernel,
                                   cudaStream_t st);
// buf[i*len + e]: tiles staged through shared memory with bulk async (TMA) copies; im_kernel picks the kernel: auto, a warp
// per tile (lanes = (instance, contact) grid), or a CTA per tile (one warp per contact, lanes = instances).
#define CPLB_IM_AUTO 0
#define CPLB_IM_WARP_TILE 1
#define CPLB_IM_CTA_TILE 2
cudaError_t launch_instance_major(const CplbParams& P, const CplbIo& io, unsigned flags, const CplbInstParams* per_instance, int im_kernel,
                                  cudaStream_t st);

}  // namespace cplb
#endif
