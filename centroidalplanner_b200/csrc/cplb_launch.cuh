// cplb_launch.cuh -- what the two kernel families share: programmatic-dependent-launch helpers, the launch wrapper and the
// parameter-source selection.  Included by cplb_kernels_cm.cuh and cplb_kernels_im.cuh only.
#ifndef CPLB_LAUNCH_CUH
#define CPLB_LAUNCH_CUH

#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>

#include "cplb_device.cuh"
#include "cplb_kernels.h"

namespace cplb {

// Programmatic dependent launch (sm_90+).  Every evaluation kernel (1) lets the NEXT kernel in the stream start
// launching right away and (2) waits for the PREVIOUS kernel to complete and flush before touching global memory,
// so stream-order semantics are unchanged for any producer/consumer of the buffers; what overlaps is the launch
// latency and CTA scheduling of back-to-back evaluations (~1 us of a ~20 us kernel).
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// With CPLB_INPUTS_READY (the caller vouches that x and the per-instance arrays were complete before the preceding
// kernel started, i.e. they are not its outputs) the loads are issued BEFORE the wait: the ~3 us HBM latency of a
// kernel's first wave of loads then overlaps the tail of the previous evaluation.  Stores always come after the wait.
__device__ __forceinline__ void pdl_prologue(unsigned flags)
{
    pdl_trigger();
    if (!(flags & CPLB_INPUTS_READY)) pdl_wait();
}
__device__ __forceinline__ void pdl_after_loads(unsigned flags)
{
    if (flags & CPLB_INPUTS_READY) pdl_wait();
}

template <class... KArgs, class... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), unsigned blocks, unsigned threads, size_t smem, cudaStream_t st, Args... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(blocks);
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

template <bool PERINST, bool COMPONENT_MAJOR>
struct ParamSource {
    using type = SharedParams;
    __device__ __forceinline__ static type make(const CplbParams& P, const CplbInstParams&, long long, long long) { return type{P}; }
};
template <bool COMPONENT_MAJOR>
struct ParamSource<true, COMPONENT_MAJOR> {
    using type = InstanceParams<COMPONENT_MAJOR>;
    __device__ __forceinline__ static type make(const CplbParams& P, const CplbInstParams& Q, long long i, long long ld)
    {
        return type{P, Q, i, ld};
    }
};

static const CplbInstParams kNoInstParams = {};

}  // namespace cplb
#endif
