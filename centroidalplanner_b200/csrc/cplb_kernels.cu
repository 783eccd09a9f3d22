// cplb_kernels.cu -- the two evaluation kernels (sm_100a, fp64, no tensor cores: the path is
// ~0.1 flop/byte, HBM-bound) and their launchers.
//
//  eval_component_major : struct-of-arrays buffers buf[e*ld + i].  One thread per instance; every
//      load and store of a warp is one contiguous 256-byte segment.
//  eval_instance_major  : per-instance contiguous slices buf[i*len + e] (what an IPOPT thread
//      consumes).  A warp owns a tile of 32/LPI consecutive instances, LPI lanes per instance
//      (one lane per contact).  The tile's x slice is fetched with ONE bulk async copy
//      (cp.async.bulk, TMA unit) into shared memory, the lanes scatter their results into shared
//      memory tiles laid out exactly like the output slices, and the tiles leave with bulk async
//      stores -- HBM only ever sees full, contiguous, 16-byte aligned bursts.
#include <cuda_runtime.h>
#include <stdint.h>

#include "cplb_device.cuh"
#include "cplb_kernels.h"

namespace cplb {

// ================================================================================================
// component-major (SoA): one thread per instance
// ================================================================================================

struct SoaEmitter {
    double* gp;
    double* jp;
    double* gradp;
    long long ld;
    long long i;
    // streaming stores: every output element is written once and never re-read by this kernel
    __device__ __forceinline__ void put(double* base, int e, double v) const { __stcs(base + (long long)e * ld + i, v); }
    __device__ __forceinline__ void g(int row, double v) const { put(gp, row, v); }
    __device__ __forceinline__ void j(int slot, double v) const { put(jp, slot, v); }
    __device__ __forceinline__ void grad(int col, double v) const { put(gradp, col, v); }
};

template <int ENV, int NC>
__global__ void __launch_bounds__(128) eval_component_major(const __grid_constant__ CplbParams P, const CplbIo io,
                                                             const unsigned flags)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= io.N) return;
    const int nc = NC > 0 ? NC : P.nc;
    const long long ld = io.ld;
    const double* __restrict__ x = io.x + i;
    SoaEmitter em{io.g, io.jac, io.grad, ld, i};

    double c[3];
#pragma unroll
    for (int q = 0; q < 3; q++) c[q] = __ldcs(x + q * ld);

    // CentroidalStatics::GetValues accumulators (CentroidalStatics.cpp:39-54) and the CoM block of
    // FillJacobianBlock (:121-135), all summed over contacts in sorted-name order.
    double v0 = 0.0, v1 = 0.0, v2 = 0.0, v3 = 0.0, v4 = 0.0, v5 = 0.0;
    double a31 = 0.0, a32 = 0.0, a40 = 0.0, a42 = 0.0, a50 = 0.0, a51 = 0.0;
    double cost = 0.0;
    const bool need_n = flags & (CPLB_WANT_G | CPLB_WANT_J);

#pragma unroll(NC > 0 ? NC : 1)
    for (int j = 0; j < nc; j++) {
        const int k = P.perm[j];
        const double* xk = x + (long long)(3 + 9 * k) * ld;
        double F[3], p[3], n[3] = {0.0, 0.0, 0.0};
#pragma unroll
        for (int q = 0; q < 3; q++) {
            F[q] = __ldcs(xk + q * ld);
            p[q] = __ldcs(xk + (3 + q) * ld);
        }
        if (need_n) {
#pragma unroll
            for (int q = 0; q < 3; q++) n[q] = __ldcs(xk + (6 + q) * ld);
        }
        const double d0 = p[0] - c[0], d1 = p[1] - c[1], d2 = p[2] - c[2];
        v0 += F[0];
        v1 += F[1];
        v2 += F[2];
        v3 += d1 * F[2] - d2 * F[1];
        v4 += d2 * F[0] - d0 * F[2];
        v5 += d0 * F[1] - d1 * F[0];
        a31 -= F[2];
        a32 -= -F[1];
        a40 -= -F[2];
        a42 -= F[0];
        a50 -= F[1];
        a51 -= -F[0];
        contact_rows<ENV>(P, em, nc, j, k, c, F, p, n, flags);
        if (flags & CPLB_WANT_COST) cost += contact_cost(P, k, F, p);
    }

    if (flags & CPLB_WANT_G) {  // :56-57  value -= wrench; value.head<3>() += m*g
        em.g(0, (v0 - P.wrench[0]) + P.mg[0]);
        em.g(1, (v1 - P.wrench[1]) + P.mg[1]);
        em.g(2, (v2 - P.wrench[2]) + P.mg[2]);
        em.g(3, v3 - P.wrench[3]);
        em.g(4, v4 - P.wrench[4]);
        em.g(5, v5 - P.wrench[5]);
    }
    if (flags & CPLB_WANT_J) {
        const int L = jac_moment_row_len(nc);
        const int s3 = 3 * nc;
        em.j(s3 + 0, a31);
        em.j(s3 + 1, a32);
        em.j(s3 + L + 0, a40);
        em.j(s3 + L + 1, a42);
        em.j(s3 + 2 * L + 0, a50);
        em.j(s3 + 2 * L + 1, a51);
    }
    if (flags & CPLB_WANT_COST) {
        cost += com_cost(P, c);
        __stcs(io.cost + i, cost);
    }
    if (flags & CPLB_WANT_GRAD) {
#pragma unroll
        for (int q = 0; q < 3; q++) em.grad(q, P.W_com * (c[q] - P.com_ref[q]));
    }
}

// ================================================================================================
// instance-major (AoS): warp tile, LPI lanes per instance, bulk async copies in and out
// ================================================================================================

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// global -> shared bulk copy (TMA unit), completion signalled on the mbarrier
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// shared -> global bulk copy
__device__ __forceinline__ void bulk_s2g(void* gdst, const void* smem_src, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

struct TileEmitter {
    double* gp;  // this instance's slices inside the warp's shared-memory tiles
    double* jp;
    double* gradp;
    __device__ __forceinline__ void g(int row, double v) const { gp[row] = v; }
    __device__ __forceinline__ void j(int slot, double v) const { jp[slot] = v; }
    __device__ __forceinline__ void grad(int col, double v) const { gradp[col] = v; }
};

// shared memory per warp (doubles): [x: T*n][g: T*m][jac: T*nnz][grad: T*n][cost: T] + one mbarrier
__host__ __device__ inline size_t tile_doubles(int T, int n, int m, int nnz, unsigned flags)
{
    size_t d = (size_t)T * n;
    if (flags & CPLB_WANT_G) d += (size_t)T * m;
    if (flags & CPLB_WANT_J) d += (size_t)T * nnz;
    if (flags & CPLB_WANT_GRAD) d += (size_t)T * n;
    if (flags & CPLB_WANT_COST) d += (size_t)T;
    return (d + 1) & ~(size_t)1;  // keep every warp's region 16-byte aligned
}

// warp-cooperative contiguous copy, used for ragged / misaligned tiles instead of the bulk engine
__device__ __forceinline__ void warp_copy(double* dst, const double* src, int count, int lane)
{
    for (int e = lane; e < count; e += 32) dst[e] = src[e];
}

template <int ENV, int LPI, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) eval_instance_major(const __grid_constant__ CplbParams P, const CplbIo io,
                                                                   const unsigned flags, const int aligned16)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int T = 32 / LPI;  // instances per warp tile
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nc = P.nc, n = P.n, m = P.m, nnz = P.nnz;
    const long long tile = (long long)blockIdx.x * WARPS + warp;
    const long long i0 = tile * T;
    if (i0 >= io.N) return;  // whole warp leaves together; no block-wide barrier is used below
    const int cnt = (io.N - i0) < T ? (int)(io.N - i0) : T;

    const size_t per_warp = tile_doubles(T, n, m, nnz, flags);
    double* xs = reinterpret_cast<double*>(smem_raw) + (size_t)warp * per_warp;
    double* cur = xs + (size_t)T * n;
    double* gs = nullptr;
    double* js = nullptr;
    double* grads = nullptr;
    double* costs = nullptr;
    if (flags & CPLB_WANT_G) { gs = cur; cur += (size_t)T * m; }
    if (flags & CPLB_WANT_J) { js = cur; cur += (size_t)T * nnz; }
    if (flags & CPLB_WANT_GRAD) { grads = cur; cur += (size_t)T * n; }
    if (flags & CPLB_WANT_COST) { costs = cur; }
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<double*>(smem_raw) + (size_t)WARPS * per_warp);
    uint64_t* bar = bars + warp;

    // bulk copies need 16-byte aligned addresses and sizes: full tiles of 16B-aligned buffers only
    const bool bulk = aligned16 && cnt == T;

    // ---- fetch the tile's x slice -----------------------------------------------------------
    const double* xsrc = io.x + i0 * n;
    if (bulk) {
        if (lane == 0) {
            mbar_init(bar, 1);
            fence_proxy_async_smem();
            mbar_expect_tx(bar, (uint32_t)(T * n * sizeof(double)));
            bulk_g2s(xs, xsrc, (uint32_t)(T * n * sizeof(double)), bar);
        }
        __syncwarp();
        mbar_wait(bar, 0);
    } else {
        warp_copy(xs, xsrc, cnt * n, lane);
        __syncwarp();
    }

    // ---- compute: lane (inst, s) handles contacts s, s+LPI, ... and statics rows s, s+LPI, ... ----
    const int inst = lane / LPI, s = lane % LPI;
    if (inst < cnt) {
        const double* xi = xs + (size_t)inst * n;
        TileEmitter em{gs ? gs + (size_t)inst * m : nullptr, js ? js + (size_t)inst * nnz : nullptr,
                       grads ? grads + (size_t)inst * n : nullptr};
        const double c[3] = {xi[0], xi[1], xi[2]};
        for (int j = s; j < nc; j += LPI) {
            const int k = P.perm[j];
            const double* xk = xi + 3 + 9 * k;
            const double F[3] = {xk[0], xk[1], xk[2]};
            const double p[3] = {xk[3], xk[4], xk[5]};
            const double nn[3] = {xk[6], xk[7], xk[8]};
            contact_rows<ENV>(P, em, nc, j, k, c, F, p, nn, flags);
        }
        // CentroidalStatics rows: row r's running sum visits the contacts in sorted-name order
        // (CentroidalStatics.cpp:44-54); the six rows are independent, so they are dealt to the lanes.
        if (flags & (CPLB_WANT_G | CPLB_WANT_J)) {
            const int L = jac_moment_row_len(nc);
            for (int r = s; r < 6; r += LPI) {
                double v = 0.0, a = 0.0, b = 0.0;
                for (int j = 0; j < nc; j++) {
                    const double* xk = xi + 3 + 9 * P.perm[j];
                    const double F0 = xk[0], F1 = xk[1], F2 = xk[2];
                    const double d0 = xk[3] - c[0], d1 = xk[4] - c[1], d2 = xk[5] - c[2];
                    switch (r) {
                    case 0: v += F0; break;
                    case 1: v += F1; break;
                    case 2: v += F2; break;
                    case 3: v += d1 * F2 - d2 * F1; a -= F2; b -= -F1; break;   // :128-129
                    case 4: v += d2 * F0 - d0 * F2; a -= -F2; b -= F0; break;   // :130-131
                    default: v += d0 * F1 - d1 * F0; a -= F1; b -= -F0; break;  // :132-133
                    }
                }
                if (flags & CPLB_WANT_G) em.g(r, r < 3 ? (v - P.wrench[r]) + P.mg[r] : v - P.wrench[r]);
                if ((flags & CPLB_WANT_J) && r >= 3) {
                    em.j(3 * nc + (r - 3) * L + 0, a);
                    em.j(3 * nc + (r - 3) * L + 1, b);
                }
            }
        }
        if (s == 0) {
            if (flags & CPLB_WANT_COST) {  // MinimizeCentroidalVariables.cpp:126-147, sorted order
                double cost = 0.0;
                for (int j = 0; j < nc; j++) {
                    const int k = P.perm[j];
                    const double* xk = xi + 3 + 9 * k;
                    const double F[3] = {xk[0], xk[1], xk[2]};
                    const double p[3] = {xk[3], xk[4], xk[5]};
                    cost += contact_cost(P, k, F, p);
                }
                cost += com_cost(P, c);
                costs[inst] = cost;
            }
            if (flags & CPLB_WANT_GRAD) {
#pragma unroll
                for (int q = 0; q < 3; q++) em.grad(q, P.W_com * (c[q] - P.com_ref[q]));
            }
        }
    }

    // ---- ship the tiles ------------------------------------------------------------------------
    if (bulk) {
        fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the async (TMA) proxy
        __syncwarp();
        if (lane == 0) {
            if (gs) bulk_s2g(io.g + i0 * m, gs, (uint32_t)(T * m * sizeof(double)));
            if (js) bulk_s2g(io.jac + i0 * nnz, js, (uint32_t)(T * nnz * sizeof(double)));
            if (grads) bulk_s2g(io.grad + i0 * n, grads, (uint32_t)(T * n * sizeof(double)));
            bulk_commit();
        }
        if (costs && lane < T) io.cost[i0 + lane] = costs[lane];
        if (lane == 0) bulk_wait_read_all();  // shared memory must outlive the engine's reads
        __syncwarp();
    } else {
        __syncwarp();
        if (gs) warp_copy(io.g + i0 * m, gs, cnt * m, lane);
        if (js) warp_copy(io.jac + i0 * nnz, js, cnt * nnz, lane);
        if (grads) warp_copy(io.grad + i0 * n, grads, cnt * n, lane);
        if (costs && lane < cnt) io.cost[i0 + lane] = costs[lane];
    }
}

// ================================================================================================
// launchers
// ================================================================================================

template <int ENV>
static cudaError_t launch_cm_env(const CplbParams& P, const CplbIo& io, unsigned flags, cudaStream_t st)
{
    const int threads = 128;
    const unsigned blocks = (unsigned)((io.N + threads - 1) / threads);
    switch (P.nc) {
    case 4: eval_component_major<ENV, 4><<<blocks, threads, 0, st>>>(P, io, flags); break;
    case 8: eval_component_major<ENV, 8><<<blocks, threads, 0, st>>>(P, io, flags); break;
    default: eval_component_major<ENV, 0><<<blocks, threads, 0, st>>>(P, io, flags); break;
    }
    return cudaGetLastError();
}

cudaError_t launch_component_major(const CplbParams& P, const CplbIo& io, unsigned flags, cudaStream_t st)
{
    if (io.N <= 0) return cudaSuccess;
    switch (P.env) {
    case CPLB_ENV_NONE_K: return launch_cm_env<CPLB_ENV_NONE_K>(P, io, flags, st);
    case CPLB_ENV_GROUND_K: return launch_cm_env<CPLB_ENV_GROUND_K>(P, io, flags, st);
    default: return launch_cm_env<CPLB_ENV_SUPERQUADRIC_K>(P, io, flags, st);
    }
}

template <int ENV, int LPI>
static cudaError_t launch_im_cfg(const CplbParams& P, const CplbIo& io, unsigned flags, cudaStream_t st)
{
    constexpr int WARPS = 4;
    constexpr int T = 32 / LPI;
    const size_t smem = tile_doubles(T, P.n, P.m, P.nnz, flags) * sizeof(double) * WARPS + WARPS * sizeof(uint64_t);
    auto kern = eval_instance_major<ENV, LPI, WARPS>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const long long tiles = (io.N + T - 1) / T;
    const unsigned blocks = (unsigned)((tiles + WARPS - 1) / WARPS);
    auto al16 = [](const void* p) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    const int aligned16 = al16(io.x) && al16(io.g) && al16(io.jac) && al16(io.grad) && (T % 2 == 0);
    kern<<<blocks, WARPS * 32, smem, st>>>(P, io, flags, aligned16);
    return cudaGetLastError();
}

template <int ENV>
static cudaError_t launch_im_env(const CplbParams& P, const CplbIo& io, unsigned flags, cudaStream_t st)
{
    // lanes per instance: the smallest power of two >= nc (capped at 8; more contacts loop)
    if (P.nc <= 1) return launch_im_cfg<ENV, 1>(P, io, flags, st);
    if (P.nc <= 2) return launch_im_cfg<ENV, 2>(P, io, flags, st);
    if (P.nc <= 4) return launch_im_cfg<ENV, 4>(P, io, flags, st);
    return launch_im_cfg<ENV, 8>(P, io, flags, st);
}

cudaError_t launch_instance_major(const CplbParams& P, const CplbIo& io, unsigned flags, cudaStream_t st)
{
    if (io.N <= 0) return cudaSuccess;
    switch (P.env) {
    case CPLB_ENV_NONE_K: return launch_im_env<CPLB_ENV_NONE_K>(P, io, flags, st);
    case CPLB_ENV_GROUND_K: return launch_im_env<CPLB_ENV_GROUND_K>(P, io, flags, st);
    default: return launch_im_env<CPLB_ENV_SUPERQUADRIC_K>(P, io, flags, st);
    }
}

}  // namespace cplb
