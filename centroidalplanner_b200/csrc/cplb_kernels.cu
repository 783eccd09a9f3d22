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
static cudaError_t launch_pdl(void (*kernel)(KArgs...), unsigned blocks, unsigned threads, size_t smem, cudaStream_t st, Args... args)
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

// ================================================================================================
// component-major (SoA)
// ================================================================================================

// Addressing: element (e, i) lives at base + e*ld + i.  With the row pitch in BYTES held in 32 bits
// (ld < 2^29) every address is one IMAD.WIDE.U32 (e * pitch + pointer): the kernel issues ~250
// loads/stores per instance, so the address arithmetic is as hot as the fp64 arithmetic.
struct SoaEmitter {
    char* gp;  // already offset to this thread's instance column
    char* jp;
    char* gradp;
    unsigned pitch;  // ld * sizeof(double)
    // streaming stores: every output element is written once and never re-read by this kernel
    __device__ __forceinline__ void put(char* base, int e, double v) const
    {
        __stcs(reinterpret_cast<double*>(base + (unsigned long long)(unsigned)e * pitch), v);
    }
    __device__ __forceinline__ void g(int row, double v) const { put(gp, row, v); }
    __device__ __forceinline__ void j(int slot, double v) const { put(jp, slot, v); }
    __device__ __forceinline__ void grad(int col, double v) const { put(gradp, col, v); }
};

__device__ __forceinline__ double ld_stream(const char* base, int e, unsigned pitch)
{
    return __ldcs(reinterpret_cast<const double*>(base + (unsigned long long)(unsigned)e * pitch));
}

// One thread per (instance, contact): warp w of a CTA owns the contact of sorted rank w for 32
// consecutive instances, so a CTA is nc warps and every global access of a warp is one contiguous
// 256-byte segment.  Compared with one thread per instance this puts nc times more warps in
// flight and cuts each thread's dependent instruction stream by nc -- what matters at 65,536
// instances, where the whole batch is less than one wave of threads and latency, not bandwidth,
// is the limit.  The only cross-contact quantities are the six CentroidalStatics sums; each warp
// leaves its contact's force and moment term in shared memory and, after one barrier, warp r adds
// row r's terms in sorted-name order (CentroidalStatics.cpp:44-54), the order that fixes rounding.
template <int ENV, unsigned FLAGS, int MAX_WARPS, bool PERINST>
__global__ void __launch_bounds__(MAX_WARPS * 32, (MAX_WARPS == 8 ? 4 : 1))  // <= 64 registers: 32 warps per SM
    eval_component_major_split(const __grid_constant__ CplbParams P, const CplbIo io, const unsigned flags_rt,
                               const __grid_constant__ CplbInstParams Q)
{
    extern __shared__ double sh_all[];  // [sub-block][nc][6 + 1][32]
    pdl_prologue(flags_rt);
    const unsigned flags = FLAGS ? FLAGS : (flags_rt & 15u);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nc = P.nc;
    const int sub = warp / nc, j = warp - sub * nc;  // a CTA holds blockDim/(32 nc) sub-blocks of 32 instances
    const int subs = blockDim.x / (32 * nc);
    double* sh = sh_all + (size_t)sub * nc * (192 + 32);
    const long long i_raw = ((long long)blockIdx.x * subs + sub) * 32 + lane;
    const bool active = i_raw < io.N;
    const long long i = active ? i_raw : io.N - 1;  // inactive lanes recompute the last instance, store nothing
    const unsigned pitch = (unsigned)(io.ld * (long long)sizeof(double));
    const char* x = reinterpret_cast<const char*>(io.x + i);
    const int k = P.perm[j];
    const bool need_n = flags & (CPLB_WANT_G | CPLB_WANT_J);
    const auto ps = ParamSource<PERINST, true>::make(P, Q, i, io.ld);

    double c[3], F[3], p[3], n[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int q = 0; q < 3; q++) c[q] = ld_stream(x, q, pitch);
#pragma unroll
    for (int q = 0; q < 3; q++) F[q] = ld_stream(x, 3 + 9 * k + q, pitch);
#pragma unroll
    for (int q = 0; q < 3; q++) p[q] = ld_stream(x, 6 + 9 * k + q, pitch);
    if (need_n) {
#pragma unroll
        for (int q = 0; q < 3; q++) n[q] = ld_stream(x, 9 + 9 * k + q, pitch);
    }
    pdl_after_loads(flags_rt);  // the loads above are in flight; everything below may store

    double* mine = sh + (size_t)j * 192 + lane;
    if (flags & (CPLB_WANT_G | CPLB_WANT_J)) {
        const double d0 = p[0] - c[0], d1 = p[1] - c[1], d2 = p[2] - c[2];
        mine[0 * 32] = F[0];
        mine[1 * 32] = F[1];
        mine[2 * 32] = F[2];
        mine[3 * 32] = d1 * F[2] - d2 * F[1];  // (p - CoM).cross(F), CentroidalStatics.cpp:53
        mine[4 * 32] = d2 * F[0] - d0 * F[2];
        mine[5 * 32] = d0 * F[1] - d1 * F[0];
    }
    if (flags & CPLB_WANT_COST) {
        if (!(flags & (CPLB_WANT_G | CPLB_WANT_J))) mine[0] = contact_cost(ps, P.reduction_order, k, F, p);
        else sh[(size_t)nc * 192 + (size_t)j * 32 + lane] = contact_cost(ps, P.reduction_order, k, F, p);
    }

    SoaEmitter em{reinterpret_cast<char*>(io.g + i), reinterpret_cast<char*>(io.jac + i),
                  reinterpret_cast<char*>(io.grad + i), pitch};
    if (active) contact_rows<ENV>(P, ps, em, nc, j, k, c, F, p, n, flags);

    __syncthreads();
    if (!active) return;

    if (flags & (CPLB_WANT_G | CPLB_WANT_J)) {
        const int L = jac_moment_row_len(nc);
        for (int r = j; r < 6; r += nc) {
            const double* col = sh + r * 32 + lane;
            double v = 0.0;
            for (int jj = 0; jj < nc; jj++) v += col[(size_t)jj * 192];
            if (flags & CPLB_WANT_G) em.g(r, r < 3 ? (v - ps.wrench(r)) + ps.mg(r) : v - ps.wrench(r));  // :56-57
            if ((flags & CPLB_WANT_J) && r >= 3) {
                // CoM block (:128-133): row 3 <- (Fz, -Fy), row 4 <- (-Fz, Fx), row 5 <- (Fy, -Fx), each "acc -= term"
                const int ia = r == 3 ? 2 : (r == 4 ? 2 : 1), ib = r == 3 ? 1 : (r == 4 ? 0 : 0);
                const bool nega = (r == 4), negb = (r != 4);
                double a = 0.0, b = 0.0;
                for (int jj = 0; jj < nc; jj++) {
                    const double fa = sh[(size_t)jj * 192 + ia * 32 + lane], fb = sh[(size_t)jj * 192 + ib * 32 + lane];
                    a -= nega ? -fa : fa;
                    b -= negb ? -fb : fb;
                }
                em.j(3 * nc + (r - 3) * L + 0, a);
                em.j(3 * nc + (r - 3) * L + 1, b);
            }
        }
    }
    if (j == 0) {
        if (flags & CPLB_WANT_COST) {  // MinimizeCentroidalVariables.cpp:126-147: contacts in sorted order, then the CoM term
            const double* cc = (flags & (CPLB_WANT_G | CPLB_WANT_J)) ? sh + (size_t)nc * 192 + lane : sh + lane;
            const size_t stride = (flags & (CPLB_WANT_G | CPLB_WANT_J)) ? 32 : 192;
            double cost = 0.0;
            for (int jj = 0; jj < nc; jj++) cost += cc[(size_t)jj * stride];
            cost += com_cost(ps, P.reduction_order, c);
            __stcs(io.cost + i, cost);
        }
        if (flags & CPLB_WANT_GRAD) {
#pragma unroll
            for (int q = 0; q < 3; q++) em.grad(q, ps.W_com() * (c[q] - ps.com_ref(q)));
        }
    }
}

// One thread per instance, for LARGER batches (where several waves of CTAs keep HBM busy and the per-contact split's
// extra threads, shared-memory exchange and barrier only cost): all 3 + 9 nc loads of a thread are issued up front, the
// statics sums stay in registers (164 / 248 registers for nc = 4 / 8; capping them spills and loses 8 % at 1M).
// Measured at 1,048,576 instances: 97.7-98.3 % of the HBM roofline against 95 % for the split kernel; at 65,536
// four-contact instances it is the other way round (80 % vs 88 %).
template <int ENV, int NC, unsigned FLAGS>
__global__ void __launch_bounds__(128) eval_component_major_whole(const __grid_constant__ CplbParams P, const CplbIo io,
                                                                   const unsigned flags_rt)
{
    pdl_prologue(flags_rt);
    const unsigned flags = FLAGS ? FLAGS : (flags_rt & 15u);
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= io.N) {
        pdl_after_loads(flags_rt);
        return;
    }
    const unsigned pitch = (unsigned)(io.ld * (long long)sizeof(double));
    const char* x = reinterpret_cast<const char*>(io.x + i);
    const SharedParams ps{P};
    SoaEmitter em{reinterpret_cast<char*>(io.g + i), reinterpret_cast<char*>(io.jac + i), reinterpret_cast<char*>(io.grad + i), pitch};

    double c[3], F[NC][3], p[NC][3], n[NC][3];
#pragma unroll
    for (int q = 0; q < 3; q++) c[q] = ld_stream(x, q, pitch);
#pragma unroll
    for (int j = 0; j < NC; j++) {
        const int k = P.perm[j];
#pragma unroll
        for (int q = 0; q < 3; q++) {
            F[j][q] = ld_stream(x, 3 + 9 * k + q, pitch);
            p[j][q] = ld_stream(x, 6 + 9 * k + q, pitch);
            n[j][q] = ld_stream(x, 9 + 9 * k + q, pitch);
        }
    }
    pdl_after_loads(flags_rt);
    double v[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    double a31 = 0.0, a32 = 0.0, a40 = 0.0, a42 = 0.0, a50 = 0.0, a51 = 0.0, cost = 0.0;
#pragma unroll
    for (int j = 0; j < NC; j++) {  // sorted-name order (CentroidalStatics.cpp:44-54, :121-135)
        const int k = P.perm[j];
        const double d0 = p[j][0] - c[0], d1 = p[j][1] - c[1], d2 = p[j][2] - c[2];
        v[0] += F[j][0];
        v[1] += F[j][1];
        v[2] += F[j][2];
        v[3] += d1 * F[j][2] - d2 * F[j][1];
        v[4] += d2 * F[j][0] - d0 * F[j][2];
        v[5] += d0 * F[j][1] - d1 * F[j][0];
        a31 -= F[j][2];
        a32 -= -F[j][1];
        a40 -= -F[j][2];
        a42 -= F[j][0];
        a50 -= F[j][1];
        a51 -= -F[j][0];
        contact_rows<ENV>(P, ps, em, NC, j, k, c, F[j], p[j], n[j], flags);
        if (flags & CPLB_WANT_COST) cost += contact_cost(ps, P.reduction_order, k, F[j], p[j]);
    }
    if (flags & CPLB_WANT_G) {
#pragma unroll
        for (int r = 0; r < 6; r++) em.g(r, r < 3 ? (v[r] - P.wrench[r]) + P.mg[r] : v[r] - P.wrench[r]);  // :56-57
    }
    if (flags & CPLB_WANT_J) {
        const int L = jac_moment_row_len(NC), s3 = 3 * NC;
        em.j(s3 + 0, a31);
        em.j(s3 + 1, a32);
        em.j(s3 + L + 0, a40);
        em.j(s3 + L + 1, a42);
        em.j(s3 + 2 * L + 0, a50);
        em.j(s3 + 2 * L + 1, a51);
    }
    if (flags & CPLB_WANT_COST) {
        cost += com_cost(ps, P.reduction_order, c);
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

// shared memory per warp (doubles): [x0: T*n][x1: T*n][g: T*m][jac: T*nnz][grad: T*n][cost: T]; then 2 mbarriers per warp
__host__ __device__ inline size_t tile_doubles(int T, int n, int m, int nnz, unsigned flags)
{
    size_t d = 2 * (size_t)T * n;
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

// Persistent warps.  Warp w of the grid owns tiles w, w + W, w + 2W, ... (W = warps in the grid, the grid is
// sized to what is resident at once).  Per tile of T = 32/LPI consecutive instances:
//   - the x slice (T*n contiguous doubles) arrives by one bulk async copy into one of two buffers; the copy of
//     the NEXT tile is issued before the current one is consumed, so its HBM latency hides behind compute;
//   - LPI lanes per instance (one lane per contact; the six statics rows are dealt to the same lanes) scatter
//     results into shared-memory tiles laid out exactly like the output slices;
//   - the tiles leave with bulk async stores; the warp only waits for the engine to have READ the tiles
//     right before it overwrites them with the next tile's results.
// No block-wide barrier: every warp runs its own pipeline (mbarriers + __syncwarp only).
template <int ENV, int LPI, int WARPS, unsigned FLAGS, bool PERINST>
__global__ void __launch_bounds__(WARPS * 32) eval_instance_major(const __grid_constant__ CplbParams P, const CplbIo io,
                                                                   const unsigned flags_rt, const int aligned16,
                                                                   const __grid_constant__ CplbInstParams Q,
                                                                   const __grid_constant__ CplbParamTile PT)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    pdl_trigger();
    if (!(flags_rt & CPLB_INPUTS_READY)) pdl_wait();
    constexpr int T = 32 / LPI;  // instances per warp tile
    const unsigned flags = FLAGS ? FLAGS : (flags_rt & 15u);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nc = P.nc, n = P.n, m = P.m, nnz = P.nnz;
    const long long tiles = (io.N + T - 1) / T;
    const long long stride = (long long)gridDim.x * WARPS;
    long long tile = (long long)blockIdx.x * WARPS + warp;
    if (tile >= tiles) return;  // whole warp leaves together

    const int ptot = PERINST ? PT.total : 0;  // staged per-instance parameter slices, double-buffered like x
    const size_t per_warp = tile_doubles(T, n, m, nnz, flags) + 2 * (size_t)ptot;
    double* xbuf = reinterpret_cast<double*>(smem_raw) + (size_t)warp * per_warp;
    double* pbuf = xbuf + 2 * (size_t)T * n;
    double* cur = pbuf + 2 * (size_t)ptot;
    double* gs = nullptr;
    double* js = nullptr;
    double* grads = nullptr;
    double* costs = nullptr;
    if (flags & CPLB_WANT_G) { gs = cur; cur += (size_t)T * m; }
    if (flags & CPLB_WANT_J) { js = cur; cur += (size_t)T * nnz; }
    if (flags & CPLB_WANT_GRAD) { grads = cur; cur += (size_t)T * n; }
    if (flags & CPLB_WANT_COST) { costs = cur; }
    uint64_t* bar = reinterpret_cast<uint64_t*>(reinterpret_cast<double*>(smem_raw) + (size_t)WARPS * per_warp) + 2 * warp;

    const uint32_t xbytes = (uint32_t)(T * n * sizeof(double));
    // bulk copies need 16-byte aligned addresses and sizes: full tiles of 16B-aligned buffers only
    auto is_bulk = [&](long long t) { return aligned16 && (t + 1) * T <= io.N; };
    const double* const parr[CPLB_NUM_INST_ARRAYS] = {Q.mass, Q.wrench, Q.mu, Q.F_thr, Q.ground_z, Q.com_ref, Q.W_com,
                                                       Q.p_ref, Q.F_ref, Q.W_p, Q.W_F};
    // one elected lane: x tile + every staged parameter slice of tile t into buffer `buf`, all on one mbarrier
    auto issue_loads = [&](long long t, int buf) {
        mbar_expect_tx(&bar[buf], xbytes + (uint32_t)(ptot * sizeof(double)));
        bulk_g2s(xbuf + (size_t)buf * T * n, io.x + t * T * n, xbytes, &bar[buf]);
        if (PERINST) {
#pragma unroll
            for (int a = 0; a < CPLB_NUM_INST_ARRAYS; a++)
                if (PT.off[a] >= 0)
                    bulk_g2s(pbuf + (size_t)buf * ptot + PT.off[a], parr[a] + t * T * PT.len[a],
                             (uint32_t)(T * PT.len[a] * sizeof(double)), &bar[buf]);
        }
    };

    if (lane == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        fence_proxy_async_smem();
        if (is_bulk(tile)) issue_loads(tile, 0);
    }
    __syncwarp();
    if (flags_rt & CPLB_INPUTS_READY) pdl_wait();  // the first tile's x is already on its way; nothing is stored before this

    const int inst = lane / LPI, s = lane % LPI;
    bool stores_in_flight = false;
    for (int it = 0; tile < tiles; tile += stride, it++) {
        const int b = it & 1;
        double* xs = xbuf + (size_t)b * T * n;
        const long long i0 = tile * T;
        const int cnt = (io.N - i0) < T ? (int)(io.N - i0) : T;
        const bool bulk = is_bulk(tile);

        // prefetch the next tile's x into the other buffer (its previous contents were consumed an iteration ago)
        const long long next = tile + stride;
        if (lane == 0 && next < tiles && is_bulk(next)) {
            fence_proxy_async_smem();
            issue_loads(next, b ^ 1);
        }
        const double* ptile = pbuf + (size_t)b * ptot;
        if (bulk) {
            mbar_wait(&bar[b], (uint32_t)((it >> 1) & 1));
        } else {
            warp_copy(xs, io.x + i0 * n, cnt * n, lane);
            if (PERINST) {
#pragma unroll
                for (int a = 0; a < CPLB_NUM_INST_ARRAYS; a++)
                    if (PT.off[a] >= 0)
                        warp_copy(pbuf + (size_t)b * ptot + PT.off[a], parr[a] + i0 * PT.len[a], cnt * PT.len[a], lane);
            }
        }
        // the output tiles are about to be overwritten: the engine must have finished reading the previous ones
        if (stores_in_flight) {
            if (lane == 0) bulk_wait_read_all();
            stores_in_flight = false;
        }
        __syncwarp();

        // ---- compute: lane (inst, s) handles contacts s, s+LPI, ... and statics rows s, s+LPI, ... ----
        // The body is instantiated per parameter source: shared block, staged per-instance tiles, or per-instance arrays read
        // with read-only global loads (kept free of any shared-memory alternative so that the compiler can hoist and batch them).
        const unsigned live = __ballot_sync(0xffffffffu, inst < cnt);  // lanes that hold an instance of this tile (shuffle mask)
        auto compute_instance = [&](const auto& ps) {
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
                contact_rows<ENV>(P, ps, em, nc, j, k, c, F, p, nn, flags);
            }
            // CentroidalStatics rows: row r's running sum visits the contacts in sorted-name order
            // (CentroidalStatics.cpp:44-54); the six rows are independent, so they are dealt to the lanes.
            // Lanes of a warp hold different r, so the row is selected by indices, not by a switch: row 3+q is
            //   v += d_{q+1} F_{q+2} - d_{q+2} F_{q+1}   [(p - CoM) x F]
            // and its CoM block is  a -= sa*F[ia],  b -= sb*F[ib]  (:128-133; multiplying by +-1 is exact).
            if (flags & (CPLB_WANT_G | CPLB_WANT_J)) {
                const int L = jac_moment_row_len(nc);
                for (int r = s; r < 6; r += LPI) {
                    const bool mom = r >= 3;
                    const int q = mom ? r - 3 : 0;
                    const int i1 = q == 2 ? 0 : q + 1, i2 = q == 0 ? 2 : q - 1;  // (q+1)%3, (q+2)%3
                    const int ia = q == 2 ? 1 : 2, ib = q == 0 ? 1 : 0;
                    const double sa = q == 1 ? -1.0 : 1.0, sb = q == 1 ? 1.0 : -1.0;
                    double v = 0.0, a = 0.0, bb = 0.0;
                    for (int j = 0; j < nc; j++) {
                        const double* xk = xi + 3 + 9 * P.perm[j];
                        if (mom) {
                            const double x1 = xk[3 + i1] - xi[i1], x2 = xk[3 + i2] - xi[i2];  // p - CoM
                            v += x1 * xk[i2] - x2 * xk[i1];
                            a -= sa * xk[ia];
                            bb -= sb * xk[ib];
                        } else {
                            v += xk[r];
                        }
                    }
                    if (flags & CPLB_WANT_G) em.g(r, mom ? v - ps.wrench(r) : (v - ps.wrench(r)) + ps.mg(r));
                    if ((flags & CPLB_WANT_J) && mom) {
                        em.j(3 * nc + q * L + 0, a);
                        em.j(3 * nc + q * L + 1, bb);
                    }
                }
            }
            if (flags & CPLB_WANT_COST) {  // MinimizeCentroidalVariables.cpp:126-147, sorted order
                // every lane evaluates the terms of its own contacts; lane 0 of the instance collects them with shuffles
                // in sorted order j = 0, 1, ... -- the same running sum as the reference's loop (0.0 + t0 is exact)
                double cost = 0.0;
                for (int r = 0; r * LPI < nc; r++) {
                    const int j = r * LPI + s;
                    double t = 0.0;
                    if (j < nc) {
                        const int k = P.perm[j];
                        const double* xk = xi + 3 + 9 * k;
                        const double F[3] = {xk[0], xk[1], xk[2]};
                        const double p[3] = {xk[3], xk[4], xk[5]};
                        t = contact_cost(ps, P.reduction_order, k, F, p);
                    }
#pragma unroll
                    for (int q = 0; q < LPI; q++) {
                        const double tq = __shfl_sync(live, t, inst * LPI + q);
                        if (r * LPI + q < nc) cost += tq;
                    }
                }
                if (s == 0) {
                    cost += com_cost(ps, P.reduction_order, c);
                    costs[inst] = cost;
                }
            }
            if (s == 0) {
                if (flags & CPLB_WANT_GRAD) {
#pragma unroll
                    for (int q = 0; q < 3; q++) em.grad(q, ps.W_com() * (c[q] - ps.com_ref(q)));
                }
            }
        };
        if (inst < cnt) {
            if constexpr (PERINST) {
                if (ptot)
                    compute_instance(TileInstanceParams{P, Q, PT, ptile, inst});
                else
                    compute_instance(InstanceParams<false>{P, Q, i0 + inst, 0});
            } else {
                compute_instance(SharedParams{P});
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
            stores_in_flight = true;
            if (costs && lane < T) io.cost[i0 + lane] = costs[lane];
        } else {
            __syncwarp();
            if (gs) warp_copy(io.g + i0 * m, gs, cnt * m, lane);
            if (js) warp_copy(io.jac + i0 * nnz, js, cnt * nnz, lane);
            if (grads) warp_copy(io.grad + i0 * n, grads, cnt * n, lane);
            if (costs && lane < cnt) io.cost[i0 + lane] = costs[lane];
        }
        __syncwarp();  // all lanes done with xs and the tiles before the next iteration touches them
    }
    if (stores_in_flight && lane == 0) bulk_wait_read_all();  // shared memory must outlive the engine's reads
    __syncwarp();
}

// ================================================================================================
// launchers
// ================================================================================================

static const CplbInstParams kNoInstParams = {};

template <int ENV>
static cudaError_t launch_cm_env(const CplbParams& P, const CplbIo& io, unsigned flags, const CplbInstParams* Q, cudaStream_t st)
{
    // two 32-instance sub-blocks per CTA when they fit in 256 threads (measured: 22.1 vs 22.5 us on config 2)
    const int subs = (P.nc <= 4) ? 2 : 1;
    const unsigned blocks = (unsigned)((io.N + 32 * subs - 1) / (32 * subs));
    const int threads = 32 * P.nc * subs;
    const size_t smem = (size_t)subs * P.nc * (192 + 32) * sizeof(double);
    const unsigned gj = CPLB_WANT_G | CPLB_WANT_J;
    // shared-parameter batches of the two benchmark shapes, once they are large enough: one thread per instance.
    // Measured crossovers (B200): nc = 4: 88 % (split) vs 80 % (whole) at 65,536 but 89 % vs 95 % at 98,304;
    // nc = 8: 90.5 % vs 95 % already at 65,536.  Both reach 97-98 % at 1,048,576 (split: 95 %).
    // Superquadric with 8 contacts stays with the split kernel at every size: one thread holding 75 inputs plus the
    // closed-form normal Jacobian needs 250 registers and 320 B of local memory (2 CTAs per SM) and measures 64-65 % of the
    // roofline at 65,536 and 1,048,576 instances (profiles/r01_variants.md).
    const long long whole_from = P.nc == 4 ? 90112 : 49152;
    static const int forced = [] {  // CPLB_CM_KERNEL=split|whole: dispatch experiments only (tools/variant_table.py)
        const char* e = std::getenv("CPLB_CM_KERNEL");
        return !e ? 0 : (e[0] == 's' ? 1 : (e[0] == 'w' ? 2 : 0));
    }();
    const bool whole_ok = !Q && (P.nc == 4 || P.nc == 8);
    const bool whole_auto = io.N >= whole_from && !(ENV == CPLB_ENV_SUPERQUADRIC_K && P.nc == 8);
    if (whole_ok && (forced == 2 || (forced == 0 && whole_auto))) {
        const unsigned wb = (unsigned)((io.N + 127) / 128);
        if (P.nc == 4) {
            if ((flags & 15u) == gj) return launch_pdl(eval_component_major_whole<ENV, 4, gj>, wb, 128u, 0, st, P, io, flags);
            return launch_pdl(eval_component_major_whole<ENV, 4, 0u>, wb, 128u, 0, st, P, io, flags);
        }
        if ((flags & 15u) == gj) return launch_pdl(eval_component_major_whole<ENV, 8, gj>, wb, 128u, 0, st, P, io, flags);
        return launch_pdl(eval_component_major_whole<ENV, 8, 0u>, wb, 128u, 0, st, P, io, flags);
    }
    if (smem > 48 * 1024) {  // more than ~26 contacts: opt in to the larger dynamic shared memory (57 KB at 32 contacts)
        cudaError_t e = cudaFuncSetAttribute(eval_component_major_split<ENV, 0u, 32, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(eval_component_major_split<ENV, 0u, 32, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    if (Q) {  // per-instance parameter arrays
        if (P.nc <= 8) return launch_pdl(eval_component_major_split<ENV, 0u, 8, true>, blocks, threads, smem, st, P, io, flags, *Q);
        return launch_pdl(eval_component_major_split<ENV, 0u, 32, true>, blocks, threads, smem, st, P, io, flags, *Q);
    }
    if (P.nc <= 8) {
        if ((flags & 15u) == gj)
            return launch_pdl(eval_component_major_split<ENV, gj, 8, false>, blocks, threads, smem, st, P, io, flags, kNoInstParams);
        else
            return launch_pdl(eval_component_major_split<ENV, 0u, 8, false>, blocks, threads, smem, st, P, io, flags, kNoInstParams);
    }
    return launch_pdl(eval_component_major_split<ENV, 0u, 32, false>, blocks, threads, smem, st, P, io, flags, kNoInstParams);
}

cudaError_t launch_component_major(const CplbParams& P, const CplbIo& io, unsigned flags, const CplbInstParams* Q, cudaStream_t st)
{
    if (io.N <= 0) return cudaSuccess;
    if (io.ld >= (1LL << 29)) return cudaErrorInvalidValue;  // row pitch must fit 32 bits of bytes
    switch (P.env) {
    case CPLB_ENV_NONE_K: return launch_cm_env<CPLB_ENV_NONE_K>(P, io, flags, Q, st);
    case CPLB_ENV_GROUND_K: return launch_cm_env<CPLB_ENV_GROUND_K>(P, io, flags, Q, st);
    default: return launch_cm_env<CPLB_ENV_SUPERQUADRIC_K>(P, io, flags, Q, st);
    }
}

template <int ENV, int LPI, int WARPS, unsigned FLAGS, bool PERINST>
static cudaError_t launch_im_kernel(const CplbParams& P, const CplbIo& io, unsigned flags, size_t smem, const CplbInstParams* Q,
                                    const CplbParamTile& PT, cudaStream_t st)
{
    constexpr int T = 32 / LPI;
    auto kern = eval_instance_major<ENV, LPI, WARPS, FLAGS, PERINST>;
    // resident CTAs per SM and the SM count are fixed per (kernel, smem, device): looked up once
    struct Cfg { int device = -1; size_t smem = 0; int resident = 0; };
    static thread_local Cfg cfg;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (cfg.device != dev || cfg.smem != smem) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        int per_sm = 0, sms = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, WARPS * 32, smem);
        if (e != cudaSuccess) return e;
        e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (e != cudaSuccess) return e;
        cfg.device = dev;
        cfg.smem = smem;
        cfg.resident = (per_sm > 0 ? per_sm : 1) * sms;
    }
    const long long tiles = (io.N + T - 1) / T;
    const long long want = (tiles + WARPS - 1) / WARPS;
    const unsigned blocks = (unsigned)(want < cfg.resident ? want : cfg.resident);
    auto al16 = [](const void* p) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    int aligned16 = al16(io.x) && al16(io.g) && al16(io.jac) && al16(io.grad) && (T % 2 == 0);
    if (Q) {
        const double* const parr[CPLB_NUM_INST_ARRAYS] = {Q->mass, Q->wrench, Q->mu, Q->F_thr, Q->ground_z, Q->com_ref, Q->W_com,
                                                           Q->p_ref, Q->F_ref, Q->W_p, Q->W_F};
        for (int a = 0; a < CPLB_NUM_INST_ARRAYS; a++)
            if (PT.off[a] >= 0) aligned16 = aligned16 && al16(parr[a]);
    }
    return launch_pdl(kern, blocks, WARPS * 32, smem, st, P, io, flags, aligned16, Q ? *Q : kNoInstParams, PT);
}

// Which per-instance arrays the instance-major kernel stages for the requested outputs, and where (CplbParamTile).
static CplbParamTile make_param_tile(const CplbParams& P, const CplbInstParams* Q, unsigned flags, int T)
{
    CplbParamTile PT = {};
    const int nc = P.nc;
    const int lens[CPLB_NUM_INST_ARRAYS] = {1, 6, 1, nc, 1, 3, 1, 3 * nc, 3 * nc, nc, nc};
    for (int a = 0; a < CPLB_NUM_INST_ARRAYS; a++) {
        PT.off[a] = -1;
        PT.len[a] = lens[a];
    }
    if (!Q) return PT;
    const double* const parr[CPLB_NUM_INST_ARRAYS] = {Q->mass, Q->wrench, Q->mu, Q->F_thr, Q->ground_z, Q->com_ref, Q->W_com,
                                                       Q->p_ref, Q->F_ref, Q->W_p, Q->W_F};
    // Staged for constraint-only evaluations: what the constraint rows read (9 + nc doubles per instance; measured 55 -> 61 %
    // of the roofline at 65,536 x 4 contacts, 69 -> 75 % at 1,048,576).  With the cost or the gradient requested nothing is
    // staged and every array is read with read-only global loads as before: that variant is bound by the cost lane's
    // critical path and by resident warps, and both staging everything (39 %) and staging only the constraint arrays (32 %)
    // measured below the plain loads (43 %).
    // Ground problems only: without an environment the extra 1.5 KB per warp costs a resident CTA (59 vs 53 KB: 70 -> 62 % at
    // 1,048,576 x 4 contacts) and the Superquadric rows are bound by their arithmetic (48 vs 45 %, 51 vs 52 %).
    if (flags & (CPLB_WANT_COST | CPLB_WANT_GRAD)) return PT;
    if (P.env != CPLB_ENV_GROUND_K) return PT;
    const bool GJ = flags & (CPLB_WANT_G | CPLB_WANT_J), C = false;
    const bool need[CPLB_NUM_INST_ARRAYS] = {GJ, GJ, GJ, GJ, GJ && P.env == CPLB_ENV_GROUND_K, C, C, C, C, C, C};
    for (int a = 0; a < CPLB_NUM_INST_ARRAYS; a++)
        if (parr[a] && need[a]) {
            PT.off[a] = PT.total;
            PT.total += T * lens[a];
        }
    return PT;
}

template <int ENV, int LPI>
static cudaError_t launch_im_cfg(const CplbParams& P, const CplbIo& io, unsigned flags, const CplbInstParams* Q, cudaStream_t st)
{
    constexpr int T = 32 / LPI;
    const CplbParamTile PT = make_param_tile(P, Q, flags & 15u, T);
    const size_t per_warp = (tile_doubles(T, P.n, P.m, P.nnz, flags & 15u) + 2 * (size_t)PT.total) * sizeof(double) + 2 * sizeof(uint64_t);
    const unsigned gj = CPLB_WANT_G | CPLB_WANT_J;
    if (4 * per_warp <= 72 * 1024) {  // the common shapes: 4 warps per CTA, 3 CTAs per SM
        if (Q) return launch_im_kernel<ENV, LPI, 4, 0u, true>(P, io, flags, 4 * per_warp, Q, PT, st);
        if ((flags & 15u) == gj) return launch_im_kernel<ENV, LPI, 4, gj, false>(P, io, flags, 4 * per_warp, Q, PT, st);
        return launch_im_kernel<ENV, LPI, 4, 0u, false>(P, io, flags, 4 * per_warp, Q, PT, st);
    }
    if (per_warp > 227 * 1024) return cudaErrorInvalidConfiguration;
    if (Q) return launch_im_kernel<ENV, LPI, 1, 0u, true>(P, io, flags, per_warp, Q, PT, st);
    return launch_im_kernel<ENV, LPI, 1, 0u, false>(P, io, flags, per_warp, Q, PT, st);  // many contacts: one warp per CTA
}

template <int ENV>
static cudaError_t launch_im_env(const CplbParams& P, const CplbIo& io, unsigned flags, const CplbInstParams* Q, cudaStream_t st)
{
    // lanes per instance: the smallest power of two >= nc (capped at 8; more contacts loop)
    if (P.nc <= 1) return launch_im_cfg<ENV, 1>(P, io, flags, Q, st);
    if (P.nc <= 2) return launch_im_cfg<ENV, 2>(P, io, flags, Q, st);
    if (P.nc <= 4) return launch_im_cfg<ENV, 4>(P, io, flags, Q, st);
    return launch_im_cfg<ENV, 8>(P, io, flags, Q, st);
}

cudaError_t launch_instance_major(const CplbParams& P, const CplbIo& io, unsigned flags, const CplbInstParams* Q, cudaStream_t st)
{
    if (io.N <= 0) return cudaSuccess;
    switch (P.env) {
    case CPLB_ENV_NONE_K: return launch_im_env<CPLB_ENV_NONE_K>(P, io, flags, Q, st);
    case CPLB_ENV_GROUND_K: return launch_im_env<CPLB_ENV_GROUND_K>(P, io, flags, Q, st);
    default: return launch_im_env<CPLB_ENV_SUPERQUADRIC_K>(P, io, flags, Q, st);
    }
}

}  // namespace cplb
