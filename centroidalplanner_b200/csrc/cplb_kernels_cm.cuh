// cplb_kernels_cm.cuh -- component-major (struct-of-arrays) evaluation kernels and their launcher template.  One translation
// unit per environment kind instantiates launch_cm_env<ENV> (cplb_kernels_cm_*.cu) so that the kinds compile in parallel.
#ifndef CPLB_KERNELS_CM_CUH
#define CPLB_KERNELS_CM_CUH

#include "cplb_launch.cuh"

namespace cplb {

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

// cm_kernel: CPLB_CM_AUTO (pick by shape and batch size, below), CPLB_CM_SPLIT, CPLB_CM_WHOLE (cplb_set_component_major_kernel:
// parity tests and dispatch measurements run BOTH kernels on the same inputs; a forced `whole` silently stays with the split
// kernel where the one-thread-per-instance kernel does not exist: per-instance parameters, contact counts other than 4 / 8).
template <int ENV>
cudaError_t launch_cm_env(const CplbParams& P, const CplbIo& io, unsigned flags, const CplbInstParams* Q, int cm_kernel, cudaStream_t st)
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
    const int forced = cm_kernel;
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
        // the attribute is process-wide per device and must never shrink under a concurrent launch: always the 32-contact size
        const int max_smem = (int)((size_t)CPLB_KMAX_CONTACTS * (192 + 32) * sizeof(double));
        cudaError_t e = cudaFuncSetAttribute(eval_component_major_split<ENV, 0u, 32, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(eval_component_major_split<ENV, 0u, 32, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
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

}  // namespace cplb
#endif
