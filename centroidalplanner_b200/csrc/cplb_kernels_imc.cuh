// cplb_kernels_imc.cuh -- instance-major evaluation, one CTA per tile ("cta tile" kernel).  Included by cplb_kernels_im.cuh.
//
// Same buffers as eval_instance_major (per-instance contiguous slices buf[i*len + e], what an IPOPT thread consumes), other
// thread mapping: a CTA owns a tile of TI consecutive instances and thread t of the CTA is (instance t % TI, contact slot
// t / TI) -- for TI = 32 a warp is ONE contact of 32 consecutive instances, exactly the mapping of the component-major
// per-contact kernel.  What that buys over the warp-tile kernel (measured there, profiles/r01_full_instance_major_*):
//   * shared-memory reads of x are conflict-free: lanes step through instances, the stride is n = 3 + 9 nc doubles (odd),
//     where the warp-tile kernel's (instance, contact) lane grid paid 4 wavefronts for every 64-bit access;
//   * no per-lane loops over contacts or rows and no lane-dependent row selection: the six CentroidalStatics sums cross
//     the contact warps through a small shared-memory exchange (one __syncthreads) and are added in sorted-name order
//     (CentroidalStatics.cpp:44-54) -- the exchange of the component-major kernel;
//   * every output slot is one STS with an immediate offset from three per-thread base pointers;
//   * the x-independent Jacobian slots (cplb_get_jacobian_constants: 72 of 174 for a 4-contact Ground problem) are
//     written into the CTA's output tile ONCE, before the first tile: the tile buffer persists across the CTA's tiles
//     and nothing else ever writes those slots.
// Two launch-level choices on top, both by measurement (profiles/r02_instance_major.md): full rows with shared parameters give
// every (instance, contact) pair TWO threads with complementary shares of the contact's rows (ROLES = 2, contact_rows<..., PARTS>),
// and on Ground / no environment up to 4 contacts they use 16-instance tiles on a grid of 2x (4x for long launches) the CTAs that
// fit at once instead of a strictly persistent grid (launch_imc_roles).
// Data movement is unchanged: one cp.async.bulk (TMA) load of the tile's T*n contiguous doubles of x into one of two
// buffers (the next tile's load is issued before the current one is consumed), outputs leave with cp.async.bulk stores and
// the CTA waits for the engine to have READ the tile only right before the next tile's first write to it.
#ifndef CPLB_KERNELS_IMC_CUH
#define CPLB_KERNELS_IMC_CUH

namespace cplb {

// exchange rows per contact: the three moment terms (the force terms are read back from the x tile), plus the cost term
__host__ __device__ inline int cta_exchange_rows(unsigned flags)
{
    return ((flags & (CPLB_WANT_G | CPLB_WANT_J)) ? 3 : 0) + ((flags & CPLB_WANT_COST) ? 1 : 0);
}

// shared memory of one CTA, in doubles: [x0: TI*n][x1: TI*n][g: TI*m][jac: TI*nnz][grad: TI*n][cost: TI][exchange: nc*rows*TI]; then 2 mbarriers.
// 4 contacts, g + Jacobian: 9,410 doubles = 75,280 B -- three CTAs per SM (227 KB, 1 KB reserved per CTA).
// (nnz = doubles per instance of the Jacobian slice: all structural slots, or only the x-dependent ones of a packed evaluation)
__host__ __device__ inline size_t cta_tile_doubles(int TI, int nc, int n, int m, int nnz, unsigned flags)
{
    size_t d = 2 * (size_t)TI * n;
    if (flags & CPLB_WANT_G) d += (size_t)TI * m;
    if (flags & CPLB_WANT_J) d += (size_t)TI * nnz;
    if (flags & CPLB_WANT_GRAD) d += (size_t)TI * n;
    if (flags & CPLB_WANT_COST) d += (size_t)TI;
    d += (size_t)nc * cta_exchange_rows(flags) * TI;
    return (d + 1) & ~(size_t)1;
}

__device__ __forceinline__ void cta_copy(double* dst, const double* src, int count, int tid, int nthreads)
{
    for (int e = tid; e < count; e += nthreads) dst[e] = src[e];
}

template <int ENV, int TI, unsigned FLAGS, bool PERINST, int PACKED, int ROLES>
__global__ void __launch_bounds__(128 * ROLES) eval_instance_major_cta(const __grid_constant__ CplbParams P, const CplbIo io,
                                                                 const unsigned flags_rt, const int aligned16,
                                                                 const __grid_constant__ CplbInstParams Q)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    pdl_trigger();
    if (!(flags_rt & CPLB_INPUTS_READY)) pdl_wait();
    const unsigned flags = FLAGS ? FLAGS : (flags_rt & 15u);
    const int tid = threadIdx.x, nthreads = blockDim.x;
    using M = JacMap<ENV, PACKED>;
    const int nc = P.nc, n = P.n, m = P.m, nnz = PACKED ? M::per_instance(P.nc) : P.nnz;  // nnz: doubles per instance of the jac slice
    // ROLES == 2: every (instance, contact) pair has two threads with complementary shares of the contact's rows (the first half of
    // the CTA takes CentroidalStatics' entries and the cheaper of {environment rows, FrictionCone}, the second half the other) --
    // twice the warps per SM on the same shared memory, which is what the kernel's latency chains need
    constexpr int kPartsA = ROLES == 1 ? kPartAll : (ENV == CPLB_ENV_SUPERQUADRIC_K ? (kPartStatics | kPartFriction) : (kPartStatics | kPartEnvironment));
    constexpr int kPartsB = kPartAll & ~kPartsA;
    const int per_role = nthreads / ROLES;
    const int role = ROLES == 1 ? 0 : tid / per_role, idx = ROLES == 1 ? tid : tid - role * per_role;
    const int inst = idx % TI, j = idx / TI;  // j: sorted rank of this thread's contact; j >= nc: padding thread (barriers only)
    const bool has_contact = j < nc;
    const int k = has_contact ? P.perm[j] : 0;

    double* xbuf = reinterpret_cast<double*>(smem_raw);
    double* cur = xbuf + 2 * (size_t)TI * n;
    double *gs = nullptr, *js = nullptr, *grads = nullptr, *costs = nullptr;
    if (flags & CPLB_WANT_G) { gs = cur; cur += (size_t)TI * m; }
    if (flags & CPLB_WANT_J) { js = cur; cur += (size_t)TI * nnz; }
    if (flags & CPLB_WANT_GRAD) { grads = cur; cur += (size_t)TI * n; }
    if (flags & CPLB_WANT_COST) { costs = cur; cur += (size_t)TI; }
    double* exch = cur;  // [nc][ER][TI]: moment term (3) and cost term (1) of every contact, as requested
    const int ER = cta_exchange_rows(flags);
    const int cost_row = (flags & (CPLB_WANT_G | CPLB_WANT_J)) ? 3 : 0;
    uint64_t* bar = reinterpret_cast<uint64_t*>(reinterpret_cast<double*>(smem_raw) + cta_tile_doubles(TI, nc, n, m, nnz, flags));

    // Persistent grid: CTA b walks tiles b, b + G, b + 2G, ... of TI instances each.  (Measured and not kept, profiles/r02_instance_major.md:
    // balanced contiguous ranges per CTA and shorter first tiles for the co-resident CTAs of an SM, to keep them out of lock step.)
    const int G = (int)gridDim.x;
    auto range_begin = [&](int it) -> long long { return ((long long)blockIdx.x + (long long)it * G) * TI; };
    auto range_count = [&](int it) -> int {
        const long long left = io.N - range_begin(it);
        return left < TI ? (int)left : TI;
    };
    auto range_valid = [&](int it) { return range_begin(it) < io.N; };
    // bulk copies need 16-byte aligned addresses and sizes: complete tiles (TI is even) of 16B-aligned buffers
    auto is_bulk = [&](int it) { return aligned16 && range_count(it) == TI; };
    auto issue_load = [&](int it, int buf) {
        const uint32_t bytes = (uint32_t)(range_count(it) * n * sizeof(double));
        mbar_expect_tx(&bar[buf], bytes);
        bulk_g2s(xbuf + (size_t)buf * TI * n, io.x + range_begin(it) * n, bytes, &bar[buf]);
    };
    if (tid == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        fence_proxy_async_smem();
        if (range_valid(0) && is_bulk(0)) issue_load(0, 0);
    }
    TileEmitter em{gs ? gs + (size_t)inst * m : nullptr, js ? js + (size_t)inst * nnz : nullptr, grads ? grads + (size_t)inst * n : nullptr};
    if (!PACKED && (flags & CPLB_WANT_J) && has_contact && role == 0) contact_constant_slots<ENV>(em, nc, j, k);  // once: the tile buffer persists
    __syncthreads();  // mbarrier inits visible to every thread before anyone polls them
    if (flags_rt & CPLB_INPUTS_READY) pdl_wait();  // the first tile's x is already on its way; nothing is stored before this

    bool stores_in_flight = false;
    for (int it = 0; range_valid(it); it++) {
        const int b = it & 1;
        const double* xs = xbuf + (size_t)b * TI * n;
        const long long i0 = range_begin(it);
        const int cnt = range_count(it);
        const bool bulk = is_bulk(it);
        const bool live = has_contact && inst < cnt;

        // prefetch the next tile's x into the other buffer: every thread passed the exchange barrier of the previous iteration,
        // i.e. has its operands of that buffer's tile in registers
        if (tid == 0 && range_valid(it + 1) && is_bulk(it + 1)) {
            fence_proxy_async_smem();
            issue_load(it + 1, b ^ 1);
        }
        // per-instance parameters: read-only global loads issued before the wait for x, consumed after it
        const auto ps = ParamSource<PERINST, false>::make(P, Q, i0 + (inst < cnt ? inst : 0), 0);
        double mu = 0.0, F_thr = 0.0, gz = 0.0;
        if (PERINST && live) {
            mu = ps.mu();
            F_thr = ps.F_thr(k);
            if (ENV == CPLB_ENV_GROUND_K) gz = ps.ground_z();
        }
        (void)mu; (void)F_thr; (void)gz;

        if (bulk) {
            mbar_wait(&bar[b], (uint32_t)((it >> 1) & 1));
        } else {
            cta_copy(xbuf + (size_t)b * TI * n, io.x + i0 * n, cnt * n, tid, nthreads);
            __syncthreads();
        }
        // the output tile is about to be overwritten: the engine must have finished reading the previous one
        if (stores_in_flight) {
            if (tid == 0) bulk_wait_read_all();
            stores_in_flight = false;
            __syncthreads();
        }

        const double* xi = xs + (size_t)inst * n;
        double c[3] = {0.0, 0.0, 0.0}, F[3] = {0.0, 0.0, 0.0}, p[3] = {0.0, 0.0, 0.0}, nn[3] = {0.0, 0.0, 0.0};
        if (live) {
            const double* xk = xi + 3 + 9 * k;
#pragma unroll
            for (int q = 0; q < 3; q++) {
                c[q] = xi[q];
                F[q] = xk[q];
                p[q] = xk[3 + q];
                nn[q] = xk[6 + q];
            }
            double* mine = exch + (size_t)j * ER * TI + inst;
            if (role != 0) {
                // the exchange is the first thread's
            } else if (flags & (CPLB_WANT_G | CPLB_WANT_J)) {
                const double d0 = p[0] - c[0], d1 = p[1] - c[1], d2 = p[2] - c[2];
                mine[0 * TI] = d1 * F[2] - d2 * F[1];  // (p - CoM).cross(F), CentroidalStatics.cpp:53
                mine[1 * TI] = d2 * F[0] - d0 * F[2];
                mine[2 * TI] = d0 * F[1] - d1 * F[0];
            }
            if ((flags & CPLB_WANT_COST) && role == 0) mine[cost_row * TI] = contact_cost(ps, P.reduction_order, k, F, p);
            if constexpr (PERINST) {
                // FrictionCone / Ground parameters were fetched above; the rest of ps is read where it is used
                struct Hoisted {
                    const decltype(ps)& base;
                    double mu_, F_thr_, gz_;
                    __device__ __forceinline__ double mu() const { return mu_; }
                    __device__ __forceinline__ double F_thr(int) const { return F_thr_; }
                    __device__ __forceinline__ double ground_z() const { return gz_; }
                    __device__ __forceinline__ double W_F(int kk) const { return base.W_F(kk); }
                    __device__ __forceinline__ double W_p(int kk) const { return base.W_p(kk); }
                    __device__ __forceinline__ double F_ref(int kk, int q) const { return base.F_ref(kk, q); }
                    __device__ __forceinline__ double p_ref(int kk, int q) const { return base.p_ref(kk, q); }
                } hp{ps, mu, F_thr, gz};
                if (role == 0) contact_rows<ENV, false, PACKED, kPartsA>(P, hp, em, nc, j, k, c, F, p, nn, flags);
                else contact_rows<ENV, false, PACKED, kPartsB>(P, hp, em, nc, j, k, c, F, p, nn, flags);
            } else {
                if (role == 0) contact_rows<ENV, false, PACKED, kPartsA>(P, ps, em, nc, j, k, c, F, p, nn, flags);
                else contact_rows<ENV, false, PACKED, kPartsB>(P, ps, em, nc, j, k, c, F, p, nn, flags);
            }
        }
        __syncthreads();  // exchange complete; every thread is done with this tile's x buffer

        if (live) {
            if (flags & (CPLB_WANT_G | CPLB_WANT_J)) {
                for (int r = j + role * nc; r < 6; r += ROLES * nc) {
                    // row r's running sum visits the contacts in sorted-name order (CentroidalStatics.cpp:44-54): force rows
                    // read the forces back from the x tile (still intact: the prefetch went to the other buffer)
                    double v = 0.0;
                    if (r < 3) {
                        for (int jj = 0; jj < nc; jj++) v += xi[3 + 9 * P.perm[jj] + r];
                    } else {
                        const double* col = exch + (r - 3) * TI + inst;
                        for (int jj = 0; jj < nc; jj++) v += col[(size_t)jj * ER * TI];
                    }
                    if (flags & CPLB_WANT_G) em.g(r, r < 3 ? (v - ps.wrench(r)) + ps.mg(r) : v - ps.wrench(r));  // :56-57
                    if ((flags & CPLB_WANT_J) && r >= 3) {
                        // CoM block (:128-133): row 3 <- (Fz, -Fy), row 4 <- (-Fz, Fx), row 5 <- (Fy, -Fx), each "acc -= term"
                        const int ia = r == 3 ? 2 : (r == 4 ? 2 : 1), ib = r == 3 ? 1 : (r == 4 ? 0 : 0);
                        const bool nega = (r == 4), negb = (r != 4);
                        double a = 0.0, bb = 0.0;
                        for (int jj = 0; jj < nc; jj++) {
                            const double* Fj = xi + 3 + 9 * P.perm[jj];
                            const double fa = Fj[ia], fb = Fj[ib];
                            a -= nega ? -fa : fa;
                            bb -= negb ? -fb : fb;
                        }
                        em.j(M::moment(nc, r - 3, 0), a);
                        em.j(M::moment(nc, r - 3, 1), bb);
                    }
                }
            }
            if (j == 0 && role == 0) {
                if (flags & CPLB_WANT_COST) {  // MinimizeCentroidalVariables.cpp:126-147: contacts in sorted order, then the CoM term
                    double cost = 0.0;
                    for (int jj = 0; jj < nc; jj++) cost += exch[(size_t)jj * ER * TI + cost_row * TI + inst];
                    cost += com_cost(ps, P.reduction_order, c);
                    costs[inst] = cost;
                }
                if (flags & CPLB_WANT_GRAD) {
#pragma unroll
                    for (int q = 0; q < 3; q++) em.grad(q, ps.W_com() * (c[q] - ps.com_ref(q)));
                }
            }
        }

        // ---- ship the tile --------------------------------------------------------------------------
        if (bulk) {
            fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the async (TMA) proxy
            __syncthreads();
            if (tid == 0) {
                if (gs) bulk_s2g(io.g + i0 * m, gs, (uint32_t)(cnt * m * sizeof(double)));
                if (js) bulk_s2g(io.jac + i0 * nnz, js, (uint32_t)(cnt * nnz * sizeof(double)));
                if (grads) bulk_s2g(io.grad + i0 * n, grads, (uint32_t)(cnt * n * sizeof(double)));
                bulk_commit();
            }
            stores_in_flight = true;
            if (costs && tid < cnt) io.cost[i0 + tid] = costs[tid];
        } else {
            __syncthreads();
            if (gs) cta_copy(io.g + i0 * m, gs, cnt * m, tid, nthreads);
            if (js) cta_copy(io.jac + i0 * nnz, js, cnt * nnz, tid, nthreads);
            if (grads) cta_copy(io.grad + i0 * n, grads, cnt * n, tid, nthreads);
            if (costs && tid < cnt) io.cost[i0 + tid] = costs[tid];
            __syncthreads();  // the copies read the tile; the next iteration writes it
        }
    }
    if (stores_in_flight && tid == 0) bulk_wait_read_all();  // shared memory must outlive the engine's reads
    __syncthreads();
}

// Grid size resident at once per (kernel, device, smem, threads), looked up once.  The opt-in dynamic shared-memory limit of a
// kernel on a device is only ever RAISED: the attribute belongs to the function, not to a thread, and lowering it for a smaller
// request would make a concurrent larger launch from another host thread fail.
inline cudaError_t resident_grid(const void* kern, int threads, size_t smem, int* resident_out)
{
    struct Entry { const void* kern; int device; size_t smem; int threads; int resident; };
    struct Limit { const void* kern; int device; size_t smem; };
    static std::mutex mu;
    static std::vector<Entry> seen;
    static std::vector<Limit> limits;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> lk(mu);
    Limit* lim = nullptr;
    for (auto& l : limits)
        if (l.kern == kern && l.device == dev) lim = &l;
    if (!lim) {
        limits.push_back({kern, dev, 0});
        lim = &limits.back();
    }
    if (smem > lim->smem) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        lim->smem = smem;
    }
    for (const auto& s : seen)
        if (s.kern == kern && s.device == dev && s.smem == smem && s.threads == threads) {
            *resident_out = s.resident;
            return cudaSuccess;
        }
    int per_sm = 0, sms = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem);
    if (e != cudaSuccess) return e;
    e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return e;
    const int resident = (per_sm > 0 ? per_sm : 1) * sms;
    seen.push_back({kern, dev, smem, threads, resident});
    *resident_out = resident;
    return cudaSuccess;
}

template <int ENV, int TI, unsigned FLAGS, bool PERINST, int PACKED, int ROLES>
cudaError_t launch_imc_roles(const CplbParams& P, const CplbIo& io, unsigned flags, const CplbInstParams* Q, cudaStream_t st)
{
    auto kern = eval_instance_major_cta<ENV, TI, FLAGS, PERINST, PACKED, ROLES>;
    const int threads = ROLES * (((TI * P.nc + 31) / 32) * 32);
    const int jac_len = PACKED ? JacMap<ENV, PACKED>::per_instance(P.nc) : P.nnz;
    const size_t smem = cta_tile_doubles(TI, P.nc, P.n, P.m, jac_len, flags & 15u) * sizeof(double) + 2 * sizeof(uint64_t);
    int resident = 0;
    cudaError_t e = resident_grid(reinterpret_cast<const void*>(kern), threads, smem, &resident);
    if (e != cudaSuccess) return e;
    const long long tiles = (io.N + TI - 1) / TI;
    // Full rows with shared parameters on Ground / no environment: a grid of 2x (4x for long launches) the CTAs that fit at once.
    // The CTAs of the first wave finish at different times and the waiting ones take their places, which keeps the co-resident
    // CTAs of an SM out of lock step (all computing, nobody storing) -- measured, profiles/r02_instance_major.md: 65,536 ground4
    // instances 85.8 -> 88.1 % of the roofline, 1,048,576: 89.1 -> 92 %.  The other variants lose with it and stay at 1x.
    const bool oversubscribe = PACKED == 0 && !PERINST && ENV != CPLB_ENV_SUPERQUADRIC_K && P.nc <= 4;
    const long long want = !oversubscribe ? resident : (long long)resident * (tiles >= 16LL * resident ? 4 : 2);
    const unsigned blocks = (unsigned)(tiles < want ? tiles : want);
    auto al16 = [](const void* p) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    // complete tiles hold TI (even) instances: TI*n*8, TI*m*8, TI*nnz*8 bytes are multiples of 16
    const int aligned16 = al16(io.x) && al16(io.g) && al16(io.jac) && al16(io.grad) && (TI % 2 == 0);
    return launch_pdl(kern, blocks, (unsigned)threads, smem, st, P, io, flags, aligned16, Q ? *Q : kNoInstParams);
}

template <int ENV, int TI, unsigned FLAGS, bool PERINST, int PACKED>
cudaError_t launch_imc_kernel(const CplbParams& P, const CplbIo& io, unsigned flags, const CplbInstParams* Q, cudaStream_t st)
{
    // Two threads per (instance, contact) where it pays (measured, profiles/r02_instance_major.md): full Jacobian rows with shared
    // parameters (+1 point Ground, +4 points Superquadric at 65,536 instances).  Packed / computed slices keep one thread: their
    // smaller tiles let more CTAs share an SM, which the doubled register footprint would undo (computed: 89.8 % -> 79 %), and
    // per-instance parameters would be fetched twice (73.9 % -> 64 %).
    constexpr int kRoles = (PACKED == 0 && !PERINST) ? 2 : 1;
    return launch_imc_roles<ENV, TI, FLAGS, PERINST, PACKED, kRoles>(P, io, flags, Q, st);
}

// tile size by contact count: threads = TI * nc stays <= 256, the CTA's shared memory <= ~75 KB (3 CTAs per SM)
inline int cta_tile_instances(int nc) { return nc <= 4 ? 32 : (nc <= 8 ? 16 : (nc <= 16 ? 8 : 4)); }

template <int ENV, int TI>
cudaError_t launch_imc_ti(const CplbParams& P, const CplbIo& io, unsigned flags, const CplbInstParams* Q, cudaStream_t st)
{
    const unsigned gj = CPLB_WANT_G | CPLB_WANT_J;
    if (flags & CPLB_JAC_COMPUTED_K) {
        if (Q) return launch_imc_kernel<ENV, TI, 0u, true, 2>(P, io, flags, Q, st);
        if ((flags & 15u) == gj) return launch_imc_kernel<ENV, TI, gj, false, 2>(P, io, flags, Q, st);
        return launch_imc_kernel<ENV, TI, 0u, false, 2>(P, io, flags, Q, st);
    }
    if (flags & CPLB_JAC_PACKED_K) {
        if (Q) return launch_imc_kernel<ENV, TI, 0u, true, 1>(P, io, flags, Q, st);
        if ((flags & 15u) == gj) return launch_imc_kernel<ENV, TI, gj, false, 1>(P, io, flags, Q, st);
        return launch_imc_kernel<ENV, TI, 0u, false, 1>(P, io, flags, Q, st);
    }
    if (Q) return launch_imc_kernel<ENV, TI, 0u, true, 0>(P, io, flags, Q, st);
    if ((flags & 15u) == gj) return launch_imc_kernel<ENV, TI, gj, false, 0>(P, io, flags, Q, st);
    return launch_imc_kernel<ENV, TI, 0u, false, 0>(P, io, flags, Q, st);
}

template <int ENV>
cudaError_t launch_imc_env(const CplbParams& P, const CplbIo& io, unsigned flags, const CplbInstParams* Q, cudaStream_t st)
{
    int ti = cta_tile_instances(P.nc);
    // (same measurement: with the oversubscribed grid, half-size tiles are the better granularity up to 4 contacts)
    if (!(flags & (CPLB_JAC_PACKED_K | CPLB_JAC_COMPUTED_K)) && Q == nullptr && ENV != CPLB_ENV_SUPERQUADRIC_K && P.nc <= 4) ti = 16;
    switch (ti) {
    case 32: return launch_imc_ti<ENV, 32>(P, io, flags, Q, st);
    case 16: return launch_imc_ti<ENV, 16>(P, io, flags, Q, st);
    case 8: return launch_imc_ti<ENV, 8>(P, io, flags, Q, st);
    default: return launch_imc_ti<ENV, 4>(P, io, flags, Q, st);
    }
}

}  // namespace cplb
#endif
