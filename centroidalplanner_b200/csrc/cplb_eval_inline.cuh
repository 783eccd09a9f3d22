// cplb_eval_inline.cuh -- the evaluation of ONE instance by ONE thread, as a device function.
//
// Same arithmetic, same order as the batched kernels (it is eval_component_major_whole's body with a run-time contact count and
// plain pointers: CentroidalStatics sums and CoM block in sorted-name order, contact_rows, cost in sorted order), so its
// outputs are the batched evaluator's bit for bit.  Used by the solve round's tail kernel (cplb_solver.cu), where a handful of
// straggling instances iterate on their own and a kernel launch per evaluation would cost more than the evaluation.
#ifndef CPLB_EVAL_INLINE_CUH
#define CPLB_EVAL_INLINE_CUH

#include "cplb_device.cuh"

namespace cplb {

struct PointerEmitter {
    double* gp;
    double* jp;
    double* gradp;
    __device__ __forceinline__ void g(int row, double v) const { gp[row] = v; }
    __device__ __forceinline__ void j(int slot, double v) const { jp[slot] = v; }
    __device__ __forceinline__ void grad(int col, double v) const { gradp[col] = v; }
};

template <int ENV, class PS>
__device__ void eval_one_instance_env(const CplbParams& P, const PS& ps, const double* x, double* g, double* jac, double* cost_out, double* grad,
                                      unsigned flags)
{
    const int nc = P.nc;
    PointerEmitter em{g, jac, grad};
    const double c[3] = {x[0], x[1], x[2]};
    double v[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    double a31 = 0.0, a32 = 0.0, a40 = 0.0, a42 = 0.0, a50 = 0.0, a51 = 0.0, cost = 0.0;
    for (int j = 0; j < nc; j++) {  // sorted-name order (CentroidalStatics.cpp:44-54, :121-135)
        const int k = P.perm[j];
        const double* xk = x + 3 + 9 * k;
        const double F[3] = {xk[0], xk[1], xk[2]}, p[3] = {xk[3], xk[4], xk[5]}, n[3] = {xk[6], xk[7], xk[8]};
        const double d0 = p[0] - c[0], d1 = p[1] - c[1], d2 = p[2] - c[2];
        v[0] += F[0];
        v[1] += F[1];
        v[2] += F[2];
        v[3] += d1 * F[2] - d2 * F[1];
        v[4] += d2 * F[0] - d0 * F[2];
        v[5] += d0 * F[1] - d1 * F[0];
        a31 -= F[2];
        a32 -= -F[1];
        a40 -= -F[2];
        a42 -= F[0];
        a50 -= F[1];
        a51 -= -F[0];
        contact_rows<ENV>(P, ps, em, nc, j, k, c, F, p, n, flags);
        if (flags & CPLB_WANT_COST) cost += contact_cost(ps, P.reduction_order, k, F, p);
    }
    if (flags & CPLB_WANT_G) {
#pragma unroll
        for (int r = 0; r < 6; r++) em.g(r, r < 3 ? (v[r] - ps.wrench(r)) + ps.mg(r) : v[r] - ps.wrench(r));  // :56-57
    }
    if (flags & CPLB_WANT_J) {
        const int L = jac_moment_row_len(nc), s3 = 3 * nc;
        em.j(s3 + 0, a31);
        em.j(s3 + 1, a32);
        em.j(s3 + L + 0, a40);
        em.j(s3 + L + 1, a42);
        em.j(s3 + 2 * L + 0, a50);
        em.j(s3 + 2 * L + 1, a51);
    }
    if (flags & CPLB_WANT_COST) {
        cost += com_cost(ps, P.reduction_order, c);
        *cost_out = cost;
    }
    if (flags & CPLB_WANT_GRAD) {
#pragma unroll
        for (int q = 0; q < 3; q++) em.grad(q, ps.W_com() * (c[q] - ps.com_ref(q)));
    }
}

template <class PS>
__device__ __forceinline__ void eval_one_instance_ps(const CplbParams& P, const PS& ps, const double* x, double* g, double* jac, double* cost,
                                                     double* grad, unsigned flags)
{
    switch (P.env) {
    case CPLB_ENV_NONE_K: eval_one_instance_env<CPLB_ENV_NONE_K>(P, ps, x, g, jac, cost, grad, flags); break;
    case CPLB_ENV_GROUND_K: eval_one_instance_env<CPLB_ENV_GROUND_K>(P, ps, x, g, jac, cost, grad, flags); break;
    default: eval_one_instance_env<CPLB_ENV_SUPERQUADRIC_K>(P, ps, x, g, jac, cost, grad, flags); break;
    }
}

// Q == nullptr: the problem's shared parameters; else instance `inst` of the per-instance arrays (instance-major, as
// cplb_instance_params lays them out), read the way the batched kernels read them (InstanceParams<false>)
__device__ __noinline__ void eval_one_instance(const CplbParams& P, const CplbInstParams* Q, long long inst, const double* x, double* g, double* jac,
                                               double* cost, double* grad, unsigned flags)
{
    if (Q == nullptr) {
        eval_one_instance_ps(P, SharedParams{P}, x, g, jac, cost, grad, flags);
    } else {
        eval_one_instance_ps(P, InstanceParams<false>{P, *Q, inst, 0}, x, g, jac, cost, grad, flags);
    }
}

}  // namespace cplb
#endif
