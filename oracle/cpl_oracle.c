/*
 * cpl_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (see cpl_oracle.h).
 *
 * CPU restatement, in plain C, of the reference's per-instance evaluation. Every function cites
 * the reference file:line (relative to /root/reference) whose operation order it follows.
 * Compile with -O2 -ffp-contract=off: parity with the reference is about rounding order.
 *
 * Two layers, like the reference:
 *   1. the component functions (GetValues / FillJacobianBlock of each set, the environment
 *      virtuals, the cost) -- one C function per reference function;
 *   2. an emulation of how ifopt assembles them (Composite stacking in insertion order;
 *      ConstraintSet::GetJacobian visiting EVERY variable set; row-major / column-ascending
 *      triplet order; every coeffRef'd slot structural even if its value is 0.0).
 * The Jacobian structure is DISCOVERED by running layer 2 once and recording which slots the
 * component functions touch -- it is not a hand-written table.
 */
#include "cpl_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#define MAXC CPL_ORACLE_MAX_CONTACTS

/* ---- variable sets (columns) ------------------------------------------------------------
 * CplProblem.cpp:17-34: AddVariableSet(CoM) then, for each name in VECTOR order, F_, p_, n_.
 * var-set index v: 0 = CoM; 1+3k = F_k; 2+3k = p_k; 3+3k = n_k. Column offset = 3*v. */
enum { VK_COM = 0, VK_F = 1, VK_P = 2, VK_N = 3 };

/* ---- constraint sets (rows) ------------------------------------------------------------- */
enum { SK_STATICS = 0, SK_ENVC = 1, SK_ENVN = 2, SK_FRICTION = 3 };

typedef struct {
    int kind;
    int contact; /* vector index of the contact (unused for statics) */
    int row0;
    int rows;
} set_desc;

/* a dense stand-in for the rows x 3 sparse block handed to FillJacobianBlock */
typedef struct {
    double v[6][3];
    unsigned char touched[6][3];
} jac_block;

struct cpl_oracle {
    int nc, env_kind, n, m, nnz;
    char *names[MAXC];
    int perm[MAXC]; /* sorted rank -> vector index (std::map iteration order) */
    double mass, g[3], wrench[6];
    double mu, ground_z, C[3], R[3], P[3];
    double F_thr[MAXC];
    double com_ref[3], W_com, p_ref[MAXC][3], F_ref[MAXC][3], W_p[MAXC], W_F[MAXC];
    double *x_lb, *x_ub;
    int reduction_order, call_all_pairs;
    int nsets, nvars;
    set_desc *sets;
    int *pos; /* [set][var][r(6)][c(3)] -> output slot, or -1 */
};

/* ---- helpers ------------------------------------------------------------------------------ */

static inline void block_set_zero(jac_block *b)
{ /* Eigen SparseMatrix::setZero(): drops every stored entry */
    memset(b->touched, 0, sizeof b->touched);
}

static inline double *cref(jac_block *b, int r, int c)
{ /* SparseMatrix::coeffRef: inserts an explicit 0.0 if the slot does not exist yet */
    if (!b->touched[r][c]) {
        b->touched[r][c] = 1;
        b->v[r][c] = 0.0;
    }
    return &b->v[r][c];
}

/* Eigen fixed-size 3-term reduction (dot / squaredNorm / norm). SURVEY Q2. */
static inline double sum3(const cpl_oracle *o, double a, double b, double c)
{
    return o->reduction_order == 0 ? (a + b) + c : a + (b + c);
}

static inline const double *var_com(const double *x) { return x; }
static inline const double *var_F(const double *x, int k) { return x + 3 + 9 * k; }
static inline const double *var_p(const double *x, int k) { return x + 3 + 9 * k + 3; }
static inline const double *var_n(const double *x, int k) { return x + 3 + 9 * k + 6; }

static inline int var_kind(int v) { return v == 0 ? VK_COM : 1 + (v - 1) % 3; }
static inline int var_contact(int v) { return v == 0 ? -1 : (v - 1) / 3; }

/* ---- environments ---------------------------------------------------------------------------- */

/* Ground.cpp:23-27 / Superquadric.cpp:40-49. NOTE: the superquadric version ACCUMULATES into
 * the out-parameter (+=); EnvironmentConstraint::GetValues zeroes it first (:19-24). */
static void env_value(const cpl_oracle *o, const double p[3], double *value)
{
    if (o->env_kind == CPL_ORACLE_ENV_SUPERQUADRIC) {
        for (int i = 0; i < 3; i++) *value += pow((p[i] - o->C[i]) / o->R[i], o->P[i]);
        *value -= 1.0;
    } else {
        *value = p[2] - o->ground_z;
    }
}

/* Ground.cpp:30-35 / Superquadric.cpp:51-57 */
static void env_gradient(const cpl_oracle *o, const double p[3], double jac[3])
{
    if (o->env_kind == CPL_ORACLE_ENV_SUPERQUADRIC) {
        for (int i = 0; i < 3; i++)
            jac[i] = o->P[i] / pow(o->R[i], o->P[i]) * pow(p[i] - o->C[i], o->P[i] - 1);
    } else {
        jac[0] = 0.0;
        jac[1] = 0.0;
        jac[2] = 1.0;
    }
}

/* Ground.cpp:38-43 / Superquadric.cpp:60-69 (the norm is recomputed three times there; one value) */
static void env_normal(const cpl_oracle *o, const double p[3], double nrm[3])
{
    if (o->env_kind == CPL_ORACLE_ENV_SUPERQUADRIC) {
        double jac[3];
        env_gradient(o, p, jac);
        double len = sqrt(sum3(o, jac[0] * jac[0], jac[1] * jac[1], jac[2] * jac[2]));
        nrm[0] = -jac[0] / len;
        nrm[1] = -jac[1] / len;
        nrm[2] = -jac[2] / len;
    } else {
        nrm[0] = 0.0;
        nrm[1] = 0.0;
        nrm[2] = 1.0;
    }
}

/* Ground.cpp:46-50 / Superquadric.cpp:72-210.
 *
 * The superquadric body is machine-generated: nine closed-form entries, each preceded by its own
 * block of temporaries. Restated here as two rules, with the products in the reference's
 * left-to-right order (that order differs between entries and is part of the contract):
 *
 *  diagonal (a,a), u < v the two other axes               (:78-100, :132-154, :186-208)
 *    chain = P_a * R_a^-P_a, then for q = x,y,z in axis order:
 *              q == a : * d_a^P_a * 1/(C_a-p_a)^2     q != a : * R_q^-2P_q * 1/(C_q-p_q)^2
 *            * (P_a - 1) * 1.0 / pow(S, 3/2) * Q
 *    S = (R_u^-2P_u * i_u * P_u^2 * d_u^2P_u + R_v^-2P_v * i_v * P_v^2 * d_v^2P_v)
 *        + P_a^2 * R_a^-2P_a * d_a^2P_a * i_a
 *    Q = C_u^2 W_u + C_v^2 W_v + p_u^2 W_u + p_v^2 W_v - C_u p_u W_u 2 - C_v p_v W_v 2   (expanded
 *        square), W_u = P_v^2 * d_v^2P_v * R_u^2P_u, W_v = P_u^2 * d_u^2P_u * R_v^2P_v,
 *        each product written out factor by factor.
 *
 *  off-diagonal (r,c), o the third axis                   (:102-130, :156-184)
 *    chain = P_r * R_r^-P_r, then the r-factor d_r^(P_r-1) and the c-factors
 *            P_c^2 * d_c^(2P_c-3) * (2P_c-2) in AXIS order (r-factor first iff r < c),
 *            * R_c^-2P_c * 1.0 / pow(S, 3/2) * (-1/2)
 *    S = (P_o^2 * R_o^-2P_o * d_o^(2P_o-2) + P_r^2 * R_r^-2P_r * d_r^(2P_r-2))
 *        + P_c^2 * d_c^(2P_c-2) * R_c^-2P_c
 */
static void env_normal_jacobian(const cpl_oracle *o, const double p[3], double J[3][3])
{
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) J[r][c] = 0.0; /* setZero(3,3) */
    if (o->env_kind != CPL_ORACLE_ENV_SUPERQUADRIC) return;

    const double *C = o->C, *R = o->R, *P = o->P;

    for (int a = 0; a < 3; a++) {
        const int u = (a == 0) ? 1 : 0;
        const int v = (a == 2) ? 1 : 2;
        double inv[3], rm2p[3], pp[3], b[3], r2p[3];
        for (int q = 0; q < 3; q++) {
            double e = C[q] - p[q];
            double twoP = P[q] * 2.0;
            inv[q] = 1.0 / (e * e);
            rm2p[q] = pow(R[q], -twoP);
            pp[q] = P[q] * P[q];
            b[q] = pow(-C[q] + p[q], twoP);
            r2p[q] = pow(R[q], twoP);
        }
        double chain = P[a] * pow(R[a], -P[a]);
        for (int q = 0; q < 3; q++) {
            if (q == a) {
                chain = chain * pow(-C[a] + p[a], P[a]);
                chain = chain * inv[a];
            } else {
                chain = chain * rm2p[q];
                chain = chain * inv[q];
            }
        }
        chain = chain * (P[a] - 1.0);
        chain = chain * 1.0;
        double S = rm2p[u] * inv[u] * pp[u] * b[u] + rm2p[v] * inv[v] * pp[v] * b[v] +
                   pp[a] * pow(R[a], P[a] * -2.0) * pow(-C[a] + p[a], P[a] * 2.0) * inv[a];
        chain = chain / pow(S, 3.0 / 2.0);
        double Q = (C[u] * C[u]) * pp[v] * b[v] * r2p[u] + (C[v] * C[v]) * pp[u] * b[u] * r2p[v] +
                   (p[u] * p[u]) * pp[v] * b[v] * r2p[u] + (p[v] * p[v]) * pp[u] * b[u] * r2p[v] -
                   C[u] * p[u] * pp[v] * b[v] * r2p[u] * 2.0 - C[v] * p[v] * pp[u] * b[u] * r2p[v] * 2.0;
        J[a][a] = chain * Q;
    }

    for (int r = 0; r < 3; r++) {
        for (int c = 0; c < 3; c++) {
            if (r == c) continue;
            const int q = 3 - r - c; /* third axis */
            double t2 = P[c] * 2.0;
            double ppc = P[c] * P[c];
            double dr = -C[r] + p[r];
            double dc = -C[c] + p[c];
            double t6 = t2 - 2.0;
            double t7 = pow(R[c], -t2);
            double chain = P[r] * pow(R[r], -P[r]);
            if (r < c) {
                chain = chain * pow(dr, P[r] - 1.0);
                chain = chain * ppc;
                chain = chain * pow(dc, t2 - 3.0);
                chain = chain * t6;
            } else {
                chain = chain * ppc;
                chain = chain * pow(dc, t2 - 3.0);
                chain = chain * t6;
                chain = chain * pow(dr, P[r] - 1.0);
            }
            chain = chain * t7;
            chain = chain * 1.0;
            double S = (P[q] * P[q]) * pow(R[q], P[q] * -2.0) * pow(-C[q] + p[q], P[q] * 2.0 - 2.0) +
                       (P[r] * P[r]) * pow(R[r], P[r] * -2.0) * pow(dr, P[r] * 2.0 - 2.0) +
                       ppc * pow(dc, t6) * t7;
            J[r][c] = chain / pow(S, 3.0 / 2.0) * (-1.0 / 2.0);
        }
    }
}

/* ---- CentroidalStatics ---------------------------------------------------------------------- */

/* CentroidalStatics.cpp:37-61 (sorted-name order, force and moment per contact,
 * then -= wrench, then += m*g on rows 0..2) */
static void statics_values(const cpl_oracle *o, const double *x, double value[6])
{
    for (int i = 0; i < 6; i++) value[i] = 0.0;
    const double *CoM = var_com(x);
    for (int j = 0; j < o->nc; j++) {
        int k = o->perm[j];
        const double *Fi = var_F(x, k), *pi = var_p(x, k);
        double d[3] = {pi[0] - CoM[0], pi[1] - CoM[1], pi[2] - CoM[2]};
        value[0] += Fi[0];
        value[1] += Fi[1];
        value[2] += Fi[2];
        value[3] += d[1] * Fi[2] - d[2] * Fi[1];
        value[4] += d[2] * Fi[0] - d[0] * Fi[2];
        value[5] += d[0] * Fi[1] - d[1] * Fi[0];
    }
    for (int i = 0; i < 6; i++) value[i] -= o->wrench[i];
    for (int i = 0; i < 3; i++) value[i] += o->mass * o->g[i];
}

/* CentroidalStatics.cpp:75-138 */
static void statics_fill(const cpl_oracle *o, const double *x, int var, jac_block *b)
{
    block_set_zero(b);
    const double *CoM = var_com(x);
    const int vk = var_kind(var), vc = var_contact(var);
    for (int j = 0; j < o->nc; j++) {
        int k = o->perm[j];
        const double *pi = var_p(x, k), *Fi = var_F(x, k);
        if (vk == VK_F && vc == k) { /* var_set == "F_" + name   (:90-103) */
            *cref(b, 0, 0) = 1.0;
            *cref(b, 1, 1) = 1.0;
            *cref(b, 2, 2) = 1.0;
            *cref(b, 3, 1) = -(pi[2] - CoM[2]);
            *cref(b, 3, 2) = pi[1] - CoM[1];
            *cref(b, 4, 0) = pi[2] - CoM[2];
            *cref(b, 4, 2) = -(pi[0] - CoM[0]);
            *cref(b, 5, 0) = -(pi[1] - CoM[1]);
            *cref(b, 5, 1) = pi[0] - CoM[0];
        }
        if (vk == VK_P && vc == k) { /* var_set == "p_" + name   (:105-115) */
            *cref(b, 3, 1) = Fi[2];
            *cref(b, 3, 2) = -Fi[1];
            *cref(b, 4, 0) = -Fi[2];
            *cref(b, 4, 2) = Fi[0];
            *cref(b, 5, 0) = Fi[1];
            *cref(b, 5, 1) = -Fi[0];
        }
    }
    if (vk == VK_COM) { /* :119-136 */
        for (int j = 0; j < o->nc; j++) {
            const double *Fi = var_F(x, o->perm[j]);
            *cref(b, 3, 1) -= Fi[2];
            *cref(b, 3, 2) -= -Fi[1];
            *cref(b, 4, 0) -= -Fi[2];
            *cref(b, 4, 2) -= Fi[0];
            *cref(b, 5, 0) -= Fi[1];
            *cref(b, 5, 1) -= -Fi[0];
        }
    }
}

/* ---- FrictionCone ---------------------------------------------------------------------------- */

/* FrictionCone.cpp:30-45 */
static void friction_values(const cpl_oracle *o, const double *x, int k, double value[2])
{
    const double *F = var_F(x, k), *n = var_n(x, k);
    double Fn = sum3(o, F[0] * n[0], F[1] * n[1], F[2] * n[2]); /* F.dot(n) */
    double nF = sum3(o, n[0] * F[0], n[1] * F[1], n[2] * F[2]); /* n.dot(F) */
    double u0 = F[0] - nF * n[0], u1 = F[1] - nF * n[1], u2 = F[2] - nF * n[2];
    value[0] = -Fn + o->F_thr[k];
    value[1] = sqrt(sum3(o, u0 * u0, u1 * u1, u2 * u2)) - o->mu * Fn;
}

/* FrictionCone.cpp:60-103 (no guard on the tangential norm: 0/0 -> NaN is a legal output) */
static void friction_fill(const cpl_oracle *o, const double *x, int k, int var, jac_block *b)
{
    double mu = o->mu;
    block_set_zero(b);
    const double *F = var_F(x, k), *n = var_n(x, k);
    double t1 = sum3(o, F[0] * n[0], F[1] * n[1], F[2] * n[2]);
    double t2 = F[0] - n[0] * t1;
    double t3 = F[1] - n[1] * t1;
    double t4 = F[2] - n[2] * t1;
    double t5 = F[0] * n[0];
    double t6 = F[1] * n[1];
    double t7 = F[2] * n[2];
    const int vk = var_kind(var), vc = var_contact(var);
    if (vk == VK_F && vc == k) { /* :79-89 */
        *cref(b, 0, 0) = -n[0];
        *cref(b, 0, 1) = -n[1];
        *cref(b, 0, 2) = -n[2];
        *cref(b, 1, 0) = (t2 * (n[0] * n[0] - 1.0) * 2.0 + n[0] * n[1] * t3 * 2.0 + n[0] * n[2] * t4 * 2.0) * 1.0 /
                             sqrt(t2 * t2 + t3 * t3 + t4 * t4) * (-1.0 / 2.0) -
                         mu * n[0];
        *cref(b, 1, 1) = (t3 * (n[1] * n[1] - 1.0) * 2.0 + n[0] * n[1] * t2 * 2.0 + n[1] * n[2] * t4 * 2.0) * 1.0 /
                             sqrt(t2 * t2 + t3 * t3 + t4 * t4) * (-1.0 / 2.0) -
                         mu * n[1];
        *cref(b, 1, 2) = (t4 * (n[2] * n[2] - 1.0) * 2.0 + n[0] * n[2] * t2 * 2.0 + n[1] * n[2] * t3 * 2.0) * 1.0 /
                             sqrt(t2 * t2 + t3 * t3 + t4 * t4) * (-1.0 / 2.0) -
                         mu * n[2];
    }
    if (vk == VK_N && vc == k) { /* :91-101 */
        *cref(b, 0, 0) = -F[0];
        *cref(b, 0, 1) = -F[1];
        *cref(b, 0, 2) = -F[2];
        *cref(b, 1, 0) = (t2 * (t6 + t7 + t5 * 2.0) * 2.0 + F[0] * n[1] * t3 * 2.0 + F[0] * n[2] * t4 * 2.0) * 1.0 /
                             sqrt(t2 * t2 + t3 * t3 + t4 * t4) * (-1.0 / 2.0) -
                         mu * F[0];
        *cref(b, 1, 1) = (t3 * (t5 + t7 + t6 * 2.0) * 2.0 + F[1] * n[0] * t2 * 2.0 + F[1] * n[2] * t4 * 2.0) * 1.0 /
                             sqrt(t2 * t2 + t3 * t3 + t4 * t4) * (-1.0 / 2.0) -
                         mu * F[1];
        *cref(b, 1, 2) = (t4 * (t5 + t6 + t7 * 2.0) * 2.0 + F[2] * n[0] * t2 * 2.0 + F[2] * n[1] * t3 * 2.0) * 1.0 /
                             sqrt(t2 * t2 + t3 * t3 + t4 * t4) * (-1.0 / 2.0) -
                         mu * F[2];
    }
}

/* ---- EnvironmentConstraint ------------------------------------------------------------------- */

/* EnvironmentConstraint.cpp:16-28 */
static void envc_values(const cpl_oracle *o, const double *x, int k, double value[1])
{
    value[0] = 0.0; /* value.setZero(1) */
    env_value(o, var_p(x, k), &value[0]);
}

/* EnvironmentConstraint.cpp:42-62 (the gradient is evaluated for EVERY var set) */
static void envc_fill(const cpl_oracle *o, const double *x, int k, int var, jac_block *b)
{
    block_set_zero(b);
    double jac[3];
    env_gradient(o, var_p(x, k), jac);
    if (var_kind(var) == VK_P && var_contact(var) == k) {
        *cref(b, 0, 0) = jac[0];
        *cref(b, 0, 1) = jac[1];
        *cref(b, 0, 2) = jac[2];
    }
}

/* ---- EnvironmentNormal ----------------------------------------------------------------------- */

/* EnvironmentNormal.cpp:16-33 */
static void envn_values(const cpl_oracle *o, const double *x, int k, double value[3])
{
    double en[3];
    const double *n = var_n(x, k);
    env_normal(o, var_p(x, k), en);
    for (int i = 0; i < 3; i++) value[i] = n[i] - en[i];
}

/* EnvironmentNormal.cpp:52-87 */
static void envn_fill(const cpl_oracle *o, const double *x, int k, int var, jac_block *b)
{
    block_set_zero(b);
    double J[3][3];
    env_normal_jacobian(o, var_p(x, k), J);
    if (var_kind(var) == VK_N && var_contact(var) == k) {
        *cref(b, 0, 0) = 1.0;
        *cref(b, 1, 1) = 1.0;
        *cref(b, 2, 2) = 1.0;
    }
    if (var_kind(var) == VK_P && var_contact(var) == k) {
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 3; c++) *cref(b, r, c) = J[r][c];
    }
}

/* ---- MinimizeCentroidalVariables ------------------------------------------------------------ */

/* MinimizeCentroidalVariables.cpp:124-148 */
static double cost_value(const cpl_oracle *o, const double *x)
{
    double value = 0;
    const double *CoM = var_com(x);
    for (int j = 0; j < o->nc; j++) {
        int k = o->perm[j];
        const double *Fi = var_F(x, k), *pi = var_p(x, k);
        double dp[3], dF[3];
        for (int i = 0; i < 3; i++) {
            dp[i] = pi[i] - o->p_ref[k][i];
            dF[i] = Fi[i] - o->F_ref[k][i];
        }
        value += 0.5 * o->W_p[k] * sum3(o, dp[0] * dp[0], dp[1] * dp[1], dp[2] * dp[2]) +
                 0.5 * o->W_F[k] * sum3(o, dF[0] * dF[0], dF[1] * dF[1], dF[2] * dF[2]);
    }
    double dc[3];
    for (int i = 0; i < 3; i++) dc[i] = CoM[i] - o->com_ref[i];
    value += 0.5 * o->W_com * sum3(o, dc[0] * dc[0], dc[1] * dc[1], dc[2] * dc[2]);
    return value;
}

/* MinimizeCentroidalVariables.cpp:151-193 (1 x 3 block) */
static void cost_fill(const cpl_oracle *o, const double *x, int var, jac_block *b)
{
    block_set_zero(b);
    const double *CoM = var_com(x);
    const int vk = var_kind(var), vc = var_contact(var);
    for (int j = 0; j < o->nc; j++) {
        int k = o->perm[j];
        if (vk == VK_F && vc == k) {
            const double *Fi = var_F(x, k);
            for (int i = 0; i < 3; i++) *cref(b, 0, i) = o->W_F[k] * (Fi[i] - o->F_ref[k][i]);
        }
        if (vk == VK_P && vc == k) {
            const double *pi = var_p(x, k);
            for (int i = 0; i < 3; i++) *cref(b, 0, i) = o->W_p[k] * (pi[i] - o->p_ref[k][i]);
        }
    }
    if (vk == VK_COM)
        for (int i = 0; i < 3; i++) *cref(b, 0, i) = o->W_com * (CoM[i] - o->com_ref[i]);
}

/* ---- ifopt assembly emulation ---------------------------------------------------------------- */

static void set_values(const cpl_oracle *o, const set_desc *s, const double *x, double *out)
{
    switch (s->kind) {
    case SK_STATICS: statics_values(o, x, out); break;
    case SK_ENVC: envc_values(o, x, s->contact, out); break;
    case SK_ENVN: envn_values(o, x, s->contact, out); break;
    default: friction_values(o, x, s->contact, out); break;
    }
}

static void set_fill(const cpl_oracle *o, const set_desc *s, const double *x, int var, jac_block *b)
{
    switch (s->kind) {
    case SK_STATICS: statics_fill(o, x, var, b); break;
    case SK_ENVC: envc_fill(o, x, s->contact, var, b); break;
    case SK_ENVN: envn_fill(o, x, s->contact, var, b); break;
    default: friction_fill(o, x, s->contact, var, b); break;
    }
}

#define POS(o, s, v, r, c) ((o)->pos[((((size_t)(s)) * (o)->nvars + (v)) * 6 + (r)) * 3 + (c)])

/* Runs ConstraintSet::GetJacobian for every set over every variable set and numbers the touched
 * slots row-major / column-ascending (Composite::GetJacobian + IpoptAdapter::eval_jac_g order). */
static void discover_structure(cpl_oracle *o)
{
    double *x = calloc((size_t)o->n, sizeof(double));
    jac_block b;
    size_t cnt = (size_t)o->nsets * o->nvars * 18;
    o->pos = malloc(cnt * sizeof(int));
    for (size_t i = 0; i < cnt; i++) o->pos[i] = -1;
    unsigned char *touched = calloc(cnt, 1);
    for (int s = 0; s < o->nsets; s++)
        for (int v = 0; v < o->nvars; v++) {
            set_fill(o, &o->sets[s], x, v, &b);
            for (int r = 0; r < o->sets[s].rows; r++)
                for (int c = 0; c < 3; c++)
                    if (b.touched[r][c]) touched[(((size_t)s * o->nvars + v) * 6 + r) * 3 + c] = 1;
        }
    int e = 0;
    for (int s = 0; s < o->nsets; s++)
        for (int r = 0; r < o->sets[s].rows; r++)
            for (int v = 0; v < o->nvars; v++)
                for (int c = 0; c < 3; c++)
                    if (touched[(((size_t)s * o->nvars + v) * 6 + r) * 3 + c]) POS(o, s, v, r, c) = e++;
    o->nnz = e;
    free(touched);
    free(x);
}

static int name_less(const char *a, const char *b)
{ /* std::string operator< : lexicographic on unsigned char, shorter prefix first */
    return strcmp(a, b) < 0;
}

cpl_oracle *cpl_oracle_new(int nc, const char *const *names, int env_kind, double mass)
{
    if (nc < 1 || nc > MAXC) return NULL;
    cpl_oracle *o = calloc(1, sizeof *o);
    o->nc = nc;
    o->env_kind = env_kind;
    for (int k = 0; k < nc; k++) {
        o->names[k] = strdup(names[k]);
        o->perm[k] = k;
    }
    /* std::map<std::string, ContactVars> iteration order (CplProblem.cpp:29,42) */
    for (int i = 1; i < nc; i++) {
        int key = o->perm[i], j = i - 1;
        while (j >= 0 && name_less(o->names[key], o->names[o->perm[j]])) {
            o->perm[j + 1] = o->perm[j];
            j--;
        }
        o->perm[j + 1] = key;
    }
    /* defaults: CentroidalStatics.cpp:12-15, FrictionCone.cpp:14, Environment.h:46, Ground.cpp:7,
     * Superquadric.cpp:7-9, MinimizeCentroidalVariables.cpp:11-26, Variable3D.cpp:12-13 */
    o->mass = 100.0;
    o->g[0] = 0.0;
    o->g[1] = 0.0;
    o->g[2] = -9.81;
    o->mu = 1.0;
    o->ground_z = 0.0;
    o->C[0] = 0.0, o->C[1] = 0.0, o->C[2] = 10.0;
    o->R[0] = o->R[1] = o->R[2] = 10.0;
    o->P[0] = o->P[1] = o->P[2] = 10.0;
    o->com_ref[2] = 1.0;
    o->W_com = 1.0;
    for (int k = 0; k < nc; k++) {
        o->W_p[k] = 1.0;
        o->W_F[k] = 1.0;
    }
    o->mass = mass; /* CplProblem.cpp:39 SetMass(_robot_mass) */
    o->n = 3 + 9 * nc;
    o->nvars = 1 + 3 * nc;
    o->x_lb = malloc(sizeof(double) * o->n);
    o->x_ub = malloc(sizeof(double) * o->n);
    for (int i = 0; i < o->n; i++) {
        o->x_lb[i] = -1000.0;
        o->x_ub[i] = 1000.0;
    }
    o->reduction_order = 0;
    o->call_all_pairs = 1;

    /* constraint sets in AddConstraintSet order (CplProblem.cpp:37-75) */
    o->nsets = 1 + (env_kind != CPL_ORACLE_ENV_NONE ? 3 : 1) * nc;
    o->sets = calloc((size_t)o->nsets, sizeof(set_desc));
    int s = 0, row = 0;
    o->sets[s++] = (set_desc){SK_STATICS, -1, row, 6};
    row += 6;
    for (int j = 0; j < nc; j++) {
        int k = o->perm[j];
        if (env_kind != CPL_ORACLE_ENV_NONE) {
            o->sets[s++] = (set_desc){SK_ENVC, k, row, 1};
            row += 1;
            o->sets[s++] = (set_desc){SK_ENVN, k, row, 3};
            row += 3;
        }
        o->sets[s++] = (set_desc){SK_FRICTION, k, row, 2};
        row += 2;
    }
    o->m = row;
    discover_structure(o);
    return o;
}

void cpl_oracle_free(cpl_oracle *o)
{
    if (!o) return;
    for (int k = 0; k < o->nc; k++) free(o->names[k]);
    free(o->x_lb);
    free(o->x_ub);
    free(o->sets);
    free(o->pos);
    free(o);
}

void cpl_oracle_dims(const cpl_oracle *o, int *n, int *m, int *nnz)
{
    if (n) *n = o->n;
    if (m) *m = o->m;
    if (nnz) *nnz = o->nnz;
}

void cpl_oracle_sorted_order(const cpl_oracle *o, int *perm)
{
    for (int j = 0; j < o->nc; j++) perm[j] = o->perm[j];
}

void cpl_oracle_structure(const cpl_oracle *o, int *iRow, int *jCol)
{
    for (int s = 0; s < o->nsets; s++)
        for (int v = 0; v < o->nvars; v++)
            for (int r = 0; r < o->sets[s].rows; r++)
                for (int c = 0; c < 3; c++) {
                    int e = POS(o, s, v, r, c);
                    if (e >= 0) {
                        iRow[e] = o->sets[s].row0 + r;
                        jCol[e] = 3 * v + c;
                    }
                }
}

void cpl_oracle_var_bounds(const cpl_oracle *o, double *lb, double *ub)
{
    memcpy(lb, o->x_lb, sizeof(double) * o->n);
    memcpy(ub, o->x_ub, sizeof(double) * o->n);
}

void cpl_oracle_con_bounds(const cpl_oracle *o, double *lb, double *ub)
{
    for (int s = 0; s < o->nsets; s++)
        for (int r = 0; r < o->sets[s].rows; r++) {
            int row = o->sets[s].row0 + r;
            if (o->sets[s].kind == SK_FRICTION) { /* ifopt::BoundSmallerZero = (-inf, 0), inf = 1e20 */
                lb[row] = -1.0e20;
                ub[row] = 0.0;
            } else { /* ifopt::Bounds(0,0): CentroidalStatics.cpp:69, EnvironmentConstraint.cpp:36, EnvironmentNormal.cpp:41-46 */
                lb[row] = 0.0;
                ub[row] = 0.0;
            }
        }
}

void cpl_oracle_set_mass(cpl_oracle *o, double m) { o->mass = m; }
void cpl_oracle_set_wrench(cpl_oracle *o, const double w[6]) { memcpy(o->wrench, w, 6 * sizeof(double)); }
void cpl_oracle_set_mu(cpl_oracle *o, double mu) { o->mu = mu; }
void cpl_oracle_set_ground_z(cpl_oracle *o, double z) { o->ground_z = z; }
void cpl_oracle_set_superquadric(cpl_oracle *o, const double C[3], const double R[3], const double P[3])
{
    memcpy(o->C, C, 3 * sizeof(double));
    memcpy(o->R, R, 3 * sizeof(double));
    memcpy(o->P, P, 3 * sizeof(double));
}
void cpl_oracle_set_force_threshold(cpl_oracle *o, int k, double thr) { o->F_thr[k] = thr; }
void cpl_oracle_set_com_ref(cpl_oracle *o, const double r[3]) { memcpy(o->com_ref, r, 3 * sizeof(double)); }
void cpl_oracle_set_com_weight(cpl_oracle *o, double w) { o->W_com = w; }
void cpl_oracle_set_pos_ref(cpl_oracle *o, int k, const double r[3]) { memcpy(o->p_ref[k], r, 3 * sizeof(double)); }
void cpl_oracle_set_force_ref(cpl_oracle *o, int k, const double r[3]) { memcpy(o->F_ref[k], r, 3 * sizeof(double)); }
void cpl_oracle_set_pos_weight(cpl_oracle *o, int k, double w) { o->W_p[k] = w; }
void cpl_oracle_set_force_weight(cpl_oracle *o, int k, double w) { o->W_F[k] = w; }
void cpl_oracle_set_var_bounds(cpl_oracle *o, int block, int k, const double lb[3], const double ub[3])
{
    int col = block == 0 ? 0 : 3 + 9 * k + 3 * (block - 1);
    for (int i = 0; i < 3; i++) {
        o->x_lb[col + i] = lb[i];
        o->x_ub[col + i] = ub[i];
    }
}
void cpl_oracle_set_reduction_order(cpl_oracle *o, int order) { o->reduction_order = order; }
void cpl_oracle_set_call_all_pairs(cpl_oracle *o, int on) { o->call_all_pairs = on; }

static int pair_has_entries(const cpl_oracle *o, int s, int v)
{
    for (int r = 0; r < o->sets[s].rows; r++)
        for (int c = 0; c < 3; c++)
            if (POS(o, s, v, r, c) >= 0) return 1;
    return 0;
}

void cpl_oracle_eval(const cpl_oracle *o, const double *x, double *g, double *jac_vals, double *cost,
                     double *grad)
{
    jac_block b;
    if (g) /* Problem::EvaluateConstraints -> Composite::GetValues: concatenation in set order */
        for (int s = 0; s < o->nsets; s++) set_values(o, &o->sets[s], x, g + o->sets[s].row0);
    if (jac_vals) /* Problem::EvalNonzerosOfJacobian */
        for (int s = 0; s < o->nsets; s++)
            for (int v = 0; v < o->nvars; v++) {
                if (!o->call_all_pairs && !pair_has_entries(o, s, v)) continue;
                set_fill(o, &o->sets[s], x, v, &b);
                for (int r = 0; r < o->sets[s].rows; r++)
                    for (int c = 0; c < 3; c++)
                        if (b.touched[r][c]) jac_vals[POS(o, s, v, r, c)] = b.v[r][c];
            }
    if (cost) *cost = cost_value(o, x); /* Problem::EvaluateCostFunction */
    if (grad) { /* Problem::EvaluateCostFunctionGradient: dense row 0 of the cost Jacobian */
        for (int i = 0; i < o->n; i++) grad[i] = 0.0;
        for (int v = 0; v < o->nvars; v++) {
            cost_fill(o, x, v, &b);
            for (int c = 0; c < 3; c++)
                if (b.touched[0][c]) grad[3 * v + c] = b.v[0][c];
        }
    }
}

typedef struct {
    const cpl_oracle *o;
    long long i0, i1;
    const double *x;
    double *g, *jac, *cost, *grad;
} batch_job;

static void *batch_worker(void *arg)
{
    batch_job *j = arg;
    const cpl_oracle *o = j->o;
    for (long long i = j->i0; i < j->i1; i++)
        cpl_oracle_eval(o, j->x + i * o->n, j->g ? j->g + i * o->m : NULL, j->jac ? j->jac + i * o->nnz : NULL,
                        j->cost ? j->cost + i : NULL, j->grad ? j->grad + i * o->n : NULL);
    return NULL;
}

int cpl_oracle_eval_batch(const cpl_oracle *o, long long N, const double *x, double *g, double *jac_vals,
                          double *cost, double *grad, int nthreads)
{
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    if ((long long)nthreads > N) nthreads = N > 0 ? (int)N : 1;
    pthread_t th[256];
    batch_job jobs[256];
    for (int t = 0; t < nthreads; t++) {
        jobs[t] = (batch_job){o, N * t / nthreads, N * (t + 1) / nthreads, x, g, jac_vals, cost, grad};
        if (t > 0) pthread_create(&th[t], NULL, batch_worker, &jobs[t]);
    }
    batch_worker(&jobs[0]);
    for (int t = 1; t < nthreads; t++) pthread_join(th[t], NULL);
    return nthreads;
}

void cpl_oracle_env_value(const cpl_oracle *o, const double p[3], double *value)
{
    *value = 0.0;
    env_value(o, p, value);
}
void cpl_oracle_env_gradient(const cpl_oracle *o, const double p[3], double grad[3]) { env_gradient(o, p, grad); }
void cpl_oracle_env_normal(const cpl_oracle *o, const double p[3], double normal[3]) { env_normal(o, p, normal); }
void cpl_oracle_env_normal_jacobian(const cpl_oracle *o, const double p[3], double jac_rowmajor[9])
{
    double J[3][3];
    env_normal_jacobian(o, p, J);
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) jac_rowmajor[3 * r + c] = J[r][c];
}
