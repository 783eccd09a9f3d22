/*
 * cpl_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C, single-instance CPU restatement of the evaluation path of
 * ADVRHumanoids/CentroidalPlanner (the IFOPT components under src/Constraints,
 * src/Ground.cpp, src/Superquadric.cpp, src/MinimizeCentroidalVariables.cpp and the
 * layout fixed by src/CplProblem.cpp:6-82), plus an emulation of the ifopt
 * assembly (Composite stacking, ConstraintSet::GetJacobian, row-major triplet order).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library, and only as the checker. The product library
 * (centroidalplanner_b200/csrc -> libcplb.so) never links, includes or calls it.
 *
 * PARITY PINNING: the reference's own tests hold no evaluation-level golden vectors
 * (tests/TestBasic.cpp only asserts post-IPOPT invariants). The arithmetic of this
 * oracle is pinned against the reference's OWN SOURCE FILES compiled in place against
 * minimal Eigen/ifopt stand-in headers (oracle/refshim, recipe oracle/Makefile target
 * _ref/libcpl_ref.so). What stays "parity unpinned" are the third-party semantics the
 * stand-ins encode from memory: ifopt's stacking / triplet order and Eigen's 3-term
 * reduction order.
 */
#ifndef CPL_ORACLE_H
#define CPL_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

#define CPL_ORACLE_ENV_NONE 0
#define CPL_ORACLE_ENV_GROUND 1
#define CPL_ORACLE_ENV_SUPERQUADRIC 2

#define CPL_ORACLE_MAX_CONTACTS 64

typedef struct cpl_oracle cpl_oracle;

/* CplProblem::CplProblem (src/CplProblem.cpp:6-82). names are in the caller's vector order. */
cpl_oracle *cpl_oracle_new(int nc, const char *const *names, int env_kind, double mass);
void cpl_oracle_free(cpl_oracle *o);

/* dimensions: n variables, m constraint rows, nnz structural Jacobian entries */
void cpl_oracle_dims(const cpl_oracle *o, int *n, int *m, int *nnz);
/* sorted-name rank j -> index in the caller's vector */
void cpl_oracle_sorted_order(const cpl_oracle *o, int *perm);
/* (iRow, jCol) in the order IpoptAdapter::eval_jac_g(values == NULL) would emit them */
void cpl_oracle_structure(const cpl_oracle *o, int *iRow, int *jCol);
/* variable and constraint bounds (Variable3D::GetBounds, each ConstraintSet::GetBounds) */
void cpl_oracle_var_bounds(const cpl_oracle *o, double *lb, double *ub);
void cpl_oracle_con_bounds(const cpl_oracle *o, double *lb, double *ub);

/* setters (no validation here: the oracle restates arithmetic, the facade validates) */
void cpl_oracle_set_mass(cpl_oracle *o, double m);
void cpl_oracle_set_wrench(cpl_oracle *o, const double w[6]);
void cpl_oracle_set_mu(cpl_oracle *o, double mu);
void cpl_oracle_set_ground_z(cpl_oracle *o, double z);
void cpl_oracle_set_superquadric(cpl_oracle *o, const double C[3], const double R[3], const double P[3]);
/* contact index k is the index in the caller's vector */
void cpl_oracle_set_force_threshold(cpl_oracle *o, int k, double thr);
void cpl_oracle_set_com_ref(cpl_oracle *o, const double r[3]);
void cpl_oracle_set_com_weight(cpl_oracle *o, double w);
void cpl_oracle_set_pos_ref(cpl_oracle *o, int k, const double r[3]);
void cpl_oracle_set_force_ref(cpl_oracle *o, int k, const double r[3]);
void cpl_oracle_set_pos_weight(cpl_oracle *o, int k, double w);
void cpl_oracle_set_force_weight(cpl_oracle *o, int k, double w);
/* block: 0 = CoM (k ignored), 1 = F_k, 2 = p_k, 3 = n_k */
void cpl_oracle_set_var_bounds(cpl_oracle *o, int block, int k, const double lb[3], const double ub[3]);

/* 3-term reduction order (Eigen-version dependent): 0 = (v0+v1)+v2 (default), 1 = v0+(v1+v2) */
void cpl_oracle_set_reduction_order(cpl_oracle *o, int order);
/* 1 (default): FillJacobianBlock is invoked for every (constraint set, variable set) pair like ifopt does.
 * 0: pairs known to produce no entry are skipped (same results; used for a faster, conservative CPU baseline). */
void cpl_oracle_set_call_all_pairs(cpl_oracle *o, int on);

/* One instance. Any output pointer may be NULL. x[n] in column-map order. */
void cpl_oracle_eval(const cpl_oracle *o, const double *x, double *g, double *jac_vals, double *cost,
                     double *grad);

/* N instances, instance-major buffers (x[i*n + c], g[i*m + r], jac[i*nnz + e], cost[i], grad[i*n + c]),
 * instances split contiguously over nthreads pthreads. Returns the number of threads actually used. */
int cpl_oracle_eval_batch(const cpl_oracle *o, long long N, const double *x, double *g, double *jac_vals,
                          double *cost, double *grad, int nthreads);

/* Environment-level entry points (Ground.cpp / Superquadric.cpp restated), for direct unit checks. */
void cpl_oracle_env_value(const cpl_oracle *o, const double p[3], double *value);
void cpl_oracle_env_gradient(const cpl_oracle *o, const double p[3], double grad[3]);
void cpl_oracle_env_normal(const cpl_oracle *o, const double p[3], double normal[3]);
void cpl_oracle_env_normal_jacobian(const cpl_oracle *o, const double p[3], double jac_rowmajor[9]);

#ifdef __cplusplus
}
#endif
#endif
