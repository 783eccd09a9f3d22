// oracle/ref_driver.cpp -- TEST INFRASTRUCTURE.  C entry points around the reference's OWN
// cpl::solver::CplProblem (compiled from /root/reference/src against oracle/refshim), driven the way
// ifopt's IpoptAdapter drives it: eval_g -> Problem::EvaluateConstraints, eval_jac_g ->
// Problem::EvalNonzerosOfJacobian / GetJacobianOfConstraints, eval_f / eval_grad_f.
#include <CentroidalPlanner/Environment/Ground.h>
#include <CentroidalPlanner/Environment/Superquadric.h>
#include <CentroidalPlanner/Ifopt/CplProblem.h>

#include <pthread.h>

#include <cstring>
#include <memory>
#include <string>
#include <vector>

using cpl::solver::CplProblem;

struct cpl_ref {
    std::vector<std::string> names;
    int env_kind;
    double mass;
    std::shared_ptr<cpl::env::Ground> ground;
    std::shared_ptr<cpl::env::Superquadric> sq;
    std::shared_ptr<CplProblem> prob;
    // recorded parameters so that per-thread clones can be configured identically
    double wrench[6] = {0, 0, 0, 0, 0, 0};
    double mu = 1.0, ground_z = 0.0;
    double C[3] = {0, 0, 10}, R[3] = {10, 10, 10}, P[3] = {10, 10, 10};
    std::vector<double> F_thr, W_p, W_F;
    std::vector<std::vector<double>> p_ref, F_ref;
    double com_ref[3] = {0, 0, 1}, W_com = 1.0;
};

static Eigen::Vector3d v3(const double* p) { return Eigen::Vector3d(p[0], p[1], p[2]); }

static void configure(cpl_ref* r, std::shared_ptr<cpl::env::Ground>& ground, std::shared_ptr<cpl::env::Superquadric>& sq,
                      std::shared_ptr<CplProblem>& prob)
{
    cpl::env::EnvironmentClass::Ptr env;
    if (r->env_kind == 1) {
        ground = std::make_shared<cpl::env::Ground>();
        ground->SetGroundZ(r->ground_z);
        env = ground;
    } else if (r->env_kind == 2) {
        sq = std::make_shared<cpl::env::Superquadric>();
        sq->SetParameters(v3(r->C), v3(r->R), v3(r->P));
        env = sq;
    }
    prob = std::make_shared<CplProblem>(r->names, r->mass, env);
    prob->SetMu(r->mu);
    Eigen::VectorXd w(6);
    for (int i = 0; i < 6; i++) w[i] = r->wrench[i];
    prob->SetManipulationWrench(w);
    prob->SetCoMRef(v3(r->com_ref));
    prob->SetCoMWeight(r->W_com);
    for (size_t k = 0; k < r->names.size(); k++) {
        prob->SetForceThreshold(r->names[k], r->F_thr[k]);
        prob->SetPosRef(r->names[k], v3(r->p_ref[k].data()));
        prob->SetForceRef(r->names[k], v3(r->F_ref[k].data()));
        prob->SetContactPosWeight(r->names[k], r->W_p[k]);
        prob->SetContactForceWeight(r->names[k], r->W_F[k]);
    }
}

extern "C" {

cpl_ref* cpl_ref_new(int nc, const char* const* names, int env_kind, double mass)
{
    cpl_ref* r = new cpl_ref();
    for (int k = 0; k < nc; k++) r->names.emplace_back(names[k]);
    r->env_kind = env_kind;
    r->mass = mass;
    r->F_thr.assign(nc, 0.0);
    r->W_p.assign(nc, 1.0);
    r->W_F.assign(nc, 1.0);
    r->p_ref.assign(nc, std::vector<double>(3, 0.0));
    r->F_ref.assign(nc, std::vector<double>(3, 0.0));
    configure(r, r->ground, r->sq, r->prob);
    return r;
}
void cpl_ref_free(cpl_ref* r) { delete r; }

static void rebuild(cpl_ref* r) { configure(r, r->ground, r->sq, r->prob); }

void cpl_ref_set_wrench(cpl_ref* r, const double* w) { std::memcpy(r->wrench, w, 48); rebuild(r); }
void cpl_ref_set_mu(cpl_ref* r, double mu) { r->mu = mu; rebuild(r); }
void cpl_ref_set_ground_z(cpl_ref* r, double z) { r->ground_z = z; rebuild(r); }
void cpl_ref_set_superquadric(cpl_ref* r, const double* C, const double* R, const double* P)
{
    std::memcpy(r->C, C, 24);
    std::memcpy(r->R, R, 24);
    std::memcpy(r->P, P, 24);
    rebuild(r);
}
void cpl_ref_set_force_threshold(cpl_ref* r, int k, double t) { r->F_thr[k] = t; rebuild(r); }
void cpl_ref_set_com_ref(cpl_ref* r, const double* v) { std::memcpy(r->com_ref, v, 24); rebuild(r); }
void cpl_ref_set_com_weight(cpl_ref* r, double w) { r->W_com = w; rebuild(r); }
void cpl_ref_set_pos_ref(cpl_ref* r, int k, const double* v) { r->p_ref[k].assign(v, v + 3); rebuild(r); }
void cpl_ref_set_force_ref(cpl_ref* r, int k, const double* v) { r->F_ref[k].assign(v, v + 3); rebuild(r); }
void cpl_ref_set_pos_weight(cpl_ref* r, int k, double w) { r->W_p[k] = w; rebuild(r); }
void cpl_ref_set_force_weight(cpl_ref* r, int k, double w) { r->W_F[k] = w; rebuild(r); }

void cpl_ref_dims(cpl_ref* r, int* n, int* m, int* nnz)
{
    *n = r->prob->GetNumberOfOptimizationVariables();
    *m = r->prob->GetNumberOfConstraints();
    *nnz = (int)r->prob->GetJacobianOfConstraints().nonZeros();  // IpoptAdapter::get_nlp_info, at the start point x = 0
}

// IpoptAdapter::eval_jac_g(values == NULL)
void cpl_ref_structure(cpl_ref* r, int* iRow, int* jCol)
{
    auto jac = r->prob->GetJacobianOfConstraints();
    int nele = 0;
    for (int k = 0; k < jac.outerSize(); ++k)
        for (ifopt::Component::Jacobian::InnerIterator it(jac, k); it; ++it) {
            iRow[nele] = (int)it.row();
            jCol[nele] = (int)it.col();
            nele++;
        }
}

void cpl_ref_bounds(cpl_ref* r, double* xl, double* xu, double* gl, double* gu)
{
    auto bx = r->prob->GetBoundsOnOptimizationVariables();
    for (size_t i = 0; i < bx.size(); i++) {
        xl[i] = bx[i].lower_;
        xu[i] = bx[i].upper_;
    }
    auto bg = r->prob->GetBoundsOnConstraints();
    for (size_t i = 0; i < bg.size(); i++) {
        gl[i] = bg[i].lower_;
        gu[i] = bg[i].upper_;
    }
}

static void eval_one(CplProblem& prob, int n, const double* x, double* g, double* jac, double* cost, double* grad)
{
    if (g) {
        Eigen::VectorXd v = prob.EvaluateConstraints(x);
        for (int i = 0; i < v.size(); i++) g[i] = v[i];
    }
    if (jac) prob.EvalNonzerosOfJacobian(x, jac);
    if (cost) *cost = prob.EvaluateCostFunction(x);
    if (grad) {
        Eigen::VectorXd v = prob.EvaluateCostFunctionGradient(x);
        for (int i = 0; i < n; i++) grad[i] = v[i];
    }
}

void cpl_ref_eval(cpl_ref* r, const double* x, double* g, double* jac, double* cost, double* grad)
{
    eval_one(*r->prob, r->prob->GetNumberOfOptimizationVariables(), x, g, jac, cost, grad);
}

struct job {
    cpl_ref* r;
    long long i0, i1;
    const double* x;
    double *g, *jac, *cost, *grad;
    int n, m, nnz;
};

static void* worker(void* arg)
{
    job* j = static_cast<job*>(arg);
    // one CplProblem per thread: the reference's objects are "mutate then read" (Variable3D::SetVariables)
    std::shared_ptr<cpl::env::Ground> ground;
    std::shared_ptr<cpl::env::Superquadric> sq;
    std::shared_ptr<CplProblem> prob;
    configure(j->r, ground, sq, prob);
    for (long long i = j->i0; i < j->i1; i++)
        eval_one(*prob, j->n, j->x + i * j->n, j->g ? j->g + i * j->m : nullptr, j->jac ? j->jac + i * j->nnz : nullptr,
                 j->cost ? j->cost + i : nullptr, j->grad ? j->grad + i * j->n : nullptr);
    return nullptr;
}

int cpl_ref_eval_batch(cpl_ref* r, long long N, const double* x, double* g, double* jac, double* cost, double* grad, int nthreads)
{
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    if ((long long)nthreads > N) nthreads = N > 0 ? (int)N : 1;
    int n, m, nnz;
    cpl_ref_dims(r, &n, &m, &nnz);
    std::vector<pthread_t> th(nthreads);
    std::vector<job> jobs(nthreads);
    for (int t = 0; t < nthreads; t++) {
        jobs[t] = job{r, N * t / nthreads, N * (t + 1) / nthreads, x, g, jac, cost, grad, n, m, nnz};
        if (t > 0) pthread_create(&th[t], nullptr, worker, &jobs[t]);
    }
    worker(&jobs[0]);
    for (int t = 1; t < nthreads; t++) pthread_join(th[t], nullptr);
    return nthreads;
}

}  // extern "C"
