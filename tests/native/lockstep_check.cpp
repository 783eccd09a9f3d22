// Built by tests/test_gpu_native.py: T solver threads, each owning one instance's ifopt::Problem view, run K rounds of
// the callback sequence IPOPT issues per iteration (eval_f, eval_grad_f, eval_g, eval_jac_g) followed by an x update.
// In lock-step mode the whole batch must cost ONE evaluation per round, and every thread must read exactly what a
// sequential CPU replay (oracle) computes.  Threads leave at different rounds, like solvers converging at different times.
#include <cplb/ifopt_views.hpp>

#include <atomic>
#include <cstdio>
#include <cstring>
#include <thread>

extern "C" {
#include "cpl_oracle.h"
}

static std::atomic<int> fails{0};
static bool same(double a, double b) { return (a != a && b != b) || std::memcmp(&a, &b, 8) == 0; }

int main()
{
    const int T = 96, K = 12;
    std::vector<std::string> names = {"contact1", "contact2", "contact3", "contact4"};
    const char* cn[4] = {"contact1", "contact2", "contact3", "contact4"};
    auto ground = std::make_shared<cplb::env::Ground>();
    ground->SetGroundZ(0.1);
    ground->SetMu(0.5);
    auto bp = std::make_shared<cplb::BatchedProblem>(names, 100.0, ground);
    bp->SetCoMWeight(2.0);
    cpl_oracle* o = cpl_oracle_new(4, cn, CPL_ORACLE_ENV_GROUND, 100.0);
    cpl_oracle_set_ground_z(o, 0.1);
    cpl_oracle_set_mu(o, 0.5);
    cpl_oracle_set_com_weight(o, 2.0);
    int n, m, nnz;
    cpl_oracle_dims(o, &n, &m, &nnz);

    auto batch = std::make_shared<cplb::solver::InstanceBatch>(bp, T);
    std::vector<cplb::solver::CplProblem::Ptr> probs;
    for (int i = 0; i < T; i++) probs.push_back(std::make_shared<cplb::solver::CplProblem>(batch, i));
    batch->EnableLockStep(T);
    const long long ev0 = batch->evaluations();

    auto solver = [&](int i) {
        std::vector<double> x(n), g(m), jac(nnz), grad(n), vals(nnz);
        for (int c = 0; c < n; c++) x[c] = 0.01 * (i + 1) + 0.1 * c;  // distinct start per thread
        const int rounds = K - (i % 4);                               // threads "converge" at different rounds
        for (int k = 0; k < rounds; k++) {
            double cost;
            cpl_oracle_eval(o, x.data(), g.data(), jac.data(), &cost, grad.data());
            const double f = probs[i]->EvaluateCostFunction(x.data());                 // eval_f
            auto gr = probs[i]->EvaluateCostFunctionGradient(x.data());                // eval_grad_f
            auto gv = probs[i]->EvaluateConstraints(x.data());                         // eval_g
            probs[i]->EvalNonzerosOfJacobian(x.data(), vals.data());                   // eval_jac_g
            if (!same(f, cost)) fails++;
            for (int c = 0; c < n; c++) if (!same(gr(c), grad[c])) fails++;
            for (int r = 0; r < m; r++) if (!same(gv(r), g[r])) fails++;
            for (int e = 0; e < nnz; e++) if (!same(vals[e], jac[e])) fails++;
            for (int c = 0; c < n; c++) x[c] -= 1e-3 * grad[c] + 1e-6 * (c + 1);       // the "step"
        }
        batch->Leave();
    };
    std::vector<std::thread> th;
    for (int i = 0; i < T; i++) th.emplace_back(solver, i);
    for (auto& t : th) t.join();
    const long long evals = batch->evaluations() - ev0;
    printf("lock-step: %d threads x up to %d rounds -> %lld batched evaluations, %d mismatches\n", T, K, evals, fails.load());
    cpl_oracle_free(o);
    if (evals != K) { printf("expected exactly %d evaluations\n", K); return 1; }
    return fails.load() ? 1 : 0;
}
