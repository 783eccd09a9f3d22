// Built by tests/test_gpu_native.py: the IFOPT-surface views (cplb/ifopt_views.hpp) driven the way IpoptAdapter
// drives ifopt::Problem, compared bit for bit with the CPU oracle, for Ground / no-environment problems
// (pow-free, so bit-exact) with vector order != sorted order.  ifopt/Eigen here are the stand-ins of oracle/refshim.
#include <cplb/ifopt_views.hpp>

#include <cstdio>
#include <cstring>
#include <random>

extern "C" {
#include "cpl_oracle.h"
}

static int fails = 0;
#define EXPECT(cond, ...)                                   \
    do {                                                    \
        if (!(cond)) {                                      \
            if (fails < 10) { printf("FAIL: " __VA_ARGS__); printf("\n"); } \
            fails++;                                        \
        }                                                   \
    } while (0)

static bool same(double a, double b) { return (a != a && b != b) || std::memcmp(&a, &b, 8) == 0; }

static void run(bool with_env, int N)
{
    std::vector<std::string> names = {"r_foot", "l_foot", "r_hand", "l_hand"};
    cplb::env::Ground::Ptr ground;
    if (with_env) {
        ground = std::make_shared<cplb::env::Ground>();
        ground->SetGroundZ(0.1);
    }
    auto bp = std::make_shared<cplb::BatchedProblem>(names, 100.0, ground);
    bp->SetMu(0.5);
    bp->SetManipulationWrench({{100.0, 0, 0, 0, 0, 100.0}});
    bp->SetCoMWeight(2.0);
    bp->SetForceThreshold("l_hand", 12.5);
    bp->SetPosRef("r_foot", {{0.1, 0.2, 0.3}});

    const char* cn[4] = {"r_foot", "l_foot", "r_hand", "l_hand"};
    cpl_oracle* o = cpl_oracle_new(4, cn, with_env ? CPL_ORACLE_ENV_GROUND : CPL_ORACLE_ENV_NONE, 100.0);
    cpl_oracle_set_ground_z(o, 0.1);
    cpl_oracle_set_mu(o, 0.5);
    const double w[6] = {100.0, 0, 0, 0, 0, 100.0};
    cpl_oracle_set_wrench(o, w);
    cpl_oracle_set_com_weight(o, 2.0);
    cpl_oracle_set_force_threshold(o, 3, 12.5);
    const double pr[3] = {0.1, 0.2, 0.3};
    cpl_oracle_set_pos_ref(o, 0, pr);
    int n, m, nnz;
    cpl_oracle_dims(o, &n, &m, &nnz);

    auto batch = std::make_shared<cplb::solver::InstanceBatch>(bp, N);
    std::vector<cplb::solver::CplProblem::Ptr> probs;
    for (int i = 0; i < N; i++) probs.push_back(std::make_shared<cplb::solver::CplProblem>(batch, i));

    EXPECT(probs[0]->GetNumberOfOptimizationVariables() == n, "n");
    EXPECT(probs[0]->GetNumberOfConstraints() == m, "m");

    std::mt19937_64 rng(7);
    std::uniform_real_distribution<double> U(-1.0, 1.0);
    std::vector<double> X((size_t)N * n);
    for (int i = 0; i < N; i++)
        for (int c = 0; c < n; c++) {
            int blk = c < 3 ? -1 : ((c - 3) % 9) / 3;
            double v = U(rng);
            X[(size_t)i * n + c] = blk == 0 ? 100.0 * v + (c % 3 == 2 ? 200.0 : 0.0) : (blk == 2 ? (c % 3 == 2 ? 1.0 : 0.0) + 0.1 * v : v);
        }
    // x = 0 start: structure as IpoptAdapter::get_nlp_info / eval_jac_g(values == NULL) see it
    {
        auto J = probs[0]->GetJacobianOfConstraints();
        EXPECT((int)J.nonZeros() == nnz, "nnz %d vs %d", (int)J.nonZeros(), nnz);
        std::vector<int> iRow(nnz), jCol(nnz);
        cpl_oracle_structure(o, iRow.data(), jCol.data());
        int e = 0;
        for (int k = 0; k < J.outerSize(); ++k)
            for (ifopt::Component::Jacobian::InnerIterator it(J, k); it; ++it, ++e)
                EXPECT(e < nnz && it.row() == iRow[e] && it.col() == jCol[e], "triplet %d", e);
    }
    // lock-step: every instance receives its x, then everyone reads -> exactly one more batched evaluation
    const long long ev0 = batch->evaluations();
    for (int i = 0; i < N; i++) probs[i]->SetVariables(&X[(size_t)i * n]);
    std::vector<double> g(m), jac(nnz), grad(n), vals(nnz);
    double cost;
    for (int i = 0; i < N; i++) {
        const double* x = &X[(size_t)i * n];
        cpl_oracle_eval(o, x, g.data(), jac.data(), &cost, grad.data());
        auto gv = probs[i]->EvaluateConstraints(x);
        for (int r = 0; r < m; r++) EXPECT(same(gv(r), g[r]), "g[%d][%d] %.17g vs %.17g", i, r, gv(r), g[r]);
        probs[i]->EvalNonzerosOfJacobian(x, vals.data());
        for (int e = 0; e < nnz; e++) EXPECT(same(vals[e], jac[e]), "jac[%d][%d] %.17g vs %.17g", i, e, vals[e], jac[e]);
        EXPECT(same(probs[i]->EvaluateCostFunction(x), cost), "cost[%d]", i);
        auto gr = probs[i]->EvaluateCostFunctionGradient(x);
        for (int c = 0; c < n; c++) EXPECT(same(gr(c), grad[c]), "grad[%d][%d]", i, c);
    }
    EXPECT(batch->evaluations() == ev0 + 1, "lock-step use must cost one batched evaluation, took %lld", batch->evaluations() - ev0);
    // CplProblem::GetSolution (src/CplProblem.cpp:85-106): what finalize_solution left in the variables, keyed by name
    {
        const int i = N / 2;
        const double* x = &X[(size_t)i * n];
        cplb::solver::Solution sol;
        probs[i]->GetSolution(sol);
        EXPECT(sol.contact_values_map.size() == names.size() && sol.contact_values_map.begin()->first == "l_foot", "solution map order");
        for (int c = 0; c < 3; c++) EXPECT(same(sol.com_sol[c], x[c]), "com_sol[%d]", c);
        for (size_t k = 0; k < names.size(); k++) {
            const auto& v = sol.contact_values_map.at(names[k]);
            for (int c = 0; c < 3; c++) {
                EXPECT(same(v.force_value[c], x[3 + 9 * k + c]), "force_value %s[%d]", names[k].c_str(), c);
                EXPECT(same(v.position_value[c], x[6 + 9 * k + c]), "position_value %s[%d]", names[k].c_str(), c);
                EXPECT(same(v.normal_value[c], x[9 + 9 * k + c]), "normal_value %s[%d]", names[k].c_str(), c);
            }
        }
    }
    // bounds as IpoptAdapter::get_bounds_info reads them
    auto bg = probs[0]->GetBoundsOnConstraints();
    std::vector<double> gl(m), gu(m);
    cpl_oracle_con_bounds(o, gl.data(), gu.data());
    for (int r = 0; r < m; r++) EXPECT(bg[r].lower_ == gl[r] && bg[r].upper_ == gu[r], "bound row %d", r);
    auto bx = probs[0]->GetBoundsOnOptimizationVariables();
    EXPECT((int)bx.size() == n && bx[0].lower_ == -1000.0 && bx[n - 1].upper_ == 1000.0, "variable bounds");
    bool threw = false;
    try {
        probs[0]->SetPosBounds("l_foot", Eigen::Vector3d(0, 0, 1), Eigen::Vector3d(1, 1, 0));
    } catch (const std::invalid_argument&) {
        threw = true;
    }
    EXPECT(threw, "Inconsistent bounds must throw std::invalid_argument");
    threw = false;
    try {
        bp->SetForceThreshold("nose", 1.0);
    } catch (const std::out_of_range&) {
        threw = true;
    }
    EXPECT(threw, "unknown contact must throw std::out_of_range");
    cpl_oracle_free(o);
}

int main()
{
    try {
        run(true, 257);
        run(false, 64);
    } catch (const std::exception& e) {
        printf("exception: %s\n", e.what());
        return 2;
    }
    printf("ifopt views: %d failures\n", fails);
    return fails ? 1 : 0;
}
