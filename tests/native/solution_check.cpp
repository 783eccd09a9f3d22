// Built by tests/test_abi_host.py (CPU, no GPU needed): CplProblem::GetSolution (src/CplProblem.cpp:85-106) and the Solution
// stream operator (:319-346) of the C++ facade.  Prints the solution of the x given on the command line so that the
// Python side can compare it with its own rendering.
#include <cplb/batched_problem.hpp>

#include <cstdlib>
#include <iostream>

int main(int argc, char** argv)
{
    std::vector<std::string> names = {"r_foot", "l_foot", "r_hand"};  // vector order != sorted order
    auto ground = std::make_shared<cplb::env::Ground>();
    cplb::BatchedProblem prob(names, 100.0, ground);
    const int n = prob.GetNumberOfOptimizationVariables();
    if (argc != n + 1) return 2;
    std::vector<double> x(n);
    for (int i = 0; i < n; i++) x[i] = std::atof(argv[1 + i]);
    cplb::solver::Solution sol;
    prob.GetSolution(x.data(), sol);
    if (sol.contact_values_map.size() != 3 || sol.contact_values_map.begin()->first != "l_foot") return 3;
    if (sol.contact_values_map.at("r_hand").normal_value[2] != x[3 + 9 * 2 + 8]) return 4;
    std::cout << sol;
    return 0;
}
