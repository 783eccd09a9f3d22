// cplb_layout.hpp -- host-side layout generator: the row/column/triplet contract that
// cpl::solver::CplProblem::CplProblem fixes (src/CplProblem.cpp:6-82) and ifopt turns into the
// (iRow, jCol) list IPOPT sees.  Closed-form in (names, env); runs once per problem.
#ifndef CPLB_LAYOUT_HPP
#define CPLB_LAYOUT_HPP

#include <algorithm>
#include <cstdint>
#include <numeric>
#include <string>
#include <vector>

namespace cplb {

struct Layout {
    int nc = 0;
    bool has_env = false;
    int n = 0, m = 0, nnz = 0;
    std::vector<int32_t> perm;      // sorted-name rank -> index in the caller's vector
    std::vector<int32_t> rank;      // index in the caller's vector -> sorted-name rank
    std::vector<int32_t> iRow, jCol;
    std::vector<uint8_t> is_const;  // slot value independent of x (ground == true adds Ground's constants)
    std::vector<double> const_value;
    struct Run { int begin, end; };  // [begin, end) runs of x-dependent slots
    std::vector<Run> var_runs;
    std::vector<int32_t> packed_to_slot;  // the x-dependent slots in slot order: element q of a PACKED Jacobian slice is slot packed_to_slot[q]
    // Where the value of a slot comes from: 0 constant (const_value), 1 copy x[source], 2 negated copy -x[source], 3 computed
    // (source = its index in a COMPUTED Jacobian slice, which holds the computed slots in slot order: computed_to_slot).
    enum SlotKind { kConstant = 0, kCopy = 1, kNegatedCopy = 2, kComputed = 3 };
    std::vector<int32_t> slot_kind, slot_source, computed_to_slot;

    static int col_com() { return 0; }
    static int col_F(int k) { return 3 + 9 * k; }  // AddVariableSet order: F_, p_, n_ per name (CplProblem.cpp:31-33)
    static int col_p(int k) { return 3 + 9 * k + 3; }
    static int col_n(int k) { return 3 + 9 * k + 6; }
    int rows_per_contact() const { return has_env ? 6 : 2; }
    int contact_row(int sorted_rank) const { return 6 + rows_per_contact() * sorted_rank; }

    void build(const std::vector<std::string>& names, bool env_present, bool ground = false)
    {
        nc = (int)names.size();
        has_env = env_present;
        n = 3 + 9 * nc;
        m = 6 + rows_per_contact() * nc;
        perm.resize(nc);
        std::iota(perm.begin(), perm.end(), 0);
        // iteration order of std::map<std::string, ContactVars> (CplProblem.cpp:29,42)
        std::sort(perm.begin(), perm.end(), [&](int a, int b) { return names[a] < names[b]; });
        rank.assign(nc, 0);
        for (int j = 0; j < nc; j++) rank[perm[j]] = j;

        iRow.clear();
        jCol.clear();
        is_const.clear();
        const_value.clear();
        slot_kind.clear();
        slot_source.clear();
        auto put = [&](int r, int c, bool constant = false, double value = 0.0) {
            iRow.push_back(r);
            jCol.push_back(c);
            is_const.push_back(constant ? 1 : 0);
            const_value.push_back(value);
            slot_kind.push_back(constant ? kConstant : kComputed);
            slot_source.push_back(-1);
        };
        auto put_copy = [&](int r, int c, bool negated, int source_col) {  // value is exactly +-x[source_col]
            put(r, c);
            slot_kind.back() = negated ? kNegatedCopy : kCopy;
            slot_source.back() = source_col;
        };
        // CentroidalStatics rows 0..2: identity on every F block (CentroidalStatics.cpp:93-95)
        for (int r = 0; r < 3; r++)
            for (int k = 0; k < nc; k++) put(r, col_F(k) + r, true, 1.0);
        // rows 3..5: the two columns != q of CoM, F_k, p_k (:96-101, :108-113, :128-133)
        for (int q = 0; q < 3; q++) {
            const int c0 = (q == 0) ? 1 : 0, c1 = (q == 2) ? 1 : 2;
            put(3 + q, col_com() + c0);
            put(3 + q, col_com() + c1);
            for (int k = 0; k < nc; k++) {
                put(3 + q, col_F(k) + c0);
                put(3 + q, col_F(k) + c1);
                // -skew(F_k) (:108-113): row 3 <- (F.z, -F.y), row 4 <- (-F.z, F.x), row 5 <- (F.y, -F.x)
                put_copy(3 + q, col_p(k) + c0, q == 1, col_F(k) + (q == 2 ? 1 : 2));
                put_copy(3 + q, col_p(k) + c1, q != 1, col_F(k) + (q == 0 ? 1 : 0));
            }
        }
        for (int j = 0; j < nc; j++) {
            const int k = perm[j];
            int row = contact_row(j);
            if (has_env) {
                for (int c = 0; c < 3; c++) put(row, col_p(k) + c, ground, c == 2 ? 1.0 : 0.0);  // EnvironmentConstraint.cpp:56-58; Ground.cpp:33-34
                row++;
                for (int i = 0; i < 3; i++) {                        // EnvironmentNormal.cpp:66-68, :75-83
                    for (int c = 0; c < 3; c++) put(row + i, col_p(k) + c, ground, 0.0);  // Ground.cpp:49
                    put(row + i, col_n(k) + i, true, 1.0);
                }
                row += 3;
            }
            for (int c = 0; c < 3; c++) put_copy(row, col_F(k) + c, true, col_n(k) + c);  // FrictionCone.cpp:82-84: -n
            for (int c = 0; c < 3; c++) put_copy(row, col_n(k) + c, true, col_F(k) + c);  // :93-95: -F
            for (int c = 0; c < 3; c++) put(row + 1, col_F(k) + c);                       // :85-87
            for (int c = 0; c < 3; c++) put(row + 1, col_n(k) + c);                       // :97-99
        }
        nnz = (int)iRow.size();
        packed_to_slot.clear();
        for (int s = 0; s < nnz; s++)
            if (!is_const[s]) packed_to_slot.push_back(s);
        computed_to_slot.clear();
        for (int s = 0; s < nnz; s++)
            if (slot_kind[s] == kComputed) {
                slot_source[s] = (int32_t)computed_to_slot.size();
                computed_to_slot.push_back(s);
            }
        var_runs.clear();
        for (int s = 0; s < nnz;) {
            if (is_const[s]) {
                s++;
                continue;
            }
            int e = s;
            while (e < nnz && !is_const[e]) e++;
            var_runs.push_back(Run{s, e});
            s = e;
        }
    }
};

}  // namespace cplb
#endif
