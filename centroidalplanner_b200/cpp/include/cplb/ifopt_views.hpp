// cplb/ifopt_views.hpp -- the IFOPT surface (VariableSet / ConstraintSet / CostTerm: GetValues, GetBounds,
// FillJacobianBlock, SetVariables, GetCost) on top of the batched evaluator.
//
// Each class below has the name, constructor role, row count, component name string and sparsity behaviour of its
// namesake under include/CentroidalPlanner/Ifopt/, but is a VIEW of instance i of an InstanceBatch: SetVariables
// writes the instance's slice of the batched x buffer, GetValues / FillJacobianBlock / GetCost read the
// instance's slices of the batched outputs, which one cplb_eval_host call produces for every instance whose x
// changed.  cplb::solver::CplProblem assembles them exactly like the reference's constructor
// (src/CplProblem.cpp:6-82), so ifopt's IpoptAdapter sees the same n, m, bounds, (iRow, jCol) and values.
//
// Intended use is lock-step: N solver threads each own one CplProblem view; all call SetVariables(x_i), then the
// first read evaluates the whole dirty range in ONE kernel launch and every thread consumes its slice.
// Needs <ifopt/...> and Eigen (the reference's own dependencies); in this repository's tests they are the
// stand-ins under oracle/refshim, passed with -I by the test only.
#ifndef CPLB_IFOPT_VIEWS_HPP
#define CPLB_IFOPT_VIEWS_HPP

#include <ifopt/constraint_set.h>
#include <ifopt/cost_term.h>
#include <ifopt/problem.h>
#include <ifopt/variable_set.h>

#include <condition_variable>
#include <cstring>
#include <exception>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "cplb/batched_problem.hpp"

namespace cplb {
namespace solver {

// Host buffers of N instances (instance-major) plus the lazily refreshed outputs.  The Jacobian buffer holds COMPUTED slices
// (CPLB_JAC_COMPUTED): only the slots that take arithmetic travel device -> host; a view's FillJacobianBlock reads those through
// the slot-source map, takes the x-independent ones (the 1.0 identities, a Ground's zeros: CentroidalStatics.cpp:93-95,
// EnvironmentNormal.cpp:66-68, Ground.cpp:33-34,49) from cplb_get_jacobian_constants and the plain copies (+-F in the moment rows,
// -n / -F in FrictionCone's first row: CentroidalStatics.cpp:108-113, FrictionCone.cpp:82-84,93-95) from the instance's own x,
// which the batch holds -- the same bits the reference writes, a negation being exact.  The problem may be sharded over
// several GPUs (BatchedProblem's device-list constructor): the batch is then evaluated by all of them at once.
class InstanceBatch {
public:
    typedef std::shared_ptr<InstanceBatch> Ptr;
    struct Entry {
        int r, c, slot;
        int kind;         // CPLB_SLOT_*
        int source;       // COMPUTED: index into the instance's computed Jacobian slice; (NEGATED_)COPY: column of x
        double constant;  // CONSTANT: the value
        double value(const double* computed, const double* x) const
        {
            switch (kind) {
            case CPLB_SLOT_COMPUTED: return computed[source];
            case CPLB_SLOT_COPY: return x[source];
            case CPLB_SLOT_NEGATED_COPY: return -x[source];
            default: return constant;
            }
        }
    };

    InstanceBatch(BatchedProblem::Ptr problem, int64_t num_instances)
        : _prob(std::move(problem)), _N(num_instances), _n(_prob->GetNumberOfOptimizationVariables()),
          _m(_prob->GetNumberOfConstraints()), _nnz(_prob->GetNumberOfJacobianNonzeros())
    {
        _nv = _prob->GetJacobianSlotSources(_kind, _source);
        std::vector<uint8_t> is_const;
        _prob->GetJacobianConstants(is_const, _const_value);
        _x = pinned((size_t)_N * _n);
        _g = pinned((size_t)_N * _m);
        _jac = pinned((size_t)_N * _nv);
        _grad = pinned((size_t)_N * _n);
        _cost = pinned((size_t)_N);
        std::memset(_x, 0, sizeof(double) * _N * _n);  // Variable3D starts at 0 (src/Variable3D.cpp:8-10)
        _prob->GetJacobianStructure(_iRow, _jCol);
        _dirty.assign((size_t)_N, 1);
        _dirty_lo = 0;
        _dirty_hi = _N;
        const auto& names = _prob->contact_names();
        _var_index["CoM"] = 0;
        for (size_t k = 0; k < names.size(); k++) {
            _var_index["F_" + names[k]] = 1 + 3 * (int)k;
            _var_index["p_" + names[k]] = 2 + 3 * (int)k;
            _var_index["n_" + names[k]] = 3 + 3 * (int)k;
        }
    }
    ~InstanceBatch()
    {
        for (double* p : {_x, _g, _jac, _grad, _cost}) cplb_host_free(p);
    }
    InstanceBatch(const InstanceBatch&) = delete;
    InstanceBatch& operator=(const InstanceBatch&) = delete;

    const BatchedProblem::Ptr& problem() const { return _prob; }
    int64_t size() const { return _N; }
    int var_index(const std::string& var_set) const
    {
        auto it = _var_index.find(var_set);
        return it == _var_index.end() ? -1 : it->second;
    }

    // writes 3 values of instance i at column col; the instance becomes dirty only if a bit changed
    void SetBlock(int64_t i, int col, const double* v)
    {
        double* dst = _x + i * _n + col;
        if (std::memcmp(dst, v, 3 * sizeof(double)) == 0) return;
        std::lock_guard<std::mutex> lk(_mu);
        std::memcpy(dst, v, 3 * sizeof(double));
        _dirty[(size_t)i] = 1;
        if (_dirty_lo >= _dirty_hi) {
            _dirty_lo = i;
            _dirty_hi = i + 1;
        } else {
            if (i < _dirty_lo) _dirty_lo = i;
            if (i + 1 > _dirty_hi) _dirty_hi = i + 1;
        }
    }
    const double* x(int64_t i) const { return _x + i * _n; }
    const double* g(int64_t i)
    {
        Refresh(i);
        return _g + i * _m;
    }
    // instance i's COMPUTED Jacobian slice (jac_computed_size() doubles); Entry::source indexes it for the computed slots
    const double* jac(int64_t i)
    {
        Refresh(i);
        return _jac + i * _nv;
    }
    int jac_computed_size() const { return _nv; }
    // one structural slot of instance i's Jacobian, constants and copies included (what IpoptAdapter::eval_jac_g's values[slot] holds)
    double jac_value(int64_t i, int slot)
    {
        const Entry e{0, 0, slot, _kind[(size_t)slot], _source[(size_t)slot], _const_value[(size_t)slot]};
        return e.value(jac(i), x(i));
    }
    const double* grad(int64_t i)
    {
        Refresh(i);
        return _grad + i * _n;
    }
    double cost(int64_t i)
    {
        Refresh(i);
        return _cost[i];
    }
    int64_t evaluations() const { return _evaluations; }

    // Lock-step mode for `participants` solver threads (one IPOPT per thread, each owning one instance at a time):
    // a thread that reads an output of an instance whose x changed waits until every participant has done the same
    // (or has called Leave()), and the last one to arrive evaluates the whole dirty range in ONE launch.  IPOPT's
    // callbacks (eval_f / eval_grad_f / eval_g / eval_jac_g) then cost one kernel launch per round, not one per thread.
    void EnableLockStep(int participants)
    {
        std::lock_guard<std::mutex> lk(_mu);
        _participants = participants;
        _arrived = 0;
    }
    // a solver thread that has finished (converged / failed) stops taking part in the rounds
    void Leave()
    {
        std::unique_lock<std::mutex> lk(_mu);
        if (_participants > 0) _participants--;
        if (_arrived > 0 && _arrived >= _participants) Release();
    }

    // structural entries of the (rows [row0,row0+rows) x variable set v) block, in ifopt order
    const std::vector<Entry>& Block(int row0, int rows, int v)
    {
        std::lock_guard<std::mutex> lk(_mu);
        const long long key = ((long long)row0 << 32) | (unsigned)v;
        auto it = _blocks.find(key);
        if (it != _blocks.end()) return it->second;
        std::vector<Entry> e;
        for (int s = 0; s < _nnz; s++)
            if (_iRow[s] >= row0 && _iRow[s] < row0 + rows && _jCol[s] >= 3 * v && _jCol[s] < 3 * v + 3)
                e.push_back(Entry{_iRow[s] - row0, _jCol[s] - 3 * v, s, _kind[(size_t)s], _source[(size_t)s], _const_value[(size_t)s]});
        return _blocks.emplace(key, std::move(e)).first->second;
    }

private:
    static double* pinned(size_t count)
    {
        void* p = nullptr;
        check(cplb_host_alloc(count * sizeof(double), &p));
        return static_cast<double*>(p);
    }
    void EvaluateDirtyRange()  // _mu held
    {
        if (_dirty_lo >= _dirty_hi) return;
        const int64_t lo = _dirty_lo, cnt = _dirty_hi - _dirty_lo;
        _prob->EvaluateHost(cnt, _x + lo * _n, _g + lo * _m, _jac + lo * _nv, _cost + lo, _grad + lo * _n, nullptr, CPLB_JAC_COMPUTED);
        std::fill(_dirty.begin() + lo, _dirty.begin() + lo + cnt, 0);
        _dirty_lo = _dirty_hi = 0;
        _evaluations++;
    }
    // _mu held: the round is complete.  A failing evaluation (CUDA error, out of memory) must not strand the other solver
    // threads in their wait: the round is released whatever happens, the exception is kept with the round's generation and
    // rethrown in EVERY participant (the evaluating thread included).
    void Release()
    {
        std::exception_ptr failure;
        try {
            EvaluateDirtyRange();
        } catch (...) {
            failure = std::current_exception();
        }
        _arrived = 0;
        _generation++;
        _failure = failure;
        _failure_generation = _generation;
        _cv.notify_all();
        if (failure) std::rethrow_exception(failure);
    }
    void Refresh(int64_t i)
    {
        std::unique_lock<std::mutex> lk(_mu);
        if (!_dirty[(size_t)i]) return;
        if (_participants <= 1) {
            EvaluateDirtyRange();
            return;
        }
        _arrived++;
        if (_arrived >= _participants) {
            Release();
        } else {
            const int64_t gen = _generation;
            _cv.wait(lk, [&] { return _generation != gen; });
            if (_failure && _failure_generation == gen + 1) std::rethrow_exception(_failure);
        }
    }

    BatchedProblem::Ptr _prob;
    int64_t _N;
    int _n, _m, _nnz, _nv = 0;
    std::vector<int32_t> _kind, _source;
    std::vector<double> _const_value;
    std::exception_ptr _failure;
    int64_t _failure_generation = -1;
    double *_x = nullptr, *_g = nullptr, *_jac = nullptr, *_grad = nullptr, *_cost = nullptr;
    std::vector<int32_t> _iRow, _jCol;
    std::map<std::string, int> _var_index;
    std::map<long long, std::vector<Entry>> _blocks;
    std::mutex _mu;
    std::condition_variable _cv;
    std::vector<char> _dirty;
    int _participants = 0, _arrived = 0;
    int64_t _generation = 0;
    int64_t _dirty_lo = 0, _dirty_hi = 0, _evaluations = 0;
};

// cpl::solver::Variable3D (include/CentroidalPlanner/Ifopt/Variable3D.h:12-34, src/Variable3D.cpp)
class Variable3D : public ifopt::VariableSet {
public:
    typedef std::shared_ptr<Variable3D> Ptr;
    Variable3D(const std::string& var_name, InstanceBatch::Ptr batch, int64_t instance, int column)
        : VariableSet(3, var_name), _batch(std::move(batch)), _i(instance), _col(column)
    {
        for (int q = 0; q < 3; q++) {
            _lb[q] = -1000.0;
            _ub[q] = 1000.0;
        }
    }
    void SetVariables(const VectorXd& x) override
    {
        const double v[3] = {x(0), x(1), x(2)};
        _batch->SetBlock(_i, _col, v);
    }
    void SetBounds(const Eigen::Vector3d& lower, const Eigen::Vector3d& upper)
    {
        bool bad = false;
        for (int q = 0; q < 3; q++) {
            _lb[q] = lower(q);
            _ub[q] = upper(q);
            if (_ub[q] - _lb[q] < 0) bad = true;
        }
        if (bad) throw std::invalid_argument("Inconsistent bounds");
    }
    VecBound GetBounds() const override
    {
        VecBound bounds(GetRows());
        for (int q = 0; q < 3; q++) bounds.at(q) = ifopt::Bounds(_lb[q], _ub[q]);
        return bounds;
    }
    VectorXd GetValues() const override
    {
        VectorXd v(3);
        const double* x = _batch->x(_i) + _col;
        for (int q = 0; q < 3; q++) v(q) = x[q];
        return v;
    }

private:
    InstanceBatch::Ptr _batch;
    int64_t _i;
    int _col;
    double _lb[3], _ub[3];
};

// common part of the four constraint-set views: a run of rows of instance i
class ConstraintRowsView : public ifopt::ConstraintSet {
public:
    ConstraintRowsView(int rows, const std::string& name, InstanceBatch::Ptr batch, int64_t instance, int row0, bool smaller_zero)
        : ConstraintSet(rows, name), _batch(std::move(batch)), _i(instance), _row0(row0), _smaller_zero(smaller_zero)
    {
    }

private:
    VectorXd GetValues() const override
    {
        VectorXd v(GetRows());
        const double* g = _batch->g(_i) + _row0;
        for (int r = 0; r < GetRows(); r++) v(r) = g[r];
        return v;
    }
    VecBound GetBounds() const override
    {
        VecBound b(GetRows());
        for (int r = 0; r < GetRows(); r++) b.at(r) = _smaller_zero ? ifopt::BoundSmallerZero : ifopt::Bounds(.0, .0);
        return b;
    }
    void FillJacobianBlock(std::string var_set, Jacobian& jac_block) const override
    {
        jac_block.setZero();
        const int v = _batch->var_index(var_set);
        if (v < 0) return;
        const double* vals = _batch->jac(_i);  // the computed slice (refreshes the instance if it is dirty)
        const double* xi = _batch->x(_i);
        for (const auto& e : _batch->Block(_row0, GetRows(), v)) jac_block.coeffRef(e.r, e.c) = e.value(vals, xi);
    }
    InstanceBatch::Ptr _batch;
    int64_t _i;
    int _row0;
    bool _smaller_zero;
};

// names and row counts of the reference's sets (CentroidalStatics.cpp:7, FrictionCone.cpp:9,
// EnvironmentConstraint.cpp:8, EnvironmentNormal.cpp:8)
struct CentroidalStatics : ConstraintRowsView {
    typedef std::shared_ptr<CentroidalStatics> Ptr;
    CentroidalStatics(InstanceBatch::Ptr b, int64_t i) : ConstraintRowsView(6, "Centroidal statics constraint", std::move(b), i, 0, false) {}
};
struct EnvironmentConstraint : ConstraintRowsView {
    typedef std::shared_ptr<EnvironmentConstraint> Ptr;
    EnvironmentConstraint(const std::string& contact, InstanceBatch::Ptr b, int64_t i, int row0)
        : ConstraintRowsView(1, "Environment constraint: " + contact, std::move(b), i, row0, false) {}
};
struct EnvironmentNormal : ConstraintRowsView {
    typedef std::shared_ptr<EnvironmentNormal> Ptr;
    EnvironmentNormal(const std::string& contact, InstanceBatch::Ptr b, int64_t i, int row0)
        : ConstraintRowsView(3, "Environment normal: " + contact, std::move(b), i, row0, false) {}
};
struct FrictionCone : ConstraintRowsView {
    typedef std::shared_ptr<FrictionCone> Ptr;
    FrictionCone(const std::string& contact, InstanceBatch::Ptr b, int64_t i, int row0)
        : ConstraintRowsView(2, "Friction cone: " + contact, std::move(b), i, row0, true) {}
};

// cpl::solver::MinimizeCentroidalVariables (MinimizeCentroidalVariables.h:13-68): cost value and gradient blocks.
class MinimizeCentroidalVariables : public ifopt::CostTerm {
public:
    typedef std::shared_ptr<MinimizeCentroidalVariables> Ptr;
    MinimizeCentroidalVariables(InstanceBatch::Ptr batch, int64_t instance)
        : CostTerm("Minimize centroidal variables"), _batch(std::move(batch)), _i(instance) {}

private:
    double GetCost() const override { return _batch->cost(_i); }
    void FillJacobianBlock(std::string var_set, Jacobian& jac) const override
    {
        jac.setZero();
        const int v = _batch->var_index(var_set);
        if (v < 0) return;
        if (v != 0 && (v - 1) % 3 == 2) return;  // n_ blocks are never written (MinimizeCentroidalVariables.cpp:151-193)
        const double* g = _batch->grad(_i) + 3 * v;
        for (int q = 0; q < 3; q++) jac.coeffRef(0, q) = g[q];
    }
    InstanceBatch::Ptr _batch;
    int64_t _i;
};

// The per-instance ifopt::Problem, assembled like cpl::solver::CplProblem::CplProblem (src/CplProblem.cpp:17-80).
class CplProblem : public ifopt::Problem {
public:
    typedef std::shared_ptr<CplProblem> Ptr;
    CplProblem(InstanceBatch::Ptr batch, int64_t instance) : _batch(std::move(batch)), _i(instance)
    {
        const BatchedProblem::Ptr& bp = _batch->problem();
        const auto& names = bp->contact_names();
        _com_var = std::make_shared<Variable3D>("CoM", _batch, _i, 0);
        AddVariableSet(_com_var);
        for (size_t k = 0; k < names.size(); k++) {  // VECTOR order
            const int c = 3 + 9 * (int)k;
            auto F = std::make_shared<Variable3D>("F_" + names[k], _batch, _i, c);
            auto p = std::make_shared<Variable3D>("p_" + names[k], _batch, _i, c + 3);
            auto n = std::make_shared<Variable3D>("n_" + names[k], _batch, _i, c + 6);
            _vars[names[k]] = {F, p, n};
            AddVariableSet(F);
            AddVariableSet(p);
            AddVariableSet(n);
        }
        AddConstraintSet(std::make_shared<CentroidalStatics>(_batch, _i));
        for (auto& elem : _vars) {  // std::map: SORTED order
            int row = bp->GetContactRow(elem.first);
            if (bp->has_environment()) {
                AddConstraintSet(std::make_shared<EnvironmentConstraint>(elem.first, _batch, _i, row));
                AddConstraintSet(std::make_shared<EnvironmentNormal>(elem.first, _batch, _i, row + 1));
                row += 4;
            }
            AddConstraintSet(std::make_shared<FrictionCone>(elem.first, _batch, _i, row));
        }
        AddCostSet(std::make_shared<MinimizeCentroidalVariables>(_batch, _i));
    }
    // per-instance variable bounds, like CplProblem::Set*Bounds (src/CplProblem.cpp:109-172)
    void SetForceBounds(const std::string& n, const Eigen::Vector3d& lb, const Eigen::Vector3d& ub) { _vars.at(n)[0]->SetBounds(lb, ub); }
    void SetPosBounds(const std::string& n, const Eigen::Vector3d& lb, const Eigen::Vector3d& ub) { _vars.at(n)[1]->SetBounds(lb, ub); }
    void SetNormalBounds(const std::string& n, const Eigen::Vector3d& lb, const Eigen::Vector3d& ub) { _vars.at(n)[2]->SetBounds(lb, ub); }

    // CplProblem::GetSolution (src/CplProblem.cpp:85-106): this instance's current variables -- after a solve, what
    // IpoptAdapter::finalize_solution wrote back through SetVariables
    void GetSolution(Solution& sol) const { _batch->problem()->GetSolution(_batch->x(_i), sol); }

private:
    InstanceBatch::Ptr _batch;
    int64_t _i;
    Variable3D::Ptr _com_var;
    std::map<std::string, std::array<Variable3D::Ptr, 3>> _vars;
};

}  // namespace solver
}  // namespace cplb
#endif
