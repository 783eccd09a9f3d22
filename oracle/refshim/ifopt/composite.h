// oracle/refshim/ifopt/composite.h -- TEST INFRASTRUCTURE.  A from-scratch stand-in for the part of the
// ifopt modelling layer that CentroidalPlanner's sources compile against and that IPOPT's adapter would
// call (ifopt itself is absent from this image).  Everything here is stated from memory of ifopt 2.x
// ([M] in SURVEY.md Appendix B) and is exactly the set of third-party semantics the parity claim
// cannot pin:
//   * Composite stacks components in insertion order; values are concatenated (summed for costs);
//   * ConstraintSet::GetJacobian visits EVERY variable set, hands FillJacobianBlock a freshly emptied
//     rows x set_rows block, and places the block's entries at the set's column offset;
//   * Jacobian = row-major sparse matrix; explicit zeros are kept; duplicates are summed;
//   * CostTerm is a 1-row ConstraintSet whose value is GetCost() and whose bound is NoBound;
//   * Problem::Evaluate* call SetVariables(x) first; the cost gradient is the dense row 0 of the cost Jacobian.
#ifndef CPL_REFSHIM_IFOPT_COMPOSITE_H
#define CPL_REFSHIM_IFOPT_COMPOSITE_H

#include <Eigen/Geometry>

#include <cassert>
#include <iostream>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace ifopt {

struct Bounds {
    Bounds(double lower = 0.0, double upper = 0.0) : lower_(lower), upper_(upper) {}
    double lower_;
    double upper_;
};

static const double inf = 1.0e20;
static const Bounds NoBound = Bounds(-inf, +inf);
static const Bounds BoundZero = Bounds(0.0, 0.0);
static const Bounds BoundGreaterZero = Bounds(0.0, +inf);
static const Bounds BoundSmallerZero = Bounds(-inf, 0.0);

class Component {
public:
    using Ptr = std::shared_ptr<Component>;
    using Jacobian = Eigen::SparseMatrix<double, Eigen::RowMajor>;
    using VectorXd = Eigen::VectorXd;
    using VecBound = std::vector<Bounds>;

    static const int kSpecifyLater = -1;

    Component(int num_rows, const std::string& name) : num_rows_(num_rows), name_(name) {}
    virtual ~Component() = default;

    virtual VectorXd GetValues() const = 0;
    virtual VecBound GetBounds() const = 0;
    virtual void SetVariables(const VectorXd& x) = 0;
    virtual Jacobian GetJacobian() const = 0;

    int GetRows() const { return num_rows_; }
    std::string GetName() const { return name_; }
    void SetRows(int num_rows) { num_rows_ = num_rows; }

private:
    int num_rows_;
    std::string name_;
};

class Composite : public Component {
public:
    using Ptr = std::shared_ptr<Composite>;
    using ComponentVec = std::vector<Component::Ptr>;

    Composite(const std::string& name, bool is_cost) : Component(0, name), is_cost_(is_cost) {}

    void AddComponent(const Component::Ptr& c)
    {
        components_.push_back(c);
        if (is_cost_)
            SetRows(1);
        else
            SetRows(GetRows() + c->GetRows());
    }
    void ClearComponents()
    {
        components_.clear();
        SetRows(0);
    }
    const Component::Ptr GetComponent(std::string name) const
    {
        for (const auto& c : components_)
            if (c->GetName() == name) return c;
        assert(false);
        return Component::Ptr();
    }
    template <typename T>
    std::shared_ptr<T> GetComponent(const std::string& name) const
    {
        return std::dynamic_pointer_cast<T>(GetComponent(name));
    }
    const ComponentVec GetComponents() const { return components_; }
    int GetComponentCount() const { return (int)components_.size(); }

    VectorXd GetValues() const override
    {
        VectorXd g_all = VectorXd::Zero(GetRows());
        int row = 0;
        for (const auto& c : components_) {
            int n_rows = c->GetRows();
            VectorXd g = c->GetValues();
            for (int i = 0; i < n_rows; i++) g_all[row + i] += g[i];
            if (!is_cost_) row += n_rows;
        }
        return g_all;
    }
    void SetVariables(const VectorXd& x) override
    {
        int row = 0;
        for (auto& c : components_) {
            int n_rows = c->GetRows();
            c->SetVariables(x.segment(row, n_rows));
            row += n_rows;
        }
    }
    Jacobian GetJacobian() const override
    {
        int n_var = components_.empty() ? 0 : (int)components_.front()->GetJacobian().cols();
        Jacobian jacobian(GetRows(), n_var);
        if (n_var == 0) return jacobian;
        int row = 0;
        std::vector<Eigen::Triplet<double>> triplet_list;
        for (const auto& c : components_) {
            const Jacobian jac = c->GetJacobian();
            for (int k = 0; k < jac.outerSize(); ++k)
                for (Jacobian::InnerIterator it(jac, k); it; ++it)
                    triplet_list.push_back(Eigen::Triplet<double>(row + it.row(), it.col(), it.value()));
            if (!is_cost_) row += c->GetRows();
        }
        jacobian.setFromTriplets(triplet_list.begin(), triplet_list.end());
        return jacobian;
    }
    VecBound GetBounds() const override
    {
        VecBound bounds_;
        for (const auto& c : components_) {
            VecBound b = c->GetBounds();
            bounds_.insert(bounds_.end(), b.begin(), b.end());
        }
        return bounds_;
    }

private:
    ComponentVec components_;
    bool is_cost_;
};

class VariableSet : public Component {
public:
    VariableSet(int n_var, const std::string& name) : Component(n_var, name) {}
    virtual ~VariableSet() = default;
    Jacobian GetJacobian() const final { throw std::runtime_error("not implemented for variables"); }
};

class ConstraintSet : public Component {
public:
    using Ptr = std::shared_ptr<ConstraintSet>;
    using VariablesPtr = Composite::Ptr;

    ConstraintSet(int n_constraints, const std::string& name) : Component(n_constraints, name) {}
    virtual ~ConstraintSet() = default;

    void LinkWithVariables(const VariablesPtr& x)
    {
        variables_ = x;
        InitVariableDependedQuantities(x);
    }
    Jacobian GetJacobian() const final
    {
        Jacobian jacobian(GetRows(), variables_->GetRows());
        int col = 0;
        Jacobian jac;
        std::vector<Eigen::Triplet<double>> triplet_list;
        for (const auto& vars : variables_->GetComponents()) {
            int n = vars->GetRows();
            jac.resize(GetRows(), n);
            FillJacobianBlock(vars->GetName(), jac);
            for (int k = 0; k < jac.outerSize(); ++k)
                for (Jacobian::InnerIterator it(jac, k); it; ++it)
                    triplet_list.push_back(Eigen::Triplet<double>(it.row(), col + it.col(), it.value()));
            col += n;
        }
        jacobian.setFromTriplets(triplet_list.begin(), triplet_list.end());
        return jacobian;
    }
    virtual void FillJacobianBlock(std::string var_set, Jacobian& jac_block) const = 0;

protected:
    const VariablesPtr GetVariables() const { return variables_; }

private:
    VariablesPtr variables_;
    virtual void InitVariableDependedQuantities(const VariablesPtr&) {}
    void SetVariables(const VectorXd&) final { assert(false); }
};

class CostTerm : public ConstraintSet {
public:
    CostTerm(const std::string& name) : ConstraintSet(1, name) {}
    virtual ~CostTerm() = default;

private:
    virtual double GetCost() const = 0;
    VectorXd GetValues() const final
    {
        VectorXd cost(1);
        cost(0) = GetCost();
        return cost;
    }
    VecBound GetBounds() const final { return VecBound(GetRows(), NoBound); }
};

class Problem {
public:
    using VecBound = Component::VecBound;
    using Jacobian = Component::Jacobian;
    using VectorXd = Component::VectorXd;

    Problem() : constraints_("constraint-sets", false), costs_("cost-terms", true) { variables_ = std::make_shared<Composite>("variable-sets", false); }
    virtual ~Problem() = default;

    void AddVariableSet(VariableSet::Ptr variable_set) { variables_->AddComponent(variable_set); }
    void AddConstraintSet(ConstraintSet::Ptr constraint_set)
    {
        constraint_set->LinkWithVariables(variables_);
        constraints_.AddComponent(constraint_set);
    }
    void AddCostSet(ConstraintSet::Ptr cost_set)
    {
        cost_set->LinkWithVariables(variables_);
        costs_.AddComponent(cost_set);
    }
    void SetVariables(const double* x)
    {
        VectorXd v(GetNumberOfOptimizationVariables());
        for (int i = 0; i < v.size(); i++) v[i] = x[i];
        variables_->SetVariables(v);
    }
    int GetNumberOfOptimizationVariables() const { return variables_->GetRows(); }
    bool HasCostTerms() const { return costs_.GetRows() > 0; }
    VecBound GetBoundsOnOptimizationVariables() const { return variables_->GetBounds(); }
    VectorXd GetVariableValues() const { return variables_->GetValues(); }
    double EvaluateCostFunction(const double* x)
    {
        VectorXd g = VectorXd::Zero(1);
        if (HasCostTerms()) {
            SetVariables(x);
            g = costs_.GetValues();
        }
        return g(0);
    }
    VectorXd EvaluateCostFunctionGradient(const double* x)
    {
        Jacobian jac = Jacobian(1, GetNumberOfOptimizationVariables());
        if (HasCostTerms()) {
            SetVariables(x);
            jac = costs_.GetJacobian();
        }
        return jac.row(0).transpose();
    }
    int GetNumberOfConstraints() const { return (int)GetBoundsOnConstraints().size(); }
    VecBound GetBoundsOnConstraints() const { return constraints_.GetBounds(); }
    VectorXd EvaluateConstraints(const double* x)
    {
        SetVariables(x);
        return constraints_.GetValues();
    }
    void EvalNonzerosOfJacobian(const double* x, double* values)
    {
        SetVariables(x);
        Jacobian jac = GetJacobianOfConstraints();
        jac.makeCompressed();
        int nele = 0;  // copy of valuePtr(): row-major, column ascending
        for (int k = 0; k < jac.outerSize(); ++k)
            for (Jacobian::InnerIterator it(jac, k); it; ++it) values[nele++] = it.value();
    }
    Jacobian GetJacobianOfConstraints() const { return constraints_.GetJacobian(); }
    Composite::Ptr GetOptVariables() const { return variables_; }
    const Composite& GetConstraints() const { return constraints_; }
    const Composite& GetCosts() const { return costs_; }

private:
    Composite::Ptr variables_;
    Composite constraints_;
    Composite costs_;
};

}  // namespace ifopt
#endif
