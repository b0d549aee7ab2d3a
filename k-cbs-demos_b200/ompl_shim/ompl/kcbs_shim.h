// kcbs_shim.h -- the slice of the OMPL (K-CBS fork) API that the K-CBS-Demos plugin headers and demo mains
// use, so that the host-side mirror in ../../includes/ compiles, links and runs without OMPL and Boost.
//
// Data-structure classes (states, spaces, bounds, problem definitions, paths) are plain re-implementations of
// the public OMPL interface.  Everything on the hot path is NOT computed here: validity, propagation, the
// low-level planner (oc::RRT) and the multi-robot planners (omrc::KCBS, omrc::PP) forward to the C ABI of
// libkcbs_b200 (include/kcbs_b200.h) through the hooks declared at the end of this file, which
// includes/kcbs_adaptors.h implements.  With a real OMPL install this directory is simply left off the
// include path.
#pragma once

#include <cmath>
#include <cstring>
#include <fstream>
#include <functional>
#include <iostream>
#include <limits>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#define KCBS_OMPL_SHIM 1

namespace ompl {
namespace base {

// ---- states -----------------------------------------------------------------------------------------------
class State
{
public:
    State() = default;
    State(const State &) = delete;
    State &operator=(const State &) = delete;
    virtual ~State() = default;
    template <class T> const T *as() const { return static_cast<const T *>(this); }
    template <class T> T *as() { return static_cast<T *>(this); }
};

class CompoundState : public State
{
public:
    template <class T> const T *as(unsigned int index) const { return static_cast<const T *>(components[index]); }
    template <class T> T *as(unsigned int index) { return static_cast<T *>(components[index]); }
    State *operator[](unsigned int i) const { return components[i]; }
    State **components = nullptr;
};

// ---- state spaces -------------------------------------------------------------------------------------------
class StateSpace
{
public:
    virtual ~StateSpace() = default;
    template <class T> T *as() { return static_cast<T *>(this); }
    template <class T> const T *as() const { return static_cast<const T *>(this); }
    const std::string &getName() const { return name_; }
    void setName(const std::string &name) { name_ = name; }
    virtual unsigned int getDimension() const = 0;
    virtual State *allocState() const = 0;
    virtual void freeState(State *state) const = 0;
    virtual void copyState(State *destination, const State *source) const = 0;
    virtual double *getValueAddressAtIndex(State *state, unsigned int index) const = 0;
    // OMPL's bounds test (RealVector: value - eps <= high && value + eps >= low with eps = DBL_EPSILON; SO2: strictly
    // inside (-pi - eps, pi + eps); compound: every component).  The GPU kernels use the exact float image of this test.
    virtual bool satisfiesBounds(const State *) const { return true; }
    void copyToReals(std::vector<double> &reals, const State *source) const
    {
        reals.resize(getDimension());
        for (unsigned int i = 0; i < reals.size(); ++i) reals[i] = *getValueAddressAtIndex(const_cast<State *>(source), i);
    }
    void copyFromReals(State *destination, const std::vector<double> &reals) const
    {
        for (unsigned int i = 0; i < reals.size(); ++i) *getValueAddressAtIndex(destination, i) = reals[i];
    }

protected:
    std::string name_;
};
using StateSpacePtr = std::shared_ptr<StateSpace>;

class RealVectorBounds
{
public:
    RealVectorBounds(unsigned int dim) : low(dim, 0.0), high(dim, 0.0) {}
    void setLow(double value) { std::fill(low.begin(), low.end(), value); }
    void setHigh(double value) { std::fill(high.begin(), high.end(), value); }
    void setLow(unsigned int index, double value) { low.at(index) = value; }
    void setHigh(unsigned int index, double value) { high.at(index) = value; }
    std::vector<double> low, high;
};

class RealVectorStateSpace : public StateSpace
{
public:
    class StateType : public State
    {
    public:
        double operator[](unsigned int i) const { return values[i]; }
        double &operator[](unsigned int i) { return values[i]; }
        double *values = nullptr;
    };
    RealVectorStateSpace(unsigned int dim = 0) : dimension_(dim), bounds_(dim) { name_ = "RealVector"; }
    unsigned int getDimension() const override { return dimension_; }
    void setBounds(const RealVectorBounds &bounds) { bounds_ = bounds; }
    void setBounds(double low, double high) { bounds_.setLow(low); bounds_.setHigh(high); }
    const RealVectorBounds &getBounds() const { return bounds_; }
    State *allocState() const override
    {
        auto *s = new StateType();
        // one double of slack: the reference's TwoSecondOrderCarStatePropagator copies 5 values out of (and into) these
        // 4-vectors (StatePropagatorDatabase.h:140, :148); with the slack that stays a benign over-read / over-write
        s->values = new double[dimension_ + 1]();
        return s;
    }
    bool satisfiesBounds(const State *state) const override
    {
        const double *v = static_cast<const StateType *>(state)->values;
        const double eps = std::numeric_limits<double>::epsilon();
        for (unsigned int i = 0; i < dimension_; ++i)
            if (v[i] - eps > bounds_.high[i] || v[i] + eps < bounds_.low[i]) return false;
        return true;
    }
    void freeState(State *state) const override
    {
        auto *s = static_cast<StateType *>(state);
        delete[] s->values;
        delete s;
    }
    void copyState(State *d, const State *s) const override
    {
        std::memcpy(static_cast<StateType *>(d)->values, static_cast<const StateType *>(s)->values, sizeof(double) * dimension_);
    }
    double *getValueAddressAtIndex(State *state, unsigned int index) const override
    {
        return index < dimension_ ? static_cast<StateType *>(state)->values + index : nullptr;
    }

private:
    unsigned int dimension_;
    RealVectorBounds bounds_;
};

class SO2StateSpace : public StateSpace
{
public:
    class StateType : public State
    {
    public:
        void setIdentity() { value = 0.0; }
        double value = 0.0;
    };
    SO2StateSpace() { name_ = "SO2"; }
    unsigned int getDimension() const override { return 1; }
    State *allocState() const override { return new StateType(); }
    void freeState(State *state) const override { delete static_cast<StateType *>(state); }
    void copyState(State *d, const State *s) const override
    {
        static_cast<StateType *>(d)->value = static_cast<const StateType *>(s)->value;
    }
    double *getValueAddressAtIndex(State *state, unsigned int index) const override
    {
        return index == 0 ? &static_cast<StateType *>(state)->value : nullptr;
    }
    bool satisfiesBounds(const State *state) const override
    {
        const double pi = 3.14159265358979323846, eps = std::numeric_limits<double>::epsilon();
        const double v = static_cast<const StateType *>(state)->value;
        return v < pi + eps && v > -pi - eps;
    }
    // OMPL: wraps to [-pi, pi).  Only reached through the scalar post-integration hook; batches wrap on the GPU.
    void enforceBounds(State *state) const
    {
        const double pi = 3.14159265358979323846;
        double v = std::fmod(static_cast<StateType *>(state)->value, 2.0 * pi);
        if (v < -pi)
            v += 2.0 * pi;
        else if (v >= pi)
            v -= 2.0 * pi;
        static_cast<StateType *>(state)->value = v;
    }
};

class CompoundStateSpace : public StateSpace
{
public:
    using StateType = CompoundState;
    CompoundStateSpace() { name_ = "Compound"; }
    template <class T> T *as(unsigned int index) const { return static_cast<T *>(components_.at(index).get()); }
    using StateSpace::as;
    void addSubspace(const StateSpacePtr &component, double weight)
    {
        if (locked_) throw std::runtime_error("This state space is locked. No further components can be added");
        components_.push_back(component);
        weights_.push_back(weight);
    }
    void lock() { locked_ = true; }
    unsigned int getSubspaceCount() const { return (unsigned int)components_.size(); }
    const StateSpacePtr &getSubspace(unsigned int index) const { return components_.at(index); }
    double getSubspaceWeight(unsigned int index) const { return weights_.at(index); }
    unsigned int getDimension() const override
    {
        unsigned int d = 0;
        for (const auto &c : components_) d += c->getDimension();
        return d;
    }
    State *allocState() const override
    {
        auto *s = new CompoundState();
        s->components = new State *[components_.size()];
        for (size_t i = 0; i < components_.size(); ++i) s->components[i] = components_[i]->allocState();
        return s;
    }
    void freeState(State *state) const override
    {
        auto *s = static_cast<CompoundState *>(state);
        for (size_t i = 0; i < components_.size(); ++i) components_[i]->freeState(s->components[i]);
        delete[] s->components;
        delete s;
    }
    void copyState(State *d, const State *s) const override
    {
        for (size_t i = 0; i < components_.size(); ++i)
            components_[i]->copyState(static_cast<CompoundState *>(d)->components[i], static_cast<const CompoundState *>(s)->components[i]);
    }
    bool satisfiesBounds(const State *state) const override
    {
        const auto *s = static_cast<const CompoundState *>(state);
        for (size_t i = 0; i < components_.size(); ++i)
            if (!components_[i]->satisfiesBounds(s->components[i])) return false;
        return true;
    }
    double *getValueAddressAtIndex(State *state, unsigned int index) const override
    {
        auto *s = static_cast<CompoundState *>(state);
        for (size_t i = 0; i < components_.size(); ++i) {
            const unsigned int d = components_[i]->getDimension();
            if (index < d) return components_[i]->getValueAddressAtIndex(s->components[i], index);
            index -= d;
        }
        return nullptr;
    }

private:
    std::vector<StateSpacePtr> components_;
    std::vector<double> weights_;
    bool locked_ = false;
};

template <class T = StateSpace>
class ScopedState
{
public:
    explicit ScopedState(const StateSpacePtr &space) : space_(space), state_(space->allocState()) {}
    ScopedState(const ScopedState &o) : space_(o.space_), state_(o.space_->allocState()) { space_->copyState(state_, o.state_); }
    ScopedState &operator=(const ScopedState &) = delete;
    ~ScopedState() { space_->freeState(state_); }
    double &operator[](unsigned int index)
    {
        double *p = space_->getValueAddressAtIndex(state_, index);
        if (!p) throw std::out_of_range("Index out of bounds");
        return *p;
    }
    double operator[](unsigned int index) const { return *space_->getValueAddressAtIndex(state_, index); }
    State *get() { return state_; }
    const State *get() const { return state_; }
    const StateSpacePtr &getSpace() const { return space_; }

private:
    StateSpacePtr space_;
    State *state_;
};

// ---- space information, validity, goals, problems, planners ----------------------------------------------------
class SpaceInformation;
using SpaceInformationPtr = std::shared_ptr<SpaceInformation>;

class StateValidityChecker
{
public:
    StateValidityChecker(SpaceInformation *si) : si_(si) {}
    StateValidityChecker(const SpaceInformationPtr &si) : si_(si.get()) {}
    virtual ~StateValidityChecker() = default;
    virtual bool isValid(const State *state) const = 0;
    // added by the K-CBS fork: does `state1` of this robot avoid the robot of state2.first at state2.second?
    virtual bool areStatesValid(const State *, const std::pair<const SpaceInformationPtr, const State *>) const { return true; }

protected:
    SpaceInformation *si_;
};
using StateValidityCheckerPtr = std::shared_ptr<StateValidityChecker>;

class SpaceInformation
{
public:
    SpaceInformation(StateSpacePtr space) : stateSpace_(std::move(space)) {}
    virtual ~SpaceInformation() = default;
    const StateSpacePtr &getStateSpace() const { return stateSpace_; }
    void setStateValidityChecker(const StateValidityCheckerPtr &svc) { svc_ = svc; }
    const StateValidityCheckerPtr &getStateValidityChecker() const { return svc_; }
    bool isValid(const State *state) const { return svc_->isValid(state); }
    bool satisfiesBounds(const State *state) const { return stateSpace_->satisfiesBounds(state); }
    State *allocState() const { return stateSpace_->allocState(); }
    void freeState(State *s) const { stateSpace_->freeState(s); }
    void copyState(State *d, const State *s) const { stateSpace_->copyState(d, s); }
    State *cloneState(const State *s) const
    {
        State *c = allocState();
        copyState(c, s);
        return c;
    }
    // fork addition (SystemMergerDatabase.h:91): drops the K-CBS constraints of this individual
    void clearDynamicObstacles() {}
    void setup() {}

protected:
    StateSpacePtr stateSpace_;
    StateValidityCheckerPtr svc_;
};

class Goal
{
public:
    Goal(const SpaceInformationPtr &si) : si_(si) {}
    virtual ~Goal() = default;
    template <class T> T *as() { return static_cast<T *>(this); }
    template <class T> const T *as() const { return static_cast<const T *>(this); }
    virtual bool isSatisfied(const State *st) const = 0;
    virtual bool isSatisfied(const State *st, double *distance) const
    {
        if (distance) *distance = std::numeric_limits<double>::max();
        return isSatisfied(st);
    }
    const SpaceInformationPtr &getSpaceInformation() const { return si_; }

protected:
    SpaceInformationPtr si_;
};
using GoalPtr = std::shared_ptr<Goal>;

class GoalRegion : public Goal
{
public:
    GoalRegion(const SpaceInformationPtr &si) : Goal(si) {}
    virtual double distanceGoal(const State *st) const = 0;
    bool isSatisfied(const State *st) const override { return distanceGoal(st) <= threshold_; }
    bool isSatisfied(const State *st, double *distance) const override
    {
        const double d = distanceGoal(st);
        if (distance) *distance = d;
        return d <= threshold_;
    }
    void setThreshold(double t) { threshold_ = t; }
    double getThreshold() const { return threshold_; }

protected:
    double threshold_ = std::numeric_limits<double>::epsilon();
};

class Path
{
public:
    virtual ~Path() = default;
    template <class T> T *as() { return static_cast<T *>(this); }
    template <class T> const T *as() const { return static_cast<const T *>(this); }
};
using PathPtr = std::shared_ptr<Path>;

class ProblemDefinition
{
public:
    ProblemDefinition(SpaceInformationPtr si) : si_(std::move(si)) {}
    ~ProblemDefinition()
    {
        for (State *s : starts_) si_->freeState(s);
    }
    void addStartState(const State *state) { starts_.push_back(si_->cloneState(state)); }
    template <class T> void addStartState(const ScopedState<T> &state) { addStartState(state.get()); }
    unsigned int getStartStateCount() const { return (unsigned int)starts_.size(); }
    const State *getStartState(unsigned int i) const { return starts_.at(i); }
    void setGoal(const GoalPtr &goal) { goal_ = goal; }
    const GoalPtr &getGoal() const { return goal_; }
    void clearSolutionPaths() { solutions_.clear(); }
    void addSolutionPath(const PathPtr &p) { solutions_.push_back(p); }
    bool hasSolution() const { return !solutions_.empty(); }
    PathPtr getSolutionPath() const { return solutions_.empty() ? nullptr : solutions_.back(); }
    const SpaceInformationPtr &getSpaceInformation() const { return si_; }

private:
    SpaceInformationPtr si_;
    std::vector<State *> starts_;
    GoalPtr goal_;
    std::vector<PathPtr> solutions_;
};
using ProblemDefinitionPtr = std::shared_ptr<ProblemDefinition>;

struct PlannerStatus
{
    enum StatusType { UNKNOWN = 0, INVALID_START, INVALID_GOAL, UNRECOGNIZED_GOAL_TYPE, TIMEOUT, APPROXIMATE_SOLUTION,
                      EXACT_SOLUTION, CRASH, ABORT, TYPE_COUNT };
    PlannerStatus(StatusType s = UNKNOWN) : status_(s) {}
    PlannerStatus(bool hasSolution, bool isApproximate)
        : status_(hasSolution ? (isApproximate ? APPROXIMATE_SOLUTION : EXACT_SOLUTION) : TIMEOUT) {}
    operator StatusType() const { return status_; }
    explicit operator bool() const { return status_ == APPROXIMATE_SOLUTION || status_ == EXACT_SOLUTION; }
    std::string asString() const
    {
        static const char *n[] = {"Unknown status", "Invalid start", "Invalid goal", "Unrecognized goal type", "Timeout",
                                  "Approximate solution", "Exact solution", "Crash", "Abort"};
        return n[status_];
    }
    StatusType status_;
};

class Planner
{
public:
    Planner(SpaceInformationPtr si, std::string name) : si_(std::move(si)), name_(std::move(name)) {}
    virtual ~Planner() = default;
    template <class T> T *as() { return static_cast<T *>(this); }
    template <class T> const T *as() const { return static_cast<const T *>(this); }
    virtual void setProblemDefinition(const ProblemDefinitionPtr &pdef) { pdef_ = pdef; }
    const ProblemDefinitionPtr &getProblemDefinition() const { return pdef_; }
    const SpaceInformationPtr &getSpaceInformation() const { return si_; }
    const std::string &getName() const { return name_; }
    virtual PlannerStatus solve(double solveTime) = 0;
    virtual void clear() {}
    virtual void setup() {}

protected:
    SpaceInformationPtr si_;
    ProblemDefinitionPtr pdef_;
    std::string name_;
};
using PlannerPtr = std::shared_ptr<Planner>;
using PlannerAllocator = std::function<PlannerPtr(const SpaceInformationPtr &)>;

}  // namespace base

// ---- ompl::control ---------------------------------------------------------------------------------------------
namespace control {

class Control
{
public:
    Control() = default;
    virtual ~Control() = default;
    template <class T> const T *as() const { return static_cast<const T *>(this); }
    template <class T> T *as() { return static_cast<T *>(this); }
};

class ControlSpace
{
public:
    ControlSpace(base::StateSpacePtr stateSpace) : stateSpace_(std::move(stateSpace)) {}
    virtual ~ControlSpace() = default;
    template <class T> T *as() { return static_cast<T *>(this); }
    virtual unsigned int getDimension() const = 0;
    virtual Control *allocControl() const = 0;
    virtual void freeControl(Control *c) const = 0;
    virtual void copyControl(Control *d, const Control *s) const = 0;
    const base::StateSpacePtr &getStateSpace() const { return stateSpace_; }

protected:
    base::StateSpacePtr stateSpace_;
};
using ControlSpacePtr = std::shared_ptr<ControlSpace>;

class RealVectorControlSpace : public ControlSpace
{
public:
    class ControlType : public Control
    {
    public:
        double operator[](unsigned int i) const { return values[i]; }
        double &operator[](unsigned int i) { return values[i]; }
        double *values = nullptr;
    };
    RealVectorControlSpace(const base::StateSpacePtr &stateSpace, unsigned int dim)
        : ControlSpace(stateSpace), dimension_(dim), bounds_(dim) {}
    unsigned int getDimension() const override { return dimension_; }
    void setBounds(const base::RealVectorBounds &bounds) { bounds_ = bounds; }
    const base::RealVectorBounds &getBounds() const { return bounds_; }
    Control *allocControl() const override
    {
        auto *c = new ControlType();
        c->values = new double[dimension_ ? dimension_ : 1]();
        return c;
    }
    void freeControl(Control *c) const override
    {
        delete[] static_cast<ControlType *>(c)->values;
        delete static_cast<ControlType *>(c);
    }
    void copyControl(Control *d, const Control *s) const override
    {
        std::memcpy(static_cast<ControlType *>(d)->values, static_cast<const ControlType *>(s)->values, sizeof(double) * dimension_);
    }

private:
    unsigned int dimension_;
    base::RealVectorBounds bounds_;
};

class SpaceInformation;
using SpaceInformationPtr = std::shared_ptr<SpaceInformation>;

class StatePropagator
{
public:
    StatePropagator(SpaceInformation *si) : si_(si) {}
    StatePropagator(const SpaceInformationPtr &si) : si_(si.get()) {}
    virtual ~StatePropagator() = default;
    virtual void propagate(const base::State *state, const Control *control, double duration, base::State *result) const = 0;
    virtual bool canPropagateBackward() const { return true; }

protected:
    SpaceInformation *si_;
};
using StatePropagatorPtr = std::shared_ptr<StatePropagator>;

class SpaceInformation : public base::SpaceInformation
{
public:
    SpaceInformation(const base::StateSpacePtr &stateSpace, ControlSpacePtr controlSpace)
        : base::SpaceInformation(stateSpace), controlSpace_(std::move(controlSpace)) {}
    const ControlSpacePtr &getControlSpace() const { return controlSpace_; }
    Control *allocControl() const { return controlSpace_->allocControl(); }
    void freeControl(Control *c) const { controlSpace_->freeControl(c); }
    void setStatePropagator(const StatePropagatorPtr &sp) { statePropagator_ = sp; }
    const StatePropagatorPtr &getStatePropagator() const { return statePropagator_; }
    void setPropagationStepSize(double stepSize) { stepSize_ = stepSize; }
    double getPropagationStepSize() const { return stepSize_; }
    void setMinMaxControlDuration(unsigned int minSteps, unsigned int maxSteps) { minSteps_ = minSteps; maxSteps_ = maxSteps; }
    unsigned int getMinControlDuration() const { return minSteps_; }
    unsigned int getMaxControlDuration() const { return maxSteps_; }
    void propagate(const base::State *state, const Control *control, int steps, base::State *result) const
    {
        if (steps == 0) {
            if (result != state) copyState(result, state);
            return;
        }
        statePropagator_->propagate(state, control, steps * stepSize_, result);
    }
    // OMPL's loop over the scalar virtuals: one propagate + one isValid per step, stop at the first invalid state
    unsigned int propagateWhileValid(const base::State *state, const Control *control, int steps, base::State *result) const
    {
        if (steps <= 0) {
            if (result != state) copyState(result, state);
            return 0;
        }
        base::State *cur = cloneState(state), *next = allocState();
        unsigned int done = 0;
        for (int i = 0; i < steps; ++i) {
            statePropagator_->propagate(cur, control, stepSize_, next);
            if (!isValid(next)) break;
            std::swap(cur, next);
            ++done;
        }
        copyState(result, cur);
        freeState(cur);
        freeState(next);
        return done;
    }

protected:
    ControlSpacePtr controlSpace_;
    StatePropagatorPtr statePropagator_;
    double stepSize_ = 0.05;
    unsigned int minSteps_ = 1, maxSteps_ = 2;
};

class ODESolver
{
public:
    typedef std::vector<double> StateType;
    typedef std::function<void(const StateType &, const Control *, StateType &)> ODE;
    typedef std::function<void(const base::State *, const Control *, double, base::State *)> PostPropagationEvent;
    ODESolver(SpaceInformationPtr si, ODE ode, double intStep) : si_(std::move(si)), ode_(std::move(ode)), intStep_(intStep) {}
    virtual ~ODESolver() = default;
    double getIntegrationStepSize() const { return intStep_; }
    void setIntegrationStepSize(double s) { intStep_ = s; }
    const SpaceInformationPtr &getSpaceInformation() const { return si_; }
    const ODE &getODE() const { return ode_; }
    // implemented by the adaptors: returns a propagator that integrates on the GPU through the C ABI
    static StatePropagatorPtr getStatePropagator(std::shared_ptr<ODESolver> solver, const PostPropagationEvent &postEvent = nullptr);

protected:
    SpaceInformationPtr si_;
    ODE ode_;
    double intStep_;
};
using ODESolverPtr = std::shared_ptr<ODESolver>;

template <class Solver = void>
class ODEBasicSolver : public ODESolver
{
public:
    ODEBasicSolver(const SpaceInformationPtr &si, const ODESolver::ODE &ode, double intStep = 1e-2) : ODESolver(si, ode, intStep) {}
};

// A path of states connected by (control, duration) edges.
class PathControl : public base::Path
{
public:
    PathControl(base::SpaceInformationPtr si) : si_(std::move(si)) {}
    ~PathControl() override
    {
        for (auto *s : states_) si_->freeState(s);
        auto *siC = static_cast<SpaceInformation *>(si_.get());
        for (auto *c : controls_) siC->freeControl(c);
    }
    void append(const base::State *state) { states_.push_back(si_->cloneState(state)); }
    void append(const base::State *state, const Control *control, double duration)
    {
        auto *siC = static_cast<SpaceInformation *>(si_.get());
        states_.push_back(si_->cloneState(state));
        Control *c = siC->allocControl();
        siC->getControlSpace()->copyControl(c, control);
        controls_.push_back(c);
        durations_.push_back(duration);
    }
    std::vector<base::State *> &getStates() { return states_; }
    std::vector<Control *> &getControls() { return controls_; }
    std::vector<double> &getControlDurations() { return durations_; }
    std::size_t getStateCount() const { return states_.size(); }
    std::size_t getControlCount() const { return controls_.size(); }
    base::State *getState(unsigned int i) { return states_[i]; }
    double length() const
    {
        double l = 0;
        for (double d : durations_) l += d;
        return l;
    }
    // paths produced by the GPU planners already hold one state per propagation step
    void interpolate() {}
    void printAsMatrix(std::ostream &out) const
    {
        std::vector<double> reals;
        for (std::size_t i = 0; i < states_.size(); ++i) {
            si_->getStateSpace()->copyToReals(reals, states_[i]);
            for (double r : reals) out << r << ' ';
            if (i < controls_.size()) {
                auto *siC = static_cast<SpaceInformation *>(si_.get());
                const unsigned int cd = siC->getControlSpace()->getDimension();
                const double *v = controls_[i]->as<RealVectorControlSpace::ControlType>()->values;
                for (unsigned int j = 0; j < cd; ++j) out << v[j] << ' ';
                out << durations_[i];
            }
            out << std::endl;
        }
    }

private:
    base::SpaceInformationPtr si_;
    std::vector<base::State *> states_;
    std::vector<Control *> controls_;
    std::vector<double> durations_;
};

// oc::RRT: the low-level planner allocateControlRRT returns (PlannerAllocatorDatabase.h:40-45).  solve() runs the
// batched GPU RRT (kcbs_rrt_plan); implemented by the adaptors.
class RRT : public base::Planner
{
public:
    RRT(const SpaceInformationPtr &si) : base::Planner(si, "RRT"), siC_(si.get()) {}
    base::PlannerStatus solve(double solveTime) override;
    void setGoalBias(double b) { goalBias_ = b; }
    double getGoalBias() const { return goalBias_; }

protected:
    SpaceInformation *siC_;
    double goalBias_ = 0.05;
};

}  // namespace control

// ---- ompl::multirobot ------------------------------------------------------------------------------------------
namespace multirobot {
namespace control {
class SystemMerger;
using SystemMergerPtr = std::shared_ptr<SystemMerger>;

class SpaceInformation
{
public:
    void addIndividual(const ompl::control::SpaceInformationPtr &individual)
    {
        if (locked_) throw std::runtime_error("This SpaceInformation is locked. No further individuals can be added");
        individuals_.push_back(individual);
    }
    const ompl::control::SpaceInformationPtr &getIndividual(unsigned int index) const { return individuals_.at(index); }
    unsigned int getIndividualCount() const { return (unsigned int)individuals_.size(); }
    void lock() { locked_ = true; }
    void setPlannerAllocator(const base::PlannerAllocator &pa) { pa_ = pa; }
    const base::PlannerAllocator &getPlannerAllocator() const { return pa_; }
    void setSystemMerger(const SystemMergerPtr &merger) { merger_ = merger; }
    const SystemMergerPtr &getSystemMerger() const { return merger_; }

private:
    std::vector<ompl::control::SpaceInformationPtr> individuals_;
    base::PlannerAllocator pa_;
    SystemMergerPtr merger_;
    bool locked_ = false;
};
using SpaceInformationPtr = std::shared_ptr<SpaceInformation>;
}  // namespace control

namespace base {
class Plan
{
public:
    virtual ~Plan() = default;
    template <class T> T *as() { return static_cast<T *>(this); }
    template <class T> const T *as() const { return static_cast<const T *>(this); }
};
using PlanPtr = std::shared_ptr<Plan>;

class ProblemDefinition
{
public:
    ProblemDefinition(control::SpaceInformationPtr si) : si_(std::move(si)) {}
    void addIndividual(const ompl::base::ProblemDefinitionPtr &pdef)
    {
        if (locked_) throw std::runtime_error("This ProblemDefinition is locked. No further individuals can be added");
        individuals_.push_back(pdef);
    }
    const ompl::base::ProblemDefinitionPtr &getIndividual(unsigned int index) const { return individuals_.at(index); }
    unsigned int getIndividualCount() const { return (unsigned int)individuals_.size(); }
    void lock() { locked_ = true; }
    void addSolutionPlan(const PlanPtr &plan) { plans_.push_back(plan); }
    bool hasSolution() const { return !plans_.empty(); }
    PlanPtr getSolutionPlan() const { return plans_.empty() ? nullptr : plans_.back(); }
    void clearSolutionPaths()
    {
        plans_.clear();
        for (auto &p : individuals_) p->clearSolutionPaths();
    }

private:
    control::SpaceInformationPtr si_;
    std::vector<ompl::base::ProblemDefinitionPtr> individuals_;
    std::vector<PlanPtr> plans_;
    bool locked_ = false;
};
using ProblemDefinitionPtr = std::shared_ptr<ProblemDefinition>;

class Planner
{
public:
    Planner(control::SpaceInformationPtr si, std::string name) : si_(std::move(si)), name_(std::move(name)) {}
    virtual ~Planner() = default;
    template <class T> T *as() { return static_cast<T *>(this); }
    template <class T> const T *as() const { return static_cast<const T *>(this); }
    virtual void setProblemDefinition(const ProblemDefinitionPtr &pdef) { pdef_ = pdef; }
    const ProblemDefinitionPtr &getProblemDefinition() const { return pdef_; }
    virtual ompl::base::PlannerStatus solve(double solveTime) = 0;
    virtual void clear() {}
    const std::string &getName() const { return name_; }

protected:
    control::SpaceInformationPtr si_;
    ProblemDefinitionPtr pdef_;
    std::string name_;
};
using PlannerPtr = std::shared_ptr<Planner>;
}  // namespace base

namespace control {

class SystemMerger
{
public:
    SystemMerger(const SpaceInformationPtr &si, const base::ProblemDefinitionPtr &pdef) : si_(si), pdef_(pdef) {}
    virtual ~SystemMerger() = default;
    virtual std::pair<const SpaceInformationPtr, const base::ProblemDefinitionPtr> merge(const int index1, const int index2) const = 0;

protected:
    const SpaceInformationPtr si_;
    const base::ProblemDefinitionPtr pdef_;
};

// The multi-robot plan: one PathControl per individual, in addIndividual order.
class PlanControl : public base::Plan
{
public:
    void append(const std::shared_ptr<ompl::control::PathControl> &path) { paths_.push_back(path); }
    std::size_t getPathCount() const { return paths_.size(); }
    const std::shared_ptr<ompl::control::PathControl> &getPath(unsigned int i) const { return paths_.at(i); }
    void interpolate() {}
    // one block per robot: "<prefix> <index>" followed by PathControl::printAsMatrix
    void printAsMatrix(std::ostream &out, const std::string &prefix = "Robot") const
    {
        for (std::size_t i = 0; i < paths_.size(); ++i) {
            out << prefix << ' ' << i << std::endl;
            paths_[i]->printAsMatrix(out);
        }
    }

private:
    std::vector<std::shared_ptr<ompl::control::PathControl>> paths_;
};

// omrc::KCBS (demo_*.cpp:155-168, Benchmark.h:151-173).  solve() runs kcbs_kcbs_solve; implemented by the adaptors.
class KCBS : public base::Planner
{
public:
    KCBS(const SpaceInformationPtr &si) : base::Planner(si, "K-CBS") {}
    void setLowLevelSolveTime(double t) { llSolveTime_ = t; }
    double getLowLevelSolveTime() const { return llSolveTime_; }
    void setMergeBound(int b) { mergeBound_ = b; }
    int getMergeBound() const { return mergeBound_; }
    double getRootSolveTime() const { return rootSolveTime_; }
    unsigned int getNumberOfNodesExpanded() const { return nodesExpanded_; }
    unsigned int getNumberOfApproximateSolutions() const { return approximateSolutions_; }
    ompl::base::PlannerStatus solve(double solveTime) override;
    void printConstraintTree(std::ostream &out) const { out << treeLog_; }
    void clear() override
    {
        nodesExpanded_ = approximateSolutions_ = 0;
        rootSolveTime_ = 0;
        treeLog_.clear();
    }

protected:
    double llSolveTime_ = 1.0, rootSolveTime_ = 0.0;
    int mergeBound_ = std::numeric_limits<int>::max();
    unsigned int nodesExpanded_ = 0, approximateSolutions_ = 0;
    std::string treeLog_;
};

// omrc::PP: prioritized planning (Benchmark.h:177-198); implemented by the adaptors on kcbs_rrt_plan.
class PP : public base::Planner
{
public:
    PP(const SpaceInformationPtr &si) : base::Planner(si, "PP") {}
    ompl::base::PlannerStatus solve(double solveTime) override;
};

}  // namespace control
}  // namespace multirobot
}  // namespace ompl
