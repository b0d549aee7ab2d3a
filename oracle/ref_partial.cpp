// ref_partial.cpp -- the part of the reference that DOES compile here, used to pin the oracle (test infrastructure).
//
// The reference (aria-systems-group/K-CBS-Demos) needs the OMPL K-CBS fork and Boost, neither of which is installed,
// so its hot path cannot be built as a whole.  Its headers, however, depend on OMPL only through declarations that the
// API shim of this repository provides (k-cbs-demos_b200/ompl_shim) and on Boost.Geometry only through a handful of
// primitives (oracle/boost_standin: point, polygon, correct, matrix transform, disjoint); everything else -- the
// arithmetic AND the control flow -- is the reference's own code:
//
//   includes/StatePropagatorDatabase.h   SecondOrderCarODE :44-64, TwoSecondOrderCarsODE :77-102  (the ODE right-hand sides)
//                                        TwoSecondOrderCarStatePropagator::propagate :130-153      (the freeze-at-goal rule)
//   includes/StateSpaceDatabase.h        createBounded[Two]2ndOrderCarStateSpace                   (the state bounds)
//   includes/ControlSpaceDatabase.h      createUniform{2,4}DRealVectorControlSpace                 (the control bounds)
//   includes/GoalRegionDatabase.h        GoalRegion2ndOrderCar :42-60, GoalRegionTwo2ndOrderCars :61-108
//   includes/Robot.h, Obstacle.h         shapes, centres, bounding radii
//   includes/StateValidityCheckerDatabase.h   isValid :54-75, areStatesValid :78-125, the quick check :134-141, the exact
//                                        checks :144-205 (the rotation matrix handed to transform), and the merged
//                                        pair's isValid :225-261 / areStatesValid :264-287
//   includes/SystemMergerDatabase.h      homogeneous2ndOrderCarSystemMerger::merge :55-168          (the wiring)
//
// This file #includes those headers WHERE THEY LIE under /root/reference (nothing is copied) and exports their functions
// through a C ABI.  What stays restated (not reference code): OMPL's satisfiesBounds / SO2 wrap (the shim), the RK4
// driver behind ODEBasicSolver (rk4_propagator below, odeint's published algorithm), and the Boost.Geometry primitives
// (the stand-in).  Built by oracle/Makefile into oracle/_ref/ (git-ignored) when /root/reference exists.
#include <cmath>
#include <cstring>
#include <memory>
#include <vector>

#include <map>
#include <set>
#include <string>
#include <unordered_map>

#include <ompl/control/SpaceInformation.h>
#include <ompl/multirobot/control/SpaceInformation.h>

#include <StatePropagatorDatabase.h>
#include <StateSpaceDatabase.h>
#include <ControlSpaceDatabase.h>
#include <GoalRegionDatabase.h>
#include <StateValidityCheckerDatabase.h>
#include <PlannerAllocatorDatabase.h>
#include <SystemMergerDatabase.h>

// ---- hooks the shim leaves to its user ----------------------------------------------------------------------------
// The product implements these on the GPU (includes/kcbs_adaptors.h).  Here ODEBasicSolver gets a CPU stand-in: classical
// RK4 with odeint's integrate_const stepping rule over the REFERENCE's ODE functor; planners are never run.
namespace {
class Rk4Propagator : public oc::StatePropagator
{
public:
    Rk4Propagator(const std::shared_ptr<oc::ODESolver> &solver, oc::ODESolver::PostPropagationEvent post)
        : oc::StatePropagator(solver->getSpaceInformation()), solver_(solver), post_(std::move(post)) {}
    void propagate(const ob::State *state, const oc::Control *control, double duration, ob::State *result) const override
    {
        const auto &space = solver_->getSpaceInformation()->getStateSpace();
        std::vector<double> x;
        space->copyToReals(x, state);
        const double h = solver_->getIntegrationStepSize();
        const auto &f = solver_->getODE();
        const size_t n = x.size();
        std::vector<double> k1(n), k2(n), k3(n), k4(n), xt(n);
        // odeint::integrate_const(stepper, system, x, 0, duration, h): steps while (time + h) <= duration (with the
        // library's epsilon slack), time recomputed as start + step * h
        double time = 0.0;
        int step = 0;
        while ((time + h) - duration <= std::numeric_limits<double>::epsilon()) {
            f(x, control, k1);
            for (size_t i = 0; i < n; ++i) xt[i] = x[i] + h * 0.5 * k1[i];
            f(xt, control, k2);
            for (size_t i = 0; i < n; ++i) xt[i] = x[i] + h * 0.5 * k2[i];
            f(xt, control, k3);
            for (size_t i = 0; i < n; ++i) xt[i] = x[i] + h * k3[i];
            f(xt, control, k4);
            for (size_t i = 0; i < n; ++i) x[i] += h / 6.0 * k1[i] + h / 3.0 * k2[i] + h / 3.0 * k3[i] + h / 6.0 * k4[i];
            ++step;
            time = 0.0 + (double)step * h;
        }
        space->copyFromReals(result, x);
        if (post_) post_(state, control, duration, result);
    }

private:
    std::shared_ptr<oc::ODESolver> solver_;
    oc::ODESolver::PostPropagationEvent post_;
};
}  // namespace

namespace ompl {
namespace control {
inline StatePropagatorPtr ODESolver::getStatePropagator(std::shared_ptr<ODESolver> solver, const PostPropagationEvent &postEvent)
{
    return std::make_shared<Rk4Propagator>(solver, postEvent);
}
inline base::PlannerStatus RRT::solve(double) { throw std::runtime_error("planners are not part of the reference pin"); }
}  // namespace control
namespace multirobot {
namespace control {
inline ompl::base::PlannerStatus KCBS::solve(double) { throw std::runtime_error("planners are not part of the reference pin"); }
inline ompl::base::PlannerStatus PP::solve(double) { throw std::runtime_error("planners are not part of the reference pin"); }
}  // namespace control
}  // namespace multirobot
}  // namespace ompl

namespace {
struct Ctl : oc::RealVectorControlSpace::ControlType {
    explicit Ctl(const double *u, int n) : store(u, u + n) { values = store.data(); }
    std::vector<double> store;
};
}  // namespace

extern "C" {

// SecondOrderCarODE, StatePropagatorDatabase.h:44-64
void ref_second_order_car_ode(const double q[5], const double u[2], double qdot[5])
{
    oc::ODESolver::StateType qs(q, q + 5), qd;
    Ctl c(u, 2);
    SecondOrderCarODE(qs, &c, qd);
    std::memcpy(qdot, qd.data(), sizeof(double) * 5);
}

// TwoSecondOrderCarsODE, StatePropagatorDatabase.h:77-102
void ref_two_second_order_cars_ode(const double q[10], const double u[4], double qdot[10])
{
    oc::ODESolver::StateType qs(q, q + 10), qd;
    Ctl c(u, 4);
    TwoSecondOrderCarsODE(qs, &c, qd);
    std::memcpy(qdot, qd.data(), sizeof(double) * 10);
}

// createBounded2ndOrderCarStateSpace, StateSpaceDatabase.h:42-62: bounds of (x, y, v, phi) and the subspace count
int ref_state_bounds(unsigned x_max, unsigned y_max, double lo[4], double hi[4])
{
    ob::StateSpacePtr space = createBounded2ndOrderCarStateSpace(x_max, y_max);
    const ob::RealVectorBounds &b = space->as<ob::CompoundStateSpace>()->as<ob::RealVectorStateSpace>(0)->getBounds();
    for (int i = 0; i < 4; ++i) {
        lo[i] = b.low[i];
        hi[i] = b.high[i];
    }
    return (int)space->as<ob::CompoundStateSpace>()->getSubspaceCount();
}

// createUniform2DRealVectorControlSpace, ControlSpaceDatabase.h:43-54
void ref_control_bounds(double lo[2], double hi[2])
{
    ob::StateSpacePtr space = createBounded2ndOrderCarStateSpace(10, 10);
    oc::ControlSpacePtr cs = createUniform2DRealVectorControlSpace(space);
    const ob::RealVectorBounds &b = cs->as<oc::RealVectorControlSpace>()->getBounds();
    for (int i = 0; i < 2; ++i) {
        lo[i] = b.low[i];
        hi[i] = b.high[i];
    }
}

// GoalRegion2ndOrderCar::distanceGoal and its threshold, GoalRegionDatabase.h:42-60
double ref_goal_distance(double x, double y, double gx, double gy, double *threshold)
{
    ob::StateSpacePtr space = createBounded2ndOrderCarStateSpace(32, 32);
    auto si = std::make_shared<ob::SpaceInformation>(space);
    GoalRegion2ndOrderCar goal(si, gx, gy);
    ob::State *st = space->allocState();
    double *pos = st->as<ob::CompoundStateSpace::StateType>()->as<ob::RealVectorStateSpace::StateType>(0)->values;
    pos[0] = x;
    pos[1] = y;
    const double d = goal.distanceGoal(st);
    if (threshold) *threshold = goal.getThreshold();
    space->freeState(st);
    return d;
}

}  // extern "C"

// ---- scene helpers for the validity checkers ----------------------------------------------------------------------
namespace {
struct RefScene {
    std::unordered_map<std::string, Robot *> robots;
    std::set<Obstacle *> obstacles;
    std::vector<std::unique_ptr<Robot>> robot_store;
    std::vector<std::unique_ptr<Obstacle>> obstacle_store;
    ob::StateSpacePtr space;
    oc::SpaceInformationPtr si;
};

// one single-car individual named `name` in a workspace with the given obstacles; `others` = further robots of the map
std::unique_ptr<RefScene> make_scene(unsigned x_max, unsigned y_max, int n_obs, const double *obs4, const char *name, double L, double W,
                                     const char *name2 = nullptr, double L2 = 0, double W2 = 0)
{
    auto sc = std::make_unique<RefScene>();
    for (int o = 0; o < n_obs; ++o) {
        sc->obstacle_store.push_back(std::make_unique<RectangularObstacle>(obs4[4 * o], obs4[4 * o + 1], obs4[4 * o + 2], obs4[4 * o + 3]));
        sc->obstacles.insert(sc->obstacle_store.back().get());
    }
    sc->robot_store.push_back(std::make_unique<RectangularRobot>(name, L, W));
    sc->robots[name] = sc->robot_store.back().get();
    if (name2) {
        sc->robot_store.push_back(std::make_unique<RectangularRobot>(name2, L2, W2));
        sc->robots[name2] = sc->robot_store.back().get();
    }
    sc->space = createBounded2ndOrderCarStateSpace(x_max, y_max);
    sc->space->setName(name);
    sc->si = std::make_shared<oc::SpaceInformation>(sc->space, createUniform2DRealVectorControlSpace(sc->space));
    return sc;
}

void set_car(ob::State *st, unsigned car, const double q[5])
{
    auto *cs = st->as<ob::CompoundStateSpace::StateType>();
    double *v = cs->as<ob::RealVectorStateSpace::StateType>(2 * car)->values;
    v[0] = q[0]; v[1] = q[1]; v[2] = q[2]; v[3] = q[3];
    cs->as<ob::SO2StateSpace::StateType>(2 * car + 1)->value = q[4];
}
void get_car(const ob::State *st, unsigned car, double q[5])
{
    const auto *cs = st->as<ob::CompoundStateSpace::StateType>();
    const double *v = cs->as<ob::RealVectorStateSpace::StateType>(2 * car)->values;
    q[0] = v[0]; q[1] = v[1]; q[2] = v[2]; q[3] = v[3];
    q[4] = cs->as<ob::SO2StateSpace::StateType>(2 * car + 1)->value;
}
}  // namespace

extern "C" {

// homogeneous2ndOrderCarSystemSVC::isValid, StateValidityCheckerDatabase.h:54-75, for n states (x, y, v, phi, theta) of a
// robot of size L x W among n_obs rectangles (x, y, length, width).
void ref_is_valid(unsigned x_max, unsigned y_max, int n_obs, const double *obs4, double L, double W, int n, const double *q, int *out)
{
    auto sc = make_scene(x_max, y_max, n_obs, obs4, "Robot 1", L, W);
    homogeneous2ndOrderCarSystemSVC svc(sc->si, sc->robots, sc->obstacles);
    ob::State *st = sc->space->allocState();
    for (int i = 0; i < n; ++i) {
        set_car(st, 0, q + 5 * i);
        out[i] = svc.isValid(st) ? 1 : 0;
    }
    sc->space->freeState(st);
}

// homogeneous2ndOrderCarSystemSVC::areStatesValid (non-composed branch), StateValidityCheckerDatabase.h:108-124: robot 1
// (L1 x W1) at q1[i] against robot 2 (L2 x W2) at q2[i].
void ref_are_states_valid(double L1, double W1, double L2, double W2, int n, const double *q1, const double *q2, int *out)
{
    auto sc = make_scene(32, 32, 0, nullptr, "Robot 1", L1, W1, "Robot 2", L2, W2);
    auto space2 = createBounded2ndOrderCarStateSpace(32, 32);
    space2->setName("Robot 2");
    auto si2 = std::make_shared<oc::SpaceInformation>(space2, createUniform2DRealVectorControlSpace(space2));
    homogeneous2ndOrderCarSystemSVC svc(sc->si, sc->robots, sc->obstacles);
    ob::State *s1 = sc->space->allocState(), *s2 = space2->allocState();
    for (int i = 0; i < n; ++i) {
        set_car(s1, 0, q1 + 5 * i);
        set_car(s2, 0, q2 + 5 * i);
        out[i] = svc.areStatesValid(s1, {si2, s2}) ? 1 : 0;
    }
    sc->space->freeState(s1);
    space2->freeState(s2);
}

// homogeneousTwo2ndOrderCarSystemSVC::isValid, StateValidityCheckerDatabase.h:225-261, for n 10-d states, and
// ::areStatesValid :264-287 against a third robot (L3 x W3) at q3[i] (q3 == NULL: skipped, out_pair untouched).
void ref_pair_is_valid(unsigned x_max, unsigned y_max, int n_obs, const double *obs4, double L1, double W1, double L2, double W2,
                       int n, const double *q10, int *out, double L3, double W3, const double *q3, int *out_pair)
{
    auto sc = make_scene(x_max, y_max, n_obs, obs4, "Robot 1", L1, W1, "Robot 2", L2, W2);
    sc->robot_store.push_back(std::make_unique<RectangularRobot>("Robot 3", L3 > 0 ? L3 : 1.0, W3 > 0 ? W3 : 1.0));
    sc->robots["Robot 3"] = sc->robot_store.back().get();
    auto space = createBoundedTwo2ndOrderCarStateSpace(x_max, y_max);
    auto si = std::make_shared<oc::SpaceInformation>(space, createUniform4DRealVectorControlSpace(space));
    homogeneousTwo2ndOrderCarSystemSVC svc(si, sc->robots.at("Robot 1"), sc->robots.at("Robot 2"), sc->robots, sc->obstacles);
    auto space3 = createBounded2ndOrderCarStateSpace(x_max, y_max);
    space3->setName("Robot 3");
    auto si3 = std::make_shared<oc::SpaceInformation>(space3, createUniform2DRealVectorControlSpace(space3));
    ob::State *st = space->allocState(), *s3 = space3->allocState();
    for (int i = 0; i < n; ++i) {
        set_car(st, 0, q10 + 10 * i);
        set_car(st, 1, q10 + 10 * i + 5);
        out[i] = svc.isValid(st) ? 1 : 0;
        if (q3) {
            set_car(s3, 0, q3 + 5 * i);
            out_pair[i] = svc.areStatesValid(st, {si3, s3}) ? 1 : 0;
        }
    }
    space->freeState(st);
    space3->freeState(s3);
}

// TwoSecondOrderCarStatePropagator::propagate, StatePropagatorDatabase.h:130-153 (the reference's freeze-at-goal rule on
// top of the RK4 stand-in over the reference's TwoSecondOrderCarsODE + its post-integration wrap), `steps` propagation
// steps of 0.1 s each, one call per step as propagateWhileValid does.
void ref_pair_propagate(const double q10[10], const double u4[4], const double goals[4], int steps, double *out /* [steps][10] */)
{
    auto space = createBoundedTwo2ndOrderCarStateSpace(32, 32);
    auto si = std::make_shared<oc::SpaceInformation>(space, createUniform4DRealVectorControlSpace(space));
    TwoSecondOrderCarStatePropagator sp(si, goals[0], goals[1], goals[2], goals[3]);
    ob::State *a = space->allocState(), *b = space->allocState();
    set_car(a, 0, q10);
    set_car(a, 1, q10 + 5);
    Ctl c(u4, 4);
    for (int j = 0; j < steps; ++j) {
        sp.propagate(a, &c, 0.1, b);
        get_car(b, 0, out + 10 * j);
        get_car(b, 1, out + 10 * j + 5);
        space->copyState(a, b);
    }
    space->freeState(a);
    space->freeState(b);
}

// SecondOrderCarODE through the same stand-in + SecondOrderCarODEPostIntegration (:67-74): the single-car propagator of
// demo_*.cpp:111-112, one call per step.
void ref_car_propagate(const double q5[5], const double u2[2], int steps, double *out /* [steps][5] */)
{
    auto space = createBounded2ndOrderCarStateSpace(32, 32);
    auto si = std::make_shared<oc::SpaceInformation>(space, createUniform2DRealVectorControlSpace(space));
    auto solver = std::make_shared<oc::ODEBasicSolver<>>(si, &SecondOrderCarODE);
    auto sp = oc::ODESolver::getStatePropagator(solver, &SecondOrderCarODEPostIntegration);
    ob::State *a = space->allocState(), *b = space->allocState();
    set_car(a, 0, q5);
    Ctl c(u2, 2);
    for (int j = 0; j < steps; ++j) {
        sp->propagate(a, &c, 0.1, b);
        get_car(b, 0, out + 5 * j);
        space->copyState(a, b);
    }
    space->freeState(a);
    space->freeState(b);
}

// GoalRegionTwo2ndOrderCars::isSatisfied(st, &distance), GoalRegionDatabase.h:61-108
int ref_two_car_goal(const double q10[10], const double goals[4], double *distance)
{
    auto space = createBoundedTwo2ndOrderCarStateSpace(32, 32);
    auto si = std::make_shared<ob::SpaceInformation>(space);
    GoalRegionTwo2ndOrderCars goal(si, goals[0], goals[1], goals[2], goals[3]);
    ob::State *st = space->allocState();
    set_car(st, 0, q10);
    set_car(st, 1, q10 + 5);
    double d = 0;
    const bool a = goal.isSatisfied(st, &d), b = goal.isSatisfied(st);
    space->freeState(st);
    if (distance) *distance = d;
    return (a ? 1 : 0) | (b ? 2 : 0);
}

// Geometry constants of the reference's shapes: obstacle centre + bounding radius (Obstacle.h:82-100), robot bounding
// radius (Robot.h:82-99).
void ref_shape_constants(double ox, double oy, double ol, double ow, double L, double W, double out[4])
{
    RectangularObstacle o(ox, oy, ol, ow);
    RectangularRobot r("r", L, W);
    out[0] = o.getCenterPoint()[0];
    out[1] = o.getCenterPoint()[1];
    out[2] = o.getBoundingRadius();
    out[3] = r.getBoundingRadius();
}

// homogeneous2ndOrderCarSystemMerger::merge(index1, index2), SystemMergerDatabase.h:55-168, on a problem of n_robots
// single cars named "Robot 1".."Robot n" (starts / goals given as (x, y) pairs).  Reports the wiring of the result:
//   returns  -1 when merge() returned {nullptr, nullptr}, else the number of individuals of the merged problem
//   names    '|'-separated state-space names of the individuals, in order
//   dims     state-space dimension of every individual
//   lo / hi  the 8 real bounds of the merged individual (two RealVector(4) subspaces), control bounds in clo / chi (4)
//   start    the merged individual's 10 start reals;  steps = (min, max) control duration; step_size
//   remerge  result of merging the merged individual (last) with individual 0 again: 1 when that returns null
int ref_merge(int n_robots, const double *starts_xy, const double *goals_xy, unsigned bound, int index1, int index2, char *names,
              int names_cap, int *dims, double lo[8], double hi[8], double clo[4], double chi[4], double start[10], int steps[2],
              double *step_size, int *remerge)
{
    std::unordered_map<std::string, Robot *> robots;
    std::vector<std::unique_ptr<Robot>> store;
    std::set<Obstacle *> obstacles;
    std::map<std::string, std::pair<double, double>> start_map, goal_map;
    auto ma_si = std::make_shared<omrc::SpaceInformation>();
    auto ma_pdef = std::make_shared<omrb::ProblemDefinition>(ma_si);
    for (int r = 0; r < n_robots; ++r) {
        const std::string name = "Robot " + std::to_string(r + 1);
        store.push_back(std::make_unique<RectangularRobot>(name, 1.0, 1.0));
        robots[name] = store.back().get();
        start_map[name] = {starts_xy[2 * r], starts_xy[2 * r + 1]};
        goal_map[name] = {goals_xy[2 * r], goals_xy[2 * r + 1]};
    }
    for (int r = 0; r < n_robots; ++r) {
        const std::string name = "Robot " + std::to_string(r + 1);
        auto space = createBounded2ndOrderCarStateSpace(bound, bound);
        space->setName(name);
        auto si = std::make_shared<oc::SpaceInformation>(space, createUniform2DRealVectorControlSpace(space));
        si->setStateValidityChecker(std::make_shared<homogeneous2ndOrderCarSystemSVC>(si, robots, obstacles));
        auto pdef = std::make_shared<ob::ProblemDefinition>(si);
        ob::ScopedState<> st(space);
        st[0] = starts_xy[2 * r];
        st[1] = starts_xy[2 * r + 1];
        st[2] = st[3] = st[4] = 0.0;
        pdef->addStartState(st);
        pdef->setGoal(std::make_shared<GoalRegion2ndOrderCar>(si, goals_xy[2 * r], goals_xy[2 * r + 1]));
        ma_si->addIndividual(si);
        ma_pdef->addIndividual(pdef);
    }
    ma_si->lock();
    ma_pdef->lock();
    homogeneous2ndOrderCarSystemMerger merger(ma_si, ma_pdef, robots, obstacles, bound, start_map, goal_map);
    auto res = merger.merge(index1, index2);
    if (!res.first) return -1;
    const int n = (int)res.first->getIndividualCount();
    std::string all;
    for (int i = 0; i < n; ++i) {
        const auto &ind = res.first->getIndividual(i);
        all += (i ? "|" : "") + ind->getStateSpace()->getName();
        dims[i] = (int)ind->getStateSpace()->getDimension();
    }
    std::snprintf(names, (size_t)names_cap, "%s", all.c_str());
    const auto &m = res.first->getIndividual(n - 1);
    const auto *cs = m->getStateSpace()->as<ob::CompoundStateSpace>();
    for (int car = 0; car < 2; ++car) {
        const ob::RealVectorBounds &b = cs->as<ob::RealVectorStateSpace>(2 * car)->getBounds();
        for (int i = 0; i < 4; ++i) {
            lo[4 * car + i] = b.low[i];
            hi[4 * car + i] = b.high[i];
        }
    }
    const ob::RealVectorBounds &cb = m->getControlSpace()->as<oc::RealVectorControlSpace>()->getBounds();
    for (int i = 0; i < 4; ++i) {
        clo[i] = cb.low[i];
        chi[i] = cb.high[i];
    }
    std::vector<double> reals;
    m->getStateSpace()->copyToReals(reals, res.second->getIndividual(n - 1)->getStartState(0));
    for (int i = 0; i < 10; ++i) start[i] = reals[i];
    steps[0] = (int)m->getMinControlDuration();
    steps[1] = (int)m->getMaxControlDuration();
    *step_size = m->getPropagationStepSize();
    // merging the merged individual again is refused (:103-104)
    std::unordered_map<std::string, Robot *> robots2;
    for (auto &kv : robots) robots2[kv.first] = kv.second;
    homogeneous2ndOrderCarSystemMerger again(res.first, res.second, robots, obstacles, bound, start_map, goal_map);
    *remerge = 0;
    try {
        auto r2 = again.merge(n - 1, 0);
        *remerge = r2.first ? 0 : 1;
    } catch (const std::exception &) {
        *remerge = 2;   // (robots_.at on the merged name throws before the dimension test is reached)
    }
    return n;
}

}  // extern "C"
