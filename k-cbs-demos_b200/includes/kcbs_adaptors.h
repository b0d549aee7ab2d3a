// kcbs_adaptors.h -- glue between the OMPL-facing plugin objects (mirror of the reference's includes/*.h) and the
// C ABI of libkcbs_b200.  Everything numeric goes through kcbs_* calls; this file only moves values between OMPL
// objects and flat float buffers, and translates error conventions.
//
//   scalar virtuals   isValid / areStatesValid / propagate  ->  one-element batches (kcbs_validity_batch,
//                     kcbs_first_conflict, kcbs_propagate_batch); they exist for API completeness, the speed-up comes
//                     from the batched entry points below
//   planners          oc::RRT::solve -> kcbs_rrt_plan, omrc::KCBS::solve -> kcbs_kcbs_solve, omrc::PP::solve ->
//                     sequential kcbs_rrt_plan calls with the earlier robots as constraints
#pragma once

#include <array>
#include <chrono>
#include <map>
#include <mutex>
#include <set>
#include <sstream>
#include <unordered_map>

#include <ompl/control/ODESolver.h>
#include <ompl/control/SpaceInformation.h>
#include <ompl/control/planners/rrt/RRT.h>
#include <ompl/multirobot/control/planners/kcbs/KCBS.h>
#include <ompl/multirobot/control/planners/pp/PP.h>

#include "Obstacle.h"
#include "Robot.h"
#include "kcbs_b200.hpp"

namespace kcbs {

namespace ob = ompl::base;
namespace oc = ompl::control;

// What the planners need to know about an individual; implemented by the state validity checkers.
class SceneProvider
{
public:
    virtual ~SceneProvider() = default;
    virtual const std::unordered_map<std::string, Robot *> &sceneRobots() const = 0;
    virtual const std::set<Obstacle *> &sceneObstacles() const = 0;
};

// Goals that are a disc around a position; implemented by GoalRegion2ndOrderCar.
class GoalDisc
{
public:
    virtual ~GoalDisc() = default;
    virtual double goalX() const = 0;
    virtual double goalY() const = 0;
    virtual double goalRadius() const = 0;
};

// ODE functors the GPU integrator knows (the functor itself is only a tag, see StatePropagatorDatabase.h).
enum class OdeModel { Unknown, SecondOrderCar, TwoSecondOrderCars };
using OdeFn = void (*)(const oc::ODESolver::StateType &, const oc::Control *, oc::ODESolver::StateType &);
inline std::map<OdeFn, OdeModel> &ode_registry()
{
    static std::map<OdeFn, OdeModel> r;
    return r;
}
struct OdeRegistrar {
    OdeRegistrar(OdeFn f, OdeModel m) { ode_registry()[f] = m; }
};

// reals of a 2nd-order-car state in OMPL order (x, y, v, phi, theta), StatePropagatorDatabase.h:46
inline void car_reals(const ob::State *state, unsigned int car, float out[5])
{
    const auto *cs = state->as<ob::CompoundStateSpace::StateType>();
    const double *v = cs->as<ob::RealVectorStateSpace::StateType>(2 * car)->values;
    out[0] = (float)v[0]; out[1] = (float)v[1]; out[2] = (float)v[2]; out[3] = (float)v[3];
    out[4] = (float)cs->as<ob::SO2StateSpace::StateType>(2 * car + 1)->value;
}
inline void set_car_reals(ob::State *state, unsigned int car, const float in[5])
{
    auto *cs = state->as<ob::CompoundStateSpace::StateType>();
    double *v = cs->as<ob::RealVectorStateSpace::StateType>(2 * car)->values;
    v[0] = in[0]; v[1] = in[1]; v[2] = in[2]; v[3] = in[3];
    cs->as<ob::SO2StateSpace::StateType>(2 * car + 1)->value = in[4];
}

// Scene of one or more rectangular robots; contexts are cached per distinct scene (they own device memory).
inline SceneDesc make_scene(const ob::StateSpacePtr &space, const std::vector<const Robot *> &robots,
                            const std::set<Obstacle *> &obstacles, double step = 0.1, double int_step = 1e-2)
{
    SceneDesc s;
    const auto &b = space->as<ob::CompoundStateSpace>()->as<ob::RealVectorStateSpace>(0)->getBounds();
    s.x_max = b.high[0];
    s.y_max = b.high[1];
    s.step_size = step;
    s.integration_step = int_step;
    for (const Robot *r : robots) s.robots.push_back({r->getLength(), r->getWidth()});
    for (const Obstacle *o : obstacles) {
        const double *q = o->getRectangle();
        s.obstacles.push_back({q[0], q[1], q[2], q[3]});
    }
    return s;
}

inline std::shared_ptr<Context> cached_context(const SceneDesc &s)
{
    static std::mutex mu;
    static std::map<std::vector<double>, std::shared_ptr<Context>> cache;
    std::vector<double> key = {s.x_max, s.y_max, s.step_size, s.integration_step, s.car_length, (double)s.robots.size()};
    for (const auto &r : s.robots) key.insert(key.end(), r.begin(), r.end());
    for (const auto &o : s.obstacles) key.insert(key.end(), o.begin(), o.end());
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find(key);
    if (it != cache.end()) return it->second;
    auto ctx = std::make_shared<Context>(s, 0);
    cache[key] = ctx;
    return ctx;
}

// ---- scalar calls (one-element batches) -----------------------------------------------------------------------
inline bool gpu_is_valid(const Context &ctx, int robot, const float q[5])
{
    uint32_t mask = 0;
    ctx.isValid(robot, 1, q, &mask);
    return (mask & 1u) != 0;
}

// Robots 0 and 1 of `ctx` at the given (x, y, theta): true when they do NOT overlap.
inline bool gpu_pair_valid(const Context &ctx, const float a[5], const float b[5])
{
    const float traj[6] = {a[0], a[1], a[4], b[0], b[1], b[4]};  // [R = 2][3][T = 1]
    kcbs_conflict c;
    ctx.firstConflict(1, 2, 1, traj, nullptr, &c);
    return c.r1 < 0;
}

// `steps` propagation steps of one car without validity checking (world with unbounded state space)
inline void gpu_propagate(const Context &ctx, const float q[5], const float u[2], int steps, float out[5])
{
    int32_t st = steps, valid = 0;
    ctx.propagateWhileValid(0, 1, 1, q, nullptr, u, &st, steps, nullptr, out, &valid);
}

// StatePropagator returned by oc::ODESolver::getStatePropagator: integrates on the GPU.
class GpuOdePropagator : public oc::StatePropagator
{
public:
    GpuOdePropagator(const oc::SpaceInformationPtr &si, OdeModel model, double int_step, oc::ODESolver::PostPropagationEvent post)
        : oc::StatePropagator(si), model_(model), int_step_(int_step), post_(std::move(post)) {}

    void propagate(const ob::State *state, const oc::Control *control, double duration, ob::State *result) const override
    {
        if (model_ == OdeModel::Unknown)
            throw std::runtime_error("ODEBasicSolver: only SecondOrderCarODE / TwoSecondOrderCarsODE are integrated (on the GPU); "
                                     "there is no CPU integrator in this build");
        // the integrator runs whole propagation steps; an arbitrary duration is one step of that size
        std::shared_ptr<Context> ctx;
        {
            std::lock_guard<std::mutex> lock(mu_);
            auto it = by_duration_.find(duration);
            if (it == by_duration_.end()) it = by_duration_.emplace(duration, make_unbounded(duration)).first;
            ctx = it->second;
        }
        const unsigned int cars = model_ == OdeModel::TwoSecondOrderCars ? 2 : 1;
        const double *u = control->as<oc::RealVectorControlSpace::ControlType>()->values;
        for (unsigned int c = 0; c < cars; ++c) {
            float q[5], out[5];
            car_reals(state, c, q);
            const float uc[2] = {(float)u[2 * c], (float)u[2 * c + 1]};
            gpu_propagate(*ctx, q, uc, 1, out);
            set_car_reals(result, c, out);
        }
        if (post_) post_(state, control, duration, result);
    }
    bool canPropagateBackward() const override { return false; }

private:
    std::shared_ptr<Context> make_unbounded(double duration) const
    {
        // state bounds far outside any workspace: every propagated state is "valid", so one call = one propagate()
        kcbs_world w;
        kcbs_world_defaults(&w, 1e30, 1e30);
        for (int i = 0; i < 4; ++i) { w.lo[i] = -1e30; w.hi[i] = 1e30; }
        w.step_size = duration;
        w.integration_step = int_step_;
        const double dims[2] = {1.0, 1.0};
        w.n_robots = 1;
        w.robot_dims = dims;
        kcbs_ctx *h = nullptr;
        if (kcbs_create(&w, 0, &h) != KCBS_OK) throw std::runtime_error(std::string("kcbs_create: ") + kcbs_last_error(nullptr));
        return std::shared_ptr<Context>(new Context(h, 1));
    }
    OdeModel model_;
    double int_step_;
    oc::ODESolver::PostPropagationEvent post_;
    mutable std::mutex mu_;
    mutable std::map<double, std::shared_ptr<Context>> by_duration_;
};

// ---- planner plumbing -------------------------------------------------------------------------------------------
struct Individual {
    std::string name;
    const Robot *robot = nullptr;
    float start[5] = {0, 0, 0, 0, 0};
    float goal[2] = {0, 0};
    float goal_radius = 0.5f;
};

// Collects scene + starts + goals of a set of individuals (addIndividual order) and opens a context for them.
inline std::shared_ptr<Context> open_problem(const std::vector<oc::SpaceInformationPtr> &sis,
                                             const std::vector<ob::ProblemDefinitionPtr> &pdefs, std::vector<Individual> &out)
{
    if (sis.empty()) throw std::runtime_error("no individuals");
    std::vector<const Robot *> robots;
    const SceneProvider *scene = nullptr;
    out.clear();
    for (size_t i = 0; i < sis.size(); ++i) {
        if (sis[i]->getStateSpace()->getDimension() != 5)
            throw std::runtime_error("the GPU planners handle single 2nd-order cars (merged individuals are not supported)");
        const auto *sp = dynamic_cast<const SceneProvider *>(sis[i]->getStateValidityChecker().get());
        if (!sp) throw std::runtime_error("the state validity checker must be a homogeneous2ndOrderCarSystemSVC");
        if (!scene) scene = sp;
        Individual ind;
        ind.name = sis[i]->getStateSpace()->getName();
        ind.robot = sp->sceneRobots().at(ind.name);  // throws std::out_of_range like the reference's robots_.at
        car_reals(pdefs[i]->getStartState(0), 0, ind.start);
        const auto *disc = dynamic_cast<const GoalDisc *>(pdefs[i]->getGoal().get());
        if (!disc) throw std::runtime_error("the goal must be a GoalRegion2ndOrderCar");
        ind.goal[0] = (float)disc->goalX();
        ind.goal[1] = (float)disc->goalY();
        ind.goal_radius = (float)disc->goalRadius();
        robots.push_back(ind.robot);
        out.push_back(ind);
    }
    const double step = sis[0]->getPropagationStepSize();
    return cached_context(make_scene(sis[0]->getStateSpace(), robots, scene->sceneObstacles(), step));
}

// [5][T] planes with stride `stride` -> PathControl with one state per propagation step
inline std::shared_ptr<oc::PathControl> make_path(const oc::SpaceInformationPtr &si, const float *traj, int stride, int T)
{
    auto path = std::make_shared<oc::PathControl>(si);
    ob::State *s = si->allocState();
    oc::Control *c = si->allocControl();
    const double dt = si->getPropagationStepSize();
    for (int k = 0; k < T; ++k) {
        const float q[5] = {traj[0 * stride + k], traj[1 * stride + k], traj[2 * stride + k], traj[3 * stride + k], traj[4 * stride + k]};
        set_car_reals(s, 0, q);
        if (k + 1 < T) {
            double *u = c->as<oc::RealVectorControlSpace::ControlType>()->values;
            u[0] = (traj[2 * stride + k + 1] - q[2]) / dt;  // a = dv/dt, phidot = dphi/dt (both exactly linear)
            u[1] = (traj[3 * stride + k + 1] - q[3]) / dt;
            path->append(s, c, dt);
        } else {
            path->append(s);
        }
    }
    si->freeState(s);
    si->freeControl(c);
    return path;
}

constexpr int kMaxPlanSteps = 8192;

}  // namespace kcbs

// ---- hooks declared by the OMPL shim ----------------------------------------------------------------------------
#ifdef KCBS_OMPL_SHIM
namespace ompl {
namespace control {

inline StatePropagatorPtr ODESolver::getStatePropagator(std::shared_ptr<ODESolver> solver, const PostPropagationEvent &postEvent)
{
    kcbs::OdeModel model = kcbs::OdeModel::Unknown;
    if (auto *fp = solver->getODE().target<kcbs::OdeFn>()) {
        auto it = kcbs::ode_registry().find(*fp);
        if (it != kcbs::ode_registry().end()) model = it->second;
    }
    return std::make_shared<kcbs::GpuOdePropagator>(solver->getSpaceInformation(), model, solver->getIntegrationStepSize(), postEvent);
}

inline base::PlannerStatus RRT::solve(double solveTime)
{
    std::vector<kcbs::Individual> ind;
    auto siC = std::static_pointer_cast<SpaceInformation>(si_);
    auto ctx = kcbs::open_problem({siC}, {pdef_}, ind);
    kcbs_rrt_params p;
    kcbs_rrt_defaults(&p);
    p.min_steps = (int)siC->getMinControlDuration();
    p.max_steps = (int)siC->getMaxControlDuration();
    p.goal_bias = (float)goalBias_;
    p.goal_radius = ind[0].goal_radius;
    p.time_limit_s = solveTime;
    std::vector<float> traj((size_t)5 * kcbs::kMaxPlanSteps);
    int32_t T = 0;
    kcbs_rrt_stats st;
    ctx->check(kcbs_rrt_plan(ctx->get(), 0, ind[0].start, ind[0].goal, &p, 0, nullptr, nullptr, nullptr, nullptr, nullptr,
                             kcbs::kMaxPlanSteps, traj.data(), &T, &st));
    if (!st.solved) return base::PlannerStatus::TIMEOUT;
    pdef_->addSolutionPath(kcbs::make_path(siC, traj.data(), kcbs::kMaxPlanSteps, T));
    return base::PlannerStatus::EXACT_SOLUTION;
}

}  // namespace control

namespace multirobot {
namespace control {

inline ompl::base::PlannerStatus KCBS::solve(double solveTime)
{
    const unsigned int R = si_->getIndividualCount();
    std::vector<ompl::control::SpaceInformationPtr> sis;
    std::vector<ompl::base::ProblemDefinitionPtr> pdefs;
    for (unsigned int r = 0; r < R; ++r) {
        sis.push_back(si_->getIndividual(r));
        pdefs.push_back(pdef_->getIndividual(r));
    }
    std::vector<kcbs::Individual> ind;
    auto ctx = kcbs::open_problem(sis, pdefs, ind);
    kcbs_kcbs_params p;
    kcbs_kcbs_defaults(&p);
    p.time_limit_s = solveTime;
    p.merge_bound = mergeBound_;
    p.low_level.time_limit_s = llSolveTime_;
    p.low_level.min_steps = (int)sis[0]->getMinControlDuration();
    p.low_level.max_steps = (int)sis[0]->getMaxControlDuration();
    p.low_level.goal_radius = ind[0].goal_radius;
    std::vector<float> starts((size_t)5 * R), goals((size_t)2 * R);
    for (unsigned int r = 0; r < R; ++r) {
        for (int c = 0; c < 5; ++c) starts[(size_t)c * R + r] = ind[r].start[c];
        goals[2 * r] = ind[r].goal[0];
        goals[2 * r + 1] = ind[r].goal[1];
    }
    const int max_T = kcbs::kMaxPlanSteps;
    std::vector<float> plan((size_t)R * 5 * max_T);
    std::vector<int32_t> lengths(R);
    kcbs_kcbs_stats st;
    ctx->check(kcbs_kcbs_solve(ctx->get(), (int)R, starts.data(), goals.data(), &p, max_T, plan.data(), lengths.data(), &st));
    rootSolveTime_ = st.root_seconds;
    nodesExpanded_ = (unsigned int)st.nodes_expanded;
    approximateSolutions_ = (unsigned int)st.approximate_solutions;
    std::ostringstream log;
    log << "K-CBS constraint tree: " << st.nodes_expanded << " nodes expanded, " << st.low_level_calls << " low-level plans, "
        << st.conflict_checks << " plans validated, " << st.approximate_solutions << " approximate solutions, " << st.merges
        << " merges, root "
        << st.root_seconds << " s, total " << st.seconds << " s" << std::endl;
    treeLog_ = log.str();
    if (!st.solved) return ompl::base::PlannerStatus::TIMEOUT;
    auto result = std::make_shared<PlanControl>();
    for (unsigned int r = 0; r < R; ++r) {
        auto path = kcbs::make_path(sis[r], plan.data() + (size_t)r * 5 * max_T, max_T, lengths[r]);
        pdefs[r]->addSolutionPath(path);
        result->append(path);
    }
    pdef_->addSolutionPlan(result);
    return ompl::base::PlannerStatus::EXACT_SOLUTION;
}

// Prioritized planning: robots are planned in index order, each against the full (padded) paths of its
// predecessors as dynamic obstacles.
inline ompl::base::PlannerStatus PP::solve(double solveTime)
{
    const auto t0 = std::chrono::steady_clock::now();
    const unsigned int R = si_->getIndividualCount();
    std::vector<ompl::control::SpaceInformationPtr> sis;
    std::vector<ompl::base::ProblemDefinitionPtr> pdefs;
    for (unsigned int r = 0; r < R; ++r) {
        sis.push_back(si_->getIndividual(r));
        pdefs.push_back(pdef_->getIndividual(r));
    }
    std::vector<kcbs::Individual> ind;
    auto ctx = kcbs::open_problem(sis, pdefs, ind);
    const int max_T = kcbs::kMaxPlanSteps;
    std::vector<std::vector<float>> trajs(R, std::vector<float>((size_t)5 * max_T));
    std::vector<int32_t> lengths(R, 0);
    for (unsigned int r = 0; r < R; ++r) {
        const double left = solveTime - std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if (left <= 0) return ompl::base::PlannerStatus::TIMEOUT;
        kcbs_rrt_params p;
        kcbs_rrt_defaults(&p);
        p.time_limit_s = left;
        p.seed = 77 + r;
        p.goal_radius = ind[r].goal_radius;
        // constraints: every earlier robot's whole path as a PERSISTENT constraint (negative length: its last state --
        // the robot parked at its goal -- keeps holding for every later time index, however long this robot's path gets)
        std::vector<int32_t> c_robot, c_k0, c_len, c_off;
        size_t total = 0;
        for (unsigned int q = 0; q < r; ++q) total += (size_t)lengths[q];
        std::vector<float> states((size_t)3 * std::max<size_t>(1, total));
        size_t at = 0;
        for (unsigned int q = 0; q < r; ++q) {
            c_robot.push_back((int)q); c_k0.push_back(0); c_len.push_back(-lengths[q]); c_off.push_back((int)at);
            for (int k = 0; k < lengths[q]; ++k) {
                states[at + k] = trajs[q][(size_t)0 * max_T + k];
                states[total + at + k] = trajs[q][(size_t)1 * max_T + k];
                states[2 * total + at + k] = trajs[q][(size_t)4 * max_T + k];
            }
            at += (size_t)lengths[q];
        }
        kcbs_rrt_stats st;
        ctx->check(kcbs_rrt_plan(ctx->get(), (int)r, ind[r].start, ind[r].goal, &p, (int)r, c_robot.data(), c_k0.data(), c_len.data(),
                                 c_off.data(), states.data(), max_T, trajs[r].data(), &lengths[r], &st));
        if (!st.solved) return ompl::base::PlannerStatus::TIMEOUT;
    }
    {
        // the finished plan is validated as a whole before it is reported (plan scan with the padding rule)
        int T = 1;
        for (unsigned int r = 0; r < R; ++r) T = std::max(T, (int)lengths[r]);
        const int Tp = (T + 3) & ~3;
        std::vector<float> traj3((size_t)R * 3 * Tp, 0.0f);
        for (unsigned int r = 0; r < R; ++r)
            for (int k = 0; k < lengths[r]; ++k) {
                traj3[((size_t)r * 3 + 0) * Tp + k] = trajs[r][(size_t)0 * max_T + k];
                traj3[((size_t)r * 3 + 1) * Tp + k] = trajs[r][(size_t)1 * max_T + k];
                traj3[((size_t)r * 3 + 2) * Tp + k] = trajs[r][(size_t)4 * max_T + k];
            }
        kcbs_conflict cf;
        ctx->check(kcbs_first_conflict(ctx->get(), 1, (int)R, Tp, traj3.data(), lengths.data(), &cf));
        if (cf.r1 >= 0) return ompl::base::PlannerStatus::APPROXIMATE_SOLUTION;
    }
    auto result = std::make_shared<PlanControl>();
    for (unsigned int r = 0; r < R; ++r) {
        auto path = kcbs::make_path(sis[r], trajs[r].data(), max_T, lengths[r]);
        pdefs[r]->addSolutionPath(path);
        result->append(path);
    }
    pdef_->addSolutionPlan(result);
    return ompl::base::PlannerStatus::EXACT_SOLUTION;
}

}  // namespace control
}  // namespace multirobot
}  // namespace ompl
#endif  // KCBS_OMPL_SHIM
