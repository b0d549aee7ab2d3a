// test_mirror.cpp -- exercises the C++ host-side mirror of the reference's plugin headers end to end on a GPU:
// the scalar OMPL virtuals, the batched extras, the planners and the Benchmark driver.  Hand cases are the same as
// the oracle's known-answer tests (tests/test_oracle_kat.py).  Exit code 0 = all checks passed.
#include "Robot.h"
#include "Obstacle.h"
#include "StateSpaceDatabase.h"
#include "ControlSpaceDatabase.h"
#include "StateValidityCheckerDatabase.h"
#include "StatePropagatorDatabase.h"
#include "GoalRegionDatabase.h"
#include "PlannerAllocatorDatabase.h"
#include "SystemMergerDatabase.h"
#include "Benchmark.h"

#include <array>
#include <cstdio>

static int failures = 0;
#define CHECK(cond)                                                        \
    do {                                                                   \
        if (!(cond)) {                                                     \
            std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond);  \
            ++failures;                                                    \
        }                                                                  \
    } while (0)

struct Problem {
    std::map<std::string, std::pair<double, double>> start_map, goal_map;
    std::set<Obstacle *> obs_set;
    std::unordered_map<std::string, Robot *> robot_map;
    omrc::SpaceInformationPtr ma_si;
    omrb::ProblemDefinitionPtr ma_pdef;
};

// Builds the multi-robot problem exactly like the reference mains do (src/demo_Corridor10x10_4robots_dyn2ndOrderCars.cpp:61-146)
static Problem make_problem(std::map<std::string, std::pair<double, double>> starts, std::map<std::string, std::pair<double, double>> goals,
                            std::vector<std::array<double, 4>> obstacles)
{
    Problem p;
    p.start_map = std::move(starts);
    p.goal_map = std::move(goals);
    for (const auto &o : obstacles) p.obs_set.insert(new RectangularObstacle(o[0], o[1], o[2], o[3]));
    for (auto &kv : p.start_map) p.robot_map[kv.first] = new RectangularRobot(kv.first, 1.0, 1.0);
    p.ma_si = std::make_shared<omrc::SpaceInformation>();
    p.ma_pdef = std::make_shared<omrb::ProblemDefinition>(p.ma_si);
    for (auto &kv : p.start_map) {
        auto space = createBounded2ndOrderCarStateSpace(10, 10);
        space->setName(kv.first);
        auto cspace = createUniform2DRealVectorControlSpace(space);
        auto si = std::make_shared<oc::SpaceInformation>(space, cspace);
        si->setStateValidityChecker(std::make_shared<homogeneous2ndOrderCarSystemSVC>(si, p.robot_map, p.obs_set));
        auto odeSolver = std::make_shared<oc::ODEBasicSolver<>>(si, &SecondOrderCarODE);
        si->setStatePropagator(oc::ODESolver::getStatePropagator(odeSolver, &SecondOrderCarODEPostIntegration));
        si->setPropagationStepSize(0.1);
        si->setMinMaxControlDuration(1, 10);
        ob::ScopedState<> start(space);
        start[0] = kv.second.first; start[1] = kv.second.second; start[2] = 0.0; start[3] = 0.0; start[4] = 0.0;
        auto pdef = std::make_shared<ob::ProblemDefinition>(si);
        pdef->addStartState(start);
        pdef->setGoal(std::make_shared<GoalRegion2ndOrderCar>(si, p.goal_map.at(kv.first).first, p.goal_map.at(kv.first).second));
        p.ma_si->addIndividual(si);
        p.ma_pdef->addIndividual(pdef);
    }
    ompl::base::PlannerAllocator allocator = &allocateControlRRT;
    p.ma_si->setPlannerAllocator(allocator);
    p.ma_si->lock();
    p.ma_pdef->lock();
    return p;
}

// Corridor10x10 with 4 robots (src/demo_Corridor10x10_4robots_dyn2ndOrderCars.cpp:61-78)
static Problem corridor()
{
    return make_problem({{"Robot 1", {1.0, 0.5}}, {"Robot 2", {1.0, 3.5}}, {"Robot 3", {9.0, 0.5}}, {"Robot 4", {9.0, 3.5}}},
                        {{"Robot 1", {9.0, 0.5}}, {"Robot 2", {9.0, 9.0}}, {"Robot 3", {1.0, 0.5}}, {"Robot 4", {1.0, 9.0}}},
                        {{0.5, 1.8, 9.0, 1.0}, {4.0, 5.0, 2.0, 2.0}});
}

// three robots crossing an open 10 x 10 workspace: easy enough for prioritized planning
static Problem crossing()
{
    return make_problem({{"Robot 1", {1.0, 1.0}}, {"Robot 2", {9.0, 1.0}}, {"Robot 3", {5.0, 9.0}}},
                        {{"Robot 1", {9.0, 9.0}}, {"Robot 2", {1.0, 9.0}}, {"Robot 3", {5.0, 1.0}}}, {});
}

static void set_state(ob::ScopedState<> &s, double x, double y, double v, double phi, double th)
{
    s[0] = x; s[1] = y; s[2] = v; s[3] = phi; s[4] = th;
}

int main()
{
    Problem corridor_problem = corridor();
    Problem &p = corridor_problem;
    auto si = p.ma_si->getIndividual(0);
    auto space = si->getStateSpace();
    CHECK(space->getDimension() == 5 && space->getName() == "Robot 1");

    // ---- geometry objects (Robot.h / Obstacle.h) ---------------------------------------------------------------
    RectangularRobot rr("r", 0.7, 0.5);
    CHECK(std::fabs(rr.getBoundingRadius() - std::hypot(0.35, 0.25)) < 1e-15 && rr.getShape().outer().size() == 5);
    RectangularObstacle ro(0.5, 1.8, 9.0, 1.0);
    CHECK(ro.getCenterPoint()[0] == 5.0 && ro.getCenterPoint()[1] == 2.3);

    // ---- isValid (scalar virtual -> kcbs_validity_batch) ---------------------------------------------------------
    ob::ScopedState<> s(space);
    auto valid = [&](double x, double y, double v, double phi, double th) {
        set_state(s, x, y, v, phi, th);
        return si->isValid(s.get());
    };
    CHECK(valid(1.0, 0.5, 0, 0, 0));                    // demo start state
    CHECK(!valid(5.0, 2.3, 0, 0, 0));                   // centre inside the long obstacle
    CHECK(!valid(5.0, 1.3, 0, 0, 0));                   // touching counts as collision
    CHECK(valid(5.0, 1.0, 0, 0, 0));
    CHECK(!valid(5.0, 1.2, 0, 0, M_PI / 4));            // rotated footprint reaches the obstacle
    CHECK(valid(5.0, 1.05, 0, 0, M_PI / 4));
    CHECK(!valid(-0.001, 1.0, 0, 0, 0) && !valid(1.0, 1.0, 1.001, 0, 0) && !valid(1.0, 1.0, 0, 1.06, 0));
    CHECK(valid(0.1, 0.1, 0, 0, 0.3));                  // the footprint may leave the workspace, the centre may not

    // batched extra: 64 states at once
    {
        const int n = 64;
        std::vector<float> st(5 * n, 0.0f);
        for (int i = 0; i < n; ++i) { st[i] = 0.1f + 0.15f * i; st[n + i] = 1.0f; }
        uint32_t mask[2] = {0, 0};
        static_cast<homogeneous2ndOrderCarSystemSVC *>(si->getStateValidityChecker().get())->isValidBatch(n, st.data(), mask);
        for (int i = 0; i < n; ++i) CHECK((((mask[i >> 5] >> (i & 31)) & 1u) != 0) == valid(st[i], 1.0, 0, 0, 0));
    }

    // ---- areStatesValid (scalar virtual -> kcbs_first_conflict) --------------------------------------------------
    auto si2 = p.ma_si->getIndividual(1);
    ob::ScopedState<> a(space), b(si2->getStateSpace());
    auto pair_ok = [&](double x1, double y1, double t1, double x2, double y2, double t2) {
        set_state(a, x1, y1, 0, 0, t1);
        set_state(b, x2, y2, 0, 0, t2);
        return si->getStateValidityChecker()->areStatesValid(a.get(), {si2, b.get()});
    };
    CHECK(pair_ok(1.0, 0.5, 0, 1.0, 3.5, 0));           // demo start states
    CHECK(!pair_ok(2.0, 5.0, 0, 2.5, 5.0, 0));          // overlapping
    CHECK(!pair_ok(2.0, 5.0, 0, 3.0, 5.0, 0));          // touching
    CHECK(pair_ok(2.0, 5.0, 0, 3.001, 5.0, 0));
    CHECK(!pair_ok(2.0, 5.0, M_PI / 4, 3.2, 5.0, M_PI / 4));  // corners of the rotated squares meet
    {
        auto stranger_space = createBounded2ndOrderCarStateSpace(10, 10);
        stranger_space->setName("Robot 99");
        auto cs = createUniform2DRealVectorControlSpace(stranger_space);
        auto stranger = std::make_shared<oc::SpaceInformation>(stranger_space, cs);
        bool threw = false;
        try {
            si->getStateValidityChecker()->areStatesValid(a.get(), {stranger, b.get()});
        } catch (const std::out_of_range &) {
            threw = true;
        }
        CHECK(threw);  // robots_.at(name) of the reference
    }

    // ---- propagate / propagateWhileValid (scalar virtuals -> kcbs_propagate_batch) ---------------------------------
    {
        auto *c = si->allocControl();
        double *u = c->as<oc::RealVectorControlSpace::ControlType>()->values;
        ob::ScopedState<> q(space), r(space);
        // straight line: phi = 0, u = (a, 0): x = x0 + cos(th) (v0 t + a t^2 / 2), RK4 is exact for it
        set_state(q, 3.0, 4.0, 0.2, 0.0, 0.6);
        u[0] = 0.5; u[1] = 0.0;
        si->getStatePropagator()->propagate(q.get(), c, 0.1, r.get());
        const double sdist = 0.2 * 0.1 + 0.5 * 0.5 * 0.01;
        CHECK(std::fabs(r[0] - (3.0 + std::cos(0.6) * sdist)) < 2e-6 && std::fabs(r[1] - (4.0 + std::sin(0.6) * sdist)) < 2e-6);
        CHECK(std::fabs(r[2] - 0.25) < 1e-6 && r[3] == 0.0 && std::fabs(r[4] - 0.6) < 1e-6);
        si->propagate(q.get(), c, 10, r.get());   // 10 steps in one call
        CHECK(std::fabs(r[2] - 0.7) < 1e-6);
        // heading wraps into [-pi, pi)
        set_state(q, 5.0, 5.0, 1.0, 1.0, M_PI - 0.05);
        u[0] = 0.0;
        si->getStatePropagator()->propagate(q.get(), c, 0.1, r.get());
        CHECK(r[4] < 0 && r[4] >= -M_PI);
        // drives into the obstacle above: stops at the last valid state
        set_state(q, 5.0, 0.9, 1.0, 0.0, M_PI / 2);
        const unsigned int done = si->propagateWhileValid(q.get(), c, 10, r.get());
        CHECK(done > 0 && done < 10 && si->isValid(r.get()));
        CHECK(!si->getStatePropagator()->canPropagateBackward());
        bool threw = false;
        try {
            std::vector<double> qq(5, 0.0), qd;
            SecondOrderCarODE(qq, c, qd);
        } catch (const std::logic_error &) {
            threw = true;
        }
        CHECK(threw);  // the functor is a model tag: no CPU integrator exists
        si->freeControl(c);
    }

    // ---- goal region --------------------------------------------------------------------------------------------
    {
        auto goal = p.ma_pdef->getIndividual(0)->getGoal();
        set_state(s, 8.7, 0.5, 0, 0, 0);
        CHECK(goal->isSatisfied(s.get()));
        set_state(s, 8.4, 0.5, 0, 0, 0);
        CHECK(!goal->isSatisfied(s.get()));
        CHECK(std::fabs(goal->as<ob::GoalRegion>()->distanceGoal(s.get()) - 0.6) < 1e-12);
    }

    // ---- low-level planner through the allocator ---------------------------------------------------------------------
    {
        auto planner = p.ma_si->getPlannerAllocator()(si);
        planner->setProblemDefinition(p.ma_pdef->getIndividual(0));
        CHECK((bool)planner->solve(5.0));
        auto path = p.ma_pdef->getIndividual(0)->getSolutionPath()->as<oc::PathControl>();
        CHECK(path->getStateCount() > 10);
        CHECK(p.ma_pdef->getIndividual(0)->getGoal()->isSatisfied(path->getStates().back()));
        for (auto *st : path->getStates()) CHECK(si->isValid(st));
        p.ma_pdef->getIndividual(0)->clearSolutionPaths();
    }

    // ---- K-CBS and PP ----------------------------------------------------------------------------------------------
    Problem open_space = crossing();
    for (const char *name : {"K-CBS", "PP"}) {
        // K-CBS on the corridor scene; prioritized planning (incomplete by nature) on the open crossing scene
        Problem &p = std::string(name) == "K-CBS" ? corridor_problem : open_space;
        const unsigned int R = p.ma_si->getIndividualCount();
        omrb::PlannerPtr planner;
        if (std::string(name) == "K-CBS") {
            omrc::SystemMergerPtr merger = std::make_shared<homogeneous2ndOrderCarSystemMerger>(p.ma_si, p.ma_pdef, p.robot_map, p.obs_set, 10,
                                                                                                p.start_map, p.goal_map);
            p.ma_si->setSystemMerger(merger);
            auto k = std::make_shared<omrc::KCBS>(p.ma_si);
            k->setLowLevelSolveTime(5.);
            k->setMergeBound(10);
            planner = k;
        } else {
            planner = std::make_shared<omrc::PP>(p.ma_si);
        }
        planner->setProblemDefinition(p.ma_pdef);
        const bool solved = (bool)planner->as<omrb::Planner>()->solve(120.0);
        CHECK(solved);
        if (solved) {
            auto plan = p.ma_pdef->getSolutionPlan()->as<omrc::PlanControl>();
            CHECK(plan->getPathCount() == R);
            // pairwise validity of the plan through the scalar virtual, with the padding rule
            size_t T = 0;
            for (unsigned int r = 0; r < R; ++r) T = std::max(T, plan->getPath(r)->getStateCount());
            int conflicts = 0;
            for (size_t k = 0; k < T; ++k)
                for (unsigned int r1 = 0; r1 < R; ++r1)
                    for (unsigned int r2 = r1 + 1; r2 < R; ++r2) {
                        auto &p1 = plan->getPath(r1)->getStates();
                        auto &p2 = plan->getPath(r2)->getStates();
                        const ob::State *s1 = p1[std::min(k, p1.size() - 1)], *s2 = p2[std::min(k, p2.size() - 1)];
                        if (!p.ma_si->getIndividual(r1)->getStateValidityChecker()->areStatesValid(s1, {p.ma_si->getIndividual(r2), s2})) ++conflicts;
                    }
            CHECK(conflicts == 0);
            std::ofstream out(std::string("plan_") + name + ".txt");
            plan->printAsMatrix(out, "Robot");
            std::printf("%s solved its %u-robot problem: %zu steps\n", name, R, T);
        }
        if (std::string(name) == "K-CBS") {
            auto *k = planner->as<omrc::KCBS>();
            CHECK(k->getNumberOfNodesExpanded() >= 1 && k->getRootSolveTime() > 0);
            k->printConstraintTree(std::cout);
        }
        p.ma_pdef->clearSolutionPaths();
    }

    // ---- Benchmark driver ---------------------------------------------------------------------------------------------
    {
        std::remove("mirror-benchmark.csv");
        Benchmark bm;
        bm.setSpaceInformation(p.ma_si);
        bm.setPoroblemDefinition(p.ma_pdef);
        bm.setSolveTime(120.0);
        bm.setKCBSMergeBound(10);
        bm.setFileName("mirror-benchmark");
        bm.runKCBS();
        std::ifstream csv("mirror-benchmark.csv");
        std::string header, row;
        std::getline(csv, header);
        std::getline(csv, row);
        CHECK(header.find("K-CBS Computation Times (s)") == 0 && header.find("K-CBS Success (Bool)") != std::string::npos);
        CHECK(!row.empty());
    }

    // ---- system merger ------------------------------------------------------------------------------------------------
    {
        auto merged = p.ma_si->getSystemMerger()->merge(0, 1);
        CHECK(merged.first != nullptr && merged.second != nullptr);
        if (merged.first) {
            CHECK(merged.first->getIndividualCount() == 3);
            auto msi = merged.first->getIndividual(2);
            CHECK(msi->getStateSpace()->getDimension() == 10 && msi->getStateSpace()->getName() == "Robot 1 and Robot 2");
            ob::ScopedState<> m(msi->getStateSpace());
            for (unsigned int i = 0; i < 10; ++i) m[i] = 0.0;
            m[0] = 1.0; m[1] = 0.5; m[5] = 1.0; m[6] = 3.5;
            CHECK(msi->isValid(m.get()));                       // the two start states
            m[5] = 1.5; m[6] = 0.5;
            CHECK(!msi->isValid(m.get()));                      // the two cars overlap each other
            m[5] = 5.0; m[6] = 2.3;
            CHECK(!msi->isValid(m.get()));                      // second car inside the obstacle
            // a third robot against the merged pair (composite branch of areStatesValid)
            auto si3 = merged.first->getIndividual(0);          // Robot 3
            ob::ScopedState<> q3(si3->getStateSpace());
            set_state(q3, 1.2, 0.6, 0, 0, 0);
            m[5] = 1.0; m[6] = 3.5;
            CHECK(!si3->getStateValidityChecker()->areStatesValid(q3.get(), {msi, m.get()}));
            CHECK(!msi->getStateValidityChecker()->areStatesValid(m.get(), {si3, q3.get()}));
            set_state(q3, 9.0, 0.5, 0, 0, 0);
            CHECK(si3->getStateValidityChecker()->areStatesValid(q3.get(), {msi, m.get()}));
            // merged propagator: car 1 parked inside its goal region stays put, car 2 moves
            auto *c = msi->allocControl();
            double *u = c->as<oc::RealVectorControlSpace::ControlType>()->values;
            u[0] = 1.0; u[1] = 0.0; u[2] = 1.0; u[3] = 0.0;
            for (unsigned int i = 0; i < 10; ++i) m[i] = 0.0;
            m[0] = 8.5; m[1] = 0.5; m[2] = 0.5;                  // within 1.5 of Robot 1's goal (9, 0.5)
            m[5] = 1.0; m[6] = 3.5; m[7] = 0.5;
            ob::ScopedState<> out(msi->getStateSpace());
            msi->getStatePropagator()->propagate(m.get(), c, 0.1, out.get());
            CHECK(out[0] == 8.5 && out[2] == 0.5);
            CHECK(out[5] > 1.04 && std::fabs(out[7] - 0.6) < 1e-6);
            msi->freeControl(c);
            // merging an already merged individual is refused
            auto again = std::make_shared<homogeneous2ndOrderCarSystemMerger>(merged.first, merged.second, p.robot_map, p.obs_set, 10, p.start_map,
                                                                              p.goal_map);
            bool refused = false;
            try {
                auto r = again->merge(0, 2);
                refused = r.first == nullptr && r.second == nullptr;
            } catch (const std::out_of_range &) {
                refused = true;  // the reference's robots_.at on the compound name
            }
            CHECK(refused);
        }
    }

    std::printf(failures ? "test_mirror: %d check(s) FAILED\n" : "test_mirror: all checks passed\n", failures);
    return failures ? 1 : 0;
}
