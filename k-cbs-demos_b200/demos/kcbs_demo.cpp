// kcbs_demo.cpp -- one driver for every scene of the reference's src/demo_*.cpp: builds the per-robot
// oc::SpaceInformation objects, plugs the plugin classes of includes/*.h in, runs K-CBS (or PP) and writes plan.txt /
// tree.txt, exactly the call sequence of the reference mains (src/demo_Corridor10x10_4robots_dyn2ndOrderCars.cpp:58-194)
// but with the scene chosen on the command line:
//     kcbs_demo <scene> [K-CBS|PP] [solve_time_s] [--benchmark]
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
#include "scenes_generated.h"

#include <chrono>
#include <cstdio>
#include <fstream>

int main(int argc, char **argv)
{
    const std::string scene_name = argc > 1 ? argv[1] : "Corridor10x10_4robots";
    const std::string planner_name = argc > 2 ? argv[2] : "K-CBS";
    const kcbs_demo::SceneSpec *scene = nullptr;
    for (const auto &s : kcbs_demo::scenes())
        if (s.name == scene_name) scene = &s;
    if (!scene) {
        std::printf("unknown scene %s; known scenes:\n", scene_name.c_str());
        for (const auto &s : kcbs_demo::scenes()) std::printf("  %s\n", s.name.c_str());
        return 2;
    }
    const double solve_time = argc > 3 ? std::atof(argv[3]) : scene->solve_time;
    const bool benchmark = argc > 4 && std::string(argv[4]) == "--benchmark";

    std::map<std::string, std::pair<double, double>> start_map, goal_map;
    std::set<Obstacle *> obs_set;
    for (const auto &o : scene->obstacles) obs_set.insert(new RectangularObstacle(o.x, o.y, o.length, o.width));
    std::unordered_map<std::string, Robot *> robot_map;
    for (const auto &r : scene->robots) {
        start_map[r.name] = {r.sx, r.sy};
        goal_map[r.name] = {r.gx, r.gy};
        robot_map[r.name] = new RectangularRobot(r.name, scene->robot_length, scene->robot_width);
    }

    auto ma_si = std::make_shared<omrc::SpaceInformation>();
    auto ma_pdef = std::make_shared<omrb::ProblemDefinition>(ma_si);
    for (const auto &r : scene->robots) {
        auto space = createBounded2ndOrderCarStateSpace(scene->x_max, scene->y_max);
        space->setName(r.name);
        auto cspace = createUniform2DRealVectorControlSpace(space);
        auto si = std::make_shared<oc::SpaceInformation>(space, cspace);
        si->setStateValidityChecker(std::make_shared<homogeneous2ndOrderCarSystemSVC>(si, robot_map, obs_set));
        auto odeSolver = std::make_shared<oc::ODEBasicSolver<>>(si, &SecondOrderCarODE);
        si->setStatePropagator(oc::ODESolver::getStatePropagator(odeSolver, &SecondOrderCarODEPostIntegration));
        si->setPropagationStepSize(0.1);
        si->setMinMaxControlDuration(1, 10);
        ob::ScopedState<> start(space);
        start[0] = r.sx; start[1] = r.sy; start[2] = 0.0; start[3] = 0.0; start[4] = 0.0;
        auto pdef = std::make_shared<ob::ProblemDefinition>(si);
        pdef->addStartState(start);
        pdef->setGoal(std::make_shared<GoalRegion2ndOrderCar>(si, r.gx, r.gy));
        ma_si->addIndividual(si);
        ma_pdef->addIndividual(pdef);
    }
    ompl::base::PlannerAllocator allocator = &allocateControlRRT;
    ma_si->setPlannerAllocator(allocator);
    ma_si->lock();
    ma_pdef->lock();

    if (benchmark) {
        Benchmark b;
        b.setSpaceInformation(ma_si);
        b.setPoroblemDefinition(ma_pdef);
        b.setSolveTime(solve_time);
        if (scene->merge_bound >= 0) b.setKCBSMergeBound(scene->merge_bound);
        b.setFileName("KCBS-benchmark-" + scene->name);
        if (planner_name == "K-CBS") b.runKCBS(); else b.runPP();
        return 0;
    }

    omrb::PlannerPtr planner;
    if (planner_name == "K-CBS") {
        if (scene->merge_bound >= 0) {
            omrc::SystemMergerPtr merger = std::make_shared<homogeneous2ndOrderCarSystemMerger>(ma_si, ma_pdef, robot_map, obs_set, scene->x_max,
                                                                                                start_map, goal_map);
            ma_si->setSystemMerger(merger);
        }
        auto kcbs_planner = std::make_shared<omrc::KCBS>(ma_si);
        kcbs_planner->setLowLevelSolveTime(scene->low_level_solve_time);
        if (scene->merge_bound >= 0) kcbs_planner->setMergeBound(scene->merge_bound);
        planner = kcbs_planner;
    } else {
        planner = std::make_shared<omrc::PP>(ma_si);
    }
    planner->setProblemDefinition(ma_pdef);

    std::printf("Planning for %zu 2nd order cars in %s with %s.\n", scene->robots.size(), scene->name.c_str(), planner_name.c_str());
    const auto t0 = std::chrono::high_resolution_clock::now();
    const bool solved = (bool)planner->as<omrb::Planner>()->solve(solve_time);
    const double seconds = std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - t0).count();
    if (solved) {
        std::printf("Found Solution in %0.3f seconds!\n", seconds);
        omrb::PlanPtr solution = ma_pdef->getSolutionPlan();
        std::ofstream plan_file("plan.txt");
        solution->as<omrc::PlanControl>()->printAsMatrix(plan_file, "Robot");
    } else {
        std::printf("No solution within %0.1f seconds (%0.3f s used).\n", solve_time, seconds);
    }
    if (planner_name == "K-CBS") {
        std::ofstream tree_file("tree.txt");
        planner->as<omrc::KCBS>()->printConstraintTree(tree_file);
        planner->as<omrc::KCBS>()->printConstraintTree(std::cout);
    }
    return solved ? 0 : 1;
}
