// SystemMergerDatabase.h -- mirror of the reference's includes/SystemMergerDatabase.h (:47-176): the SystemMerger that
// K-CBS calls when one pair of robots has conflicted more often than the merge bound.  Pure object wiring: the two
// individuals become one 10-d system (two cars), the others keep their spaces but get validity checkers that know the
// new robot map.  Like the reference it must be included after the other *Database.h headers.
#pragma once
#include <unordered_map>

#include <ompl/multirobot/control/SpaceInformation.h>

#include "Obstacle.h"
#include "Robot.h"

namespace omr = ompl::multirobot;
namespace omrb = ompl::multirobot::base;
namespace omrc = ompl::multirobot::control;

class homogeneous2ndOrderCarSystemMerger : public omrc::SystemMerger
{
public:
    homogeneous2ndOrderCarSystemMerger(const omrc::SpaceInformationPtr &si, const omrb::ProblemDefinitionPtr pdef,
                                       std::unordered_map<std::string, Robot *> robots, std::set<Obstacle *> obstacles,
                                       const unsigned int space_bound, const std::map<std::string, std::pair<double, double>> start_map,
                                       const std::map<std::string, std::pair<double, double>> goal_map)
        : omrc::SystemMerger(si, pdef), robots_(std::move(robots)), obstacles_(std::move(obstacles)), bound_(space_bound),
          starts_(start_map), goals_(goal_map)
    {
    }

    std::pair<const omrc::SpaceInformationPtr, const omrb::ProblemDefinitionPtr> merge(const int index1, const int index2) const override
    {
        auto merged_si = std::make_shared<omrc::SpaceInformation>();
        auto merged_pdef = std::make_shared<omrb::ProblemDefinition>(merged_si);

        const std::string name1 = si_->getIndividual(index1)->getStateSpace()->getName();
        const std::string name2 = si_->getIndividual(index2)->getStateSpace()->getName();
        const std::string merged_name = name1 + " and " + name2;
        Robot *robot1 = robots_.at(name1);
        Robot *robot2 = robots_.at(name2);

        // robot map after the merge: everybody else plus the compound robot
        std::unordered_map<std::string, Robot *> new_map;
        for (const auto &kv : robots_)
            if (kv.first != name1 && kv.first != name2) new_map[kv.first] = kv.second;
        new_map[merged_name] = new CompoundRobot(merged_name, robot1, robot2);

        // untouched individuals: constraints and old paths dropped, validity checker rebuilt on the new map
        for (unsigned int idx = 0; idx < si_->getIndividualCount(); idx++) {
            if ((int)idx == index1 || (int)idx == index2) continue;
            auto ind = si_->getIndividual(idx);
            ind->clearDynamicObstacles();
            pdef_->getIndividual(idx)->clearSolutionPaths();
            ind->setStateValidityChecker(std::make_shared<homogeneous2ndOrderCarSystemSVC>(ind, new_map, obstacles_));
            merged_si->addIndividual(ind);
            merged_pdef->addIndividual(pdef_->getIndividual(idx));
        }

        // merging something that is already a merged pair is not supported
        if (si_->getIndividual(index1)->getStateSpace()->getDimension() != 5 || si_->getIndividual(index2)->getStateSpace()->getDimension() != 5)
            return {nullptr, nullptr};

        // the composed 10-d system, placed last
        auto space = createBoundedTwo2ndOrderCarStateSpace(bound_, bound_);
        space->setName(merged_name);
        auto cspace = createUniform4DRealVectorControlSpace(space);
        auto si = std::make_shared<oc::SpaceInformation>(space, cspace);
        si->setStateValidityChecker(std::make_shared<homogeneousTwo2ndOrderCarSystemSVC>(si, robot1, robot2, robots_, obstacles_));
        const auto &g1 = goals_.at(name1), &g2 = goals_.at(name2);
        si->setStatePropagator(std::make_shared<TwoSecondOrderCarStatePropagator>(si, g1.first, g1.second, g2.first, g2.second));
        si->setPropagationStepSize(0.1);
        si->setMinMaxControlDuration(1, 10);

        ob::ScopedState<> start(space);
        const auto &s1 = starts_.at(name1), &s2 = starts_.at(name2);
        for (unsigned int i = 0; i < 10; ++i) start[i] = 0.0;
        start[0] = s1.first;  start[1] = s1.second;
        start[5] = s2.first;  start[6] = s2.second;

        auto pdef = std::make_shared<ob::ProblemDefinition>(si);
        pdef->addStartState(start);
        pdef->setGoal(std::make_shared<GoalRegionTwo2ndOrderCars>(si, g1.first, g1.second, g2.first, g2.second));

        merged_si->addIndividual(si);
        merged_pdef->addIndividual(pdef);
        merged_si->lock();
        merged_pdef->lock();
        ompl::base::PlannerAllocator allocator = &allocateControlRRT;
        merged_si->setPlannerAllocator(allocator);
        return {merged_si, merged_pdef};
    }

private:
    std::unordered_map<std::string, Robot *> robots_;
    std::set<Obstacle *> obstacles_;
    const unsigned int bound_;
    const std::map<std::string, std::pair<double, double>> starts_;
    const std::map<std::string, std::pair<double, double>> goals_;
};
