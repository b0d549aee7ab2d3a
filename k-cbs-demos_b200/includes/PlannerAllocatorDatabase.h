// PlannerAllocatorDatabase.h -- mirror of the reference's includes/PlannerAllocatorDatabase.h (:40-45).
// allocateControlRRT is the ob::PlannerAllocator the demos install (demo_*.cpp:141-142); the oc::RRT it returns is the
// batched GPU low-level planner (kcbs_rrt_plan) when built against the shim.
#pragma once
#include <ompl/control/SpaceInformation.h>
#include <ompl/control/planners/rrt/RRT.h>

#include "kcbs_adaptors.h"

inline ompl::base::PlannerPtr allocateControlRRT(const ompl::base::SpaceInformationPtr &si)
{
    const auto siC = std::static_pointer_cast<ompl::control::SpaceInformation>(si);
    return std::make_shared<ompl::control::RRT>(siC);
}
