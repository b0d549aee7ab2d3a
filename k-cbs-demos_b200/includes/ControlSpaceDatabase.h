// ControlSpaceDatabase.h -- mirror of the reference's includes/ControlSpaceDatabase.h (:43-67): real-vector control
// spaces with every component in [-1, 1]; 2-D (a, phidot) for one car, 4-D for a merged pair.
#pragma once
#include <ompl/base/spaces/RealVectorStateSpace.h>
#include <ompl/control/spaces/RealVectorControlSpace.h>

namespace ob = ompl::base;
namespace oc = ompl::control;

namespace kcbs {
inline oc::ControlSpacePtr unit_control_space(ob::StateSpacePtr &space, unsigned int dim)
{
    auto cspace = std::make_shared<oc::RealVectorControlSpace>(space, dim);
    ob::RealVectorBounds b(dim);
    b.setLow(-1);
    b.setHigh(1);
    cspace->setBounds(b);
    return cspace;
}
}  // namespace kcbs

inline oc::ControlSpacePtr createUniform2DRealVectorControlSpace(ob::StateSpacePtr &space) { return kcbs::unit_control_space(space, 2); }
inline oc::ControlSpacePtr createUniform4DRealVectorControlSpace(ob::StateSpacePtr &space) { return kcbs::unit_control_space(space, 4); }
