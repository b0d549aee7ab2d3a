// StateSpaceDatabase.h -- mirror of the reference's includes/StateSpaceDatabase.h: same factory functions, same
// spaces and bounds (StateSpaceDatabase.h:42-90 of K-CBS-Demos).  These bounds are what kcbs_world_defaults hands to
// the GPU validity test.
#pragma once
#include <ompl/base/spaces/RealVectorStateSpace.h>
#include <ompl/base/spaces/SO2StateSpace.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

namespace ob = ompl::base;

namespace kcbs {
// bounds of one 2nd-order car: x in [0, x_max], y in [0, y_max], v in [-1, 1], phi in [-pi/3, pi/3]
inline ob::RealVectorBounds car_bounds(unsigned int x_max, unsigned int y_max)
{
    ob::RealVectorBounds b(4);
    b.setLow(0, 0);           b.setHigh(0, x_max);
    b.setLow(1, 0);           b.setHigh(1, y_max);
    b.setLow(2, -1);          b.setHigh(2, 1);
    b.setLow(3, -M_PI / 3);   b.setHigh(3, M_PI / 3);
    return b;
}
inline ob::StateSpacePtr car_space(unsigned int cars, unsigned int x_max, unsigned int y_max)
{
    auto space = std::make_shared<ob::CompoundStateSpace>();
    for (unsigned int c = 0; c < cars; ++c) {
        space->addSubspace(std::make_shared<ob::RealVectorStateSpace>(4), 1.0);  // x, y, v, phi
        space->addSubspace(std::make_shared<ob::SO2StateSpace>(), 1.0);          // theta
    }
    space->lock();
    for (unsigned int c = 0; c < cars; ++c) space->as<ob::RealVectorStateSpace>(2 * c)->setBounds(car_bounds(x_max, y_max));
    return space;
}
}  // namespace kcbs

inline ob::StateSpacePtr createBounded2ndOrderCarStateSpace(const unsigned int x_max, const unsigned int y_max)
{
    return kcbs::car_space(1, x_max, y_max);
}

inline ob::StateSpacePtr createBoundedTwo2ndOrderCarStateSpace(const unsigned int x_max, const unsigned int y_max)
{
    return kcbs::car_space(2, x_max, y_max);
}
