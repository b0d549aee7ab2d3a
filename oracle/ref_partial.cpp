// ref_partial.cpp -- the part of the reference that DOES compile here, used to pin the oracle (test infrastructure).
//
// The reference (aria-systems-group/K-CBS-Demos) needs the OMPL K-CBS fork and Boost, neither of which is installed,
// so its hot path cannot be built as a whole.  Four of its headers, however, depend on OMPL only through declarations
// that the API shim of this repository provides (k-cbs-demos_b200/ompl_shim), and contain the arithmetic themselves:
//
//   includes/StatePropagatorDatabase.h   SecondOrderCarODE :44-64, TwoSecondOrderCarsODE :77-102  (the ODE right-hand sides)
//   includes/StateSpaceDatabase.h        createBounded2ndOrderCarStateSpace :42-62                 (the state bounds)
//   includes/ControlSpaceDatabase.h      createUniform2DRealVectorControlSpace :43-54              (the control bounds)
//   includes/GoalRegionDatabase.h        GoalRegion2ndOrderCar::distanceGoal :48-55, threshold :46 (the goal test)
//
// This file #includes those headers WHERE THEY LIE under /root/reference (nothing is copied) and exports their functions
// through a C ABI.  What runs inside these entry points is the reference's own code; the OMPL containers it fills
// (bounds vectors, control values) are the shim's.  The RK4 driver, the SO2 wrap and the Boost.Geometry predicates are
// NOT reachable this way and stay restated from the published upstream behaviour (kcbs_oracle.c, "UPSTREAM-RECALLED").
// Built by oracle/Makefile into oracle/_ref/ (git-ignored) when /root/reference exists.
#include <cmath>
#include <cstring>
#include <memory>
#include <vector>

#include <StatePropagatorDatabase.h>
#include <StateSpaceDatabase.h>
#include <ControlSpaceDatabase.h>
#include <GoalRegionDatabase.h>

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
