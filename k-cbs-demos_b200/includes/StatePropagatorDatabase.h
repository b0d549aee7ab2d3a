// StatePropagatorDatabase.h -- mirror of the reference's includes/StatePropagatorDatabase.h (:44-171): same function
// and class names.  The dynamics
//     xdot = v cos(theta), ydot = v sin(theta), vdot = a, phidot = w, thetadot = (v / 0.5) tan(phi)      (:59-63)
// are integrated by k_propagate (csrc/propagate.cu) with the reference's RK4 stages, step 0.01, and the theta wrap of
// the post-integration event after every propagation step.  The ODE functors below are therefore *model tags*: handing
// &SecondOrderCarODE to oc::ODEBasicSolver selects the GPU integrator (ODESolver::getStatePropagator in
// kcbs_adaptors.h).  They are never evaluated on the host; there is no CPU integrator in this build.
#pragma once
#include <ompl/base/spaces/SO2StateSpace.h>
#include <ompl/control/ODESolver.h>
#include <ompl/control/spaces/RealVectorControlSpace.h>

#include "kcbs_adaptors.h"

namespace ob = ompl::base;
namespace oc = ompl::control;

// q = x, y, v, phi, theta ; control = a, phidot
inline void SecondOrderCarODE(const oc::ODESolver::StateType & /*q*/, const oc::Control * /*control*/, oc::ODESolver::StateType & /*qdot*/)
{
    throw std::logic_error("SecondOrderCarODE is a model tag: the right-hand side is evaluated on the GPU (k_propagate)");
}

// theta -> [-pi, pi) after every propagation step (SO2StateSpace::enforceBounds); k_propagate already wraps, so this
// hook only normalises states that reach it through other routes
inline void SecondOrderCarODEPostIntegration(const ob::State * /*state*/, const oc::Control * /*control*/, const double /*duration*/, ob::State *result)
{
    ob::SO2StateSpace SO2;
    SO2.enforceBounds(result->as<ob::CompoundState>()->as<ob::SO2StateSpace::StateType>(1));
}

// q = x1, y1, v1, phi1, theta1, x2, y2, v2, phi2, theta2 ; control = a1, phidot1, a2, phidot2
inline void TwoSecondOrderCarsODE(const oc::ODESolver::StateType & /*q*/, const oc::Control * /*control*/, oc::ODESolver::StateType & /*qdot*/)
{
    throw std::logic_error("TwoSecondOrderCarsODE is a model tag: the right-hand side is evaluated on the GPU (k_propagate)");
}

inline void TwoSecondOrderCarODEPostIntegration(const ob::State * /*state*/, const oc::Control * /*control*/, const double /*duration*/, ob::State *result)
{
    ob::SO2StateSpace SO2;
    for (unsigned int idx = 0; idx < 2; idx++) SO2.enforceBounds(result->as<ob::CompoundState>()->as<ob::SO2StateSpace::StateType>(2 * idx + 1));
}

namespace kcbs {
static const OdeRegistrar register_second_order_car(&SecondOrderCarODE, OdeModel::SecondOrderCar);
static const OdeRegistrar register_two_second_order_cars(&TwoSecondOrderCarsODE, OdeModel::TwoSecondOrderCars);
}  // namespace kcbs

// Propagator of a merged pair: both cars move under their own controls, but a car that was already inside its goal
// region (distance < 1.5) before the step stays where it is (reference :130-153).
class TwoSecondOrderCarStatePropagator : public oc::StatePropagator
{
public:
    TwoSecondOrderCarStatePropagator(const oc::SpaceInformationPtr &si, double gx1, double gy1, double gx2, double gy2)
        : oc::StatePropagator(si.get()), gx1_(gx1), gy1_(gy1), gx2_(gx2), gy2_(gy2)
    {
        auto odeSolver(std::make_shared<oc::ODEBasicSolver<>>(si, &TwoSecondOrderCarsODE));
        systemSP_ = oc::ODESolver::getStatePropagator(odeSolver, &TwoSecondOrderCarODEPostIntegration);
    }

    void propagate(const ob::State *state, const oc::Control *control, double duration, ob::State *result) const override
    {
        systemSP_->propagate(state, control, duration, result);
        const double g[2][2] = {{gx1_, gy1_}, {gx2_, gy2_}};
        for (unsigned int car = 0; car < 2; ++car) {
            float q[5];
            kcbs::car_reals(state, car, q);
            if (std::hypot(q[0] - g[car][0], q[1] - g[car][1]) < threshold_) {
                // copy this car's part of `state` into `result` unchanged (the reference copies 5 values out of the
                // 4-d real vector here, :140-141 -- an out-of-bounds read that is not reproduced)
                const auto *src = state->as<ob::CompoundStateSpace::StateType>();
                auto *dst = result->as<ob::CompoundStateSpace::StateType>();
                for (unsigned int i = 0; i < 4; ++i)
                    dst->as<ob::RealVectorStateSpace::StateType>(2 * car)->values[i] = src->as<ob::RealVectorStateSpace::StateType>(2 * car)->values[i];
                dst->as<ob::SO2StateSpace::StateType>(2 * car + 1)->value = src->as<ob::SO2StateSpace::StateType>(2 * car + 1)->value;
            }
        }
    }

    bool canPropagateBackward() const override { return false; }

private:
    oc::StatePropagatorPtr systemSP_ = nullptr;
    const double gx1_, gy1_, gx2_, gy2_;
    const double threshold_ = 1.5;
};
