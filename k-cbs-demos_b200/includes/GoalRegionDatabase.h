// GoalRegionDatabase.h -- mirror of the reference's includes/GoalRegionDatabase.h (:42-108).  The goal regions are
// discs around a position; the planners read centre and radius through kcbs::GoalDisc and test the goal on the GPU
// (k_rrt_select), the virtuals below serve scalar callers.
#pragma once
#include <ompl/base/goals/GoalRegion.h>
#include <ompl/base/spaces/RealVectorStateSpace.h>

#include "kcbs_adaptors.h"

namespace ob = ompl::base;

class GoalRegion2ndOrderCar : public ob::GoalRegion, public kcbs::GoalDisc
{
public:
    GoalRegion2ndOrderCar(const ob::SpaceInformationPtr &si, double gx, double gy) : ob::GoalRegion(si), gx_(gx), gy_(gy)
    {
        threshold_ = 0.5;  // GoalRegionDatabase.h:48
    }
    double distanceGoal(const ob::State *st) const override
    {
        const double *p = st->as<ob::CompoundStateSpace::StateType>()->as<ob::RealVectorStateSpace::StateType>(0)->values;
        return std::hypot(p[0] - gx_, p[1] - gy_);
    }
    double goalX() const override { return gx_; }
    double goalY() const override { return gy_; }
    double goalRadius() const override { return threshold_; }

private:
    const double gx_, gy_;
};

// both cars of a merged pair within 1.5 of their goals (GoalRegionDatabase.h:61-108)
class GoalRegionTwo2ndOrderCars : public ob::Goal
{
public:
    GoalRegionTwo2ndOrderCars(const ob::SpaceInformationPtr &si, double gx1, double gy1, double gx2, double gy2)
        : ob::Goal(si), gx1_(gx1), gy1_(gy1), gx2_(gx2), gy2_(gy2) {}
    bool isSatisfied(const ob::State *st) const override
    {
        double d;
        return isSatisfied(st, &d);
    }
    bool isSatisfied(const ob::State *st, double *distance) const override
    {
        const auto *cs = st->as<ob::CompoundStateSpace::StateType>();
        const double *p1 = cs->as<ob::RealVectorStateSpace::StateType>(0)->values;
        const double *p2 = cs->as<ob::RealVectorStateSpace::StateType>(2)->values;
        const double d1 = std::hypot(p1[0] - gx1_, p1[1] - gy1_), d2 = std::hypot(p2[0] - gx2_, p2[1] - gy2_);
        *distance = d1 + d2;
        return d1 < threshold_ && d2 < threshold_;
    }

private:
    const double gx1_, gy1_, gx2_, gy2_;
    const double threshold_ = 1.5;
};
