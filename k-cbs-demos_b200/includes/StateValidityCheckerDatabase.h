// StateValidityCheckerDatabase.h -- mirror of the reference's includes/StateValidityCheckerDatabase.h: same class
// names, constructor signatures and virtuals (:44-212 single car, :214-368 merged pair), evaluated on the GPU.
//
//   isValid         bounds + rotated rectangle vs obstacle rectangles      -> kcbs_validity_batch  (k_validity)
//   areStatesValid  rectangle vs rectangle of another robot                -> kcbs_first_conflict  (k_conflict_keys)
// The scalar virtuals submit one-element batches; isValidBatch / the planners use the batched entry points.
// Deviation kept on purpose: against a merged ("A and B") robot the reference hands the shapeless
// CompoundRobot to the exact test (:90,102, undefined behaviour); here each constituent is tested with its own shape.
#pragma once
#include <set>
#include <unordered_map>

#include <ompl/control/SpaceInformation.h>

#include "kcbs_adaptors.h"

namespace ob = ompl::base;

class homogeneous2ndOrderCarSystemSVC : public ob::StateValidityChecker, public kcbs::SceneProvider
{
public:
    homogeneous2ndOrderCarSystemSVC(const ob::SpaceInformationPtr &si, std::unordered_map<std::string, Robot *> robots,
                                    std::set<Obstacle *> obs_set = {})
        : ob::StateValidityChecker(si), robot1_name_(si->getStateSpace()->getName()), robot1_(robots.at(robot1_name_)),
          rad1_(robot1_->getBoundingRadius()), robots_(std::move(robots)), obstacles_(std::move(obs_set)), space_(si->getStateSpace())
    {
    }

    // is the robot described by `si_` at `state` valid?  (reference :54-75)
    bool isValid(const ob::State *state) const override
    {
        float q[5];
        kcbs::car_reals(state, 0, q);
        return kcbs::gpu_is_valid(*self_context(), 0, q);
    }

    // n states at once: 5 planes of n floats -> packed validity bits (kcbs_validity_batch)
    void isValidBatch(int64_t n, const float *states, uint32_t *bitmask) const { self_context()->isValid(0, n, states, bitmask); }

    // does this robot at state1 avoid the robot of state2.first at state2.second?  (reference :78-125)
    bool areStatesValid(const ob::State *state1, const std::pair<const ob::SpaceInformationPtr, const ob::State *> state2) const override
    {
        const std::string robot2_name = state2.first->getStateSpace()->getName();
        Robot *other = robots_.at(robot2_name);  // std::out_of_range for an unknown robot, like the reference
        float a[5];
        kcbs::car_reals(state1, 0, a);
        if (robot2_name.find("and") != robot2_name.npos) {
            for (unsigned int idx = 0; idx < 2; ++idx) {
                float b[5];
                kcbs::car_reals(state2.second, idx, b);
                if (!kcbs::gpu_pair_valid(*pair_context(other->as<CompoundRobot>()->getRobot(idx)), a, b)) return false;
            }
            return true;
        }
        float b[5];
        kcbs::car_reals(state2.second, 0, b);
        return kcbs::gpu_pair_valid(*pair_context(other), a, b);
    }

    void updateRobotMap(std::unordered_map<std::string, Robot *> new_robots) { robots_ = std::move(new_robots); }

    const std::unordered_map<std::string, Robot *> &sceneRobots() const override { return robots_; }
    const std::set<Obstacle *> &sceneObstacles() const override { return obstacles_; }

private:
    std::shared_ptr<kcbs::Context> self_context() const
    {
        std::lock_guard<std::mutex> lock(mu_);
        if (!self_) self_ = kcbs::cached_context(kcbs::make_scene(space_, {robot1_}, obstacles_));
        return self_;
    }
    // world with robot 0 = this robot and robot 1 = `other` (obstacles play no role in the pair test)
    std::shared_ptr<kcbs::Context> pair_context(const Robot *other) const
    {
        return kcbs::cached_context(kcbs::make_scene(space_, {robot1_, other}, {}));
    }

    const std::string robot1_name_;
    const Robot *robot1_;
    const double rad1_;
    std::unordered_map<std::string, Robot *> robots_;
    std::set<Obstacle *> obstacles_;
    ob::StateSpacePtr space_;
    mutable std::mutex mu_;
    mutable std::shared_ptr<kcbs::Context> self_;
};

// two cars planned as one system after a merge (reference :214-368); state = (x1, y1, v1, phi1, theta1, x2, ...)
class homogeneousTwo2ndOrderCarSystemSVC : public ob::StateValidityChecker, public kcbs::SceneProvider
{
public:
    homogeneousTwo2ndOrderCarSystemSVC(const ob::SpaceInformationPtr &si, const Robot *robot1, const Robot *robot2,
                                       std::unordered_map<std::string, Robot *> robots, std::set<Obstacle *> obs_set = {})
        : ob::StateValidityChecker(si), robot1_(robot1), robot2_(robot2), rad1_(robot1->getBoundingRadius()),
          rad2_(robot2->getBoundingRadius()), robots_(std::move(robots)), obstacles_(std::move(obs_set)), space_(si->getStateSpace())
    {
    }

    // both cars inside the bounds and clear of every obstacle, and not overlapping each other (reference :225-261)
    bool isValid(const ob::State *state) const override
    {
        float q[10];
        kcbs::car_reals(state, 0, q);
        kcbs::car_reals(state, 1, q + 5);
        uint32_t mask = 0;
        context()->isValidPair(0, 1, 1, q, &mask);  // robots of the context: 0 = robot1, 1 = robot2 (k_validity_pair)
        return (mask & 1u) != 0;
    }

    // do both cars avoid the (single) robot of state2.first at state2.second?  (reference :264-287)
    bool areStatesValid(const ob::State *state1, const std::pair<const ob::SpaceInformationPtr, const ob::State *> state2) const override
    {
        Robot *other = robots_.at(state2.first->getStateSpace()->getName());
        float o[5];
        kcbs::car_reals(state2.second, 0, o);
        const Robot *mine[2] = {robot1_, robot2_};
        for (unsigned int idx = 0; idx < 2; ++idx) {
            float a[5];
            kcbs::car_reals(state1, idx, a);
            auto ctx = kcbs::cached_context(kcbs::make_scene(space_, {mine[idx], other}, {}));
            if (!kcbs::gpu_pair_valid(*ctx, a, o)) return false;
        }
        return true;
    }

    const std::unordered_map<std::string, Robot *> &sceneRobots() const override { return robots_; }
    const std::set<Obstacle *> &sceneObstacles() const override { return obstacles_; }

private:
    std::shared_ptr<kcbs::Context> context() const { return kcbs::cached_context(kcbs::make_scene(space_, {robot1_, robot2_}, obstacles_)); }

    const Robot *robot1_;
    const Robot *robot2_;
    const double rad1_, rad2_;
    std::unordered_map<std::string, Robot *> robots_;
    std::set<Obstacle *> obstacles_;
    ob::StateSpacePtr space_;
};
