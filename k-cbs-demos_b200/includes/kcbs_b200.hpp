// kcbs_b200.hpp -- C++ RAII layer over the C ABI (include/kcbs_b200.h).  Host-side plumbing for the
// mirror of the reference's plugin headers; no compute happens here.  Status codes are translated to
// the reference's conventions: unknown robot -> std::out_of_range (unordered_map::at,
// StateValidityCheckerDatabase.h:111), everything else -> std::runtime_error.
#pragma once

#include <array>
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/kcbs_b200.h"

namespace kcbs {

struct SceneDesc {
    double x_max = 10, y_max = 10;
    std::vector<std::array<double, 4>> obstacles;  // (x, y, length, width), Obstacle.h:82
    std::vector<std::array<double, 2>> robots;     // (length, width), Robot.h:82, addIndividual order
    double step_size = 0.1, integration_step = 1e-2, car_length = 0.5;
};

class Context {
public:
    explicit Context(const SceneDesc &scene, int device = 0)
    {
        kcbs_world w;
        kcbs_world_defaults(&w, scene.x_max, scene.y_max);
        w.step_size = scene.step_size;
        w.integration_step = scene.integration_step;
        w.car_length = scene.car_length;
        std::vector<double> obs, dims;
        for (const auto &o : scene.obstacles) obs.insert(obs.end(), o.begin(), o.end());
        for (const auto &r : scene.robots) dims.insert(dims.end(), r.begin(), r.end());
        w.n_obstacles = (int32_t)scene.obstacles.size();
        w.n_robots = (int32_t)scene.robots.size();
        w.obstacles = obs.empty() ? nullptr : obs.data();
        w.robot_dims = dims.data();
        kcbs_ctx *h = nullptr;
        const int st = kcbs_create(&w, device, &h);
        if (st != KCBS_OK) throw std::runtime_error(std::string("kcbs_create: ") + kcbs_last_error(nullptr));
        h_.reset(h, kcbs_destroy);
        n_robots_ = w.n_robots;
    }

    // adopts a context created directly through the C ABI
    Context(kcbs_ctx *h, int n_robots) : h_(h, kcbs_destroy), n_robots_(n_robots) {}

    kcbs_ctx *get() const { return h_.get(); }
    int robots() const { return n_robots_; }

    void check(int st) const
    {
        if (st == KCBS_OK) return;
        const std::string msg = kcbs_last_error(h_.get());
        if (st == KCBS_ERR_UNKNOWN_ROBOT) throw std::out_of_range(msg);
        throw std::runtime_error(msg);
    }

    // oc::SpaceInformation::propagateWhileValid for n samples (host buffers, SoA planes)
    void propagateWhileValid(int robot, int64_t n, int64_t n_parents, const float *parents, const int32_t *parent_idx,
                             const float *controls, const int32_t *steps, int max_steps, float *seg, float *fin,
                             int32_t *valid_steps) const
    {
        check(kcbs_propagate_batch(h_.get(), robot, n, n_parents, parents, parent_idx, controls, steps, max_steps, seg,
                                   fin, valid_steps));
    }

    // homogeneous2ndOrderCarSystemSVC::isValid for n states
    void isValid(int robot, int64_t n, const float *states, uint32_t *bitmask) const
    {
        check(kcbs_validity_batch(h_.get(), robot, n, states, bitmask));
    }

    // homogeneousTwo2ndOrderCarSystemSVC::isValid for n 10-d states (merged pair robot1 + robot2)
    void isValidPair(int robot1, int robot2, int64_t n, const float *states, uint32_t *bitmask) const
    {
        check(kcbs_validity_pair_batch(h_.get(), robot1, robot2, n, states, bitmask));
    }

    // KCBS::findConflicts for n_plans plans
    void firstConflict(int n_plans, int R, int T, const float *traj, const int32_t *lengths, kcbs_conflict *out) const
    {
        check(kcbs_first_conflict(h_.get(), n_plans, R, T, traj, lengths, out));
    }

private:
    std::shared_ptr<kcbs_ctx> h_;
    int n_robots_ = 0;
};

}  // namespace kcbs
