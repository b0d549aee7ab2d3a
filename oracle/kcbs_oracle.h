/*
 * kcbs_oracle.h -- CPU (double precision, plain C) restatement of the K-CBS low-level hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (k-cbs-demos_b200/) may include, link or
 * call this.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs use it, and there only as the checker / the reported CPU baseline.
 *
 * PARITY PARTLY PINNED, OTHERWISE UNPINNED: the reference (/root/reference) ships no tests or golden
 * vectors, and its integrator, bounds test, angle wrap and conflict scan loop live in un-vendored
 * dependencies (an unpinned OMPL fork with K-CBS, Boost.odeint, Boost.Geometry) that are absent from
 * this image, so the reference's path cannot be compiled or run here as a whole.  What does compile --
 * the ODE right-hand sides, the state and control bounds and the goal test, which the reference's
 * headers compute themselves -- is built from /root/reference/includes by ref_partial.cpp and pins
 * orc_car_ode and the scene constants bit for bit (tests/test_ref_partial.py, tests/golden/ref_partial.npz).
 * Everything else is unpinned: every function below cites the reference file:line it follows;
 * semantics that live upstream are marked [UPSTREAM] and are restated from the published algorithms of
 * those libraries, anchored on the reference's own call sites.  Each [UPSTREAM] rule is a named switch
 * so it can be re-pinned in one place.
 */
#ifndef KCBS_ORACLE_H
#define KCBS_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MAX_OBSTACLES 64
#define ORC_MAX_ROBOTS 256

/* [UPSTREAM] SO2StateSpace::enforceBounds variants (StatePropagatorDatabase.h:72-73 calls it). */
#define ORC_WRAP_HALF_OPEN 0 /* [-pi, pi)  -- current OMPL (default)              */
#define ORC_WRAP_LEGACY 1    /* (-pi, pi]  -- pre-2016 OMPL, kept as a switch      */

typedef struct orc_world {
    /* StateSpaceDatabase.h:49-59: RealVector(4) bounds in reals order (x, y, v, phi). */
    double lo[4];
    double hi[4];
    double step_size;        /* demo_*.cpp:115  setPropagationStepSize(0.1)                */
    double integration_step; /* [UPSTREAM] ODEBasicSolver default intStep = 1e-2           */
    double car_length;       /* StatePropagatorDatabase.h:53                                */
    int32_t wrap_mode;       /* ORC_WRAP_*                                                  */
    int32_t n_obstacles;
    /* Obstacle.h:82-99: lower-left corner (x, y) + extents (length along x, width along y). */
    double obstacles[ORC_MAX_OBSTACLES][4];
    int32_t n_robots;
    /* Robot.h:82-94: rectangle length (local x) and width (local y), centred at the origin. */
    double robot_dims[ORC_MAX_ROBOTS][2];
} orc_world;

typedef struct orc_conflict {
    int32_t r1, r2;   /* robot indices, r1 < r2; -1 when the plan is conflict free */
    int32_t k_start;  /* first time index at which (r1, r2) collide                 */
    int32_t k_end;    /* last consecutive colliding index of that same pair         */
} orc_conflict;

/* ---- scalar building blocks (one reference call each) --------------------------------- */

/* SecondOrderCarODE, StatePropagatorDatabase.h:44-64.  q = (x, y, v, phi, theta), u = (a, phidot). */
void orc_car_ode(const double q[5], const double u[2], double car_length, double qdot[5]);

/* [UPSTREAM] boost::numeric::odeint::runge_kutta4::do_step on orc_car_ode. */
void orc_rk4_step(double q[5], const double u[2], double h, double car_length);

/* [UPSTREAM] odeint::integrate_const(stepper, f, q, 0, duration, h): number of h-steps taken. */
int orc_integrate_const(double q[5], const double u[2], double duration, double h, double car_length);

/* [UPSTREAM] SO2StateSpace::enforceBounds (called from StatePropagatorDatabase.h:67-74). */
double orc_wrap_angle(double theta, int wrap_mode);

/* oc::StatePropagator::propagate as wired by demo_*.cpp:111-112: integrate + post-event wrap. */
void orc_propagate(const orc_world *w, const double q_in[5], const double u[2], double duration, double q_out[5]);

/* [UPSTREAM] CompoundStateSpace::satisfiesBounds (StateValidityCheckerDatabase.h:57).
 * *slack (optional) = min distance to any bound, negative when violated. */
int orc_satisfies_bounds(const orc_world *w, const double q[5], double *slack);

/* homogeneous2ndOrderCarSystemSVC::isValid, StateValidityCheckerDatabase.h:54-75.
 * *margin (optional): signed distance to the nearest decision boundary (min of bound slack and
 * SAT gaps); tests use it to mask the 1e-6 band north_star excludes. */
int orc_is_valid(const orc_world *w, int robot, const double q[5], double *margin);

/* homogeneous2ndOrderCarSystemSVC::areStatesValid (non-composed branch),
 * StateValidityCheckerDatabase.h:108-124.  s = (x, y, theta).  Returns 1 when NOT in collision. */
int orc_pair_valid(const orc_world *w, int r1, const double s1[3], int r2, const double s2[3], double *margin);

/* [UPSTREAM] oc::SpaceInformation::propagateWhileValid (demo_*.cpp:115,118 configure it).
 * seg (optional) receives up to `steps` states of 5 doubles; returns the number of valid steps;
 * q_final = last valid state (= q0 when 0). */
int orc_propagate_while_valid(const orc_world *w, int robot, const double q0[5], const double u[2], int steps,
                              double *seg, double q_final[5]);

/* Robot footprint corners in world frame, matrix_transformer of StateValidityCheckerDatabase.h:152-155
 * (rotation by -theta).  out = 4 corners (x, y) in Robot.h:85-94 append order. */
void orc_robot_corners(double length, double width, double x, double y, double theta, double out[4][2]);

/* [UPSTREAM] boost::geometry::disjoint(polygon, polygon) restated for simple polygons: no
 * boundary segments intersect or touch AND neither polygon contains a vertex of the other. */
int orc_polygons_disjoint(int n1, const double (*p1)[2], int n2, const double (*p2)[2]);

/* Separating-axis gap (> 0 separated, <= 0 overlapping/touching) for two rectangles; equal in sign
 * to orc_polygons_disjoint on them. */
double orc_sat_gap_robot_robot(double la, double wa, const double sa[3], double lb, double wb, const double sb[3]);
double orc_sat_gap_robot_obstacle(double l, double wd, const double s[3], const double obs[4]);

/* ---- batched drivers (float32 I/O identical to what the GPU path receives) -------------- */

/* One expansion: n (parent, control, steps) triples.  parents = 5 planes of n_parents floats
 * (x, y, v, phi, theta); parent_idx optional; controls = 2 planes of n; seg_out (optional) =
 * [5][max_steps][n] doubles; final_out (optional) = [5][n]; valid_steps = [n].
 * Returns the number of propagate() calls performed.  Uses n_threads pthreads. */
int64_t orc_propagate_batch(const orc_world *w, int robot, int64_t n, int64_t n_parents, const float *parents,
                            const int32_t *parent_idx, const float *controls, const int32_t *steps,
                            int max_steps, double *seg_out, double *final_out, int32_t *valid_steps,
                            int n_threads);

/* states = 5 planes of n floats; valid (optional) = n bytes; margin (optional) = n doubles. */
int64_t orc_validity_batch(const orc_world *w, int robot, int64_t n, const float *states, uint8_t *valid,
                           double *margin, int n_threads);

/* [UPSTREAM] KCBS::findConflicts scan order (call sites demo_*.cpp:155,168; Benchmark.h:151-158).
 * traj = [R][3][T] floats (x, y, theta planes per robot); lengths (optional) = R path lengths,
 * shorter paths are padded with their last state.  min_margin (optional) = smallest |SAT gap|
 * seen over every pair-time predicate evaluated up to and including the conflict.
 * Returns the number of pair predicates evaluated. */
int64_t orc_first_conflict(const orc_world *w, int R, int T, const float *traj, const int32_t *lengths,
                           orc_conflict *out, double *min_margin);

/* Same over P plans ([P][R][3][T]) with n_threads pthreads over plans; full_scan != 0 disables the
 * early return (every pair x time predicate is evaluated; the result is still the first conflict). */
int64_t orc_first_conflict_batch(const orc_world *w, int P, int R, int T, const float *traj, const int32_t *lengths,
                                 orc_conflict *out, int full_scan, int n_threads);

/* ---- merged pair (two cars planned as one 10-d system after a K-CBS merge) ------------------------------ */

/* homogeneousTwo2ndOrderCarSystemSVC::isValid, StateValidityCheckerDatabase.h:225-261: both cars inside the bounds,
 * the two cars not overlapping, both cars clear of every obstacle.  q = (x1, y1, v1, phi1, theta1, x2, ...).
 * Each car is tested with its own rectangle (the reference uses robot1_'s shape and radius for both in the obstacle
 * loop, :245-256, which is the same thing in its homogeneous scenes). */
int orc_pair_is_valid(const orc_world *w, int r1, int r2, const double q[10], double *margin);

/* TwoSecondOrderCarStatePropagator::propagate, StatePropagatorDatabase.h:130-153: both cars are integrated
 * (TwoSecondOrderCarsODE :77-102 + post-integration wrap :106-117), then a car whose position BEFORE the step was
 * within hold_radius (1.5, :170) of its goal keeps its previous state.  goals = (gx1, gy1, gx2, gy2). */
void orc_pair_propagate(const orc_world *w, const double q_in[10], const double u[4], double duration, const double goals[4],
                        double hold_radius, double q_out[10]);

/* propagateWhileValid of the merged system. seg (optional) = steps x 10 doubles. */
int orc_pair_propagate_while_valid(const orc_world *w, int r1, int r2, const double q0[10], const double u[4], int steps,
                                   const double goals[4], double hold_radius, double *seg, double q_final[10]);

/* Batched drivers, float32 I/O: parents [10][n_parents], controls [4][n], seg_out [10][max_steps][n] doubles,
 * final_out [10][n], hold_margin (optional) [n] = min over (step, car) of |distance to goal - hold_radius|. */
int64_t orc_pair_propagate_batch(const orc_world *w, int r1, int r2, int64_t n, int64_t n_parents, const float *parents,
                                 const int32_t *parent_idx, const float *controls, const int32_t *steps, int max_steps,
                                 const double goals[4], double hold_radius, double *seg_out, double *final_out,
                                 int32_t *valid_steps, double *hold_margin, int n_threads);
int64_t orc_pair_validity_batch(const orc_world *w, int r1, int r2, int64_t n, const float *states, uint8_t *valid,
                                double *margin, int n_threads);

/* findConflicts over individuals some of which are merged: bodies with the same group id belong to one individual
 * and are never tested against each other (their mutual collision is part of that individual's isValid).
 * Results name bodies (rows of traj). */
int64_t orc_first_conflict_groups(const orc_world *w, int R, int T, const float *traj, const int32_t *lengths,
                                  const int32_t *group, orc_conflict *out, double *min_margin);

#ifdef __cplusplus
}
#endif
#endif
