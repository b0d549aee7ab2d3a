/*
 * kcbs_oracle.c -- CPU (double precision) restatement of the K-CBS low-level hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see kcbs_oracle.h).  PARITY UNPINNED: the reference has no golden
 * vectors and cannot be built here; this file restates the cited reference lines plus the
 * published behaviour of OMPL / Boost.odeint / Boost.Geometry ([UPSTREAM]).
 *
 * All `file:line` citations are relative to /root/reference/.
 */
#include "kcbs_oracle.h"

#include <float.h>
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* ------------------------------------------------------------------------------------------
 * Dynamics: includes/StatePropagatorDatabase.h:44-64
 * ---------------------------------------------------------------------------------------- */
void orc_car_ode(const double q[5], const double u[2], double car_length, double qdot[5])
{
    const double v = q[2];     /* :50 */
    const double phi = q[3];   /* :51 */
    const double theta = q[4]; /* :52 */
    qdot[0] = v * cos(theta);            /* :59 */
    qdot[1] = v * sin(theta);            /* :60 */
    qdot[2] = u[0];                      /* :61 */
    qdot[3] = u[1];                      /* :62 */
    qdot[4] = (v / car_length) * tan(phi); /* :63 */
}

/* [UPSTREAM] odeint runge_kutta4 = explicit_generic_rk<4> with the classical tableau
 * c = (0, 1/2, 1/2, 1), a = ((1/2), (0, 1/2), (0, 0, 1)), b = (1/6, 1/3, 1/3, 1/6).
 * The generic algorithm forms every stage as x + dt*a_i1*k1 + ... including the zero
 * coefficients, and the result as x + dt*b1*k1 + dt*b2*k2 + dt*b3*k3 + dt*b4*k4. */
void orc_rk4_step(double q[5], const double u[2], double h, double L)
{
    double k1[5], k2[5], k3[5], k4[5], t[5];
    int i;
    orc_car_ode(q, u, L, k1);
    for (i = 0; i < 5; ++i) t[i] = q[i] + (h * 0.5) * k1[i];
    orc_car_ode(t, u, L, k2);
    for (i = 0; i < 5; ++i) t[i] = q[i] + (h * 0.0) * k1[i] + (h * 0.5) * k2[i];
    orc_car_ode(t, u, L, k3);
    for (i = 0; i < 5; ++i) t[i] = q[i] + (h * 0.0) * k1[i] + (h * 0.0) * k2[i] + (h * 1.0) * k3[i];
    orc_car_ode(t, u, L, k4);
    for (i = 0; i < 5; ++i)
        q[i] = q[i] + (h * (1.0 / 6.0)) * k1[i] + (h * (1.0 / 3.0)) * k2[i] + (h * (1.0 / 3.0)) * k3[i] +
               (h * (1.0 / 6.0)) * k4[i];
}

/* [UPSTREAM] odeint::integrate_const for a plain stepper: while ((t + dt) - t_end <= eps) step;
 * t is recomputed as start + n*dt (no accumulation).  ODEBasicSolver::solve calls it with
 * start 0, end = duration, dt = intStep (1e-2).  duration 0.1 => exactly 10 steps. */
int orc_integrate_const(double q[5], const double u[2], double duration, double h, double L)
{
    double time = 0.0;
    int step = 0;
    while (((time + h) - duration) <= DBL_EPSILON) {
        orc_rk4_step(q, u, h, L);
        ++step;
        time = 0.0 + (double)step * h;
    }
    return step;
}

/* [UPSTREAM] ompl::base::SO2StateSpace::enforceBounds; reached through
 * SecondOrderCarODEPostIntegration, StatePropagatorDatabase.h:67-74. */
double orc_wrap_angle(double theta, int wrap_mode)
{
    double v = fmod(theta, 2.0 * M_PI);
    if (wrap_mode == ORC_WRAP_LEGACY) {
        if (v <= -M_PI)
            v += 2.0 * M_PI;
        else if (v > M_PI)
            v -= 2.0 * M_PI;
    } else {
        if (v < -M_PI)
            v += 2.0 * M_PI;
        else if (v >= M_PI)
            v -= 2.0 * M_PI;
    }
    return v;
}

/* [UPSTREAM] ODESolver::getStatePropagator(solver, postEvent)::propagate:
 * copyToReals -> solve -> copyFromReals -> postEvent (demo_Corridor...:111-112). */
void orc_propagate(const orc_world *w, const double q_in[5], const double u[2], double duration, double q_out[5])
{
    double q[5];
    memcpy(q, q_in, sizeof q);
    orc_integrate_const(q, u, duration, w->integration_step, w->car_length);
    q[4] = orc_wrap_angle(q[4], w->wrap_mode);
    memcpy(q_out, q, sizeof q);
}

/* ------------------------------------------------------------------------------------------
 * Bounds: [UPSTREAM] RealVectorStateSpace::satisfiesBounds with eps = DBL_EPSILON and
 * SO2StateSpace::satisfiesBounds; bounds from includes/StateSpaceDatabase.h:49-59.
 * ---------------------------------------------------------------------------------------- */
int orc_satisfies_bounds(const orc_world *w, const double q[5], double *slack)
{
    int ok = 1, i;
    double s = INFINITY;
    for (i = 0; i < 4; ++i) {
        if (q[i] - DBL_EPSILON > w->hi[i] || q[i] + DBL_EPSILON < w->lo[i]) ok = 0;
        if (w->hi[i] - q[i] < s) s = w->hi[i] - q[i];
        if (q[i] - w->lo[i] < s) s = q[i] - w->lo[i];
    }
    /* SO2: (value < pi + eps) && (value > -pi - eps) */
    if (!((q[4] < M_PI + DBL_EPSILON) && (q[4] > -M_PI - DBL_EPSILON))) ok = 0;
    if (M_PI - q[4] < s) s = M_PI - q[4];
    if (q[4] + M_PI < s) s = q[4] + M_PI;
    if (slack) *slack = s;
    return ok;
}

/* ------------------------------------------------------------------------------------------
 * Geometry.  Robot.h:79-100 (rectangle centred at origin, radius hypot(L/2, W/2));
 * Obstacle.h:79-103 (lower-left + extents, centre, radius);
 * footprint transform StateValidityCheckerDatabase.h:152-155 / :188-191:
 *   matrix_transformer( cos t, sin t, x ; -sin t, cos t, y )  =>  rotation by -theta.
 * ---------------------------------------------------------------------------------------- */
void orc_robot_corners(double length, double width, double x, double y, double theta, double out[4][2])
{
    /* Robot.h:85-94 append order: bottom-left, bottom-right, top-right, top-left */
    const double px[4] = {-length / 2, length / 2, length / 2, -length / 2};
    const double py[4] = {-width / 2, -width / 2, width / 2, width / 2};
    const double c = cos(theta), s = sin(theta);
    int i;
    for (i = 0; i < 4; ++i) {
        out[i][0] = c * px[i] + s * py[i] + x;
        out[i][1] = -s * px[i] + c * py[i] + y;
    }
}

static void obstacle_corners(const double obs[4], double out[4][2])
{
    /* Obstacle.h:88-97 */
    out[0][0] = obs[0];          out[0][1] = obs[1];
    out[1][0] = obs[0] + obs[2]; out[1][1] = obs[1];
    out[2][0] = obs[0] + obs[2]; out[2][1] = obs[1] + obs[3];
    out[3][0] = obs[0];          out[3][1] = obs[1] + obs[3];
}

static double cross3(const double a[2], const double b[2], const double c[2])
{
    return (b[0] - a[0]) * (c[1] - a[1]) - (b[1] - a[1]) * (c[0] - a[0]);
}

static int sgn(double v) { return (v > 0) - (v < 0); }

static int on_segment(const double a[2], const double b[2], const double p[2])
{
    return fmin(a[0], b[0]) <= p[0] && p[0] <= fmax(a[0], b[0]) && fmin(a[1], b[1]) <= p[1] &&
           p[1] <= fmax(a[1], b[1]);
}

/* closed segments share at least one point */
static int segments_touch(const double a[2], const double b[2], const double c[2], const double d[2])
{
    const int o1 = sgn(cross3(a, b, c)), o2 = sgn(cross3(a, b, d));
    const int o3 = sgn(cross3(c, d, a)), o4 = sgn(cross3(c, d, b));
    if (o1 != o2 && o3 != o4) return 1;
    if (o1 == 0 && on_segment(a, b, c)) return 1;
    if (o2 == 0 && on_segment(a, b, d)) return 1;
    if (o3 == 0 && on_segment(c, d, a)) return 1;
    if (o4 == 0 && on_segment(c, d, b)) return 1;
    return 0;
}

/* point strictly inside a simple polygon (crossing number); boundary cases are handled by the
 * segment test that runs first. */
static int point_in_polygon(int n, const double (*p)[2], const double q[2])
{
    int i, j, in = 0;
    for (i = 0, j = n - 1; i < n; j = i++) {
        if (((p[i][1] > q[1]) != (p[j][1] > q[1])) &&
            (q[0] < (p[j][0] - p[i][0]) * (q[1] - p[i][1]) / (p[j][1] - p[i][1]) + p[i][0]))
            in = !in;
    }
    return in;
}

/* [UPSTREAM] boost::geometry::disjoint(polygon, polygon) (called at
 * StateValidityCheckerDatabase.h:173,201): areal/areal disjoint = no boundary intersection
 * (touching counts as intersecting) and neither ring has a point within the other. */
int orc_polygons_disjoint(int n1, const double (*p1)[2], int n2, const double (*p2)[2])
{
    int i, j;
    for (i = 0; i < n1; ++i)
        for (j = 0; j < n2; ++j)
            if (segments_touch(p1[i], p1[(i + 1) % n1], p2[j], p2[(j + 1) % n2])) return 0;
    if (point_in_polygon(n2, p2, p1[0])) return 0;
    if (point_in_polygon(n1, p1, p2[0])) return 0;
    return 1;
}

/* Separating-axis gap of two rectangles given as centre + two orthonormal axes + half extents. */
static double sat_gap(const double ca[2], const double a1[2], const double a2[2], double ha1, double ha2,
                      const double cb[2], const double b1[2], const double b2[2], double hb1, double hb2)
{
    const double d[2] = {cb[0] - ca[0], cb[1] - ca[1]};
    const double *axes[4] = {a1, a2, b1, b2};
    double best = -INFINITY;
    int i;
    for (i = 0; i < 4; ++i) {
        const double *n = axes[i];
        const double ra = ha1 * fabs(a1[0] * n[0] + a1[1] * n[1]) + ha2 * fabs(a2[0] * n[0] + a2[1] * n[1]);
        const double rb = hb1 * fabs(b1[0] * n[0] + b1[1] * n[1]) + hb2 * fabs(b2[0] * n[0] + b2[1] * n[1]);
        const double g = fabs(d[0] * n[0] + d[1] * n[1]) - (ra + rb);
        if (g > best) best = g;
    }
    return best;
}

static void robot_axes(double theta, double e1[2], double e2[2])
{
    /* local x axis -> (cos, -sin), local y axis -> (sin, cos): columns of the :152-155 matrix */
    const double c = cos(theta), s = sin(theta);
    e1[0] = c; e1[1] = -s;
    e2[0] = s; e2[1] = c;
}

double orc_sat_gap_robot_robot(double la, double wa, const double sa[3], double lb, double wb, const double sb[3])
{
    double a1[2], a2[2], b1[2], b2[2];
    robot_axes(sa[2], a1, a2);
    robot_axes(sb[2], b1, b2);
    return sat_gap(sa, a1, a2, la / 2, wa / 2, sb, b1, b2, lb / 2, wb / 2);
}

double orc_sat_gap_robot_obstacle(double l, double wd, const double s[3], const double obs[4])
{
    double a1[2], a2[2];
    const double b1[2] = {1, 0}, b2[2] = {0, 1};
    const double cb[2] = {obs[0] + obs[2] / 2, obs[1] + obs[3] / 2};
    robot_axes(s[2], a1, a2);
    return sat_gap(s, a1, a2, l / 2, wd / 2, cb, b1, b2, obs[2] / 2, obs[3] / 2);
}

static double bounding_radius(double length, double width)
{
    /* Robot.h:99, Obstacle.h:102 */
    return sqrt(pow(length / 2, 2) + pow(width / 2, 2));
}

/* performQuickCollisionCheck, StateValidityCheckerDatabase.h:134-141: 1 = no collision possible */
static int quick_check(const double a[2], const double b[2], double rad1, double rad2)
{
    const double distance = sqrt(pow(a[0] - b[0], 2) + pow(a[1] - b[1], 2));
    return distance > (rad1 + rad2);
}

/* ------------------------------------------------------------------------------------------
 * isValid: includes/StateValidityCheckerDatabase.h:54-75
 * ---------------------------------------------------------------------------------------- */
int orc_is_valid(const orc_world *w, int robot, const double q[5], double *margin)
{
    double slack;
    const double L = w->robot_dims[robot][0], W = w->robot_dims[robot][1];
    const double rad1 = bounding_radius(L, W);
    const double s[3] = {q[0], q[1], q[4]};
    int o, ok = 1;
    double m;
    if (!orc_satisfies_bounds(w, q, &slack)) { /* :57 */
        if (margin) *margin = slack;
        return 0;
    }
    m = slack;
    for (o = 0; o < w->n_obstacles; ++o) { /* :62 */
        const double *ob = w->obstacles[o];
        const double centre[2] = {ob[0] + ob[2] / 2, ob[1] + ob[3] / 2}; /* Obstacle.h:85-86 */
        const double orad = bounding_radius(ob[2], ob[3]);
        if (margin) {
            const double g = orc_sat_gap_robot_obstacle(L, W, s, ob);
            if (g < m) m = g;
        }
        if (!quick_check(s, centre, rad1, orad)) { /* :66 */
            double rc[4][2], oc[4][2];
            orc_robot_corners(L, W, q[0], q[1], q[4], rc); /* :182-192 */
            obstacle_corners(ob, oc);
            if (!orc_polygons_disjoint(4, rc, 4, oc)) { /* :201 */
                ok = 0;
                if (!margin) return 0; /* :70-71 */
            }
        }
    }
    if (margin) *margin = m;
    return ok;
}

/* ------------------------------------------------------------------------------------------
 * areStatesValid, non-composed branch: includes/StateValidityCheckerDatabase.h:108-124
 * ---------------------------------------------------------------------------------------- */
int orc_pair_valid(const orc_world *w, int r1, const double s1[3], int r2, const double s2[3], double *margin)
{
    const double L1 = w->robot_dims[r1][0], W1 = w->robot_dims[r1][1];
    const double L2 = w->robot_dims[r2][0], W2 = w->robot_dims[r2][1];
    const double rad1 = bounding_radius(L1, W1), rad2 = bounding_radius(L2, W2); /* :112 */
    if (margin) *margin = orc_sat_gap_robot_robot(L1, W1, s1, L2, W2, s2);
    if (!quick_check(s1, s2, rad1, rad2)) { /* :117 */
        double a[4][2], b[4][2];
        orc_robot_corners(L1, W1, s1[0], s1[1], s1[2], a); /* :146-157 */
        orc_robot_corners(L2, W2, s2[0], s2[1], s2[2], b); /* :159-170 */
        if (!orc_polygons_disjoint(4, a, 4, b)) return 0;   /* :173 */
    }
    return 1;
}

/* ------------------------------------------------------------------------------------------
 * [UPSTREAM] oc::SpaceInformation::propagateWhileValid(state, control, steps, result)
 * ---------------------------------------------------------------------------------------- */
int orc_propagate_while_valid(const orc_world *w, int robot, const double q0[5], const double u[2], int steps,
                              double *seg, double q_final[5])
{
    double cur[5], next[5];
    int i;
    memcpy(cur, q0, sizeof cur);
    for (i = 0; i < steps; ++i) {
        orc_propagate(w, cur, u, w->step_size, next);
        if (!orc_is_valid(w, robot, next, NULL)) break;
        memcpy(cur, next, sizeof cur);
        if (seg) memcpy(seg + 5 * i, cur, sizeof cur);
    }
    if (q_final) memcpy(q_final, cur, sizeof cur);
    return i;
}

/* ------------------------------------------------------------------------------------------
 * Batched drivers
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    void (*fn)(void *, int64_t, int64_t);
    void *arg;
    int64_t lo, hi;
} task_t;

static void *task_main(void *p)
{
    task_t *t = (task_t *)p;
    t->fn(t->arg, t->lo, t->hi);
    return NULL;
}

static void parallel_for(int64_t n, int n_threads, void (*fn)(void *, int64_t, int64_t), void *arg)
{
    int t;
    pthread_t *th;
    task_t *tasks;
    if (n_threads <= 1 || n < 2) {
        fn(arg, 0, n);
        return;
    }
    if (n_threads > n) n_threads = (int)n;
    th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)n_threads);
    tasks = (task_t *)malloc(sizeof(task_t) * (size_t)n_threads);
    for (t = 0; t < n_threads; ++t) {
        tasks[t].fn = fn;
        tasks[t].arg = arg;
        tasks[t].lo = n * t / n_threads;
        tasks[t].hi = n * (t + 1) / n_threads;
        pthread_create(&th[t], NULL, task_main, &tasks[t]);
    }
    for (t = 0; t < n_threads; ++t) pthread_join(th[t], NULL);
    free(th);
    free(tasks);
}

typedef struct {
    const orc_world *w;
    int robot;
    int64_t n, n_parents;
    const float *parents;
    const int32_t *parent_idx;
    const float *controls;
    const int32_t *steps;
    int max_steps;
    double *seg_out, *final_out;
    int32_t *valid_steps;
    int64_t *calls; /* per-chunk accumulation through atomic add */
    pthread_mutex_t mu;
} prop_job;

static void prop_range(void *p, int64_t lo, int64_t hi)
{
    prop_job *j = (prop_job *)p;
    int64_t i, calls = 0;
    double seg[5 * 64];
    for (i = lo; i < hi; ++i) {
        const int64_t pi = j->parent_idx ? j->parent_idx[i] : (j->n_parents == 1 ? 0 : i);
        double q0[5], qf[5];
        const double u[2] = {(double)j->controls[i], (double)j->controls[j->n + i]};
        int c, s, nv, st = j->steps ? j->steps[i] : j->max_steps;
        if (st > j->max_steps) st = j->max_steps;
        if (st > 64) st = 64;
        for (c = 0; c < 5; ++c) q0[c] = (double)j->parents[c * j->n_parents + pi];
        nv = orc_propagate_while_valid(j->w, j->robot, q0, u, st, seg, qf);
        calls += (nv < st) ? nv + 1 : nv;
        if (j->valid_steps) j->valid_steps[i] = nv;
        if (j->seg_out)
            for (s = 0; s < nv; ++s)
                for (c = 0; c < 5; ++c) j->seg_out[((int64_t)c * j->max_steps + s) * j->n + i] = seg[5 * s + c];
        if (j->final_out)
            for (c = 0; c < 5; ++c) j->final_out[c * j->n + i] = qf[c];
    }
    pthread_mutex_lock(&j->mu);
    *j->calls += calls;
    pthread_mutex_unlock(&j->mu);
}

int64_t orc_propagate_batch(const orc_world *w, int robot, int64_t n, int64_t n_parents, const float *parents,
                            const int32_t *parent_idx, const float *controls, const int32_t *steps,
                            int max_steps, double *seg_out, double *final_out, int32_t *valid_steps,
                            int n_threads)
{
    int64_t calls = 0;
    prop_job j;
    j.w = w; j.robot = robot; j.n = n; j.n_parents = n_parents; j.parents = parents;
    j.parent_idx = parent_idx; j.controls = controls; j.steps = steps; j.max_steps = max_steps;
    j.seg_out = seg_out; j.final_out = final_out; j.valid_steps = valid_steps; j.calls = &calls;
    pthread_mutex_init(&j.mu, NULL);
    parallel_for(n, n_threads, prop_range, &j);
    pthread_mutex_destroy(&j.mu);
    return calls;
}

typedef struct {
    const orc_world *w;
    int robot;
    int64_t n;
    const float *states;
    uint8_t *valid;
    double *margin;
} val_job;

static void val_range(void *p, int64_t lo, int64_t hi)
{
    val_job *j = (val_job *)p;
    int64_t i;
    for (i = lo; i < hi; ++i) {
        double q[5], m;
        int c, ok;
        for (c = 0; c < 5; ++c) q[c] = (double)j->states[c * j->n + i];
        ok = orc_is_valid(j->w, j->robot, q, j->margin ? &m : NULL);
        if (j->valid) j->valid[i] = (uint8_t)ok;
        if (j->margin) j->margin[i] = m;
    }
}

int64_t orc_validity_batch(const orc_world *w, int robot, int64_t n, const float *states, uint8_t *valid,
                           double *margin, int n_threads)
{
    val_job j;
    j.w = w; j.robot = robot; j.n = n; j.states = states; j.valid = valid; j.margin = margin;
    parallel_for(n, n_threads, val_range, &j);
    return n;
}

/* [UPSTREAM] KCBS::findConflicts: paths interpolated to one state per propagation step, shorter
 * paths padded with their last state; scan k ascending, then r1 ascending, then r2 > r1
 * ascending; the first pair failing areStatesValid is the conflict; it is extended forward
 * while that same pair keeps colliding. */
static void traj_state(const float *traj, int T, int r, int k, const int32_t *lengths, double s[3])
{
    int kk = k;
    if (lengths && kk > lengths[r] - 1) kk = lengths[r] - 1;
    if (kk < 0) kk = 0;
    s[0] = (double)traj[((int64_t)r * 3 + 0) * T + kk];
    s[1] = (double)traj[((int64_t)r * 3 + 1) * T + kk];
    s[2] = (double)traj[((int64_t)r * 3 + 2) * T + kk];
}

static int64_t first_conflict_impl(const orc_world *w, int R, int T, const float *traj, const int32_t *lengths,
                                   orc_conflict *out, double *min_margin, int full_scan, const int32_t *group)
{
    int k, r1, r2, found = 0;
    int64_t evals = 0;
    double mm = INFINITY;
    out->r1 = out->r2 = out->k_start = out->k_end = -1;
    for (k = 0; k < T; ++k) {
        for (r1 = 0; r1 < R; ++r1) {
            double s1[3];
            traj_state(traj, T, r1, k, lengths, s1);
            for (r2 = r1 + 1; r2 < R; ++r2) {
                double s2[3], m;
                int ok;
                if (group && group[r1] == group[r2]) continue;
                traj_state(traj, T, r2, k, lengths, s2);
                ok = orc_pair_valid(w, r1, s1, r2, s2, min_margin ? &m : NULL);
                ++evals;
                if (min_margin && !found && fabs(m) < mm) mm = fabs(m);
                if (!ok && !found) {
                    int kk = k;
                    found = 1;
                    out->r1 = r1;
                    out->r2 = r2;
                    out->k_start = k;
                    while (kk + 1 < T) {
                        double a[3], b[3];
                        traj_state(traj, T, r1, kk + 1, lengths, a);
                        traj_state(traj, T, r2, kk + 1, lengths, b);
                        if (orc_pair_valid(w, r1, a, r2, b, NULL)) break;
                        ++kk;
                    }
                    out->k_end = kk;
                    if (!full_scan) {
                        if (min_margin) *min_margin = mm;
                        return evals;
                    }
                }
            }
        }
    }
    if (min_margin) *min_margin = mm;
    return evals;
}

int64_t orc_first_conflict(const orc_world *w, int R, int T, const float *traj, const int32_t *lengths,
                           orc_conflict *out, double *min_margin)
{
    return first_conflict_impl(w, R, T, traj, lengths, out, min_margin, 0, NULL);
}

int64_t orc_first_conflict_groups(const orc_world *w, int R, int T, const float *traj, const int32_t *lengths,
                                  const int32_t *group, orc_conflict *out, double *min_margin)
{
    return first_conflict_impl(w, R, T, traj, lengths, out, min_margin, 0, group);
}

typedef struct {
    const orc_world *w;
    int R, T, full_scan;
    const float *traj;
    const int32_t *lengths;
    orc_conflict *out;
    int64_t evals;
    pthread_mutex_t mu;
} fc_job;

static void fc_range(void *p, int64_t lo, int64_t hi)
{
    fc_job *j = (fc_job *)p;
    int64_t i, ev = 0;
    for (i = lo; i < hi; ++i)
        ev += first_conflict_impl(j->w, j->R, j->T, j->traj + i * (int64_t)j->R * 3 * j->T,
                                  j->lengths ? j->lengths + i * j->R : NULL, &j->out[i], NULL, j->full_scan, NULL);
    pthread_mutex_lock(&j->mu);
    j->evals += ev;
    pthread_mutex_unlock(&j->mu);
}

int64_t orc_first_conflict_batch(const orc_world *w, int P, int R, int T, const float *traj, const int32_t *lengths,
                                 orc_conflict *out, int full_scan, int n_threads)
{
    fc_job j;
    j.w = w; j.R = R; j.T = T; j.full_scan = full_scan; j.traj = traj; j.lengths = lengths; j.out = out;
    j.evals = 0;
    pthread_mutex_init(&j.mu, NULL);
    parallel_for(P, n_threads, fc_range, &j);
    pthread_mutex_destroy(&j.mu);
    return j.evals;
}

/* ------------------------------------------------------------------------------------------
 * Merged pair: StatePropagatorDatabase.h:77-171, StateValidityCheckerDatabase.h:214-261
 * ---------------------------------------------------------------------------------------- */
int orc_pair_is_valid(const orc_world *w, int r1, int r2, const double q[10], double *margin)
{
    double m1, m2, mp;
    int ok = 1;
    const double sa[3] = {q[0], q[1], q[4]}, sb[3] = {q[5], q[6], q[9]};
    /* si_->satisfiesBounds on the 10-d compound space (:228) followed by the pair test (:237-241) and both cars
     * against the obstacles (:244-258); orc_is_valid covers bounds + obstacles of one car */
    if (!orc_is_valid(w, r1, q, &m1)) ok = 0;
    if (!orc_is_valid(w, r2, q + 5, &m2)) ok = 0;
    if (!orc_pair_valid(w, r1, sa, r2, sb, &mp)) ok = 0;
    if (margin) {
        double m = m1 < m2 ? m1 : m2;
        *margin = m < mp ? m : mp;
    }
    return ok;
}

void orc_pair_propagate(const orc_world *w, const double q_in[10], const double u[4], double duration, const double goals[4],
                        double hold_radius, double q_out[10])
{
    int c, i;
    for (c = 0; c < 2; ++c) {
        /* TwoSecondOrderCarsODE integrates the cars independently (:87-101): same RK4 per car */
        orc_propagate(w, q_in + 5 * c, u + 2 * c, duration, q_out + 5 * c);
        /* override when the car was in its goal region before the step (:137-152) */
        if (sqrt(pow(q_in[5 * c] - goals[2 * c], 2) + pow(q_in[5 * c + 1] - goals[2 * c + 1], 2)) < hold_radius)
            for (i = 0; i < 5; ++i) q_out[5 * c + i] = q_in[5 * c + i];
    }
}

int orc_pair_propagate_while_valid(const orc_world *w, int r1, int r2, const double q0[10], const double u[4], int steps,
                                   const double goals[4], double hold_radius, double *seg, double q_final[10])
{
    double cur[10], next[10];
    int i;
    memcpy(cur, q0, sizeof cur);
    for (i = 0; i < steps; ++i) {
        orc_pair_propagate(w, cur, u, w->step_size, goals, hold_radius, next);
        if (!orc_pair_is_valid(w, r1, r2, next, NULL)) break;
        memcpy(cur, next, sizeof cur);
        if (seg) memcpy(seg + 10 * i, cur, sizeof cur);
    }
    if (q_final) memcpy(q_final, cur, sizeof cur);
    return i;
}

typedef struct {
    const orc_world *w;
    int r1, r2;
    int64_t n, n_parents;
    const float *parents;
    const int32_t *parent_idx;
    const float *controls;
    const int32_t *steps;
    int max_steps;
    const double *goals;
    double hold_radius;
    double *seg_out, *final_out, *hold_margin;
    int32_t *valid_steps;
    int64_t calls;
    pthread_mutex_t mu;
} pair_job;

static void pair_range(void *p, int64_t lo, int64_t hi)
{
    pair_job *j = (pair_job *)p;
    int64_t i, calls = 0;
    double seg[10 * 64];
    for (i = lo; i < hi; ++i) {
        const int64_t pi = j->parent_idx ? j->parent_idx[i] : (j->n_parents == 1 ? 0 : i);
        double q0[10], qf[10], u[4];
        int c, s, nv, st = j->steps ? j->steps[i] : j->max_steps;
        if (st > j->max_steps) st = j->max_steps;
        if (st > 64) st = 64;
        for (c = 0; c < 10; ++c) q0[c] = (double)j->parents[c * j->n_parents + pi];
        for (c = 0; c < 4; ++c) u[c] = (double)j->controls[c * j->n + i];
        nv = orc_pair_propagate_while_valid(j->w, j->r1, j->r2, q0, u, st, j->goals, j->hold_radius, seg, qf);
        calls += (nv < st) ? nv + 1 : nv;
        if (j->valid_steps) j->valid_steps[i] = nv;
        if (j->seg_out)
            for (s = 0; s < nv; ++s)
                for (c = 0; c < 10; ++c) j->seg_out[((int64_t)c * j->max_steps + s) * j->n + i] = seg[10 * s + c];
        if (j->final_out)
            for (c = 0; c < 10; ++c) j->final_out[c * j->n + i] = qf[c];
        if (j->hold_margin) {
            /* how close any freeze decision of this sample came to the hold radius */
            double mm = INFINITY;
            const double *cur = q0;
            for (s = 0; s <= nv && s < st; ++s) {
                for (c = 0; c < 2; ++c) {
                    const double d = sqrt(pow(cur[5 * c] - j->goals[2 * c], 2) + pow(cur[5 * c + 1] - j->goals[2 * c + 1], 2));
                    if (fabs(d - j->hold_radius) < mm) mm = fabs(d - j->hold_radius);
                }
                if (s < nv) cur = seg + 10 * s;
            }
            j->hold_margin[i] = mm;
        }
    }
    pthread_mutex_lock(&j->mu);
    j->calls += calls;
    pthread_mutex_unlock(&j->mu);
}

int64_t orc_pair_propagate_batch(const orc_world *w, int r1, int r2, int64_t n, int64_t n_parents, const float *parents,
                                 const int32_t *parent_idx, const float *controls, const int32_t *steps, int max_steps,
                                 const double goals[4], double hold_radius, double *seg_out, double *final_out,
                                 int32_t *valid_steps, double *hold_margin, int n_threads)
{
    pair_job j;
    j.w = w; j.r1 = r1; j.r2 = r2; j.n = n; j.n_parents = n_parents; j.parents = parents; j.parent_idx = parent_idx;
    j.controls = controls; j.steps = steps; j.max_steps = max_steps; j.goals = goals; j.hold_radius = hold_radius;
    j.seg_out = seg_out; j.final_out = final_out; j.hold_margin = hold_margin; j.valid_steps = valid_steps; j.calls = 0;
    pthread_mutex_init(&j.mu, NULL);
    parallel_for(n, n_threads, pair_range, &j);
    pthread_mutex_destroy(&j.mu);
    return j.calls;
}

typedef struct {
    const orc_world *w;
    int r1, r2;
    int64_t n;
    const float *states;
    uint8_t *valid;
    double *margin;
} pval_job;

static void pval_range(void *p, int64_t lo, int64_t hi)
{
    pval_job *j = (pval_job *)p;
    int64_t i;
    for (i = lo; i < hi; ++i) {
        double q[10], m;
        int c, ok;
        for (c = 0; c < 10; ++c) q[c] = (double)j->states[c * j->n + i];
        ok = orc_pair_is_valid(j->w, j->r1, j->r2, q, &m);
        if (j->valid) j->valid[i] = (uint8_t)ok;
        if (j->margin) j->margin[i] = m;
    }
}

int64_t orc_pair_validity_batch(const orc_world *w, int r1, int r2, int64_t n, const float *states, uint8_t *valid,
                                double *margin, int n_threads)
{
    pval_job j;
    j.w = w; j.r1 = r1; j.r2 = r2; j.n = n; j.states = states; j.valid = valid; j.margin = margin;
    parallel_for(n, n_threads, pval_range, &j);
    return n;
}
