// kcbs_device.cuh -- device-side scene constants and the geometric predicates shared by the
// three kernels (propagate / validity / conflict).  sm_100a only.
//
// Reference semantics restated here (file:line relative to the K-CBS-Demos tree):
//   bounds            StateSpaceDatabase.h:49-59 + OMPL RealVectorStateSpace::satisfiesBounds
//   broad phase       StateValidityCheckerDatabase.h:134-141
//   robot footprint   StateValidityCheckerDatabase.h:152-155 / :188-191  (rotation by -theta)
//   narrow phase      boost::geometry::disjoint at :173 / :201 == SAT with touching = collision
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/kcbs_b200.h"

namespace kcbs {

// Passed to every kernel as a __grid_constant__ parameter: it lives in the constant bank, so the
// warp-uniform obstacle loop reads it through the uniform datapath (no LSU traffic).
struct DevWorld {
    // Exact float thresholds: a float value f passes the reference's double test
    // (f - DBL_EPSILON <= hi && f + DBL_EPSILON >= lo) iff lo_t <= f <= hi_t.  Computed on the host.
    float lo_t[4];
    float hi_t[4];
    float th_lo_t, th_hi_t;  // SO2StateSpace::satisfiesBounds thresholds
    float h;                 // RK4 sub-step (ODEBasicSolver intStep)
    float step;              // propagation step size
    float inv_len;           // 1 / car_length
    int n_sub;               // RK4 sub-steps per propagation step (integrate_const rule)
    int n_obs;
    int n_robots;
    float max_rad;           // largest robot bounding radius
    // Obstacle cull grid (built on the host at kcbs_create, lives in device memory): cell (ix, iy) holds a
    // 64-bit mask of the obstacles whose rectangle, grown by max_rad, overlaps the cell.
    int gdim;                // cells per side
    float gx0, gy0;          // grid origin = lower workspace bound
    float ginvx, ginvy;      // cells per metre
    const unsigned long long *grid;  // [gdim * gdim]
    // Fine three-valued obstacle grid (built on the host at kcbs_create): one 16-bit code per cell, see cell_code().
    int fdim;                        // cells per side (kFineDim)
    float finvx, finvy;              // fine cells per metre
    float foffx, foffy;              // -origin * cells per metre: cell = (int)fma(x, finvx, foffx)
    const unsigned short *fine;      // [kFineRows][kFinePitch], see fine_cell()
    int sym_bounds;                  // v, phi and theta thresholds are symmetric about 0 (one compare each)
    const float4 *obs4;              // [n_obs] (cx, cy, hx, hy), same values as the arrays below
    float obs_cx[KCBS_MAX_OBSTACLES];  // Obstacle.h:85-86 centre
    float obs_cy[KCBS_MAX_OBSTACLES];
    float obs_hx[KCBS_MAX_OBSTACLES];  // half extents
    float obs_hy[KCBS_MAX_OBSTACLES];
    float rob_hl[KCBS_MAX_ROBOTS];     // Robot.h:85-88 half length / half width
    float rob_hw[KCBS_MAX_ROBOTS];
    float rob_rad[KCBS_MAX_ROBOTS];    // Robot.h:99
};

constexpr float kPi = 3.14159265358979323846f;
constexpr float kTwoPi = 6.28318530717958647692f;

// SO2StateSpace::enforceBounds restated for float: result in [-pi, pi).
__device__ __forceinline__ float wrap_angle(float th)
{
    if (th >= kPi)
        th -= kTwoPi;
    else if (th < -kPi)
        th += kTwoPi;
    if (!(th >= -kPi && th < kPi)) {  // more than one turn away: rare general path
        th = fmodf(th, kTwoPi);
        if (th < -kPi)
            th += kTwoPi;
        else if (th >= kPi)
            th -= kTwoPi;
    }
    return th;
}

__device__ __forceinline__ bool in_bounds(const DevWorld &w, float x, float y, float v, float phi, float th)
{
    bool ok = (x >= w.lo_t[0]) & (x <= w.hi_t[0]);
    ok &= (y >= w.lo_t[1]) & (y <= w.hi_t[1]);
    ok &= (v >= w.lo_t[2]) & (v <= w.hi_t[2]);
    ok &= (phi >= w.lo_t[3]) & (phi <= w.hi_t[3]);
    ok &= (th >= w.th_lo_t) & (th <= w.th_hi_t);
    return ok;
}

// Same test when the v, phi and theta thresholds are symmetric about zero (DevWorld::sym_bounds; true for every demo).
__device__ __forceinline__ bool in_bounds_sym(const DevWorld &w, float x, float y, float v, float phi, float th)
{
    bool ok = (x >= w.lo_t[0]) & (x <= w.hi_t[0]);
    ok &= (y >= w.lo_t[1]) & (y <= w.hi_t[1]);
    ok &= (fabsf(v) <= w.hi_t[2]) & (fabsf(phi) <= w.hi_t[3]) & (fabsf(th) <= w.th_hi_t);
    return ok;
}

// SAT gap, robot rectangle (half extents hl, hw, centre (x, y), axes e1 = (c, -s), e2 = (s, c)) against
// an axis-aligned rectangle (centre (cx, cy), half extents (hx, hy)).  > 0 separated.
__device__ __forceinline__ float sat_gap_obstacle(float x, float y, float c, float s, float hl, float hw, float cx,
                                                  float cy, float hx, float hy)
{
    const float dx = cx - x, dy = cy - y;
    const float ac = fabsf(c), as = fabsf(s);
    const float g0 = fabsf(dx) - (hx + fmaf(hl, ac, hw * as));              // world x axis
    const float g1 = fabsf(dy) - (hy + fmaf(hl, as, hw * ac));              // world y axis
    const float g2 = fabsf(fmaf(dx, c, -dy * s)) - (hl + fmaf(hx, ac, hy * as));  // robot e1
    const float g3 = fabsf(fmaf(dx, s, dy * c)) - (hw + fmaf(hx, as, hy * ac));   // robot e2
    return fmaxf(fmaxf(g0, g1), fmaxf(g2, g3));
}

// SAT gap of two robot rectangles, each rotated by -theta (StateValidityCheckerDatabase.h:152-168).
__device__ __forceinline__ float sat_gap_robots(float xa, float ya, float ca, float sa, float hla, float hwa, float xb,
                                                float yb, float cb, float sb, float hlb, float hwb)
{
    const float dx = xb - xa, dy = yb - ya;
    // |a1.b1| = |a2.b2| = |cos(tb - ta)|, |a1.b2| = |a2.b1| = |sin(tb - ta)|
    const float cd = fabsf(fmaf(ca, cb, sa * sb));
    const float sd = fabsf(fmaf(ca, sb, -sa * cb));
    const float g0 = fabsf(fmaf(dx, ca, -dy * sa)) - (hla + fmaf(hlb, cd, hwb * sd));  // a1 = (ca, -sa)
    const float g1 = fabsf(fmaf(dx, sa, dy * ca)) - (hwa + fmaf(hlb, sd, hwb * cd));   // a2 = (sa, ca)
    const float g2 = fabsf(fmaf(dx, cb, -dy * sb)) - (hlb + fmaf(hla, cd, hwa * sd));  // b1
    const float g3 = fabsf(fmaf(dx, sb, dy * cb)) - (hwb + fmaf(hla, sd, hwa * cd));   // b2
    return fmaxf(fmaxf(g0, g1), fmaxf(g2, g3));
}

// Obstacle part of isValid (StateValidityCheckerDatabase.h:62-73).
// The reference's circle test only skips obstacles whose SAT gap is positive anyway, so a cheaper
// conservative cull (obstacle AABB grown by the robot's bounding radius) followed by the exact SAT
// gives the same boolean.  cos/sin of the heading are only evaluated when some obstacle is near.
__device__ __forceinline__ bool obstacles_clear(const DevWorld &w, float x, float y, float th, float hl, float hw,
                                                float rad)
{
    bool ok = true, have = false;
    float c = 1.0f, s = 0.0f;
    for (int o = 0; o < w.n_obs; ++o) {
        const float cx = w.obs_cx[o], cy = w.obs_cy[o], hx = w.obs_hx[o], hy = w.obs_hy[o];
        if (fabsf(cx - x) <= hx + rad && fabsf(cy - y) <= hy + rad) {
            if (!have) {
                sincosf(th, &s, &c);
                have = true;
            }
            ok &= sat_gap_obstacle(x, y, c, s, hl, hw, cx, cy, hx, hy) > 0.0f;
        }
    }
    return ok;
}

constexpr int kGridDim = 32;

// Obstacle part of isValid with the cull grid: only obstacles flagged for the state's cell are tested.
// `grid` / `obs` may point to shared memory (k_validity stages them) or to global memory (k_propagate).
__device__ __forceinline__ bool obstacles_clear_grid(const DevWorld &w, const unsigned long long *grid,
                                                     const float4 *obs, float x, float y, float th, float hl,
                                                     float hw)
{
    const int ix = min(w.gdim - 1, max(0, (int)((x - w.gx0) * w.ginvx)));
    const int iy = min(w.gdim - 1, max(0, (int)((y - w.gy0) * w.ginvy)));
    unsigned long long m = grid[iy * w.gdim + ix];
    if (m == 0ull) return true;
    float s, c;
    sincosf(th, &s, &c);
    const float ac = fabsf(c), as = fabsf(s);
    const float ex = fmaf(hl, ac, hw * as), ey = fmaf(hl, as, hw * ac);  // robot extent along world x / y
    bool ok = true;
    while (m) {
        const int o = __ffsll((long long)m) - 1;
        m &= m - 1;
        const float4 ob = obs[o];
        const float dx = ob.x - x, dy = ob.y - y;
        // separating axis found <=> SAT gap > 0 (touching counts as collision)
        const bool sep = (fabsf(dx) > ob.z + ex) | (fabsf(dy) > ob.w + ey) |
                         (fabsf(fmaf(dx, c, -dy * s)) > hl + fmaf(ob.z, ac, ob.w * as)) |
                         (fabsf(fmaf(dx, s, dy * c)) > hw + fmaf(ob.z, as, ob.w * ac));
        ok &= sep;
    }
    return ok;
}

// ---- fine obstacle grid ------------------------------------------------------------------------------------------
// Every cell of a kFineDim x kFineDim grid over the workspace is classified on the host, in double, against every
// obstacle for ANY heading of ANY robot of the scene:
//   free   the cell is farther from the obstacle than the largest bounding radius: isValid cannot fail on it
//   hit    every point of the cell is closer to the obstacle than the smallest inscribed radius min(L, W) / 2: the
//          footprint contains that disc, so the exact test (StateValidityCheckerDatabase.h:180-205) reports a collision
//          whatever the heading
//   maybe  the exact rectangle test decides
// Cell code: 0 = all obstacles free; kCellHit = some obstacle hits; kCellMany = more than two maybes (decided through
// the coarse 64-bit mask grid); otherwise the one or two maybe obstacles as (index + 1) in the low / high byte.
constexpr int kFineDim = 128;
// Device layout: one extra row and column (copies of the last ones) so that a state sitting exactly on the upper bound
// indexes without a clamp; pitch padded to a multiple of 16 bytes.
constexpr int kFineRows = kFineDim + 1, kFinePitch = 136;
constexpr unsigned kCellHit = 0xffffu, kCellMany = 0xfffeu;

__device__ __forceinline__ int fine_cell(const DevWorld &w, float x, float y)
{
    const int ix = min(kFineDim - 1, max(0, (int)fmaf(x, w.finvx, w.foffx)));
    const int iy = min(kFineDim - 1, max(0, (int)fmaf(y, w.finvy, w.foffy)));
    return iy * kFinePitch + ix;
}

// true when a separating axis exists between the robot rectangle (centre (x, y), heading cos / sin c, s, half extents
// hl, hw; ex, ey = its extents along world x / y) and the obstacle rectangle ob = (cx, cy, hx, hy).
__device__ __forceinline__ bool separated_from(const float4 ob, float x, float y, float c, float s, float ac, float as,
                                               float ex, float ey, float hl, float hw)
{
    const float dx = ob.x - x, dy = ob.y - y;
    return (fabsf(dx) > ob.z + ex) | (fabsf(dy) > ob.w + ey) |
           (fabsf(fmaf(dx, c, -dy * s)) > hl + fmaf(ob.z, ac, ob.w * as)) |
           (fabsf(fmaf(dx, s, dy * c)) > hw + fmaf(ob.z, as, ob.w * ac));
}

// Exact part of the obstacle test for a state whose cell code is neither 0 nor kCellHit.
__device__ __forceinline__ bool obstacles_clear_code(const DevWorld &w, unsigned code, const float4 *obs, float x, float y,
                                                     float th, float hl, float hw)
{
    if (code == kCellMany) return obstacles_clear_grid(w, w.grid, w.obs4, x, y, th, hl, hw);
    float s, c;
    sincosf(th, &s, &c);
    const float ac = fabsf(c), as = fabsf(s);
    const float ex = fmaf(hl, ac, hw * as), ey = fmaf(hl, as, hw * ac);
    bool ok = separated_from(obs[(code & 0xffu) - 1], x, y, c, s, ac, as, ex, ey, hl, hw);
    if (code >> 8) ok &= separated_from(obs[(code >> 8) - 1], x, y, c, s, ac, as, ex, ey, hl, hw);
    return ok;
}

// Obstacle part of isValid through the fine grid read from global memory (k_propagate, pair kernels).
__device__ __forceinline__ bool obstacles_clear_fine(const DevWorld &w, float x, float y, float th, float hl, float hw)
{
    const unsigned code = __ldg(w.fine + fine_cell(w, x, y));
    if (code == 0u) return true;
    if (code == kCellHit) return false;
    return obstacles_clear_code(w, code, w.obs4, x, y, th, hl, hw);
}

// One dynamic-obstacle constraint of the K-CBS low level: at absolute time index k in [k0, k0 + |len|) the planned robot must
// not overlap robot `other` placed at cons_states[plane * total + offset + (k - k0)]; len < 0 marks a persistent
// constraint whose last state keeps holding for every later time index.
struct ConsDesc {
    int other, k0, len, offset;
    float cx, cy, r;   // a circle around the positions of all its states (filled by the host): lets a sample that cannot
    int pad;           // come near them skip the constraint for all of its steps (cons_mask)
};

// Which of the first 32 constraints can matter to a sample at all: it starts at (x0, y0) at time index t0 and drives `steps`
// steps at a speed of at most v_max, so it stays within reach = v_max * steps * step of its parent; a constraint whose time
// window misses [t0 + 1, t0 + steps] or whose circle is farther away than reach + both bounding radii can never fail the
// exact test (constraints_clear applies the same circle test to the state itself first).  Constraints past the 32nd are
// always tested.
__device__ __forceinline__ unsigned cons_mask(const DevWorld &w, const ConsDesc *cons, int n_cons, int t0, int steps, float x0, float y0,
                                              float rad)
{
    unsigned mask = 0u;
    const float v_max = fmaxf(fabsf(w.lo_t[2]), fabsf(w.hi_t[2]));
    const float reach = v_max * ((float)steps * w.step) * 1.01f + 0.05f;
    const int nq = min(n_cons, 32);
    for (int q = 0; q < nq; ++q) {
        const int4 h = __ldg(reinterpret_cast<const int4 *>(cons + q));
        const float4 c = __ldg(reinterpret_cast<const float4 *>(cons + q) + 1);
        const bool in_time = t0 + steps >= h.y && (h.z < 0 || t0 + 1 < h.y + h.z);
        const float dx = c.x - x0, dy = c.y - y0, R = c.z + rad + w.rob_rad[h.x] + reach;
        if (in_time && fmaf(dy, dy, dx * dx) <= R * R) mask |= 1u << q;
    }
    return mask;
}

// Dynamic-obstacle constraints of the K-CBS low level (the fork tests `next` against the constraining robot's
// state at that time through areStatesValid, StateValidityCheckerDatabase.h:78-125).
__device__ __forceinline__ bool constraints_clear(const DevWorld &w, const ConsDesc *cons, int n_cons, const float *cons_states,
                                                  int cons_total, int k, float x, float y, float th, float hl, float hw, float rad,
                                                  unsigned mask = 0xffffffffu)
{
    bool ok = true, have = false;
    float c = 1.0f, s = 0.0f;
    for (int q = 0; q < n_cons; ++q) {
        if (q < 32 && !((mask >> q) & 1u)) continue;   // (cons_mask: out of this sample's reach, in time or in space)
        const int4 h = __ldg(reinterpret_cast<const int4 *>(cons + q));
        ConsDesc d;
        d.other = h.x; d.k0 = h.y; d.len = h.z; d.offset = h.w;
        int rel = k - d.k0;
        if (rel < 0) continue;
        if (d.len < 0)
            rel = min(rel, -d.len - 1);  // persistent constraint: its last state holds for every later time index
        else if (rel >= d.len)
            continue;
        const float *st = cons_states + d.offset + rel;
        const float ox = st[0], oy = st[cons_total], dx = ox - x, dy = oy - y;
        const float rs = rad + w.rob_rad[d.other];
        if (fmaf(dy, dy, dx * dx) > rs * rs * 1.000001f) continue;
        if (!have) {
            sincosf(th, &s, &c);
            have = true;
        }
        float oc, os;
        sincosf(st[2 * cons_total], &os, &oc);
        ok &= sat_gap_robots(x, y, c, s, hl, hw, ox, oy, oc, os, w.rob_hl[d.other], w.rob_hw[d.other]) > 0.0f;
    }
    return ok;
}

// Packed first-conflict key: lexicographic (k, r1, r2) order == integer order.
__host__ __device__ __forceinline__ unsigned long long pack_key(int k, int r1, int r2)
{
    return ((unsigned long long)(unsigned)k << 16) | ((unsigned long long)r1 << 8) | (unsigned long long)r2;
}
constexpr unsigned long long kNoConflict = ~0ull;

}  // namespace kcbs
