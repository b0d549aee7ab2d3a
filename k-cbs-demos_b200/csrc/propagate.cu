// propagate.cu -- K1: fused multi-step RK4 expansion kernel (one thread per control sample).
//
// Replaces, per sample, oc::SpaceInformation::propagateWhileValid driving
//   SecondOrderCarODE                      StatePropagatorDatabase.h:44-64
//   odeint runge_kutta4, dt = 0.01         (ODEBasicSolver, call site demo_*.cpp:111-112)
//   SecondOrderCarODEPostIntegration       StatePropagatorDatabase.h:67-74 (theta wrap per step)
//   homogeneous2ndOrderCarSystemSVC::isValid  StateValidityCheckerDatabase.h:54-75
//
// The kernel is classical RK4 with the reference's stages, evaluated with what the ODE's structure gives for free
// (car_dynamics.cuh, car_step2, has the derivation and the cost model):
//   * vdot = a and phidot = w are constant: the v / phi stage values are exact linear functions of time, stages 2 and
//     3 share v, phi, tan(phi) and k_theta, and stage 4 of a sub-step is stage 1 of the next one.
//   * tan(phi) at the half and the full sub-step follows from the start of the sub-step by the addition formula
//     tan(phi + d) = (T + t) / (1 - T t); the two denominators share one MUFU.RCP.  Re-anchored every propagation step.
//   * the four stage headings of a sub-step differ by at most 0.035 rad, so the weighted sum of the four stage velocities
//     collapses to ONE amplitude times (cos, sin) of ONE heading: one MUFU sin / cos per sample and sub-step.
//   * positions are accumulated as offsets from the parent and the heading as an offset from the step start, so FP32
//     rounding at |x| ~ 32 happens once per written state, not once per sub-step.
//   * every thread integrates TWO samples at once in packed FP32 pairs (fma / mul / add .f32x2, SASS FFMA2 ...):
//     21 packed instructions + 2 FMUL + 6 MUFU per sub-step of two samples.
//
// Work distribution: requested durations are U{1..10} (SimpleDirectedControlSampler), which would idle 45 % of the
// lanes of a warp; and because v, phi are linear in time, the step at which a sample leaves the v / phi bounds is known
// up front, so its effective duration is too (the doomed step is never integrated).  Each CTA (128 threads) therefore
// counting-sorts its 512 samples by effective duration in shared memory and every thread runs two packed pairs of
// neighbouring ranks: ranks (2t, 2t+1) (long) and then (510-2t, 511-2t) (short).  Both halves of a pair and the lanes
// of a warp then hold (almost) equal durations in both rounds and every thread of the CTA has the same total work, so
// neither lanes nor warps idle.  (Launches that cannot fill the GPU run one pair per thread, 256 samples per CTA.)
// Segments are staged in shared memory (x, y, theta per step; v and phi are recomputed from their exact linear form, or
// staged too when the packed layout leaves room) and leave in one coalesced pass: see the output modes below.
#include <cstddef>
#include <type_traits>

#include "car_dynamics.cuh"
#include "kernels.h"

namespace kcbs {

#ifndef KCBS_PROP_MINB
#define KCBS_PROP_MINB 4
#endif
#ifndef KCBS_PROP_THREADS
#define KCBS_PROP_THREADS 128
#endif
constexpr int kPropThreads = KCBS_PROP_THREADS;  // threads per CTA
constexpr int kStageSteps = 10;    // segments up to this many steps are staged in shared memory

// Output modes of the kernel.
//   kDirect  x, y, theta stored straight to the sample's slot of the dense segment buffer (horizons longer than
//            kStageSteps, final-state-only calls), v / phi and the zero rows by a coalesced epilogue
//   kDense   segments staged in shared memory, dense [5][max_steps][n] layout with zero rows past the last valid state
//   kPacked  segments staged in shared memory, CSR layout: sample i owns a run of rows from offsets[i] in five planes, as
//            many as it can have valid states at all (its effective duration, known before anything is integrated); the
//            first valid_steps[i] of them hold its states, the rest (a sample that left the workspace early) are zero.
//            Because the row count does not depend on the integration, a CTA gets its rows from the allocator (one
//            atomic, first come, first served) while it integrates, and the write-out is a straight, 16-byte vectorised
//            copy of the stage.
enum PropMode { kDirect = 0, kDense = 1, kPacked = 2 };

// ROUNDS = packed pairs per thread.  2: ranks (2t, 2t+1) (long) then (S-2-2t, S-1-2t) (short) -- every thread has the same
// total work, the choice for launches that fill the GPU.  1: one pair per thread, twice the threads per sample -- the
// choice for a launch that cannot fill the GPU on its own (one 64K-control expansion = 512 warps on 592 warp schedulers
// with ROUNDS = 2; a thread's 10-step chain of dependent RK4 sub-steps is the critical path, so more, shorter threads win).
template <int ROUNDS>
struct PropGeom {
    static constexpr int kSamples = 2 * ROUNDS * kPropThreads;   // samples per CTA
    static constexpr int kPer = 2 * ROUNDS;                      // samples per thread
    static constexpr int kCtasPerSm = 384 / kPropThreads;        // 12 warps per SM: what 150 registers per thread allow
};

// Shared-memory staging of one CTA's segments: x, y, theta of every (step, slot), plus the four numbers from which
// v and phi of any step follow exactly (they are linear in time).  70 KB per 512 samples: three CTAs per SM.
template <int SAMPLES>
struct __align__(16) PropStage {
    float xyt[3][kStageSteps * SAMPLES];   // kDense: [step][slot]; kPacked: [cap_off[slot] + step]
    float v0[SAMPLES], ua[SAMPLES], p0[SAMPLES], uw[SAMPLES];
    __align__(16) unsigned char nv[SAMPLES];
};

// Bookkeeping of the packed layout (static shared memory of the kPacked kernels).
template <int SAMPLES>
struct PackTables {
    unsigned short cap_off[SAMPLES + 4];   // slot -> first stage position (= first row inside the CTA's range) of its run
    int cap_total;                         // stage positions in use = rows of the CTA (before rounding up to 4)
    long long base;                        // first packed row of the CTA
};

// Where state j of the sample in `slot` lives in the stage, and where the sort pass stashes (x0, y0, theta0): in the
// sample's own LAST position, which the sample overwrites only when it completes its final step.  Component c of stage
// position p is flat[c * pstride + p].  kDense: three planes (x, y, theta) of kStageSteps * SAMPLES positions.  kPacked:
// when the CTA's rows fit into 3/5 of that (U{1..10} durations: 55 % on average) the same memory holds FIVE planes and
// v, phi are staged too (`five`), so the write-out is one straight copy; otherwise three planes, and v, phi are laid out
// in a second pass.
template <int MODE, int SAMPLES>
struct StageIndex {
    static constexpr int kPositions = kStageSteps * SAMPLES;
    static constexpr int kFivePositions = kPositions * 3 / 5;   // (a multiple of 4)
    const unsigned short *cap_off;
    const unsigned char *steps;
    float *flat;
    int pstride;
    bool five;
    __device__ __forceinline__ int at(int j, int slot) const
    {
        return MODE == kPacked ? (int)cap_off[slot] + j : j * SAMPLES + slot;
    }
    __device__ __forceinline__ int stash(int slot) const
    {
        return MODE == kPacked ? (int)cap_off[slot] + (int)steps[slot] - 1 : (kStageSteps - 1) * SAMPLES + slot;
    }
    __device__ __forceinline__ float &ref(int c, int pos) const { return flat[c * pstride + pos]; }
};

// tan(phi) of a parent state: accurate (it seeds a hundred sub-steps of the addition formula).
__device__ __forceinline__ float tan_phi(float phi)
{
    float sp, cp;
    sincosf(phi, &sp, &cp);
    return __fdividef(sp, cp);
}

// propagateWhileValid for the two samples i[0], i[1] (effective durations st[0], st[1]; st = 0 for a slot past the end
// of the batch): writes their segments, final states and valid-step counts.
// FIVE (packed layout only): the CTA's rows fit five stage planes; the plane stride is a compile-time constant either way,
// so that a store into plane c is the position plus an immediate.
template <bool HAS_OBS, bool HAS_CONS, int MODE, int SAMPLES, bool FIVE>
__device__ __forceinline__ void propagate_two(const DevWorld &w, const PropArgs &a, const long long i[2], const int st[2],
                                              PropStage<SAMPLES> *stage, const StageIndex<MODE, SAMPLES> &six,
                                              unsigned char *s_nv, const int slot[2], float tan_shared)
{
    constexpr bool STAGED = MODE != kDirect;
    constexpr int kPlane = FIVE ? StageIndex<MODE, SAMPLES>::kFivePositions : StageIndex<MODE, SAMPLES>::kPositions;
    float *const flat = six.flat;
    auto ref = [&](int c, int pos) -> float & { return flat[c * kPlane + pos]; };
    float x0[2], y0[2], v0[2], phi0[2], th0[2], ua[2], uw[2];
    int t0[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
        const long long pi = parent_of(a, i[e]);
        t0[e] = (HAS_CONS && a.parent_time) ? a.parent_time[pi] : 0;
        if (STAGED) {  // the sort pass left the sample's inputs in shared memory (slot -1: any slot, nothing is used)
            const int sl = max(slot[e], 0);
            if (MODE == kPacked && st[e] == 0) {
                // a sample without a single admissible step owns no stage position: its parent comes from global memory
                x0[e] = a.parents[0 * a.n_parents + pi];
                y0[e] = a.parents[1 * a.n_parents + pi];
                th0[e] = a.parents[4 * a.n_parents + pi];
            } else {
                const int at = six.stash(sl);
                x0[e] = ref(0, at);
                y0[e] = ref(1, at);
                th0[e] = ref(2, at);
            }
            v0[e] = stage->v0[sl];
            phi0[e] = stage->p0[sl];
            ua[e] = stage->ua[sl];
            uw[e] = stage->uw[sl];
            continue;
        }
        x0[e] = a.parents[0 * a.n_parents + pi];
        y0[e] = a.parents[1 * a.n_parents + pi];
        v0[e] = a.parents[2 * a.n_parents + pi];
        phi0[e] = a.parents[3 * a.n_parents + pi];
        th0[e] = a.parents[4 * a.n_parents + pi];
        ua[e] = a.controls[i[e]];
        uw[e] = a.controls[a.n + i[e]];
    }
    const float hl = w.rob_hl[a.robot], hw = w.rob_hw[a.robot];
    const CarConsts2 k = car_consts2(w, ua, uw, v0);
    unsigned cmask[2] = {0u, 0u};   // constraints that can matter to the sample at all (most samples: none)
    if (HAS_CONS) {
#pragma unroll
        for (int e = 0; e < 2; ++e)
            if (st[e] > 0) cmask[e] = cons_mask(w, a.cons, a.n_cons, t0[e], st[e], x0[e], y0[e], w.rob_rad[a.robot]);
    }

    // tan(phi) at the parent: accurate once per sample; from then on it advances step by step with the addition formula
    CarPair2 p;
    {
        float Ts[2] = {tan_shared, tan_shared};
        if (a.n_parents != 1) {   // (one shared parent: the thread took its tangent once, tan_phi)
#pragma unroll
            for (int e = 0; e < 2; ++e) Ts[e] = tan_phi(phi0[e]);
        }
        p.T = pk(Ts[0], Ts[1]);
        // the heading enters wrapped (the general wrap, once per sample): every step then moves it by less than pi, so
        // the per-step wrap is two compares
        p.TH = pk(wrap_angle(th0[0]), wrap_angle(th0[1]));
        p.offx = p.offy = pk1(0.0f);   // position offsets from the parents (in units of k.cpos)
    }
    const f32x2 X0 = pk(x0[0], x0[1]), Y0 = pk(y0[0], y0[1]), CP = pk1(k.cpos);
    float xl[2] = {x0[0], x0[1]}, yl[2] = {y0[0], y0[1]}, tl[2] = {th0[0], th0[1]};   // last valid state (direct path only)
    int nv[2] = {0, 0};
    bool alive[2] = {st[0] > 0, st[1] > 0};
    const long long plane = (long long)a.max_steps * a.n;
    const int jmax = max(st[0], st[1]);
    int at0[2] = {0, 0};
    if (STAGED) {
        at0[0] = six.at(0, max(slot[0], 0));
        at0[1] = six.at(0, max(slot[1], 0));
    }
    constexpr int kStride = MODE == kPacked ? 1 : SAMPLES;   // stage positions between consecutive steps of a sample

    float jf = 0.0f;   // (float)j: small integers are exact
    for (int j = 0; j < jmax && (alive[0] | alive[1]); ++j, jf += 1.0f) {
        const float ts = jf * w.step, t1 = (jf + 1.0f) * w.step;
        const f32x2 ang = w.n_sub == 10 ? car_step2<10>(k, 10, ts, p) : car_step2<0>(k, w.n_sub, ts, p);
        float xn[2], yn[2], tn[2];
        upk(fma2(CP, p.offx, X0), xn[0], xn[1]);
        upk(fma2(CP, p.offy, Y0), yn[0], yn[1]);
        upk(ang, tn[0], tn[1]);
        bool oks[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            // straight-line code with predicated commits, so that the two halves interleave; a half that has failed keeps
            // integrating from whatever it holds and nothing of it is used again
            tn[e] = tn[e] >= kPi ? tn[e] - kTwoPi : tn[e];     // wrap_angle for |theta| < 3 pi
            tn[e] = tn[e] < -kPi ? tn[e] + kTwoPi : tn[e];
            // v and phi of every step below st were tested when st was derived (same expressions): x, y, theta remain
            oks[e] = alive[e] & (xn[e] >= w.lo_t[0]) & (xn[e] <= w.hi_t[0]) & (yn[e] >= w.lo_t[1]) & (yn[e] <= w.hi_t[1]) &
                     (tn[e] >= w.th_lo_t) & (tn[e] <= w.th_hi_t);
        }
        if (HAS_OBS) {
            // Both samples' cell codes are looked up together (two independent loads), free and hit cells are decided by
            // them; the exact rectangle test of the "maybe" cells is ONE divergent region for both samples: its first
            // pass serves every lane's first candidate, a second pass only the lanes whose both samples need it (in an
            // obstacle scene some lane of a warp needs the test at nearly every step, so two separate regions both ran).
            unsigned code[2];
            bool need[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                // (a state inside the bounds indexes [0, kFineDim]: the table repeats its last row and column, no clamps)
                const float xc = oks[e] ? xn[e] : w.gx0, yc = oks[e] ? yn[e] : w.gy0;
                const int ix = (int)fmaf(xc, w.finvx, w.foffx), iy = (int)fmaf(yc, w.finvy, w.foffy);
                code[e] = __ldg(w.fine + iy * kFinePitch + ix);
            }
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                oks[e] &= code[e] != kCellHit;   // (a state outside the bounds looked the origin's cell up: oks stays false)
                need[e] = oks[e] & (code[e] != 0u);
            }
            if (need[0] | need[1]) {
                const bool first = need[0];   // which sample the first pass tests in this lane
                const bool clear = obstacles_clear_code(w, first ? code[0] : code[1], w.obs4, first ? xn[0] : xn[1], first ? yn[0] : yn[1],
                                                        first ? tn[0] : tn[1], hl, hw);
                oks[0] &= clear | !first;
                oks[1] &= clear | first;
                if (need[0] & need[1]) oks[1] &= obstacles_clear_code(w, code[1], w.obs4, xn[1], yn[1], tn[1], hl, hw);
            }
        }
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            bool ok = oks[e];
            if (HAS_CONS) {
                if (ok && (cmask[e] != 0u || a.n_cons > 32))
                    ok = constraints_clear(w, a.cons, a.n_cons, a.cons_states, a.cons_total, t0[e] + j + 1, xn[e], yn[e], tn[e], hl, hw,
                                           w.rob_rad[a.robot], cmask[e]);
            }
            nv[e] = ok ? j + 1 : nv[e];
            alive[e] = ok & (j + 1 < st[e]);
            if (!STAGED) {
                xl[e] = ok ? xn[e] : xl[e];
                yl[e] = ok ? yn[e] : yl[e];
                tl[e] = ok ? tn[e] : tl[e];
            }
            if (!ok) continue;
            if (STAGED) {
                const int at = at0[e] + j * kStride;
                ref(0, at) = xn[e];
                ref(1, at) = yn[e];
                ref(2, at) = tn[e];
                if (FIVE) {   // the expressions st was derived with (and the dense write-out's)
                    ref(3, at) = fmaf(ua[e], t1, v0[e]);
                    ref(4, at) = fmaf(uw[e], t1, phi0[e]);
                }
            } else if (a.seg) {
                // straight to the sample's own slot: the CTA's stores of a row stay inside one 1 KB window, L2 merges
                // them (and the epilogue's) into full lines; v and phi are written by the epilogue
                float *o = a.seg + (long long)j * a.n + i[e];
                __stcs(o, xn[e]);
                __stcs(o + plane, yn[e]);
                __stcs(o + 4 * plane, tn[e]);
            }
        }
        p.TH = pk(tn[0], tn[1]);
    }

#pragma unroll
    for (int e = 0; e < 2; ++e) {
        if (slot[e] < 0) continue;  // past the end of the batch
        if (MODE == kPacked) {
            // rows the sample owns but did not reach (it left the workspace / hit an obstacle early) leave as zeros
            for (int j = nv[e]; j < st[e]; ++j) {
                ref(0, at0[e] + j) = 0.0f;
                ref(1, at0[e] + j) = 0.0f;
                ref(2, at0[e] + j) = 0.0f;
                if (FIVE) {
                    ref(3, at0[e] + j) = 0.0f;
                    ref(4, at0[e] + j) = 0.0f;
                }
            }
        }
        if (STAGED) {
            // final states and valid counts leave with the write-out (coalesced, from the stage)
            stage->nv[slot[e]] = (unsigned char)nv[e];
            continue;
        }
        s_nv[slot[e]] = (unsigned char)nv[e];
        if (a.valid_steps) a.valid_steps[i[e]] = nv[e];
        if (a.fin) {
            // the last valid state: x, y, theta as tracked, v and phi from their linear form
            const float tfin = (float)nv[e] * w.step;
            a.fin[i[e]] = xl[e];
            a.fin[a.n + i[e]] = yl[e];
            a.fin[2 * a.n + i[e]] = nv[e] > 0 ? fmaf(ua[e], tfin, v0[e]) : v0[e];
            a.fin[3 * a.n + i[e]] = nv[e] > 0 ? fmaf(uw[e], tfin, phi0[e]) : phi0[e];
            a.fin[4 * a.n + i[e]] = tl[e];
        }
    }
}

// Final states and valid counts of the CTA's samples, from the stage (staged variants): thread t owns the four consecutive
// slots 4 t .. 4 t + 3, so every plane leaves as 16-byte stores.  (x, y, theta) = the sample's last valid row (`row`
// gives its stage position), v and phi from their linear form; a sample that never moved returns its parent.
template <int SAMPLES, class Stage, class Row, class Ref>
__device__ __forceinline__ void write_final_states(const DevWorld &w, const PropArgs &a, const Stage *stage, long long b0, int tid, Row row,
                                                   Ref ref)
{
    if (!a.fin && !a.valid_steps) return;
    for (int slot = 4 * tid; slot < SAMPLES; slot += 4 * kPropThreads) {
        const long long i0 = b0 + slot;
        if (i0 >= a.n) break;
        float f[5][4];
        int nvs[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const long long i = min(i0 + e, a.n - 1);
            const int sl = slot + e;
            const int nv = i0 + e < a.n ? (int)stage->nv[sl] : 0;
            nvs[e] = nv;
            if (!a.fin) continue;
            if (nv > 0) {
                const int at = row(sl, nv - 1);
                const float tfin = (float)nv * w.step;
                f[0][e] = ref(0, at);
                f[1][e] = ref(1, at);
                f[2][e] = fmaf(stage->ua[sl], tfin, stage->v0[sl]);
                f[3][e] = fmaf(stage->uw[sl], tfin, stage->p0[sl]);
                f[4][e] = ref(2, at);
            } else {
                const long long pi = parent_of(a, i);
#pragma unroll
                for (int c = 0; c < 5; ++c) f[c][e] = a.parents[c * a.n_parents + pi];
            }
        }
        const bool vec = i0 + 3 < a.n && (a.n & 3) == 0;
        if (a.fin) {
            if (vec && (reinterpret_cast<uintptr_t>(a.fin) & 15) == 0) {
#pragma unroll
                for (int c = 0; c < 5; ++c)
                    *reinterpret_cast<float4 *>(a.fin + c * a.n + i0) = make_float4(f[c][0], f[c][1], f[c][2], f[c][3]);
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    if (i0 + e < a.n) {
#pragma unroll
                        for (int c = 0; c < 5; ++c) a.fin[c * a.n + i0 + e] = f[c][e];
                    }
            }
        }
        if (a.valid_steps) {
            if (i0 + 3 < a.n && (reinterpret_cast<uintptr_t>(a.valid_steps + i0) & 15) == 0) {
                *reinterpret_cast<int4 *>(a.valid_steps + i0) = make_int4(nvs[0], nvs[1], nvs[2], nvs[3]);
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    if (i0 + e < a.n) a.valid_steps[i0 + e] = nvs[e];
            }
        }
    }
}

// Exclusive prefix sums over the CTA's slots in slot order: thread t owns the PER consecutive slots [PER t, PER t + PER).
// Returns the CTA total; `s_warp` is scratch for one value per warp.
template <int PER, class Get, class Put>
__device__ __forceinline__ int block_exclusive_scan(int tid, int *s_warp, Get get, Put put)
{
    int v[PER], sum = 0;
#pragma unroll
    for (int u = 0; u < PER; ++u) {
        v[u] = get(PER * tid + u);
        sum += v[u];
    }
    const int lane = tid & 31, wid = tid >> 5;
    int incl = sum;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const int up = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += up;
    }
    if (lane == 31) s_warp[wid] = incl;
    __syncthreads();
    int before = 0, total = 0;
#pragma unroll
    for (int q = 0; q < kPropThreads / 32; ++q) {
        before += q < wid ? s_warp[q] : 0;
        total += s_warp[q];
    }
    int run = before + incl - sum;
#pragma unroll
    for (int u = 0; u < PER; ++u) {
        put(PER * tid + u, run);
        run += v[u];
    }
    __syncthreads();   // s_warp may be reused, and the results are visible to every thread
    return total;
}

template <bool HAS_OBS, bool HAS_CONS, int MODE, int ROUNDS>
__global__ void __launch_bounds__(kPropThreads, MODE != kDirect ? PropGeom<ROUNDS>::kCtasPerSm : KCBS_PROP_MINB)
k_propagate(const __grid_constant__ DevWorld w, const __grid_constant__ PropArgs a_in)
{
    constexpr bool STAGED = MODE != kDirect;
    // Planner rounds (direct path only): what the launch cannot know -- it is the same launch for every round of every
    // plan -- comes from the plan block on the device: the robot, the constraint counts, the parity of the round.
    PropArgs a_plan;
    if (MODE == kDirect) {
        a_plan = a_in;
        if (a_in.plan) {
            const int4 h = __ldcg(reinterpret_cast<const int4 *>(a_in.plan));   // round, robot[0], robot[1], n_cons
            a_plan.robot = h.y;
            a_plan.controls = a_in.controls + (h.x & 1) * a_in.parity_ctrl;
            if (a_in.steps) a_plan.steps = a_in.steps + (h.x & 1) * a_in.parity_steps;
            if (a_in.parent_idx) a_plan.parent_idx = a_in.parent_idx + (h.x & 1) * a_in.parity_parent;
            if (HAS_CONS) {
                a_plan.n_cons = h.w;
                a_plan.cons_total = __ldcg(&a_in.plan->cons_total);
            }
        }
    }
    const PropArgs &a = MODE == kDirect ? a_plan : a_in;
    constexpr int kSamples = PropGeom<ROUNDS>::kSamples, kPer = PropGeom<ROUNDS>::kPer;
    typedef PropStage<kSamples> Stage;
    __shared__ int s_hist[KCBS_MAX_STEPS + 1], s_first[KCBS_MAX_STEPS + 1];
    __shared__ unsigned short s_perm[kSamples];  // rank -> slot
    __shared__ unsigned char s_steps[kSamples];  // slot -> effective duration
    __shared__ unsigned char s_nv[MODE == kDirect ? kSamples : 4];     // slot -> valid steps (direct path)
    __shared__ PackTables<MODE == kPacked ? kSamples : 32> s_pk;    // (a stub unless kPacked)
    __shared__ int s_warp[kPropThreads / 32];
    extern __shared__ __align__(16) unsigned char s_dyn[];  // the stage (staged variants only)
    Stage *stage = STAGED ? reinterpret_cast<Stage *>(s_dyn) : nullptr;

    const int tid = threadIdx.x;
    if (tid <= KCBS_MAX_STEPS) s_hist[tid] = 0;
    const long long b0 = (long long)blockIdx.x * kSamples;
    StageIndex<MODE, kSamples> six;
    six.cap_off = s_pk.cap_off;
    six.steps = s_steps;
    six.flat = STAGED ? &stage->xyt[0][0] : nullptr;
    six.pstride = StageIndex<MODE, kSamples>::kPositions;
    six.five = false;
    unsigned long long alloc = 0;   // (kPacked, thread 0) the row allocator's answer, see below

    // ---- counting sort of the CTA's samples by effective duration (longest first) ----------------
    int st4[kPer], within[kPer];
    {
        // all global loads of the thread's samples are issued before anything depends on them (indices clamped
        // instead of branched on): one or two memory round trips per CTA instead of a dozen serialised ones
        long long ii[kPer], pi[kPer];
        int st_req[kPer];
        float v0[kPer], phi0[kPer], ua[kPer], uw[kPer], x0[kPer], y0[kPer], th0[kPer];
        auto load_samples = [&]() {
#pragma unroll
            for (int u = 0; u < kPer; ++u) {
                ii[u] = min(b0 + tid + u * kPropThreads, a.n - 1);
                pi[u] = parent_of(a, ii[u]);
                st_req[u] = a.steps ? a.steps[ii[u]] : a.max_steps;
                ua[u] = a.controls[ii[u]];
                uw[u] = a.controls[a.n + ii[u]];
            }
#pragma unroll
            for (int u = 0; u < kPer; ++u) {
                v0[u] = a.parents[2 * a.n_parents + pi[u]];
                phi0[u] = a.parents[3 * a.n_parents + pi[u]];
                if (STAGED) {
                    x0[u] = a.parents[0 * a.n_parents + pi[u]];
                    y0[u] = a.parents[1 * a.n_parents + pi[u]];
                    th0[u] = a.parents[4 * a.n_parents + pi[u]];
                }
            }
        };
        load_samples();
        __syncthreads();   // the histogram is clear
#pragma unroll
        for (int u = 0; u < kPer; ++u) {
            const int slot = tid + u * kPropThreads;
            int st = 0;
            if (b0 + slot < a.n) {
                st = min(max(st_req[u], 0), a.max_steps);
                // v and phi are linear in time, so the first step at which they leave their bounds is known
                // before integrating anything: that step would fail isValid, so the segment ends just before
                // it (same result as propagating it and testing).  Sorting by this effective duration keeps
                // warps uniform even when many samples leave the v / phi bounds early.
                // fmaf(u, t, v0) is monotonic in t, so the valid steps form a prefix.  The crossing estimated from
                // the linear form is right to within one step, so five independent evaluations of the exact test
                // (no data-dependent branch: the four samples of a thread interleave) settle it; anything they do
                // not settle (a degenerate estimate) falls back to counting the prefix.
                auto inside = [&](int j) {
                    const float t1 = (float)(j + 1) * w.step;
                    const float vn = fmaf(ua[u], t1, v0[u]), pn = fmaf(uw[u], t1, phi0[u]);
                    return (vn >= w.lo_t[2]) & (vn <= w.hi_t[2]) & (pn >= w.lo_t[3]) & (pn <= w.hi_t[3]);
                };
                const float cv = __fdividef((ua[u] > 0.0f ? w.hi_t[2] : w.lo_t[2]) - v0[u], ua[u] * w.step);
                const float cp = __fdividef((uw[u] > 0.0f ? w.hi_t[3] : w.lo_t[3]) - phi0[u], uw[u] * w.step);
                const float c = fminf(ua[u] != 0.0f ? cv : 1.0e9f, uw[u] != 0.0f ? cp : 1.0e9f);
                const int k = min(max((int)c, 1), st);
                const bool e0 = inside(0), em2 = inside(k - 2), em1 = inside(k - 1), ek = inside(k),
                           ek1 = inside(k + 1);
                int ok_steps = k + 1;
                bool unsettled = !(k + 1 == st || !ek1);
                if (k == st || !ek) { ok_steps = k; unsettled = false; }
                if (!em1) { ok_steps = k - 1; unsettled = !em2; }
                if (st == 0 || !e0) { ok_steps = 0; unsettled = false; }
                if (unsettled) {
                    ok_steps = 0;
                    while (ok_steps < st && inside(ok_steps)) ++ok_steps;
                }
                st = ok_steps;
                if (STAGED) {
                    // inputs of the sample for propagate_two and the write-out
                    stage->v0[slot] = v0[u];
                    stage->p0[slot] = phi0[u];
                    stage->ua[slot] = ua[u];
                    stage->uw[slot] = uw[u];
                }
            }
            st4[u] = st;
            s_steps[slot] = (unsigned char)st;
            if (STAGED) stage->nv[slot] = 0;
            else s_nv[slot] = 0;
            within[u] = atomicAdd(&s_hist[st], 1);
        }
        if (MODE == kPacked) {
            // stage positions = rows: every slot gets a run of st positions in slot order
            __syncthreads();
            const int total = block_exclusive_scan<kPer>(
                tid, s_warp, [&](int slot) { return (int)s_steps[slot]; },
                [&](int slot, int pre) { s_pk.cap_off[slot] = (unsigned short)pre; });
            if (tid == 0) {
                s_pk.cap_total = total;
                s_pk.cap_off[kSamples] = (unsigned short)total;
                // The CTA's rows are allocated first come, first served: one atomic on (CTAs served << 40 | rows handed
                // out); nobody waits for anybody.  The answer stays in a register until the write-out, so thread 0 does
                // not wait for it here.
                alloc = atomicAdd(reinterpret_cast<unsigned long long *>(a.scan_ctl), (1ull << 40) | (unsigned long long)((total + 3) & ~3));
            }
            // rows of the CTA fit into 3/5 of the stage: five planes, v and phi are staged too
            six.five = total <= StageIndex<MODE, kSamples>::kFivePositions;
            if (six.five) six.pstride = StageIndex<MODE, kSamples>::kFivePositions;
        }
        if (STAGED) {
            // x0, y0, theta0 borrow the sample's own last stage position, which the sample overwrites only when it
            // completes its final step
#pragma unroll
            for (int u = 0; u < kPer; ++u) {
                const int slot = tid + u * kPropThreads;
                if (b0 + slot < a.n && (MODE != kPacked || st4[u] > 0)) {
                    const int at = six.stash(slot);
                    six.ref(0, at) = x0[u];
                    six.ref(1, at) = y0[u];
                    six.ref(2, at) = th0[u];
                }
            }
        }
    }
    __syncthreads();
    if (tid < 32) {  // s_first[d] = rank of the first sample of duration d (longer durations come first): suffix sums
        static_assert(KCBS_MAX_STEPS == 32, "one warp scans bins 1..32");
        int sum = s_hist[tid + 1];
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int up = __shfl_down_sync(0xffffffffu, sum, off);
            if (tid + off < 32) sum += up;
        }
        s_first[tid] = sum;  // hist[tid + 1] + ... + hist[32]
        if (tid == 0) s_first[KCBS_MAX_STEPS] = 0;
    }
    __syncthreads();
#pragma unroll
    for (int u = 0; u < kPer; ++u) s_perm[s_first[st4[u]] + within[u]] = (unsigned short)(tid + u * kPropThreads);
    __syncthreads();

    // ---- packed pairs of neighbouring ranks: (2t, 2t+1) (long) and, with two rounds, (S-2-2t, S-1-2t) (short) ------
    const float tan_shared = a.n_parents == 1 ? tan_phi(a.parents[3]) : 0.0f;
#pragma unroll 1
    for (int round = 0; round < ROUNDS; ++round) {
        const int r0 = round == 0 ? 2 * tid : kSamples - 2 - 2 * tid;
        int slot[2], st[2];
        long long i[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            slot[e] = s_perm[r0 + e];
            i[e] = b0 + slot[e];
            st[e] = (int)s_steps[slot[e]];
            if (i[e] >= a.n) {  // past the end of the batch: integrate nothing, write nothing
                i[e] = a.n - 1;
                st[e] = 0;
                slot[e] = -1;
            }
        }
        if (slot[0] >= 0 || slot[1] >= 0) {
            if (MODE == kPacked && six.five)
                propagate_two<HAS_OBS, HAS_CONS, MODE, kSamples, MODE == kPacked>(w, a, i, st, stage, six, s_nv, slot, tan_shared);
            else
                propagate_two<HAS_OBS, HAS_CONS, MODE, kSamples, false>(w, a, i, st, stage, six, s_nv, slot, tan_shared);
        }
    }

    if (MODE == kDirect && a.seg) {
        // ---- coalesced epilogue of the direct path: every thread owns the slots it sorted; it writes v and phi of the
        // valid rows (exact linear functions of time, the very expressions the integrator tested) and zero-fills all
        // five planes of the rows past the last valid state.  Disjoint from the x / y / theta stores above.
        __syncthreads();
        const long long plane = (long long)a.max_steps * a.n;
#pragma unroll
        for (int u = 0; u < kPer; ++u) {
            const int slot = tid + u * kPropThreads;
            const long long i = b0 + slot;
            if (i >= a.n) continue;
            const int nv = s_nv[slot];
            const long long pi = parent_of(a, i);
            const float v0 = a.parents[2 * a.n_parents + pi], p0 = a.parents[3 * a.n_parents + pi];
            const float ua = a.controls[i], uw = a.controls[a.n + i];
            for (int j = 0; j < a.max_steps; ++j) {
                float *o = a.seg + (long long)j * a.n + i;
                const float t1 = (float)(j + 1) * w.step;
                if (j < nv) {
                    __stcs(o + 2 * plane, fmaf(ua, t1, v0));
                    __stcs(o + 3 * plane, fmaf(uw, t1, p0));
                } else {
#pragma unroll
                    for (int c = 0; c < 5; ++c) __stcs(o + c * plane, 0.0f);
                }
            }
        }
    }

    if (MODE == kDense) {
        // ---- coalesced write-out: every row of every plane leaves as full 128 B lines; rows past the last valid
        // state are zero, so no sector of the segment buffer is ever partially written --------------------------
        __syncthreads();
        const long long plane = (long long)a.max_steps * a.n;
        if ((a.n & 3) == 0 && (reinterpret_cast<uintptr_t>(a.seg) & 15) == 0) {
            // 16 B stores: the thread owns four consecutive slots
            for (int slot = 4 * tid; slot < kSamples; slot += 4 * kPropThreads) {
                const long long i = b0 + slot;
                if (i >= a.n) break;
                const uchar4 nv4 = *reinterpret_cast<const uchar4 *>(&stage->nv[slot]);
                const int nv[4] = {nv4.x, nv4.y, nv4.z, nv4.w};
                const float4 v0 = *reinterpret_cast<const float4 *>(&stage->v0[slot]);
                const float4 ua = *reinterpret_cast<const float4 *>(&stage->ua[slot]);
                const float4 p0 = *reinterpret_cast<const float4 *>(&stage->p0[slot]);
                const float4 uw = *reinterpret_cast<const float4 *>(&stage->uw[slot]);
                float *o = a.seg + i;
                // (every lane runs all rows: a loop bound that differs between the lanes would split the row stores of
                // the warp into partial sectors)
                for (int j = 0; j < a.max_steps; ++j, o += a.n) {
                    const float t1 = (float)(j + 1) * w.step;
                    const bool l0 = j < nv[0], l1 = j < nv[1], l2 = j < nv[2], l3 = j < nv[3];
                    const float4 X = *reinterpret_cast<const float4 *>(&stage->xyt[0][j * kSamples + slot]);
                    const float4 Y = *reinterpret_cast<const float4 *>(&stage->xyt[1][j * kSamples + slot]);
                    const float4 T = *reinterpret_cast<const float4 *>(&stage->xyt[2][j * kSamples + slot]);
                    __stcs(reinterpret_cast<float4 *>(o),
                           make_float4(l0 ? X.x : 0.0f, l1 ? X.y : 0.0f, l2 ? X.z : 0.0f, l3 ? X.w : 0.0f));
                    __stcs(reinterpret_cast<float4 *>(o + plane),
                           make_float4(l0 ? Y.x : 0.0f, l1 ? Y.y : 0.0f, l2 ? Y.z : 0.0f, l3 ? Y.w : 0.0f));
                    // same expressions as the integrator's
                    __stcs(reinterpret_cast<float4 *>(o + 2 * plane),
                           make_float4(l0 ? fmaf(ua.x, t1, v0.x) : 0.0f, l1 ? fmaf(ua.y, t1, v0.y) : 0.0f,
                                       l2 ? fmaf(ua.z, t1, v0.z) : 0.0f, l3 ? fmaf(ua.w, t1, v0.w) : 0.0f));
                    __stcs(reinterpret_cast<float4 *>(o + 3 * plane),
                           make_float4(l0 ? fmaf(uw.x, t1, p0.x) : 0.0f, l1 ? fmaf(uw.y, t1, p0.y) : 0.0f,
                                       l2 ? fmaf(uw.z, t1, p0.z) : 0.0f, l3 ? fmaf(uw.w, t1, p0.w) : 0.0f));
                    __stcs(reinterpret_cast<float4 *>(o + 4 * plane),
                           make_float4(l0 ? T.x : 0.0f, l1 ? T.y : 0.0f, l2 ? T.z : 0.0f, l3 ? T.w : 0.0f));
                }
            }
        } else {
#pragma unroll
            for (int u = 0; u < kPer; ++u) {
                const int slot = tid + u * kPropThreads;
                const long long i = b0 + slot;
                if (i >= a.n) continue;
                const int nv = stage->nv[slot];
                const float v0 = stage->v0[slot], ua = stage->ua[slot], p0 = stage->p0[slot], uw = stage->uw[slot];
                float *o = a.seg + i;
                for (int j = 0; j < a.max_steps; ++j, o += a.n) {
                    const bool live = j < nv;
                    const float t1 = (float)(j + 1) * w.step;
                    __stcs(o, live ? stage->xyt[0][j * kSamples + slot] : 0.0f);
                    __stcs(o + plane, live ? stage->xyt[1][j * kSamples + slot] : 0.0f);
                    __stcs(o + 2 * plane, live ? fmaf(ua, t1, v0) : 0.0f);
                    __stcs(o + 3 * plane, live ? fmaf(uw, t1, p0) : 0.0f);
                    __stcs(o + 4 * plane, live ? stage->xyt[2][j * kSamples + slot] : 0.0f);
                }
            }
        }
        write_final_states<kSamples>(w, a, stage, b0, tid, [&](int sl, int j) { return j * kSamples + sl; },
                                     [&](int c, int at) { return six.ref(c, at); });
    }

    if (MODE == kPacked) {
        // ---- packed write-out: the CTA's rows leave as ONE contiguous, 16-byte aligned range of every plane, a straight
        // copy of the stage (row order already).  With three stage planes (the CTA's rows did not fit five), v and phi
        // are laid out in the freed stage by the threads that own the samples and leave the same way ------------------
        const int cap_total = s_pk.cap_total, rows4 = (cap_total + 3) & ~3;
        if (tid == 0) {
            // the allocator's answer is needed now; the last CTA served also reports the total and leaves the
            // allocator zeroed for the next launch (every other CTA of this launch has been served)
            const long long first = (long long)(alloc & ((1ull << 40) - 1));
            s_pk.base = first;
            if ((unsigned)(alloc >> 40) == gridDim.x - 1) {
                a.offsets[a.n] = (int)(first + rows4);
                *reinterpret_cast<unsigned long long *>(a.scan_ctl) = 0ull;
            }
        }
        __syncthreads();
        const long long base = s_pk.base;
        float *out = a.seg;
        const long long cap = a.packed_cap;
        const bool vec = (cap & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 && base + rows4 <= cap;
        // copies stage planes s[0..NP) to the output planes o[0..NP); positions past cap_total (the round-up) are zero
        auto copy_planes = [&](auto np, const int (&sp)[5], const int (&op)[5]) {
            constexpr int NP = decltype(np)::value;
            for (int p4 = 4 * tid; p4 < rows4; p4 += 4 * kPropThreads) {
                float4 q[NP];
#pragma unroll
                for (int c = 0; c < NP; ++c) q[c] = *reinterpret_cast<const float4 *>(&six.ref(sp[c], p4));
                if (p4 + 4 > cap_total) {
#pragma unroll
                    for (int c = 0; c < NP; ++c) {
                        if (p4 + 1 >= cap_total) q[c].y = 0.0f;
                        if (p4 + 2 >= cap_total) q[c].z = 0.0f;
                        q[c].w = 0.0f;
                    }
                }
                const long long o = base + p4;
#pragma unroll
                for (int c = 0; c < NP; ++c) {
                    float *dst = out + op[c] * cap + o;
                    if (vec) {
                        __stcs(reinterpret_cast<float4 *>(dst), q[c]);
                    } else {
                        const float e4[4] = {q[c].x, q[c].y, q[c].z, q[c].w};
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            if (o + e < cap) __stcs(dst + e, e4[e]);
                    }
                }
            }
        };
        if (six.five) {
            copy_planes(std::integral_constant<int, 5>(), {0, 1, 2, 3, 4}, {0, 1, 4, 2, 3});
        } else {
            copy_planes(std::integral_constant<int, 3>(), {0, 1, 2, 0, 0}, {0, 1, 4, 0, 0});
        }
        // row offsets (16-byte stores: the thread owns four consecutive slots), final states, valid counts
        for (int slot = 4 * tid; slot < kSamples; slot += 4 * kPropThreads) {
            const long long i0 = b0 + slot;
            if (i0 >= a.n) break;
            const int o4[4] = {(int)(base + s_pk.cap_off[slot]), (int)(base + s_pk.cap_off[slot + 1]),
                               (int)(base + s_pk.cap_off[slot + 2]), (int)(base + s_pk.cap_off[slot + 3])};
            if (i0 + 3 < a.n && (reinterpret_cast<uintptr_t>(a.offsets + i0) & 15) == 0) {
                *reinterpret_cast<int4 *>(a.offsets + i0) = make_int4(o4[0], o4[1], o4[2], o4[3]);
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    if (i0 + e < a.n) a.offsets[i0 + e] = o4[e];
            }
        }
        write_final_states<kSamples>(w, a, stage, b0, tid, [&](int sl, int j) { return (int)s_pk.cap_off[sl] + j; },
                                     [&](int c, int at) { return six.ref(c, at); });
        if (!six.five) {
            __syncthreads();   // x, y, theta have left the stage
            for (int slot = 4 * tid; slot < kSamples; slot += 4 * kPropThreads) {
                if (b0 + slot >= a.n) break;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int sl = slot + e;
                    const int at = s_pk.cap_off[sl], st = (int)s_pk.cap_off[sl + 1] - at, nv = stage->nv[sl];
                    const float v0 = stage->v0[sl], ua = stage->ua[sl], p0 = stage->p0[sl], uw = stage->uw[sl];
                    for (int j = 0; j < st; ++j) {
                        const float t1 = (float)(j + 1) * w.step;
                        six.ref(0, at + j) = j < nv ? fmaf(ua, t1, v0) : 0.0f;   // the integrator's expressions
                        six.ref(1, at + j) = j < nv ? fmaf(uw, t1, p0) : 0.0f;
                    }
                }
            }
            __syncthreads();
            copy_planes(std::integral_constant<int, 2>(), {0, 1, 0, 0, 0}, {2, 3, 0, 0, 0});
        }
    }
}

template <int MODE, int ROUNDS, class F>
static cudaError_t for_each_variant(F f)
{
    cudaError_t e;
    if ((e = f(k_propagate<false, false, MODE, ROUNDS>)) != cudaSuccess) return e;
    if ((e = f(k_propagate<true, false, MODE, ROUNDS>)) != cudaSuccess) return e;
    if (MODE == kPacked) return cudaSuccess;   // the planner (the only user of constraints) keeps no segments
    if ((e = f(k_propagate<false, true, MODE == kPacked ? kDense : MODE, ROUNDS>)) != cudaSuccess) return e;
    return f(k_propagate<true, true, MODE == kPacked ? kDense : MODE, ROUNDS>);
}

cudaError_t init_propagate_kernels()
{
    cudaError_t e;
    auto set1 = [](auto *k) { return cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PropStage<PropGeom<1>::kSamples>)); };
    auto set2 = [](auto *k) { return cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PropStage<PropGeom<2>::kSamples>)); };
    if ((e = for_each_variant<kDense, 1>(set1)) != cudaSuccess) return e;
    if ((e = for_each_variant<kPacked, 1>(set1)) != cudaSuccess) return e;
    if ((e = for_each_variant<kDense, 2>(set2)) != cudaSuccess) return e;
    if ((e = for_each_variant<kPacked, 2>(set2)) != cudaSuccess) return e;
    // the direct variants are launched from inside stream captures (planner.cu): load them now, outside any capture
    auto load = [](auto *k) {
        cudaFuncAttributes at;
        return cudaFuncGetAttributes(&at, k);
    };
    if ((e = for_each_variant<kDirect, 1>(load)) != cudaSuccess) return e;
    return for_each_variant<kDirect, 2>(load);
}

template <int MODE, int ROUNDS>
static cudaError_t launch_mode(const DevWorld &w, const PropArgs &a, cudaStream_t stream)
{
    constexpr int kSamples = PropGeom<ROUNDS>::kSamples;
    const unsigned grid = (unsigned)((a.n + kSamples - 1) / kSamples);
    const size_t smem = MODE != kDirect ? sizeof(PropStage<kSamples>) : 0;
    if (a.n_cons > 0) {
        constexpr int M = MODE == kPacked ? kDense : MODE;   // (never launched: packed calls carry no constraints)
        if (MODE == kPacked) return cudaErrorInvalidValue;
        if (w.n_obs > 0)
            k_propagate<true, true, M, ROUNDS><<<grid, kPropThreads, smem, stream>>>(w, a);
        else
            k_propagate<false, true, M, ROUNDS><<<grid, kPropThreads, smem, stream>>>(w, a);
    } else if (w.n_obs > 0) {
        k_propagate<true, false, MODE, ROUNDS><<<grid, kPropThreads, smem, stream>>>(w, a);
    } else {
        k_propagate<false, false, MODE, ROUNDS><<<grid, kPropThreads, smem, stream>>>(w, a);
    }
    return cudaGetLastError();
}

// Samples up to which one launch runs with one packed pair per thread: the grid of 256-sample CTAs is then a single wave
// (148 SMs x 3 CTAs), i.e. the launch cannot fill the GPU with two pairs per thread anyway.
constexpr long long kLatencySamples = 148ll * 3 * PropGeom<1>::kSamples;

int propagate_grid(const PropArgs &a)
{
    const int rounds = a.rounds == 1 || a.rounds == 2 ? a.rounds : (a.n <= kLatencySamples ? 1 : 2);
    const int samples = rounds == 1 ? PropGeom<1>::kSamples : PropGeom<2>::kSamples;
    return (int)((a.n + samples - 1) / samples);
}

cudaError_t launch_propagate(const DevWorld &w, const PropArgs &a, cudaStream_t stream)
{
    if (a.n <= 0) return cudaSuccess;
    const int rounds = a.rounds == 1 || a.rounds == 2 ? a.rounds : (a.n <= kLatencySamples ? 1 : 2);
    // whole segments of up to kStageSteps steps leave through the shared-memory stage (coalesced);
    // longer horizons and final-state-only calls write directly
    if (a.offsets != nullptr) {
        if (a.max_steps > kStageSteps || a.seg == nullptr || a.scan == nullptr) return cudaErrorInvalidValue;
        return rounds == 1 ? launch_mode<kPacked, 1>(w, a, stream) : launch_mode<kPacked, 2>(w, a, stream);
    }
    // (planner launches -- a.plan -- take the direct path: it is the one that reads the plan block)
    if (a.seg != nullptr && a.max_steps <= kStageSteps && a.plan == nullptr)
        return rounds == 1 ? launch_mode<kDense, 1>(w, a, stream) : launch_mode<kDense, 2>(w, a, stream);
    return rounds == 1 ? launch_mode<kDirect, 1>(w, a, stream) : launch_mode<kDirect, 2>(w, a, stream);
}

}  // namespace kcbs
