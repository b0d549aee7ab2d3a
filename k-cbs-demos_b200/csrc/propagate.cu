// propagate.cu -- K1: fused multi-step RK4 expansion kernel (one thread per control sample).
//
// Replaces, per sample, oc::SpaceInformation::propagateWhileValid driving
//   SecondOrderCarODE                      StatePropagatorDatabase.h:44-64
//   odeint runge_kutta4, dt = 0.01         (ODEBasicSolver, call site demo_*.cpp:111-112)
//   SecondOrderCarODEPostIntegration       StatePropagatorDatabase.h:67-74 (theta wrap per step)
//   homogeneous2ndOrderCarSystemSVC::isValid  StateValidityCheckerDatabase.h:54-75
//
// The kernel is classical RK4 with the reference's stages, evaluated with what the ODE's structure
// gives for free:
//   * vdot = a and phidot = w are constant: the v / phi stage values are exact linear functions of
//     time, stages 2 and 3 share v, phi, tan(phi) and k_theta, and stage 4 of a sub-step is stage 1
//     of the next one.
//   * phi advances by the constant delta = (h/2) w per half sub-step, so tan(phi) follows the addition
//     formula tan(phi + delta) = (T + t) / (1 - T t) (FADD + FFMA + MUFU.RCP + FMUL); re-anchored by an
//     accurate sincosf every propagation step.
//   * the four stage headings of a sub-step differ by at most h*|k_theta| <= 0.035 rad, so their
//     (cos, sin) follow from the previous stage by a 2nd-order rotation (FFMA only, truncation
//     < 1e-6 which enters the position with weight h/6); re-anchored by MUFU sincos every step.
//     This moves the kernel off the MUFU pipe (10 -> 2 MUFU per sub-step) onto the FMA pipe.
//   * k_theta is carried pre-scaled by h/2 and v pre-scaled by h/6, both as exact linear sequences, so a
//     sub-step is ~45 FMA-pipe instructions + 2 MUFU.
//   * positions and the heading are accumulated as small offsets from the parent / the step start,
//     so FP32 rounding at |x| ~ 32 happens once per written state, not once per sub-step.
//
// Work distribution: requested durations are U{1..10} (SimpleDirectedControlSampler), which would
// idle 45 % of the lanes of a warp; and because v, phi are linear in time, the step at which a sample
// leaves the v / phi bounds is known up front, so its effective duration is too (the doomed step is never
// integrated).  Each CTA (128 threads) therefore counting-sorts its 256 samples by
// duration in shared memory and every thread runs two of them: rank t (long) and rank 255 - t (short).
// Lanes of a warp then hold (almost) equal durations in both rounds and every thread of the CTA has the
// same total work, so neither lanes nor warps idle.  Segments are staged in shared memory at their original
// slots (x, y, theta per step; v and phi are recomputed from their exact linear form) and leave in one
// coalesced pass that also zero-fills the rows past the last valid state: every 32 B sector is written
// completely (no partial-sector traffic at L2, no read-modify-write at HBM).
#include <cstddef>

#include "car_dynamics.cuh"
#include "kernels.h"

namespace kcbs {

// Dynamic-obstacle constraints of the K-CBS low level (the fork tests `next` against the constraining robot's
// state at that time through areStatesValid, StateValidityCheckerDatabase.h:78-125).
__device__ __forceinline__ bool constraints_clear(const DevWorld &w, const PropArgs &a, int k, float x, float y,
                                                  float th, float hl, float hw, float rad)
{
    bool ok = true, have = false;
    float c = 1.0f, s = 0.0f;
    for (int q = 0; q < a.n_cons; ++q) {
        const ConsDesc d = a.cons[q];
        const int rel = k - d.k0;
        if (rel < 0 || rel >= d.len) continue;
        const float *st = a.cons_states + d.offset + rel;
        const float ox = st[0], oy = st[a.cons_total], dx = ox - x, dy = oy - y;
        const float rs = rad + w.rob_rad[d.other];
        if (fmaf(dy, dy, dx * dx) > rs * rs * 1.000001f) continue;
        if (!have) {
            sincosf(th, &s, &c);
            have = true;
        }
        float oc, os;
        sincosf(st[2 * a.cons_total], &os, &oc);
        ok &= sat_gap_robots(x, y, c, s, hl, hw, ox, oy, oc, os, w.rob_hl[d.other], w.rob_hw[d.other]) > 0.0f;
    }
    return ok;
}

constexpr int kPropThreads = 128;  // threads per CTA
constexpr int kPropSamples = 256;  // samples per CTA (two per thread)
constexpr int kStageSteps = 10;    // segments up to this many steps are staged in shared memory

// Shared-memory staging of one CTA's segments: x, y, theta of every (step, slot), plus the four numbers from which
// v and phi of any step follow exactly (they are linear in time).  35 KB -> 6 CTAs (24 warps) per SM.
struct PropStage {
    float xyt[3][kStageSteps][kPropSamples];
    float v0[kPropSamples], ua[kPropSamples], p0[kPropSamples], uw[kPropSamples];
    unsigned char nv[kPropSamples];
};

template <bool STAGED>
struct StageStorage {
    using type = PropStage;
    __device__ static PropStage *get(type &t) { return &t; }
};
template <>
struct StageStorage<false> {  // direct-store variants carry no stage
    using type = int;
    __device__ static PropStage *get(type &) { return nullptr; }
};

// propagateWhileValid for one sample; writes the segment, the final state and the valid-step count.
template <bool HAS_OBS, bool HAS_CONS, bool STAGED>
__device__ __forceinline__ void propagate_sample(const DevWorld &w, const PropArgs &a, long long i, int st, PropStage *stage,
                                                 int slot)
{
    const long long pi = a.parent_idx ? (long long)a.parent_idx[i] : (a.n_parents == 1 ? 0 : i);
    const int t0 = (HAS_CONS && a.parent_time) ? a.parent_time[pi] : 0;
    const float x0 = a.parents[0 * a.n_parents + pi];
    const float y0 = a.parents[1 * a.n_parents + pi];
    const float v0 = a.parents[2 * a.n_parents + pi];
    const float phi0 = a.parents[3 * a.n_parents + pi];
    const float th0 = a.parents[4 * a.n_parents + pi];
    const float ua = a.controls[i];
    const float uw = a.controls[a.n + i];

    const float hl = w.rob_hl[a.robot], hw = w.rob_hw[a.robot];
    const CarConsts k = car_consts(w, ua, uw);

    float offx = 0.0f, offy = 0.0f;                         // position offset from the parent
    float xs = x0, ys = y0, vs = v0, ps = phi0, ts = th0;   // last valid state
    int nv = 0;
    const long long plane = (long long)a.max_steps * a.n;

    for (int j = 0; j < st; ++j) {
        const float toff = car_step(k, w.n_sub, vs, ps, ts, offx, offy);

        const float t1 = (float)(j + 1) * w.step;
        const float xn = x0 + offx, yn = y0 + offy;
        const float vn = fmaf(ua, t1, v0), pn = fmaf(uw, t1, phi0);
        const float tn = wrap_angle(ts + toff);

        bool ok = in_bounds(w, xn, yn, vn, pn, tn);
        if (HAS_OBS) {
            if (ok) ok = obstacles_clear_fine(w, xn, yn, tn, hl, hw);
        }
        if (HAS_CONS) {
            if (ok) ok = constraints_clear(w, a, t0 + j + 1, xn, yn, tn, hl, hw, w.rob_rad[a.robot]);
        }
        if (!ok) break;

        xs = xn; ys = yn; vs = vn; ps = pn; ts = tn;
        nv = j + 1;
        if (STAGED) {
            stage->xyt[0][j][slot] = xn;
            stage->xyt[1][j][slot] = yn;
            stage->xyt[2][j][slot] = tn;
        } else if (a.seg) {
            float *o = a.seg + (long long)j * a.n + i;
            __stcs(o, xn);
            __stcs(o + plane, yn);
            __stcs(o + 2 * plane, vn);
            __stcs(o + 3 * plane, pn);
            __stcs(o + 4 * plane, tn);
        }
    }

    if (STAGED) {
        stage->v0[slot] = v0;
        stage->ua[slot] = ua;
        stage->p0[slot] = phi0;
        stage->uw[slot] = uw;
        stage->nv[slot] = (unsigned char)nv;
    } else if (a.seg) {
        // rows past the last valid state are zero-filled (as the staged path does)
        for (int j = nv; j < a.max_steps; ++j) {
            float *o = a.seg + (long long)j * a.n + i;
#pragma unroll
            for (int c = 0; c < 5; ++c) __stcs(o + c * plane, 0.0f);
        }
    }
    if (a.valid_steps) a.valid_steps[i] = nv;
    if (a.fin) {
        a.fin[i] = xs;
        a.fin[a.n + i] = ys;
        a.fin[2 * a.n + i] = vs;
        a.fin[3 * a.n + i] = ps;
        a.fin[4 * a.n + i] = ts;
    }
}

template <bool HAS_OBS, bool HAS_CONS, bool STAGED>
__global__ void __launch_bounds__(kPropThreads)
k_propagate(const __grid_constant__ DevWorld w, const __grid_constant__ PropArgs a)
{
    __shared__ int s_hist[KCBS_MAX_STEPS + 1];
    __shared__ unsigned char s_perm[kPropSamples];   // rank -> slot
    __shared__ unsigned char s_steps[kPropSamples];  // slot -> effective duration
    __shared__ typename StageStorage<STAGED>::type s_stage_mem;
    PropStage *stage = StageStorage<STAGED>::get(s_stage_mem);

    const int tid = threadIdx.x;
    const long long b0 = (long long)blockIdx.x * kPropSamples;

    // ---- counting sort of the CTA's samples by effective duration (longest first) ----------------
    if (tid <= KCBS_MAX_STEPS) s_hist[tid] = 0;
    __syncthreads();
    int st2[2], within[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        const int slot = tid + u * kPropThreads;
        const long long i = b0 + slot;
        int st = 0;
        if (i < a.n) {
            st = a.steps ? a.steps[i] : a.max_steps;
            st = min(max(st, 0), a.max_steps);
            // v and phi are linear in time, so the first step at which they leave their bounds is known
            // before integrating anything: that step would fail isValid, so the segment ends just before
            // it (same result as propagating it and testing).  Sorting by this effective duration keeps
            // warps uniform even when many samples leave the v / phi bounds early.
            const long long pi = a.parent_idx ? (long long)a.parent_idx[i] : (a.n_parents == 1 ? 0 : i);
            const float v0 = a.parents[2 * a.n_parents + pi], phi0 = a.parents[3 * a.n_parents + pi];
            const float ua = a.controls[i], uw = a.controls[a.n + i];
            int ok_steps = 0;
            for (int j = 0; j < st; ++j) {
                const float t1 = (float)(j + 1) * w.step;
                const float vn = fmaf(ua, t1, v0), pn = fmaf(uw, t1, phi0);
                if (!((vn >= w.lo_t[2]) & (vn <= w.hi_t[2]) & (pn >= w.lo_t[3]) & (pn <= w.hi_t[3]))) break;
                ok_steps = j + 1;
            }
            st = ok_steps;
        }
        st2[u] = st;
        s_steps[slot] = (unsigned char)st;
        if (STAGED) stage->nv[slot] = 0;
        within[u] = atomicAdd(&s_hist[st], 1);
    }
    __syncthreads();
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        int before = 0;  // samples with a longer duration come first
        for (int s = KCBS_MAX_STEPS; s > st2[u]; --s) before += s_hist[s];
        s_perm[before + within[u]] = (unsigned char)(tid + u * kPropThreads);
    }
    __syncthreads();

    // ---- two samples per thread: rank tid (long) and rank 255 - tid (short) ------------------------
#pragma unroll 1
    for (int round = 0; round < 2; ++round) {
        const int rank = round == 0 ? tid : kPropSamples - 1 - tid;
        const int slot = s_perm[rank];
        const long long i = b0 + slot;
        if (i < a.n) propagate_sample<HAS_OBS, HAS_CONS, STAGED>(w, a, i, (int)s_steps[slot], stage, slot);
    }

    if (STAGED) {
        // ---- coalesced write-out: every row of every plane leaves as full 128 B lines; rows past the last valid
        // state are zero, so no sector of the segment buffer is ever partially written --------------------------
        __syncthreads();
        const long long plane = (long long)a.max_steps * a.n;
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int slot = tid + u * kPropThreads;
            const long long i = b0 + slot;
            if (i >= a.n) continue;
            const int nv = stage->nv[slot];
            const float v0 = stage->v0[slot], ua = stage->ua[slot], p0 = stage->p0[slot], uw = stage->uw[slot];
            for (int j = 0; j < a.max_steps; ++j) {
                float *o = a.seg + (long long)j * a.n + i;
                const bool live = j < nv;
                const float t1 = (float)(j + 1) * w.step;
                __stcs(o, live ? stage->xyt[0][j][slot] : 0.0f);
                __stcs(o + plane, live ? stage->xyt[1][j][slot] : 0.0f);
                __stcs(o + 2 * plane, live ? fmaf(ua, t1, v0) : 0.0f);   // same expression as the integrator's
                __stcs(o + 3 * plane, live ? fmaf(uw, t1, p0) : 0.0f);
                __stcs(o + 4 * plane, live ? stage->xyt[2][j][slot] : 0.0f);
            }
        }
    }
}

cudaError_t init_propagate_kernels() { return cudaSuccess; }

template <bool STAGED>
static cudaError_t launch_staged(const DevWorld &w, const PropArgs &a, cudaStream_t stream)
{
    const long long grid = (a.n + kPropSamples - 1) / kPropSamples;
    if (a.n_cons > 0) {
        if (w.n_obs > 0)
            k_propagate<true, true, STAGED><<<(unsigned)grid, kPropThreads, 0, stream>>>(w, a);
        else
            k_propagate<false, true, STAGED><<<(unsigned)grid, kPropThreads, 0, stream>>>(w, a);
    } else if (w.n_obs > 0) {
        k_propagate<true, false, STAGED><<<(unsigned)grid, kPropThreads, 0, stream>>>(w, a);
    } else {
        k_propagate<false, false, STAGED><<<(unsigned)grid, kPropThreads, 0, stream>>>(w, a);
    }
    return cudaGetLastError();
}

cudaError_t launch_propagate(const DevWorld &w, const PropArgs &a, cudaStream_t stream)
{
    if (a.n <= 0) return cudaSuccess;
    // whole segments of up to kStageSteps steps leave through the shared-memory stage (coalesced);
    // longer horizons and final-state-only calls write directly
    if (a.seg != nullptr && a.max_steps <= kStageSteps) return launch_staged<true>(w, a, stream);
    return launch_staged<false>(w, a, stream);
}

}  // namespace kcbs
