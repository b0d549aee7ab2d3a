// pair.cu -- the merged-pair variants of the hot path: after a K-CBS merge two cars are
// planned as one 10-d system, q = (x1, y1, v1, phi1, theta1, x2, y2, v2, phi2, theta2), control = (a1, w1, a2, w2).
//
//   k_propagate_pair  propagateWhileValid of the merged system:
//                       TwoSecondOrderCarsODE + RK4 + wrap            StatePropagatorDatabase.h:77-117
//                       TwoSecondOrderCarStatePropagator::propagate   StatePropagatorDatabase.h:130-153
//                         (a car that was within 1.5 of its goal before the step keeps its state)
//                       homogeneousTwo2ndOrderCarSystemSVC::isValid   StateValidityCheckerDatabase.h:225-261
//   k_validity_pair   that isValid alone for batches of 10-d states
// Both cars run the single-car integrator of car_dynamics.cuh; one thread per sample.  A merge is rare (only after
// more than 10 / 20 conflicts of one pair), so these kernels favour simplicity over the sorting machinery of K1.
#include "car_dynamics.cuh"
#include "kernels.h"

namespace kcbs {

// isValid of the merged system for one 10-d state
__device__ __forceinline__ bool pair_state_valid(const DevWorld &w, int r1, int r2, const float a[5], const float b[5])
{
    if (!in_bounds(w, a[0], a[1], a[2], a[3], a[4]) || !in_bounds(w, b[0], b[1], b[2], b[3], b[4])) return false;
    // the two cars against each other: circle cull, then rectangle SAT (touching counts as collision)
    const float dx = b[0] - a[0], dy = b[1] - a[1];
    const float rs = w.rob_rad[r1] + w.rob_rad[r2];
    if (fmaf(dy, dy, dx * dx) <= rs * rs * 1.000001f) {
        float ca, sa, cb, sb;
        sincosf(a[4], &sa, &ca);
        sincosf(b[4], &sb, &cb);
        if (sat_gap_robots(a[0], a[1], ca, sa, w.rob_hl[r1], w.rob_hw[r1], b[0], b[1], cb, sb, w.rob_hl[r2], w.rob_hw[r2]) <= 0.0f)
            return false;
    }
    if (w.n_obs > 0) {
        if (!obstacles_clear_fine(w, a[0], a[1], a[4], w.rob_hl[r1], w.rob_hw[r1])) return false;
        if (!obstacles_clear_fine(w, b[0], b[1], b[4], w.rob_hl[r2], w.rob_hw[r2])) return false;
    }
    return true;
}

__global__ void __launch_bounds__(128) k_propagate_pair(const __grid_constant__ DevWorld w, const __grid_constant__ PairArgs a_in)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a_in.n) return;
    // planner rounds: robots, goals, constraint counts and the round's parity come from the plan block (kernels.h)
    PairArgs a = a_in;
    if (a_in.plan) {
        const RrtPlan &p = *a_in.plan;
        const int round = __ldcg(&p.round);
        a.robot1 = __ldcg(&p.robot[0]);
        a.robot2 = __ldcg(&p.robot[1]);
        a.n_cons = __ldcg(&p.n_cons);
        a.cons_total = __ldcg(&p.cons_total);
#pragma unroll
        for (int c = 0; c < 4; ++c) a.goals[c] = __ldcg(&p.goal[c]);
        a.controls = a_in.controls + (round & 1) * a_in.parity_ctrl;
        if (a_in.steps) a.steps = a_in.steps + (round & 1) * a_in.parity_steps;
        if (a_in.parent_idx) a.parent_idx = a_in.parent_idx + (round & 1) * a_in.parity_parent;
    }
    const long long pi = parent_of(a, i);
    int st = a.steps ? a.steps[i] : a.max_steps;
    st = min(max(st, 0), a.max_steps);

    float q0[2][5], cur[2][5], u[2][2], off[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
    CarConsts k[2];
    bool frozen[2] = {false, false};
#pragma unroll
    for (int c = 0; c < 2; ++c) {
#pragma unroll
        for (int d = 0; d < 5; ++d) q0[c][d] = cur[c][d] = a.parents[(long long)(5 * c + d) * a.n_parents + pi];
        u[c][0] = a.controls[(long long)(2 * c) * a.n + i];
        u[c][1] = a.controls[(long long)(2 * c + 1) * a.n + i];
        k[c] = car_consts(w, u[c][0], u[c][1]);
    }
    const float hold2 = a.hold_radius * a.hold_radius;
    const int t0 = (a.n_cons > 0 && a.parent_time) ? a.parent_time[pi] : 0;
    const long long plane = (long long)a.max_steps * a.n;
    int nv = 0;
    for (int j = 0; j < st; ++j) {
        float nxt[2][5];
        const float t1 = (float)(j + 1) * w.step;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            // in its goal region before the step: the car keeps its state from now on (the region is absorbing)
            const float gx = cur[c][0] - a.goals[2 * c], gy = cur[c][1] - a.goals[2 * c + 1];
            if (!frozen[c]) frozen[c] = fmaf(gy, gy, gx * gx) < hold2;
            if (frozen[c]) {
#pragma unroll
                for (int d = 0; d < 5; ++d) nxt[c][d] = cur[c][d];
            } else {
                const float toff = car_step(k[c], w.n_sub, cur[c][2], cur[c][3], cur[c][4], off[c][0], off[c][1]);
                nxt[c][0] = q0[c][0] + off[c][0];
                nxt[c][1] = q0[c][1] + off[c][1];
                nxt[c][2] = fmaf(u[c][0], t1, q0[c][2]);
                nxt[c][3] = fmaf(u[c][1], t1, q0[c][3]);
                nxt[c][4] = wrap_angle(cur[c][4] + toff);
            }
        }
        if (!pair_state_valid(w, a.robot1, a.robot2, nxt[0], nxt[1])) break;
        if (a.n_cons > 0) {
            const int k = t0 + j + 1;
            if (!constraints_clear(w, a.cons, a.n_cons, a.cons_states, a.cons_total, k, nxt[0][0], nxt[0][1], nxt[0][4], w.rob_hl[a.robot1],
                                   w.rob_hw[a.robot1], w.rob_rad[a.robot1]) ||
                !constraints_clear(w, a.cons, a.n_cons, a.cons_states, a.cons_total, k, nxt[1][0], nxt[1][1], nxt[1][4], w.rob_hl[a.robot2],
                                   w.rob_hw[a.robot2], w.rob_rad[a.robot2]))
                break;
        }
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
            for (int d = 0; d < 5; ++d) cur[c][d] = nxt[c][d];
        nv = j + 1;
        if (a.seg) {
            float *o = a.seg + (long long)j * a.n + i;
#pragma unroll
            for (int c = 0; c < 10; ++c) __stcs(o + c * plane, cur[c / 5][c % 5]);
        }
    }
    if (a.seg)
        for (int j = nv; j < a.max_steps; ++j) {
            float *o = a.seg + (long long)j * a.n + i;
#pragma unroll
            for (int c = 0; c < 10; ++c) __stcs(o + c * plane, 0.0f);
        }
    if (a.valid_steps) a.valid_steps[i] = nv;
    if (a.fin) {
#pragma unroll
        for (int c = 0; c < 10; ++c) a.fin[(long long)c * a.n + i] = cur[c / 5][c % 5];
    }
}

__global__ void __launch_bounds__(256) k_validity_pair(const __grid_constant__ DevWorld w, const __grid_constant__ PairValidArgs a)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    bool ok = false;
    if (i < a.n) {
        float q[2][5];
#pragma unroll
        for (int c = 0; c < 10; ++c) q[c / 5][c % 5] = __ldcs(a.states + (long long)c * a.n + i);
        ok = pair_state_valid(w, a.robot1, a.robot2, q[0], q[1]);
    }
    const unsigned m = __ballot_sync(0xffffffffu, ok);
    if (lane == 0 && i < a.n) a.bitmask[i >> 5] = m;
}

cudaError_t launch_propagate_pair(const DevWorld &w, const PairArgs &a, cudaStream_t stream)
{
    if (a.n <= 0) return cudaSuccess;
    k_propagate_pair<<<(unsigned)((a.n + 127) / 128), 128, 0, stream>>>(w, a);
    return cudaGetLastError();
}

cudaError_t init_pair_kernels()   // (loaded outside any stream capture, see planner.cu)
{
    cudaFuncAttributes at;
    cudaError_t e = cudaFuncGetAttributes(&at, k_propagate_pair);
    return e != cudaSuccess ? e : cudaFuncGetAttributes(&at, k_validity_pair);
}

cudaError_t launch_validity_pair(const DevWorld &w, const PairValidArgs &a, cudaStream_t stream)
{
    if (a.n <= 0) return cudaSuccess;
    k_validity_pair<<<(unsigned)((a.n + 255) / 256), 256, 0, stream>>>(w, a);
    return cudaGetLastError();
}

}  // namespace kcbs
