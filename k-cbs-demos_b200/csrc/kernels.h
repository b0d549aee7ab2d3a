// kernels.h -- argument blocks and launchers of the three hot-path kernels (internal).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/kcbs_b200.h"

namespace kcbs {

struct DevWorld;

struct ConsDesc;   // kcbs_device.cuh

// What changes from one low-level plan to the next, and from one RRT round to the next, lives on the DEVICE (one 128-byte
// block per planner workspace, uploaded once per plan): the kernels of a round read it instead of taking it as launch
// arguments, so the same launches -- a CUDA graph with a device-driven loop -- serve every round of every plan.
struct RrtPlan {
    // (the first 16 bytes are what the propagation kernels read, as one word)
    int round;                 // index of the current round = rounds completed (kept by k_rrt_append)
    int robot[2];              // robot of car 0 / car 1 (merged pair; = robot[0] for one car)
    int n_cons;
    int cons_total, max_T, max_iterations, node_cap;
    float goal[4];             // (gx, gy) per car
    float goal_radius, goal_bias;
    unsigned long long seed;
    unsigned long long limit_ns, t_start;   // time limit of the device-driven loop (0 = none); start stamp (k_rrt_init)
    float start[10];           // the root state (5 per car)
    int pad[2];
};

struct PropArgs {
    long long n, n_parents;
    const float *parents;      // [5][n_parents]
    const int *parent_idx;     // [n] or null
    const float *controls;     // [2][n]
    const int *steps;          // [n] or null
    float *seg;                // [5][max_steps][n] or null
    float *fin;                // [5][n] or null
    int *valid_steps;          // [n] or null
    int max_steps;
    int robot;
    // dynamic-obstacle constraints (planner only; n_cons == 0 otherwise)
    const int *parent_time;    // [n_parents] absolute time index of each parent, or null (= 0)
    const ConsDesc *cons;      // [n_cons]
    const float *cons_states;  // [3][cons_total]
    int n_cons, cons_total;
    int rounds;                // packed pairs per thread: 1, 2, or 0 = chosen from n (propagate.cu)
    // parent of sample i = parent_idx[(i / parent_group) * parent_stride] (0 = 1: one entry per sample).  The planner hands
    // over its nearest-neighbour keys as they stand: one 64-bit key per target, node index in the low word
    int parent_group, parent_stride;
    // packed (CSR) segment output (offsets != null): seg = [5][packed_cap] planes, offsets = [n + 1] rows
    int *offsets;
    long long packed_cap;
    unsigned long long *scan;  // (spare workspace words behind scan_ctl)
    unsigned *scan_ctl;        // 8-byte aligned row allocator (CTAs served << 40 | rows handed out); zeroed, left zeroed
    // planner rounds (null otherwise): robot, n_cons, cons_total come from the plan block, and controls / steps / parent
    // keys are the buffers of the round's parity: controls + (round & 1) * parity_ctrl, and so on
    const RrtPlan *plan;
    long long parity_ctrl, parity_steps, parity_parent;
};

struct ValidArgs {
    long long n;
    const float *states;       // [5][n]
    uint32_t *bitmask;         // [(n + 31) / 32]
    int robot;
};

struct PairArgs {              // merged pair: 10-d states, 4-d controls
    long long n, n_parents;
    const float *parents;      // [10][n_parents]
    const int *parent_idx;
    const float *controls;     // [4][n]
    const int *steps;
    float *seg;                // [10][max_steps][n] or null
    float *fin;                // [10][n] or null
    int *valid_steps;
    int max_steps, robot1, robot2;
    int parent_group, parent_stride;   // as in PropArgs
    float goals[4];            // gx1, gy1, gx2, gy2
    float hold_radius;         // 1.5, StatePropagatorDatabase.h:170
    // dynamic-obstacle constraints (planner only; n_cons == 0 otherwise): they bind BOTH cars, as the merged
    // individual's areStatesValid does (StateValidityCheckerDatabase.h:264-287)
    const int *parent_time;
    const ConsDesc *cons;
    const float *cons_states;
    int n_cons, cons_total;
    // planner rounds (null otherwise): robots, goals, n_cons, cons_total and the round's parity come from the plan block
    const RrtPlan *plan;
    long long parity_ctrl, parity_steps, parity_parent;
};

struct PairValidArgs {
    long long n;
    const float *states;       // [10][n]
    uint32_t *bitmask;
    int robot1, robot2;
};

struct ConflictArgs {
    const float *traj;         // [P][R][3][T]
    const int *lengths;        // [P][R] or null
    unsigned long long *keys;  // [P], pre-set to kNoConflict
    const int *group;          // [R] individual of each body (same group = never tested against each other) or null
    int *tickets;              // [P] zeroed CTA tickets (fused finish) or null
    kcbs_conflict *out;        // [P] results written by the last CTA of each plan (fused finish) or null
    int P, R, T;
    int k_lo, k_hi;            // scanned time range
    // sharded scan (kcbs_first_conflict_sharded_dev; world <= 1 otherwise): per-plan keys are exchanged through
    // mailboxes in peer memory, box[parity][source rank][box_plans] tagged 64-bit words
    unsigned long long *box_peer[KCBS_MAX_PEERS];  // every rank's mailbox as mapped into this process ([rank] = own)
    unsigned *epoch;           // device word: call counter of this context (tag of this call's messages)
    int *done;                 // device word: plans finished in this call (the last one advances the epoch)
    int *error;                // host-mapped word: set when a wait timed out
    int rank, world, box_plans;
};

struct FinishArgs {
    const float *traj;
    const int *lengths;
    const unsigned long long *keys;
    kcbs_conflict *out;
    int P, R, T;
};

// index of the parent state of sample i (PropArgs / PairArgs)
template <class A>
__host__ __device__ inline long long parent_of(const A &a, long long i)
{
    if (!a.parent_idx) return a.n_parents == 1 ? 0 : i;
    const long long g = a.parent_group > 1 ? i / a.parent_group : i;
    return (long long)a.parent_idx[a.parent_stride > 1 ? g * a.parent_stride : g];
}

cudaError_t launch_propagate(const DevWorld &w, const PropArgs &a, cudaStream_t stream);
int propagate_grid(const PropArgs &a);   // CTAs launch_propagate will use (size of the look-back workspace)
cudaError_t launch_validity(const DevWorld &w, const ValidArgs &a, cudaStream_t stream);
cudaError_t launch_propagate_pair(const DevWorld &w, const PairArgs &a, cudaStream_t stream);
cudaError_t launch_validity_pair(const DevWorld &w, const PairValidArgs &a, cudaStream_t stream);
cudaError_t launch_conflict_keys(const DevWorld &w, const ConflictArgs &a, cudaStream_t stream);
void conflict_time_slab(int T, int rank, int world, int *k_lo, int *k_hi);
cudaError_t launch_conflict_finish(const DevWorld &w, const FinishArgs &a, cudaStream_t stream);
cudaError_t launch_fill_keys(unsigned long long *keys, int *tickets, int n, cudaStream_t stream);
// FFMA saturation probe: returns flops executed; elapsed time is measured by the caller.
cudaError_t launch_fp32_probe(float *sink, int iters, cudaStream_t stream, double *flops);
cudaError_t launch_fp32x2_probe(float *sink, int iters, cudaStream_t stream, double *flops);
cudaError_t init_propagate_kernels();
cudaError_t init_pair_kernels();
cudaError_t init_planner_kernels();
cudaError_t init_validity_kernels();
cudaError_t init_conflict_kernels();  // per-device function attributes; call once per context

}  // namespace kcbs
