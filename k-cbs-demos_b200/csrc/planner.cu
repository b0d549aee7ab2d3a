// planner.cu -- "next" rows of the scope table: a batched kinodynamic RRT low level and a
// minimal K-CBS high level, built on the three hot-path kernels.
//
// They stand in for the planners the reference obtains from the un-vendored OMPL fork:
//   oc::RRT            allocated by allocateControlRRT, PlannerAllocatorDatabase.h:40-45
//   omrc::KCBS         demo_*.cpp:155-168, Benchmark.h:151-173
// with the reference's plugin semantics: controls uniform in [-1, 1]^2 (ControlSpaceDatabase.h:48-51),
// durations UniformInt{1..10} x 0.1 s (demo_*.cpp:115,118), state distance of the compound space
// (StateSpaceDatabase.h:45-46: L2 over (x, y, v, phi) + SO2 arc, weights 1), goal = disc of 0.5 around the goal
// position (GoalRegionDatabase.h:48-55), conflicts = kcbs_first_conflict.
//
// Low level: ONE CUDA graph launch per plan -- k_rrt_init, a WHILE node over the round below (its condition is set on the
// device by k_rrt_append: goal found, tree full, iteration or time limit), then the path extraction, the re-expansion of
// the path's edges and the copies of the results into pinned memory.  What differs from plan to plan is read by the
// kernels from a 128-byte plan block (RrtPlan), so a lane re-uses its instantiated graph for every plan of the same
// launch shape.  KCBS_RRT_ROUNDS=stream queues the same kernels round by round from the host instead (also the fallback
// when a graph cannot be built).  A round (all state in HBM; targets are a pure function of (seed, round, index)):
//   k_rrt_nearest   tiled all-pairs nearest tree node per target (four targets per thread in packed FP32 pairs, nodes
//                   broadcast from shared memory), 64-bit atomicMin on (distance bits, index).  Two launches: the nodes
//                   that existed before the previous round are scanned for the NEXT round's targets on a parallel
//                   branch, the nodes the previous round appended are scanned on the critical path
//   k_propagate     the K1 expansion kernel with dynamic-obstacle constraints and per-parent time indices; the parent of a
//                   sample is its target's nearest node (the keys of k_rrt_nearest as they stand), its control and
//                   duration were drawn one round ahead
//   k_rrt_select    per target the sample that ends closest to it is chosen; the next round's controls and durations
//   k_rrt_append    the chosen samples become tree nodes in target order (goal test; decides whether the loop goes on)
// A found path is re-expanded edge by edge with k_propagate (identical arithmetic) to emit every 0.1 s state.
// Independent plans run on planner lanes (host thread + stream + workspace + graphs each); stream captures and device-wide
// synchronisation exclude each other process wide (capture_gate, ctx.h).
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <queue>
#include <thread>
#include <vector>

#include "ctx.h"

namespace kcbs {

constexpr int kRrtDepth = 4;   // RRT iterations in flight on the stream before the host looks at the oldest one's outcome

// What the host sees of one finished iteration (host-mapped memory, written by k_rrt_append).
struct RrtProgress {
    unsigned long long goal;   // packed (time << 32 | node) of the best goal node, ~0 when none yet
    int nodes, rounds;
};

struct RrtWorkspace {
    int max_nodes = 0, n_samples = 0, n_targets = 0, cons_cap = 0, cons_desc_cap = 0, path_cap = 0, dim = 0;
    float *nodes = nullptr;        // [dim][max_nodes] states (dim = 5 per car)
    int *parent = nullptr;         // [max_nodes]
    int *time = nullptr;           // [max_nodes] absolute time index (propagation steps from the root)
    float *ctrl = nullptr;         // [2 per car][max_nodes] control of the edge into the node
    int *steps = nullptr;          // [max_nodes] duration of the edge into the node
    float *targets = nullptr;      // [2 rounds (parity)][dim][n_targets]
    unsigned long long *nn = nullptr;  // [2 rounds (parity)][n_targets] packed (distance bits << 32 | node); kNNEmpty when unused
    float *s_ctrl = nullptr;       // [2 rounds (parity)][2 per car][n_samples]: drawn one round ahead (k_rrt_select)
    int *s_steps = nullptr;        // [2 rounds (parity)][n_samples]
    float *s_fin = nullptr;        // [dim][n_samples]
    int *s_valid = nullptr;        // [n_samples]
    int *sel = nullptr;            // [n_targets] sample chosen for each target this round, or -1
    float4 *nn_r = nullptr;        // [max_nodes] (x, y, v, phi) of the first car: what the nearest-neighbour scan streams
    float *nn_t = nullptr;         // [max_nodes] its heading
    int *counters = nullptr;       // [0], [1] = node count (ring by round parity), [2] = rounds completed, [3] = append ticket,
                                   // [4] = edges of the extracted path, [5] = a goal node exists (set between rounds),
                                   // [6], [7] = round and CTA tickets of the side scan (k_rrt_nearest)
    unsigned long long *goal_key = nullptr;  // packed (time << 32 | node) of the best goal node
    ConsDesc *cons = nullptr;
    float *cons_states = nullptr;
    // path extraction
    int *path_nodes = nullptr;     // [path_cap]
    float *path_ctrl = nullptr;    // [2 per car][path_cap]
    int *path_steps = nullptr, *path_parent = nullptr, *path_valid = nullptr;
    float *path_seg = nullptr;     // [dim][KCBS_MAX_STEPS][path_cap]
    RrtProgress *h_progress = nullptr, *d_progress = nullptr;   // [kRrtDepth] host-mapped ring
    int *h_counters = nullptr;                                  // pinned: [8] counters, then the goal key (2 words)
    float *h_path_seg = nullptr;                                // pinned: the re-expanded path [dim][max_steps][e_max]
    int *h_path_valid = nullptr;                                // pinned [e_max]
    size_t h_path_cap = 0;                                      // floats h_path_seg holds
    cudaEvent_t ev[kRrtDepth] = {};
    RrtPlan *plan = nullptr, *h_plan = nullptr;   // the plan block on the device (kernels.h) and its pinned staging copy
    unsigned long long *trace = nullptr;          // (KCBS_RRT_TRACE builds) [kTraceRounds][4 kernels][start, end] globaltimer stamps
    void *slab = nullptr, *h_slab = nullptr;      // the one device and the one pinned (mapped) allocation everything above is carved from
    bool own_cons_states = false, own_cons = false, own_h_path_seg = false;   // outgrew the slab: allocated on their own
    cudaStream_t side = nullptr;                  // second branch of a round (the scan of the old nodes for the next round)
    cudaEvent_t fork = nullptr, join = nullptr;
    // device-driven rounds: one instantiated graph per (cars, constrained or not), rebuilt when the launch shape changes
    struct Loop {
        cudaGraphExec_t exec = nullptr;
        int key[6] = {0, 0, 0, 0, 0, 0};   // n_targets, n_controls, min_steps, max_steps, expansion shape, path edge slots
        bool failed = false;               // the graph of this key could not be built: rounds are queued by the host instead
    } loop[2][2][3];   // [cars - 1][constrained][expansion shape]: a lane that alternates between shapes keeps both graphs
};

namespace {

__device__ __forceinline__ unsigned hash32(unsigned x)
{
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}
// counter-based uniform in [0, 1): (seed, round, index, stream) -> 24 random bits
__device__ __forceinline__ float urand(unsigned long long seed, unsigned round, unsigned idx, unsigned stream)
{
    unsigned h = hash32((unsigned)seed ^ 0x9e3779b9u);
    h = hash32(h ^ (unsigned)(seed >> 32));
    h = hash32(h ^ round * 0x85ebca6bu);
    h = hash32(h ^ idx * 0xc2b2ae35u);
    h = hash32(h ^ stream * 0x27d4eb2fu);
    return (float)(h >> 8) * (1.0f / 16777216.0f);
}

// "No neighbour yet": larger than every real key, and its low word is a valid node index (the root), so a cancelled round's
// propagation -- the only kernel of a round that cannot test rrt_done -- still reads a parent that exists.
constexpr unsigned long long kNNEmpty = 0xffffffff00000000ull;

// Launch arguments of the planner kernels.  Everything from `goal` on is NOT filled by the host: it is loaded from the
// plan block on the device at the top of every kernel (rrt_load_plan), see RrtPlan.
struct RrtArgs {
    RrtWorkspace ws;
    float lo[4], hi[4];
    int n_targets, n_controls, min_steps, max_steps;
    cudaGraphConditionalHandle loop;   // the WHILE node of the device-driven loop (in_loop != 0), set by k_rrt_append
    int in_loop;
    float goal[4];             // (gx, gy) per car
    float goal_radius;         // one car: distanceGoal <= radius (GoalRegionDatabase.h:48-55); merged pair: both < radius (:61-108)
    float goal_bias;
    unsigned long long seed;
    unsigned round;
    int robot[2];
    int n_cons, cons_total;
    int max_T;  // states the caller's trajectory buffer holds: a goal node must have time index < max_T
    int max_iterations;
    unsigned long long limit_ns, t_start;
};

static_assert(offsetof(RrtPlan, goal) == 32 && offsetof(RrtPlan, goal_radius) == 48 && offsetof(RrtPlan, seed) == 56 &&
                  offsetof(RrtPlan, limit_ns) == 64 && offsetof(RrtPlan, start) == 80 && sizeof(RrtPlan) == 128,
              "rrt_load_plan reads the block as five 16-byte words");
__device__ __forceinline__ void rrt_load_plan(RrtArgs &a)
{
    const int4 *p = reinterpret_cast<const int4 *>(a.ws.plan);
    const int4 h0 = __ldcg(p), h1 = __ldcg(p + 1), h2 = __ldcg(p + 2), h3 = __ldcg(p + 3), h4 = __ldcg(p + 4);
    a.round = (unsigned)h0.x; a.robot[0] = h0.y; a.robot[1] = h0.z; a.n_cons = h0.w;
    a.cons_total = h1.x; a.max_T = h1.y; a.max_iterations = h1.z;
    a.goal[0] = __int_as_float(h2.x); a.goal[1] = __int_as_float(h2.y); a.goal[2] = __int_as_float(h2.z); a.goal[3] = __int_as_float(h2.w);
    a.goal_radius = __int_as_float(h3.x); a.goal_bias = __int_as_float(h3.y);
    a.seed = (unsigned long long)(unsigned)h3.z | ((unsigned long long)(unsigned)h3.w << 32);
    a.limit_ns = (unsigned long long)(unsigned)h4.x | ((unsigned long long)(unsigned)h4.y << 32);
    a.t_start = (unsigned long long)(unsigned)h4.z | ((unsigned long long)(unsigned)h4.w << 32);
}
__device__ __forceinline__ unsigned long long global_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
constexpr int kTraceRounds = 128;
// (debug builds with -DKCBS_RRT_TRACE) first start / last end of kernel k of a round over its CTAs
__device__ __forceinline__ void trace_stamp(const RrtArgs &a, unsigned round, int k, int end)
{
#ifdef KCBS_RRT_TRACE
    if (threadIdx.x == 0 && round < (unsigned)kTraceRounds) {
        unsigned long long *w = a.ws.trace + ((size_t)round * 4 + k) * 2 + end;
        if (end) atomicMax(w, global_ns());
        else atomicMin(w, global_ns());
    }
#endif
}

// Every kernel of an iteration starts with this test: once a goal node exists the remaining iterations already queued
// on the stream do nothing, so the outcome is that of the first iteration that found one -- however far the host ran ahead.
// The flag is raised by the last CTA of the round that found a goal node (k_rrt_append), never while a kernel that
// tests it is running: every thread of every kernel of a round sees the same answer (the goal key itself changes while
// k_rrt_append runs).
__device__ __forceinline__ bool rrt_done(const RrtArgs &a) { return *(volatile const int *)(a.ws.counters + 5) != 0; }

// ob::CompoundStateSpace::distance with unit weights, per car: L2 over (x, y, v, phi) + SO2 arc length.
__device__ __forceinline__ float car_distance(const float *a, const float *b)
{
    const float dx = a[0] - b[0], dy = a[1] - b[1], dv = a[2] - b[2], dp = a[3] - b[3];
    float d = fabsf(a[4] - b[4]);
    d = d > kPi ? kTwoPi - d : d;
    return sqrtf(fmaf(dx, dx, fmaf(dy, dy, fmaf(dv, dv, dp * dp)))) + d;
}

// Random target state of a round: uniform over the state bounds, or (goal bias) the goal positions with random
// v, phi, theta.  A pure function of (seed, round, t), so any kernel can recompute it.
template <int NC>
__device__ __forceinline__ void rrt_target(const RrtArgs &a, unsigned round, int t, float q[5 * NC])
{
    const bool to_goal = urand(a.seed, round, t, 5) < a.goal_bias;
#pragma unroll
    for (int car = 0; car < NC; ++car) {
#pragma unroll
        for (int c = 0; c < 4; ++c) q[5 * car + c] = a.lo[c] + (a.hi[c] - a.lo[c]) * urand(a.seed, round, t, 16 * car + c);
        q[5 * car + 4] = -kPi + kTwoPi * urand(a.seed, round, t, 16 * car + 4);
        if (to_goal) {
            q[5 * car + 0] = a.goal[2 * car];
            q[5 * car + 1] = a.goal[2 * car + 1];
        }
    }
}

constexpr int kNNThreads = 128;
constexpr int kNNPer = 4;                            // targets per thread (two packed pairs)
constexpr int kNNTile = kNNThreads * kNNPer;         // targets per CTA
constexpr int kNNChunkMax = 512;   // nodes per CTA at most (the chunk shrinks while the tree is small, so that the grid fills 148 SMs)

// nodes per CTA for a tree of at most `bound` nodes and `tiles` target tiles: about 600 CTAs, multiples of 32, within
// [32, kNNChunkMax]
__host__ __device__ inline int nn_chunk(int n_nodes, int tiles)
{
    const int chunks = 600 / tiles > 1 ? 600 / tiles : 1;
    const int c = (n_nodes / chunks + 31) / 32 * 32;
    return c < 32 ? 32 : (c > kNNChunkMax ? kNNChunkMax : c);
}
// chunks of the largest tree: the grid of every round (CTAs past the round's node count return at once)
inline int nn_grid_y(int max_nodes, int tiles)
{
    return std::max(std::max(1, 600 / tiles), (max_nodes + kNNChunkMax - 1) / kNNChunkMax);
}

// node count of the tree at the start of `round`: counters live in a two-entry ring indexed by the round's parity, so a
// round's kernels read one entry while k_rrt_append writes the other
__device__ __forceinline__ int rrt_node_count(const RrtArgs &a) { return min(a.ws.counters[a.round & 1u], a.ws.max_nodes); }

typedef unsigned long long nn_f32x2;
__device__ __forceinline__ nn_f32x2 nn_pk(float lo, float hi)
{
    nn_f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ nn_f32x2 nn_sub2(nn_f32x2 a, nn_f32x2 b)
{
    nn_f32x2 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ nn_f32x2 nn_mul2(nn_f32x2 a, nn_f32x2 b)
{
    nn_f32x2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ nn_f32x2 nn_fma2(nn_f32x2 a, nn_f32x2 b, nn_f32x2 c)
{
    nn_f32x2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ float nn_sqrt(float x)
{
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// grid = (ceil(n_targets / 512), chunks of the largest tree, nn_grid_y): every thread owns FOUR targets (two
// packed FP32 pairs) and scans one chunk of nodes broadcast from shared memory (car 0 with every component duplicated, so
// that a node is a packed operand as it stands; the node count is read on the device: the host does not know it).
// Straight-line code per node -- squared L2 over (x, y, v, phi), MUFU square root, SO2 arc -- and one unsigned minimum on
// (distance bits with the low 9 bits replaced by the node's place in the chunk): the distance is compared to 2^-14
// relative, ties go to the smallest index, nothing diverges.
template <int NC>
__global__ void __launch_bounds__(kNNThreads) k_rrt_nearest(RrtArgs a, int fresh)
{
    constexpr int D = 5 * NC;
    static_assert(kNNChunkMax <= 512, "the node's place in the chunk takes 9 bits of the key");
    __shared__ float2 s_n[5][kNNChunkMax];             // (c, c) for c = x, y, v, phi, theta of car 0
    __shared__ float s_o[NC == 2 ? 5 : 1][kNNChunkMax];   // car 1 (merged pair only)
    rrt_load_plan(a);
    // The scan of a round is split in two.  fresh = 0: the targets of the NEXT round against the nodes the tree has when
    // this round starts -- nothing of this round is needed for it, so it runs beside the round's other kernels.
    // fresh = 1: the targets of THIS round against the nodes the previous round appended (a few thousand at most).
    // Together they cover the whole tree; both write the keys / targets of their round's parity.
    // The side scan may still be running when k_rrt_append moves the plan block to the next round, so it counts rounds
    // for itself: counters[6], advanced by the last of its CTAs to finish (counters[7] = their tickets).
    __shared__ int s_side_round;
    if (!fresh) {
        if (threadIdx.x == 0) s_side_round = *(volatile int *)(a.ws.counters + 6);
        __syncthreads();
        a.round = (unsigned)s_side_round;
    }
    trace_stamp(a, a.round, fresh ? 1 : 0, 0);
    auto leave = [&]() {   // every CTA of the side scan, whatever it did
        trace_stamp(a, a.round, fresh ? 1 : 0, 1);
        if (!fresh) {
            __syncthreads();
            if (threadIdx.x == 0 && atomicAdd(a.ws.counters + 7, 1) == (int)(gridDim.x * gridDim.y) - 1) {
                a.ws.counters[7] = 0;
                *(volatile int *)(a.ws.counters + 6) = (int)a.round + 1;
            }
        }
    };
    if (rrt_done(a)) return leave();
    const unsigned tr = fresh ? a.round : a.round + 1u;
    const int n_now = rrt_node_count(a);
    const int lo = fresh ? min(a.ws.counters[(a.round + 1u) & 1u], n_now) : 0;   // (ring: the other entry still holds the previous count)
    const int n_nodes = n_now - lo;
    const int chunk = nn_chunk(n_nodes, (int)gridDim.x);
    const int n0 = lo + blockIdx.y * chunk, n1 = min(n0 + chunk, n_now);
    if (n0 >= n_now) return leave();
    const int B = a.n_targets, M = a.ws.max_nodes;
    const int m = n1 - n0;
    for (int j = threadIdx.x; j < m; j += kNNThreads) {
        const float4 r = a.ws.nn_r[n0 + j];
        const float th = a.ws.nn_t[n0 + j];
        s_n[0][j] = make_float2(r.x, r.x);
        s_n[1][j] = make_float2(r.y, r.y);
        s_n[2][j] = make_float2(r.z, r.z);
        s_n[3][j] = make_float2(r.w, r.w);
        s_n[4][j] = make_float2(th, th);
        if (NC == 2)
            for (int c = 0; c < 5; ++c) s_o[c][j] = a.ws.nodes[(5 + c) * M + n0 + j];
    }
    // the thread's targets: t0 + {0, 1, 2, 3} * kNNThreads (coalesced stores of the target planes)
    const int t0 = blockIdx.x * kNNTile + threadIdx.x;
    float q[kNNPer][D];
#pragma unroll
    for (int u = 0; u < kNNPer; ++u) {
        const int t = t0 + u * kNNThreads;
        rrt_target<NC>(a, tr, min(t, B - 1), q[u]);
        // (the targets of round 0 are written by its fresh scan: no earlier round scanned for it)
        if (t < B && blockIdx.y == 0 && (!fresh || a.round == 0u))
            for (int c = 0; c < D; ++c) a.ws.targets[(size_t)(tr & 1u) * D * B + c * B + t] = q[u][c];
    }
    nn_f32x2 Q[2][5];
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int c = 0; c < 5; ++c) Q[h][c] = nn_pk(q[2 * h][c], q[2 * h + 1][c]);
    __syncthreads();
    unsigned best[kNNPer] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu};
#pragma unroll 4
    for (int j = 0; j < m; ++j) {
        nn_f32x2 N[5];
#pragma unroll
        for (int c = 0; c < 5; ++c) N[c] = *reinterpret_cast<const nn_f32x2 *>(&s_n[c][j]);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const nn_f32x2 dx = nn_sub2(N[0], Q[h][0]), dy = nn_sub2(N[1], Q[h][1]), dv = nn_sub2(N[2], Q[h][2]), dp = nn_sub2(N[3], Q[h][3]);
            const nn_f32x2 l2 = nn_fma2(dx, dx, nn_fma2(dy, dy, nn_fma2(dv, dv, nn_mul2(dp, dp))));
            const nn_f32x2 dt = nn_sub2(N[4], Q[h][4]);
            float l[2], t[2];
            asm("mov.b64 {%0, %1}, %2;" : "=f"(l[0]), "=f"(l[1]) : "l"(l2));
            asm("mov.b64 {%0, %1}, %2;" : "=f"(t[0]), "=f"(t[1]) : "l"(dt));
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const float at = fabsf(t[e]);
                float d = nn_sqrt(l[e]) + fminf(at, kTwoPi - at);
                if (NC == 2) {
                    float o[5];
#pragma unroll
                    for (int c = 0; c < 5; ++c) o[c] = s_o[c][j];
                    d += car_distance(o, q[2 * h + e] + 5 * (NC - 1));
                }
                best[2 * h + e] = min(best[2 * h + e], (__float_as_uint(d) & 0xfffffe00u) | (unsigned)j);
            }
        }
    }
#pragma unroll
    for (int u = 0; u < kNNPer; ++u) {
        const int t = t0 + u * kNNThreads;
        if (t < B) atomicMin(a.ws.nn + (size_t)(tr & 1u) * B + t, ((unsigned long long)(best[u] & 0xfffffe00u) << 32) | (unsigned)(n0 + (int)(best[u] & 0x1ffu)));
    }
    leave();
}

// Control and duration of sample s of round `round` (sample s belongs to target s / n_controls; its parent is that
// target's nearest node): a pure function of (seed, round, s), written into the buffer of the round's parity.
template <int NC>
__device__ __forceinline__ void rrt_write_sample(const RrtArgs &a, unsigned round, int s)
{
    const int n = a.n_targets * a.n_controls;
    float *ctrl = a.ws.s_ctrl + (size_t)(round & 1u) * 2 * NC * n;
#pragma unroll
    for (int car = 0; car < NC; ++car) {
        ctrl[(2 * car) * n + s] = -1.0f + 2.0f * urand(a.seed, round, s, 8 + 16 * car);       // ControlSpaceDatabase.h:48-51
        ctrl[(2 * car + 1) * n + s] = -1.0f + 2.0f * urand(a.seed, round, s, 9 + 16 * car);
    }
    const int span = a.max_steps - a.min_steps + 1;
    a.ws.s_steps[(size_t)(round & 1u) * n + s] = a.min_steps + min(span - 1, (int)(urand(a.seed, round, s, 10) * (float)span));
}

template <int NC>
__global__ void k_rrt_init(RrtArgs a)
{
    rrt_load_plan(a);
    const float *start = a.ws.plan->start;
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        a.ws.plan->round = 0;
        a.ws.plan->t_start = global_ns();
        const int M = a.ws.max_nodes;
        for (int c = 0; c < 5 * NC; ++c) a.ws.nodes[c * M] = start[c];
        a.ws.nn_r[0] = make_float4(start[0], start[1], start[2], start[3]);
        a.ws.nn_t[0] = start[4];
        for (int c = 0; c < 2 * NC; ++c) a.ws.ctrl[c * M] = 0.f;
        a.ws.parent[0] = -1; a.ws.time[0] = 0; a.ws.steps[0] = 0;
        a.ws.counters[0] = 1;   // node count seen by round 0 (parity ring, see rrt_node_count)
        a.ws.counters[1] = 0;   // ... and "the count before": round 0's fresh scan covers the root
        a.ws.counters[2] = 0;   // rounds completed
        a.ws.counters[3] = 0;   // CTA ticket of k_rrt_append
        a.ws.counters[5] = 0;   // a goal node exists (rrt_done)
        a.ws.counters[6] = 0;   // round of the side scan (k_rrt_nearest)
        a.ws.counters[7] = 0;   // ... and the tickets of its CTAs
        *a.ws.goal_key = ~0ull;
    }
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < 2 * a.n_targets) a.ws.nn[t] = kNNEmpty;   // (both parities)
    if (t < a.n_targets * a.n_controls) rrt_write_sample<NC>(a, 0u, t);   // (grid = one thread per sample)
}

// A robot waiting at its final state must also respect the constraints that come later.
__device__ bool parked_state_clear(const DevWorld &w, const RrtArgs &a, int robot, int t_node, float x, float y, float th)
{
    const float hl = w.rob_hl[robot], hw = w.rob_hw[robot], rad = w.rob_rad[robot];
    float s, c;
    sincosf(th, &s, &c);
    for (int q = 0; q < a.n_cons; ++q) {
        const ConsDesc d = a.ws.cons[q];
        const int k_end = d.k0 + abs(d.len);
        int k_begin = max(t_node + 1, d.k0);
        if (d.len < 0) k_begin = min(k_begin, k_end - 1);  // persistent: the last state holds from k_end - 1 on
        for (int k = k_begin; k < k_end; ++k) {
            const float *st = a.ws.cons_states + d.offset + (k - d.k0);
            const float ox = st[0], oy = st[a.cons_total], dx = ox - x, dy = oy - y;
            const float rs = rad + w.rob_rad[d.other];
            if (fmaf(dy, dy, dx * dx) > rs * rs * 1.000001f) continue;
            float oc, os;
            sincosf(st[2 * a.cons_total], &os, &oc);
            if (sat_gap_robots(x, y, c, s, hl, hw, ox, oy, oc, os, w.rob_hl[d.other], w.rob_hw[d.other]) <= 0.0f)
                return false;
        }
    }
    return true;
}

// One warp per target: the valid sample that ends closest to the target becomes a tree node.
template <int NC>
__global__ void __launch_bounds__(256) k_rrt_select(RrtArgs a)
{
    constexpr int D = 5 * NC;
    if (rrt_done(a)) return;
    rrt_load_plan(a);
    trace_stamp(a, a.round, 2, 0);
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= a.n_targets) return;
    const int B = a.n_targets, C = a.n_controls, n = B * C;
    float q[D];
    for (int c = 0; c < D; ++c) q[c] = a.ws.targets[(size_t)(a.round & 1u) * D * B + c * B + warp];
    float best = 3.0e38f;
    int best_s = -1;
    for (int j = lane; j < C; j += 32) {
        const int s = warp * C + j;
        if (a.ws.s_valid[s] >= a.min_steps) {
            float f[D];
            for (int c = 0; c < D; ++c) f[c] = a.ws.s_fin[c * n + s];
            float d = car_distance(f, q);
            if (NC == 2) d += car_distance(f + 5 * (NC - 1), q + 5 * (NC - 1));
            if (d < best) {
                best = d;
                best_s = s;
            }
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, off);
        const int os = __shfl_xor_sync(0xffffffffu, best_s, off);
        if (ob < best || (ob == best && os >= 0 && (best_s < 0 || os < best_s))) {
            best = ob;
            best_s = os;
        }
    }
    if (lane == 0) a.ws.sel[warp] = best_s;
    // the controls and durations of the next round's samples of this target (other parity: this round's are still needed)
    for (int j = lane; j < C; j += 32) rrt_write_sample<NC>(a, a.round + 1u, warp * C + j);
    trace_stamp(a, a.round, 2, 1);
}

// Appends the chosen samples to the tree in TARGET ORDER, so that node indices -- and with them nearest-neighbour ties,
// the goal node and the whole plan -- depend only on the seed, never on the order in which warps of k_rrt_select happen
// to finish.  CTA b owns targets [128 b, 128 b + 128); the position of a target's node is the node count plus the number of
// chosen samples before it, which every CTA counts for itself over the (tiny) selection array.  The kernel is latency
// bound (a few thousand scattered nodes), so every level of loads is issued for all of a thread's work before anything
// waits: the selection array, then the chosen sample, then the parent's time index.  The last CTA to finish publishes the
// new node count (other parity, see rrt_node_count) and reports the round to the host.
constexpr int kAppendThreads = 128;
constexpr int kAppendScan = 16;   // selection entries a thread counts per pass (in flight together)
template <int NC>
__global__ void __launch_bounds__(kAppendThreads) k_rrt_append(const __grid_constant__ DevWorld w, RrtArgs a)
{
    constexpr int D = 5 * NC, U = 2 * NC, kWarps = kAppendThreads / 32;
    __shared__ int s_before[kWarps], s_total[kWarps], s_own[kWarps];
    __shared__ int s_last;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int B = a.n_targets, C = a.n_controls, n = B * C, M = a.ws.max_nodes;
    rrt_load_plan(a);
    RrtProgress *slot = a.ws.d_progress + (a.round % kRrtDepth);
    trace_stamp(a, a.round, 3, 0);
    if (rrt_done(a)) {   // (uniform, see rrt_done)
        if (tid == 0 && blockIdx.x == 0) {
            if (a.in_loop) cudaGraphSetConditional(a.loop, 0u);
            slot->goal = *a.ws.goal_key;
            slot->nodes = a.ws.counters[a.round & 1u];
            slot->rounds = a.ws.counters[2];
            a.ws.counters[(a.round + 1u) & 1u] = a.ws.counters[a.round & 1u];
        }
        return;
    }
    const int n_nodes = rrt_node_count(a);
    const int first = blockIdx.x * kAppendThreads;   // this CTA's first target
    const int t = first + tid;
    const int s_idx = t < B ? a.ws.sel[t] : -1;
    // chosen samples before this CTA's targets, and in all targets
    int before = 0, total = 0;
    for (int j0 = 0; j0 < B; j0 += kAppendScan * kAppendThreads) {
        int v[kAppendScan];
#pragma unroll
        for (int q = 0; q < kAppendScan; ++q) {
            const int j = j0 + q * kAppendThreads + tid;
            v[q] = j < B ? a.ws.sel[j] : -1;
        }
#pragma unroll
        for (int q = 0; q < kAppendScan; ++q) {
            const int j = j0 + q * kAppendThreads + tid;
            total += v[q] >= 0;
            before += v[q] >= 0 && j < first;
        }
    }
    // the chosen sample, while the counts are reduced
    const int s = max(s_idx, 0);
    // (the parent of a target's samples = the node index in the low word of its nearest-neighbour key)
    unsigned long long *nn = a.ws.nn + (size_t)(a.round & 1u) * B;   // this round's keys
    const int par = t < B ? (int)(nn[t] & 0xffffffffu) : 0, nv = a.ws.s_valid[s];
    const float *ctrl = a.ws.s_ctrl + (size_t)(a.round & 1u) * U * n;
    float f[D], u[U];
#pragma unroll
    for (int c = 0; c < D; ++c) f[c] = a.ws.s_fin[c * n + s];
#pragma unroll
    for (int c = 0; c < U; ++c) u[c] = ctrl[c * n + s];
    if (t < B) nn[t] = kNNEmpty;   // clean for the round after the next (same parity)
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        before += __shfl_xor_sync(0xffffffffu, before, off);
        total += __shfl_xor_sync(0xffffffffu, total, off);
    }
    const unsigned bal = __ballot_sync(0xffffffffu, s_idx >= 0);
    if (lane == 0) {
        s_before[wid] = before;
        s_total[wid] = total;
        s_own[wid] = __popc(bal);
    }
    const int t_par = a.ws.time[par];
    __syncthreads();
    int bsum = 0, tsum = 0, in_cta = 0;
#pragma unroll
    for (int q = 0; q < kWarps; ++q) {
        bsum += s_before[q];
        tsum += s_total[q];
        in_cta += q < wid ? s_own[q] : 0;
    }
    const int idx = n_nodes + bsum + in_cta + __popc(bal & ((1u << lane) - 1u));
    if (s_idx >= 0 && idx < M) {
#pragma unroll
        for (int c = 0; c < D; ++c) a.ws.nodes[c * M + idx] = f[c];
        a.ws.nn_r[idx] = make_float4(f[0], f[1], f[2], f[3]);
        a.ws.nn_t[idx] = f[4];
        a.ws.parent[idx] = par;
        const int t_node = t_par + nv;
        a.ws.time[idx] = t_node;
#pragma unroll
        for (int c = 0; c < U; ++c) a.ws.ctrl[c * M + idx] = u[c];
        a.ws.steps[idx] = nv;
        // (a path that would not fit the caller's buffer is never a solution: it would come back truncated)
        bool in_goal = t_node < a.max_T;
#pragma unroll
        for (int car = 0; car < NC; ++car) {
            const float gx = f[5 * car] - a.goal[2 * car], gy = f[5 * car + 1] - a.goal[2 * car + 1];
            const float d = sqrtf(fmaf(gx, gx, gy * gy));
            // GoalRegion2ndOrderCar: distanceGoal <= threshold (GoalRegionDatabase.h:48-55);
            // GoalRegionTwo2ndOrderCars: both distances < threshold (:61-108)
            in_goal &= NC == 1 ? d <= a.goal_radius : d < a.goal_radius;
        }
        if (in_goal && a.n_cons > 0) {
#pragma unroll
            for (int car = 0; car < NC; ++car)
                in_goal = in_goal && parked_state_clear(w, a, a.robot[car], t_node, f[5 * car], f[5 * car + 1], f[5 * car + 4]);
        }
        if (in_goal) atomicMin(a.ws.goal_key, ((unsigned long long)(unsigned)t_node << 32) | (unsigned)idx);
    }
    // the last CTA of the round publishes the node count and reports to the host
    trace_stamp(a, a.round, 3, 1);
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        s_last = atomicAdd(a.ws.counters + 3, 1) == (int)gridDim.x - 1;
    }
    __syncthreads();
    if (s_last && tid == 0) {
        __threadfence();
        const int count = min(n_nodes + tsum, M);
        const unsigned long long goal = *(volatile unsigned long long *)a.ws.goal_key;
        a.ws.counters[(a.round + 1u) & 1u] = count;
        a.ws.counters[2] = (int)a.round + 1;
        a.ws.counters[3] = 0;
        a.ws.counters[5] = goal != ~0ull;
        a.ws.plan->round = (int)a.round + 1;
        if (a.in_loop) {
            // another round?  (what the host decides between rounds when it queues them itself)
            const bool more = goal == ~0ull && count < M && (int)a.round + 1 < a.max_iterations &&
                              (a.limit_ns == 0 || global_ns() - a.t_start < a.limit_ns);
            cudaGraphSetConditional(a.loop, more ? 1u : 0u);
        }
        slot->goal = goal;
        slot->nodes = count;
        slot->rounds = (int)a.round + 1;
    }
}

// Walks goal -> root (one lane: a chain of dependent loads) and emits the edges root -> goal as propagation inputs, one
// lane per edge: [e_max] entries with planes of stride e_max, the entries past the path's edges with duration 0 (they
// expand to nothing), so that the launch that re-expands the path is the same for every plan.  One warp.
template <int NC>
__global__ void k_rrt_path(RrtArgs a, int e_max)
{
    const int M = a.ws.max_nodes, lane = threadIdx.x;
    const unsigned long long goal = *a.ws.goal_key;
    int n = 0;
    if (lane == 0 && goal != ~0ull)
        for (int i = (int)(goal & 0xffffffffu); a.ws.parent[i] >= 0 && n < e_max; i = a.ws.parent[i]) a.ws.path_nodes[n++] = i;
    n = __shfl_sync(0xffffffffu, n, 0);
    __syncwarp();
    for (int e = lane; e < e_max; e += 32) {
        const bool on = e < n;
        const int i = on ? a.ws.path_nodes[n - 1 - e] : 0;
        a.ws.path_parent[e] = on ? a.ws.parent[i] : 0;
        for (int c = 0; c < 2 * NC; ++c) a.ws.path_ctrl[c * e_max + e] = on ? a.ws.ctrl[c * M + i] : 0.0f;
        a.ws.path_steps[e] = on ? a.ws.steps[i] : 0;
    }
    if (lane == 0) a.ws.counters[4] = n;
}

template <class T>
int dev_alloc(kcbs_ctx *ctx, T *&p, size_t n)
{
    KCBS_CUDA(ctx, cudaMalloc((void **)&p, sizeof(T) * (n ? n : 1)));
    return KCBS_OK;
}

template <int NC>
cudaError_t load_planner_kernels()
{
    cudaFuncAttributes at;
    cudaError_t e;
    if ((e = cudaFuncGetAttributes(&at, k_rrt_init<NC>)) != cudaSuccess || (e = cudaFuncGetAttributes(&at, k_rrt_nearest<NC>)) != cudaSuccess ||
        (e = cudaFuncGetAttributes(&at, k_rrt_select<NC>)) != cudaSuccess || (e = cudaFuncGetAttributes(&at, k_rrt_append<NC>)) != cudaSuccess)
        return e;
    return cudaFuncGetAttributes(&at, k_rrt_path<NC>);
}

// the instantiated round loops hold the workspace's pointers: they go whenever a buffer moves
void drop_loops(RrtWorkspace &ws)
{
    for (auto &per_cars : ws.loop)
        for (auto &per_cons : per_cars)
            for (RrtWorkspace::Loop &l : per_cons) {
                if (l.exec) cudaGraphExecDestroy(l.exec);
                l = RrtWorkspace::Loop();
            }
}

void free_workspace(RrtWorkspace &ws)
{
    drop_loops(ws);
    // one device slab and one pinned slab hold everything; what outgrew its share was allocated on its own
    if (ws.own_cons_states && ws.cons_states) cudaFree(ws.cons_states);
    if (ws.own_cons && ws.cons) cudaFree(ws.cons);
    if (ws.own_h_path_seg && ws.h_path_seg) cudaFreeHost(ws.h_path_seg);
    if (ws.slab) cudaFree(ws.slab);
    if (ws.h_slab) cudaFreeHost(ws.h_slab);
    if (ws.side) cudaStreamDestroy(ws.side);
    if (ws.fork) cudaEventDestroy(ws.fork);
    if (ws.join) cudaEventDestroy(ws.join);
    for (cudaEvent_t e : ws.ev)
        if (e) cudaEventDestroy(e);
    ws = RrtWorkspace();
}

// Hands out 256-byte aligned pieces of a slab; with a null base it only adds up the sizes.
struct Carver {
    char *base;
    size_t off = 0;
    explicit Carver(void *b) : base((char *)b) {}
    template <class T>
    void take(T *&p, size_t n)
    {
        p = base ? reinterpret_cast<T *>(base + off) : nullptr;
        off += (sizeof(T) * (n ? n : 1) + 255) & ~(size_t)255;
    }
};

// Everything that touches a planner workspace runs on the lane's own stream or on the workspace's side stream: waiting
// for those two is all a re-allocation needs (never the device: other lanes are at work, or capturing, see capture_gate).
cudaError_t sync_lane(kcbs_ctx *ctx, RrtWorkspace &ws)
{
    const cudaError_t e = cudaStreamSynchronize(ctx->stream);
    return e != cudaSuccess || !ws.side ? e : cudaStreamSynchronize(ws.side);
}

// A lane's first plan used to cost 8 ms of allocation calls (28 of them, serialised by the driver over 16 lanes): the
// workspace is ONE device allocation and ONE pinned (mapped) allocation, carved up here.
int ensure_workspace(kcbs_ctx *ctx, const kcbs_rrt_params &p, int n_cars, int cons_total, int n_cons, size_t seg_floats)
{
    if (!ctx->rrt) ctx->rrt = new RrtWorkspace();
    RrtWorkspace &ws = *ctx->rrt;
    const int n_samples = p.n_targets * p.n_controls, dim = 5 * n_cars;
    if (ws.max_nodes != p.max_nodes || ws.n_samples != n_samples || ws.n_targets != p.n_targets || ws.dim < dim) {
        KCBS_CUDA(ctx, sync_lane(ctx, ws));
        std::lock_guard<std::mutex> gate(capture_gate());   // (frees and allocations stay away from captures in progress)
        free_workspace(ws);
        const size_t M = (size_t)p.max_nodes, S = (size_t)n_samples, B = (size_t)p.n_targets, D = (size_t)dim, U = (size_t)2 * n_cars;
        ws.path_cap = 8192;
        const size_t PC = (size_t)ws.path_cap;
        const int cons_cap = std::max(4096, cons_total + cons_total / 2 + 1024), cons_desc_cap = std::max(64, n_cons + 64);
        auto device_layout = [&](Carver &c) {
            c.take(ws.nodes, D * M); c.take(ws.parent, M); c.take(ws.time, M); c.take(ws.ctrl, U * M); c.take(ws.steps, M);
            c.take(ws.targets, 2 * D * B); c.take(ws.nn, 2 * B); c.take(ws.s_ctrl, 2 * U * S); c.take(ws.s_steps, 2 * S);
            c.take(ws.s_fin, D * S); c.take(ws.s_valid, S); c.take(ws.sel, B); c.take(ws.counters, 8); c.take(ws.goal_key, 1);
            c.take(ws.path_nodes, PC); c.take(ws.path_ctrl, U * PC); c.take(ws.path_steps, PC); c.take(ws.path_parent, PC);
            c.take(ws.path_valid, PC); c.take(ws.path_seg, D * (size_t)KCBS_MAX_STEPS * PC); c.take(ws.nn_r, M); c.take(ws.nn_t, M);
            c.take(ws.plan, 1); c.take(ws.cons_states, 3 * (size_t)cons_cap); c.take(ws.cons, (size_t)cons_desc_cap);
#ifdef KCBS_RRT_TRACE
            c.take(ws.trace, (size_t)kTraceRounds * 8);
#endif
        };
        auto host_layout = [&](Carver &c) {
            c.take(ws.h_plan, 1); c.take(ws.h_progress, kRrtDepth); c.take(ws.h_counters, 16); c.take(ws.h_path_valid, PC);
            c.take(ws.h_path_seg, seg_floats);
        };
        Carver dev_size(nullptr), host_size(nullptr);
        device_layout(dev_size);
        host_layout(host_size);
        KCBS_CUDA(ctx, cudaMalloc(&ws.slab, dev_size.off));
        KCBS_CUDA(ctx, cudaHostAlloc(&ws.h_slab, host_size.off, cudaHostAllocMapped));
        Carver dev(ws.slab), host(ws.h_slab);
        device_layout(dev);
        host_layout(host);
        ws.h_path_cap = seg_floats;
        ws.cons_cap = cons_cap;
        ws.cons_desc_cap = cons_desc_cap;
        char *h_slab_dev = nullptr;   // the pinned slab as the device sees it (the progress ring is written by k_rrt_append)
        KCBS_CUDA(ctx, cudaHostGetDevicePointer((void **)&h_slab_dev, ws.h_slab, 0));
        ws.d_progress = reinterpret_cast<RrtProgress *>(h_slab_dev + ((char *)ws.h_progress - (char *)ws.h_slab));
        KCBS_CUDA(ctx, cudaStreamCreateWithFlags(&ws.side, cudaStreamNonBlocking));
        KCBS_CUDA(ctx, cudaEventCreateWithFlags(&ws.fork, cudaEventDisableTiming));
        KCBS_CUDA(ctx, cudaEventCreateWithFlags(&ws.join, cudaEventDisableTiming));
        for (cudaEvent_t &e : ws.ev) KCBS_CUDA(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        ws.max_nodes = p.max_nodes;
        ws.n_samples = n_samples;
        ws.n_targets = p.n_targets;
        ws.dim = dim;
    }
    if (cons_total > ws.cons_cap) {
        KCBS_CUDA(ctx, sync_lane(ctx, ws));
        std::lock_guard<std::mutex> gate(capture_gate());
        drop_loops(ws);
        if (ws.own_cons_states && ws.cons_states) cudaFree(ws.cons_states);
        ws.cons_states = nullptr;
        ws.own_cons_states = true;
        ws.cons_cap = cons_total + cons_total / 2 + 1024;
        int st = dev_alloc(ctx, ws.cons_states, 3 * (size_t)ws.cons_cap);
        if (st) return st;
    }
    if (n_cons > ws.cons_desc_cap) {
        KCBS_CUDA(ctx, sync_lane(ctx, ws));
        std::lock_guard<std::mutex> gate(capture_gate());
        drop_loops(ws);
        if (ws.own_cons && ws.cons) cudaFree(ws.cons);
        ws.cons = nullptr;
        ws.own_cons = true;
        ws.cons_desc_cap = n_cons + 64;
        int st = dev_alloc(ctx, ws.cons, (size_t)ws.cons_desc_cap);
        if (st) return st;
    }
    if (seg_floats > ws.h_path_cap) {   // (the loops copy the re-expanded path into this buffer)
        KCBS_CUDA(ctx, sync_lane(ctx, ws));
        std::lock_guard<std::mutex> gate(capture_gate());
        drop_loops(ws);
        if (ws.own_h_path_seg && ws.h_path_seg) cudaFreeHost(ws.h_path_seg);
        ws.h_path_seg = nullptr;
        ws.own_h_path_seg = true;
        ws.h_path_cap = 0;
        KCBS_CUDA(ctx, cudaHostAlloc((void **)&ws.h_path_seg, sizeof(float) * seg_floats, cudaHostAllocDefault));
        ws.h_path_cap = seg_floats;
    }
    return KCBS_OK;
}

// Rounds driven by the device (one graph launch per plan) unless KCBS_RRT_ROUNDS=stream asks for host-queued rounds.
bool device_rounds()
{
    static const bool on = [] {
        const char *e = std::getenv("KCBS_RRT_ROUNDS");
        return !(e && std::strcmp(e, "stream") == 0);
    }();
    return on;
}

#ifndef KCBS_LOCAL_REPAIRS
#define KCBS_LOCAL_REPAIRS 1
#endif
constexpr int kLocalRepairs = KCBS_LOCAL_REPAIRS;   // conflicts of one pair that are answered by re-plans from the split

// Steps of the parent's path a child's re-plan gives up before the newest constraint begins (KCBS_REPLAN_BACK; 0 = every
// re-plan starts from the robot's start state).
int replan_back()
{
    static const int v = [] {
        const char *e = std::getenv("KCBS_REPLAN_BACK");
        return e ? std::max(0, std::atoi(e)) : 40;
    }();
    return v;
}

bool trace_solve()
{
    static const bool on = std::getenv("KCBS_TRACE_SOLVE") != nullptr;   // per-expansion and per-plan timing on stderr
    return on;
}

double now_s()
{
    using namespace std::chrono;
    return duration<double>(steady_clock::now().time_since_epoch()).count();
}

int check_rrt_params(kcbs_ctx *ctx, const kcbs_rrt_params &p)
{
    if (p.n_targets < 1 || p.n_controls < 1 || (p.max_nodes != 0 && p.max_nodes < 2) || p.max_iterations < 1 || p.min_steps < 1 ||
        p.max_steps < p.min_steps || p.max_steps > KCBS_MAX_STEPS || !(p.goal_radius > 0) ||
        (long long)p.n_targets * p.n_controls > (1ll << 26))
        return fail(ctx, KCBS_ERR_INVALID_ARGUMENT, "kcbs_rrt_params out of range");
    return KCBS_OK;
}

}  // namespace

// The planner's kernels are launched from inside stream captures: they are loaded when the context is created, outside
// any capture (a kernel's first use loads it, which a capture in progress does not always survive).
cudaError_t init_planner_kernels()
{
    const cudaError_t e = load_planner_kernels<1>();
    return e != cudaSuccess ? e : load_planner_kernels<2>();
}

void destroy_rrt_workspace(kcbs_ctx *ctx)
{
    for (kcbs_ctx *w : ctx->workers) {  // planner workers share the scene buffers and own only what they created
        if (w->stream) cudaStreamSynchronize(w->stream);
        destroy_rrt_workspace(w);
        if (w->stream) cudaStreamDestroy(w->stream);
        delete w;
    }
    ctx->workers.clear();
    if (ctx->rrt) {
        free_workspace(*ctx->rrt);
        delete ctx->rrt;
        ctx->rrt = nullptr;
    }
}

// The low-level planner proper for NC cars planned as one system (NC = 1: a robot; NC = 2: a merged pair, the 10-d
// system of SystemMergerDatabase.h:108-156).  Host orchestration only; every arithmetic step is a kernel.  By default the
// whole search is one graph launch (the rounds are driven by the device, see the top of this file).  With host-queued
// rounds the host queues iterations without waiting for them: each round reports (goal, node count) into a host-mapped
// ring, the host looks at round i - kRrtDepth before it queues round i, and rounds queued after the goal was found cancel
// themselves on the device.  start = 5 NC reals, goal_xy = 2 NC, traj_out = [5 NC][max_T].
// Tree capacity when the caller leaves it to the planner (max_nodes = 0): the rounds a path across the workspace needs at
// full speed (one round extends the tree by at most max_steps * step seconds of driving) with 50 % slack, between 32 and
// 128 rounds.  A re-plan that cannot succeed -- K-CBS produces them: a constraint can park another robot on the goal --
// costs the whole capacity (the nearest-neighbour scan grows with the tree), so the capacity is what bounds a failure.
static int auto_max_nodes(const kcbs_ctx *ctx, const kcbs_rrt_params &p)
{
    const DevWorld &w = ctx->world;
    const float dx = w.hi_t[0] - w.lo_t[0], dy = w.hi_t[1] - w.lo_t[1];
    const float v_max = std::max(std::fabs(w.lo_t[2]), std::fabs(w.hi_t[2]));
    const float per_round = std::max(1e-3f, v_max * w.step * (float)p.max_steps);
    const int rounds = std::min(128, std::max(32, (int)std::ceil(1.5f * std::sqrt(dx * dx + dy * dy) / per_round)));
    return (int)std::min<long long>((long long)rounds * p.n_targets + 1, 1ll << 22);
}

template <int NC>
static int rrt_plan(kcbs_ctx *ctx, const int *robots, const float *start, const float *goal_xy, const kcbs_rrt_params &p_in,
                    int n_cons, const int32_t *cons_robot, const int32_t *cons_k0, const int32_t *cons_len,
                    const int32_t *cons_offset, const float *cons_states, int max_T, float *traj_out, int32_t *T_out,
                    kcbs_rrt_stats *stats)
{
    constexpr int D = 5 * NC;
    const double t_begin = now_s();
    int st = check_rrt_params(ctx, p_in);
    if (st) return st;
    kcbs_rrt_params p = p_in;
    if (p.max_nodes == 0) p.max_nodes = auto_max_nodes(ctx, p);
    for (int c = 0; c < NC; ++c)
        if (robots[c] < 0 || robots[c] >= ctx->n_robots) return fail(ctx, KCBS_ERR_UNKNOWN_ROBOT, "robot index out of range");
    if (!start || !goal_xy || !traj_out || !T_out || max_T < 1) return fail(ctx, KCBS_ERR_INVALID_ARGUMENT, "null argument");
    int cons_total = 0;
    for (int c = 0; c < n_cons; ++c) {
        if (cons_robot[c] < 0 || cons_robot[c] >= ctx->n_robots || cons_offset[c] < 0)
            return fail(ctx, KCBS_ERR_INVALID_ARGUMENT, "bad constraint");
        cons_total = std::max(cons_total, cons_offset[c] + std::abs(cons_len[c]));
    }
    KCBS_CUDA(ctx, cudaSetDevice(ctx->device));
    const int e_max = std::min(p.max_iterations, 8192);   // edge slots of the path extraction (= RrtWorkspace::path_cap at most)
    const size_t seg_floats = (size_t)D * p.max_steps * e_max;
    if ((st = ensure_workspace(ctx, p, NC, cons_total, n_cons, seg_floats))) return st;
    RrtWorkspace &ws = *ctx->rrt;
    cudaStream_t s = ctx->stream;

    std::vector<ConsDesc> desc(n_cons);
    if (n_cons > 0) {
        for (int c = 0; c < n_cons; ++c) {
            // circle around the constraint's positions (centre of their bounding box): the expansion kernel skips the
            // constraint for samples that cannot come near it (cons_mask, kcbs_device.cuh)
            const int len = std::abs(cons_len[c]);
            const float *xs = cons_states + cons_offset[c], *ys = xs + cons_total;
            float x_lo = xs[0], x_hi = xs[0], y_lo = ys[0], y_hi = ys[0];
            for (int j = 1; j < len; ++j) {
                x_lo = std::min(x_lo, xs[j]); x_hi = std::max(x_hi, xs[j]);
                y_lo = std::min(y_lo, ys[j]); y_hi = std::max(y_hi, ys[j]);
            }
            const float cx = 0.5f * (x_lo + x_hi), cy = 0.5f * (y_lo + y_hi);
            float r2 = 0.0f;
            for (int j = 0; j < len; ++j) r2 = std::max(r2, (xs[j] - cx) * (xs[j] - cx) + (ys[j] - cy) * (ys[j] - cy));
            desc[c] = ConsDesc{cons_robot[c], cons_k0[c], cons_len[c], cons_offset[c], cx, cy, std::sqrt(r2) * 1.0001f + 1e-4f, 0};
        }
        KCBS_CUDA(ctx, cudaMemcpyAsync(ws.cons, desc.data(), sizeof(ConsDesc) * n_cons, cudaMemcpyHostToDevice, s));
        // host planes have stride cons_total; keep the same stride on the device
        KCBS_CUDA(ctx, cudaMemcpyAsync(ws.cons_states, cons_states, sizeof(float) * 3 * (size_t)cons_total, cudaMemcpyHostToDevice, s));
    }

    // what changes from plan to plan goes to the device once (RrtPlan, kernels.h); the launches below are the same for
    // every round of every plan
    RrtPlan &hp = *ws.h_plan;
    std::memset(&hp, 0, sizeof(hp));
    hp.robot[0] = robots[0]; hp.robot[1] = robots[NC - 1]; hp.n_cons = n_cons; hp.cons_total = cons_total; hp.max_T = max_T;
    hp.max_iterations = p.max_iterations; hp.node_cap = p.max_nodes;
    for (int c = 0; c < 2 * NC; ++c) hp.goal[c] = goal_xy[c];
    hp.goal_radius = NC == 1 ? p.goal_radius : KCBS_PAIR_GOAL_RADIUS;
    hp.goal_bias = p.goal_bias;
    hp.seed = p.seed;
    for (int c = 0; c < D; ++c) hp.start[c] = start[c];
    hp.limit_ns = p.time_limit_s > 0 ? (unsigned long long)(p.time_limit_s * 1e9) : 0ull;
    KCBS_CUDA(ctx, cudaMemcpyAsync(ws.plan, ws.h_plan, sizeof(RrtPlan), cudaMemcpyHostToDevice, s));

    RrtArgs a = {};
    a.ws = ws;
    // sampling box = the state bounds the validity test uses (thresholds are within one ulp of them)
    for (int c = 0; c < 4; ++c) { a.lo[c] = ctx->world.lo_t[c]; a.hi[c] = ctx->world.hi_t[c]; }
    a.n_targets = p.n_targets; a.n_controls = p.n_controls; a.min_steps = p.min_steps; a.max_steps = p.max_steps;
    const int n = p.n_targets * p.n_controls;

    // the round's expansion: parents = the low words of the nearest-neighbour keys (little endian), controls and durations
    // (drawn one round ahead) = the buffers of the round's parity; robot(s) and constraint counts come from the plan block
    // (n_cons here only picks the kernel variant)
    PropArgs pa = {};
    PairArgs qa = {};
    if (NC == 1) {
        pa.rounds = ctx->prop_rounds;
        pa.n = n; pa.n_parents = ws.max_nodes; pa.parents = ws.nodes; pa.parent_idx = reinterpret_cast<const int *>(ws.nn); pa.controls = ws.s_ctrl;
        pa.parent_group = p.n_controls; pa.parent_stride = 2;
        pa.steps = ws.s_steps; pa.seg = nullptr; pa.fin = ws.s_fin; pa.valid_steps = ws.s_valid; pa.max_steps = p.max_steps;
        pa.robot = robots[0]; pa.parent_time = ws.time; pa.cons = ws.cons; pa.cons_states = ws.cons_states; pa.n_cons = n_cons;
        pa.cons_total = cons_total;
        pa.plan = ws.plan; pa.parity_ctrl = (long long)2 * NC * n; pa.parity_steps = n; pa.parity_parent = 2 * (long long)p.n_targets;   // (in ints: two per key)
    } else {
        qa.n = n; qa.n_parents = ws.max_nodes; qa.parents = ws.nodes; qa.parent_idx = reinterpret_cast<const int *>(ws.nn); qa.controls = ws.s_ctrl;
        qa.parent_group = p.n_controls; qa.parent_stride = 2;
        qa.steps = ws.s_steps; qa.seg = nullptr; qa.fin = ws.s_fin; qa.valid_steps = ws.s_valid; qa.max_steps = p.max_steps;
        qa.robot1 = robots[0]; qa.robot2 = robots[NC - 1];
        for (int c = 0; c < 4; ++c) qa.goals[c] = hp.goal[c];
        qa.hold_radius = KCBS_PAIR_GOAL_RADIUS;   // StatePropagatorDatabase.h:170
        qa.parent_time = ws.time; qa.cons = ws.cons; qa.cons_states = ws.cons_states; qa.n_cons = n_cons; qa.cons_total = cons_total;
        qa.plan = ws.plan; qa.parity_ctrl = (long long)2 * NC * n; qa.parity_steps = n; qa.parity_parent = 2 * (long long)p.n_targets;
    }
    const int tiles = (p.n_targets + kNNTile - 1) / kNNTile;
    const dim3 nn_grid((unsigned)tiles, (unsigned)nn_grid_y(ws.max_nodes, tiles));
    const dim3 nn_fresh((unsigned)tiles, (unsigned)nn_grid_y(p.n_targets, tiles));   // a round appends n_targets nodes at most
    auto launch_init = [&](const RrtArgs &ra) {
        k_rrt_init<NC><<<(std::max(n, 2 * p.n_targets) + 255) / 256, 256, 0, s>>>(ra);
        return cudaGetLastError();
    };
    // One round.  The scan of the old nodes for the NEXT round depends on nothing this round does: it is a branch of its
    // own (the side stream; in the graph, a parallel branch of the loop body), so the round's critical path is the scan of
    // the few nodes the previous round appended, the expansion, the selection and the append.
    auto launch_round = [&](const RrtArgs &ra) {
        cudaError_t e;
        if ((e = cudaEventRecord(ws.fork, s)) != cudaSuccess || (e = cudaStreamWaitEvent(ws.side, ws.fork, 0)) != cudaSuccess) return e;
        k_rrt_nearest<NC><<<nn_grid, kNNThreads, 0, ws.side>>>(ra, 0);
        if ((e = cudaEventRecord(ws.join, ws.side)) != cudaSuccess) return e;
        k_rrt_nearest<NC><<<nn_fresh, kNNThreads, 0, s>>>(ra, 1);
        e = NC == 1 ? launch_propagate(ctx->world, pa, s) : launch_propagate_pair(ctx->world, qa, s);
        if (e != cudaSuccess) return e;
        k_rrt_select<NC><<<(p.n_targets * 32 + 255) / 256, 256, 0, s>>>(ra);
        k_rrt_append<NC><<<(p.n_targets + kAppendThreads - 1) / kAppendThreads, kAppendThreads, 0, s>>>(ctx->world, ra);
        if ((e = cudaStreamWaitEvent(s, ws.join, 0)) != cudaSuccess) return e;
        return cudaGetLastError();
    };
    // What follows the last round, whatever its outcome: the path's edges (none when there is no goal node), the states
    // along them re-expanded by the propagation kernel (same arithmetic as the search), and everything the host wants to
    // know in pinned memory.  Same launches for every plan: e_max edge slots, the unused ones of duration 0.
    auto launch_tail = [&](const RrtArgs &ra) {
        cudaError_t e;
        k_rrt_path<NC><<<1, 32, 0, s>>>(ra, e_max);
        if (NC == 1) {
            PropArgs pe = pa;   // (plan block: robot and constraint counts; no parity: the path has its own buffers)
            pe.n = e_max; pe.parent_idx = ws.path_parent; pe.parent_group = pe.parent_stride = 1; pe.controls = ws.path_ctrl; pe.steps = ws.path_steps;
            pe.parity_ctrl = pe.parity_steps = pe.parity_parent = 0; pe.rounds = 1;
            pe.seg = ws.path_seg; pe.fin = nullptr; pe.valid_steps = ws.path_valid; pe.max_steps = p.max_steps;
            e = launch_propagate(ctx->world, pe, s);
        } else {
            PairArgs qe = qa;
            qe.n = e_max; qe.parent_idx = ws.path_parent; qe.parent_group = qe.parent_stride = 1; qe.controls = ws.path_ctrl; qe.steps = ws.path_steps;
            qe.parity_ctrl = qe.parity_steps = qe.parity_parent = 0;
            qe.seg = ws.path_seg; qe.fin = nullptr; qe.valid_steps = ws.path_valid; qe.max_steps = p.max_steps;
            e = launch_propagate_pair(ctx->world, qe, s);
        }
        if (e != cudaSuccess) return e;
        if ((e = cudaMemcpyAsync(ws.h_counters, ws.counters, 8 * sizeof(int), cudaMemcpyDeviceToHost, s)) != cudaSuccess) return e;
        if ((e = cudaMemcpyAsync(ws.h_counters + 8, ws.goal_key, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s)) != cudaSuccess) return e;
        if ((e = cudaMemcpyAsync(ws.h_path_valid, ws.path_valid, sizeof(int) * e_max, cudaMemcpyDeviceToHost, s)) != cudaSuccess) return e;
        return cudaMemcpyAsync(ws.h_path_seg, ws.path_seg, sizeof(float) * seg_floats, cudaMemcpyDeviceToHost, s);
    };

    int n_nodes = 1, iters = 0, queued = 0;
    unsigned long long goal = ~0ull;
    // the start itself may already satisfy the goal
    {
        bool in = n_cons == 0;
        for (int c = 0; c < NC; ++c) {
            const float gx = start[5 * c] - goal_xy[2 * c], gy = start[5 * c + 1] - goal_xy[2 * c + 1];
            const float d = std::sqrt(gx * gx + gy * gy);
            in = in && (NC == 1 ? d <= hp.goal_radius : d < hp.goal_radius);
        }
        if (in) goal = 0;
    }
#ifdef KCBS_RRT_TRACE
    {
        std::vector<unsigned long long> z((size_t)kTraceRounds * 8);
        for (size_t q = 0; q < z.size(); ++q) z[q] = (q & 1) ? 0ull : ~0ull;
        KCBS_CUDA(ctx, cudaMemcpyAsync(ws.trace, z.data(), z.size() * 8, cudaMemcpyHostToDevice, s));
        KCBS_CUDA(ctx, cudaStreamSynchronize(s));
    }
#endif
    // ---- the whole search as ONE graph launch: k_rrt_init, then a WHILE node whose body is the kernels of a round and
    // whose condition k_rrt_append sets on the device (goal found, tree full, iteration or time limit), then the path
    // extraction.  Built once per launch shape; a shape whose graph cannot be built keeps host-queued rounds. ----
    cudaGraphExec_t loop_exec = nullptr;
    if (goal == ~0ull && device_rounds()) {
        RrtWorkspace::Loop &L = ctx->rrt->loop[NC - 1][n_cons > 0][NC == 1 ? std::min(std::max(pa.rounds, 0), 2) : 0];
        const int key[6] = {p.n_targets, p.n_controls, p.min_steps, p.max_steps, NC == 1 ? pa.rounds : 0, e_max};
        if (std::memcmp(L.key, key, sizeof(key)) != 0 || (!L.exec && !L.failed)) {
            const double t_build = now_s();
            // one capture at a time, process wide, and no device-wide synchronisation meanwhile (capture_gate, ctx.h)
            std::lock_guard<std::mutex> build_lock(capture_gate());
            if (L.exec) cudaGraphExecDestroy(L.exec);
            L = RrtWorkspace::Loop();
            std::memcpy(L.key, key, sizeof(key));
            cudaGraph_t g = nullptr;
            KCBS_CUDA(ctx, cudaGraphCreate(&g, 0));
            RrtArgs la = a;
            la.in_loop = 1;
            cudaGraphNode_t init_node = nullptr, while_node = nullptr;
            size_t one = 1;
            cudaGraphNodeParams wp = {};
            const char *step = "cudaGraphConditionalHandleCreate";
            bool capture_broken = false;
            cudaError_t e = cudaGraphConditionalHandleCreate(&la.loop, g, 1, cudaGraphCondAssignDefault);
            // captures the launches of `what` into graph `into`, after node `dep` (if any)
            auto capture = [&](cudaGraph_t into, cudaGraphNode_t *dep, const char *name, auto what) {
                if (e != cudaSuccess) return;
                step = name;
                e = cudaStreamBeginCaptureToGraph(s, into, dep, nullptr, dep ? 1 : 0, cudaStreamCaptureModeThreadLocal);
                if (e != cudaSuccess) return;
                const cudaError_t e1 = what(la);
                cudaGraph_t same = nullptr;
                e = cudaStreamEndCapture(s, &same);
                if (e1 != cudaSuccess) e = e1;
                capture_broken = e != cudaSuccess;
            };
            capture(g, nullptr, "capture of k_rrt_init", launch_init);
            if (e == cudaSuccess) { step = "cudaGraphGetNodes"; e = cudaGraphGetNodes(g, &init_node, &one); }
            if (e == cudaSuccess) {
                step = "cudaGraphAddNode (while)";
                wp.type = cudaGraphNodeTypeConditional;
                wp.conditional.handle = la.loop;
                wp.conditional.type = cudaGraphCondTypeWhile;
                wp.conditional.size = 1;
                e = cudaGraphAddNode(&while_node, g, &init_node, 1, &wp);
            }
            if (e == cudaSuccess) capture(wp.conditional.phGraph_out[0], nullptr, "capture of a round", launch_round);
            capture(g, &while_node, "capture of the path extraction", launch_tail);
            if (e == cudaSuccess) { step = "cudaGraphInstantiate"; e = cudaGraphInstantiate(&L.exec, g, 0); }
            // (a graph whose capture was invalidated is abandoned, not destroyed: the driver has been seen to fault on it)
            if (!capture_broken) cudaGraphDestroy(g);
            if (e != cudaSuccess) {
                L.exec = nullptr;
                L.failed = true;
                cudaGetLastError();
                fprintf(stderr, "[kcbs] round loop graph not built (%s: %s): rounds of this shape are queued by the host\n", step,
                        cudaGetErrorString(e));
            } else if (trace_solve()) {
                fprintf(stderr, "[kcbs]   round loop graph built in %.0f us\n", (now_s() - t_build) * 1e6);
            }
        }
        loop_exec = L.exec;
    }
    if (goal != ~0ull) {
        KCBS_CUDA(ctx, cudaStreamSynchronize(s));   // (the uploads above)
    } else if (loop_exec) {
        KCBS_CUDA(ctx, cudaGraphLaunch(loop_exec, s));
    } else {
        // ---- rounds queued by the host (KCBS_RRT_ROUNDS=stream), kRrtDepth of them ahead of the device ----
        for (int q = 0; q < kRrtDepth; ++q) ws.h_progress[q] = RrtProgress{~0ull, 1, 0};
        KCBS_CUDA(ctx, launch_init(a));
        while (queued < p.max_iterations) {
            if (queued >= kRrtDepth) {   // outcome of the round that used this ring slot last
                KCBS_CUDA(ctx, cudaEventSynchronize(ws.ev[queued % kRrtDepth]));
                const RrtProgress pr = ws.h_progress[queued % kRrtDepth];
                if (pr.goal != ~0ull || pr.nodes >= p.max_nodes) break;
            }
            if (p.time_limit_s > 0 && now_s() - t_begin > p.time_limit_s) break;
            KCBS_CUDA(ctx, launch_round(a));
            KCBS_CUDA(ctx, cudaEventRecord(ws.ev[queued % kRrtDepth], s));
            ++queued;
        }
        KCBS_CUDA(ctx, launch_tail(a));
    }
    int T = 0;
    if (goal == ~0ull) {
        KCBS_CUDA(ctx, cudaStreamSynchronize(s));
        std::memcpy(&goal, ws.h_counters + 8, sizeof(goal));
        iters = ws.h_counters[2];
        n_nodes = std::min(ws.h_counters[iters & 1], p.max_nodes);   // the entry the next round would have read
        ctx->launches += 1 + 5 * (loop_exec ? iters : queued) + 2;
        if (goal != ~0ull) {
            const int E = ws.h_counters[4];
            for (int c = 0; c < D; ++c) traj_out[(size_t)c * max_T] = start[c];
            T = 1;
            for (int e = 0; e < E; ++e)
                for (int j = 0; j < ws.h_path_valid[e] && T < max_T; ++j, ++T)
                    for (int c = 0; c < D; ++c) traj_out[(size_t)c * max_T + T] = ws.h_path_seg[((size_t)c * p.max_steps + j) * e_max + e];
        }
    } else {   // (the start is in the goal region)
        for (int c = 0; c < D; ++c) traj_out[(size_t)c * max_T] = start[c];
        T = 1;
    }
    *T_out = T;
#ifdef KCBS_RRT_TRACE
    if (trace_solve()) {
        std::vector<unsigned long long> z((size_t)kTraceRounds * 8);
        cudaMemcpy(z.data(), ws.trace, z.size() * 8, cudaMemcpyDeviceToHost);
        const unsigned long long t0 = z[2];   // start of round 0's fresh scan
        for (int r = 0; r < std::min(iters, kTraceRounds); ++r) {
            const unsigned long long *w = &z[(size_t)r * 8];
            auto us = [&](unsigned long long t) { return (double)((long long)(t - t0)) * 1e-3; };
            fprintf(stderr, "[kcbs]     round %2d: side scan %7.1f..%7.1f | fresh scan %7.1f..%7.1f | select %7.1f..%7.1f | append %7.1f..%7.1f us\n", r,
                    us(w[0]), us(w[1]), us(w[2]), us(w[3]), us(w[4]), us(w[5]), us(w[6]), us(w[7]));
        }
    }
#endif
    if (trace_solve())
        fprintf(stderr, "[kcbs]   plan robot %d: %.0f us, %d rounds run, %d queued, %d constraints, %d nodes, path %d\n", robots[0],
                (now_s() - t_begin) * 1e6, iters, queued, n_cons, n_nodes, T);
    if (stats) {
        stats->solved = goal != ~0ull;
        stats->iterations = iters;
        stats->nodes = n_nodes;
        stats->path_steps = T > 0 ? T - 1 : 0;
        stats->seconds = now_s() - t_begin;
    }
    return KCBS_OK;
}

}  // namespace kcbs

using namespace kcbs;

extern "C" {

void kcbs_rrt_defaults(kcbs_rrt_params *p)
{
    if (!p) return;
    p->n_targets = 1024;   // measured over the demo scenes: 2048 costs twice the work per round for no fewer rounds
    p->n_controls = 32;
    p->max_nodes = 0;       // chosen from the workspace, see auto_max_nodes
    p->max_iterations = 512;
    p->min_steps = 1;    // demo_*.cpp:118
    p->max_steps = 10;
    p->goal_bias = 0.05f;
    p->goal_radius = 0.5f;  // GoalRegionDatabase.h:48
    p->time_limit_s = 5.0;  // Benchmark.h:153
    p->seed = 1;
}

void kcbs_kcbs_defaults(kcbs_kcbs_params *p)
{
    if (!p) return;
    kcbs_rrt_defaults(&p->low_level);
    p->time_limit_s = 180.0;  // demo_*.cpp:168
    p->max_nodes = 0;
    p->merge_bound = 2147483647;  // KCBS default: never merge (the demos set 10 / 20)
}

int kcbs_rrt_plan(kcbs_ctx *ctx, int robot, const float *start, const float *goal_xy, const kcbs_rrt_params *params,
                  int n_cons, const int32_t *cons_robot, const int32_t *cons_k0, const int32_t *cons_len,
                  const int32_t *cons_offset, const float *cons_states, int max_T, float *traj_out, int32_t *T_out,
                  kcbs_rrt_stats *stats)
{
    if (!ctx) return KCBS_ERR_INVALID_ARGUMENT;
    kcbs_rrt_params p;
    if (params)
        p = *params;
    else
        kcbs_rrt_defaults(&p);
    if (n_cons < 0 || (n_cons > 0 && (!cons_robot || !cons_k0 || !cons_len || !cons_offset || !cons_states)))
        return fail(ctx, KCBS_ERR_INVALID_ARGUMENT, "constraint arrays are null");
    return rrt_plan<1>(ctx, &robot, start, goal_xy, p, n_cons, cons_robot, cons_k0, cons_len, cons_offset, cons_states, max_T,
                       traj_out, T_out, stats);
}

int kcbs_rrt_plan_pair(kcbs_ctx *ctx, int robot1, int robot2, const float *start, const float *goals_xy, const kcbs_rrt_params *params,
                       int n_cons, const int32_t *cons_robot, const int32_t *cons_k0, const int32_t *cons_len,
                       const int32_t *cons_offset, const float *cons_states, int max_T, float *traj_out, int32_t *T_out,
                       kcbs_rrt_stats *stats)
{
    if (!ctx) return KCBS_ERR_INVALID_ARGUMENT;
    kcbs_rrt_params p;
    if (params)
        p = *params;
    else
        kcbs_rrt_defaults(&p);
    if (n_cons < 0 || (n_cons > 0 && (!cons_robot || !cons_k0 || !cons_len || !cons_offset || !cons_states)))
        return fail(ctx, KCBS_ERR_INVALID_ARGUMENT, "constraint arrays are null");
    const int robots[2] = {robot1, robot2};
    return rrt_plan<2>(ctx, robots, start, goals_xy, p, n_cons, cons_robot, cons_k0, cons_len, cons_offset, cons_states, max_T,
                       traj_out, T_out, stats);
}

// ---- K-CBS high level ------------------------------------------------------------------------------------

namespace {

struct Constraint {
    int other, k0;
    std::vector<float> x, y, th;  // states of `other` at k0 .. k0 + len - 1
    bool persistent = false;      // the last state holds for every later time index (a robot parked at its goal)
};

struct CbsNode {
    std::vector<std::vector<float>> traj;        // per robot [5][T_r] planes, stride T_r
    std::vector<int> len;                        // T_r
    std::vector<std::vector<std::shared_ptr<Constraint>>> cons;  // per robot
    long long cost = 0;                          // sum of path lengths (propagation steps)
    long long id = 0;
};

struct NodeOrder {
    bool operator()(const std::shared_ptr<CbsNode> &a, const std::shared_ptr<CbsNode> &b) const
    {
        return a->cost != b->cost ? a->cost > b->cost : a->id > b->id;
    }
};

extern "C++" {
// Context the i-th concurrent planner job runs on: the caller's own context for i == 0, else a worker clone.
kcbs_ctx *planner_lane(kcbs_ctx *ctx, int i)
{
    if (i == 0) return ctx;
    while ((int)ctx->workers.size() < i) {
        kcbs_ctx *w = new kcbs_ctx();
        w->device = ctx->device;
        w->world = ctx->world;  // device pointers of the scene are shared, read only
        w->n_robots = ctx->n_robots;
        if (cudaSetDevice(ctx->device) != cudaSuccess || cudaStreamCreateWithFlags(&w->stream, cudaStreamNonBlocking) != cudaSuccess) {
            delete w;
            return nullptr;
        }
        ctx->workers.push_back(w);
    }
    return ctx->workers[i - 1];
}

// Runs job(lane context, j) for j in [0, n_jobs) on up to max_lanes host threads, each driving its own stream and
// planner workspace, so that independent low-level plans overlap on the GPU (a plan is a chain of small dependent
// kernels that leaves most of the GPU idle).  Returns the first non-OK status.
template <class Job>
int run_planner_jobs(kcbs_ctx *ctx, int n_jobs, int max_lanes, Job job)
{
    const int lanes = std::max(1, std::min(n_jobs, max_lanes));
    std::vector<kcbs_ctx *> lane_ctx(lanes);
    for (int i = 0; i < lanes; ++i)
        if (!(lane_ctx[i] = planner_lane(ctx, i))) return fail(ctx, KCBS_ERR_CUDA, "planner worker creation failed");
    // launch shape of the lanes' expansions (kcbs_set_expansion_mode of the caller's context wins): a handful of lanes
    // cannot fill the GPU, so each expansion takes the latency shape; many concurrent lanes take the throughput shape
    const int user_mode = ctx->user_prop_rounds;
    for (int i = 0; i < lanes; ++i) lane_ctx[i]->prop_rounds = user_mode ? user_mode : (lanes > 4 ? 2 : 1);
    std::atomic<int> next(0);
    std::vector<int> status(lanes, KCBS_OK);
    auto body = [&](int lane) {
        cudaSetDevice(ctx->device);
        for (int j = next.fetch_add(1); j < n_jobs; j = next.fetch_add(1)) {
            const int st = job(lane_ctx[lane], j);
            if (st != KCBS_OK && status[lane] == KCBS_OK) status[lane] = st;
        }
    };
    std::vector<std::thread> threads;
    for (int i = 1; i < lanes; ++i) threads.emplace_back(body, i);
    body(0);
    for (auto &t : threads) t.join();
    ctx->prop_rounds = user_mode;
    for (int i = 0; i < lanes; ++i) {
        if (i > 0) {
            ctx->launches += lane_ctx[i]->launches;
            lane_ctx[i]->launches = 0;
        }
        if (status[i] != KCBS_OK) {
            if (i > 0) ctx->err = lane_ctx[i]->err;
            return status[i];
        }
    }
    return KCBS_OK;
}

#ifndef KCBS_PLANNER_LANES
#define KCBS_PLANNER_LANES 16
#endif
constexpr int kPlannerLanes = KCBS_PLANNER_LANES;
}  // extern "C++"

// flat constraint arrays of the C ABI from the constraint objects of a tree node
// (shift > 0: the plan starts at absolute time index `shift`; constraint times are taken relative to it, what lies before
// it is cut off, and of a persistent constraint that ended earlier the parked state remains)
struct FlatConstraints {
    std::vector<int32_t> robot, k0, ln, off;
    std::vector<float> states;
    explicit FlatConstraints(const std::vector<std::shared_ptr<Constraint>> &cons, int shift = 0)
    {
        std::vector<int> first;   // first state of every kept constraint
        std::vector<const Constraint *> kept;
        int total = 0;
        for (const auto &c : cons) {
            const int n = (int)c->x.size(), ks = c->k0 - shift;
            int f = std::max(0, -ks);
            if (f >= n) {
                if (!c->persistent) continue;   // over before the plan starts
                f = n - 1;
            }
            kept.push_back(c.get()); first.push_back(f);
            robot.push_back(c->other); k0.push_back(std::max(ks, 0)); off.push_back(total);
            ln.push_back(c->persistent ? -(n - f) : n - f);
            total += n - f;
        }
        states.assign((size_t)3 * std::max(total, 1), 0.0f);
        for (size_t q = 0; q < kept.size(); ++q)
            for (size_t j = (size_t)first[q]; j < kept[q]->x.size(); ++j) {
                const size_t at = (size_t)off[q] + j - (size_t)first[q];
                states[at] = kept[q]->x[j];
                states[(size_t)total + at] = kept[q]->y[j];
                states[(size_t)2 * total + at] = kept[q]->th[j];
            }
    }
};

// Low-level plan of robot r under `cons`.  k_split = 0: from the robot's start state.  k_split > 0 (a re-plan of a child
// node): the path `traj` of the parent node is kept up to time index k_split -- it satisfies the older constraints, and the
// new one begins later -- and only the rest is planned, from the state the old path has there; the tree then needs the
// rounds of the remaining way, not of the whole way.
int plan_robot(kcbs_ctx *ctx, int r, const float *starts, const float *goals, int R, const kcbs_rrt_params &p,
               const std::vector<std::shared_ptr<Constraint>> &cons, int max_T, std::vector<float> &traj, int &len,
               kcbs_rrt_stats &rs, int k_split = 0)
{
    float start[5];
    const int L0 = len;
    if (k_split <= 0 || k_split >= L0 || k_split >= max_T - 1) k_split = 0;
    for (int c = 0; c < 5; ++c) start[c] = k_split > 0 ? traj[(size_t)c * L0 + k_split] : starts[(size_t)c * R + r];
    FlatConstraints fc(cons, k_split);
    const int sub_T = max_T - k_split;
    std::vector<float> out((size_t)5 * sub_T);
    int32_t T = 0;
    const int st = kcbs_rrt_plan(ctx, r, start, goals + 2 * r, &p, (int)fc.robot.size(), fc.robot.data(), fc.k0.data(), fc.ln.data(),
                                 fc.off.data(), fc.states.data(), sub_T, out.data(), &T, &rs);
    if (st != KCBS_OK) return st;
    if (k_split > 0 && !rs.solved) return KCBS_OK;   // (the caller plans from the start instead; traj / len untouched)
    std::vector<float> joined((size_t)5 * std::max(k_split + T, 1), 0.0f);
    const int L = k_split + T;
    for (int c = 0; c < 5; ++c) {
        for (int k = 0; k < k_split; ++k) joined[(size_t)c * L + k] = traj[(size_t)c * L0 + k];
        for (int k = 0; k < T; ++k) joined[(size_t)c * L + k_split + k] = out[(size_t)c * sub_T + k];
    }
    traj.swap(joined);
    len = L;
    rs.path_steps = L > 0 ? L - 1 : 0;
    return KCBS_OK;
}

// A merged pair planned as ONE 10-d system (SystemMergerDatabase.h:108-156): joint RRT over both cars' states and
// controls through k_propagate_pair (freeze-at-goal rule, the two cars kept apart by the merged isValid), goal = both cars
// within 1.5 of their goals.  Both trajectories come back with the same length.
int plan_pair(kcbs_ctx *ctx, int r1, int r2, const float *starts, const float *goals, int R, const kcbs_rrt_params &p,
              const std::vector<std::shared_ptr<Constraint>> &cons, int max_T, std::vector<float> &traj1, std::vector<float> &traj2,
              int &len, kcbs_rrt_stats &rs)
{
    float start[10], goal[4] = {goals[2 * r1], goals[2 * r1 + 1], goals[2 * r2], goals[2 * r2 + 1]};
    for (int c = 0; c < 5; ++c) {
        start[c] = starts[(size_t)c * R + r1];
        start[5 + c] = starts[(size_t)c * R + r2];
    }
    FlatConstraints fc(cons);
    std::vector<float> out((size_t)10 * max_T);
    int32_t T = 0;
    const int st = kcbs_rrt_plan_pair(ctx, r1, r2, start, goal, &p, (int)cons.size(), fc.robot.data(), fc.k0.data(), fc.ln.data(),
                                      fc.off.data(), fc.states.data(), max_T, out.data(), &T, &rs);
    if (st != KCBS_OK) return st;
    len = T;
    traj1.assign((size_t)5 * std::max(T, 1), 0.0f);
    traj2.assign((size_t)5 * std::max(T, 1), 0.0f);
    for (int c = 0; c < 5; ++c)
        for (int k = 0; k < T; ++k) {
            traj1[(size_t)c * T + k] = out[(size_t)c * max_T + k];
            traj2[(size_t)c * T + k] = out[(size_t)(5 + c) * max_T + k];
        }
    return KCBS_OK;
}

}  // namespace

int kcbs_rrt_plan_many(kcbs_ctx *ctx, int n_jobs, const int32_t *robots, const float *starts, const float *goals,
                       const kcbs_rrt_params *params, const uint64_t *seeds, int max_T, float *traj_out, int32_t *T_out,
                       kcbs_rrt_stats *stats)
{
    if (!ctx) return KCBS_ERR_INVALID_ARGUMENT;
    if (n_jobs < 0 || max_T < 1) return fail(ctx, KCBS_ERR_INVALID_ARGUMENT, "n_jobs or max_T out of range");
    if (n_jobs == 0) return KCBS_OK;
    if (!robots || !starts || !goals || !traj_out || !T_out) return fail(ctx, KCBS_ERR_INVALID_ARGUMENT, "null argument");
    kcbs_rrt_params base;
    if (params)
        base = *params;
    else
        kcbs_rrt_defaults(&base);
    return run_planner_jobs(ctx, n_jobs, kPlannerLanes, [&](kcbs_ctx *lane, int j) {
        kcbs_rrt_params lp = base;
        if (seeds) lp.seed = seeds[j];
        float start[5];
        for (int c = 0; c < 5; ++c) start[c] = starts[(size_t)c * n_jobs + j];
        kcbs_rrt_stats rs;
        const int robot = robots[j];
        const int st = rrt_plan<1>(lane, &robot, start, goals + 2 * (size_t)j, lp, 0, nullptr, nullptr, nullptr, nullptr, nullptr,
                                   max_T, traj_out + (size_t)j * 5 * max_T, T_out + j, &rs);
        if (stats) stats[j] = rs;
        return st;
    });
}

// ---- K-CBS search over a fixed set of individuals (1 robot, or 2 after a merge) -------------------------------------

extern "C++" {
namespace {

struct Search {
    kcbs_ctx *ctx;
    int R;
    const float *starts, *goals;
    kcbs_kcbs_params P;
    int max_T;
    double t_begin;
    kcbs_kcbs_stats *S;
    std::vector<std::vector<int>> ind;  // robots of every individual
    std::vector<int> ind_of;            // robot -> individual
};

// Low-level plan of one individual into `node`: a single robot through the 5-d RRT, a merged pair through the joint 10-d
// RRT (a few seeds are tried before a merged pair is given up: the joint space is large).  Returns KCBS_OK and sets `solved`.
int plan_individual(kcbs_ctx *lane, const Search &q, CbsNode &node, int I, unsigned long long seed, bool &solved,
                    bool keep_prefix = true)
{
    const std::vector<int> &members = q.ind[I];
    kcbs_rrt_params lp = q.P.low_level;
    kcbs_rrt_stats rs;
    if (members.size() == 1) {
        const int r = members[0];
        lp.seed = q.P.low_level.seed * 1000003ull + seed * 131ull + (unsigned long long)r;  // seed = salt of this plan
        // a child node's re-plan keeps the parent's path until replan_back() steps before the newest constraint begins
        // (or before the old path ends, when the robot is already parked by then); when that cannot be completed the
        // robot is planned from its start state, as the root does
        int k_split = 0;
        if (keep_prefix && node.len[r] > 1 && !node.cons[I].empty() && replan_back() > 0)
            k_split = std::max(0, std::min(node.cons[I].back()->k0, node.len[r] - 1) - replan_back());
        int st = plan_robot(lane, r, q.starts, q.goals, q.R, lp, node.cons[I], q.max_T, node.traj[r], node.len[r], rs, k_split);
        if (st == KCBS_OK && k_split > 0 && !rs.solved)
            st = plan_robot(lane, r, q.starts, q.goals, q.R, lp, node.cons[I], q.max_T, node.traj[r], node.len[r], rs, 0);
        solved = st == KCBS_OK && rs.solved;
        return st;
    }
    const int r1 = members[0], r2 = members[1];
    solved = false;
    // the joint space has twice the dimensions: twice the targets per round (the corridor swap needs them)
    lp.n_targets = (int)std::min<long long>(2ll * lp.n_targets, (1ll << 26) / std::max(1, lp.n_controls));
    for (int attempt = 0; attempt < 3 && !solved; ++attempt) {
        lp.seed = q.P.low_level.seed * 1000003ull + (seed + 7ull * attempt) * 131ull + (unsigned long long)(64 * r1 + r2 + 4096);
        int len = 0;
        const int st = plan_pair(lane, r1, r2, q.starts, q.goals, q.R, lp, node.cons[I], q.max_T, node.traj[r1], node.traj[r2], len, rs);
        if (st != KCBS_OK) return st;
        node.len[r1] = node.len[r2] = len;
        solved = rs.solved != 0;
        if (q.P.time_limit_s > 0 && now_s() - q.t_begin > q.P.time_limit_s) break;
    }
    return KCBS_OK;
}

// Best-first search of the constraint tree.  Returns KCBS_OK with `solution` set, or with merge_a >= 0 when a pair of
// single-robot individuals exceeded the merge bound (the caller merges and restarts), or with neither (time / node
// budget exhausted, or the root has no plan).
int search_tree(Search &q, std::shared_ptr<CbsNode> &solution, int &merge_a, int &merge_b)
{
    kcbs_ctx *ctx = q.ctx;
    kcbs_kcbs_stats &S = *q.S;
    const int R = q.R, NI = (int)q.ind.size();
    solution.reset();
    merge_a = merge_b = -1;
    long long next_id = 0;
    std::vector<int> pair_conflicts((size_t)NI * NI, 0);

    // root: every individual planned on its own (KCBS::solve root node), all at once on the planner lanes
    auto root = std::make_shared<CbsNode>();
    root->traj.resize(R); root->len.assign(R, 0); root->cons.resize(NI);
    bool root_ok = true;
    {
        std::vector<char> ok(NI, 0);
        const int st = run_planner_jobs(ctx, NI, kPlannerLanes, [&](kcbs_ctx *lane, int I) {
            bool solved = false;
            const int rc = plan_individual(lane, q, *root, I, 7919ull * (unsigned long long)S.merges, solved);
            ok[I] = solved;
            return rc;
        });
        if (st != KCBS_OK) return st;
        for (int I = 0; I < NI; ++I) {
            S.low_level_calls += (int)q.ind[I].size();
            if (!ok[I]) root_ok = false;
        }
        for (int r = 0; r < R; ++r) root->cost += root->len[r];
    }
    if (S.merges == 0) S.root_seconds = now_s() - q.t_begin;
    std::priority_queue<std::shared_ptr<CbsNode>, std::vector<std::shared_ptr<CbsNode>>, NodeOrder> open;
    if (root_ok) {
        root->id = next_id++;
        open.push(root);
    }

    std::vector<float> traj3;
    std::vector<int32_t> lens(R), group(R);
    for (int r = 0; r < R; ++r) group[r] = q.ind_of[r];
    while (!open.empty()) {
        if (q.P.time_limit_s > 0 && now_s() - q.t_begin > q.P.time_limit_s) break;
        if (q.P.max_nodes > 0 && S.nodes_expanded >= q.P.max_nodes) break;
        auto node = open.top();
        open.pop();
        ++S.nodes_expanded;
        // validate the plan: [R][3][T] with short paths padded (the kernel applies the padding rule)
        int T = 1;
        for (int r = 0; r < R; ++r) T = std::max(T, node->len[r]);
        const int Tp = (T + 3) & ~3;  // keeps rows 16-byte aligned for the bulk-copy path
        traj3.assign((size_t)R * 3 * Tp, 0.0f);
        for (int r = 0; r < R; ++r) {
            const int L = node->len[r];
            lens[r] = L;
            for (int k = 0; k < L; ++k) {
                traj3[((size_t)r * 3 + 0) * Tp + k] = node->traj[r][(size_t)0 * L + k];
                traj3[((size_t)r * 3 + 1) * Tp + k] = node->traj[r][(size_t)1 * L + k];
                traj3[((size_t)r * 3 + 2) * Tp + k] = node->traj[r][(size_t)4 * L + k];
            }
        }
        const double t_exp0 = now_s();
        kcbs_conflict cf;
        // time indices beyond T are never scanned: every path is padded to T, and k_hi = Tp only repeats the
        // padded states of [T, Tp), which cannot introduce an earlier conflict than the ones in [0, T).  Bodies of
        // one (merged) individual are never tested against each other: its low level keeps them apart.
        int st = kcbs_first_conflict_groups(ctx, 1, R, Tp, traj3.data(), lens.data(), NI < R ? group.data() : nullptr, &cf);
        if (st != KCBS_OK) return st;
        ++S.conflict_checks;
        const double t_exp1 = now_s();
        if (cf.r1 < 0) {
            solution = node;
            return KCBS_OK;
        }
        cf.k_end = std::min(cf.k_end, T - 1);
        // merge bound (KCBS::setMergeBound): a pair of single robots that keeps colliding becomes one individual
        const int I1 = q.ind_of[cf.r1], I2 = q.ind_of[cf.r2];
        const int n_conf = ++pair_conflicts[(size_t)std::min(I1, I2) * NI + std::max(I1, I2)];
        // (merging something that is already a merged pair is not supported, SystemMergerDatabase.h:103-104)
        if (n_conf > q.P.merge_bound && q.ind[I1].size() == 1 && q.ind[I2].size() == 1) {
            merge_a = std::min(I1, I2);
            merge_b = std::max(I1, I2);
            return KCBS_OK;
        }
        // two children: each re-plans one individual against the other robot's states during the conflict
        const int pair[2][2] = {{cf.r1, cf.r2}, {cf.r2, cf.r1}};
        std::shared_ptr<CbsNode> child[2];
        bool solved[2] = {false, false};
        st = run_planner_jobs(ctx, 2, 2, [&](kcbs_ctx *lane, int ch) {
            const int me = q.ind_of[pair[ch][0]], other = pair[ch][1];
            auto c = std::make_shared<Constraint>();
            c->other = other;
            c->k0 = cf.k_start;
            const int Lo = node->len[other];
            for (int k = cf.k_start; k <= cf.k_end; ++k) {
                const int kk = std::min(k, Lo - 1);
                c->x.push_back(node->traj[other][(size_t)0 * Lo + kk]);
                c->y.push_back(node->traj[other][(size_t)1 * Lo + kk]);
                c->th.push_back(node->traj[other][(size_t)4 * Lo + kk]);
            }
            child[ch] = std::make_shared<CbsNode>(*node);
            child[ch]->cons[me].push_back(c);
            // the seed depends on the expansion and the branch only, so both branches can run at once
            // a pair that keeps colliding is not helped by repairs near the conflict (two cars that must pass each other in a
            // corridor): from its kLocalRepairs-th conflict on, its re-plans start from the robots' start states again
            return plan_individual(lane, q, *child[ch], me,
                                   (unsigned long long)(S.nodes_expanded * 2 + ch) + 7919ull * (unsigned long long)S.merges, solved[ch],
                                   n_conf <= kLocalRepairs);
        });
        if (st != KCBS_OK) return st;
        if (trace_solve())
            fprintf(stderr, "[kcbs] expansion %d: validation %.0f us, two re-plans %.0f us (robots %d / %d, window %d..%d, %zu / %zu constraints)\n",
                    S.nodes_expanded, (t_exp1 - t_exp0) * 1e6, (now_s() - t_exp1) * 1e6, cf.r1, cf.r2, cf.k_start, cf.k_end,
                    child[0]->cons[q.ind_of[cf.r1]].size(), child[1]->cons[q.ind_of[cf.r2]].size());
        for (int ch = 0; ch < 2; ++ch) {
            S.low_level_calls += (int)q.ind[q.ind_of[pair[ch][0]]].size();
            if (!solved[ch]) {
                ++S.approximate_solutions;  // the fork counts low-level time-outs as approximate solutions
                continue;
            }
            child[ch]->cost = 0;
            for (int r = 0; r < R; ++r) child[ch]->cost += child[ch]->len[r];
            child[ch]->id = next_id++;
            open.push(child[ch]);
        }
    }
    return KCBS_OK;
}

}  // namespace
}  // extern "C++"

int kcbs_kcbs_solve(kcbs_ctx *ctx, int R, const float *starts, const float *goals, const kcbs_kcbs_params *params,
                    int max_T, float *plan_out, int32_t *lengths_out, kcbs_kcbs_stats *stats)
{
    if (!ctx) return KCBS_ERR_INVALID_ARGUMENT;
    if (R < 1 || R > ctx->n_robots) return fail(ctx, KCBS_ERR_UNKNOWN_ROBOT, "R exceeds the robots of this world");
    if (!starts || !goals || !plan_out || !lengths_out || max_T < 2) return fail(ctx, KCBS_ERR_INVALID_ARGUMENT, "null argument");
    kcbs_kcbs_stats S;
    std::memset(&S, 0, sizeof S);
    Search q;
    q.ctx = ctx; q.R = R; q.starts = starts; q.goals = goals; q.max_T = max_T; q.S = &S;
    if (params)
        q.P = *params;
    else
        kcbs_kcbs_defaults(&q.P);
    q.t_begin = now_s();
    q.ind.resize(R);
    q.ind_of.resize(R);
    for (int r = 0; r < R; ++r) {
        q.ind[r] = {r};
        q.ind_of[r] = r;
    }

    std::shared_ptr<CbsNode> solution;
    for (;;) {
        int ma = -1, mb = -1;
        const int st = search_tree(q, solution, ma, mb);
        if (st != KCBS_OK) return st;
        if (solution || ma < 0) break;
        // SystemMerger::merge(ma, mb): the pair becomes one individual, placed last, and the search starts over (as the
        // fork does); from now on it is planned as one 10-d system
        std::vector<int> merged = q.ind[ma];
        merged.insert(merged.end(), q.ind[mb].begin(), q.ind[mb].end());
        q.ind.erase(q.ind.begin() + mb);
        q.ind.erase(q.ind.begin() + ma);
        q.ind.push_back(merged);
        ++S.merges;
        for (size_t I = 0; I < q.ind.size(); ++I)
            for (int r : q.ind[I]) q.ind_of[r] = (int)I;
        if (q.P.time_limit_s > 0 && now_s() - q.t_begin > q.P.time_limit_s) break;
    }

    if (solution) {
        for (int r = 0; r < R; ++r) {
            const int L = solution->len[r];
            lengths_out[r] = L;
            for (int c = 0; c < 5; ++c)
                for (int k = 0; k < max_T; ++k)
                    plan_out[((size_t)r * 5 + c) * max_T + k] = k < L ? solution->traj[r][(size_t)c * L + k] : 0.0f;
        }
        S.solved = 1;
    } else {
        for (int r = 0; r < R; ++r) lengths_out[r] = 0;
    }
    S.seconds = now_s() - q.t_begin;
    if (stats) *stats = S;
    return KCBS_OK;
}

}  // extern "C"
