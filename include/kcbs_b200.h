/*
 * kcbs_b200.h -- C ABI of the B200-native K-CBS low-level hot path (libkcbs_b200.so).
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types, integer status
 * returns, never throws.  Each entry point names the reference interface it replaces
 * (file:line relative to the K-CBS-Demos tree).  The host-side mirror of the reference's OMPL
 * plugin classes (k-cbs-demos_b200/includes/) is written on top of these calls.
 *
 * Conventions
 *   state  s = (x, y, v, phi, theta)   OMPL "reals" order, StatePropagatorDatabase.h:46
 *   control u = (a, phidot)            StatePropagatorDatabase.h:47, ControlSpaceDatabase.h:43-54
 *   All batches are structure-of-arrays float32 "planes": plane c of a batch of n starts at
 *   base + c * n.  Entry points without suffix take HOST pointers and copy through the context's
 *   device workspace; `_dev` entry points take DEVICE pointers on the context's device and are
 *   asynchronous on `stream` (a cudaStream_t passed as void*; NULL = the context's own stream, pass
 *   cudaStreamLegacy for the legacy default stream).
 *   There is no CPU fallback: every compute call runs sm_100a kernels or returns an error.
 */
#ifndef KCBS_B200_H
#define KCBS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define KCBS_API
#else
#define KCBS_API __attribute__((visibility("default")))
#endif

#define KCBS_ABI_VERSION 1
#define KCBS_MAX_OBSTACLES 64 /* Congested10x10 uses 9 (demo_Congested10x10_7robots...:82-90) */
#define KCBS_MAX_ROBOTS 64    /* largest demo has 30 (demo_Benchmark_Empty32x32_30robots...)   */
#define KCBS_MAX_STEPS 32     /* setMinMaxControlDuration(1, 10), demo_*.cpp:118               */

typedef enum kcbs_status {
    KCBS_OK = 0,
    KCBS_ERR_INVALID_ARGUMENT = 1, /* null pointer, negative size, index out of range          */
    KCBS_ERR_CUDA = 2,             /* CUDA runtime/driver failure; see kcbs_last_error         */
    KCBS_ERR_NO_DEVICE = 3,        /* no sm_100-class device: there is no CPU fallback         */
    KCBS_ERR_OUT_OF_MEMORY = 4,
    KCBS_ERR_UNKNOWN_ROBOT = 5,    /* mirrors unordered_map::at throwing,                       *
                                    * StateValidityCheckerDatabase.h:111                        */
    KCBS_ERR_PEER_TIMEOUT = 6      /* a sharded scan waited KCBS_PEER_TIMEOUT_S for a peer      */
} kcbs_status;

/* Immutable scene description, copied by kcbs_create (mirrors the by-value copies of the robot
 * map and obstacle set in homogeneous2ndOrderCarSystemSVC's constructor,
 * StateValidityCheckerDatabase.h:47-49). */
typedef struct kcbs_world {
    double lo[4];            /* StateSpaceDatabase.h:51-58 lower bounds of (x, y, v, phi)       */
    double hi[4];            /* upper bounds                                                    */
    double step_size;        /* si->setPropagationStepSize(0.1), demo_*.cpp:115                 */
    double integration_step; /* ODEBasicSolver intStep (1e-2), call site demo_*.cpp:111         */
    double car_length;       /* StatePropagatorDatabase.h:53 (0.5)                              */
    int32_t n_obstacles;
    int32_t n_robots;
    const double *obstacles;  /* n_obstacles x (x, y, length, width): RectangularObstacle ctor,  *
                               * Obstacle.h:82                                                   */
    const double *robot_dims; /* n_robots x (length, width): RectangularRobot ctor, Robot.h:82   */
} kcbs_world;

/* First conflict of one plan; replaces the result of KCBS::findConflicts (call sites
 * demo_*.cpp:155,168, Benchmark.h:151-158). r1 = -1 when the plan is conflict free. */
typedef struct kcbs_conflict {
    int32_t r1, r2;  /* robot indices in addIndividual order, r1 < r2                           */
    int32_t k_start; /* earliest colliding time index                                           */
    int32_t k_end;   /* last consecutive colliding index of the same pair                       */
} kcbs_conflict;

typedef struct kcbs_ctx kcbs_ctx;

/* ---- life cycle ------------------------------------------------------------------------- */

KCBS_API int kcbs_abi_version(void);
KCBS_API const char *kcbs_status_string(int status);

/* Fills *w with the constants every demo uses: bounds of createBounded2ndOrderCarStateSpace
 * (StateSpaceDatabase.h:42-62) for x_max, y_max; step 0.1, integration step 0.01, car length 0.5. */
KCBS_API void kcbs_world_defaults(kcbs_world *w, double x_max, double y_max);

/* Host-only introspection (needs no device): the three-valued obstacle grid kcbs_create derives from `world` and the
 * validity / propagation kernels consult before the exact footprint test of
 * StateValidityCheckerDatabase.h:62-73,180-205.  codes[iy * dim + ix] over a dim x dim grid spanning [lo, hi] of (x, y):
 * 0 = no obstacle can touch a robot centred in the cell, 0xffff = every heading collides, 0xfffe = more than two
 * obstacles need the exact test, else (o1 + 1) | (o2 + 1) << 8 = the one or two obstacles that need it.
 * Pass codes = NULL to query *dim only. */
KCBS_API int kcbs_scene_grid(const kcbs_world *world, int32_t *dim, uint16_t *codes);

/* Creates a context on CUDA device `device`.  Fails with KCBS_ERR_NO_DEVICE when the device is
 * absent or not compute capability 10.x. */
KCBS_API int kcbs_create(const kcbs_world *world, int device, kcbs_ctx **out);
KCBS_API void kcbs_destroy(kcbs_ctx *ctx);
/* Last error text of this context (or of the failed kcbs_create when ctx == NULL). */
KCBS_API const char *kcbs_last_error(const kcbs_ctx *ctx);
/* Blocks until all work queued on the context stream has finished. */
KCBS_API int kcbs_sync(kcbs_ctx *ctx);
/* Pinned host memory for callers that want full-rate transfers through the host entry points. */
KCBS_API int kcbs_host_alloc(kcbs_ctx *ctx, size_t bytes, void **out);
KCBS_API int kcbs_host_free(kcbs_ctx *ctx, void *p);

/* ---- (1) batched tree expansion ------------------------------------------------------------
 * Replaces n calls of oc::SpaceInformation::propagateWhileValid(state, control, steps, result)
 * with the propagator of demo_*.cpp:111-112 (SecondOrderCarODE + RK4 + theta wrap,
 * StatePropagatorDatabase.h:44-74) and the validity checker of
 * StateValidityCheckerDatabase.h:54-75, for robot `robot`.
 *
 *   parents     5 planes of n_parents floats.
 *   parent_idx  n indices into parents, or NULL: then n_parents must be 1 (one shared parent,
 *               a tree-node expansion) or n (sample i starts at parent i).
 *   controls    2 planes of n floats (a, phidot).
 *   steps       n requested step counts in [0, max_steps], or NULL for max_steps everywhere.
 *   seg_out     [5][max_steps][n] floats: seg_out[(c*max_steps + j)*n + i] is component c of
 *               sample i after j+1 steps for j < valid_steps[i]; rows j >= valid_steps[i] are zero.  May be NULL.
 *   final_out   [5][n] floats, last valid state (the parent when 0 steps are valid). May be NULL.
 *   valid_steps n ints: the propagateWhileValid return value.
 */
KCBS_API int kcbs_propagate_batch(kcbs_ctx *ctx, int robot, int64_t n, int64_t n_parents, const float *parents,
                                  const int32_t *parent_idx, const float *controls, const int32_t *steps,
                                  int max_steps, float *seg_out, float *final_out, int32_t *valid_steps);
KCBS_API int kcbs_propagate_batch_dev(kcbs_ctx *ctx, int robot, int64_t n, int64_t n_parents,
                                      const float *parents, const int32_t *parent_idx, const float *controls,
                                      const int32_t *steps, int max_steps, float *seg_out, float *final_out,
                                      int32_t *valid_steps, void *stream);

/* Packed (CSR) segments: the same expansion, without the padding of the dense layout.  Sample i owns a run of rows
 * that starts at offsets_out[i], in five planes (packed_out[c * packed_cap + row] is component c); the first
 * valid_steps[i] rows of the run are its states, in step order.  A sample owns as many rows as it can have valid states
 * at all: its requested duration cut where v or phi leave their bounds (both are linear in time, so that is known before
 * anything is integrated); rows it owns but did not reach -- it left the workspace or hit an obstacle earlier -- are
 * zero, as are the up to 3 rows that round a block of samples up to a 16-byte boundary.  On the Empty32x32 expansion
 * workload that is 1.05 rows per valid state: ~21 bytes leave the GPU per propagated state instead of 20 * max_steps per
 * sample.  The runs of a block of 256 / 512 consecutive samples follow each other in sample order; the blocks themselves
 * are placed first come, first served (one atomic per block, no block waits for another), so offsets_out is NOT
 * monotone across blocks and the placement may differ from call to call; the states do not.
 *   packed_out   [5][packed_cap] floats; n * max_steps + 4 * (n / 256 + 1) rows always suffice.  Rows that do not fit
 *                are dropped (offsets stay exact: offsets_out[n] > packed_cap tells).
 *   offsets_out  n + 1 ints: the first row of every sample, and in offsets_out[n] the number of rows emitted.
 *   max_steps    at most KCBS_PACKED_MAX_STEPS (the reference's setMinMaxControlDuration(1, 10), demo_*.cpp:118).
 *   total_out    (host call) rows emitted = offsets_out[n]; may be NULL.
 * The host call cuts the batch into chunks and pipelines upload / kernel / download of consecutive chunks; the
 * download is exactly the emitted rows.  final_out may be NULL; valid_steps may be NULL in the device call.  One context
 * runs one packed call at a time (the row allocator lives in the context). */
#define KCBS_PACKED_MAX_STEPS 10
KCBS_API int kcbs_propagate_packed(kcbs_ctx *ctx, int robot, int64_t n, int64_t n_parents, const float *parents,
                                   const int32_t *parent_idx, const float *controls, const int32_t *steps, int max_steps,
                                   float *packed_out, int64_t packed_cap, int32_t *offsets_out, float *final_out,
                                   int32_t *valid_steps, int64_t *total_out);
KCBS_API int kcbs_propagate_packed_dev(kcbs_ctx *ctx, int robot, int64_t n, int64_t n_parents, const float *parents,
                                       const int32_t *parent_idx, const float *controls, const int32_t *steps,
                                       int max_steps, float *packed_out, int64_t packed_cap, int32_t *offsets_out,
                                       float *final_out, int32_t *valid_steps, void *stream);
/* Launch shape of the expansion kernel.  0 (default): chosen per call from n -- a launch that cannot fill the GPU on
 * its own (up to ~110K samples, e.g. ONE 64K-control expansion) runs one packed pair per thread (twice the threads,
 * half the critical path), larger ones two pairs per thread (balanced threads).  1 / 2 force the first / second shape:
 * 2 is the better choice when many small expansions run concurrently on different streams.  Results are
 * bit-identical in every shape. */
KCBS_API int kcbs_set_expansion_mode(kcbs_ctx *ctx, int mode);

/* ---- (2) batched state validity ------------------------------------------------------------
 * Replaces n calls of homogeneous2ndOrderCarSystemSVC::isValid
 * (StateValidityCheckerDatabase.h:54-75) for robot `robot`.
 *   states   5 planes of n floats.
 *   bitmask  (n + 31) / 32 words; bit (i & 31) of word i >> 5 is 1 when state i is valid.
 */
KCBS_API int kcbs_validity_batch(kcbs_ctx *ctx, int robot, int64_t n, const float *states, uint32_t *bitmask);
KCBS_API int kcbs_validity_batch_dev(kcbs_ctx *ctx, int robot, int64_t n, const float *states, uint32_t *bitmask,
                                     void *stream);

/* ---- (3) plan validation ---------------------------------------------------------------------
 * Replaces KCBS::findConflicts' scan over homogeneous2ndOrderCarSystemSVC::areStatesValid
 * (StateValidityCheckerDatabase.h:78-125) for n_plans candidate plans at once.
 *   traj     [n_plans][R][3][T] floats: per robot the x, y and theta planes, one state per
 *            propagation step (already interpolated).
 *   lengths  [n_plans][R] path lengths (1..T), shorter paths are padded with their last state;
 *            NULL when every path has T states.
 *   out      n_plans results: the earliest (k, r1, r2) in lexicographic order, extended to k_end.
 * R is the number of robots of the world given to kcbs_create (robot_dims rows are used by index).
 */
KCBS_API int kcbs_first_conflict(kcbs_ctx *ctx, int n_plans, int R, int T, const float *traj,
                                 const int32_t *lengths, kcbs_conflict *out);
KCBS_API int kcbs_first_conflict_dev(kcbs_ctx *ctx, int n_plans, int R, int T, const float *traj,
                                     const int32_t *lengths, kcbs_conflict *out, void *stream);

/* Partial scan for sharding one plan set across GPUs: only time indices [k_lo, k_hi) are
 * examined and the result is the packed key (k << 16 | r1 << 8 | r2), UINT64_MAX when none;
 * the minimum over all shards is the first conflict.  keys = n_plans device words. */
KCBS_API int kcbs_conflict_keys_dev(kcbs_ctx *ctx, int n_plans, int R, int T, int k_lo, int k_hi,
                                    const float *traj, const int32_t *lengths, uint64_t *keys, void *stream);
/* Turns reduced keys into kcbs_conflict records (runs the k_end extension). */
KCBS_API int kcbs_conflict_finish_dev(kcbs_ctx *ctx, int n_plans, int R, int T, const float *traj,
                                      const int32_t *lengths, const uint64_t *keys, kcbs_conflict *out,
                                      void *stream);

/* ---- (3c) plan validation sharded over the GPUs of one node (BASELINE config 5) ------------
 * One process per GPU.  Every rank holds the same plan batch (the all-gathered planned trajectories) and calls
 * kcbs_first_conflict_sharded_dev with the same arguments: rank r scans the r-th slab of time indices
 * (kcbs_time_slab) of every plan -- 1 / world of the HBM traffic and of the pair tests -- and the only exchange of the
 * path, the 8-byte packed first-conflict key per plan and rank (an unsigned-min all-reduce), happens INSIDE the scan
 * kernel over NVLink peer memory: the last CTA of a plan on a rank stores the rank's key straight into every peer's
 * mailbox (one tagged 64-bit store per peer, the tag is the call's epoch, so the data is its own flag), waits for the
 * peers' keys of that plan, reduces, extends the conflict (k_end) and writes the record.  After the call `out` holds
 * all n_plans results on EVERY rank.  No NCCL call and no second launch is on the critical path.
 * Semantics = kcbs_first_conflict_dev (scan order of StateValidityCheckerDatabase.h:78-125 through the key order).
 *
 * Set-up (once per context): kcbs_peer_export allocates the mailbox and fills `out` with an opaque handle; the
 * ranks exchange the handles (any transport; they are plain bytes) and call kcbs_peer_connect with all `world`
 * handles in rank order.  Ranks may be processes (CUDA IPC) or contexts of one process.  A wait that sees no
 * peer for KCBS_PEER_TIMEOUT_S gives up: the results of that plan become (-2, -2, -2, -2) and
 * kcbs_peer_status returns KCBS_ERR_PEER_TIMEOUT -- a dead peer never hangs the GPU. */
#define KCBS_MAX_PEERS 16
#define KCBS_PEER_TIMEOUT_S 4
typedef struct kcbs_peer_handle {
    unsigned char bytes[128];
} kcbs_peer_handle;
KCBS_API int kcbs_peer_export(kcbs_ctx *ctx, int max_plans, kcbs_peer_handle *out);
KCBS_API int kcbs_peer_connect(kcbs_ctx *ctx, int rank, int world, const kcbs_peer_handle *handles);
KCBS_API int kcbs_peer_disconnect(kcbs_ctx *ctx);
/* KCBS_OK, or KCBS_ERR_PEER_TIMEOUT when a sharded scan since the last call gave up waiting (host-visible flag, no sync). */
KCBS_API int kcbs_peer_status(kcbs_ctx *ctx);
/* Time indices [*k_lo, *k_hi) of a T-step plan that `rank` of `world` scans (whole tiles of the scan kernel, so every
 * rank keeps the bulk-copy path; empty when T has fewer tiles than ranks). */
KCBS_API int kcbs_time_slab(int T, int rank, int world, int32_t *k_lo, int32_t *k_hi);
KCBS_API int kcbs_first_conflict_sharded_dev(kcbs_ctx *ctx, int n_plans, int R, int T, const float *traj,
                                             const int32_t *lengths, kcbs_conflict *out, void *stream);

/* ---- (3b) merged pair (two cars planned as one 10-d system after SystemMerger::merge) -------------------
 * State q = (x1, y1, v1, phi1, theta1, x2, y2, v2, phi2, theta2), control = (a1, phidot1, a2, phidot2):
 * 10 / 4 SoA planes.  robot1 / robot2 index the world's robot_dims rows of the two cars.
 *
 * kcbs_propagate_pair_batch replaces propagateWhileValid over TwoSecondOrderCarStatePropagator::propagate
 * (StatePropagatorDatabase.h:130-153: both cars integrated by TwoSecondOrderCarsODE :77-102 with the theta wrap of
 * :106-117, then a car that was within hold_radius (1.5, :170) of its goal before the step keeps its state) and
 * homogeneousTwo2ndOrderCarSystemSVC::isValid (StateValidityCheckerDatabase.h:225-261).  goals = (gx1, gy1, gx2, gy2).
 * kcbs_validity_pair_batch replaces that isValid alone.  Layouts as in (1) / (2) with 10 planes. */
KCBS_API int kcbs_propagate_pair_batch(kcbs_ctx *ctx, int robot1, int robot2, int64_t n, int64_t n_parents,
                                       const float *parents, const int32_t *parent_idx, const float *controls,
                                       const int32_t *steps, int max_steps, const float *goals, float hold_radius,
                                       float *seg_out, float *final_out, int32_t *valid_steps);
KCBS_API int kcbs_propagate_pair_batch_dev(kcbs_ctx *ctx, int robot1, int robot2, int64_t n, int64_t n_parents,
                                           const float *parents, const int32_t *parent_idx, const float *controls,
                                           const int32_t *steps, int max_steps, const float *goals, float hold_radius,
                                           float *seg_out, float *final_out, int32_t *valid_steps, void *stream);
KCBS_API int kcbs_validity_pair_batch(kcbs_ctx *ctx, int robot1, int robot2, int64_t n, const float *states,
                                      uint32_t *bitmask);
KCBS_API int kcbs_validity_pair_batch_dev(kcbs_ctx *ctx, int robot1, int robot2, int64_t n, const float *states,
                                          uint32_t *bitmask, void *stream);
/* Plan validation when some individuals are merged pairs: traj rows are BODIES (a merged individual contributes
 * two consecutive rows), group[R] names the individual of every body; bodies of one individual are never tested
 * against each other (areStatesValid of StateValidityCheckerDatabase.h:87-107 / :264-287 tests a merged
 * individual's two cars against the other robot only).  Results name bodies.  group is a HOST pointer in both. */
KCBS_API int kcbs_first_conflict_groups(kcbs_ctx *ctx, int n_plans, int R, int T, const float *traj,
                                        const int32_t *lengths, const int32_t *group, kcbs_conflict *out);
KCBS_API int kcbs_first_conflict_groups_dev(kcbs_ctx *ctx, int n_plans, int R, int T, const float *traj,
                                            const int32_t *lengths, const int32_t *group, kcbs_conflict *out,
                                            void *stream);

/* ---- (4) "next" rows: batched low-level planner and K-CBS high level -----------
 * These replace the un-vendored planners the reference allocates: oc::RRT through allocateControlRRT
 * (PlannerAllocatorDatabase.h:40-45) and omrc::KCBS (demo_*.cpp:155-168, Benchmark.h:151-173).  They are
 * built from the three kernels above plus tree bookkeeping kernels; all planning state lives in HBM. */

typedef struct kcbs_rrt_params {
    int32_t n_targets;      /* tree nodes expanded per iteration                  (default 1024;   *
                             * kcbs_kcbs_solve doubles it for a merged pair: 10-d joint space)     */
    int32_t n_controls;     /* sampled controls per expanded node                 (default 32)     *
                             * n_targets * n_controls = 32,768 controls per expansion round        */
    int32_t max_nodes;      /* tree capacity; 0 = from the workspace: 1.5 x the rounds a path across it needs,
                               32 .. 128 rounds of n_targets nodes                 (default 0)     */
    int32_t max_iterations; /*                                                    (default 512)    */
    int32_t min_steps;      /* setMinMaxControlDuration(1, 10), demo_*.cpp:118                      */
    int32_t max_steps;
    float goal_bias;        /* oc::RRT goal bias                                  (default 0.05)   */
    float goal_radius;      /* GoalRegion2ndOrderCar threshold, GoalRegionDatabase.h:48 (0.5)       */
    double time_limit_s;    /* setLowLevelSolveTime(5.), Benchmark.h:153                            */
    uint64_t seed;
} kcbs_rrt_params;

typedef struct kcbs_rrt_stats {
    int32_t solved;      /* 1 when a node inside the goal region was found */
    int32_t iterations;
    int32_t nodes;
    int32_t path_steps;  /* propagation steps of the returned path */
    double seconds;
} kcbs_rrt_stats;

typedef struct kcbs_kcbs_params {
    kcbs_rrt_params low_level;
    double time_limit_s;  /* planner->solve(180.0), demo_*.cpp:168 */
    int32_t max_nodes;    /* constraint-tree nodes to expand at most (0 = unlimited) */
    int32_t merge_bound;  /* KCBS::setMergeBound (demo_*.cpp: 10 / 20): a pair of robots with more conflicts than this is
                           * merged into one individual; default INT32_MAX (never), as in the fork */
} kcbs_kcbs_params;

typedef struct kcbs_kcbs_stats {
    int32_t solved;
    int32_t nodes_expanded;         /* KCBS::getNumberOfNodesExpanded, Benchmark.h:167            */
    int32_t approximate_solutions;  /* KCBS::getNumberOfApproximateSolutions, Benchmark.h:168     */
    int32_t low_level_calls;
    double root_seconds;            /* KCBS::getRootSolveTime, Benchmark.h:166                     */
    double seconds;
    int64_t conflict_checks;        /* plans validated by kcbs_first_conflict                      */
    int32_t merges;                 /* pairs merged into one individual (SystemMerger::merge)      */
    int32_t reserved_;
} kcbs_kcbs_stats;

KCBS_API void kcbs_rrt_defaults(kcbs_rrt_params *p);
KCBS_API void kcbs_kcbs_defaults(kcbs_kcbs_params *p);

/* Kinodynamic RRT for robot `robot` from `start` (5 reals) to the disc of goal_radius around goal_xy,
 * expanding n_targets tree nodes x n_controls controls per round through k_propagate.  Dynamic-obstacle
 * constraints (the K-CBS low-level input): constraint c forbids overlapping robot cons_robot[c] placed at
 * cons_states[(s*total) + cons_offset[c] + (k - cons_k0[c])] (s = x, y, theta plane; total = sum of
 * |cons_len|) at absolute time index k in [cons_k0[c], cons_k0[c] + cons_len[c]).  A NEGATIVE cons_len[c] marks a
 * persistent constraint of -cons_len[c] states whose last state keeps holding for every later time index (a robot
 * parked at its goal).  The planned robot itself is taken to wait at its last state after its path ends.  A goal
 * node whose path would not fit max_T states is never accepted (a solved plan is never truncated).  traj_out = [5][max_T] planes (state at every propagation
 * step, index 0 = start); *T_out = number of states.  All pointers are HOST pointers. */
KCBS_API int kcbs_rrt_plan(kcbs_ctx *ctx, int robot, const float *start, const float *goal_xy,
                           const kcbs_rrt_params *params, int n_cons, const int32_t *cons_robot,
                           const int32_t *cons_k0, const int32_t *cons_len, const int32_t *cons_offset,
                           const float *cons_states, int max_T, float *traj_out, int32_t *T_out,
                           kcbs_rrt_stats *stats);

/* The same planner for a MERGED PAIR: robots robot1 and robot2 planned as one 10-d system (what the reference plans
 * after SystemMerger::merge, SystemMergerDatabase.h:108-156): joint tree over q = (car 1, car 2), controls (a1, w1, a2, w2),
 * expansion through the merged propagator (TwoSecondOrderCarStatePropagator, StatePropagatorDatabase.h:119-171: a car within
 * KCBS_PAIR_GOAL_RADIUS of its goal keeps its state) and the merged validity checker
 * (homogeneousTwo2ndOrderCarSystemSVC, StateValidityCheckerDatabase.h:225-261: the two cars never overlap each other);
 * the goal is GoalRegionTwo2ndOrderCars (GoalRegionDatabase.h:61-108): both cars strictly within KCBS_PAIR_GOAL_RADIUS of
 * their goals.  start = 10 reals, goals_xy = (gx1, gy1, gx2, gy2), constraints bind both cars, traj_out = [10][max_T]. */
#define KCBS_PAIR_GOAL_RADIUS 1.5f
KCBS_API int kcbs_rrt_plan_pair(kcbs_ctx *ctx, int robot1, int robot2, const float *start, const float *goals_xy,
                                const kcbs_rrt_params *params, int n_cons, const int32_t *cons_robot,
                                const int32_t *cons_k0, const int32_t *cons_len, const int32_t *cons_offset,
                                const float *cons_states, int max_T, float *traj_out, int32_t *T_out,
                                kcbs_rrt_stats *stats);

/* n_jobs independent, unconstrained low-level plans at once (the re-plans of different robots / constraint-tree nodes
 * that one GPU owns when they are sharded over GPUs; also what the root of kcbs_kcbs_solve does): job j plans robot
 * robots[j] from starts[(c * n_jobs) + j] (c = 0..4) to goals[2 j .. 2 j + 1] with seeds[j].  The jobs run concurrently
 * on up to 16 planner lanes of the context (host thread + stream + workspace each); results do not depend on the
 * interleaving.  traj_out = [n_jobs][5][max_T], T_out / stats = [n_jobs] (stats may be NULL). */
KCBS_API int kcbs_rrt_plan_many(kcbs_ctx *ctx, int n_jobs, const int32_t *robots, const float *starts, const float *goals,
                                const kcbs_rrt_params *params, const uint64_t *seeds, int max_T, float *traj_out,
                                int32_t *T_out, kcbs_rrt_stats *stats);

/* K-CBS over the R robots of the world: root = independent low-level plans; best-first search over the
 * constraint tree with kcbs_first_conflict as the plan validator; every conflict spawns two children that
 * re-plan one robot each against the other's colliding states (after a pair's first conflict a re-plan keeps the parent
 * node's path until 4 s before its newest constraint begins and plans the rest from there; from the start state when
 * that fails and for every later conflict of the pair).
 * starts = [5][R] planes, goals = [R][2];
 * plan_out = [R][5][max_T], lengths_out = [R].  A pair with more than merge_bound conflicts is merged (the role of
 * SystemMerger::merge, SystemMergerDatabase.h:55-168) and the search restarts with it as one individual, placed last,
 * whose bodies are never tested against each other by the high level; its low level is the joint 10-d planner
 * kcbs_rrt_plan_pair.  A merged pair is never merged again (SystemMergerDatabase.h:103-104). */
KCBS_API int kcbs_kcbs_solve(kcbs_ctx *ctx, int R, const float *starts, const float *goals,
                             const kcbs_kcbs_params *params, int max_T, float *plan_out, int32_t *lengths_out,
                             kcbs_kcbs_stats *stats);

/* ---- measurement support ------------------------------------------------------------------- */

/* Runs an FFMA-saturating kernel and reports the device's sustained FP32 rate in TFLOP/s
 * (2 flops per FMA); the denominator of the propagation roofline. */
KCBS_API int kcbs_measure_fp32_peak(kcbs_ctx *ctx, double *tflops);
/* Same probe with the packed fma.rn.f32x2 instruction (two FMAs per lane per instruction). */
KCBS_API int kcbs_measure_fp32x2_peak(kcbs_ctx *ctx, double *tflops);
/* Number of kernels this context has launched since creation. */
KCBS_API int64_t kcbs_launch_count(const kcbs_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* KCBS_B200_H */
