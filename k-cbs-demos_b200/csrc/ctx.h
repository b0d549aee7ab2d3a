// ctx.h -- the context object behind kcbs_ctx and the error / workspace helpers shared by api.cu and
// planner.cu (internal).
#pragma once

#include <mutex>
#include <string>
#include <vector>

#include "kcbs_device.cuh"
#include "kernels.h"

namespace kcbs {

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
};

struct RrtWorkspace;  // planner.cu
struct HostPipeline;  // api_propagate.cu

// Sharded plan validation (kcbs_peer_* / kcbs_first_conflict_sharded_dev): this context's key mailbox and the peers'
// mailboxes as mapped into this process.
struct PeerState {
    unsigned long long *box = nullptr;  // own mailbox [2 parities][KCBS_MAX_PEERS sources][box_plans]
    int box_plans = 0;
    unsigned *epoch = nullptr;          // device: call counter (message tag)
    int *done = nullptr;                // device: plans finished in the running call
    int *h_error = nullptr, *d_error = nullptr;  // host-mapped timeout flag
    int rank = 0, world = 0;            // world == 0: not connected
    unsigned long long *peer[KCBS_MAX_PEERS] = {};
    bool ipc_opened[KCBS_MAX_PEERS] = {};
};

}  // namespace kcbs

struct kcbs_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    kcbs::DevWorld world;
    int n_robots = 0;
    long long launches = 0;
    std::string err;
    // grow-only device workspaces for the host-buffer entry points
    kcbs::DevBuf in0, in1, in2, in3, out0, out1, out2, keys;
    int key_plans = 0;      // plans the keys/tickets workspace was laid out for
    kcbs::DevBuf scanws;    // look-back workspace of packed expansions: [4] control words, then one status word per CTA
    int scan_ctas = 0;      // CTAs the workspace holds (kept zeroed by the kernel)
    int prop_rounds = 0;    // launch shape of the next expansion: 0 = from n, 1 = latency shape, 2 = throughput shape
    int user_prop_rounds = 0;  // what kcbs_set_expansion_mode asked for (the planner picks per batch of jobs when 0)
    kcbs::HostPipeline *pipe = nullptr;  // chunked H2D / kernel / D2H pipeline of the host entry points (api_propagate.cu)
    kcbs::PeerState peers;
    void *scene = nullptr;  // cull grid + obstacle table in device memory
    std::vector<unsigned long long> h_grid;
    std::vector<float> h_obs4;
    std::vector<unsigned short> h_fine;  // fine three-valued obstacle grid
    kcbs::RrtWorkspace *rrt = nullptr;  // planner state, created on first use
    // planner workers: shallow clones (same device, world and scene buffers; own stream and planner workspace) on which
    // kcbs_kcbs_solve runs independent low-level plans concurrently; owned by this context
    std::vector<kcbs_ctx *> workers;
};

namespace kcbs {

int fail(kcbs_ctx *ctx, int status, const char *what, cudaError_t e = cudaSuccess);
// Stream captures and device-wide synchronisation exclude each other, process wide: synchronising the device while ANY of
// its streams is capturing is invalid and kills that capture (the planner lanes capture their round loops while other
// threads work).  The gate is held for the whole of a capture sequence and around every cudaDeviceSynchronize().
std::mutex &capture_gate();
cudaError_t device_sync();   // cudaDeviceSynchronize() under the gate
int reserve(kcbs_ctx *ctx, DevBuf &b, size_t bytes);
void destroy_rrt_workspace(kcbs_ctx *ctx);
void destroy_peers(kcbs_ctx *ctx);
void destroy_pipeline(kcbs_ctx *ctx);
cudaStream_t pick(kcbs_ctx *ctx, void *stream);
int check_robot(kcbs_ctx *ctx, int robot);

#define KCBS_CUDA(ctx, call)                                                                                    \
    do {                                                                                                        \
        cudaError_t e_ = (call);                                                                                \
        if (e_ != cudaSuccess)                                                                                  \
            return kcbs::fail(ctx, e_ == cudaErrorMemoryAllocation ? KCBS_ERR_OUT_OF_MEMORY : KCBS_ERR_CUDA, #call, e_); \
    } while (0)

}  // namespace kcbs
