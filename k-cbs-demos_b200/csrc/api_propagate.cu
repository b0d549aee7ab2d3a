// api_propagate.cu -- entry points of the batched tree expansion (include/kcbs_b200.h, section 1): the device-pointer
// calls, and the host-buffer calls as a chunked pipeline (H2D of chunk k+1 and the kernel of chunk k run while chunk
// k-1 is copied back), dense or packed (CSR) segments.
#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "ctx.h"

using namespace kcbs;

namespace kcbs {

constexpr int kPipeSlots = 3;             // chunks in flight
constexpr int kPipeChunks = 4;            // chunks a large call is cut into at least
// samples per chunk at least: smaller chunks cost more in launches and copy set-up than they hide (KCBS_PIPE_MIN_CHUNK
// overrides it for experiments)
static int64_t pipe_min_chunk()
{
    static const int64_t v = [] {
        const char *e = std::getenv("KCBS_PIPE_MIN_CHUNK");
        const long long x = e ? std::atoll(e) : 0;
        return (int64_t)(x >= 2048 ? x : 32768);
    }();
    return v;
}

// Host calls of the process in flight (all contexts): see propagate_host.
static std::atomic<int> g_host_calls{0};
constexpr int kCrowdedCallers = 3;
constexpr int64_t kCrowdedChunk = 131072;
struct ActiveHostCall {
    int others;
    ActiveHostCall() : others(g_host_calls.fetch_add(1, std::memory_order_relaxed)) {}
    ~ActiveHostCall() { g_host_calls.fetch_sub(1, std::memory_order_relaxed); }
};

struct PipeSlot {
    cudaStream_t stream = nullptr;
    cudaEvent_t done = nullptr;
    DevBuf par, ctl, idx, steps, seg, fin, valid, offsets, scan;
    int scan_ctas = 0;
    int32_t *h_offsets = nullptr;  // pinned: chunk-local offsets [cn + 1], then the valid counts [cn] (packed mode: one copy)
    size_t h_offsets_cap = 0;
};

struct HostPipeline {
    PipeSlot slot[kPipeSlots];
    cudaEvent_t start = nullptr;
};

void destroy_pipeline(kcbs_ctx *ctx)
{
    if (!ctx->pipe) return;
    for (PipeSlot &s : ctx->pipe->slot) {
        if (s.stream) cudaStreamSynchronize(s.stream);
        DevBuf *bufs[] = {&s.par, &s.ctl, &s.idx, &s.steps, &s.seg, &s.fin, &s.valid, &s.offsets, &s.scan};
        for (DevBuf *b : bufs)
            if (b->p) cudaFree(b->p);
        if (s.h_offsets) cudaFreeHost(s.h_offsets);
        if (s.done) cudaEventDestroy(s.done);
        if (s.stream) cudaStreamDestroy(s.stream);
    }
    if (ctx->pipe->start) cudaEventDestroy(ctx->pipe->start);
    delete ctx->pipe;
    ctx->pipe = nullptr;
}

static int ensure_pipeline(kcbs_ctx *ctx)
{
    if (ctx->pipe) return KCBS_OK;
    HostPipeline *p = new (std::nothrow) HostPipeline();
    if (!p) return fail(ctx, KCBS_ERR_OUT_OF_MEMORY, "host allocation failed");
    ctx->pipe = p;
    KCBS_CUDA(ctx, cudaEventCreateWithFlags(&p->start, cudaEventDisableTiming));
    for (PipeSlot &s : p->slot) {
        KCBS_CUDA(ctx, cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
        KCBS_CUDA(ctx, cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
    }
    return KCBS_OK;
}

// Zeroed look-back workspace for `ctas` CTAs in `buf` ([4] control words + status words); the kernel keeps it zeroed.
static int ensure_scan(kcbs_ctx *ctx, DevBuf &buf, int &have, int ctas, cudaStream_t s)
{
    if (ctas <= have) return KCBS_OK;
    const int want = ctas + ctas / 2 + 64;
    if (have > 0) KCBS_CUDA(ctx, device_sync());
    int st = reserve(ctx, buf, 16 + sizeof(unsigned long long) * (size_t)want);
    if (st != KCBS_OK) return st;
    KCBS_CUDA(ctx, cudaMemsetAsync(buf.p, 0, 16 + sizeof(unsigned long long) * (size_t)want, s));
    have = want;
    return KCBS_OK;
}

static int check_prop_args(kcbs_ctx *ctx, int robot, int64_t n, int64_t n_parents, const float *parents,
                           const int32_t *parent_idx, const float *controls, int max_steps)
{
    int st = check_robot(ctx, robot);
    if (st != KCBS_OK) return st;
    if (n < 0 || n_parents < 1 || max_steps < 0 || max_steps > KCBS_MAX_STEPS)
        return fail(ctx, KCBS_ERR_INVALID_ARGUMENT, "n, n_parents or max_steps out of range");
    if (n == 0) return KCBS_OK;
    if (!parents || !controls) return fail(ctx, KCBS_ERR_INVALID_ARGUMENT, "parents / controls is null");
    if (!parent_idx && n_parents != 1 && n_parents != n)
        return fail(ctx, KCBS_ERR_INVALID_ARGUMENT, "without parent_idx, n_parents must be 1 or n");
    if (n > (int64_t)0x7fffffff * 256) return fail(ctx, KCBS_ERR_INVALID_ARGUMENT, "n too large");
    return KCBS_OK;
}

static PropArgs make_args(kcbs_ctx *ctx, int robot, int64_t n, int64_t n_parents, const float *parents,
                          const int32_t *parent_idx, const float *controls, const int32_t *steps, int max_steps,
                          float *seg, float *fin, int32_t *valid)
{
    PropArgs a = {};
    a.n = n; a.n_parents = n_parents; a.parents = parents; a.parent_idx = parent_idx; a.controls = controls;
    a.steps = steps; a.seg = seg; a.fin = fin; a.valid_steps = valid; a.max_steps = max_steps;
    a.robot = robot; a.rounds = ctx->prop_rounds;
    return a;
}

}  // namespace kcbs

extern "C" {

int kcbs_set_expansion_mode(kcbs_ctx *ctx, int mode)
{
    if (!ctx) return KCBS_ERR_INVALID_ARGUMENT;
    if (mode < 0 || mode > 2) return fail(ctx, KCBS_ERR_INVALID_ARGUMENT, "mode must be 0 (auto), 1 (latency) or 2 (throughput)");
    ctx->prop_rounds = ctx->user_prop_rounds = mode;
    return KCBS_OK;
}

int kcbs_propagate_batch_dev(kcbs_ctx *ctx, int robot, int64_t n, int64_t n_parents, const float *parents,
                             const int32_t *parent_idx, const float *controls, const int32_t *steps, int max_steps,
                             float *seg_out, float *final_out, int32_t *valid_steps, void *stream)
{
    if (!ctx) return KCBS_ERR_INVALID_ARGUMENT;
    int st = check_prop_args(ctx, robot, n, n_parents, parents, parent_idx, controls, max_steps);
    if (st != KCBS_OK || n == 0) return st;
    const PropArgs a = make_args(ctx, robot, n, n_parents, parents, parent_idx, controls, steps, max_steps, seg_out, final_out, valid_steps);
    KCBS_CUDA(ctx, cudaSetDevice(ctx->device));
    KCBS_CUDA(ctx, launch_propagate(ctx->world, a, pick(ctx, stream)));
    ctx->launches += 1;
    return KCBS_OK;
}

int kcbs_propagate_packed_dev(kcbs_ctx *ctx, int robot, int64_t n, int64_t n_parents, const float *parents,
                              const int32_t *parent_idx, const float *controls, const int32_t *steps, int max_steps,
                              float *packed_out, int64_t packed_cap, int32_t *offsets_out, float *final_out,
                              int32_t *valid_steps, void *stream)
{
    if (!ctx) return KCBS_ERR_INVALID_ARGUMENT;
    int st = check_prop_args(ctx, robot, n, n_parents, parents, parent_idx, controls, max_steps);
    if (st != KCBS_OK) return st;
    if (max_steps > KCBS_PACKED_MAX_STEPS) return fail(ctx, KCBS_ERR_INVALID_ARGUMENT, "packed segments need max_steps <= KCBS_PACKED_MAX_STEPS");
    if (!offsets_out || packed_cap < 0 || (packed_cap > 0 && !packed_out)) return fail(ctx, KCBS_ERR_INVALID_ARGUMENT, "offsets_out / packed_out is null");
    if (n * (int64_t)max_steps > 0x7fffffff) return fail(ctx, KCBS_ERR_INVALID_ARGUMENT, "n * max_steps exceeds 32-bit row offsets");
    cudaStream_t s = pick(ctx, stream);
    KCBS_CUDA(ctx, cudaSetDevice(ctx->device));
    if (n == 0) {
        KCBS_CUDA(ctx, cudaMemsetAsync(offsets_out, 0, sizeof(int32_t), s));
        return KCBS_OK;
    }
    PropArgs a = make_args(ctx, robot, n, n_parents, parents, parent_idx, controls, steps, max_steps, packed_out, final_out, valid_steps);
    a.offsets = offsets_out;
    a.packed_cap = packed_cap;
    if ((st = ensure_scan(ctx, ctx->scanws, ctx->scan_ctas, propagate_grid(a), s))) return st;
    a.scan_ctl = (unsigned *)ctx->scanws.p;
    a.scan = (unsigned long long *)((char *)ctx->scanws.p + 16);
    KCBS_CUDA(ctx, launch_propagate(ctx->world, a, s));
    ctx->launches += 1;
    return KCBS_OK;
}

}  // extern "C"

namespace kcbs {

// The host-buffer calls.  The batch is cut into chunks; chunk c runs on pipeline slot c % kPipeSlots (own stream and
// device buffers): inputs up, kernel, results down.  Packed mode needs the chunk's row count on the host before its
// rows can be copied to their place, so chunk c's rows are issued after chunk c+1's kernel is queued.
static int propagate_host(kcbs_ctx *ctx, int robot, int64_t n, int64_t n_parents, const float *parents,
                          const int32_t *parent_idx, const float *controls, const int32_t *steps, int max_steps,
                          float *seg_out, float *packed_out, int64_t packed_cap, int32_t *offsets_out, float *final_out,
                          int32_t *valid_steps, int64_t *total_out)
{
    const bool packed = offsets_out != nullptr;
    int st = check_prop_args(ctx, robot, n, n_parents, parents, parent_idx, controls, max_steps);
    if (st != KCBS_OK) return st;
    if (packed) {
        if (max_steps > KCBS_PACKED_MAX_STEPS) return fail(ctx, KCBS_ERR_INVALID_ARGUMENT, "packed segments need max_steps <= KCBS_PACKED_MAX_STEPS");
        if (packed_cap < 0 || (packed_cap > 0 && !packed_out)) return fail(ctx, KCBS_ERR_INVALID_ARGUMENT, "packed_out is null");
        if (n * (int64_t)max_steps > 0x7fffffff) return fail(ctx, KCBS_ERR_INVALID_ARGUMENT, "n * max_steps exceeds 32-bit row offsets");
        offsets_out[0] = 0;
        if (total_out) *total_out = 0;
    }
    if (n == 0) return KCBS_OK;
    KCBS_CUDA(ctx, cudaSetDevice(ctx->device));
    if ((st = ensure_pipeline(ctx))) return st;
    HostPipeline &P = *ctx->pipe;

    // chunks: multiples of 2048 samples (whole CTAs of either shape, 16-byte aligned planes).  The pipeline inside a call
    // pays when the call has the PCIe link to itself; when other host threads of the process are inside host calls of
    // their own contexts (one planner thread per robot), their calls overlap each other and bigger copies use the link
    // better (measured: 43 -> 46 GB/s device to host with ten callers), so such a call goes up and down in one piece
    ActiveHostCall active_call;
    const bool crowded = active_call.others >= kCrowdedCallers;
    int64_t chunk = std::max<int64_t>(crowded ? kCrowdedChunk : pipe_min_chunk(), (n + kPipeChunks - 1) / kPipeChunks);
    chunk = (chunk + 2047) / 2048 * 2048;
    const int n_chunks = (int)((n + chunk - 1) / chunk);
    const bool par_chunked = !parent_idx && n_parents == n && n_parents != 1;   // sample i starts at parent i

    // parents that are not chunked with the samples go up once (ctx->in0), ahead of every chunk
    const float *d_parents = nullptr;
    if (!par_chunked) {
        if ((st = reserve(ctx, ctx->in0, sizeof(float) * 5 * (size_t)n_parents))) return st;
        KCBS_CUDA(ctx, cudaMemcpyAsync(ctx->in0.p, parents, sizeof(float) * 5 * (size_t)n_parents, cudaMemcpyHostToDevice, ctx->stream));
        KCBS_CUDA(ctx, cudaEventRecord(P.start, ctx->stream));
        d_parents = (const float *)ctx->in0.p;
    }

    // rows a chunk of cn samples can emit at most: its requested steps + the rows that align every block of samples
    auto chunk_rows = [&](int64_t cn) { return ((cn * max_steps + 4 * (cn / 256 + 2)) + 3) & ~(int64_t)3; };
    struct Pending {
        int slot = -1;
        int64_t c0 = 0, cn = 0;
    } pending;
    int64_t running = 0;   // packed rows emitted so far
    auto finish_packed_chunk = [&](const Pending &q) -> int {
        PipeSlot &S = P.slot[q.slot];
        KCBS_CUDA(ctx, cudaEventSynchronize(S.done));   // the chunk's offsets are on the host
        const int64_t total = S.h_offsets[q.cn];
        for (int64_t i = 0; i < q.cn; ++i) offsets_out[q.c0 + i] = (int32_t)(running + S.h_offsets[i]);
        offsets_out[q.c0 + q.cn] = (int32_t)(running + total);
        if (valid_steps) std::memcpy(valid_steps + q.c0, S.h_offsets + q.cn + 1, sizeof(int32_t) * (size_t)q.cn);
        const int64_t room = std::max<int64_t>(0, std::min<int64_t>(total, packed_cap - running));
        if (room > 0)
            KCBS_CUDA(ctx, cudaMemcpy2DAsync(packed_out + running, sizeof(float) * (size_t)packed_cap, S.seg.p,
                                             sizeof(float) * (size_t)chunk_rows(q.cn), sizeof(float) * (size_t)room, 5,
                                             cudaMemcpyDeviceToHost, S.stream));
        running += total;
        return KCBS_OK;
    };

    for (int c = 0; c < n_chunks; ++c) {
        const int64_t c0 = (int64_t)c * chunk, cn = std::min<int64_t>(chunk, n - c0);
        PipeSlot &S = P.slot[c % kPipeSlots];
        cudaStream_t s = S.stream;
        if (packed && c >= kPipeSlots) {
            // the slot's previous chunk must have handed over its rows before its buffers are reused (its D2H is queued on
            // this very stream, so stream order does the rest): chunks finish in order, at most one is pending
            if (pending.slot == c % kPipeSlots) {
                if ((st = finish_packed_chunk(pending))) return st;
                pending.slot = -1;
            }
        }
        const size_t b_ctl = sizeof(float) * 2 * (size_t)cn, b_idx = sizeof(int32_t) * (size_t)cn;
        const size_t b_fin = sizeof(float) * 5 * (size_t)cn;
        const size_t b_seg = packed ? sizeof(float) * 5 * (size_t)chunk_rows(cn) : sizeof(float) * 5 * (size_t)max_steps * (size_t)cn;
        if ((st = reserve(ctx, S.ctl, b_ctl)) || (st = reserve(ctx, S.idx, parent_idx ? b_idx : 0)) ||
            (st = reserve(ctx, S.steps, steps ? b_idx : 0)) || (st = reserve(ctx, S.par, par_chunked ? b_fin : 0)) ||
            (st = reserve(ctx, S.seg, (seg_out || packed) ? b_seg : 0)) || (st = reserve(ctx, S.fin, final_out ? b_fin : 0)) ||
            (st = reserve(ctx, S.valid, packed ? 0 : b_idx)) || (st = reserve(ctx, S.offsets, packed ? 2 * b_idx + sizeof(int32_t) : 0)))
            return st;
        if (packed && S.h_offsets_cap < 2 * (size_t)cn + 1) {
            if (S.h_offsets) KCBS_CUDA(ctx, cudaFreeHost(S.h_offsets));
            S.h_offsets = nullptr;
            KCBS_CUDA(ctx, cudaHostAlloc((void **)&S.h_offsets, sizeof(int32_t) * (2 * (size_t)chunk + 1), cudaHostAllocDefault));
            S.h_offsets_cap = 2 * (size_t)chunk + 1;
        }
        // inputs of the chunk: planes of the host batch have stride n, planes of the chunk stride cn
        KCBS_CUDA(ctx, cudaMemcpy2DAsync(S.ctl.p, sizeof(float) * (size_t)cn, controls + c0, sizeof(float) * (size_t)n,
                                         sizeof(float) * (size_t)cn, 2, cudaMemcpyHostToDevice, s));
        if (parent_idx) KCBS_CUDA(ctx, cudaMemcpyAsync(S.idx.p, parent_idx + c0, b_idx, cudaMemcpyHostToDevice, s));
        if (steps) KCBS_CUDA(ctx, cudaMemcpyAsync(S.steps.p, steps + c0, b_idx, cudaMemcpyHostToDevice, s));
        const float *dp = d_parents;
        int64_t np = n_parents;
        if (par_chunked) {
            KCBS_CUDA(ctx, cudaMemcpy2DAsync(S.par.p, sizeof(float) * (size_t)cn, parents + c0, sizeof(float) * (size_t)n,
                                             sizeof(float) * (size_t)cn, 5, cudaMemcpyHostToDevice, s));
            dp = (const float *)S.par.p;
            np = cn;
        } else {
            KCBS_CUDA(ctx, cudaStreamWaitEvent(s, P.start, 0));
        }
        PropArgs a = make_args(ctx, robot, cn, np, dp, parent_idx ? (const int32_t *)S.idx.p : nullptr, (const float *)S.ctl.p,
                               steps ? (const int32_t *)S.steps.p : nullptr, max_steps,
                               (seg_out || packed) ? (float *)S.seg.p : nullptr, final_out ? (float *)S.fin.p : nullptr,
                               packed ? (int32_t *)S.offsets.p + cn + 1 : (int32_t *)S.valid.p);
        if (packed) {
            a.offsets = (int32_t *)S.offsets.p;
            a.packed_cap = chunk_rows(cn);
            if ((st = ensure_scan(ctx, S.scan, S.scan_ctas, propagate_grid(a), s))) return st;
            a.scan_ctl = (unsigned *)S.scan.p;
            a.scan = (unsigned long long *)((char *)S.scan.p + 16);
        }
        KCBS_CUDA(ctx, launch_propagate(ctx->world, a, s));
        ctx->launches += 1;
        // results of the chunk
        if (valid_steps && !packed) KCBS_CUDA(ctx, cudaMemcpyAsync(valid_steps + c0, S.valid.p, b_idx, cudaMemcpyDeviceToHost, s));
        if (final_out)
            KCBS_CUDA(ctx, cudaMemcpy2DAsync(final_out + c0, sizeof(float) * (size_t)n, S.fin.p, sizeof(float) * (size_t)cn,
                                             sizeof(float) * (size_t)cn, 5, cudaMemcpyDeviceToHost, s));
        if (packed) {
            // offsets and (right behind them) the valid counts come down in one copy
            KCBS_CUDA(ctx, cudaMemcpyAsync(S.h_offsets, S.offsets.p, (valid_steps ? 2 : 1) * b_idx + sizeof(int32_t), cudaMemcpyDeviceToHost, s));
            KCBS_CUDA(ctx, cudaEventRecord(S.done, s));
            if (pending.slot >= 0 && (st = finish_packed_chunk(pending))) return st;
            pending.slot = c % kPipeSlots;
            pending.c0 = c0;
            pending.cn = cn;
        } else if (seg_out) {
            KCBS_CUDA(ctx, cudaMemcpy2DAsync(seg_out + c0, sizeof(float) * (size_t)n, S.seg.p, sizeof(float) * (size_t)cn,
                                             sizeof(float) * (size_t)cn, (size_t)5 * max_steps, cudaMemcpyDeviceToHost, s));
        }
    }
    if (packed && pending.slot >= 0 && (st = finish_packed_chunk(pending))) return st;
    for (int q = 0; q < std::min(n_chunks, kPipeSlots); ++q) KCBS_CUDA(ctx, cudaStreamSynchronize(P.slot[q].stream));
    if (packed && total_out) *total_out = running;
    return KCBS_OK;
}

}  // namespace kcbs

extern "C" {

int kcbs_propagate_batch(kcbs_ctx *ctx, int robot, int64_t n, int64_t n_parents, const float *parents,
                         const int32_t *parent_idx, const float *controls, const int32_t *steps, int max_steps,
                         float *seg_out, float *final_out, int32_t *valid_steps)
{
    if (!ctx) return KCBS_ERR_INVALID_ARGUMENT;
    return propagate_host(ctx, robot, n, n_parents, parents, parent_idx, controls, steps, max_steps, seg_out, nullptr, 0, nullptr,
                          final_out, valid_steps, nullptr);
}

int kcbs_propagate_packed(kcbs_ctx *ctx, int robot, int64_t n, int64_t n_parents, const float *parents,
                          const int32_t *parent_idx, const float *controls, const int32_t *steps, int max_steps,
                          float *packed_out, int64_t packed_cap, int32_t *offsets_out, float *final_out,
                          int32_t *valid_steps, int64_t *total_out)
{
    if (!ctx) return KCBS_ERR_INVALID_ARGUMENT;
    if (!offsets_out) return fail(ctx, KCBS_ERR_INVALID_ARGUMENT, "offsets_out is null");
    return propagate_host(ctx, robot, n, n_parents, parents, parent_idx, controls, steps, max_steps, nullptr, packed_out, packed_cap,
                          offsets_out, final_out, valid_steps, total_out);
}

}  // extern "C"
