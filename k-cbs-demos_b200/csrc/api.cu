// api.cu -- the C ABI of include/kcbs_b200.h: context life cycle, scene upload, host-buffer and
// device-pointer entry points.  No CPU fallback anywhere: a context cannot be created without an
// sm_100-class device and every compute entry point launches the kernels of this library.
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <algorithm>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include <unistd.h>

#include "ctx.h"

using namespace kcbs;

namespace {
thread_local std::string g_create_error;
}  // namespace

namespace kcbs {

int fail(kcbs_ctx *ctx, int status, const char *what, cudaError_t e)
{
    std::string msg = what;
    if (e != cudaSuccess) {
        msg += ": ";
        msg += cudaGetErrorString(e);
    }
    if (ctx)
        ctx->err = msg;
    else
        g_create_error = msg;
    return status;
}

std::mutex &capture_gate()
{
    static std::mutex gate;
    return gate;
}

cudaError_t device_sync()
{
    std::lock_guard<std::mutex> lock(capture_gate());
    return cudaDeviceSynchronize();
}

int reserve(kcbs_ctx *ctx, DevBuf &b, size_t bytes)
{
    if (bytes <= b.cap) return KCBS_OK;
    if (b.p) {
        KCBS_CUDA(ctx, device_sync());  // callers may still use it on their own stream
        KCBS_CUDA(ctx, cudaFree(b.p));
        b.p = nullptr;
        b.cap = 0;
    }
    const size_t want = bytes + bytes / 4 + 256;
    KCBS_CUDA(ctx, cudaMalloc(&b.p, want));
    b.cap = want;
    return KCBS_OK;
}

}  // namespace kcbs

namespace {

// Largest float f with ((double)f - eps <= hi)  /  smallest float f with ((double)f + eps >= lo):
// the float image of OMPL's RealVectorStateSpace::satisfiesBounds test.
float hi_threshold(double hi, double eps)
{
    float f = (float)(hi + eps);  // within an ulp or two of the answer, so the exact searches below are short
    while ((double)f - eps > hi) f = nextafterf(f, -INFINITY);
    while ((double)nextafterf(f, INFINITY) - eps <= hi) f = nextafterf(f, INFINITY);
    return f;
}
float lo_threshold(double lo, double eps)
{
    float f = (float)(lo - eps);
    while ((double)f + eps < lo) f = nextafterf(f, INFINITY);
    while ((double)nextafterf(f, -INFINITY) + eps >= lo) f = nextafterf(f, -INFINITY);
    return f;
}

// Number of h-steps odeint::integrate_const takes over [0, duration] (ODEBasicSolver::solve).
int count_substeps(double duration, double h)
{
    double time = 0.0;
    int step = 0;
    while (((time + h) - duration) <= DBL_EPSILON && step < 1000000) {
        ++step;
        time = 0.0 + (double)step * h;
    }
    return step;
}

int build_world(const kcbs_world *src, DevWorld &w)
{
    if (!src) return fail(nullptr, KCBS_ERR_INVALID_ARGUMENT, "world is null");
    if (src->n_obstacles < 0 || src->n_obstacles > KCBS_MAX_OBSTACLES)
        return fail(nullptr, KCBS_ERR_INVALID_ARGUMENT, "n_obstacles out of range");
    if (src->n_robots < 1 || src->n_robots > KCBS_MAX_ROBOTS)
        return fail(nullptr, KCBS_ERR_INVALID_ARGUMENT, "n_robots out of range");
    if ((src->n_obstacles > 0 && !src->obstacles) || !src->robot_dims)
        return fail(nullptr, KCBS_ERR_INVALID_ARGUMENT, "obstacles / robot_dims is null");
    if (!(src->step_size > 0) || !(src->integration_step > 0) || !(src->car_length > 0))
        return fail(nullptr, KCBS_ERR_INVALID_ARGUMENT, "step_size, integration_step and car_length must be > 0");
    std::memset(&w, 0, sizeof w);
    for (int i = 0; i < 4; ++i) {
        if (!(src->lo[i] <= src->hi[i])) return fail(nullptr, KCBS_ERR_INVALID_ARGUMENT, "lo > hi");
        w.lo_t[i] = lo_threshold(src->lo[i], DBL_EPSILON);
        w.hi_t[i] = hi_threshold(src->hi[i], DBL_EPSILON);
    }
    // SO2StateSpace::satisfiesBounds is strict: value < pi + eps && value > -pi - eps
    {
        const double hi = M_PI + DBL_EPSILON, lo = -M_PI - DBL_EPSILON;
        float f = (float)hi;
        while (!((double)f < hi)) f = nextafterf(f, -INFINITY);
        w.th_hi_t = f;
        f = (float)lo;
        while (!((double)f > lo)) f = nextafterf(f, INFINITY);
        w.th_lo_t = f;
    }
    w.sym_bounds = (w.lo_t[2] == -w.hi_t[2]) && (w.lo_t[3] == -w.hi_t[3]) && (w.th_lo_t == -w.th_hi_t);
    w.h = (float)src->integration_step;
    w.step = (float)src->step_size;
    w.inv_len = (float)(1.0 / src->car_length);
    w.n_sub = count_substeps(src->step_size, src->integration_step);
    w.n_obs = src->n_obstacles;
    w.n_robots = src->n_robots;
    for (int o = 0; o < src->n_obstacles; ++o) {
        const double *ob = src->obstacles + 4 * o;  // Obstacle.h:82-99
        if (!(ob[2] >= 0) || !(ob[3] >= 0)) return fail(nullptr, KCBS_ERR_INVALID_ARGUMENT, "negative obstacle size");
        w.obs_cx[o] = (float)(ob[0] + ob[2] / 2);
        w.obs_cy[o] = (float)(ob[1] + ob[3] / 2);
        w.obs_hx[o] = (float)(ob[2] / 2);
        w.obs_hy[o] = (float)(ob[3] / 2);
    }
    float max_rad = 0.0f;
    for (int r = 0; r < src->n_robots; ++r) {
        const double *d = src->robot_dims + 2 * r;  // Robot.h:82-99
        if (!(d[0] > 0) || !(d[1] > 0)) return fail(nullptr, KCBS_ERR_INVALID_ARGUMENT, "robot size must be > 0");
        w.rob_hl[r] = (float)(d[0] / 2);
        w.rob_hw[r] = (float)(d[1] / 2);
        w.rob_rad[r] = (float)std::sqrt(d[0] * d[0] / 4 + d[1] * d[1] / 4);
        // the cull radii must never be smaller than the true radius
        if ((double)w.rob_rad[r] < std::sqrt(d[0] * d[0] / 4 + d[1] * d[1] / 4)) w.rob_rad[r] = nextafterf(w.rob_rad[r], INFINITY);
        max_rad = fmaxf(max_rad, w.rob_rad[r]);
    }
    w.max_rad = max_rad;
    return KCBS_OK;
}

// Conservative obstacle cull grid: bit o of cell (ix, iy) is set when obstacle o, grown by the largest robot
// bounding radius (plus slack for the float cell lookup), overlaps the cell.
void build_grid(const kcbs_world *src, DevWorld &w, std::vector<unsigned long long> &grid, std::vector<float> &obs4)
{
    const int G = kGridDim;
    const double x0 = src->lo[0], y0 = src->lo[1];
    const double sx = std::max(src->hi[0] - src->lo[0], 1e-9), sy = std::max(src->hi[1] - src->lo[1], 1e-9);
    const double cw = sx / G, ch = sy / G;
    w.gdim = G;
    w.gx0 = (float)x0;
    w.gy0 = (float)y0;
    w.ginvx = (float)(G / sx);
    w.ginvy = (float)(G / sy);
    grid.assign((size_t)G * G, 0ull);
    obs4.assign((size_t)4 * KCBS_MAX_OBSTACLES, 0.0f);
    for (int o = 0; o < src->n_obstacles; ++o) {
        const double *ob = src->obstacles + 4 * o;
        obs4[4 * o + 0] = w.obs_cx[o];
        obs4[4 * o + 1] = w.obs_cy[o];
        obs4[4 * o + 2] = w.obs_hx[o];
        obs4[4 * o + 3] = w.obs_hy[o];
        const double grow = (double)w.max_rad * (1.0 + 1e-5) + 1e-4 * std::max(cw, ch) + 1e-5;
        const double ax = ob[0] - grow, bx = ob[0] + ob[2] + grow, ay = ob[1] - grow, by = ob[1] + ob[3] + grow;
        int ix0 = (int)std::floor((ax - x0) / cw), ix1 = (int)std::floor((bx - x0) / cw);
        int iy0 = (int)std::floor((ay - y0) / ch), iy1 = (int)std::floor((by - y0) / ch);
        // (`grow` already holds 1e-4 of a cell of slack for the float cell lookup on the device)
        ix0 = std::max(ix0, 0); iy0 = std::max(iy0, 0);
        ix1 = std::min(ix1, G - 1); iy1 = std::min(iy1, G - 1);
        for (int iy = iy0; iy <= iy1; ++iy)
            for (int ix = ix0; ix <= ix1; ++ix) grid[(size_t)iy * G + ix] |= 1ull << o;
    }
}

// Euclidean distance from the point (px, py) to the axis-aligned rectangle [x0, x1] x [y0, y1].
double point_rect_distance(double px, double py, double x0, double y0, double x1, double y1)
{
    const double dx = std::max(0.0, std::max(x0 - px, px - x1)), dy = std::max(0.0, std::max(y0 - py, py - y1));
    return std::sqrt(dx * dx + dy * dy);
}

// Fine three-valued grid (kcbs_device.cuh, cell_code): classification of every cell against every obstacle, valid for
// any heading of any robot of the scene.  The margins (1e-5 relative + 1e-5 absolute on the radii, 1e-3 of a cell for
// the float cell lookup on the device) keep "free" and "hit" away from the decision boundary of the exact test.
void build_fine_grid(const kcbs_world *src, DevWorld &w, std::vector<unsigned short> &fine)
{
    const int G = kFineDim;
    const double x0 = src->lo[0], y0 = src->lo[1];
    const double sx = std::max(src->hi[0] - src->lo[0], 1e-9), sy = std::max(src->hi[1] - src->lo[1], 1e-9);
    const double cw = sx / G, ch = sy / G;
    w.fdim = G;
    w.finvx = (float)(G / sx);
    w.finvy = (float)(G / sy);
    w.foffx = (float)(-x0 * (G / sx));
    w.foffy = (float)(-y0 * (G / sy));
    fine.assign((size_t)G * G, 0);
    if (src->n_obstacles == 0) return;
    double max_rad = 0.0, min_in = INFINITY;
    for (int r = 0; r < src->n_robots; ++r) {
        const double *d = src->robot_dims + 2 * r;
        max_rad = std::max(max_rad, std::sqrt(d[0] * d[0] / 4 + d[1] * d[1] / 4));
        min_in = std::min(min_in, std::min(d[0], d[1]) / 2);
    }
    const double r_free = max_rad * (1.0 + 1e-5) + 1e-5, r_hit = min_in * (1.0 - 1e-5) - 1e-5;
    for (int iy = 0; iy < G; ++iy)
        for (int ix = 0; ix < G; ++ix) {
            // the cell as the device may see it: in-bounds states just outside the workspace clamp into the border cells
            const double sl = 1e-3;
            const double ax = x0 + (ix - sl) * cw - (ix == 0 ? 1e-3 : 0), bx = x0 + (ix + 1 + sl) * cw + (ix == G - 1 ? 1e-3 : 0);
            const double ay = y0 + (iy - sl) * ch - (iy == 0 ? 1e-3 : 0), by = y0 + (iy + 1 + sl) * ch + (iy == G - 1 ? 1e-3 : 0);
            bool hit = false;
            int n_maybe = 0, cand[2] = {0, 0};
            for (int o = 0; o < src->n_obstacles && !hit; ++o) {
                const double *ob = src->obstacles + 4 * o;
                const double ox0 = ob[0], oy0 = ob[1], ox1 = ob[0] + ob[2], oy1 = ob[1] + ob[3];
                const double gx = std::max(0.0, std::max(ox0 - bx, ax - ox1)), gy = std::max(0.0, std::max(oy0 - by, ay - oy1));
                if (std::sqrt(gx * gx + gy * gy) > r_free) continue;  // free
                const double far = std::max(std::max(point_rect_distance(ax, ay, ox0, oy0, ox1, oy1),
                                                     point_rect_distance(bx, ay, ox0, oy0, ox1, oy1)),
                                            std::max(point_rect_distance(ax, by, ox0, oy0, ox1, oy1),
                                                     point_rect_distance(bx, by, ox0, oy0, ox1, oy1)));
                if (far < r_hit) {
                    hit = true;
                } else {
                    if (n_maybe < 2) cand[n_maybe] = o + 1;
                    ++n_maybe;
                }
            }
            unsigned code = 0;
            if (hit)
                code = kCellHit;
            else if (n_maybe > 2)
                code = kCellMany;
            else
                code = (unsigned)cand[0] | ((unsigned)cand[1] << 8);
            fine[(size_t)iy * G + ix] = (unsigned short)code;
        }
}

}  // namespace

namespace kcbs {
cudaStream_t pick(kcbs_ctx *ctx, void *stream) { return stream ? (cudaStream_t)stream : ctx->stream; }

int check_robot(kcbs_ctx *ctx, int robot)
{
    if (robot < 0 || robot >= ctx->n_robots) return fail(ctx, KCBS_ERR_UNKNOWN_ROBOT, "robot index out of range");
    return KCBS_OK;
}
}  // namespace kcbs

extern "C" {

int kcbs_abi_version(void) { return KCBS_ABI_VERSION; }

const char *kcbs_status_string(int status)
{
    switch (status) {
        case KCBS_OK: return "ok";
        case KCBS_ERR_INVALID_ARGUMENT: return "invalid argument";
        case KCBS_ERR_CUDA: return "CUDA error";
        case KCBS_ERR_NO_DEVICE: return "no sm_100 device (no CPU fallback)";
        case KCBS_ERR_OUT_OF_MEMORY: return "out of device memory";
        case KCBS_ERR_UNKNOWN_ROBOT: return "unknown robot";
        case KCBS_ERR_PEER_TIMEOUT: return "peer timeout in a sharded scan";
        default: return "unknown status";
    }
}

void kcbs_world_defaults(kcbs_world *w, double x_max, double y_max)
{
    if (!w) return;
    std::memset(w, 0, sizeof *w);
    // StateSpaceDatabase.h:49-59
    w->lo[0] = 0.0;          w->hi[0] = x_max;
    w->lo[1] = 0.0;          w->hi[1] = y_max;
    w->lo[2] = -1.0;         w->hi[2] = 1.0;
    w->lo[3] = -M_PI / 3;    w->hi[3] = M_PI / 3;
    w->step_size = 0.1;          // demo_*.cpp:115
    w->integration_step = 1e-2;  // ODEBasicSolver default
    w->car_length = 0.5;         // StatePropagatorDatabase.h:53
}

int kcbs_scene_grid(const kcbs_world *world, int32_t *dim, uint16_t *codes)
{
    DevWorld w;
    int st = build_world(world, w);
    if (st != KCBS_OK) return st;
    if (!dim) return fail(nullptr, KCBS_ERR_INVALID_ARGUMENT, "dim is null");
    *dim = kFineDim;
    if (!codes) return KCBS_OK;
    std::vector<unsigned short> fine;
    build_fine_grid(world, w, fine);
    std::memcpy(codes, fine.data(), sizeof(unsigned short) * fine.size());
    return KCBS_OK;
}

int kcbs_create(const kcbs_world *world, int device, kcbs_ctx **out)
{
    if (!out) return fail(nullptr, KCBS_ERR_INVALID_ARGUMENT, "out is null");
    *out = nullptr;
    DevWorld w;
    int st = build_world(world, w);
    if (st != KCBS_OK) return st;

    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0)
        return fail(nullptr, KCBS_ERR_NO_DEVICE, "no CUDA device available (this library has no CPU fallback)", e);
    if (device < 0 || device >= count) return fail(nullptr, KCBS_ERR_NO_DEVICE, "device index out of range");
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) return fail(nullptr, KCBS_ERR_CUDA, "cudaGetDeviceProperties", e);
    if (prop.major != 10)
        return fail(nullptr, KCBS_ERR_NO_DEVICE, "device is not compute capability 10.x (kernels are sm_100a only)");

    kcbs_ctx *ctx = new (std::nothrow) kcbs_ctx();
    if (!ctx) return fail(nullptr, KCBS_ERR_OUT_OF_MEMORY, "host allocation failed");
    ctx->device = device;
    build_grid(world, w, ctx->h_grid, ctx->h_obs4);
    {
        // device layout of the fine grid: padded pitch, last row / column repeated (kcbs_device.cuh)
        std::vector<unsigned short> logical;
        build_fine_grid(world, w, logical);
        ctx->h_fine.assign((size_t)kFineRows * kFinePitch, 0);
        for (int iy = 0; iy < kFineRows; ++iy)
            for (int ix = 0; ix <= kFineDim; ++ix)
                ctx->h_fine[(size_t)iy * kFinePitch + ix] =
                    logical[(size_t)std::min(iy, kFineDim - 1) * kFineDim + std::min(ix, kFineDim - 1)];
    }
    ctx->world = w;
    ctx->n_robots = world->n_robots;
    if ((e = cudaSetDevice(device)) != cudaSuccess || (e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaEventCreate(&ctx->ev0)) != cudaSuccess || (e = cudaEventCreate(&ctx->ev1)) != cudaSuccess ||
        (e = init_conflict_kernels()) != cudaSuccess || (e = init_validity_kernels()) != cudaSuccess || (e = init_propagate_kernels()) != cudaSuccess ||
        (e = init_pair_kernels()) != cudaSuccess || (e = init_planner_kernels()) != cudaSuccess) {
        fail(nullptr, KCBS_ERR_CUDA, "context setup", e);
        kcbs_destroy(ctx);
        return KCBS_ERR_CUDA;
    }
    {
        const size_t b_grid = sizeof(unsigned long long) * ctx->h_grid.size(), b_obs = sizeof(float) * ctx->h_obs4.size();
        const size_t b_fine = sizeof(unsigned short) * ctx->h_fine.size();
        if ((e = cudaMalloc(&ctx->scene, b_grid + b_obs + b_fine)) != cudaSuccess ||
            (e = cudaMemcpy(ctx->scene, ctx->h_grid.data(), b_grid, cudaMemcpyHostToDevice)) != cudaSuccess ||
            (e = cudaMemcpy((char *)ctx->scene + b_grid, ctx->h_obs4.data(), b_obs, cudaMemcpyHostToDevice)) != cudaSuccess ||
            (e = cudaMemcpy((char *)ctx->scene + b_grid + b_obs, ctx->h_fine.data(), b_fine, cudaMemcpyHostToDevice)) != cudaSuccess) {
            fail(nullptr, KCBS_ERR_CUDA, "scene upload", e);
            kcbs_destroy(ctx);
            return KCBS_ERR_CUDA;
        }
        ctx->world.grid = (const unsigned long long *)ctx->scene;
        ctx->world.obs4 = (const float4 *)((char *)ctx->scene + b_grid);
        ctx->world.fine = (const unsigned short *)((char *)ctx->scene + b_grid + b_obs);
    }
    *out = ctx;
    return KCBS_OK;
}

void kcbs_destroy(kcbs_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    DevBuf *bufs[] = {&ctx->in0, &ctx->in1, &ctx->in2, &ctx->in3, &ctx->out0, &ctx->out1, &ctx->out2, &ctx->keys, &ctx->scanws};
    for (DevBuf *b : bufs)
        if (b->p) cudaFree(b->p);
    destroy_pipeline(ctx);
    destroy_rrt_workspace(ctx);
    destroy_peers(ctx);
    if (ctx->scene) cudaFree(ctx->scene);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char *kcbs_last_error(const kcbs_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int kcbs_sync(kcbs_ctx *ctx)
{
    if (!ctx) return KCBS_ERR_INVALID_ARGUMENT;
    KCBS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return KCBS_OK;
}

int kcbs_host_alloc(kcbs_ctx *ctx, size_t bytes, void **out)
{
    if (!ctx || !out) return KCBS_ERR_INVALID_ARGUMENT;
    KCBS_CUDA(ctx, cudaSetDevice(ctx->device));
    KCBS_CUDA(ctx, cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault));
    return KCBS_OK;
}

int kcbs_host_free(kcbs_ctx *ctx, void *p)
{
    if (!ctx) return KCBS_ERR_INVALID_ARGUMENT;
    if (p) KCBS_CUDA(ctx, cudaFreeHost(p));
    return KCBS_OK;
}

/* ---- (2) validity -------------------------------------------------------------------------- */

int kcbs_validity_batch_dev(kcbs_ctx *ctx, int robot, int64_t n, const float *states, uint32_t *bitmask, void *stream)
{
    if (!ctx) return KCBS_ERR_INVALID_ARGUMENT;
    int st = check_robot(ctx, robot);
    if (st != KCBS_OK) return st;
    if (n < 0) return fail(ctx, KCBS_ERR_INVALID_ARGUMENT, "n is negative");
    if (n == 0) return KCBS_OK;
    if (!states || !bitmask) return fail(ctx, KCBS_ERR_INVALID_ARGUMENT, "states / bitmask is null");
    ValidArgs a;
    a.n = n; a.states = states; a.bitmask = bitmask; a.robot = robot;
    KCBS_CUDA(ctx, cudaSetDevice(ctx->device));
    KCBS_CUDA(ctx, launch_validity(ctx->world, a, pick(ctx, stream)));
    ctx->launches += 1;
    return KCBS_OK;
}

int kcbs_validity_batch(kcbs_ctx *ctx, int robot, int64_t n, const float *states, uint32_t *bitmask)
{
    if (!ctx) return KCBS_ERR_INVALID_ARGUMENT;
    if (n < 0) return fail(ctx, KCBS_ERR_INVALID_ARGUMENT, "n is negative");
    if (n == 0) return KCBS_OK;
    if (!states || !bitmask) return fail(ctx, KCBS_ERR_INVALID_ARGUMENT, "states / bitmask is null");
    KCBS_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    const size_t b_in = sizeof(float) * 5 * (size_t)n, b_out = sizeof(uint32_t) * (size_t)((n + 31) / 32);
    int st;
    if ((st = reserve(ctx, ctx->in0, b_in)) || (st = reserve(ctx, ctx->out0, b_out))) return st;
    KCBS_CUDA(ctx, cudaMemcpyAsync(ctx->in0.p, states, b_in, cudaMemcpyHostToDevice, s));
    st = kcbs_validity_batch_dev(ctx, robot, n, (const float *)ctx->in0.p, (uint32_t *)ctx->out0.p, s);
    if (st != KCBS_OK) return st;
    KCBS_CUDA(ctx, cudaMemcpyAsync(bitmask, ctx->out0.p, b_out, cudaMemcpyDeviceToHost, s));
    KCBS_CUDA(ctx, cudaStreamSynchronize(s));
    return KCBS_OK;
}

/* ---- (3) plan validation --------------------------------------------------------------------- */

static int check_plan_args(kcbs_ctx *ctx, int n_plans, int R, int T)
{
    if (n_plans < 0 || T < 0) return fail(ctx, KCBS_ERR_INVALID_ARGUMENT, "n_plans or T is negative");
    if (R < 1 || R > ctx->n_robots) return fail(ctx, KCBS_ERR_UNKNOWN_ROBOT, "R exceeds the robots of this world");
    if ((long long)T >= (1ll << 31)) return fail(ctx, KCBS_ERR_INVALID_ARGUMENT, "T too large");
    return KCBS_OK;
}

int kcbs_conflict_keys_dev(kcbs_ctx *ctx, int n_plans, int R, int T, int k_lo, int k_hi, const float *traj,
                           const int32_t *lengths, uint64_t *keys, void *stream)
{
    if (!ctx) return KCBS_ERR_INVALID_ARGUMENT;
    int st = check_plan_args(ctx, n_plans, R, T);
    if (st != KCBS_OK) return st;
    if (n_plans == 0) return KCBS_OK;
    if (!keys || (T > 0 && !traj)) return fail(ctx, KCBS_ERR_INVALID_ARGUMENT, "traj / keys is null");
    k_lo = k_lo < 0 ? 0 : k_lo;
    k_hi = k_hi > T ? T : k_hi;
    cudaStream_t s = pick(ctx, stream);
    KCBS_CUDA(ctx, cudaSetDevice(ctx->device));
    KCBS_CUDA(ctx, launch_fill_keys((unsigned long long *)keys, nullptr, n_plans, s));
    ctx->launches += 1;
    if (k_hi > k_lo) {
        ConflictArgs a = {};
        a.traj = traj; a.lengths = lengths; a.keys = (unsigned long long *)keys; a.group = nullptr; a.tickets = nullptr; a.out = nullptr;
        a.P = n_plans; a.R = R; a.T = T; a.k_lo = k_lo; a.k_hi = k_hi;
        KCBS_CUDA(ctx, launch_conflict_keys(ctx->world, a, s));
        ctx->launches += 1;
    }
    return KCBS_OK;
}

int kcbs_conflict_finish_dev(kcbs_ctx *ctx, int n_plans, int R, int T, const float *traj, const int32_t *lengths,
                             const uint64_t *keys, kcbs_conflict *out, void *stream)
{
    if (!ctx) return KCBS_ERR_INVALID_ARGUMENT;
    int st = check_plan_args(ctx, n_plans, R, T);
    if (st != KCBS_OK) return st;
    if (n_plans == 0) return KCBS_OK;
    if (!keys || !out || (T > 0 && !traj)) return fail(ctx, KCBS_ERR_INVALID_ARGUMENT, "traj / keys / out is null");
    FinishArgs a;
    a.traj = traj; a.lengths = lengths; a.keys = (const unsigned long long *)keys; a.out = out; a.P = n_plans;
    a.R = R; a.T = T;
    KCBS_CUDA(ctx, cudaSetDevice(ctx->device));
    KCBS_CUDA(ctx, launch_conflict_finish(ctx->world, a, pick(ctx, stream)));
    ctx->launches += 1;
    return KCBS_OK;
}

// keys + tickets live in the context ([key_plans] keys, then [key_plans] tickets); the kernel leaves them clean, so
// they are filled only when the layout changes.  The layout is decided by the PLAN COUNT it was made for, never by
// the byte capacity (reserve() over-allocates): a call with more plans than that re-lays-out and re-fills both.
static int ensure_keys(kcbs_ctx *ctx, int n_plans, cudaStream_t s)
{
    if (n_plans <= ctx->key_plans) return KCBS_OK;
    const int cap_plans = n_plans + n_plans / 4 + 64;
    if (ctx->key_plans > 0) KCBS_CUDA(ctx, device_sync());  // earlier scans (any stream) still use the old layout
    int st = reserve(ctx, ctx->keys, (sizeof(uint64_t) + sizeof(int)) * (size_t)cap_plans);
    if (st != KCBS_OK) return st;
    ctx->key_plans = cap_plans;
    KCBS_CUDA(ctx, launch_fill_keys((unsigned long long *)ctx->keys.p, (int *)((uint64_t *)ctx->keys.p + cap_plans), cap_plans, s));
    ctx->launches += 1;
    return KCBS_OK;
}

static int first_conflict_dev_impl(kcbs_ctx *ctx, int n_plans, int R, int T, const float *traj, const int32_t *lengths,
                                   const int *d_group, kcbs_conflict *out, void *stream)
{
    if (!ctx) return KCBS_ERR_INVALID_ARGUMENT;
    int st = check_plan_args(ctx, n_plans, R, T);
    if (st != KCBS_OK) return st;
    if (n_plans == 0) return KCBS_OK;
    if (!out || (T > 0 && !traj)) return fail(ctx, KCBS_ERR_INVALID_ARGUMENT, "traj / out is null");
    cudaStream_t s = pick(ctx, stream);
    KCBS_CUDA(ctx, cudaSetDevice(ctx->device));
    if ((st = ensure_keys(ctx, n_plans, s))) return st;
    if (T == 0) {   // nothing to scan: every plan is conflict free
        FinishArgs f;
        f.traj = traj; f.lengths = lengths; f.keys = (const unsigned long long *)ctx->keys.p; f.out = out; f.P = n_plans;
        f.R = R; f.T = T;
        KCBS_CUDA(ctx, launch_conflict_finish(ctx->world, f, s));
        ctx->launches += 1;
        return KCBS_OK;
    }
    ConflictArgs a = {};
    a.traj = traj; a.lengths = lengths; a.keys = (unsigned long long *)ctx->keys.p;
    a.tickets = (int *)((uint64_t *)ctx->keys.p + ctx->key_plans); a.out = out; a.group = d_group;
    a.P = n_plans; a.R = R; a.T = T; a.k_lo = 0; a.k_hi = T;
    KCBS_CUDA(ctx, launch_conflict_keys(ctx->world, a, s));
    ctx->launches += 1;
    return KCBS_OK;
}

int kcbs_first_conflict_dev(kcbs_ctx *ctx, int n_plans, int R, int T, const float *traj, const int32_t *lengths,
                            kcbs_conflict *out, void *stream)
{
    return first_conflict_dev_impl(ctx, n_plans, R, T, traj, lengths, nullptr, out, stream);
}

// uploads the body -> individual table (tiny) into the context and runs the grouped scan
int kcbs_first_conflict_groups_dev(kcbs_ctx *ctx, int n_plans, int R, int T, const float *traj, const int32_t *lengths,
                                   const int32_t *group, kcbs_conflict *out, void *stream)
{
    if (!ctx) return KCBS_ERR_INVALID_ARGUMENT;
    if (!group) return first_conflict_dev_impl(ctx, n_plans, R, T, traj, lengths, nullptr, out, stream);
    if (R < 1 || R > ctx->n_robots) return fail(ctx, KCBS_ERR_UNKNOWN_ROBOT, "R exceeds the robots of this world");
    int st = reserve(ctx, ctx->in3, sizeof(int32_t) * KCBS_MAX_ROBOTS);
    if (st != KCBS_OK) return st;
    KCBS_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : ctx->stream;
    KCBS_CUDA(ctx, cudaMemcpyAsync(ctx->in3.p, group, sizeof(int32_t) * R, cudaMemcpyHostToDevice, s));
    KCBS_CUDA(ctx, cudaStreamSynchronize(s));  // `group` is the caller's host buffer
    return first_conflict_dev_impl(ctx, n_plans, R, T, traj, lengths, (const int *)ctx->in3.p, out, stream);
}

int kcbs_first_conflict_groups(kcbs_ctx *ctx, int n_plans, int R, int T, const float *traj, const int32_t *lengths,
                               const int32_t *group, kcbs_conflict *out);

int kcbs_first_conflict(kcbs_ctx *ctx, int n_plans, int R, int T, const float *traj, const int32_t *lengths,
                        kcbs_conflict *out)
{
    return kcbs_first_conflict_groups(ctx, n_plans, R, T, traj, lengths, nullptr, out);
}

int kcbs_first_conflict_groups(kcbs_ctx *ctx, int n_plans, int R, int T, const float *traj, const int32_t *lengths,
                               const int32_t *group, kcbs_conflict *out)
{
    if (!ctx) return KCBS_ERR_INVALID_ARGUMENT;
    int st = check_plan_args(ctx, n_plans, R, T);
    if (st != KCBS_OK) return st;
    if (n_plans == 0) return KCBS_OK;
    if (!out || (T > 0 && !traj)) return fail(ctx, KCBS_ERR_INVALID_ARGUMENT, "traj / out is null");
    KCBS_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    const size_t b_traj = sizeof(float) * (size_t)n_plans * R * 3 * (size_t)T;
    const size_t b_len = sizeof(int32_t) * (size_t)n_plans * R, b_out = sizeof(kcbs_conflict) * (size_t)n_plans;
    if ((st = reserve(ctx, ctx->in0, b_traj)) || (st = reserve(ctx, ctx->in2, lengths ? b_len : 0)) ||
        (st = reserve(ctx, ctx->out0, b_out)))
        return st;
    if (b_traj) KCBS_CUDA(ctx, cudaMemcpyAsync(ctx->in0.p, traj, b_traj, cudaMemcpyHostToDevice, s));
    if (lengths) KCBS_CUDA(ctx, cudaMemcpyAsync(ctx->in2.p, lengths, b_len, cudaMemcpyHostToDevice, s));
    st = kcbs_first_conflict_groups_dev(ctx, n_plans, R, T, (const float *)ctx->in0.p,
                                        lengths ? (const int32_t *)ctx->in2.p : nullptr, group, (kcbs_conflict *)ctx->out0.p, s);
    if (st != KCBS_OK) return st;
    KCBS_CUDA(ctx, cudaMemcpyAsync(out, ctx->out0.p, b_out, cudaMemcpyDeviceToHost, s));
    KCBS_CUDA(ctx, cudaStreamSynchronize(s));
    return KCBS_OK;
}

/* ---- (3c) sharded plan validation over NVLink peer memory ---------------------------------------- */

extern "C++" {
namespace {
struct PeerHandleRaw {
    uint32_t magic;
    int32_t pid, device, box_plans, ipc_ok;
    uint64_t ptr;
    cudaIpcMemHandle_t ipc;
};
static_assert(sizeof(PeerHandleRaw) <= sizeof(kcbs_peer_handle), "handle does not fit");
constexpr uint32_t kPeerMagic = 0x4b434250u;  // "KCBP"

void peer_unmap(kcbs_ctx *ctx)
{
    PeerState &ps = ctx->peers;
    for (int i = 0; i < KCBS_MAX_PEERS; ++i) {
        if (ps.ipc_opened[i] && ps.peer[i]) cudaIpcCloseMemHandle(ps.peer[i]);
        ps.peer[i] = nullptr;
        ps.ipc_opened[i] = false;
    }
    ps.world = 0;
    ps.rank = 0;
}
}  // namespace

namespace kcbs {
void destroy_peers(kcbs_ctx *ctx)
{
    peer_unmap(ctx);
    PeerState &ps = ctx->peers;
    if (ps.box) cudaFree(ps.box);
    if (ps.epoch) cudaFree(ps.epoch);
    if (ps.h_error) cudaFreeHost(ps.h_error);
    ps = PeerState();
}
}  // namespace kcbs
}  // extern "C++"

int kcbs_peer_export(kcbs_ctx *ctx, int max_plans, kcbs_peer_handle *out)
{
    if (!ctx || !out) return KCBS_ERR_INVALID_ARGUMENT;
    if (max_plans < 1 || max_plans > (1 << 24)) return fail(ctx, KCBS_ERR_INVALID_ARGUMENT, "max_plans out of range");
    KCBS_CUDA(ctx, cudaSetDevice(ctx->device));
    KCBS_CUDA(ctx, device_sync());
    destroy_peers(ctx);
    PeerState &ps = ctx->peers;
    const size_t bytes = sizeof(unsigned long long) * 2 * KCBS_MAX_PEERS * (size_t)max_plans;
    // a mailbox is written by the peers from the moment its handle leaves this call: it is cleared here, never later
    KCBS_CUDA(ctx, cudaMalloc((void **)&ps.box, bytes));
    KCBS_CUDA(ctx, cudaMemset(ps.box, 0, bytes));
    KCBS_CUDA(ctx, cudaMalloc((void **)&ps.epoch, 64));
    KCBS_CUDA(ctx, cudaMemset(ps.epoch, 0, 64));
    ps.done = (int *)(ps.epoch + 1);
    KCBS_CUDA(ctx, cudaHostAlloc((void **)&ps.h_error, sizeof(int), cudaHostAllocMapped));
    *ps.h_error = 0;
    KCBS_CUDA(ctx, cudaHostGetDevicePointer((void **)&ps.d_error, ps.h_error, 0));
    KCBS_CUDA(ctx, device_sync());
    ps.box_plans = max_plans;
    PeerHandleRaw h;
    std::memset(&h, 0, sizeof h);
    h.magic = kPeerMagic;
    h.pid = (int32_t)getpid();
    h.device = ctx->device;
    h.box_plans = max_plans;
    h.ptr = (uint64_t)(uintptr_t)ps.box;
    h.ipc_ok = cudaIpcGetMemHandle(&h.ipc, ps.box) == cudaSuccess;
    if (!h.ipc_ok) (void)cudaGetLastError();  // contexts of one process can still connect through the raw pointer
    std::memset(out, 0, sizeof *out);
    std::memcpy(out->bytes, &h, sizeof h);
    return KCBS_OK;
}

int kcbs_peer_connect(kcbs_ctx *ctx, int rank, int world, const kcbs_peer_handle *handles)
{
    if (!ctx || !handles) return KCBS_ERR_INVALID_ARGUMENT;
    if (world < 1 || world > KCBS_MAX_PEERS || rank < 0 || rank >= world)
        return fail(ctx, KCBS_ERR_INVALID_ARGUMENT, "rank / world out of range");
    PeerState &ps = ctx->peers;
    if (!ps.box) return fail(ctx, KCBS_ERR_INVALID_ARGUMENT, "kcbs_peer_export must come first");
    KCBS_CUDA(ctx, cudaSetDevice(ctx->device));
    peer_unmap(ctx);
    for (int i = 0; i < world; ++i) {
        PeerHandleRaw h;
        std::memcpy(&h, handles[i].bytes, sizeof h);
        if (h.magic != kPeerMagic) return fail(ctx, KCBS_ERR_INVALID_ARGUMENT, "not a kcbs_peer_handle");
        if (h.box_plans != ps.box_plans) return fail(ctx, KCBS_ERR_INVALID_ARGUMENT, "ranks exported different max_plans");
        if (i == rank) {
            if (h.ptr != (uint64_t)(uintptr_t)ps.box || h.pid != (int32_t)getpid())
                return fail(ctx, KCBS_ERR_INVALID_ARGUMENT, "handles[rank] is not this context's handle");
            ps.peer[i] = ps.box;
        } else if (h.pid == (int32_t)getpid()) {
            // another context of this process: the pointer is valid here; another device needs peer access
            if (h.device != ctx->device) {
                int can = 0;
                KCBS_CUDA(ctx, cudaDeviceCanAccessPeer(&can, ctx->device, h.device));
                if (!can) return fail(ctx, KCBS_ERR_CUDA, "no peer access between the devices of two ranks");
                const cudaError_t e = cudaDeviceEnablePeerAccess(h.device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail(ctx, KCBS_ERR_CUDA, "cudaDeviceEnablePeerAccess", e);
                (void)cudaGetLastError();
            }
            ps.peer[i] = (unsigned long long *)(uintptr_t)h.ptr;
        } else {
            if (!h.ipc_ok) return fail(ctx, KCBS_ERR_CUDA, "the peer could not export an IPC handle");
            void *p = nullptr;
            KCBS_CUDA(ctx, cudaIpcOpenMemHandle(&p, h.ipc, cudaIpcMemLazyEnablePeerAccess));
            ps.peer[i] = (unsigned long long *)p;
            ps.ipc_opened[i] = true;
        }
    }
    ps.rank = rank;
    ps.world = world;
    return KCBS_OK;
}

int kcbs_peer_disconnect(kcbs_ctx *ctx)
{
    if (!ctx) return KCBS_ERR_INVALID_ARGUMENT;
    KCBS_CUDA(ctx, cudaSetDevice(ctx->device));
    KCBS_CUDA(ctx, device_sync());
    peer_unmap(ctx);
    return KCBS_OK;
}

int kcbs_peer_status(kcbs_ctx *ctx)
{
    if (!ctx) return KCBS_ERR_INVALID_ARGUMENT;
    PeerState &ps = ctx->peers;
    if (ps.h_error && *(volatile int *)ps.h_error) {
        *(volatile int *)ps.h_error = 0;
        return fail(ctx, KCBS_ERR_PEER_TIMEOUT, "a sharded scan gave up waiting for a peer's keys");
    }
    return KCBS_OK;
}

int kcbs_time_slab(int T, int rank, int world, int32_t *k_lo, int32_t *k_hi)
{
    if (T < 0 || world < 1 || rank < 0 || rank >= world || !k_lo || !k_hi) return KCBS_ERR_INVALID_ARGUMENT;
    int lo, hi;
    conflict_time_slab(T, rank, world, &lo, &hi);
    *k_lo = lo;
    *k_hi = hi;
    return KCBS_OK;
}

int kcbs_first_conflict_sharded_dev(kcbs_ctx *ctx, int n_plans, int R, int T, const float *traj, const int32_t *lengths,
                                    kcbs_conflict *out, void *stream)
{
    if (!ctx) return KCBS_ERR_INVALID_ARGUMENT;
    PeerState &ps = ctx->peers;
    if (ps.world < 1) return fail(ctx, KCBS_ERR_INVALID_ARGUMENT, "kcbs_peer_connect must come first");
    // one rank, or nothing to scan: no exchange (every rank takes the same branch)
    if (ps.world == 1 || T == 0) return first_conflict_dev_impl(ctx, n_plans, R, T, traj, lengths, nullptr, out, stream);
    int st = check_plan_args(ctx, n_plans, R, T);
    if (st != KCBS_OK) return st;
    if (n_plans == 0) return KCBS_OK;
    if (n_plans > ps.box_plans) return fail(ctx, KCBS_ERR_INVALID_ARGUMENT, "n_plans exceeds the exported max_plans");
    if (!out || !traj) return fail(ctx, KCBS_ERR_INVALID_ARGUMENT, "traj / out is null");
    cudaStream_t s = pick(ctx, stream);
    KCBS_CUDA(ctx, cudaSetDevice(ctx->device));
    if ((st = ensure_keys(ctx, n_plans, s))) return st;
    ConflictArgs a = {};
    a.traj = traj; a.lengths = lengths; a.keys = (unsigned long long *)ctx->keys.p;
    a.tickets = (int *)((uint64_t *)ctx->keys.p + ctx->key_plans); a.out = out; a.group = nullptr;
    a.P = n_plans; a.R = R; a.T = T;
    conflict_time_slab(T, ps.rank, ps.world, &a.k_lo, &a.k_hi);
    for (int i = 0; i < ps.world; ++i) a.box_peer[i] = ps.peer[i];
    a.epoch = ps.epoch; a.done = ps.done; a.error = ps.d_error;
    a.rank = ps.rank; a.world = ps.world; a.box_plans = ps.box_plans;
    KCBS_CUDA(ctx, launch_conflict_keys(ctx->world, a, s));
    ctx->launches += 1;
    return KCBS_OK;
}

/* ---- (3b) merged pair ----------------------------------------------------------------------- */

int kcbs_propagate_pair_batch_dev(kcbs_ctx *ctx, int robot1, int robot2, int64_t n, int64_t n_parents, const float *parents,
                                  const int32_t *parent_idx, const float *controls, const int32_t *steps, int max_steps,
                                  const float *goals, float hold_radius, float *seg_out, float *final_out,
                                  int32_t *valid_steps, void *stream)
{
    if (!ctx) return KCBS_ERR_INVALID_ARGUMENT;
    int st;
    if ((st = check_robot(ctx, robot1)) || (st = check_robot(ctx, robot2))) return st;
    if (n < 0 || n_parents < 1 || max_steps < 0 || max_steps > KCBS_MAX_STEPS)
        return fail(ctx, KCBS_ERR_INVALID_ARGUMENT, "n, n_parents or max_steps out of range");
    if (n == 0) return KCBS_OK;
    if (!parents || !controls || !goals) return fail(ctx, KCBS_ERR_INVALID_ARGUMENT, "parents / controls / goals is null");
    if (!parent_idx && n_parents != 1 && n_parents != n)
        return fail(ctx, KCBS_ERR_INVALID_ARGUMENT, "without parent_idx, n_parents must be 1 or n");
    PairArgs a = {};
    a.n = n; a.n_parents = n_parents; a.parents = parents; a.parent_idx = parent_idx; a.controls = controls; a.steps = steps;
    a.seg = seg_out; a.fin = final_out; a.valid_steps = valid_steps; a.max_steps = max_steps; a.robot1 = robot1; a.robot2 = robot2;
    for (int i = 0; i < 4; ++i) a.goals[i] = goals[i];
    a.hold_radius = hold_radius;
    KCBS_CUDA(ctx, cudaSetDevice(ctx->device));
    KCBS_CUDA(ctx, launch_propagate_pair(ctx->world, a, pick(ctx, stream)));
    ctx->launches += 1;
    return KCBS_OK;
}

int kcbs_propagate_pair_batch(kcbs_ctx *ctx, int robot1, int robot2, int64_t n, int64_t n_parents, const float *parents,
                              const int32_t *parent_idx, const float *controls, const int32_t *steps, int max_steps,
                              const float *goals, float hold_radius, float *seg_out, float *final_out, int32_t *valid_steps)
{
    if (!ctx) return KCBS_ERR_INVALID_ARGUMENT;
    if (n < 0 || n_parents < 1 || max_steps < 0 || max_steps > KCBS_MAX_STEPS)
        return fail(ctx, KCBS_ERR_INVALID_ARGUMENT, "n, n_parents or max_steps out of range");
    if (n == 0) return KCBS_OK;
    if (!parents || !controls || !goals) return fail(ctx, KCBS_ERR_INVALID_ARGUMENT, "parents / controls / goals is null");
    KCBS_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    const size_t b_par = sizeof(float) * 10 * (size_t)n_parents, b_ctl = sizeof(float) * 4 * (size_t)n;
    const size_t b_idx = sizeof(int32_t) * (size_t)n;
    const size_t b_seg = sizeof(float) * 10 * (size_t)max_steps * (size_t)n, b_fin = sizeof(float) * 10 * (size_t)n;
    int st;
    if ((st = reserve(ctx, ctx->in0, b_par)) || (st = reserve(ctx, ctx->in1, b_ctl)) ||
        (st = reserve(ctx, ctx->in2, parent_idx ? b_idx : 0)) || (st = reserve(ctx, ctx->in3, steps ? b_idx : 0)) ||
        (st = reserve(ctx, ctx->out0, seg_out ? b_seg : 0)) || (st = reserve(ctx, ctx->out1, final_out ? b_fin : 0)) ||
        (st = reserve(ctx, ctx->out2, b_idx)))
        return st;
    KCBS_CUDA(ctx, cudaMemcpyAsync(ctx->in0.p, parents, b_par, cudaMemcpyHostToDevice, s));
    KCBS_CUDA(ctx, cudaMemcpyAsync(ctx->in1.p, controls, b_ctl, cudaMemcpyHostToDevice, s));
    if (parent_idx) KCBS_CUDA(ctx, cudaMemcpyAsync(ctx->in2.p, parent_idx, b_idx, cudaMemcpyHostToDevice, s));
    if (steps) KCBS_CUDA(ctx, cudaMemcpyAsync(ctx->in3.p, steps, b_idx, cudaMemcpyHostToDevice, s));
    st = kcbs_propagate_pair_batch_dev(ctx, robot1, robot2, n, n_parents, (const float *)ctx->in0.p,
                                       parent_idx ? (const int32_t *)ctx->in2.p : nullptr, (const float *)ctx->in1.p,
                                       steps ? (const int32_t *)ctx->in3.p : nullptr, max_steps, goals, hold_radius,
                                       seg_out ? (float *)ctx->out0.p : nullptr, final_out ? (float *)ctx->out1.p : nullptr,
                                       (int32_t *)ctx->out2.p, s);
    if (st != KCBS_OK) return st;
    if (seg_out) KCBS_CUDA(ctx, cudaMemcpyAsync(seg_out, ctx->out0.p, b_seg, cudaMemcpyDeviceToHost, s));
    if (final_out) KCBS_CUDA(ctx, cudaMemcpyAsync(final_out, ctx->out1.p, b_fin, cudaMemcpyDeviceToHost, s));
    if (valid_steps) KCBS_CUDA(ctx, cudaMemcpyAsync(valid_steps, ctx->out2.p, b_idx, cudaMemcpyDeviceToHost, s));
    KCBS_CUDA(ctx, cudaStreamSynchronize(s));
    return KCBS_OK;
}

int kcbs_validity_pair_batch_dev(kcbs_ctx *ctx, int robot1, int robot2, int64_t n, const float *states, uint32_t *bitmask,
                                 void *stream)
{
    if (!ctx) return KCBS_ERR_INVALID_ARGUMENT;
    int st;
    if ((st = check_robot(ctx, robot1)) || (st = check_robot(ctx, robot2))) return st;
    if (n < 0) return fail(ctx, KCBS_ERR_INVALID_ARGUMENT, "n is negative");
    if (n == 0) return KCBS_OK;
    if (!states || !bitmask) return fail(ctx, KCBS_ERR_INVALID_ARGUMENT, "states / bitmask is null");
    PairValidArgs a;
    a.n = n; a.states = states; a.bitmask = bitmask; a.robot1 = robot1; a.robot2 = robot2;
    KCBS_CUDA(ctx, cudaSetDevice(ctx->device));
    KCBS_CUDA(ctx, launch_validity_pair(ctx->world, a, pick(ctx, stream)));
    ctx->launches += 1;
    return KCBS_OK;
}

int kcbs_validity_pair_batch(kcbs_ctx *ctx, int robot1, int robot2, int64_t n, const float *states, uint32_t *bitmask)
{
    if (!ctx) return KCBS_ERR_INVALID_ARGUMENT;
    if (n < 0) return fail(ctx, KCBS_ERR_INVALID_ARGUMENT, "n is negative");
    if (n == 0) return KCBS_OK;
    if (!states || !bitmask) return fail(ctx, KCBS_ERR_INVALID_ARGUMENT, "states / bitmask is null");
    KCBS_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    const size_t b_in = sizeof(float) * 10 * (size_t)n, b_out = sizeof(uint32_t) * (size_t)((n + 31) / 32);
    int st;
    if ((st = reserve(ctx, ctx->in0, b_in)) || (st = reserve(ctx, ctx->out0, b_out))) return st;
    KCBS_CUDA(ctx, cudaMemcpyAsync(ctx->in0.p, states, b_in, cudaMemcpyHostToDevice, s));
    st = kcbs_validity_pair_batch_dev(ctx, robot1, robot2, n, (const float *)ctx->in0.p, (uint32_t *)ctx->out0.p, s);
    if (st != KCBS_OK) return st;
    KCBS_CUDA(ctx, cudaMemcpyAsync(bitmask, ctx->out0.p, b_out, cudaMemcpyDeviceToHost, s));
    KCBS_CUDA(ctx, cudaStreamSynchronize(s));
    return KCBS_OK;
}

/* ---- measurement support ------------------------------------------------------------------- */

static int measure_probe(kcbs_ctx *ctx, double *tflops, bool packed)
{
    if (!ctx || !tflops) return KCBS_ERR_INVALID_ARGUMENT;
    KCBS_CUDA(ctx, cudaSetDevice(ctx->device));
    int st = reserve(ctx, ctx->out2, 256);
    if (st != KCBS_OK) return st;
    auto launch = packed ? launch_fp32x2_probe : launch_fp32_probe;
    double flops = 0, best = 0;
    KCBS_CUDA(ctx, launch((float *)ctx->out2.p, 64, ctx->stream, &flops));  // warm-up
    for (int rep = 0; rep < 5; ++rep) {
        KCBS_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
        KCBS_CUDA(ctx, launch((float *)ctx->out2.p, 2048, ctx->stream, &flops));
        KCBS_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
        KCBS_CUDA(ctx, cudaEventSynchronize(ctx->ev1));
        float ms = 0;
        KCBS_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        const double t = flops / (ms * 1e-3) * 1e-12;
        if (t > best) best = t;
    }
    ctx->launches += 6;
    *tflops = best;
    return KCBS_OK;
}

int kcbs_measure_fp32_peak(kcbs_ctx *ctx, double *tflops) { return measure_probe(ctx, tflops, false); }
int kcbs_measure_fp32x2_peak(kcbs_ctx *ctx, double *tflops) { return measure_probe(ctx, tflops, true); }

int64_t kcbs_launch_count(const kcbs_ctx *ctx) { return ctx ? ctx->launches : 0; }

}  // extern "C"
