// conflict.cu -- K3: first robot-robot conflict of candidate plans.
//
// Replaces the scan of KCBS::findConflicts (OMPL K-CBS fork; call sites demo_*.cpp:155,168 and
// Benchmark.h:151-158) over homogeneous2ndOrderCarSystemSVC::areStatesValid
// (StateValidityCheckerDatabase.h:78-125, non-composed branch :108-124):
//   for k ascending, r1 ascending, r2 > r1 ascending: first pair whose rectangles (each rotated by
//   -theta, :152-168) are not disjoint (:173) is the conflict; it is extended while it persists.
//
// Mapping: one CTA = one plan x one tile of kTile time indices; one thread = one time index with
// every robot's (x, y) of that index parked in a private shared-memory column (coalesced plane
// reads, no cross-thread sharing, no __syncthreads).  Robots are processed as register tiles of 8
// ("A") against the remaining robots streamed from shared memory ("B"), tracking only the minimum
// squared centre distance: the reference's circle test (:134-141) can only clear pairs that the
// exact test would clear as well, so the rectangle SAT runs only for threads whose minimum falls
// inside the bounding circles.  Heading planes are read only there.  The lexicographic minimum
// (k, r1, r2) is a warp shuffle-min followed by one 64-bit atomicMin per warp.
#include "kcbs_device.cuh"
#include "kernels.h"

namespace kcbs {

constexpr int kTile = 128;   // time indices (threads) per CTA
constexpr int kATile = 8;    // robots held in registers

__device__ __forceinline__ int clamp_k(int k, int T, const int *len, int r)
{
    int kk = min(k, T - 1);
    if (len) kk = min(kk, len[r] - 1);
    return max(kk, 0);
}

__global__ void __launch_bounds__(kTile)
k_conflict_keys(const __grid_constant__ DevWorld w, const __grid_constant__ ConflictArgs a, int tiles_per_plan)
{
    extern __shared__ float smem[];
    const int R = a.R, T = a.T;
    const int plan = blockIdx.x / tiles_per_plan;
    const int tile = blockIdx.x - plan * tiles_per_plan;
    const int k0 = a.k_lo + tile * kTile;
    const int tid = threadIdx.x;
    const int k = k0 + tid;

    // a conflict already recorded at an earlier time index makes this tile irrelevant
    const unsigned long long seen = *(volatile const unsigned long long *)(a.keys + plan);
    if (seen != kNoConflict && (long long)(seen >> 16) < (long long)k0) return;

    const float *base = a.traj + (long long)plan * R * 3 * T;
    const int *len = a.lengths ? a.lengths + (long long)plan * R : nullptr;
    float *xs = smem;              // [R][kTile]
    float *ys = smem + R * kTile;  // [R][kTile]
    const bool active = k < a.k_hi;

#pragma unroll 4
    for (int r = 0; r < R; ++r) {
        const int kk = clamp_k(k, T, len, r);
        xs[r * kTile + tid] = __ldcs(base + ((long long)r * 3 + 0) * T + kk);
        ys[r * kTile + tid] = __ldcs(base + ((long long)r * 3 + 1) * T + kk);
    }

    float m = 3.0e38f;  // min squared centre distance over all pairs at this time index
    for (int a0 = 0; a0 < R; a0 += kATile) {
        float xa[kATile], ya[kATile];
#pragma unroll
        for (int q = 0; q < kATile; ++q) {
            const bool in = a0 + q < R;
            // distinct far-away sentinels for the missing robots of a partial tile
            xa[q] = in ? xs[(a0 + q) * kTile + tid] : 1.0e18f * (float)(q + 1);
            ya[q] = in ? ys[(a0 + q) * kTile + tid] : 0.0f;
        }
#pragma unroll
        for (int q = 0; q < kATile; ++q)
#pragma unroll
            for (int p = q + 1; p < kATile; ++p) {
                const float dx = xa[q] - xa[p], dy = ya[q] - ya[p];
                m = fminf(m, fmaf(dy, dy, dx * dx));
            }
        for (int r2 = a0 + kATile; r2 < R; ++r2) {
            const float xb = xs[r2 * kTile + tid], yb = ys[r2 * kTile + tid];
#pragma unroll
            for (int q = 0; q < kATile; ++q) {
                const float dx = xa[q] - xb, dy = ya[q] - yb;
                m = fminf(m, fmaf(dy, dy, dx * dx));
            }
        }
    }

    unsigned long long key = kNoConflict;
    const float lim = 2.0f * w.max_rad;
    if (active && m <= lim * lim * 1.000001f) {
        // rare: some pair is inside the bounding circles -> exact rectangle test in scan order
        for (int r1 = 0; r1 < R && key == kNoConflict; ++r1) {
            const float x1 = xs[r1 * kTile + tid], y1 = ys[r1 * kTile + tid];
            bool have1 = false;
            float c1 = 1.0f, s1 = 0.0f;
            for (int r2 = r1 + 1; r2 < R; ++r2) {
                const float x2 = xs[r2 * kTile + tid], y2 = ys[r2 * kTile + tid];
                const float dx = x2 - x1, dy = y2 - y1;
                const float rs = w.rob_rad[r1] + w.rob_rad[r2];
                if (fmaf(dy, dy, dx * dx) <= rs * rs * 1.000001f) {
                    if (!have1) {
                        sincosf(base[((long long)r1 * 3 + 2) * T + clamp_k(k, T, len, r1)], &s1, &c1);
                        have1 = true;
                    }
                    float c2, s2;
                    sincosf(base[((long long)r2 * 3 + 2) * T + clamp_k(k, T, len, r2)], &s2, &c2);
                    const float gap = sat_gap_robots(x1, y1, c1, s1, w.rob_hl[r1], w.rob_hw[r1], x2, y2, c2, s2,
                                                     w.rob_hl[r2], w.rob_hw[r2]);
                    if (gap <= 0.0f) {
                        key = pack_key(k, r1, r2);
                        break;
                    }
                }
            }
        }
    }

    if (__any_sync(0xffffffffu, key != kNoConflict)) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const unsigned long long o = __shfl_xor_sync(0xffffffffu, key, off);
            key = o < key ? o : key;
        }
        if ((tid & 31) == 0) atomicMin(a.keys + plan, key);
    }
}

// Pair test at one time index, used by the k_end extension.
__device__ __forceinline__ bool pair_collides(const DevWorld &w, const float *base, const int *len, int T, int r1,
                                              int r2, int k)
{
    const int k1 = clamp_k(k, T, len, r1), k2 = clamp_k(k, T, len, r2);
    const float x1 = base[((long long)r1 * 3 + 0) * T + k1], y1 = base[((long long)r1 * 3 + 1) * T + k1];
    const float x2 = base[((long long)r2 * 3 + 0) * T + k2], y2 = base[((long long)r2 * 3 + 1) * T + k2];
    const float dx = x2 - x1, dy = y2 - y1;
    const float rs = w.rob_rad[r1] + w.rob_rad[r2];
    if (fmaf(dy, dy, dx * dx) > rs * rs * 1.000001f) return false;
    float c1, s1, c2, s2;
    sincosf(base[((long long)r1 * 3 + 2) * T + k1], &s1, &c1);
    sincosf(base[((long long)r2 * 3 + 2) * T + k2], &s2, &c2);
    return sat_gap_robots(x1, y1, c1, s1, w.rob_hl[r1], w.rob_hw[r1], x2, y2, c2, s2, w.rob_hl[r2], w.rob_hw[r2]) <=
           0.0f;
}

// One warp per plan: unpack the key and extend the conflict while the same pair keeps colliding.
__global__ void __launch_bounds__(128)
k_conflict_finish(const __grid_constant__ DevWorld w, const __grid_constant__ FinishArgs a)
{
    const int plan = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (plan >= a.P) return;
    const unsigned long long key = a.keys[plan];
    kcbs_conflict c;
    if (key == kNoConflict) {
        c.r1 = c.r2 = c.k_start = c.k_end = -1;
    } else {
        const int T = a.T;
        const float *base = a.traj + (long long)plan * a.R * 3 * T;
        const int *len = a.lengths ? a.lengths + (long long)plan * a.R : nullptr;
        c.k_start = (int)(key >> 16);
        c.r1 = (int)((key >> 8) & 0xff);
        c.r2 = (int)(key & 0xff);
        int k_end = c.k_start;
        for (int b = c.k_start + 1; b < T; b += 32) {
            const int k = b + lane;
            const bool hit = (k < T) && pair_collides(w, base, len, T, c.r1, c.r2, k);
            const unsigned miss = ~__ballot_sync(0xffffffffu, hit);
            if (miss) {
                k_end = b + __ffs(miss) - 2;  // last hit before the first miss
                break;
            }
            k_end = b + 31;
        }
        c.k_end = min(k_end, T - 1);
    }
    if (lane == 0) a.out[plan] = c;
}

__global__ void k_fill_keys(unsigned long long *keys, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) keys[i] = kNoConflict;
}

cudaError_t launch_fill_keys(unsigned long long *keys, int n, cudaStream_t stream)
{
    if (n <= 0) return cudaSuccess;
    k_fill_keys<<<(n + 255) / 256, 256, 0, stream>>>(keys, n);
    return cudaGetLastError();
}

cudaError_t launch_conflict_keys(const DevWorld &w, const ConflictArgs &a, cudaStream_t stream)
{
    if (a.P <= 0 || a.k_hi <= a.k_lo) return cudaSuccess;
    const int tiles = (a.k_hi - a.k_lo + kTile - 1) / kTile;
    const size_t smem = (size_t)2 * a.R * kTile * sizeof(float);
    k_conflict_keys<<<(unsigned)((long long)tiles * a.P), kTile, smem, stream>>>(w, a, tiles);
    return cudaGetLastError();
}

cudaError_t launch_conflict_finish(const DevWorld &w, const FinishArgs &a, cudaStream_t stream)
{
    if (a.P <= 0) return cudaSuccess;
    const int warps_per_block = 4;
    k_conflict_finish<<<(a.P + warps_per_block - 1) / warps_per_block, 128, 0, stream>>>(w, a);
    return cudaGetLastError();
}

int conflict_keys_launches() { return 1; }

cudaError_t init_conflict_kernels()
{
    return cudaFuncSetAttribute(k_conflict_keys, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                2 * KCBS_MAX_ROBOTS * kTile * (int)sizeof(float));
}

}  // namespace kcbs
