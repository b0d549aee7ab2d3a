// validity.cu -- K2: batched homogeneous2ndOrderCarSystemSVC::isValid
// (StateValidityCheckerDatabase.h:54-75): state bounds (StateSpaceDatabase.h:49-59 through
// si_->satisfiesBounds, :57) and the rotated robot rectangle against the obstacle rectangles
// (:62-73, Obstacle.h:79-103).  HBM-bound: 20 B read per state, 1 bit written.
//
// Persistent CTAs stage the obstacle table (float4 per rectangle) and the 8 KB cull grid in shared
// memory once, then stream 1024-state chunks: each warp produces four 32-bit mask words per pass with
// all 20 plane loads of a thread in flight before the first use; per state only the obstacles flagged
// for its grid cell get the exact 4-axis SAT; verdicts leave through __ballot_sync as one coalesced
// 16 B store per warp.
#include "kcbs_device.cuh"
#include "kernels.h"

namespace kcbs {

constexpr int kValidUnroll = 4;
constexpr int kValidBlock = 256;

template <bool HAS_OBS>
__global__ void __launch_bounds__(kValidBlock)
k_validity(const __grid_constant__ DevWorld w, const __grid_constant__ ValidArgs a)
{
    __shared__ unsigned long long s_grid[HAS_OBS ? kGridDim * kGridDim : 1];
    __shared__ float4 s_obs[HAS_OBS ? KCBS_MAX_OBSTACLES : 1];
    if (HAS_OBS) {
        for (int t = threadIdx.x; t < kGridDim * kGridDim; t += kValidBlock) s_grid[t] = w.grid[t];
        if (threadIdx.x < KCBS_MAX_OBSTACLES) s_obs[threadIdx.x] = w.obs4[threadIdx.x];
        __syncthreads();
    }
    const int lane = threadIdx.x & 31;
    const long long n = a.n;
    const long long n_words = (n + 31) >> 5;
    const float hl = w.rob_hl[a.robot], hw = w.rob_hw[a.robot];
    const long long warps_total = ((long long)gridDim.x * kValidBlock) >> 5;

    for (long long warp = ((long long)blockIdx.x * kValidBlock + threadIdx.x) >> 5; warp * kValidUnroll < n_words;
         warp += warps_total) {
        const long long word0 = warp * kValidUnroll;
        float x[kValidUnroll], y[kValidUnroll], v[kValidUnroll], p[kValidUnroll], t[kValidUnroll];
#pragma unroll
        for (int u = 0; u < kValidUnroll; ++u) {
            const long long i = (word0 + u) * 32 + lane;
            const long long ii = i < n ? i : 0;
            x[u] = __ldcs(a.states + ii);
            y[u] = __ldcs(a.states + n + ii);
            v[u] = __ldcs(a.states + 2 * n + ii);
            p[u] = __ldcs(a.states + 3 * n + ii);
            t[u] = __ldcs(a.states + 4 * n + ii);
        }
        uint32_t mine = 0;
#pragma unroll
        for (int u = 0; u < kValidUnroll; ++u) {
            const long long i = (word0 + u) * 32 + lane;
            bool ok = (i < n) && in_bounds(w, x[u], y[u], v[u], p[u], t[u]);
            if (HAS_OBS) {
                if (ok) ok = obstacles_clear_grid(w, s_grid, s_obs, x[u], y[u], t[u], hl, hw);
            }
            const uint32_t m = __ballot_sync(0xffffffffu, ok);
            if (lane == u) mine = m;
        }
        if (lane < kValidUnroll && word0 + lane < n_words) a.bitmask[word0 + lane] = mine;
    }
}

cudaError_t launch_validity(const DevWorld &w, const ValidArgs &a, cudaStream_t stream)
{
    if (a.n <= 0) return cudaSuccess;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long per_block = (long long)kValidBlock * kValidUnroll;
    const long long need = (a.n + per_block - 1) / per_block;
    const long long grid = need < (long long)sms * 8 ? need : (long long)sms * 8;  // persistent: 8 CTAs per SM
    if (w.n_obs > 0)
        k_validity<true><<<(unsigned)grid, kValidBlock, 0, stream>>>(w, a);
    else
        k_validity<false><<<(unsigned)grid, kValidBlock, 0, stream>>>(w, a);
    return cudaGetLastError();
}

}  // namespace kcbs
