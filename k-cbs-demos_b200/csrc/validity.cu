// validity.cu -- K2: batched homogeneous2ndOrderCarSystemSVC::isValid
// (StateValidityCheckerDatabase.h:54-75): state bounds (StateSpaceDatabase.h:49-59 through
// si_->satisfiesBounds, :57) and the rotated robot rectangle against the obstacle rectangles
// (:62-73, Obstacle.h:79-103).  HBM-bound: 20 B read per state, 1 bit written.
//
// One warp produces four 32-bit mask words per pass: all 20 plane loads of a thread are issued
// before the first use, the obstacle table is read from the constant bank (warp-uniform index),
// and the verdicts leave through __ballot_sync -> one coalesced 16 B store per warp.
#include "kcbs_device.cuh"
#include "kernels.h"

namespace kcbs {

constexpr int kValidUnroll = 4;

template <bool HAS_OBS>
__global__ void __launch_bounds__(256)
k_validity(const __grid_constant__ DevWorld w, const __grid_constant__ ValidArgs a)
{
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long word0 = warp * kValidUnroll;
    const long long n = a.n;
    const float hl = w.rob_hl[a.robot], hw = w.rob_hw[a.robot], rad = w.rob_rad[a.robot];

    float x[kValidUnroll], y[kValidUnroll], v[kValidUnroll], p[kValidUnroll], t[kValidUnroll];
#pragma unroll
    for (int u = 0; u < kValidUnroll; ++u) {
        const long long i = (word0 + u) * 32 + lane;
        const bool in = i < n;
        const long long ii = in ? i : 0;
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
            if (ok) ok = obstacles_clear(w, x[u], y[u], t[u], hl, hw, rad);
        }
        const uint32_t m = __ballot_sync(0xffffffffu, ok);
        if (lane == u) mine = m;
    }
    const long long n_words = (n + 31) >> 5;
    if (lane < kValidUnroll && word0 + lane < n_words) a.bitmask[word0 + lane] = mine;
}

cudaError_t launch_validity(const DevWorld &w, const ValidArgs &a, cudaStream_t stream)
{
    if (a.n <= 0) return cudaSuccess;
    const int block = 256;
    const long long per_block = (long long)block * kValidUnroll;
    const long long grid = (a.n + per_block - 1) / per_block;
    if (w.n_obs > 0)
        k_validity<true><<<(unsigned)grid, block, 0, stream>>>(w, a);
    else
        k_validity<false><<<(unsigned)grid, block, 0, stream>>>(w, a);
    return cudaGetLastError();
}

}  // namespace kcbs
