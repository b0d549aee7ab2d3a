// validity.cu -- K2: batched homogeneous2ndOrderCarSystemSVC::isValid
// (StateValidityCheckerDatabase.h:54-75): state bounds (StateSpaceDatabase.h:49-59 through
// si_->satisfiesBounds, :57) and the rotated robot rectangle against the obstacle rectangles
// (:62-73, Obstacle.h:79-103).  HBM-bound: 20 B read per state, 1 bit written.
//
// Persistent CTAs stage the obstacle table (float4 per rectangle) and the 8 KB cull grid in shared memory once,
// then every warp streams 128-state passes (four 32-bit mask words):
//   1. 20 plane loads per thread in flight before first use; bounds = 10 float compares against exact thresholds;
//   2. the state's grid cell names the few obstacles that can touch it; (state, obstacle) candidates of the whole
//      pass are compacted into a per-warp shared-memory queue (warp prefix sums), so that
//   3. the exact 4-axis SAT runs with all 32 lanes busy on real candidates instead of diverging per state;
//      colliding states are cleared with a shared-memory atomicOr;
//   4. verdicts leave as one coalesced 16 B store per warp.
#include "kcbs_device.cuh"
#include "kernels.h"

namespace kcbs {

constexpr int kValidUnroll = 4;     // mask words (32 states each) per warp pass
constexpr int kValidBlock = 256;
constexpr int kValidWarps = kValidBlock / 32;
constexpr int kPassStates = 32 * kValidUnroll;
constexpr int kQueueCap = 512;      // (state, obstacle) candidates per warp pass; overflow is tested inline

// exact robot-rectangle vs obstacle-rectangle test given cos/sin of the heading: true = separated
__device__ __forceinline__ bool obstacle_separated(float x, float y, float c, float s, float hl, float hw, float4 ob)
{
    const float ac = fabsf(c), as = fabsf(s);
    const float dx = ob.x - x, dy = ob.y - y;
    return (fabsf(dx) > ob.z + fmaf(hl, ac, hw * as)) | (fabsf(dy) > ob.w + fmaf(hl, as, hw * ac)) |
           (fabsf(fmaf(dx, c, -dy * s)) > hl + fmaf(ob.z, ac, ob.w * as)) |
           (fabsf(fmaf(dx, s, dy * c)) > hw + fmaf(ob.z, as, ob.w * ac));
}

template <bool HAS_OBS>
__global__ void __launch_bounds__(kValidBlock)
k_validity(const __grid_constant__ DevWorld w, const __grid_constant__ ValidArgs a)
{
    __shared__ unsigned long long s_grid[HAS_OBS ? kGridDim * kGridDim : 1];
    __shared__ float4 s_obs[HAS_OBS ? KCBS_MAX_OBSTACLES : 1];
    __shared__ float s_state[HAS_OBS ? kValidWarps : 1][3][HAS_OBS ? kPassStates : 1];
    __shared__ unsigned short s_queue[HAS_OBS ? kValidWarps : 1][HAS_OBS ? kQueueCap : 1];
    __shared__ unsigned int s_bad[HAS_OBS ? kValidWarps : 1][kValidUnroll];
    if (HAS_OBS) {
        for (int t = threadIdx.x; t < kGridDim * kGridDim; t += kValidBlock) s_grid[t] = w.grid[t];
        if (threadIdx.x < KCBS_MAX_OBSTACLES) s_obs[threadIdx.x] = w.obs4[threadIdx.x];
        __syncthreads();
    }
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
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
        int total = 0;  // candidates queued by this warp in this pass (warp-uniform)
        if (HAS_OBS) {
            if (lane < kValidUnroll) s_bad[wib][lane] = 0;
            __syncwarp();
        }
#pragma unroll
        for (int u = 0; u < kValidUnroll; ++u) {
            const long long i = (word0 + u) * 32 + lane;
            const bool ok = (i < n) && in_bounds(w, x[u], y[u], v[u], p[u], t[u]);
            const uint32_t m = __ballot_sync(0xffffffffu, ok);
            if (lane == u) mine = m;
            if (HAS_OBS) {
                unsigned long long cand = 0ull;
                if (ok) {
                    const int ix = min(w.gdim - 1, max(0, (int)((x[u] - w.gx0) * w.ginvx)));
                    const int iy = min(w.gdim - 1, max(0, (int)((y[u] - w.gy0) * w.ginvy)));
                    cand = s_grid[iy * w.gdim + ix];
                }
                const int slot = u * 32 + lane;
                const int cnt = __popcll(cand);
                // exclusive prefix sum of cnt over the warp
                int incl = cnt;
#pragma unroll
                for (int off = 1; off < 32; off <<= 1) {
                    const int o = __shfl_up_sync(0xffffffffu, incl, off);
                    if (lane >= off) incl += o;
                }
                int pos = total + incl - cnt;
                total += __shfl_sync(0xffffffffu, incl, 31);
                if (cand) {
                    s_state[wib][0][slot] = x[u];
                    s_state[wib][1][slot] = y[u];
                    s_state[wib][2][slot] = t[u];
                    float c = 0.f, s = 0.f;
                    bool have = false;
                    while (cand) {
                        const int o = __ffsll((long long)cand) - 1;
                        cand &= cand - 1;
                        if (pos < kQueueCap) {
                            s_queue[wib][pos] = (unsigned short)((slot << 6) | o);
                        } else {  // queue full (pathological scenes): test inline
                            if (!have) {
                                sincosf(t[u], &s, &c);
                                have = true;
                            }
                            if (!obstacle_separated(x[u], y[u], c, s, hl, hw, s_obs[o])) atomicOr(&s_bad[wib][u], 1u << lane);
                        }
                        ++pos;
                    }
                }
            }
        }
        if (HAS_OBS) {
            __syncwarp();
            const int q_n = min(total, kQueueCap);
            for (int e = lane; e < q_n; e += 32) {
                const unsigned int ent = s_queue[wib][e];
                const int slot = ent >> 6, o = ent & 63;
                float s, c;
                sincosf(s_state[wib][2][slot], &s, &c);
                if (!obstacle_separated(s_state[wib][0][slot], s_state[wib][1][slot], c, s, hl, hw, s_obs[o]))
                    atomicOr(&s_bad[wib][slot >> 5], 1u << (slot & 31));
            }
            __syncwarp();
            if (lane < kValidUnroll) mine &= ~s_bad[wib][lane];
            __syncwarp();
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
