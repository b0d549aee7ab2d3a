// validity.cu -- K2: batched homogeneous2ndOrderCarSystemSVC::isValid
// (StateValidityCheckerDatabase.h:54-75): state bounds (StateSpaceDatabase.h:49-59 through
// si_->satisfiesBounds, :57) and the rotated robot rectangle against the obstacle rectangles
// (:62-73, Obstacle.h:79-103).  HBM-bound: 20 B read per state, 1 bit written.
//
// Persistent CTAs stage the fine three-valued obstacle grid (32 KB of 16-bit cell codes, kcbs_device.cuh) and the
// obstacle table in shared memory once, then every warp streams passes of 128 states: one 16 B load per plane and
// lane (4 consecutive states), the bounds test, one shared-memory code lookup per state.  "free" and "hit" cells are
// decided by the lookup alone; the states of "maybe" cells (the band where the heading decides, ~10 % of a
// Congested10x10 batch) are compacted into a per-warp shared-memory queue with warp prefix sums, so the exact
// rectangle test (sincos + 4 separating axes per candidate obstacle) runs once per pass with all lanes busy instead
// of once per state slot with a few lanes active.  Verdicts leave as one 16 B store per warp and pass.
#include "kcbs_device.cuh"
#include "kernels.h"

namespace kcbs {

constexpr int kValidBlock = 512;
constexpr int kValidWarps = kValidBlock / 32;
constexpr int kValidPass = 128;  // states per warp and pass (4 per lane)

struct ValidShared {
    unsigned short code[kFineDim * kFineDim];
    float4 obs[KCBS_MAX_OBSTACLES];
    float4 queue[kValidWarps][kValidPass];  // (x, y, theta, code | slot << 16) of the states that need the exact test
    uint32_t res[kValidWarps][4];
};

template <bool HAS_OBS, bool VEC>
__global__ void __launch_bounds__(kValidBlock, 2)
k_validity(const __grid_constant__ DevWorld w, const __grid_constant__ ValidArgs a)
{
    extern __shared__ __align__(16) unsigned char s_raw[];
    ValidShared &sh = *reinterpret_cast<ValidShared *>(s_raw);
    if (HAS_OBS) {
        const uint4 *src = reinterpret_cast<const uint4 *>(w.fine);
        uint4 *dst = reinterpret_cast<uint4 *>(sh.code);
        for (int t = threadIdx.x; t < kFineDim * kFineDim / 8; t += kValidBlock) dst[t] = src[t];
        if (threadIdx.x < KCBS_MAX_OBSTACLES) sh.obs[threadIdx.x] = w.obs4[threadIdx.x];
        __syncthreads();
    }
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    const long long n = a.n;
    const long long n_words = (n + 31) >> 5;
    const float hl = w.rob_hl[a.robot], hw = w.rob_hw[a.robot];
    const long long warps_total = (long long)gridDim.x * kValidWarps;
    const float *px = a.states, *py = a.states + n, *pv = a.states + 2 * n, *pp = a.states + 3 * n, *pt = a.states + 4 * n;

    for (long long pass = (long long)blockIdx.x * kValidWarps + wib; pass * kValidPass < n; pass += warps_total) {
        const long long i0 = pass * kValidPass + lane * 4;  // this lane's four consecutive states
        float x[4], y[4], v[4], p[4], t[4];
        if (VEC && i0 + 3 < n) {
            const float4 X = __ldcs(reinterpret_cast<const float4 *>(px + i0));
            const float4 Y = __ldcs(reinterpret_cast<const float4 *>(py + i0));
            const float4 V = __ldcs(reinterpret_cast<const float4 *>(pv + i0));
            const float4 P = __ldcs(reinterpret_cast<const float4 *>(pp + i0));
            const float4 T = __ldcs(reinterpret_cast<const float4 *>(pt + i0));
            x[0] = X.x; x[1] = X.y; x[2] = X.z; x[3] = X.w;
            y[0] = Y.x; y[1] = Y.y; y[2] = Y.z; y[3] = Y.w;
            v[0] = V.x; v[1] = V.y; v[2] = V.z; v[3] = V.w;
            p[0] = P.x; p[1] = P.y; p[2] = P.z; p[3] = P.w;
            t[0] = T.x; t[1] = T.y; t[2] = T.z; t[3] = T.w;
        } else {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const long long ii = i0 + u < n ? i0 + u : 0;
                x[u] = __ldcs(px + ii);
                y[u] = __ldcs(py + ii);
                v[u] = __ldcs(pv + ii);
                p[u] = __ldcs(pp + ii);
                t[u] = __ldcs(pt + ii);
            }
        }
        uint32_t nib = 0;
        int qn = 0;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            bool ok = (i0 + u < n) && in_bounds(w, x[u], y[u], v[u], p[u], t[u]);
            if (HAS_OBS) {
                const unsigned code = ok ? (unsigned)sh.code[fine_cell(w, x[u], y[u])] : 0u;
                ok = ok && code != kCellHit;
                const bool need = ok && code != 0u;
                const unsigned bal = __ballot_sync(0xffffffffu, need);
                if (need)
                    sh.queue[wib][qn + __popc(bal & lt_mask)] =
                        make_float4(x[u], y[u], t[u], __uint_as_float(code | (unsigned)(lane * 4 + u) << 16));
                qn += __popc(bal);
            }
            nib |= (ok ? 1u : 0u) << u;
        }
        // lanes 8k .. 8k+7 hold the 32 verdicts of mask word k
        uint32_t word = nib << (4 * (lane & 7));
        word |= __shfl_xor_sync(0xffffffffu, word, 1);
        word |= __shfl_xor_sync(0xffffffffu, word, 2);
        word |= __shfl_xor_sync(0xffffffffu, word, 4);
        uint32_t out;
        if (HAS_OBS && qn > 0) {
            if ((lane & 7) == 0) sh.res[wib][lane >> 3] = word;
            __syncwarp();
            for (int q = lane; q < qn; q += 32) {
                const float4 e = sh.queue[wib][q];
                const unsigned tag = __float_as_uint(e.w), slot = tag >> 16;
                if (!obstacles_clear_code(w, tag & 0xffffu, sh.obs, e.x, e.y, e.z, hl, hw))
                    atomicAnd(&sh.res[wib][slot >> 5], ~(1u << (slot & 31)));
            }
            __syncwarp();
            out = sh.res[wib][lane & 3];
            __syncwarp();
        } else {
            out = __shfl_sync(0xffffffffu, word, (lane & 3) * 8);
        }
        if (lane < 4 && pass * 4 + lane < n_words) a.bitmask[pass * 4 + lane] = out;
    }
}

cudaError_t init_validity_kernels()
{
    cudaError_t e = cudaFuncSetAttribute(k_validity<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ValidShared));
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(k_validity<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ValidShared));
}

cudaError_t launch_validity(const DevWorld &w, const ValidArgs &a, cudaStream_t stream)
{
    if (a.n <= 0) return cudaSuccess;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const bool obs = w.n_obs > 0;
    const long long per_block = (long long)kValidBlock * 4;
    const long long need = (a.n + per_block - 1) / per_block;
    const long long resident = (long long)sms * 2;  // persistent CTAs
    const unsigned grid = (unsigned)(need < resident ? need : resident);
    // 16 B loads need 16 B aligned planes: base aligned and n a multiple of 4
    const bool vec = (a.n % 4 == 0) && ((uintptr_t)a.states % 16 == 0);
    if (obs) {
        if (vec)
            k_validity<true, true><<<grid, kValidBlock, sizeof(ValidShared), stream>>>(w, a);
        else
            k_validity<true, false><<<grid, kValidBlock, sizeof(ValidShared), stream>>>(w, a);
    } else {
        if (vec)
            k_validity<false, true><<<grid, kValidBlock, 0, stream>>>(w, a);
        else
            k_validity<false, false><<<grid, kValidBlock, 0, stream>>>(w, a);
    }
    return cudaGetLastError();
}

}  // namespace kcbs
