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
    unsigned short code[kFineRows * kFinePitch];
    float4 obs[KCBS_MAX_OBSTACLES];
    float st[kValidWarps][3][kValidPass];    // x, y, theta of the warp's current pass (read back by the exact test)
    uint32_t queue[kValidWarps][kValidPass]; // cell code | slot << 16 of the states that need the exact test
    uint32_t res[kValidWarps][4];
};
static_assert(sizeof(unsigned short) * kFineRows * kFinePitch % 16 == 0, "the grid is staged with 16 B copies");

// One pass of a warp: the bounds test (StateSpaceDatabase.h:49-59) of the lane's four states and, with obstacles, their
// cell codes; states of "maybe" cells are pushed on the warp's queue.  Branch free: a state outside the bounds looks up
// the cell of the workspace origin and ignores the answer.  Returns the lane's four verdict bits.
template <bool HAS_OBS, bool SYM, bool FULL>
__device__ __forceinline__ uint32_t classify4(const DevWorld &w, const unsigned short *codes, uint32_t *queue, int lane,
                                              unsigned lt_mask, int live, const float4 X, const float4 Y, const float4 V,
                                              const float4 P, const float4 T, int &qn)
{
    const float x[4] = {X.x, X.y, X.z, X.w}, y[4] = {Y.x, Y.y, Y.z, Y.w}, v[4] = {V.x, V.y, V.z, V.w};
    const float p[4] = {P.x, P.y, P.z, P.w}, t[4] = {T.x, T.y, T.z, T.w};
    uint32_t nib = 0;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        bool ok = SYM ? in_bounds_sym(w, x[u], y[u], v[u], p[u], t[u]) : in_bounds(w, x[u], y[u], v[u], p[u], t[u]);
        if (!FULL) ok &= u < live;
        if (HAS_OBS) {
            const float xc = ok ? x[u] : w.gx0, yc = ok ? y[u] : w.gy0;
            // in-bounds coordinates index [0, kFineDim] (the upper bound itself lands on the repeated last row / column)
            const int ix = (int)fmaf(xc, w.finvx, w.foffx), iy = (int)fmaf(yc, w.finvy, w.foffy);
            const unsigned code = codes[iy * kFinePitch + ix];
            const bool need = ok & (code - 1u < kCellMany);  // neither free (0) nor hit
            ok &= code != kCellHit;
            const unsigned bal = __ballot_sync(0xffffffffu, need);
            if (need) queue[qn + __popc(bal & lt_mask)] = code | (unsigned)(lane * 4 + u) << 16;
            qn += __popc(bal);
        }
        nib |= (ok ? 1u : 0u) << u;
    }
    return nib;
}

template <bool HAS_OBS, bool VEC, bool SYM>
__global__ void __launch_bounds__(kValidBlock, 2)
k_validity(const __grid_constant__ DevWorld w, const __grid_constant__ ValidArgs a)
{
    extern __shared__ __align__(16) unsigned char s_raw[];
    ValidShared &sh = *reinterpret_cast<ValidShared *>(s_raw);
    if (HAS_OBS) {
        const uint4 *src = reinterpret_cast<const uint4 *>(w.fine);
        uint4 *dst = reinterpret_cast<uint4 *>(sh.code);
        for (int t = threadIdx.x; t < kFineRows * kFinePitch / 8; t += kValidBlock) dst[t] = src[t];
        if (threadIdx.x < KCBS_MAX_OBSTACLES) sh.obs[threadIdx.x] = w.obs4[threadIdx.x];
        __syncthreads();
    }
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    const long long n = a.n;
    const long long n_pass = (n + kValidPass - 1) / kValidPass;
    const float hl = w.rob_hl[a.robot], hw = w.rob_hw[a.robot];
    const long long warps_total = (long long)gridDim.x * kValidWarps;
    const float *px = a.states, *py = a.states + n, *pv = a.states + 2 * n, *pp = a.states + 3 * n, *pt = a.states + 4 * n;
    float *sx = sh.st[wib][0], *sy = sh.st[wib][1], *sth = sh.st[wib][2];
    uint32_t *queue = sh.queue[wib], *res = sh.res[wib];

    for (long long pass = (long long)blockIdx.x * kValidWarps + wib; pass < n_pass; pass += warps_total) {
        const long long i0 = pass * kValidPass + lane * 4;  // this lane's four consecutive states
        const bool full = (pass + 1) * kValidPass <= n;      // warp uniform
        const int live = full ? 4 : (int)max(0ll, min(4ll, n - i0));
        float4 X, Y, V, P, T;
        if (VEC && full) {
            X = __ldcs(reinterpret_cast<const float4 *>(px + i0));
            Y = __ldcs(reinterpret_cast<const float4 *>(py + i0));
            V = __ldcs(reinterpret_cast<const float4 *>(pv + i0));
            P = __ldcs(reinterpret_cast<const float4 *>(pp + i0));
            T = __ldcs(reinterpret_cast<const float4 *>(pt + i0));
        } else {
            float e[5][4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const long long ii = u < live ? i0 + u : 0;
                e[0][u] = __ldcs(px + ii);
                e[1][u] = __ldcs(py + ii);
                e[2][u] = __ldcs(pv + ii);
                e[3][u] = __ldcs(pp + ii);
                e[4][u] = __ldcs(pt + ii);
            }
            X = make_float4(e[0][0], e[0][1], e[0][2], e[0][3]);
            Y = make_float4(e[1][0], e[1][1], e[1][2], e[1][3]);
            V = make_float4(e[2][0], e[2][1], e[2][2], e[2][3]);
            P = make_float4(e[3][0], e[3][1], e[3][2], e[3][3]);
            T = make_float4(e[4][0], e[4][1], e[4][2], e[4][3]);
        }
        if (HAS_OBS) {
            reinterpret_cast<float4 *>(sx)[lane] = X;
            reinterpret_cast<float4 *>(sy)[lane] = Y;
            reinterpret_cast<float4 *>(sth)[lane] = T;
        }
        int qn = 0;
        const uint32_t nib = full ? classify4<HAS_OBS, SYM, true>(w, sh.code, queue, lane, lt_mask, live, X, Y, V, P, T, qn)
                                  : classify4<HAS_OBS, SYM, false>(w, sh.code, queue, lane, lt_mask, live, X, Y, V, P, T, qn);
        // lanes 8k .. 8k+7 hold the 32 verdicts of mask word k
        uint32_t word = nib << (4 * (lane & 7));
        word |= __shfl_xor_sync(0xffffffffu, word, 1);
        word |= __shfl_xor_sync(0xffffffffu, word, 2);
        word |= __shfl_xor_sync(0xffffffffu, word, 4);
        uint32_t out;
        if (HAS_OBS && qn > 0) {
            if ((lane & 7) == 0) res[lane >> 3] = word;
            __syncwarp();
#pragma unroll 1
            for (int q = lane; q < qn; q += 32) {
                const unsigned tag = queue[q], slot = tag >> 16;
                if (!obstacles_clear_code(w, tag & 0xffffu, sh.obs, sx[slot], sy[slot], sth[slot], hl, hw))
                    atomicAnd(&res[slot >> 5], ~(1u << (slot & 31)));
            }
            __syncwarp();
            out = res[lane & 3];
            __syncwarp();
        } else {
            out = __shfl_sync(0xffffffffu, word, (lane & 3) * 8);
        }
        if (lane < 4 && (pass * 4 + lane) * 32 < n) a.bitmask[pass * 4 + lane] = out;
    }
}

template <bool HAS_OBS>
static cudaError_t launch_validity_t(const DevWorld &w, const ValidArgs &a, unsigned grid, bool vec, cudaStream_t stream)
{
    const size_t smem = HAS_OBS ? sizeof(ValidShared) : 0;
    if (vec) {
        if (w.sym_bounds)
            k_validity<HAS_OBS, true, true><<<grid, kValidBlock, smem, stream>>>(w, a);
        else
            k_validity<HAS_OBS, true, false><<<grid, kValidBlock, smem, stream>>>(w, a);
    } else {
        if (w.sym_bounds)
            k_validity<HAS_OBS, false, true><<<grid, kValidBlock, smem, stream>>>(w, a);
        else
            k_validity<HAS_OBS, false, false><<<grid, kValidBlock, smem, stream>>>(w, a);
    }
    return cudaGetLastError();
}

cudaError_t init_validity_kernels()
{
    const int bytes = (int)sizeof(ValidShared);
    cudaError_t e;
    if ((e = cudaFuncSetAttribute(k_validity<true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(k_validity<true, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(k_validity<true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes)) != cudaSuccess) return e;
    return cudaFuncSetAttribute(k_validity<true, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
}

cudaError_t launch_validity(const DevWorld &w, const ValidArgs &a, cudaStream_t stream)
{
    if (a.n <= 0) return cudaSuccess;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long per_block = (long long)kValidBlock * 4;
    const long long need = (a.n + per_block - 1) / per_block;
    const long long resident = (long long)sms * 2;  // persistent CTAs
    const unsigned grid = (unsigned)(need < resident ? need : resident);
    // 16 B loads need 16 B aligned planes: base aligned and n a multiple of 4
    const bool vec = (a.n % 4 == 0) && ((uintptr_t)a.states % 16 == 0);
    return w.n_obs > 0 ? launch_validity_t<true>(w, a, grid, vec, stream) : launch_validity_t<false>(w, a, grid, vec, stream);
}

}  // namespace kcbs
