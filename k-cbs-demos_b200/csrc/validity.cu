// validity.cu -- K2: batched homogeneous2ndOrderCarSystemSVC::isValid
// (StateValidityCheckerDatabase.h:54-75): state bounds (StateSpaceDatabase.h:49-59 through
// si_->satisfiesBounds, :57) and the rotated robot rectangle against the obstacle rectangles
// (:62-73, Obstacle.h:79-103).  HBM-bound: 20 B read per state, 1 bit written.
//
// Persistent CTAs stage the fine three-valued obstacle grid (32 KB of 16-bit cell codes, kcbs_device.cuh) and the
// obstacle table in shared memory once, then every warp streams passes of 128 states: one 16 B load per plane and
// lane (4 consecutive states), the bounds test, one shared-memory code lookup per state.  "free" and "hit" cells are
// decided by the lookup alone; the states of "maybe" cells (the band where the heading decides, ~10 % of a
// Congested10x10 batch) are pushed on a per-warp shared-memory queue (position, heading, cell code, place of the
// verdict bit) with warp prefix sums.  The queue OUTLIVES a pass: the exact rectangle test (sincos + 4 separating axes
// per candidate obstacle) runs only when 32 entries are waiting, one per lane, and clears the verdict bit of a colliding
// state with an atomic on the mask word the warp stored earlier -- so the most expensive code of the kernel always
// runs with every lane busy (it used to run once per pass with ~13 of 32 lanes).  Verdicts leave as one 16 B store per
// warp and pass.
#include "kcbs_device.cuh"
#include "kernels.h"

namespace kcbs {

constexpr int kValidBlock = 512;
constexpr int kValidWarps = kValidBlock / 32;
constexpr int kValidPass = 128;  // states per warp and pass (4 per lane)

constexpr int kValidQueue = kValidPass + 32;   // a pass adds at most 128 entries to the < 32 left over from the pass before

// States whose cell needs the exact test, per warp: position, heading, cell code | bit << 16, and the word of the verdict
// mask that holds the state's bit.  The queue outlives a pass: entries are tested 32 at a time, as soon as there are 32.
struct ValidQueue {
    float x[kValidQueue], y[kValidQueue], th[kValidQueue];
    uint32_t meta[kValidQueue], word[kValidQueue];
};

struct ValidShared {
    unsigned short code[kFineRows * kFinePitch];
    float4 obs[KCBS_MAX_OBSTACLES];
    ValidQueue queue[kValidWarps];
};
static_assert(sizeof(unsigned short) * kFineRows * kFinePitch % 16 == 0, "the grid is staged with 16 B copies");

// One pass of a warp: the bounds test (StateSpaceDatabase.h:49-59) of the lane's four states and, with obstacles, their
// cell codes; states of "maybe" cells are pushed on the warp's queue.  Branch free: a state outside the bounds looks up
// the cell of the workspace origin and ignores the answer.  Returns the lane's four verdict bits.
template <bool HAS_OBS, bool SYM, bool FULL>
__device__ __forceinline__ uint32_t classify4(const DevWorld &w, const unsigned short *codes, ValidQueue &queue, unsigned lt_mask,
                                              int live, long long i0, const float4 X, const float4 Y, const float4 V,
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
            const bool need = ok & (code - 1u < kCellMany);  // neither free (0) nor hit: valid unless the exact test says otherwise
            ok &= code != kCellHit;
            const unsigned bal = __ballot_sync(0xffffffffu, need);
            if (need) {
                const int at = qn + __popc(bal & lt_mask);
                const long long i = i0 + u;
                queue.x[at] = x[u];
                queue.y[at] = y[u];
                queue.th[at] = t[u];
                queue.meta[at] = code | (unsigned)(i & 31) << 16;
                queue.word[at] = (uint32_t)(i >> 5);
            }
            qn += __popc(bal);
        }
        nib |= (ok ? 1u : 0u) << u;
    }
    return nib;
}

// The exact rectangle test (StateValidityCheckerDatabase.h:62-73,180-205) of queue entry e; a collision clears the state's
// bit in the verdict mask (the word was stored earlier by this warp: ordered by the __syncwarp in between).
__device__ __forceinline__ void exact_test(const DevWorld &w, const ValidShared &sh, const ValidQueue &queue, int e, float hl, float hw,
                                           uint32_t *bitmask)
{
    const unsigned meta = queue.meta[e];
    if (!obstacles_clear_code(w, meta & 0xffffu, sh.obs, queue.x[e], queue.y[e], queue.th[e], hl, hw))
        atomicAnd(bitmask + queue.word[e], ~(1u << (meta >> 16)));
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
    ValidQueue &queue = sh.queue[wib];
    int qn = 0;   // entries waiting for the exact test (warp uniform)

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
        const uint32_t nib = full ? classify4<HAS_OBS, SYM, true>(w, sh.code, queue, lt_mask, live, i0, X, Y, V, P, T, qn)
                                  : classify4<HAS_OBS, SYM, false>(w, sh.code, queue, lt_mask, live, i0, X, Y, V, P, T, qn);
        // lanes 8k .. 8k+7 hold the 32 verdicts of mask word k
        uint32_t word = nib << (4 * (lane & 7));
        word |= __shfl_xor_sync(0xffffffffu, word, 1);
        word |= __shfl_xor_sync(0xffffffffu, word, 2);
        word |= __shfl_xor_sync(0xffffffffu, word, 4);
        const uint32_t out = __shfl_sync(0xffffffffu, word, (lane & 3) * 8);
        if (lane < 4 && (pass * 4 + lane) * 32 < n) a.bitmask[pass * 4 + lane] = out;
        if (HAS_OBS) {
            // the exact test runs with every lane busy: 32 entries at a time, taken from the end of the queue (the rest
            // waits for the next pass; a Congested10x10 pass adds ~13)
            __syncwarp();
            while (qn >= 32) {
                qn -= 32;
                exact_test(w, sh, queue, qn + lane, hl, hw, a.bitmask);
            }
            __syncwarp();
        }
    }
    if (HAS_OBS && lane < qn) exact_test(w, sh, queue, lane, hl, hw, a.bitmask);
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
