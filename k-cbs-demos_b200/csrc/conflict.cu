// conflict.cu -- K3: first robot-robot conflict of candidate plans.
//
// Replaces the scan of KCBS::findConflicts (OMPL K-CBS fork; call sites demo_*.cpp:155,168 and
// Benchmark.h:151-158) over homogeneous2ndOrderCarSystemSVC::areStatesValid
// (StateValidityCheckerDatabase.h:78-125, non-composed branch :108-124):
//   for k ascending, r1 ascending, r2 > r1 ascending: first pair whose rectangles (each rotated by
//   -theta, :152-168) are not disjoint (:173) is the conflict; it is extended while it persists.
//
// Mapping: one CTA = one plan x one tile of kTile time indices; one thread = one time index.
//   * Staging: the tile's 2R trajectory rows (x and y planes, 512 B each) are fetched by the TMA engine
//     (cp.async.bulk global -> shared, one instruction per row, completion on an mbarrier), so no
//     thread holds load registers and the whole tile is one memory round trip.  Tiles that touch the
//     padding of a short path, or misaligned plans (T % 4 != 0), take a per-element path.
//   * Broad phase: robots are processed as register tiles of 8 ("A") against the remaining robots
//     streamed from shared memory ("B").  Only min over pairs of max(|dx|, |dy|) is tracked: the
//     reference's circle test (:134-141) can only clear pairs that the exact test clears as well, and a
//     pair whose centres differ by more than r1 + r2 along an axis is outside the circle, so the verdict
//     is unchanged.  The subtractions use the packed add.rn.f32x2 (two pairs per instruction), the
//     max/min chain FMNMX / FMNMX3: 3 issue slots per pair.  Robots missing from the last register tile
//     are sentinel rows far away from everything.
//   * Narrow phase (rare): threads whose minimum is inside the bounding circles run the rectangle SAT in
//     scan order; heading planes are read only here.
//   * Reduction: the packed key (k, r1, r2) is min-reduced by warp shuffles and one 64-bit atomicMin
//     per warp; the last CTA of a plan (ticket counter) extends the conflict while the same pair keeps
//     colliding, writes the record and resets the plan's key and ticket for the next call.
//   * Sharded scan (a.world > 1, kcbs_first_conflict_sharded_dev): the CTAs of a rank cover its time slab only; the
//     last CTA of a plan then exchanges the rank's key with the peers THROUGH NVLINK PEER MEMORY from inside this
//     kernel (one tagged 64-bit store into every rank's mailbox, then a poll of its own mailbox), reduces and finishes:
//     the scan, the min all-reduce and the k_end extension are one launch, and the exchange of one plan overlaps with
//     the scan of the others.
//   * Tile size: 128 time indices per CTA for large batches; 64 when the grid would otherwise be a few waves only
//     (one plan = 79 CTAs of 128 on 148 SMs), so a single plan spreads over every SM and the tail of a small grid is short.
#include "kcbs_device.cuh"
#include "kernels.h"

namespace kcbs {

constexpr int kTile = 128;      // time indices (threads) per CTA, large grids; slabs of the sharded scan align to it
constexpr int kTileSmall = 64;  // ... small grids
constexpr int kATile = 8;       // robots held in registers
#ifndef KCBS_CONFLICT_SMALL_GRID
#define KCBS_CONFLICT_SMALL_GRID 6144   // CTAs of 128 below which the 64-index tile is used (about 6 waves)
#endif

__device__ __forceinline__ int clamp_k(int k, int T, const int *len, int r)
{
    int kk = min(k, T - 1);
    if (len) kk = min(kk, len[r] - 1);
    return max(kk, 0);
}

// ---- mbarrier / bulk-copy helpers (PTX) ------------------------------------------------------------
__device__ __forceinline__ unsigned smem_addr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
{
    unsigned done = 0;
    while (!done) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(smem_addr(bar)), "r"(parity)
            : "memory");
    }
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_addr(dst)),
                 "l"(src), "r"(bytes), "r"(smem_addr(bar))
                 : "memory");
}

// packed FP32 pairs (sm_100 add.rn.f32x2): two subtractions per issued instruction
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi)
{
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float &lo, float &hi)
{
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b)
{
    f32x2 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
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

// Unpacks a key and extends the conflict while the same pair keeps colliding.  Executed by one full warp.
__device__ __forceinline__ kcbs_conflict finish_plan(const DevWorld &w, unsigned long long key, const float *base,
                                                    const int *len, int T, int lane)
{
    kcbs_conflict c;
    if (key == kNoConflict) {
        c.r1 = c.r2 = c.k_start = c.k_end = -1;
        return c;
    }
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
    return c;
}

// ---- key exchange of the sharded scan (system-scope accesses to mailboxes in peer memory) -------------------------
// A message is one 64-bit word: tag (16 bits, never 0) << 48 | key (48 bits; all ones = no conflict).  It is written
// with a single store and is its own flag; mailboxes are double buffered by call parity (a rank cannot be two calls
// ahead of a peer, because finishing a call needs every peer's message of that call).
constexpr unsigned long long kKey48 = 0xffffffffffffull;
__device__ __forceinline__ void st_sys(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ unsigned epoch_tag(unsigned epoch) { return epoch % 65535u + 1u; }

// Executed by warp 0 of the last CTA of `plan` on this rank: publishes `mine`, returns the minimum over all ranks
// (kNoConflict - 1 after a timeout: no valid key has r1 = r2 = 255 at k = 2^31 - 1).
constexpr unsigned long long kPeerTimeoutKey = kNoConflict - 1;
__device__ __forceinline__ unsigned long long exchange_keys(const ConflictArgs &a, int plan, unsigned epoch,
                                                            unsigned long long mine, int lane)
{
    const unsigned tag = epoch_tag(epoch);
    const size_t slot = ((size_t)(epoch & 1u) * KCBS_MAX_PEERS) * a.box_plans + plan;
    unsigned long long got = kNoConflict;
    bool timed_out = false;
    if (lane < a.world) {
        st_sys(a.box_peer[lane] + slot + (size_t)a.rank * a.box_plans, ((unsigned long long)tag << 48) | (mine & kKey48));
        const unsigned long long *in = a.box_peer[a.rank] + slot + (size_t)lane * a.box_plans;
        unsigned long long v = ld_sys(in);
        if ((unsigned)(v >> 48) != tag) {
            const unsigned long long t0 = global_timer_ns();
            unsigned spins = 0;
            while ((unsigned)((v = ld_sys(in)) >> 48) != tag) {
                if ((++spins & 1023u) == 0 && global_timer_ns() - t0 > (unsigned long long)KCBS_PEER_TIMEOUT_S * 1000000000ull) {
                    timed_out = true;
                    break;
                }
            }
        }
        const unsigned long long k48 = v & kKey48;
        got = k48 == kKey48 ? kNoConflict : k48;
    }
    timed_out = __any_sync(0xffffffffu, timed_out);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const unsigned long long o = __shfl_xor_sync(0xffffffffu, got, off);
        got = o < got ? o : got;
    }
    if (timed_out) {
        if (lane == 0) *(volatile int *)a.error = 1;
        return kPeerTimeoutKey;
    }
    return got;
}

template <bool FUSED, int TILE>
__global__ void __launch_bounds__(TILE)
k_conflict_keys(const __grid_constant__ DevWorld w, const __grid_constant__ ConflictArgs a, int tiles_per_plan)
{
    constexpr int kTile = TILE;
    extern __shared__ __align__(128) float smem[];
    __shared__ __align__(8) unsigned long long s_bar;
    __shared__ int s_last;
    __shared__ unsigned s_epoch;

    const int R = a.R, T = a.T;
    const int plan = blockIdx.x / tiles_per_plan;
    const int tile = blockIdx.x - plan * tiles_per_plan;
    const int k0 = a.k_lo + tile * kTile;
    const int tid = threadIdx.x;
    const int k = k0 + tid;
    const float *base = a.traj + (long long)plan * R * 3 * T;
    const int *len = a.lengths ? a.lengths + (long long)plan * R : nullptr;
    const int Rp = (R + kATile - 1) / kATile * kATile;  // rows incl. the sentinel robots of the last register tile
    float *xs = smem;               // [Rp][kTile]
    float *ys = smem + Rp * kTile;  // [Rp][kTile]

    unsigned long long key = kNoConflict;

    const int n_k = min(kTile, a.k_hi - k0);   // time indices of this tile
    // bulk path: rows 16 B aligned, sizes multiples of 16 B, no padding inside the tile
    // (n_k <= 0: a rank of the sharded scan whose slab is empty still runs one CTA per plan for the exchange)
    bool bulk = (n_k > 0) && ((T & 3) == 0) && ((k0 & 3) == 0) && ((n_k & 3) == 0) && ((((size_t)a.traj) & 15) == 0);
    if (bulk && len) {
        for (int r = 0; r < R; ++r) bulk &= len[r] >= k0 + n_k;
    }
    if (bulk) {   // start the tile's copies first; everything below overlaps with them
        if (tid == 0) {
            mbar_init(&s_bar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        // arrive + expect_tx is one atomic update, so copies may complete before or after it
        if (tid == 0) mbar_expect_tx(&s_bar, (unsigned)(2 * R * n_k * sizeof(float)));
        for (int row = tid; row < 2 * R; row += kTile) {
            const int r = row >> 1, c = row & 1;
            bulk_g2s((c ? ys : xs) + r * kTile, base + ((long long)r * 3 + c) * T + k0, (unsigned)(n_k * sizeof(float)), &s_bar);
        }
    }
    // sentinel robots: far away from everything and from each other
    for (int r = R; r < Rp; ++r) {
        xs[r * kTile + tid] = 1.0e18f * (float)(r - R + 1);
        ys[r * kTile + tid] = 0.0f;
    }

    // a conflict already recorded at an earlier time index makes this tile irrelevant
    // (one load per warp, broadcast: other CTAs update the key concurrently, and `skip` guards full-mask warp collectives,
    // so it must be warp-uniform)
    unsigned long long seen = 0;
    if ((tid & 31) == 0) seen = *(volatile const unsigned long long *)(a.keys + plan);
    seen = __shfl_sync(0xffffffffu, seen, 0);
    const bool skip = (seen != kNoConflict && (long long)(seen >> 16) < (long long)k0) || n_k <= 0;

    if (bulk) mbar_wait(&s_bar, 0);   // (also when skipping: the copies must land before the CTA exits)

    if (!skip) {
        if (!bulk) {
#pragma unroll 4
            for (int r = 0; r < R; ++r) {
                const int kk = clamp_k(k, T, len, r);
                xs[r * kTile + tid] = __ldcs(base + ((long long)r * 3 + 0) * T + kk);
                ys[r * kTile + tid] = __ldcs(base + ((long long)r * 3 + 1) * T + kk);
            }
        }
        const bool active = tid < n_k;

        // ---- broad phase: min over pairs of the Chebyshev centre distance -----------------------------
        float m = 3.0e38f;
        if (active) {
            for (int a0 = 0; a0 < R; a0 += kATile) {
                float xa[kATile], ya[kATile];
                f32x2 xa2[kATile / 2], ya2[kATile / 2];
#pragma unroll
                for (int q = 0; q < kATile; ++q) {
                    xa[q] = xs[(a0 + q) * kTile + tid];
                    ya[q] = ys[(a0 + q) * kTile + tid];
                }
#pragma unroll
                for (int q = 0; q < kATile / 2; ++q) {
                    xa2[q] = pack2(xa[2 * q], xa[2 * q + 1]);
                    ya2[q] = pack2(ya[2 * q], ya[2 * q + 1]);
                }
#pragma unroll
                for (int q = 0; q < kATile; ++q)
#pragma unroll
                    for (int p = q + 1; p < kATile; ++p)
                        m = fminf(m, fmaxf(fabsf(xa[q] - xa[p]), fabsf(ya[q] - ya[p])));
#pragma unroll 2
                for (int r2 = a0 + kATile; r2 < R; ++r2) {
                    const float xb = xs[r2 * kTile + tid], yb = ys[r2 * kTile + tid];
                    const f32x2 xb2 = pack2(xb, xb), yb2 = pack2(yb, yb);
#pragma unroll
                    for (int q = 0; q < kATile / 2; ++q) {
                        float dx0, dx1, dy0, dy1;
                        unpack2(sub2(xa2[q], xb2), dx0, dx1);
                        unpack2(sub2(ya2[q], yb2), dy0, dy1);
                        m = fminf(m, fminf(fmaxf(fabsf(dx0), fabsf(dy0)), fmaxf(fabsf(dx1), fabsf(dy1))));
                    }
                }
            }
        }

        // ---- narrow phase (rare): exact rectangle test in scan order ------------------------------------
        if (active && m <= 2.0f * w.max_rad * 1.000001f) {
            for (int r1 = 0; r1 < R && key == kNoConflict; ++r1) {
                const float x1 = xs[r1 * kTile + tid], y1 = ys[r1 * kTile + tid];
                bool have1 = false;
                float c1 = 1.0f, s1 = 0.0f;
                for (int r2 = r1 + 1; r2 < R; ++r2) {
                    if (a.group && a.group[r1] == a.group[r2]) continue;  // two bodies of one merged individual
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

    if (FUSED) {
        // last CTA of the plan finishes it and leaves key + ticket clean for the next call
        __syncthreads();
        if (tid == 0) {
            // (the epoch is read before the ticket is taken: it advances only after every plan's last ticket)
            if (a.world > 1) s_epoch = *(volatile const unsigned *)a.epoch;
            __threadfence();
            const int ticket = atomicAdd(a.tickets + plan, 1);
            s_last = (ticket == tiles_per_plan - 1);
        }
        __syncthreads();
        if (s_last && tid < 32) {
            __threadfence();
            unsigned long long best = *(volatile const unsigned long long *)(a.keys + plan);
            if (a.world > 1) best = exchange_keys(a, plan, s_epoch, best, tid);
            kcbs_conflict c;
            if (best == kPeerTimeoutKey)
                c.r1 = c.r2 = c.k_start = c.k_end = -2;
            else
                c = finish_plan(w, best, base, len, T, tid);
            if (tid == 0) {
                a.out[plan] = c;
                a.keys[plan] = kNoConflict;
                a.tickets[plan] = 0;
                if (a.world > 1) {
                    __threadfence();
                    if (atomicAdd(a.done, 1) == a.P - 1) {  // last plan of the call: next call, next epoch
                        *a.done = 0;
                        *a.epoch = s_epoch + 1u;
                    }
                }
            }
        }
    }
}

// One warp per plan: unpack the key and extend the conflict (used after a cross-GPU key reduction).
__global__ void __launch_bounds__(128)
k_conflict_finish(const __grid_constant__ DevWorld w, const __grid_constant__ FinishArgs a)
{
    const int plan = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (plan >= a.P) return;
    const float *base = a.traj + (long long)plan * a.R * 3 * a.T;
    const int *len = a.lengths ? a.lengths + (long long)plan * a.R : nullptr;
    const kcbs_conflict c = finish_plan(w, a.keys[plan], base, len, a.T, lane);
    if (lane == 0) a.out[plan] = c;
}

__global__ void k_fill_keys(unsigned long long *keys, int *tickets, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        keys[i] = kNoConflict;
        if (tickets) tickets[i] = 0;
    }
}

cudaError_t launch_fill_keys(unsigned long long *keys, int *tickets, int n, cudaStream_t stream)
{
    if (n <= 0) return cudaSuccess;
    k_fill_keys<<<(n + 255) / 256, 256, 0, stream>>>(keys, tickets, n);
    return cudaGetLastError();
}

template <int TILE>
static cudaError_t launch_conflict_tiled(const DevWorld &w, const ConflictArgs &a, cudaStream_t stream)
{
    // an empty slab (sharded scan, fewer tiles than ranks) still runs one CTA per plan: it takes part in the exchange
    const int tiles = a.k_hi > a.k_lo ? (a.k_hi - a.k_lo + TILE - 1) / TILE : 1;
    const int Rp = (a.R + kATile - 1) / kATile * kATile;
    const size_t smem = (size_t)2 * Rp * TILE * sizeof(float);
    const unsigned grid = (unsigned)((long long)tiles * a.P);
    if (a.out)
        k_conflict_keys<true, TILE><<<grid, TILE, smem, stream>>>(w, a, tiles);
    else
        k_conflict_keys<false, TILE><<<grid, TILE, smem, stream>>>(w, a, tiles);
    return cudaGetLastError();
}

cudaError_t launch_conflict_keys(const DevWorld &w, const ConflictArgs &a, cudaStream_t stream)
{
    if (a.P <= 0) return cudaSuccess;
    if (a.k_hi <= a.k_lo && !(a.world > 1 && a.out)) return cudaSuccess;
    const long long ctas128 = (long long)((a.k_hi - a.k_lo + kTile - 1) / kTile) * a.P;
    if (ctas128 < KCBS_CONFLICT_SMALL_GRID) return launch_conflict_tiled<kTileSmall>(w, a, stream);
    return launch_conflict_tiled<kTile>(w, a, stream);
}

// Time slab of `rank` (whole 128-index tiles; the 64-index tile divides them).
void conflict_time_slab(int T, int rank, int world, int *k_lo, int *k_hi)
{
    const long long tiles = (T + kTile - 1) / kTile;
    const long long lo = tiles * rank / world, hi = tiles * (rank + 1) / world;
    *k_lo = (int)(lo * kTile < T ? lo * kTile : T);
    *k_hi = (int)(hi * kTile < T ? hi * kTile : T);
}

cudaError_t launch_conflict_finish(const DevWorld &w, const FinishArgs &a, cudaStream_t stream)
{
    if (a.P <= 0) return cudaSuccess;
    const int warps_per_block = 4;
    k_conflict_finish<<<(a.P + warps_per_block - 1) / warps_per_block, 128, 0, stream>>>(w, a);
    return cudaGetLastError();
}

cudaError_t init_conflict_kernels()
{
    const int bytes = 2 * KCBS_MAX_ROBOTS * kTile * (int)sizeof(float);
    cudaError_t e;
    if ((e = cudaFuncSetAttribute(k_conflict_keys<true, kTile>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(k_conflict_keys<false, kTile>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(k_conflict_keys<true, kTileSmall>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes / 2)) != cudaSuccess) return e;
    return cudaFuncSetAttribute(k_conflict_keys<false, kTileSmall>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes / 2);
}

}  // namespace kcbs
