// pipes.cu -- microbenchmarks of the sm_100a issue / FMA / XU behaviour that K1 depends on.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o pipes pipes.cu && ./pipes
// Each test: one CTA per SM, W warps per SM sub-partition, a loop of independent packed / scalar / MUFU instructions;
// prints cycles per loop body per sub-partition and per warp.
#include <cstdio>
#include <cuda_runtime.h>

typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 r; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 r; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ float rcp(float x) { float r; asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float ffma(float a, float b, float c) { float r; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }

constexpr int N = 8;   // independent chains per thread

// MODE 0: FFMA2 r_i = a_i * b_i + r_i, all operands distinct registers
// MODE 1: FFMA2 r_i = r_i * b + c (two operands shared by all chains: reuse cache friendly)
// MODE 2: FMUL2 r_i = r_i * b_i
// MODE 3: scalar FFMA x2 per chain (same flops as MODE 0)
// MODE 4: MODE 0 + one MUFU.RCP per 4 FFMA2
// MODE 5: MODE 0 + one MUFU.RCP per 8 FFMA2
// MODE 6: MUFU.RCP only (8 per body)
// MODE 7: FADD2 r_i = r_i + b_i
// MODE 8: FFMA2 with an immediate: r_i = r_i * 1.0001 + b_i
template <int MODE>
__global__ void __launch_bounds__(1024, 1) k(float *sink, int iters, long long *cycles)
{
    // operands come from memory as 64-bit words: aligned register pairs, nothing for the compiler to fold or move
    const unsigned long long *src = reinterpret_cast<const unsigned long long *>(sink) + 16;
    f32x2 r[N], a[N], b[N];
    float x[N], y[N];
    const float s = __uint_as_float((unsigned)src[40 + (threadIdx.x & 1)]);
#pragma unroll
    for (int i = 0; i < N; ++i) {
        r[i] = src[i] + threadIdx.x;
        a[i] = src[8 + i];
        b[i] = src[16 + i];
        x[i] = __uint_as_float((unsigned)src[24 + i]);
        y[i] = __uint_as_float((unsigned)src[32 + i]);
    }
    float m = 1.0f + s;
    const f32x2 B = b[0], C = b[1];
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < N; ++i) {
            // in-place forms ("+l"): no register moves in the loop
            if (MODE == 0 || MODE == 4 || MODE == 5) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(r[i]) : "l"(a[i]), "l"(b[i]));
            if (MODE == 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(r[i]) : "l"(B), "l"(C));
            if (MODE == 2) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(r[i]) : "l"(a[i]));
            if (MODE == 3) {
                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[i]) : "f"(m), "f"(y[i]));
                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(y[i]) : "f"(m), "f"(x[i]));
            }
            if (MODE == 7) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(r[i]) : "l"(b[i]));
            if (MODE == 8) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(r[i]) : "l"(pk(1.0001f, 1.0001f)), "l"(b[i]));
            if (MODE == 9) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(r[i]) : "l"(a[i]), "l"(b[i]));
            if (MODE == 4 && (i % 4) == 3) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(m));
            if (MODE == 5 && (i % 8) == 7) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(m));
            if (MODE == 6) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
            if (MODE == 10) {   // K1-like mix: 3 packed + 1 MUFU (independent)
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(r[i]) : "l"(a[i]), "l"(b[i]));
                if (i % 3 == 0) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
            }
        }
    }
    const long long t1 = clock64();
    float acc = m;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        float lo, hi;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(r[i]));
        acc += lo + hi + x[i] + y[i];
    }
    if (acc == 12345.678f) sink[0] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char *name, int per_body)
{
    float *sink;
    long long *cyc, h[148];
    cudaMalloc(&sink, 4096); { float h0[1024]; for (int q = 0; q < 1024; ++q) h0[q] = 1.0f + 1e-6f * q; cudaMemcpy(sink, h0, 4096, cudaMemcpyHostToDevice); }
    cudaMalloc(&cyc, sizeof(h));
    const int iters = 20000;
    printf("%-58s", name);
    for (int warps = 1; warps <= 6; ++warps) {   // warps per sub-partition
        k<MODE><<<148, 128 * warps>>>(sink, iters, cyc);
        cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        const double c = (double)h[0] / iters;   // cycles per loop body (every warp runs one body per iteration)
        printf("  W=%d %6.2f c/body/SMSP (%5.2f c/instr)", warps, c, c / (per_body * warps));
        (void)per_body;
    }
    printf("\n");
    cudaFree(sink);
    cudaFree(cyc);
}

int main()
{
    printf("cycles per loop body of one warp scheduler (W warps each executing the body), and per instruction issued\n");
    run<0>("FFMA2 x8, 3 distinct register operands", 8);
    run<1>("FFMA2 x8, r = r * B + C (shared operands)", 8);
    run<8>("FFMA2 x8, r = r * imm + b_i", 8);
    run<2>("FMUL2 x8, r = r * a_i", 8);
    run<7>("FADD2 x8, r = r + b_i", 8);
    run<3>("FFMA x16 scalar", 16);
    run<6>("MUFU.RCP x8", 8);
    run<4>("FFMA2 x8 + MUFU.RCP x2", 10);
    run<5>("FFMA2 x8 + MUFU.RCP x1", 9);
    run<9>("FFMA2 x8, r = r * a_i + b_i", 8);
    run<10>("FFMA2 x8 + MUFU.RCP x3 (independent)", 11);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) printf("error: %s\n", cudaGetErrorString(e));
    return 0;
}
