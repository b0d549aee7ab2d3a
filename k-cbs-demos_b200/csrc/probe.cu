// probe.cu -- FFMA saturation probe: measures the FP32 rate this device sustains, which is the
// denominator of the propagation kernel's roofline (MEASURED_PEAKS.json has no FP32 entry).
#include "kcbs_device.cuh"
#include "kernels.h"

namespace kcbs {

constexpr int kProbeChains = 8;
constexpr int kProbeInner = 64;

__global__ void __launch_bounds__(256) k_fp32_probe(float *sink, int iters)
{
    float acc[kProbeChains];
    const float a = 1.0f + 1e-7f * (float)threadIdx.x, b = 1e-9f * (float)blockIdx.x;
#pragma unroll
    for (int c = 0; c < kProbeChains; ++c) acc[c] = (float)c;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < kProbeInner; ++u)
#pragma unroll
            for (int c = 0; c < kProbeChains; ++c) acc[c] = fmaf(acc[c], a, b);
    }
    float s = 0.0f;
#pragma unroll
    for (int c = 0; c < kProbeChains; ++c) s += acc[c];
    if (s == 123.456f) sink[0] = s;  // keeps the chains alive without a store in practice
}

// Same chains with the packed fma.rn.f32x2 instruction (sm_100): two FMAs per lane per instruction.
__global__ void __launch_bounds__(256) k_fp32x2_probe(float *sink, int iters)
{
    unsigned long long acc[kProbeChains];
    const float a = 1.0f + 1e-7f * (float)threadIdx.x, b = 1e-9f * (float)blockIdx.x;
    unsigned long long a2, b2;
    asm("mov.b64 %0, {%1, %1};" : "=l"(a2) : "f"(a));
    asm("mov.b64 %0, {%1, %1};" : "=l"(b2) : "f"(b));
#pragma unroll
    for (int c = 0; c < kProbeChains; ++c) {
        const float v = (float)c;
        asm("mov.b64 %0, {%1, %1};" : "=l"(acc[c]) : "f"(v));
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < kProbeInner; ++u)
#pragma unroll
            for (int c = 0; c < kProbeChains; ++c) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(acc[c]) : "l"(a2), "l"(b2));
    }
    float s = 0.0f;
#pragma unroll
    for (int c = 0; c < kProbeChains; ++c) {
        float lo, hi;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc[c]));
        s += lo + hi;
    }
    if (s == 123.456f) sink[0] = s;
}

cudaError_t launch_fp32x2_probe(float *sink, int iters, cudaStream_t stream, double *flops)
{
    int dev = 0, sms = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return e;
    const int block = 256, grid = sms * 8;
    k_fp32x2_probe<<<grid, block, 0, stream>>>(sink, iters);
    if (flops) *flops = 4.0 * (double)grid * block * (double)iters * kProbeInner * kProbeChains;
    return cudaGetLastError();
}

cudaError_t launch_fp32_probe(float *sink, int iters, cudaStream_t stream, double *flops)
{
    int dev = 0, sms = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return e;
    const int block = 256, grid = sms * 8;
    k_fp32_probe<<<grid, block, 0, stream>>>(sink, iters);
    if (flops) *flops = 2.0 * (double)grid * block * (double)iters * kProbeInner * kProbeChains;
    return cudaGetLastError();
}

}  // namespace kcbs
