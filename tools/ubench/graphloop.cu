// Micro-benchmark: a chain of R rounds x 4 dependent kernels (one wide, three small) issued three ways:
//   (a) plain stream launches, (b) one CUDA graph of 4 R kernel nodes, (c) a WHILE conditional node whose body is the
//   four kernels and whose condition the last kernel updates on the device.  S independent chains run at once, each on
//   its own stream (and its own host thread for (a)).
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o graphloop graphloop.cu
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

__global__ void k_wide(float *buf, int spin)
{
    float v = buf[blockIdx.x * blockDim.x + threadIdx.x];
    for (int i = 0; i < spin; ++i) v = fmaf(v, 1.0001f, 0.5f);
    buf[blockIdx.x * blockDim.x + threadIdx.x] = v;
}
__global__ void k_small(float *buf, int spin)
{
    float v = buf[threadIdx.x];
    for (int i = 0; i < spin; ++i) v = fmaf(v, 1.0001f, 0.5f);
    buf[threadIdx.x] = v;
}
__global__ void k_tail(float *buf, int *counter, int rounds, cudaGraphConditionalHandle h, int use_handle)
{
    if (threadIdx.x == 0) {
        const int r = ++*counter;
        if (use_handle) cudaGraphSetConditional(h, r < rounds ? 1u : 0u);
    }
}
__global__ void k_reset(int *counter, cudaGraphConditionalHandle h, int use_handle)
{
    *counter = 0;
    if (use_handle) cudaGraphSetConditional(h, 1u);
}

struct Chain {
    cudaStream_t s;
    float *buf;
    int *counter;
    cudaGraphExec_t flat = nullptr, loop = nullptr;
    cudaGraphConditionalHandle h;
};

static void rounds_on_stream(Chain &c, int R, int wide_ctas, int spin_wide, int spin_small)
{
    k_reset<<<1, 1, 0, c.s>>>(c.counter, c.h, 0);
    for (int r = 0; r < R; ++r) {
        k_wide<<<wide_ctas, 128, 0, c.s>>>(c.buf, spin_wide);
        k_small<<<4, 128, 0, c.s>>>(c.buf, spin_small);
        k_small<<<8, 128, 0, c.s>>>(c.buf, spin_small / 2);
        k_tail<<<1, 32, 0, c.s>>>(c.buf, c.counter, R, c.h, 0);
    }
}

int main(int argc, char **argv)
{
    const int R = argc > 1 ? atoi(argv[1]) : 16, wide_ctas = argc > 2 ? atoi(argv[2]) : 600;
    const int spin_wide = argc > 3 ? atoi(argv[3]) : 3000, spin_small = argc > 4 ? atoi(argv[4]) : 3000;
    const int maxS = 16;
    std::vector<Chain> ch(maxS);
    for (auto &c : ch) {
        CK(cudaStreamCreateWithFlags(&c.s, cudaStreamNonBlocking));
        CK(cudaMalloc(&c.buf, sizeof(float) * 128 * 2048));
        CK(cudaMemset(c.buf, 0, sizeof(float) * 128 * 2048));
        CK(cudaMalloc(&c.counter, sizeof(int)));
        // (b) flat graph by capture
        cudaGraph_t g;
        CK(cudaStreamBeginCapture(c.s, cudaStreamCaptureModeThreadLocal));
        rounds_on_stream(c, R, wide_ctas, spin_wide, spin_small);
        CK(cudaStreamEndCapture(c.s, &g));
        CK(cudaGraphInstantiate(&c.flat, g, 0));
        CK(cudaGraphDestroy(g));
        // (c) while node
        cudaGraph_t gl;
        CK(cudaGraphCreate(&gl, 0));
        CK(cudaGraphConditionalHandleCreate(&c.h, gl, 1, cudaGraphCondAssignDefault));
        cudaGraphNode_t reset_node, while_node;
        {
            cudaKernelNodeParams kp = {};
            int zero = 1;
            void *args[] = {&c.counter, &c.h, &zero};
            kp.func = (void *)k_reset; kp.gridDim = dim3(1); kp.blockDim = dim3(1); kp.kernelParams = args;
            CK(cudaGraphAddKernelNode(&reset_node, gl, nullptr, 0, &kp));
        }
        cudaGraphNodeParams wp = {};
        wp.type = cudaGraphNodeTypeConditional;
        wp.conditional.handle = c.h;
        wp.conditional.type = cudaGraphCondTypeWhile;
        wp.conditional.size = 1;
        CK(cudaGraphAddNode(&while_node, gl, &reset_node, 1, &wp));
        cudaGraph_t body = wp.conditional.phGraph_out[0];
        {
            cudaGraphNode_t n0, n1, n2, n3;
            cudaKernelNodeParams kp = {};
            void *a0[] = {&c.buf, (void *)&spin_wide};
            kp.func = (void *)k_wide; kp.gridDim = dim3(wide_ctas); kp.blockDim = dim3(128); kp.kernelParams = a0;
            CK(cudaGraphAddKernelNode(&n0, body, nullptr, 0, &kp));
            void *a1[] = {&c.buf, (void *)&spin_small};
            kp.func = (void *)k_small; kp.gridDim = dim3(4); kp.kernelParams = a1;
            CK(cudaGraphAddKernelNode(&n1, body, &n0, 1, &kp));
            int half = spin_small / 2;
            void *a2[] = {&c.buf, &half};
            kp.gridDim = dim3(8); kp.kernelParams = a2;
            CK(cudaGraphAddKernelNode(&n2, body, &n1, 1, &kp));
            int Rv = R, one = 1;
            void *a3[] = {&c.buf, &c.counter, &Rv, &c.h, &one};
            kp.func = (void *)k_tail; kp.gridDim = dim3(1); kp.blockDim = dim3(32); kp.kernelParams = a3;
            CK(cudaGraphAddKernelNode(&n3, body, &n2, 1, &kp));
        }
        CK(cudaGraphInstantiate(&c.loop, gl, 0));
        CK(cudaGraphDestroy(gl));
    }
    CK(cudaDeviceSynchronize());
    auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    printf("R=%d rounds, wide kernel %d CTAs\n", R, wide_ctas);
    for (int S : {1, 2, 4, 8, 16}) {
        double best[3] = {1e9, 1e9, 1e9};
        for (int rep = 0; rep < 6; ++rep) {
            for (int mode = 0; mode < 3; ++mode) {
                CK(cudaDeviceSynchronize());
                const double t0 = now();
                if (mode == 0) {
                    std::vector<std::thread> th;
                    for (int i = 0; i < S; ++i)
                        th.emplace_back([&, i] { rounds_on_stream(ch[i], R, wide_ctas, spin_wide, spin_small); cudaStreamSynchronize(ch[i].s); });
                    for (auto &t : th) t.join();
                } else {
                    for (int i = 0; i < S; ++i) CK(cudaGraphLaunch(mode == 1 ? ch[i].flat : ch[i].loop, ch[i].s));
                    for (int i = 0; i < S; ++i) CK(cudaStreamSynchronize(ch[i].s));
                }
                const double dt = now() - t0;
                if (rep > 0 && dt < best[mode]) best[mode] = dt;
            }
        }
        int cnt = 0;
        CK(cudaMemcpy(&cnt, ch[0].counter, sizeof(int), cudaMemcpyDeviceToHost));
        printf("S=%2d chains: stream launches %8.1f us, flat graph %8.1f us, while graph %8.1f us  (per round: %.1f / %.1f / %.1f us; counter %d)\n",
               S, best[0] * 1e6, best[1] * 1e6, best[2] * 1e6, best[0] * 1e6 / R, best[1] * 1e6 / R, best[2] * 1e6 / R, cnt);
    }
    return 0;
}
