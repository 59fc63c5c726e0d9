// MUFU.RCP issue-rate microbenchmark: 8 independent dependent chains per thread, x <- rcp(x) + c.
#include <cuda_runtime.h>
#include <cstdio>
__device__ __forceinline__ float rcp(float x) { float r; asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
template <int FMA_PER_RCP>
__global__ void k(float* out, float x, int iters)
{
    float a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = x + i;
    float b[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) b[i] = x * 0.5f + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            a[i] = rcp(a[i]) + 1.5f;
#pragma unroll
            for (int j = 0; j < FMA_PER_RCP; ++j) b[i] = fmaf(b[i], 0.999f, 0.001f);
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i] + b[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int F> void run(float* d, const char* name)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int it = 4000;
    k<F><<<148 * 8, 256>>>(d, 1.0f, 10); cudaDeviceSynchronize();
    cudaEventRecord(e0); k<F><<<148 * 8, 256>>>(d, 1.0f, it); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double rcps = (double)148 * 8 * 256 * it * 8;
    printf("%s: %.3e rcp/s  (%.2f rcp/clk/SM at 1.965 GHz)  + %d FFMA each: %.3e fma/s\n", name, rcps / (ms * 1e-3), rcps / (ms * 1e-3) / 148 / 1.965e9, F, rcps * F / (ms * 1e-3));
}
int main()
{
    float* d; cudaMalloc(&d, 148 * 8 * 256 * 4);
    run<0>(d, "rcp+fadd only");
    run<2>(d, "rcp + 2 ffma ");
    run<5>(d, "rcp + 5 ffma ");
    run<7>(d, "rcp + 7 ffma ");
    return 0;
}
