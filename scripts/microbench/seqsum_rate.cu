// Where does block_seqsum32 spend its time?  Variants of the round structure on 4 M addends, one block.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o seqsum_rate.bin seqsum_rate.cu
#include "../../wgsassign_b200/csrc/wgs_kernels.cuh"
#include <cstdio>
#include <vector>
#include <cmath>
using namespace wgs;

// V0: the product kernel.  V1: loads + chunk Q only (no fold).  V2: loads only.
template <int V>
__global__ void __launch_bounds__(kSeqWarps * 32) probe(const float* x, long n, float* out)
{
    if (V == 0) {
        const float r = block_seqsum32(0.0f, n, [&](long i) { return x[i]; });
        if (threadIdx.x == 0) out[0] = r;
        return;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __shared__ unsigned sq[kSeqWarps];
    float acc = 0.f;
    unsigned qq = 0;
    for (long r0 = 0; r0 < n; r0 += 256L * kSeqWarps) {
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { const long i = r0 + warp * 256 + 32 * j + lane; v[j] = i < n ? x[i] : 0.f; }
        if (V == 1) {
            unsigned Q; bool bad;
            seq_chunk_q(v, 100, Q, bad);
            if (lane == 0) sq[warp] = Q + bad;
            __syncthreads();
            qq += sq[0];
            __syncthreads();
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) acc += v[j];
        }
    }
    if (threadIdx.x == 0) out[0] = acc + qq;
}

template <int V> void run(const char* name, const float* d, long n, float* dout)
{
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    probe<V><<<1, kSeqWarps * 32>>>(d, n, dout);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    probe<V><<<1, kSeqWarps * 32>>>(d, n, dout);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    float h; cudaMemcpy(&h, dout, 4, cudaMemcpyDeviceToHost);
    printf("%-28s %.3f ms  (%.2f us per round of 8192)  result %g  %s\n", name, ms, ms * 1e3 / (n / 8192.0), h, cudaGetErrorString(cudaGetLastError()));
}

int main()
{
    const long n = 4 << 20;
    std::vector<float> h(n);
    unsigned s = 12345;
    for (long i = 0; i < n; ++i) { s = s * 1664525u + 1013904223u; float u = (s >> 8) * (1.0f / 16777216.0f); h[i] = u * u * 4e-8f; }
    float *d, *dout;
    cudaMalloc(&d, n * 4); cudaMalloc(&dout, 4);
    cudaMemcpy(d, h.data(), n * 4, cudaMemcpyHostToDevice);
    run<2>("loads only", d, n, dout);
    run<1>("loads + chunk Q + 2 syncs", d, n, dout);
    run<0>("block_seqsum32", d, n, dout);
    float ser = 0.f;
    for (long i = 0; i < n; ++i) ser += h[i];
    printf("serial float32 reference: %g\n", ser);
    return 0;
}
