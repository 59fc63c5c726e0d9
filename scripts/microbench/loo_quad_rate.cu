// Ceiling probe for the LOO-EM quad inner loop (wgs::loo4_quad): the same packed instruction
// mix with (0) shared-memory operands + MUFU, (1) no MUFU, (2) register operands + MUFU,
// (3) register operands, no MUFU.  Reports posterior evaluations per second on one GPU.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o loo_quad_rate loo_quad_rate.cu
#include "../../wgsassign_b200/csrc/wgs_kernels.cuh"
#include <cstdio>
using namespace wgs;

template <bool MUFU>
__device__ __forceinline__ void quad(const ulonglong2 v0, const ulonglong2 v1, const ulonglong2 v2, const Loo4Coef (&c)[4], f32x2 (&acc)[4]) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const f32x2 nu = ffma2(v1.x, c[k].B, v0.y);
        const f32x2 nv = ffma2(v2.y, c[k].B, v2.x);
        const f32x2 du = ffma2(v0.x, c[k].A, fadd2(v0.y, nu));
        const f32x2 dv = ffma2(v1.y, c[k].A, fadd2(v2.x, nv));
        const f32x2 m = fmul2(du, dv);
        const f32x2 x = ffma2(nv, du, fmul2(nu, dv));
        if (MUFU) {
            const float2 mm = unpack2(m);
            acc[k] = ffma2(x, pack2(fast_rcp(mm.x), fast_rcp(mm.y)), acc[k]);
        } else {
            acc[k] = ffma2(x, m, acc[k]);
        }
    }
}

template <int MODE>
__global__ void __launch_bounds__(256, 3) probe(float* out, int nq, int rows, int iters, float f0)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    ulonglong2* tile = reinterpret_cast<ulonglong2*>(smem_raw);
    const int stride = (3 * nq) | 1;
    for (int e = threadIdx.x; e < rows * stride; e += blockDim.x) {
        ulonglong2 v; v.x = pack2(0.3f + 1e-3f * (e % 7), 0.2f); v.y = pack2(0.25f, 0.35f + 1e-3f * (e % 5));
        tile[e] = v;
    }
    __syncthreads();
    const int t = threadIdx.x, r = (t / nq) % rows;
    Loo4Coef c[4];
    f32x2 acc[4];
    for (int k = 0; k < 4; ++k) { c[k] = loo4_coef(f0 + 0.01f * k + 1e-4f * (t & 7)); acc[k] = 0ull; }
    const ulonglong2* row = tile + r * stride;
    ulonglong2 r0 = row[0], r1 = row[1], r2 = row[2];
    for (int it = 0; it < iters; ++it) {
        if (MODE < 2) {
            int q = 0;
#pragma unroll 1
            for (; q + 1 < nq; q += 2) {
                quad<MODE == 0>(row[3 * q], row[3 * q + 1], row[3 * q + 2], c, acc);
                quad<MODE == 0>(row[3 * q + 3], row[3 * q + 4], row[3 * q + 5], c, acc);
            }
            if (q < nq) quad<MODE == 0>(row[3 * q], row[3 * q + 1], row[3 * q + 2], c, acc);
        } else {
            int q = 0;
#pragma unroll 1
            for (; q + 1 < nq; q += 2) {
                quad<MODE == 2>(r0, r1, r2, c, acc);
                quad<MODE == 2>(r1, r2, r0, c, acc);
                asm volatile("" : "+l"(r0.x), "+l"(r1.y));      // keep the loop from being collapsed
            }
            if (q < nq) quad<MODE == 2>(r0, r1, r2, c, acc);
        }
    }
    float s = 0.f;
    for (int k = 0; k < 4; ++k) { float2 a = unpack2(acc[k]); s += a.x + a.y; }
    out[blockIdx.x * blockDim.x + t] = s;
}

template <bool MUFU>
__global__ void __launch_bounds__(256, 3) probe5(float* out, int nq, int rows, int iters, float f0)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    ulonglong2* tile = reinterpret_cast<ulonglong2*>(smem_raw);
    const int nc = (nq + 1) / 2, stride = (5 * nc) | 1;
    for (int e = threadIdx.x; e < rows * stride; e += blockDim.x) {
        ulonglong2 v; v.x = pack2(0.3f + 1e-3f * (e % 7), 0.2f); v.y = pack2(0.25f, 0.35f + 1e-3f * (e % 5));
        tile[e] = v;
    }
    __syncthreads();
    const int t = threadIdx.x, r = (t / nq) % rows;
    Loo5Coef c[4];
    f32x2 acc[4];
    for (int k = 0; k < 4; ++k) { c[k] = loo5_coef(f0 + 0.01f * k + 1e-4f * (t & 7)); acc[k] = 0ull; }
    const ulonglong2* row = tile + r * stride;
    for (int it = 0; it < iters; ++it) {
        const int nfull = nq >> 1;
#pragma unroll 1
        for (int cc = 0; cc < nfull; ++cc) {
            const ulonglong2 v0 = row[5 * cc], v1 = row[5 * cc + 1], v2 = row[5 * cc + 2], v3 = row[5 * cc + 3], v4 = row[5 * cc + 4];
            loo5_quad(v0.x, v0.y, v1.x, v1.y, v2.x, c, acc);
            loo5_quad(v2.y, v3.x, v3.y, v4.x, v4.y, c, acc);
        }
        if (nq & 1) {
            const ulonglong2 v0 = row[5 * nfull], v1 = row[5 * nfull + 1], v2 = row[5 * nfull + 2];
            loo5_quad(v0.x, v0.y, v1.x, v1.y, v2.x, c, acc);
        }
    }
    float s = 0.f;
    for (int k = 0; k < 4; ++k) { float2 a = unpack2(acc[k]); s += a.x + a.y; }
    out[blockIdx.x * blockDim.x + t] = s;
}
void run5(float* d, int block, int nq, int rows, int iters)
{
    const size_t smem = (size_t)rows * ((5 * ((nq + 1) / 2)) | 1) * 16;
    cudaFuncSetAttribute(probe5<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, probe5<true>, block, smem);
    const int grid = 148 * occ;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    probe5<true><<<grid, block, smem>>>(d, nq, rows, 10, 0.3f);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    probe5<true><<<grid, block, smem>>>(d, nq, rows, iters, 0.3f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double evals = (double)grid * block * iters * nq * 16.0;
    printf("%-28s block %3d occ %d: %.3e evals/s  (%s)\n", "pair polynomials (step5)", block, occ, evals / (ms * 1e-3), cudaGetErrorString(cudaGetLastError()));
}

template <int MODE> void run(const char* name, float* d, int block, int nq, int rows, int iters)
{
    const size_t smem = (size_t)rows * ((3 * nq) | 1) * 16;
    cudaFuncSetAttribute(probe<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    cudaFuncSetAttribute(probe<MODE>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, probe<MODE>, block, smem);
    const int grid = 148 * occ;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    probe<MODE><<<grid, block, smem>>>(d, nq, rows, 10, 0.3f);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    probe<MODE><<<grid, block, smem>>>(d, nq, rows, iters, 0.3f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double evals = (double)grid * block * iters * nq * 16.0;   // 4 problems x 4 individuals per quad
    printf("%-28s block %3d occ %d: %.3e evals/s  (%s)\n", name, block, occ, evals / (ms * 1e-3), cudaGetErrorString(cudaGetLastError()));
}

int main()
{
    float* d; cudaMalloc(&d, 148 * 8 * 512 * 4);
    for (int block : {128, 224, 256}) {
        run<0>("smem operands + MUFU", d, block, 13, 34, 2000);
        run<1>("smem operands, no MUFU", d, block, 13, 34, 2000);
        run5(d, block, 13, 34, 2000);
    }
    return 0;
}
