// DMMA m8n8k4 throughput with DISTINCT operand registers (as in a real register-tiled GEMM):
// MT x NT accumulator tile per warp, A fragments a[MT], B fragments b[NT], all in registers.
// Variants: operands fixed in registers vs. refreshed from shared memory every k-step.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "CUDA error %s line %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int MT, int NT, bool SMEM>
__global__ void __launch_bounds__(256) k(double* out, int iters, double seed) {
    __shared__ double sa[8][4][32 + 1], sb[8][8][32 + 1];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double a[MT], b[NT], c[MT][NT][2];
#pragma unroll
    for (int i = 0; i < MT; i++) { a[i] = seed + i + lane * 1e-9; sa[warp][i % 4][lane] = a[i]; }
#pragma unroll
    for (int j = 0; j < NT; j++) { b[j] = 1.0 + 1e-12 * (j + lane); sb[warp][j % 8][lane] = b[j]; }
#pragma unroll
    for (int i = 0; i < MT; i++)
#pragma unroll
        for (int j = 0; j < NT; j++) { c[i][j][0] = i; c[i][j][1] = -j; }
    __syncthreads();
    for (int it = 0; it < iters; it++) {
        if (SMEM) {
#pragma unroll
            for (int i = 0; i < MT; i++) a[i] = ((volatile double*)&sa[warp][i % 4][0])[lane];
#pragma unroll
            for (int j = 0; j < NT; j++) b[j] = ((volatile double*)&sb[warp][j % 8][0])[lane];
        }
#pragma unroll
        for (int j = 0; j < NT; j++)
#pragma unroll
            for (int i = 0; i < MT; i++) dmma884(c[i][j][0], c[i][j][1], a[i], b[j]);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < MT; i++)
#pragma unroll
        for (int j = 0; j < NT; j++) s += c[i][j][0] + c[i][j][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MT, int NT, bool SMEM>
double run(int sms, int warps, double* out) {
    int iters = 4000;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    k<MT, NT, SMEM><<<sms, warps * 32>>>(out, iters, 0.5); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 3; r++) {
        CK(cudaEventRecord(e0)); k<MT, NT, SMEM><<<sms, warps * 32>>>(out, iters, 0.5); CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    return (double)sms * warps * iters * MT * NT * 256.0 * 2 / (best * 1e-3) / 1e12;
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    int sms = prop.multiProcessorCount;
    double* out; CK(cudaMalloc(&out, sizeof(double) * sms * 1024));
    printf("{\"gpu\": \"%s\",\n", prop.name);
    printf(" \"reg_2x8_w8\": %.2f, \"reg_4x8_w4\": %.2f, \"reg_4x8_w8\": %.2f, \"reg_2x4_w8\": %.2f, \"reg_1x1_w8\": %.2f,\n",
           run<2, 8, false>(sms, 8, out), run<4, 8, false>(sms, 4, out), run<4, 8, false>(sms, 8, out),
           run<2, 4, false>(sms, 8, out), run<1, 1, false>(sms, 8, out));
    printf(" \"smem_2x8_w8\": %.2f, \"smem_4x8_w4\": %.2f, \"smem_4x8_w8\": %.2f, \"smem_2x4_w8\": %.2f}\n",
           run<2, 8, true>(sms, 8, out), run<4, 8, true>(sms, 4, out), run<4, 8, true>(sms, 8, out),
           run<2, 4, true>(sms, 8, out));
    return 0;
}
