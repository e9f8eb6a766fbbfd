// DMMA m8n8k4 throughput vs resident warps per SM and independent accumulators per warp (B200).
// Answers: how many warps per scheduler does the dense kernel need to keep the FP64 tensor pipe busy?
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "CUDA error %s line %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int NACC>
__global__ void k(double* out, int iters, double seed) {
    double a = seed + threadIdx.x * 1e-9, b = 1.0 + 1e-12 * threadIdx.x;
    double c[NACC][2];
#pragma unroll
    for (int i = 0; i < NACC; i++) { c[i][0] = i; c[i][1] = -i; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 32 / NACC; r++)
#pragma unroll
            for (int i = 0; i < NACC; i++) dmma884(c[i][0], c[i][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
double run(int sms, int warps, double* out) {
    int iters = 20000;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    k<NACC><<<sms, warps * 32>>>(out, iters, 0.5); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 3; r++) {
        CK(cudaEventRecord(e0)); k<NACC><<<sms, warps * 32>>>(out, iters, 0.5); CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    return (double)sms * warps * iters * 32.0 * 256 * 2 / (best * 1e-3) / 1e12;
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    int sms = prop.multiProcessorCount;
    double* out; CK(cudaMalloc(&out, sizeof(double) * sms * 1024));
    printf("{\"gpu\": \"%s\", \"dmma_tflops_by_warps_per_sm\": {", prop.name);
    int ws[] = {4, 8, 12, 16, 24, 32};
    for (int i = 0; i < 6; i++) {
        printf("%s\"%d\": {\"acc16\": %.2f, \"acc8\": %.2f, \"acc4\": %.2f, \"acc2\": %.2f}", i ? ", " : "", ws[i],
               run<16>(sms, ws[i], out), run<8>(sms, ws[i], out), run<4>(sms, ws[i], out), run<2>(sms, ws[i], out));
    }
    printf("}}\n");
    return 0;
}
