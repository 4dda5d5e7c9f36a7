// DMMA.8x8x4 issue-rate probe: register-only loops, varying warps/SM and independent accumulators/warp.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/dmma_probe scripts/dmma_probe.cu && /tmp/dmma_probe
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int NACC>
__global__ void probe(double* out, int iters) {
    double c[NACC][2];
#pragma unroll
    for (int i = 0; i < NACC; ++i) c[i][0] = c[i][1] = threadIdx.x * 1e-3 + i;
    double a = 1.0 + threadIdx.x * 1e-6, b = 1.0 - threadIdx.x * 1e-6;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) dmma(c[i][0], c[i][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NACC>
void run(int warps_per_sm, double* out) {
    int threads = 32 * (warps_per_sm > 32 ? 32 : warps_per_sm);
    int ctas_per_sm = warps_per_sm > 32 ? warps_per_sm / 32 : 1;
    int iters = 4096;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    probe<NACC><<<148 * ctas_per_sm, threads>>>(out, 16);
    cudaEventRecord(e0);
    probe<NACC><<<148 * ctas_per_sm, threads>>>(out, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double n = 148.0 * warps_per_sm * (double)iters * NACC;
    double tf = n * 512 / (ms * 1e-3) / 1e12;
    printf("warps/SM %2d  acc/warp %2d : %6.2f TFLOP/s  (%.1f clk per DMMA per SMSP @1.965GHz)\n", warps_per_sm, NACC, tf,
           ms * 1e-3 * 1.965e9 / (n / (148.0 * 4)));
}
int main() {
    double* out; cudaMalloc(&out, sizeof(double) * 148 * 2048);
    int ws[] = {4, 8, 16, 32, 64};
    for (int w : ws) { run<1>(w, out); run<2>(w, out); run<4>(w, out); run<8>(w, out); run<16>(w, out); run<32>(w, out); }
    return 0;
}
