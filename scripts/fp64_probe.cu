// fp64_probe.cu -- latency of dependent FP64 operations on sm_100a (development aid for potf2_kernel):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/fp64_probe scripts/fp64_probe.cu && ./gpurun_out/fp64_probe
#include <cstdio>
#include <cuda_runtime.h>

__global__ void chain_dfma(double* out, double a, double b, int iters, long long* cyc) {
    double x = a + threadIdx.x;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < iters; ++i) x = fma(x, b, a);
    long long t1 = clock64();
    out[threadIdx.x + blockIdx.x * blockDim.x] = x;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void chain_rsqrt(double* out, double a, int iters, long long* cyc) {
    double x = a + threadIdx.x;
    long long t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < iters; ++i) x = rsqrt(x) + a;
    long long t1 = clock64();
    out[threadIdx.x + blockIdx.x * blockDim.x] = x;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void indep_dfma(double* out, double a, double b, int iters, long long* cyc) {
    double x[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = a + threadIdx.x + k;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) x[k] = fma(x[k], b, a);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += x[k];
    out[threadIdx.x + blockIdx.x * blockDim.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void barrier_cost(int iters, long long* cyc) {
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void sts_bar_lds(double* out, int iters, long long* cyc) {
    __shared__ double s[2][256];
    double x = threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        s[i & 1][threadIdx.x] = x;
        __syncthreads();
        x = s[i & 1][(threadIdx.x + 37) & 255] + 1.0;
    }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void lds128_round(double* out, int iters, long long* cyc, int nload) {
    __shared__ double2 s[2][64];
    __shared__ long long tmax;
    if (threadIdx.x < 64) { s[0][threadIdx.x] = make_double2(1.0, 2.0); s[1][threadIdx.x] = make_double2(0.5, 0.25); }
    __syncthreads();
    const int bi = threadIdx.x & 15, bk = (threadIdx.x >> 4) & 15;
    double acc = 0.0;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        const int buf = i & 1;
        if (bk == (i & 15)) s[buf][bi] = make_double2(acc, 1.0);
        __syncthreads();
        double2 v[10];
#pragma unroll
        for (int r = 0; r < 10; ++r) v[r] = (r < nload) ? s[buf][(16 * (r & 3) + ((r & 4) ? bk : bi) + (r >> 3)) & 63] : make_double2(0.0, 0.0);
#pragma unroll
        for (int r = 0; r < 10; ++r) acc += v[r].x * 1e-9 + v[r].y * 1e-9;
    }
    long long t1 = clock64();
    out[threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
    double* out; long long* cyc; long long h;
    cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 1024);
    const int it = 4096;
    for (int threads : {32, 256}) {
        chain_dfma<<<1, threads>>>(out, 1.0, 0.999, it, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("dependent DFMA chain, %3d threads: %.1f cycles/op\n", threads, (double)h / it);
        chain_rsqrt<<<1, threads>>>(out, 1.5, it, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("dependent rsqrt+add chain, %3d threads: %.1f cycles/op\n", threads, (double)h / it);
        indep_dfma<<<1, threads>>>(out, 1.0, 0.999, it, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("8 independent DFMA chains, %3d threads: %.2f cycles/DFMA/warp\n", threads, (double)h / it / 8);
    }
    indep_dfma<<<1, 1024>>>(out, 1.0, 0.999, it, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("8 independent DFMA chains, 1024 threads: %.2f cycles per DFMA per warp (8 warps/SMSP)\n", (double)h / it / 8);
    barrier_cost<<<1, 256>>>(it, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("__syncthreads, 256 threads: %.1f cycles\n", (double)h / it);
    sts_bar_lds<<<1, 256>>>(out, it, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("STS + __syncthreads + LDS + DADD round, 256 threads: %.1f cycles\n", (double)h / it);
    for (int threads : {32, 160, 256})
        for (int nload : {1, 10}) {
            lds128_round<<<1, threads>>>(out, it, cyc, nload); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
            printf("STS + barrier + %2d x LDS.128 + 20 dependent DP ops, %3d threads: %.1f cycles/round\n", nload, threads, (double)h / it);
        }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
