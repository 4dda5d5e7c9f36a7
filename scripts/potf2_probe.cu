// potf2_probe.cu -- phase timing (clock64) of the 64 x 64 diagonal-block kernel in isolation:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o scripts/potf2_probe.bin scripts/potf2_probe.cu
#include <cstdio>
#include <vector>
#include "../dgp_b200/csrc/potf2.cuh"
namespace dgpb { thread_local char g_err[512]; std::atomic<long long> g_launches; }
using namespace dgpb;
int main() {
    const int n = 1024, B = 8;
    Geom g = make_geom(n, false);
    std::vector<double> h(g.elems(), 0.0);
    for (int i = 0; i < g.npad; ++i)
        for (int j = 0; j <= i; ++j) h[(size_t)i * g.ld + j] = (i == j) ? 2.0 : exp(-0.05 * (i - j) * (i - j));
    Batch bt;
    double *T, *D; int* info; long long* st;
    cudaMalloc(&T, g.elems() * 8 * B); cudaMalloc(&D, diag_elems(g) * 8 * B); cudaMalloc(&info, 4 * MAXB);
    cudaMalloc(&st, 8 * 8 * B + 1024); cudaMemset(info, 0, 4 * MAXB);
    for (int b = 0; b < MAXB; ++b) { bt.T[b] = b < B ? T + g.elems() * b : nullptr; bt.diag[b] = b < B ? D + diag_elems(g) * b : nullptr; }
    for (int b = 0; b < B; ++b) cudaMemcpy(bt.T[b], h.data(), g.elems() * 8, cudaMemcpyHostToDevice);
    bt.info = info;
    cudaFuncSetAttribute(potf2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPotf2Smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 3; ++rep) potf2_kernel<<<B, 256, kPotf2Smem>>>(bt, g.ld, g.npad, 0, st);
    cudaEventRecord(e0);
    for (int rep = 0; rep < 20; ++rep) potf2_kernel<<<B, 256, kPotf2Smem>>>(bt, g.ld, g.npad, 0, nullptr);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long s[8]; cudaMemcpy(s, st, 64, cudaMemcpyDeviceToHost);
    const char* names[] = {"prologue (loads, zero fill)", "column loop (32 x 2 columns)", "16x16 diagonal inverses", "off-diagonal level 1", "off-diagonal level 2", "write-out"};
    for (int i = 0; i < 6; ++i) printf("%-32s %8lld cycles\n", names[i], s[i + 1] - s[i]);
    printf("variant %d: total %lld cycles; back-to-back launches: %.2f us each; status %s\n", DGPB_POTF2_VARIANT, s[6] - s[0], ms * 1e3 / 20, cudaGetErrorString(cudaGetLastError()));
#if DGPB_POTF2_VARIANT == 9
    { long long f[8]; cudaMemcpy(f, st + 64, 64, cudaMemcpyDeviceToHost);
      const char* fn[] = {"publish (select + STS)", "barrier", "loads (ci, ck, piv)", "w + update", "next pivot", "owner L entries"};
      for (int i = 0; i < 6; ++i) printf("  step 9: %-28s %6lld cycles\n", fn[i], f[i + 1] - f[i]); }
#endif
    int hi[MAXB]; cudaMemcpy(hi, info, 4 * B, cudaMemcpyDeviceToHost); printf("info[0] = %d\n", hi[0]);
    // correctness: L L' = A (64 x 64 block) and inv(L) L = I
    std::vector<double> dg(diag_elems(g));
    cudaMemcpy(dg.data(), bt.diag[B - 1], diag_elems(g) * 8, cudaMemcpyDeviceToHost);
    const double* Lb = dg.data() + g.npad;
    const double* Di = dg.data() + (size_t)g.npad * (1 + NB);
    double er1 = 0, er2 = 0, er3 = 0;
    for (int i = 0; i < 64; ++i)
        for (int j = 0; j < 64; ++j) {
            double s1 = 0, s2 = 0;
            for (int k = 0; k < 64; ++k) { s1 += Lb[i * 64 + k] * Lb[j * 64 + k]; s2 += Di[i * 64 + k] * Lb[k * 64 + j]; }
            if (j <= i) er1 = fmax(er1, fabs(s1 - h[(size_t)i * g.ld + j]));
            er2 = fmax(er2, fabs(s2 - (i == j ? 1.0 : 0.0)));
            if (j > i) er3 = fmax(er3, fabs(Lb[i * 64 + j]) + fabs(Di[i * 64 + j]));
        }
    for (int i = 0; i < 64; ++i) er1 = fmax(er1, fabs(dg[i] - Lb[i * 64 + i]));
    printf("max |LL' - A| = %.3e, max |inv(L) L - I| = %.3e, upper part = %.3e\n", er1, er2, er3);
    return 0;
}
