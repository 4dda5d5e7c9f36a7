// vecchia.cuh -- internal interface of vecchia.cu used by ess.cu.
#pragma once
#include "common.cuh"

namespace dgpb {

constexpr int kMaxBlock = 64;  // max conditioning-block size (m+1)

struct VKern {
    int kind, D, ard;
    double len[kMaxDim];
};
int make_vkern(int kind, int64_t D, const double* length_host, int64_t nlen, VKern* vk);

// out2_dev[0] = sum_i (L_i^-1 y_i)_last^2, out2_dev[1] = sum_i 2 log L_i,last     (vecchia.py:177-178)
int vecchia_llik_device(Workspace* ws, const VKern& vk, const double* X, const double* y, const int64_t* NN, int64_t n,
                        int64_t m1, double nugget, const double* nugget_diag, double* out2_dev, cudaStream_t st);
// out = (L/sqrt(scale))^-1 z in Vecchia order                                      (vecchia.py:133-140)
int vecchia_mvn_draw_device(Workspace* ws, const VKern& vk, const double* X, const int64_t* NN, int64_t n, int64_t m1,
                            double scale, double nugget, const double* z, double* out, cudaStream_t st);

// 1 = register-resident kernel for conditioning blocks of <= 32 points (default), 0 = shared-memory kernels only
int vecchia_set_small(int on);
// knn.cu: 1 = tensor-core screen + exact ranking (default), 0 = scalar exact kernel only
int knn_set_mma(int on);

}  // namespace dgpb
