// comm.cuh -- internal interface of the multi-GPU exchange steps (comm.cu).
#pragma once
#include "common.cuh"

namespace dgpb {

// recv (world x count) <- every rank's send (count doubles), enqueued on `st`
int comm_allgather(Workspace* ws, const double* send, double* recv, size_t count, cudaStream_t st);
// rows k with root_of_row[k] >= 0 of the row-major matrix `base` (row length row_elems) are broadcast from that
// rank; rows with a negative root are left alone (every rank computed them)
int comm_bcast_rows(Workspace* ws, double* base, int64_t row_elems, const int* root_of_row, int rows, cudaStream_t st);
// *global_flag = max over ranks of local_flag (synchronises `st`); world == 1: the local value
int comm_max_flag(Workspace* ws, int local_flag, int* global_flag, cudaStream_t st);

// Angles of one ESS wave: out[0] = th0, out[s] = the angle the bracket rule (imputation.py:111-119) draws after
// candidates 0..s-1 were rejected; u[s-1] = the uniform that rejection consumes.  Returns the number of candidates
// min(max_cand, 1 + nu_left) >= 1.  Shared by the device loop (ess.cu) and the host-only entry dgpb_ess_plan_wave.
inline int ess_plan_wave(double th0, double lmin, double lmax, const double* u, int nu_left, int max_cand, double* out) {
    const int S = std::max(1, std::min(max_cand, 1 + nu_left));
    out[0] = th0;
    for (int s = 1; s < S; ++s) {
        if (out[s - 1] < 0.0) lmin = out[s - 1]; else lmax = out[s - 1];   // imputation.py:115-118
        out[s] = lmin + (lmax - lmin) * u[s - 1];                          // imputation.py:119
    }
    return S;
}

}  // namespace dgpb
