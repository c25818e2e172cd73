// Host-side interface of the TMA-staged chain GEMM (chain_gemm_tma.cu).
#pragma once
#include "common.cuh"

namespace tn {

struct alignas(64) TmaMap {  // opaque CUtensorMap (128 bytes)
  unsigned char bytes[128];
};

bool tma_available();
int tma_encode_2d(TmaMap* out, const double* base, long long rows, long long cols, long long ld, int box_rows);
int tma_encode_3d(TmaMap* out, const double* base, long long K, int d, long long Ny, long long ld);
int tma_box_rows_a();
int tma_box_rows_b_nt();
int gemm_launch_tma(const GemmLaunch& L, const GemmSchedule& S, const ProblemDev* problems_dev, const LinkDev* links_dev,
                    const TmaMap* maps_dev, const TmaMap& psi_a, const TmaMap& psi_b, const double* dyn_in, double* dyn_out,
                    double dyn_alpha, cudaStream_t stream);

}  // namespace tn
