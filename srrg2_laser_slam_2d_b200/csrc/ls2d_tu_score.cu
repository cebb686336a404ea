// ls2d_tu_score.cu -- the scoring pass (ls2d_score_batch): score_kernel, persistent and TMA-fed (ls2d_score.cuh), for
// clouds of up to 1152 points with the plane-to-plane factor; other shapes run one linearisation of the aligner kernels.
#include "ls2d_internal.h"
#include "ls2d_score.cuh"

namespace ls2d {
namespace {

#ifndef LS2D_SCORE_T
#define LS2D_SCORE_T 288
#define LS2D_SCORE_PPT 4
#define LS2D_SCORE_MINB 3
#endif
constexpr int SCORE_T = LS2D_SCORE_T, SCORE_PPT = LS2D_SCORE_PPT, SCORE_MINB = LS2D_SCORE_MINB;  // 1152 point slots per cloud

bool score_kernel_serves(const ls2d_handle* h, int maxp) {
  // (the z-buffer cells hold rho bits or point indices: a squared-range gate that opens below 1e-30 m^2 would let
  // the two overlap -- and feed the gated square root operands it is not exact for)
  return maxp <= SCORE_T * SCORE_PPT && h->dp.factor == LS2D_FACTOR_PLANE2PLANE && h->dp.gate2.lo >= 1.0e-30f &&
         (size_t) score_map(maxp, h->dp.cam.cols, SCORE_T, SCORE_PPT).bytes() <= SMEM_LIMIT / SCORE_MINB - 1024;
}

template <bool SENSOR, bool FUSED>
int launch_score_k(ls2d_handle* h, const align_args& a, int maxp) {
  auto kern        = score_kernel<SCORE_T, SCORE_PPT, SENSOR, FUSED, SCORE_MINB>;
  const int smem   = score_map(maxp, h->dp.cam.cols, SCORE_T, SCORE_PPT).bytes();
  if (int rc = configure_kernel(h, kern, (size_t) (smem))) return rc;
  int per_sm = 0;
  CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, SCORE_T, (size_t) smem));
  if (per_sm < 1) return LS2D_ERR_UNSUPPORTED;
  const int grid = a.n_pairs < h->sm_count * per_sm ? a.n_pairs : h->sm_count * per_sm;  // persistent: one CTA per slot
  kern<<<grid, SCORE_T, (size_t) smem, h->stream>>>(h->dp, a, maxp);
  CU(cudaGetLastError());
  h->launches++;
  return LS2D_OK;
}

}  // namespace

int launch_score(ls2d_handle* h, const align_args& a) {
  if (a.n_pairs <= 0) return LS2D_OK;
  const int maxp = h->sets[0].max_points > h->sets[1].max_points ? h->sets[0].max_points : h->sets[1].max_points;
  if (!score_kernel_serves(h, maxp)) return launch_icp(h, a);  // a.score_only == 1: one linearisation of the aligner
  const bool fused = !h->prm.single_rounding_accumulation;
  if (h->dp.with_sensor) return fused ? launch_score_k<true, true>(h, a, maxp) : launch_score_k<true, false>(h, a, maxp);
  return fused ? launch_score_k<false, true>(h, a, maxp) : launch_score_k<false, false>(h, a, maxp);
}

int score_reduction_shape(const ls2d_handle* h_or_null, const dev_params& dp, bool single_rounding, int maxp) {
  (void) h_or_null;
  if (maxp <= SCORE_T * SCORE_PPT && dp.factor == LS2D_FACTOR_PLANE2PLANE && dp.gate2.lo >= 1.0e-30f &&
      (size_t) score_map(maxp, dp.cam.cols, SCORE_T, SCORE_PPT).bytes() <= SMEM_LIMIT / SCORE_MINB - 1024)
    return SCORE_T | (single_rounding ? 0 : 1 << 17);
  return -1;  // the aligner's own shape (icp_reduction_shape)
}

}  // namespace ls2d
