// ls2d_tu_general.cu -- icp_general_kernel: every option of the aligner, clouds of any size (ls2d_general.cuh)
#include "ls2d_general.cuh"
#include "ls2d_internal.h"

namespace ls2d {

int launch_general(ls2d_handle* h, const align_args& a, int max_points) {
  if (a.n_pairs <= 0) return LS2D_OK;
  constexpr int T   = 512;
  const size_t smem = icp_general_smem_bytes(h->dp.cam.cols, T, max_points);
  if (max_points > 65535 || smem > SMEM_LIMIT) return LS2D_ERR_UNSUPPORTED;
  if (h->dp.with_sensor) {
    auto kern = icp_general_kernel<T, true>;
    if (int rc = configure_kernel(h, kern, (size_t) ((int) smem), false)) return rc;
    kern<<<a.n_pairs, T, smem, h->stream>>>(h->dp, a, max_points);
  } else {
    auto kern = icp_general_kernel<T, false>;
    if (int rc = configure_kernel(h, kern, (size_t) ((int) smem), false)) return rc;
    kern<<<a.n_pairs, T, smem, h->stream>>>(h->dp, a, max_points);
  }
  CU(cudaGetLastError());
  h->launches++;
  return LS2D_OK;
}

}  // namespace ls2d
