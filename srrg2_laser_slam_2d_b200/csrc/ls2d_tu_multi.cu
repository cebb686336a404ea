// ls2d_tu_multi.cu -- the multi-slice aligners: icp_multi2_kernel (register-resident, the shape of the shipped
// configurations: two slices of up to 768 points on canvases below 768 columns; ls2d_multi2.cuh) and the general
// icp_multi_kernel (any number of slices, any size; ls2d_multi.cuh)
#include "ls2d_internal.h"
#include "ls2d_multi2.cuh"

namespace ls2d {
namespace {

constexpr int M2_T = 256, M2_PPF = 3, M2_PPM = 6, M2_CS = 768;

// two slices with their own fixed clouds (up to 768 points) aligning ONE shared moving cloud (up to 1536 points)
bool multi2_serves(const multi_args& a) {
  if (a.n_slices != 2 || a.max_fixed_points > M2_T * M2_PPF || a.max_points > M2_T * M2_PPM || a.max_cols >= M2_CS) return false;
  if (a.sl[0].moving_pts != a.sl[1].moving_pts || a.sl[0].moving_off != a.sl[1].moving_off) return false;
  for (int s = 0; s < 2; ++s)
    if (a.sl[s].P.factor != LS2D_FACTOR_PLANE2PLANE || a.sl[s].P.gate2.lo < 1.0e-30f || !a.sl[s].P.cam.edge) return false;
  return true;
}

}  // namespace

int multi_reduction_threads() { return MULTI_THREADS; }

int multi_reduction_shape(const multi_args& a) {
  return multi2_serves(a) ? (M2_T | 1 << 16 | (a.fused ? 1 << 17 : 0)) : MULTI_THREADS;
}

int launch_multi(ls2d_handle* h, const multi_args& a, const int* cols) {
  if (a.n_pairs <= 0) return LS2D_OK;
  if (multi2_serves(a)) {
    using map = multi2_map<M2_T, M2_PPF, M2_PPM, M2_CS>;
    auto kern = a.fused ? icp_multi2_kernel<M2_T, M2_PPF, M2_PPM, M2_CS, true> : icp_multi2_kernel<M2_T, M2_PPF, M2_PPM, M2_CS, false>;
    if (int rc = configure_kernel(h, kern, (size_t) (map::BYTES))) return rc;
    kern<<<a.n_pairs, M2_T, map::BYTES, h->stream>>>(a);
    CU(cudaGetLastError());
    h->launches++;
    return LS2D_OK;
  }
  constexpr int T   = MULTI_THREADS;  // 4 CTAs of 8 warps per SM measured best (512 x 2: +28 % time, 384 x 3: +16 %)
  const size_t smem = multi_smem_bytes(cols, a.n_slices, a.max_cols, a.max_points, T);
  if (smem > SMEM_LIMIT) return LS2D_ERR_UNSUPPORTED;
  auto kern = icp_multi_kernel<T, 4>;
  if (int rc = configure_kernel(h, kern, (size_t) ((int) smem))) return rc;
  kern<<<a.n_pairs, T, smem, h->stream>>>(a);
  CU(cudaGetLastError());
  h->launches++;
  return LS2D_OK;
}

}  // namespace ls2d
