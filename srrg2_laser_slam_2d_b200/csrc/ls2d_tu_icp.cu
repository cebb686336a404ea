// ls2d_tu_icp.cu -- kernel table of the single-slice aligner: (threads, points per thread, CTAs per SM) by cloud size.
// One kernel per shape; nothing here reads the environment.
#include "ls2d_icp2.cuh"
#include "ls2d_internal.h"

namespace ls2d {
namespace {

struct shape {
  int threads, ppt, minb;
  int kind;  // 0: points in registers (icp_fused_kernel); 1: streamed from global/L2 (icp_stream_kernel);
             // 3: points in registers, icp_fused2_kernel (compile-time column stride cs; needs cols < cs)
  int cs;    // kind 3: column stride (768 covers the 721-column canvas of the shipped configurations, 1152 the
             // 1081-column one); wider canvases fall back to kind 0
};

// measured on B200 (profiles/r01_variant_sweep.md, profiles/r02_*): the losers of the sweep are not compiled in
shape pick_shape(int max_points) {
  if (max_points <= 256) return {128, 2, 6, 0, 0};
  if (max_points <= 512) return {128, 4, 6, 0, 0};
  if (max_points <= 768) return {256, 3, 5, 3, 768};    // 721 beams: 0.246 ms per 4096 pairs
  if (max_points <= 1152) return {288, 4, 4, 3, 1152};  // 1081 beams: the headline shape
  if (max_points <= 1536) return {256, 6, 2, 0, 0};
  if (max_points <= 2048) return {256, 8, 2, 0, 0};
  if (max_points <= 4096) return {512, 8, 1, 0, 0};
  if (max_points <= 65535) return {512, 0, 2, 1, 0};  // streaming kernel, any size the shared-memory stash can hold
  return {0, 0, 0, -1, 0};
}

// the shape launch_icp() runs for these parameters: the compile-time-stride kernel serves the plane-to-plane factor
// on canvases narrower than its stride; everything else runs the run-time-shaped kernels
shape resolve_shape(const dev_params& dp, int max_points) {
  shape s = pick_shape(max_points);
  // (its projection takes the range gate on the squared range and a square root that is exact above 1e-30 m^2)
  if (s.kind == 3 && (dp.cam.cols >= s.cs || dp.gate2.lo < 1.0e-30f)) {
    s = s.cs == 768 ? shape{256, 3, 3, 0, 0} : shape{288, 4, 4, 0, 0};
  }
  if (s.kind == 0 && icp_smem_bytes(dp.cam.cols, s.threads, s.ppt) > SMEM_LIMIT) s = {512, 0, 2, 1, 0};
  return s;
}

template <int T, int PPT, bool SENSOR, int MINB>
int launch_icp_k(ls2d_handle* h, const align_args& a) {
  const size_t smem = icp_smem_bytes(h->dp.cam.cols, T, PPT);
  auto kern         = icp_fused_kernel<T, PPT, SENSOR, MINB>;
  if (int rc = configure_kernel(h, kern, (size_t) ((int) smem))) return rc;
  kern<<<a.n_pairs, T, smem, h->stream>>>(h->dp, a);
  CU(cudaGetLastError());
  h->launches++;
  return LS2D_OK;
}

template <int T, int PPT, int MINB>
int launch_icp_t(ls2d_handle* h, const align_args& a) {
  return h->dp.with_sensor ? launch_icp_k<T, PPT, true, MINB>(h, a) : launch_icp_k<T, PPT, false, MINB>(h, a);
}

template <int T, int PPT, bool SENSOR, int MINB, int CS, bool FUSED, bool P2P>
int launch_icp2_k(ls2d_handle* h, const align_args& a) {
  constexpr size_t smem = icp2_map<T, PPT, CS>::BYTES;
  auto kern             = icp_fused2_kernel<T, PPT, SENSOR, MINB, CS, FUSED, P2P>;
  if (int rc = configure_kernel(h, kern, (size_t) ((int) smem))) return rc;
  kern<<<a.n_pairs, T, smem, h->stream>>>(h->dp, a);
  CU(cudaGetLastError());
  h->launches++;
  return LS2D_OK;
}

template <int T, int PPT, int MINB, int CS, bool FUSED, bool P2P = false>
int launch_icp2_t(ls2d_handle* h, const align_args& a) {
  return h->dp.with_sensor ? launch_icp2_k<T, PPT, true, MINB, CS, FUSED, P2P>(h, a)
                           : launch_icp2_k<T, PPT, false, MINB, CS, FUSED, P2P>(h, a);
}

template <int T, bool SENSOR, int MINB>
int launch_stream_k(ls2d_handle* h, const align_args& a, int maxp) {
  const size_t smem = icp_stream_smem_bytes(h->dp.cam.cols, T, maxp, false);
  if (smem > SMEM_LIMIT) return LS2D_ERR_UNSUPPORTED;
  auto kern = icp_stream_kernel<T, SENSOR, false, MINB>;
  if (int rc = configure_kernel(h, kern, (size_t) ((int) smem))) return rc;
  kern<<<a.n_pairs, T, smem, h->stream>>>(h->dp, a, maxp);
  CU(cudaGetLastError());
  h->launches++;
  return LS2D_OK;
}

bool needs_general(const dev_params& dp) {
  return dp.algorithm != LS2D_ALGORITHM_GN || dp.inlier_only_runs || dp.termination_epsilon > 0.f;
}

// the compile-time-stride kernels accumulate with fused multiply-adds unless the caller asked for single-rounding
// sums (721-beam shape: 0.2553 -> 0.2457 ms per 4096 pairs)
// (plane-to-plane only: D18 restates that factor's arithmetic)
bool fused_accumulation(const shape& s, const dev_params& dp, bool single_rounding) {
  return s.kind == 3 && !single_rounding && dp.factor == LS2D_FACTOR_PLANE2PLANE;
}

}  // namespace

int launch_icp(ls2d_handle* h, const align_args& a) {
  if (a.n_pairs <= 0) return LS2D_OK;
  const int maxp = h->sets[0].max_points > h->sets[1].max_points ? h->sets[0].max_points : h->sets[1].max_points;
  if (needs_general(h->dp) && !a.score_only) return launch_general(h, a, maxp);
  const shape s = resolve_shape(h->dp, maxp);
  const bool p2p = h->dp.factor == LS2D_FACTOR_POINT2POINT;  // the factor is a template switch of the kernel
  if (s.kind == 3 && s.cs == 1152 && p2p) return launch_icp2_t<288, 4, 4, 1152, false, true>(h, a);
  if (s.kind == 3 && s.cs == 768 && p2p) return launch_icp2_t<256, 3, 5, 768, false, true>(h, a);
  if (s.kind == 3 && s.cs == 1152)
    return fused_accumulation(s, h->dp, h->prm.single_rounding_accumulation != 0) ? launch_icp2_t<288, 4, 4, 1152, true>(h, a)
                                                                                  : launch_icp2_t<288, 4, 4, 1152, false>(h, a);
  if (s.kind == 3 && s.cs == 768)
    return fused_accumulation(s, h->dp, h->prm.single_rounding_accumulation != 0) ? launch_icp2_t<256, 3, 5, 768, true>(h, a)
                                                                                  : launch_icp2_t<256, 3, 5, 768, false>(h, a);
  if (s.kind == 1) return h->dp.with_sensor ? launch_stream_k<512, true, 2>(h, a, maxp) : launch_stream_k<512, false, 2>(h, a, maxp);
#define LS2D_CASE(T, P, B) \
  if (s.kind == 0 && s.threads == T && s.ppt == P && s.minb == B) return launch_icp_t<T, P, B>(h, a);
  LS2D_CASE(128, 2, 6)
  LS2D_CASE(128, 4, 6)
  LS2D_CASE(256, 3, 3)
  LS2D_CASE(288, 4, 4)
  LS2D_CASE(256, 6, 2)
  LS2D_CASE(256, 8, 2)
  LS2D_CASE(512, 8, 1)
#undef LS2D_CASE
  return LS2D_ERR_UNSUPPORTED;
}

int icp_reduction_shape(const dev_params& dp, bool single_rounding, int max_points) {
  if (needs_general(dp)) return 512;  // icp_general_kernel: xor-butterfly, single-rounding
  const shape s = resolve_shape(dp, max_points);
  if (s.kind < 0) return LS2D_ERR_UNSUPPORTED;
  return s.threads | (s.kind == 3 ? 1 << 16 : 0) | (fused_accumulation(s, dp, single_rounding) ? 1 << 17 : 0);
}

}  // namespace ls2d
