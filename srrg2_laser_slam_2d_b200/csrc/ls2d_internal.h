// ls2d_internal.h -- what the translation units of libls2d.so share: the handle, the CUDA error plumbing and the
// kernel launchers.  Host-side bookkeeping only; there is no CPU implementation of the path anywhere in here.
//
//   ls2d_api.cu          the C ABI (include/ls2d.h): argument checks, device buffers, parameter translation
//   ls2d_tu_icp.cu       icp_fused2_kernel / icp_fused_kernel / icp_stream_kernel: kernel table by cloud size
//   ls2d_tu_general.cu   icp_general_kernel: Levenberg-Marquardt, inlier-only runs, termination criterion
//   ls2d_tu_score.cu     score_kernel: the single-linearisation pass, TMA-fed and persistent
//   ls2d_tu_multi.cu     the multi-slice aligners
//   ls2d_tu_service.cu   projector / finder / clipper / merger / best-of kernels; raw-scan pre-processor,
//                        voxelizing clipper, CSR packing
#pragma once

#include <cuda_runtime.h>

#include <vector>
#include <stdint.h>

#include "../../include/ls2d.h"
#include "ls2d_args.h"
#include "ls2d_stage.h"

namespace ls2d {

struct cloud_set {
  float4* pts     = nullptr;
  int* off        = nullptr;
  bool owned      = false;
  size_t cap_pts  = 0;  // points
  size_t cap_off  = 0;  // ints
  int n_clouds    = 0;
  int max_points  = 0;
};

struct scratch {
  void* p    = nullptr;
  size_t cap = 0;
};

}  // namespace ls2d

struct ls2d_handle {
  int device               = 0;
  cudaStream_t own_stream  = nullptr;
  cudaStream_t stream      = nullptr;
  ls2d_params prm;
  ls2d::dev_params dp;
  int pose_format = LS2D_POSE_XYT;
  ls2d::cloud_set sets[LS2D_MAX_CLOUD_SETS];
  ls2d::scratch d_fid, d_mid, d_init, d_out, d_iters, d_best, d_misc, d_prior, d_ranges, d_clip;
  ls2d::scratch d_edge;  // rounding-edge directions of the projector (polar_cam::edge), rebuilt by ls2d_set_params
  ls2d::polar_cam edge_key = {};  // camera the table in d_edge was built for
  ls2d::scratch d_edge_slice[LS2D_MAX_SLICES];  // the same for the slices of ls2d_align_multi, cached by camera
  ls2d::polar_cam edge_slice_key[LS2D_MAX_SLICES] = {};
  int64_t launches = 0;
  // host pipeline of ls2d_align_pairs_host: uploads run on their own stream, one event per chunk
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_ready     = nullptr;
  cudaEvent_t ev_chunk[8]  = {};
  // second compute lane of ls2d_track_batch (pre-processor), one "packed" event per chunk
  cudaStream_t aux_stream  = nullptr;
  cudaEvent_t ev_packed[8] = {};
  std::vector<float> h_ident;  // identity initial guesses of ls2d_track_batch (must outlive the asynchronous upload)
  ls2d::stage_pool stage;  // uploads from pageable caller buffers (ls2d_align_pairs_host, ls2d_track_batch)
  // NCCL, resolved lazily
  void* nccl_lib                                                               = nullptr;
  int (*nccl_all_gather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
  int (*nccl_comm_count)(void*, int*)                                          = nullptr;
  int sm_count = 0;
  // producers of packed cloud sets (pre-processor, clipper): look-back words and the launch epoch they carry
  ls2d::scratch d_look;
  unsigned pack_epoch = 0;
  ls2d::scratch d_ticket;  // finished-CTA counter of the pre-processor (zero between launches)
  // (cos, sin) per beam of the pre-processor, cached by (n_beams, sensor matrix)
  ls2d::scratch d_beam;
  int beam_n     = 0;
  float beam_ifx = 0.f, beam_cx = 0.f;
};

namespace ls2d {

void set_last_cuda_error(cudaError_t e, const char* what);

#define CU(call)                                   \
  do {                                             \
    cudaError_t e__ = (call);                      \
    if (e__ != cudaSuccess) {                      \
      ::ls2d::set_last_cuda_error(e__, #call);     \
      return LS2D_ERR_CUDA;                        \
    }                                              \
  } while (0)

constexpr size_t SMEM_LIMIT = 227 * 1024;  // dynamic shared memory a CTA can opt into on sm_100a

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize [, PreferredSharedMemoryCarveout]) only when a kernel asks for more
// than it was ever granted on this device: the attribute is an upper bound, and two driver calls per launch are
// measurable on the single-pair path.  (ls2d_api.cu; process-wide, per device and kernel)
int grant_shared_memory(int device, const void* kernel, size_t bytes, bool max_carveout);
template <typename K>
int configure_kernel(const ls2d_handle* h, K kernel, size_t bytes, bool max_carveout = true) {
  return grant_shared_memory(h->device, reinterpret_cast<const void*>(kernel), bytes, max_carveout);
}

int reserve(scratch& s, size_t bytes);
int pose_stride(const ls2d_handle* h);

// ---- launchers (one per translation unit; all asynchronous on h->stream) ------------------------------------------
// the aligner for LS2D_FIXED / LS2D_MOVING with the handle's parameters: picks the kernel by cloud size, canvas
// width and the options in h->dp
int launch_icp(ls2d_handle* h, const align_args& a);
// shape of the reduction launch_icp() would run (include/ls2d.h: ls2d_reduction_shape); negative: ls2d_error
int icp_reduction_shape(const dev_params& dp, bool single_rounding, int max_points);
int launch_general(ls2d_handle* h, const align_args& a, int max_points);
int launch_score(ls2d_handle* h, const align_args& a);
// shape of the scoring pass's reduction when score_kernel serves these parameters, else -1 (the aligner's shape)
int score_reduction_shape(const ls2d_handle* h_or_null, const dev_params& dp, bool single_rounding, int max_points);

int launch_multi(ls2d_handle* h, const multi_args& a, const int* cols);
int multi_reduction_threads();
int multi_reduction_shape(const multi_args& a);  // shape launch_multi() runs for these slices (ls2d_multi_reduction_shape)

int launch_project(ls2d_handle* h, const project_args& a);
int launch_correspond(ls2d_handle* h, const correspond_args& a);
int launch_clip(ls2d_handle* h, const clip_args& a, int n);
int launch_merge(ls2d_handle* h, const merge_args& a);
int launch_classify(ls2d_handle* h, const classify_args& a, const int* off_f, int cloud_f, const int* off_m, int cloud_m);
int launch_selftest_sqrt(ls2d_handle* h, unsigned lo_bits, unsigned hi_bits, unsigned long long* n_mismatch_dev);
int launch_best_of(ls2d_handle* h, const ls2d_result* res, int n, int n_guess, const ls2d_gates& g, int candidate_base,
                   ls2d_best* out);
int launch_best_of_groups(ls2d_handle* h, const ls2d_result* res, const int* group_off, int n_groups,
                          const int* moving_id, const ls2d_gates& g, ls2d_best* out);

int launch_preprocess(ls2d_handle* h, const scan_dev_params& P, const scan_args& a, int n_scans);
int launch_clip_voxel(ls2d_handle* h, const clip_args& a, int n, float inv_res);
// *out = the largest off[i + 1] - off[i], or -1 when a difference is negative
int launch_largest_cloud(ls2d_handle* h, const int* off, int n, int* out);
int launch_scan_pack(ls2d_handle* h, const float4* strided, const int* off, int stride, int n, float4* packed);

}  // namespace ls2d
