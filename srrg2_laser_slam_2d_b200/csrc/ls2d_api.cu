// ls2d_api.cu -- implementation of the C ABI declared in include/ls2d.h.
// Host-side bookkeeping only: device buffers, parameter translation, kernel selection and launches.
// There is deliberately no CPU implementation behind these entry points.
#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <new>
#include <vector>

#include <map>
#include <mutex>
#include <utility>

#include "ls2d_internal.h"

using namespace ls2d;

namespace {
thread_local char g_last_cuda[256] = "";
}

namespace ls2d {

void set_last_cuda_error(cudaError_t e, const char* what) {
  snprintf(g_last_cuda, sizeof(g_last_cuda), "CUDA error: %s (%s)", cudaGetErrorString(e), what);
}

int pose_stride(const ls2d_handle* h) { return h->pose_format == LS2D_POSE_ISO ? 4 : 3; }

int grant_shared_memory(int device, const void* kernel, size_t bytes, bool max_carveout) {
  static std::mutex mu;
  static std::map<std::pair<int, const void*>, size_t> granted;
  std::lock_guard<std::mutex> lk(mu);
  size_t& have = granted[{device, kernel}];
  if (have >= bytes && have != 0) return LS2D_OK;
  CU(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) bytes));
  if (max_carveout) CU(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  have = bytes > 0 ? bytes : 1;
  return LS2D_OK;
}

int reserve(scratch& s, size_t bytes) {
  if (bytes <= s.cap) return LS2D_OK;
  if (s.p) cudaFree(s.p);
  s.p   = nullptr;
  s.cap = 0;
  const size_t want = bytes + bytes / 4 + 256;
  CU(cudaMalloc(&s.p, want));
  s.cap = want;
  return LS2D_OK;
}

}  // namespace ls2d

namespace {

void release(scratch& s) {
  if (s.p) cudaFree(s.p);
  s.p   = nullptr;
  s.cap = 0;
}

void release(cloud_set& c) {
  if (c.owned) {
    if (c.pts) cudaFree(c.pts);
    if (c.off) cudaFree(c.off);
  }
  c = cloud_set();
}

bool params_valid(const ls2d_params& p) {
  if (p.canvas_cols < 1 || p.canvas_cols > 7680) return false;
  if (!(p.angle_col_max > p.angle_col_min)) return false;
  if (!(p.range_max > p.range_min)) return false;
  if (!(p.range_max <= 1.0e5f)) return false;  // invalid beams travel as points at 1e6 m: beyond every legal range_max
  if (p.max_iterations < 0 || p.max_iterations > 100000) return false;
  if (p.with_sensor < 0 || p.with_sensor > 2) return false;
  if (p.factor != LS2D_FACTOR_PLANE2PLANE && p.factor != LS2D_FACTOR_POINT2POINT) return false;
  if (p.algorithm != LS2D_ALGORITHM_GN && p.algorithm != LS2D_ALGORITHM_LM) return false;
  if (p.algorithm == LS2D_ALGORITHM_LM && (p.lm_iterations_max < 1 || p.lm_iterations_max > 1000)) return false;
  return true;
}

// sensor_in_robot of a WithSensor slice in either form (include/ls2d.h: with_sensor 1 / 2)
iso sensor_iso(const ls2d_params& p) {
  if (p.with_sensor == 2) {
    iso S;
    S.tx = p.sensor_in_robot[0], S.ty = p.sensor_in_robot[1], S.c = p.sensor_in_robot_cs[0], S.s = p.sensor_in_robot_cs[1];
    return S;
  }
  return iso_v2t(p.sensor_in_robot[0], p.sensor_in_robot[1], p.sensor_in_robot[2]);
}

dev_params translate(const ls2d_params& p) {
  dev_params d;
  d.cam                     = make_polar_cam(p.canvas_cols, p.angle_col_min, p.angle_col_max);
  d.range_min               = p.range_min;
  d.range_max               = p.range_max;
  d.gate2                   = make_range_gate2(p.range_min, p.range_max);
  d.point_distance          = p.point_distance;
  d.normal_cos              = p.normal_cos;
  d.tau                     = p.cauchy_chi_threshold;
  d.inv_tau                 = p.cauchy_chi_threshold > 0.f ? 1.f / p.cauchy_chi_threshold : 0.f;
  d.damping                 = p.damping;
  d.max_iterations          = p.max_iterations;
  d.min_num_correspondences = p.min_num_correspondences;
  d.min_num_inliers         = p.min_num_inliers;
  d.with_sensor             = p.with_sensor ? 1 : 0;
  d.Sinv                    = iso_identity();
  if (p.with_sensor) d.Sinv = iso_inverse(sensor_iso(p));
  d.factor                  = p.factor;
  d.algorithm               = p.algorithm;
  d.lm_user_lambda_init     = p.lm_user_lambda_init;
  d.lm_tau                  = p.lm_tau;
  d.lm_step_low             = p.lm_step_low;
  d.lm_step_high            = p.lm_step_high;
  d.lm_iterations_max       = p.lm_iterations_max;
  d.lm_variable_damping     = p.lm_variable_damping;
  d.inlier_only_runs        = p.enable_inlier_only_runs;
  d.termination_epsilon     = p.termination_epsilon;
  return d;
}

// iteration records one alignment may write (include/ls2d.h: iter_stats)
int iters_per_pair(const ls2d_params& p) { return p.max_iterations * (p.enable_inlier_only_runs ? 2 : 1); }

// table of a camera's rounding edges as the kernels read it: [cols + 1] polar_edge (binary64 directions, second tier
// of the aligner kernels) followed by [cols + 1] polar_edge_f (binary32, the scoring kernel stages it in shared memory)
size_t edge_table_bytes(const polar_cam& cam) { return ((size_t) cam.cols + 1) * (sizeof(polar_edge) + sizeof(polar_edge_f)); }

int upload_edge_table(ls2d_handle* h, scratch& buf, polar_cam& key, polar_cam& cam) {
  if (!(buf.p && key.cols == cam.cols && key.K00 == cam.K00 && key.K01 == cam.K01)) {
    const size_t n = (size_t) cam.cols + 1;
    std::vector<unsigned char> host(edge_table_bytes(cam));
    fill_polar_edges(cam, reinterpret_cast<polar_edge*>(host.data()));
    fill_polar_edges_f(cam, reinterpret_cast<polar_edge_f*>(host.data() + n * sizeof(polar_edge)));
    key.cols = 0;
    int rc   = reserve(buf, host.size());
    if (rc) return rc;
    CU(cudaMemcpyAsync(buf.p, host.data(), host.size(), cudaMemcpyHostToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));  // `host` is pageable and dies here
    key = cam;
  }
  cam.edge = static_cast<const polar_edge*>(buf.p);
  return LS2D_OK;
}

// device copy of the projector's rounding edges (second tier of the column decision, ls2d_math.cuh); the table depends
// on the camera constants only and is re-used across ls2d_set_params
int upload_edges(ls2d_handle* h) {
  h->dp.cam.edge = nullptr;
  return upload_edge_table(h, h->d_edge, h->edge_key, h->dp.cam);
}

// edge table of slice `slot` of the multi-slice aligner: uploaded when the slice's camera changes
int slice_edges(ls2d_handle* h, int slot, polar_cam& cam) {
  return upload_edge_table(h, h->d_edge_slice[slot], h->edge_slice_key[slot], cam);
}

bool ready(const ls2d_handle* h) { return h->sets[0].pts && h->sets[1].pts && h->sets[0].off && h->sets[1].off; }

int h2d(ls2d_handle* h, scratch& s, const void* src, size_t bytes) {
  int rc = reserve(s, bytes);
  if (rc) return rc;
  CU(cudaMemcpyAsync(s.p, src, bytes, cudaMemcpyHostToDevice, h->stream));
  return LS2D_OK;
}

align_args base_args(const ls2d_handle* h) {
  align_args a;
  memset(&a, 0, sizeof(a));
  a.pose_stride  = pose_stride(h);
  a.iters_stride = iters_per_pair(h->prm);
  a.fixed_pts   = h->sets[0].pts;
  a.fixed_off   = h->sets[0].off;
  a.moving_pts  = h->sets[1].pts;
  a.moving_off  = h->sets[1].off;
  a.moving_div  = 1;
  a.fixed_const = -1;
  return a;
}

}  // namespace

extern "C" {

int ls2d_version(void) { return LS2D_VERSION; }

const char* ls2d_strerror(int err) {
  switch (err) {
    case LS2D_OK: return "ok";
    case LS2D_ERR_INVALID: return "invalid argument";
    case LS2D_ERR_CUDA: return g_last_cuda[0] ? g_last_cuda : "CUDA error";
    case LS2D_ERR_NOT_READY: return "clouds or parameters not set";
    case LS2D_ERR_UNSUPPORTED: return "size outside the compiled kernel table";
    case LS2D_ERR_NCCL: return "NCCL error";
    default: return "unknown error";
  }
}

void ls2d_default_params(ls2d_params* p) {
  if (!p) return;
  memset(p, 0, sizeof(*p));
  p->canvas_cols             = 721;
  p->angle_col_min           = -3.14159f;
  p->angle_col_max           = 3.14159f;
  p->range_min               = 0.3f;
  p->range_max               = 20.f;
  p->point_distance          = 0.5f;
  p->normal_cos              = 0.8f;
  p->cauchy_chi_threshold    = 0.01f;
  p->damping                 = 0.f;
  p->max_iterations          = 10;
  p->min_num_correspondences = 0;
  p->min_num_inliers         = 10;
  p->with_sensor             = 0;
  p->factor                  = LS2D_FACTOR_PLANE2PLANE;
  p->algorithm               = LS2D_ALGORITHM_GN;
  p->lm_user_lambda_init     = 0.f;
  p->lm_tau                  = 1e-5f;
  p->lm_step_low             = 1.f / 3.f;
  p->lm_step_high            = 2.f / 3.f;
  p->lm_iterations_max       = 10;
  p->lm_variable_damping     = 1;
}

int ls2d_create(ls2d_handle** out, int device) {
  if (!out) return LS2D_ERR_INVALID;
  *out = nullptr;
  int n_dev = 0;
  CU(cudaGetDeviceCount(&n_dev));
  if (device < 0 || device >= n_dev) return LS2D_ERR_INVALID;
  CU(cudaSetDevice(device));
  ls2d_handle* h = new (std::nothrow) ls2d_handle();
  if (!h) return LS2D_ERR_INVALID;
  h->device = device;
  if (cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
    delete h;
    return LS2D_ERR_CUDA;
  }
  h->stream = h->own_stream;
  ls2d_default_params(&h->prm);
  h->dp = translate(h->prm);
  if (upload_edges(h) != LS2D_OK) {
    cudaStreamDestroy(h->own_stream);
    delete h;
    return LS2D_ERR_CUDA;
  }
  cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, device);
  *out = h;
  return LS2D_OK;
}

int ls2d_destroy(ls2d_handle* h) {
  if (!h) return LS2D_OK;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  for (cloud_set& c : h->sets) release(c);
  release(h->d_prior);
  release(h->d_fid);
  release(h->d_mid);
  release(h->d_init);
  release(h->d_out);
  release(h->d_iters);
  release(h->d_best);
  release(h->d_misc);
  release(h->d_ranges);
  release(h->d_clip);
  release(h->d_look);
  release(h->d_ticket);
  release(h->d_beam);
  release(h->d_edge);
  for (scratch& e : h->d_edge_slice) release(e);
  if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
  if (h->aux_stream) cudaStreamDestroy(h->aux_stream);
  for (cudaEvent_t e : h->ev_packed)
    if (e) cudaEventDestroy(e);
  if (h->ev_ready) cudaEventDestroy(h->ev_ready);
  for (cudaEvent_t e : h->ev_chunk)
    if (e) cudaEventDestroy(e);
  if (h->own_stream) cudaStreamDestroy(h->own_stream);
  if (h->nccl_lib) dlclose(h->nccl_lib);
  delete h;
  return LS2D_OK;
}

int ls2d_set_stream(ls2d_handle* h, void* s) {
  if (!h) return LS2D_ERR_INVALID;
  h->stream = s ? (cudaStream_t) s : h->own_stream;
  return LS2D_OK;
}

int ls2d_sync(ls2d_handle* h) {
  if (!h) return LS2D_ERR_INVALID;
  CU(cudaSetDevice(h->device));
  CU(cudaStreamSynchronize(h->stream));
  return LS2D_OK;
}

int ls2d_set_params(ls2d_handle* h, const ls2d_params* p) {
  if (!h || !p || !params_valid(*p)) return LS2D_ERR_INVALID;
  h->prm = *p;
  h->dp  = translate(*p);
  CU(cudaSetDevice(h->device));
  return upload_edges(h);
}

int ls2d_get_params(const ls2d_handle* h, ls2d_params* p) {
  if (!h || !p) return LS2D_ERR_INVALID;
  *p = h->prm;
  return LS2D_OK;
}

int ls2d_set_pose_format(ls2d_handle* h, int format) {
  if (!h || (format != LS2D_POSE_XYT && format != LS2D_POSE_ISO)) return LS2D_ERR_INVALID;
  h->pose_format = format;
  return LS2D_OK;
}

int ls2d_get_pose_format(const ls2d_handle* h) { return h ? h->pose_format : LS2D_ERR_INVALID; }

int ls2d_upload_clouds(ls2d_handle* h, int which, const float* pts, const int32_t* off, int32_t n_clouds) {
  if (!h || which < 0 || which >= LS2D_MAX_CLOUD_SETS || !off || n_clouds < 0 || (!pts && n_clouds > 0 && off[n_clouds] > 0))
    return LS2D_ERR_INVALID;
  if (off[0] != 0) return LS2D_ERR_INVALID;
  int maxp = 0;
  for (int i = 0; i < n_clouds; ++i) {
    const int n = off[i + 1] - off[i];
    if (n < 0) return LS2D_ERR_INVALID;
    if (n > maxp) maxp = n;
  }
  CU(cudaSetDevice(h->device));
  cloud_set& c = h->sets[which];
  if (!c.owned) c = cloud_set();
  c.owned              = true;
  const size_t total   = (size_t) off[n_clouds];
  const size_t n_off   = (size_t) n_clouds + 1;
  if (total > c.cap_pts || !c.pts) {
    if (c.pts) cudaFree(c.pts);
    c.pts     = nullptr;
    c.cap_pts = 0;
    CU(cudaMalloc((void**) &c.pts, (total + total / 8 + 16) * sizeof(float4)));
    c.cap_pts = total + total / 8 + 16;
  }
  if (n_off > c.cap_off || !c.off) {
    if (c.off) cudaFree(c.off);
    c.off     = nullptr;
    c.cap_off = 0;
    CU(cudaMalloc((void**) &c.off, (n_off + n_off / 8 + 16) * sizeof(int)));
    c.cap_off = n_off + n_off / 8 + 16;
  }
  if (total) CU(cudaMemcpyAsync(c.pts, pts, total * sizeof(float4), cudaMemcpyHostToDevice, h->stream));
  CU(cudaMemcpyAsync(c.off, off, n_off * sizeof(int), cudaMemcpyHostToDevice, h->stream));
  c.n_clouds   = n_clouds;
  c.max_points = maxp;
  return LS2D_OK;
}

int ls2d_set_clouds_dev(ls2d_handle* h, int which, const void* pts_dev, const int32_t* off_dev,
                        int32_t n_clouds, int32_t max_points) {
  if (!h || which < 0 || which >= LS2D_MAX_CLOUD_SETS || !pts_dev || !off_dev || n_clouds < 0 || max_points < 0)
    return LS2D_ERR_INVALID;
  if (((uintptr_t) pts_dev) & 15) return LS2D_ERR_INVALID;
  // max_points picks the kernel and its capacity: an understated value would silently drop points, so the offsets
  // (which must be complete when this is called) are checked against it once, here
  if (n_clouds > 0) {
    CU(cudaSetDevice(h->device));
    int rc, largest = 0;
    if ((rc = reserve(h->d_misc, sizeof(int)))) return rc;
    if ((rc = launch_largest_cloud(h, off_dev, n_clouds, (int*) h->d_misc.p))) return rc;
    CU(cudaMemcpyAsync(&largest, h->d_misc.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    if (largest < 0 || largest > max_points) return LS2D_ERR_INVALID;  // < 0: offsets not ascending
  }
  release(h->sets[which]);
  cloud_set& c = h->sets[which];
  c.pts        = (float4*) pts_dev;
  c.off        = (int*) off_dev;
  c.owned      = false;
  c.n_clouds   = n_clouds;
  c.max_points = max_points;
  return LS2D_OK;
}

static int align_dev_impl(ls2d_handle* h, const int32_t* fid, const int32_t* mid, const float* init,
                          int32_t n_pairs, ls2d_result* out, ls2d_iter_stats* iters, int score_only) {
  if (!h || !init || !out || n_pairs < 0) return LS2D_ERR_INVALID;
  if (!ready(h)) return LS2D_ERR_NOT_READY;
  CU(cudaSetDevice(h->device));
  align_args a = base_args(h);
  a.fixed_id   = fid;
  a.moving_id  = mid;
  a.init_pose  = init;
  a.out        = out;
  a.iters      = iters;
  a.n_pairs    = n_pairs;
  a.score_only = score_only;
  return score_only ? launch_score(h, a) : launch_icp(h, a);
}

static int check_ids(const ls2d_handle* h, const int32_t* fid, const int32_t* mid, int32_t n) {
  for (int i = 0; i < n; ++i) {
    const int f = fid ? fid[i] : i, m = mid ? mid[i] : i;
    if (f < 0 || f >= h->sets[0].n_clouds || m < 0 || m >= h->sets[1].n_clouds) return LS2D_ERR_INVALID;
  }
  return LS2D_OK;
}

static int align_host_impl(ls2d_handle* h, const int32_t* fid, const int32_t* mid, const float* init,
                           int32_t n_pairs, ls2d_result* out, ls2d_iter_stats* iters, int score_only) {
  if (!h || !init || !out || n_pairs < 0) return LS2D_ERR_INVALID;
  if (!ready(h)) return LS2D_ERR_NOT_READY;
  if (n_pairs == 0) return LS2D_OK;
  int rc = check_ids(h, fid, mid, n_pairs);
  if (rc) return rc;
  CU(cudaSetDevice(h->device));
  if (fid && (rc = h2d(h, h->d_fid, fid, sizeof(int) * (size_t) n_pairs))) return rc;
  if (mid && (rc = h2d(h, h->d_mid, mid, sizeof(int) * (size_t) n_pairs))) return rc;
  if ((rc = h2d(h, h->d_init, init, sizeof(float) * pose_stride(h) * (size_t) n_pairs))) return rc;
  if ((rc = reserve(h->d_out, sizeof(ls2d_result) * (size_t) n_pairs))) return rc;
  const size_t iter_bytes = sizeof(ls2d_iter_stats) * (size_t) n_pairs * (size_t) iters_per_pair(h->prm);
  if (iters && iter_bytes) {
    if ((rc = reserve(h->d_iters, iter_bytes))) return rc;
    CU(cudaMemsetAsync(h->d_iters.p, 0, iter_bytes, h->stream));
  }
  rc = align_dev_impl(h, fid ? (const int*) h->d_fid.p : nullptr, mid ? (const int*) h->d_mid.p : nullptr,
                      (const float*) h->d_init.p, n_pairs, (ls2d_result*) h->d_out.p,
                      (iters && iter_bytes) ? (ls2d_iter_stats*) h->d_iters.p : nullptr, score_only);
  if (rc) return rc;
  CU(cudaMemcpyAsync(out, h->d_out.p, sizeof(ls2d_result) * (size_t) n_pairs, cudaMemcpyDeviceToHost, h->stream));
  if (iters && iter_bytes) CU(cudaMemcpyAsync(iters, h->d_iters.p, iter_bytes, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return LS2D_OK;
}

int ls2d_align_batch(ls2d_handle* h, const int32_t* fid, const int32_t* mid, const float* init,
                     int32_t n_pairs, ls2d_result* out, ls2d_iter_stats* iters) {
  return align_host_impl(h, fid, mid, init, n_pairs, out, iters, 0);
}

int ls2d_align_batch_dev(ls2d_handle* h, const int32_t* fid, const int32_t* mid, const float* init,
                         int32_t n_pairs, ls2d_result* out, ls2d_iter_stats* iters) {
  return align_dev_impl(h, fid, mid, init, n_pairs, out, iters, 0);
}

int ls2d_score_batch(ls2d_handle* h, const int32_t* fid, const int32_t* mid, const float* xyt,
                     int32_t n_pairs, ls2d_result* out) {
  return align_host_impl(h, fid, mid, xyt, n_pairs, out, nullptr, 1);
}

int ls2d_score_batch_dev(ls2d_handle* h, const int32_t* fid, const int32_t* mid, const float* xyt,
                         int32_t n_pairs, ls2d_result* out) {
  return align_dev_impl(h, fid, mid, xyt, n_pairs, out, nullptr, 1);
}

// grows an owned cloud set to hold `total` points / `n_off` offsets (no copies)
static int reserve_set(cloud_set& c, size_t total, size_t n_off) {
  if (!c.owned) c = cloud_set();
  c.owned = true;
  if (total > c.cap_pts || !c.pts) {
    if (c.pts) cudaFree(c.pts);
    c.pts     = nullptr;
    c.cap_pts = 0;
    CU(cudaMalloc((void**) &c.pts, (total + total / 8 + 16) * sizeof(float4)));
    c.cap_pts = total + total / 8 + 16;
  }
  if (n_off > c.cap_off || !c.off) {
    if (c.off) cudaFree(c.off);
    c.off     = nullptr;
    c.cap_off = 0;
    CU(cudaMalloc((void**) &c.off, (n_off + n_off / 8 + 16) * sizeof(int)));
    c.cap_off = n_off + n_off / 8 + 16;
  }
  return LS2D_OK;
}

// One call from host buffers.  The batch is cut into up to 8 chunks of whole pairs: chunk k+1's clouds cross
// PCIe on the copy stream while chunk k aligns, so the call costs the upload plus one chunk's kernel.
int ls2d_align_pairs_host(ls2d_handle* h, const float* fpts, const int32_t* foff, const float* mpts,
                          const int32_t* moff, const float* init, int32_t n_pairs, ls2d_result* out) {
  if (!h || !foff || !moff || !init || !out || n_pairs < 0) return LS2D_ERR_INVALID;
  if (n_pairs == 0) return LS2D_OK;
  if (foff[0] != 0 || moff[0] != 0 || (!fpts && foff[n_pairs] > 0) || (!mpts && moff[n_pairs] > 0)) return LS2D_ERR_INVALID;
  int maxf = 0, maxm = 0;
  for (int i = 0; i < n_pairs; ++i) {
    const int nf = foff[i + 1] - foff[i], nm = moff[i + 1] - moff[i];
    if (nf < 0 || nm < 0) return LS2D_ERR_INVALID;
    if (nf > maxf) maxf = nf;
    if (nm > maxm) maxm = nm;
  }
  CU(cudaSetDevice(h->device));
  int rc;
  cloud_set& F = h->sets[LS2D_FIXED];
  cloud_set& M = h->sets[LS2D_MOVING];
  if ((rc = reserve_set(F, (size_t) foff[n_pairs], (size_t) n_pairs + 1))) return rc;
  if ((rc = reserve_set(M, (size_t) moff[n_pairs], (size_t) n_pairs + 1))) return rc;
  F.n_clouds = M.n_clouds = n_pairs;
  F.max_points = maxf, M.max_points = maxm;
  if ((rc = reserve(h->d_init, sizeof(float) * pose_stride(h) * (size_t) n_pairs))) return rc;
  if ((rc = reserve(h->d_out, sizeof(ls2d_result) * (size_t) n_pairs))) return rc;
  if (!h->copy_stream) {
    CU(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&h->ev_ready, cudaEventDisableTiming));
    for (cudaEvent_t& e : h->ev_chunk) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  }
  // earlier work on the compute stream may still read the sets: the uploads wait for it
  CU(cudaEventRecord(h->ev_ready, h->stream));
  CU(cudaStreamWaitEvent(h->copy_stream, h->ev_ready, 0));
  CU(cudaMemcpyAsync(F.off, foff, sizeof(int) * ((size_t) n_pairs + 1), cudaMemcpyHostToDevice, h->copy_stream));
  CU(cudaMemcpyAsync(M.off, moff, sizeof(int) * ((size_t) n_pairs + 1), cudaMemcpyHostToDevice, h->copy_stream));
  CU(cudaMemcpyAsync(h->d_init.p, init, sizeof(float) * pose_stride(h) * (size_t) n_pairs, cudaMemcpyHostToDevice, h->copy_stream));
  int n_chunks = n_pairs / 512;  // a chunk's upload (~0.3 ms) is far longer than its kernel: the more chunks, the
                                 // shorter the tail left after the last upload (~one wave of 4 CTAs x 148 SMs)
  n_chunks     = n_chunks < 1 ? 1 : (n_chunks > 8 ? 8 : n_chunks);
  for (int k = 0; k < n_chunks; ++k) {
    const int p0 = (int) ((long long) n_pairs * k / n_chunks), p1 = (int) ((long long) n_pairs * (k + 1) / n_chunks);
    const size_t nf = (size_t) (foff[p1] - foff[p0]), nm = (size_t) (moff[p1] - moff[p0]);
    // (pageable caller buffers go through the handle's pinned ring with a few copy threads, ls2d_stage.h)
    if (nf) CU(h->stage.upload(F.pts + foff[p0], fpts + 4 * (size_t) foff[p0], nf * sizeof(float4), h->copy_stream));
    if (nm) CU(h->stage.upload(M.pts + moff[p0], mpts + 4 * (size_t) moff[p0], nm * sizeof(float4), h->copy_stream));
    CU(cudaEventRecord(h->ev_chunk[k], h->copy_stream));
    CU(cudaStreamWaitEvent(h->stream, h->ev_chunk[k], 0));
    align_args a = base_args(h);
    a.init_pose  = (const float*) h->d_init.p;
    a.out        = (ls2d_result*) h->d_out.p;
    a.n_pairs    = p1 - p0;
    a.pair_base  = p0;
    if ((rc = launch_icp(h, a))) return rc;
  }
  CU(cudaMemcpyAsync(out, h->d_out.p, sizeof(ls2d_result) * (size_t) n_pairs, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return LS2D_OK;
}

// ---- multi-slice aligner ---------------------------------------------------------------------------
static int multi_dev_impl(ls2d_handle* h, const ls2d_params* slices, const int32_t* fset, const int32_t* mset,
                          int32_t n_slices, const ls2d_prior* prior, const float* prior_z, const int32_t* fid,
                          const int32_t* mid, const float* init, int32_t n_pairs, ls2d_result* out,
                          ls2d_iter_stats* iters) {
  if (!h || !slices || !fset || !mset || n_slices < 1 || n_slices > LS2D_MAX_SLICES || !init || !out || n_pairs < 0)
    return LS2D_ERR_INVALID;
  if ((prior == nullptr) != (prior_z == nullptr)) return LS2D_ERR_INVALID;
  multi_args a;
  memset(&a, 0, sizeof(a));
  int cols[LS2D_MAX_SLICES];
  for (int s = 0; s < n_slices; ++s) {
    if (!params_valid(slices[s])) return LS2D_ERR_INVALID;
    if (fset[s] < 0 || fset[s] >= LS2D_MAX_CLOUD_SETS || mset[s] < 0 || mset[s] >= LS2D_MAX_CLOUD_SETS)
      return LS2D_ERR_INVALID;
    const cloud_set& f = h->sets[fset[s]];
    const cloud_set& m = h->sets[mset[s]];
    if (!f.pts || !f.off || !m.pts || !m.off) return LS2D_ERR_NOT_READY;
    a.sl[s].P          = translate(slices[s]);
    if (int rc = slice_edges(h, s, a.sl[s].P.cam)) return rc;
    a.sl[s].fixed_pts  = f.pts;
    a.sl[s].fixed_off  = f.off;
    a.sl[s].moving_pts = m.pts;
    a.sl[s].moving_off = m.off;
    cols[s]            = slices[s].canvas_cols;
    if (cols[s] > a.max_cols) a.max_cols = cols[s];
    if (f.max_points > a.max_points) a.max_points = f.max_points;
    if (m.max_points > a.max_points) a.max_points = m.max_points;
    if (f.max_points > a.max_fixed_points) a.max_fixed_points = f.max_points;
  }
  if (a.max_cols > 0xFFFE) return LS2D_ERR_UNSUPPORTED;
  a.n_slices    = n_slices;
  a.fused       = 1;
  for (int s = 0; s < n_slices; ++s)
    if (slices[s].single_rounding_accumulation) a.fused = 0;
  a.fixed_id    = fid;
  a.moving_id   = mid;
  a.init_pose   = init;
  a.pose_stride = pose_stride(h);
  a.prior_z     = prior_z;
  if (prior) {
    memcpy(a.prior_info, prior->information, sizeof(float) * 6);
    a.prior_tau     = prior->cauchy_chi_threshold;
    a.prior_inv_tau = prior->cauchy_chi_threshold > 0.f ? 1.f / prior->cauchy_chi_threshold : 0.f;
  }
  a.out        = out;
  a.iters      = iters;
  a.n_pairs    = n_pairs;
  a.score_only = 0;
  if (n_pairs == 0) return LS2D_OK;
  for (int s = 0; s < n_slices; ++s)  // the multi-slice kernels run Gauss-Newton rounds only
    if (slices[s].algorithm != LS2D_ALGORITHM_GN || slices[s].enable_inlier_only_runs || slices[s].termination_epsilon > 0.f)
      return LS2D_ERR_UNSUPPORTED;
  CU(cudaSetDevice(h->device));
  return launch_multi(h, a, cols);
}

int ls2d_align_multi_dev(ls2d_handle* h, const ls2d_params* slices, const int32_t* fset, const int32_t* mset,
                         int32_t n_slices, const ls2d_prior* prior, const float* prior_z, const int32_t* fid,
                         const int32_t* mid, const float* init, int32_t n_pairs, ls2d_result* out,
                         ls2d_iter_stats* iters) {
  return multi_dev_impl(h, slices, fset, mset, n_slices, prior, prior_z, fid, mid, init, n_pairs, out, iters);
}

int ls2d_align_multi(ls2d_handle* h, const ls2d_params* slices, const int32_t* fset, const int32_t* mset,
                     int32_t n_slices, const ls2d_prior* prior, const float* prior_z, const int32_t* fid,
                     const int32_t* mid, const float* init, int32_t n_pairs, ls2d_result* out,
                     ls2d_iter_stats* iters) {
  if (!h || !slices || !fset || !mset || n_slices < 1 || n_slices > LS2D_MAX_SLICES || !init || !out || n_pairs < 0)
    return LS2D_ERR_INVALID;
  if (n_pairs == 0) return LS2D_OK;
  for (int s = 0; s < n_slices; ++s) {
    if (fset[s] < 0 || fset[s] >= LS2D_MAX_CLOUD_SETS || mset[s] < 0 || mset[s] >= LS2D_MAX_CLOUD_SETS)
      return LS2D_ERR_INVALID;
    for (int i = 0; i < n_pairs; ++i) {
      const int f = fid ? fid[i] : i, m = mid ? mid[i] : i;
      if (f < 0 || f >= h->sets[fset[s]].n_clouds || m < 0 || m >= h->sets[mset[s]].n_clouds) return LS2D_ERR_INVALID;
    }
  }
  CU(cudaSetDevice(h->device));
  int rc;
  if (fid && (rc = h2d(h, h->d_fid, fid, sizeof(int) * (size_t) n_pairs))) return rc;
  if (mid && (rc = h2d(h, h->d_mid, mid, sizeof(int) * (size_t) n_pairs))) return rc;
  if ((rc = h2d(h, h->d_init, init, sizeof(float) * pose_stride(h) * (size_t) n_pairs))) return rc;
  if (prior_z && (rc = h2d(h, h->d_prior, prior_z, sizeof(float) * pose_stride(h) * (size_t) n_pairs))) return rc;
  if ((rc = reserve(h->d_out, sizeof(ls2d_result) * (size_t) n_pairs))) return rc;
  const size_t iter_bytes = sizeof(ls2d_iter_stats) * (size_t) n_pairs * (size_t) slices[0].max_iterations;
  if (iters && iter_bytes) {
    if ((rc = reserve(h->d_iters, iter_bytes))) return rc;
    CU(cudaMemsetAsync(h->d_iters.p, 0, iter_bytes, h->stream));
  }
  rc = multi_dev_impl(h, slices, fset, mset, n_slices, prior, prior_z ? (const float*) h->d_prior.p : nullptr,
                      fid ? (const int*) h->d_fid.p : nullptr, mid ? (const int*) h->d_mid.p : nullptr,
                      (const float*) h->d_init.p, n_pairs, (ls2d_result*) h->d_out.p,
                      (iters && iter_bytes) ? (ls2d_iter_stats*) h->d_iters.p : nullptr);
  if (rc) return rc;
  CU(cudaMemcpyAsync(out, h->d_out.p, sizeof(ls2d_result) * (size_t) n_pairs, cudaMemcpyDeviceToHost, h->stream));
  if (iters && iter_bytes) CU(cudaMemcpyAsync(iters, h->d_iters.p, iter_bytes, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return LS2D_OK;
}

int ls2d_find_correspondences(ls2d_handle* h, int32_t fixed_id, int32_t moving_id, const float* xyt,
                              int32_t* fixed_idx, int32_t* moving_idx, int32_t* n_out) {
  return ls2d_find_correspondences_in(h, LS2D_FIXED, LS2D_MOVING, fixed_id, moving_id, xyt, fixed_idx, moving_idx, n_out);
}

int ls2d_find_correspondences_in(ls2d_handle* h, int32_t fixed_set, int32_t moving_set, int32_t fixed_id,
                                 int32_t moving_id, const float* xyt, int32_t* fixed_idx, int32_t* moving_idx,
                                 int32_t* n_out) {
  if (!h || !xyt || !fixed_idx || !moving_idx || !n_out) return LS2D_ERR_INVALID;
  if (fixed_set < 0 || fixed_set >= LS2D_MAX_CLOUD_SETS || moving_set < 0 || moving_set >= LS2D_MAX_CLOUD_SETS)
    return LS2D_ERR_INVALID;
  const cloud_set& fs = h->sets[fixed_set];
  const cloud_set& ms = h->sets[moving_set];
  if (!fs.pts || !fs.off || !ms.pts || !ms.off) return LS2D_ERR_NOT_READY;
  if (fixed_id < 0 || fixed_id >= fs.n_clouds || moving_id < 0 || moving_id >= ms.n_clouds) return LS2D_ERR_INVALID;
  CU(cudaSetDevice(h->device));
  const int C = h->dp.cam.cols;
  int rc      = reserve(h->d_misc, sizeof(int) * (2 * (size_t) C + 1));
  if (rc) return rc;
  correspond_args a;
  a.fixed_pts    = fs.pts;
  a.fixed_off    = fs.off;
  a.moving_pts   = ms.pts;
  a.moving_off   = ms.off;
  a.fixed_cloud  = fixed_id;
  a.moving_cloud = moving_id;
  a.pose_stride = pose_stride(h);
  memset(a.lmis_pose, 0, sizeof(a.lmis_pose));
  memcpy(a.lmis_pose, xyt, sizeof(float) * a.pose_stride);
  a.fixed_idx  = (int*) h->d_misc.p;
  a.moving_idx = a.fixed_idx + C;
  a.count      = a.moving_idx + C;
  if ((rc = launch_correspond(h, a))) return rc;
  int n = 0;
  CU(cudaMemcpyAsync(&n, a.count, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  if (n < 0 || n > C) return LS2D_ERR_CUDA;
  if (n) {
    CU(cudaMemcpyAsync(fixed_idx, a.fixed_idx, sizeof(int) * n, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaMemcpyAsync(moving_idx, a.moving_idx, sizeof(int) * n, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
  }
  *n_out = n;
  return LS2D_OK;
}

int ls2d_classify_correspondences(ls2d_handle* h, int32_t fixed_set, int32_t moving_set, int32_t fixed_id,
                                  int32_t moving_id, const float* X_pose, const int32_t* fixed_idx,
                                  const int32_t* moving_idx, int32_t n, uint8_t* is_inlier) {
  if (!h || !X_pose || n < 0 || (n > 0 && (!fixed_idx || !moving_idx || !is_inlier))) return LS2D_ERR_INVALID;
  if (fixed_set < 0 || fixed_set >= LS2D_MAX_CLOUD_SETS || moving_set < 0 || moving_set >= LS2D_MAX_CLOUD_SETS)
    return LS2D_ERR_INVALID;
  const cloud_set& fs = h->sets[fixed_set];
  const cloud_set& ms = h->sets[moving_set];
  if (!fs.pts || !fs.off || !ms.pts || !ms.off) return LS2D_ERR_NOT_READY;
  if (fixed_id < 0 || fixed_id >= fs.n_clouds || moving_id < 0 || moving_id >= ms.n_clouds) return LS2D_ERR_INVALID;
  if (n == 0) return LS2D_OK;
  CU(cudaSetDevice(h->device));
  int rc;
  if ((rc = reserve(h->d_misc, (sizeof(int) * 2 + 1) * (size_t) n + 16))) return rc;
  int* d_fi          = (int*) h->d_misc.p;
  int* d_mi          = d_fi + n;
  unsigned char* d_o = (unsigned char*) (d_mi + n);
  CU(cudaMemcpyAsync(d_fi, fixed_idx, sizeof(int) * (size_t) n, cudaMemcpyHostToDevice, h->stream));
  CU(cudaMemcpyAsync(d_mi, moving_idx, sizeof(int) * (size_t) n, cudaMemcpyHostToDevice, h->stream));
  classify_args a;
  memset(&a, 0, sizeof(a));
  a.fixed_pts   = fs.pts;
  a.moving_pts  = ms.pts;
  a.pose_stride = pose_stride(h);
  memcpy(a.X_pose, X_pose, sizeof(float) * a.pose_stride);
  a.fixed_idx  = d_fi;
  a.moving_idx = d_mi;
  a.n          = n;
  a.is_inlier  = d_o;
  if ((rc = launch_classify(h, a, fs.off, fixed_id, ms.off, moving_id))) return rc;
  CU(cudaMemcpyAsync(is_inlier, d_o, (size_t) n, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return LS2D_OK;
}

int ls2d_project(ls2d_handle* h, int which, int32_t cloud_id, const float* cam_xyt, int32_t* source_idx,
                 float* depth) {
  if (!h || which < 0 || which >= LS2D_MAX_CLOUD_SETS || !cam_xyt || !source_idx || !depth) return LS2D_ERR_INVALID;
  const cloud_set& c = h->sets[which];
  if (!c.pts || !c.off) return LS2D_ERR_NOT_READY;
  if (cloud_id < 0 || cloud_id >= c.n_clouds) return LS2D_ERR_INVALID;
  CU(cudaSetDevice(h->device));
  const int C = h->dp.cam.cols;
  int rc      = reserve(h->d_misc, sizeof(int) * (2 * (size_t) C + 1));
  if (rc) return rc;
  project_args a;
  a.pts   = c.pts;
  a.off   = c.off;
  a.cloud = cloud_id;
  a.pose_stride = pose_stride(h);
  memset(a.cam_pose, 0, sizeof(a.cam_pose));
  memcpy(a.cam_pose, cam_xyt, sizeof(float) * a.pose_stride);
  a.source_idx = (int*) h->d_misc.p;
  a.depth      = (float*) (a.source_idx + C);
  if ((rc = launch_project(h, a))) return rc;
  CU(cudaMemcpyAsync(source_idx, a.source_idx, sizeof(int) * C, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaMemcpyAsync(depth, a.depth, sizeof(float) * C, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return LS2D_OK;
}

int ls2d_verify_dev(ls2d_handle* h, int32_t query_id, const int32_t* cand_dev, int32_t n_cand,
                    const float* guesses_dev, int32_t n_guess, const ls2d_gates* gates, int32_t cand_base,
                    ls2d_best* best_dev, ls2d_result* all_dev) {
  if (!h || !guesses_dev || !gates || !best_dev || n_cand < 0 || n_guess < 1) return LS2D_ERR_INVALID;
  if (!ready(h)) return LS2D_ERR_NOT_READY;
  if (query_id < 0 || query_id >= h->sets[0].n_clouds) return LS2D_ERR_INVALID;
  if ((int64_t) n_cand * n_guess > 0x7fffffff) return LS2D_ERR_INVALID;
  CU(cudaSetDevice(h->device));
  const int n = n_cand * n_guess;
  int rc;
  if (!all_dev) {
    if ((rc = reserve(h->d_out, sizeof(ls2d_result) * (size_t) (n > 0 ? n : 1)))) return rc;
    all_dev = (ls2d_result*) h->d_out.p;
  }
  align_args a  = base_args(h);
  a.fixed_const = query_id;
  a.moving_id   = cand_dev;
  a.moving_div  = n_guess;
  a.init_pose   = guesses_dev;
  a.out         = all_dev;
  a.n_pairs     = n;
  if ((rc = launch_icp(h, a))) return rc;
  return launch_best_of(h, all_dev, n, n_guess, *gates, cand_base, best_dev);
}

int ls2d_verify(ls2d_handle* h, int32_t query_id, const int32_t* cand, int32_t n_cand, const float* guesses,
                int32_t n_guess, const ls2d_gates* gates, int32_t cand_base, ls2d_best* best,
                ls2d_result* all) {
  if (!h || !guesses || !gates || !best || n_cand < 0 || n_guess < 1) return LS2D_ERR_INVALID;
  if (!ready(h)) return LS2D_ERR_NOT_READY;
  if ((int64_t) n_cand * n_guess > 0x7fffffff) return LS2D_ERR_INVALID;
  if (cand)
    for (int i = 0; i < n_cand; ++i)
      if (cand[i] < 0 || cand[i] >= h->sets[1].n_clouds) return LS2D_ERR_INVALID;
  if (!cand && n_cand > h->sets[1].n_clouds) return LS2D_ERR_INVALID;
  CU(cudaSetDevice(h->device));
  const size_t n = (size_t) n_cand * n_guess;
  int rc;
  if (cand && n_cand && (rc = h2d(h, h->d_mid, cand, sizeof(int) * (size_t) n_cand))) return rc;
  if ((rc = h2d(h, h->d_init, guesses, sizeof(float) * pose_stride(h) * (n ? n : 1)))) return rc;
  if ((rc = reserve(h->d_out, sizeof(ls2d_result) * (n ? n : 1)))) return rc;
  if ((rc = reserve(h->d_best, sizeof(ls2d_best)))) return rc;
  rc = ls2d_verify_dev(h, query_id, cand ? (const int*) h->d_mid.p : nullptr, n_cand, (const float*) h->d_init.p,
                       n_guess, gates, cand_base, (ls2d_best*) h->d_best.p, (ls2d_result*) h->d_out.p);
  if (rc) return rc;
  CU(cudaMemcpyAsync(best, h->d_best.p, sizeof(ls2d_best), cudaMemcpyDeviceToHost, h->stream));
  if (all && n) CU(cudaMemcpyAsync(all, h->d_out.p, sizeof(ls2d_result) * n, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return LS2D_OK;
}

int ls2d_verify_pairs_dev(ls2d_handle* h, const int32_t* fid_dev, const int32_t* mid_dev, const float* guesses_dev,
                          int32_t n_pairs, const int32_t* group_off_dev, int32_t n_groups, const ls2d_gates* gates,
                          ls2d_best* best_dev, ls2d_result* all_dev) {
  if (!h || !guesses_dev || !gates || !best_dev || !group_off_dev || n_pairs < 0 || n_groups < 0) return LS2D_ERR_INVALID;
  if (!ready(h)) return LS2D_ERR_NOT_READY;
  CU(cudaSetDevice(h->device));
  int rc;
  if (!all_dev) {
    if ((rc = reserve(h->d_out, sizeof(ls2d_result) * (size_t) (n_pairs > 0 ? n_pairs : 1)))) return rc;
    all_dev = (ls2d_result*) h->d_out.p;
  }
  align_args a = base_args(h);
  a.fixed_id   = fid_dev;
  a.moving_id  = mid_dev;
  a.init_pose  = guesses_dev;
  a.out        = all_dev;
  a.n_pairs    = n_pairs;
  if ((rc = launch_icp(h, a))) return rc;
  return launch_best_of_groups(h, all_dev, group_off_dev, n_groups, mid_dev, *gates, best_dev);
}

int ls2d_verify_pairs(ls2d_handle* h, const int32_t* fid, const int32_t* mid, const float* guesses, int32_t n_pairs,
                      const int32_t* group_off, int32_t n_groups, const ls2d_gates* gates, ls2d_best* best,
                      ls2d_result* all) {
  if (!h || !guesses || !gates || !best || !group_off || n_pairs < 0 || n_groups < 0) return LS2D_ERR_INVALID;
  if (!ready(h)) return LS2D_ERR_NOT_READY;
  if (group_off[0] != 0 || group_off[n_groups] != n_pairs) return LS2D_ERR_INVALID;
  for (int g = 0; g < n_groups; ++g)
    if (group_off[g + 1] < group_off[g]) return LS2D_ERR_INVALID;
  int rc = check_ids(h, fid, mid, n_pairs);
  if (rc) return rc;
  if (n_groups == 0) return LS2D_OK;
  CU(cudaSetDevice(h->device));
  const size_t n = (size_t) n_pairs;
  if (fid && n && (rc = h2d(h, h->d_fid, fid, sizeof(int) * n))) return rc;
  if (mid && n && (rc = h2d(h, h->d_mid, mid, sizeof(int) * n))) return rc;
  if ((rc = h2d(h, h->d_init, guesses, sizeof(float) * pose_stride(h) * (n ? n : 1)))) return rc;
  if ((rc = h2d(h, h->d_misc, group_off, sizeof(int) * ((size_t) n_groups + 1)))) return rc;
  if ((rc = reserve(h->d_out, sizeof(ls2d_result) * (n ? n : 1)))) return rc;
  if ((rc = reserve(h->d_best, sizeof(ls2d_best) * (size_t) n_groups))) return rc;
  rc = ls2d_verify_pairs_dev(h, fid ? (const int*) h->d_fid.p : nullptr, mid ? (const int*) h->d_mid.p : nullptr,
                             (const float*) h->d_init.p, n_pairs, (const int*) h->d_misc.p, n_groups, gates,
                             (ls2d_best*) h->d_best.p, (ls2d_result*) h->d_out.p);
  if (rc) return rc;
  CU(cudaMemcpyAsync(best, h->d_best.p, sizeof(ls2d_best) * (size_t) n_groups, cudaMemcpyDeviceToHost, h->stream));
  if (all && n) CU(cudaMemcpyAsync(all, h->d_out.p, sizeof(ls2d_result) * n, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return LS2D_OK;
}

int ls2d_reduce_best(const ls2d_best* rec, int32_t n, ls2d_best* out) {
  if (!rec || !out || n < 0) return LS2D_ERR_INVALID;
  ls2d_best b;
  memset(&b, 0, sizeof(b));
  b.candidate = -1;
  b.guess     = -1;
  b.c         = 1.f;  // "nothing accepted" carries the identity, like best_of_kernel's empty record
  for (int i = 0; i < n; ++i) {
    const ls2d_best& r = rec[i];
    if (r.candidate < 0 || r.n_inliers <= 0) continue;
    if (b.candidate < 0) {
      b = r;
      continue;
    }
    const float cr = r.chi_inliers / (float) r.n_inliers, cb = b.chi_inliers / (float) b.n_inliers;
    bool take      = false;
    if (r.n_inliers != b.n_inliers)
      take = r.n_inliers > b.n_inliers;
    else if (cr != cb)
      take = cr < cb;
    else if (r.candidate != b.candidate)
      take = r.candidate < b.candidate;
    else
      take = r.guess < b.guess;
    if (take) b = r;
  }
  *out = b;
  return LS2D_OK;
}

int ls2d_verify_sharded_nccl(ls2d_handle* h, int32_t query_id, const int32_t* cand_dev, int32_t n_cand,
                             const float* guesses_dev, int32_t n_guess, const ls2d_gates* gates,
                             int32_t cand_base, void* comm, int32_t n_ranks, ls2d_best* best) {
  if (!h || !comm || !best || n_ranks < 1) return LS2D_ERR_INVALID;
  CU(cudaSetDevice(h->device));  // before any allocation: the gather buffer must live on the handle's device
  if (!h->nccl_all_gather) {
    // the NCCL the process already runs (torch bundles its own) before any system copy: a communicator only makes
    // sense to the library that created it
    void* sym = dlsym(RTLD_DEFAULT, "ncclAllGather");
    void* cnt = dlsym(RTLD_DEFAULT, "ncclCommCount");
    if (!sym || !cnt) {
      h->nccl_lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
      if (!h->nccl_lib) h->nccl_lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
      if (!h->nccl_lib) return LS2D_ERR_NCCL;
      sym = dlsym(h->nccl_lib, "ncclAllGather");
      cnt = dlsym(h->nccl_lib, "ncclCommCount");
    }
    if (!sym || !cnt) return LS2D_ERR_NCCL;
    h->nccl_all_gather = (int (*)(const void*, void*, size_t, int, void*, cudaStream_t)) sym;
    h->nccl_comm_count = (int (*)(void*, int*)) cnt;
  }
  int comm_size = 0;
  if (h->nccl_comm_count(comm, &comm_size) != 0) return LS2D_ERR_NCCL;
  if (comm_size != n_ranks) return LS2D_ERR_INVALID;  // the all-gather writes comm_size records
  int rc;
  if ((rc = reserve(h->d_best, sizeof(ls2d_best) * (size_t) (n_ranks + 1)))) return rc;
  ls2d_best* mine   = (ls2d_best*) h->d_best.p;
  ls2d_best* gather = mine + 1;
  if ((rc = ls2d_verify_dev(h, query_id, cand_dev, n_cand, guesses_dev, n_guess, gates, cand_base, mine, nullptr)))
    return rc;
  constexpr int kNcclInt8 = 0;  // ncclDataType_t::ncclInt8 (nccl.h); the records travel as bytes
  if (h->nccl_all_gather(mine, gather, sizeof(ls2d_best), kNcclInt8, comm, h->stream) != 0) return LS2D_ERR_NCCL;
  std::vector<ls2d_best> host((size_t) n_ranks);
  CU(cudaMemcpyAsync(host.data(), gather, sizeof(ls2d_best) * (size_t) n_ranks, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return ls2d_reduce_best(host.data(), n_ranks, best);
}

int ls2d_merge_scene_dev(ls2d_handle* h, void* scene_dev, int32_t* scene_size_dev, int32_t capacity,
                         const void* meas_dev, int32_t n_meas, const float* mis_xyt, float merge_threshold,
                         int32_t* counters_dev) {
  if (!h || !scene_dev || !scene_size_dev || capacity < 0 || (!meas_dev && n_meas > 0) || n_meas < 0 || !mis_xyt)
    return LS2D_ERR_INVALID;
  CU(cudaSetDevice(h->device));
  merge_args a;
  a.scene      = (float4*) scene_dev;
  a.scene_size = scene_size_dev;
  a.capacity   = capacity;
  a.meas       = (const float4*) meas_dev;
  a.n_meas     = n_meas;
  a.pose_stride = pose_stride(h);
  memset(a.mis_pose, 0, sizeof(a.mis_pose));
  memcpy(a.mis_pose, mis_xyt, sizeof(float) * a.pose_stride);
  a.merge_threshold = merge_threshold;
  a.counters        = counters_dev;
  return launch_merge(h, a);
}

int ls2d_merge_scene(ls2d_handle* h, float* scene, int32_t* scene_size, int32_t capacity, const float* meas,
                     int32_t n_meas, const float* mis_xyt, float merge_threshold, int32_t* counters) {
  if (!h || !scene || !scene_size || !mis_xyt || n_meas < 0 || (!meas && n_meas > 0)) return LS2D_ERR_INVALID;
  if (*scene_size < 0 || capacity < *scene_size + h->dp.cam.cols) return LS2D_ERR_INVALID;
  CU(cudaSetDevice(h->device));
  int rc;
  const size_t scene_bytes = sizeof(float4) * (size_t) capacity;
  const size_t meas_bytes  = sizeof(float4) * (size_t) n_meas;
  if ((rc = reserve(h->d_misc, scene_bytes + meas_bytes + 64))) return rc;
  char* base     = (char*) h->d_misc.p;
  float4* d_scene = (float4*) base;
  float4* d_meas  = (float4*) (base + scene_bytes);
  int* d_ints     = (int*) (base + scene_bytes + meas_bytes);  // [0] = size, [1..4] = counters
  CU(cudaMemcpyAsync(d_scene, scene, sizeof(float4) * (size_t) *scene_size, cudaMemcpyHostToDevice, h->stream));
  if (n_meas) CU(cudaMemcpyAsync(d_meas, meas, meas_bytes, cudaMemcpyHostToDevice, h->stream));
  CU(cudaMemcpyAsync(d_ints, scene_size, sizeof(int), cudaMemcpyHostToDevice, h->stream));
  if ((rc = ls2d_merge_scene_dev(h, d_scene, d_ints, capacity, d_meas, n_meas, mis_xyt, merge_threshold, d_ints + 1)))
    return rc;
  int ints[5];
  CU(cudaMemcpyAsync(ints, d_ints, sizeof(ints), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  if (ints[4]) return LS2D_ERR_INVALID;  // cannot happen with the capacity check above
  CU(cudaMemcpyAsync(scene, d_scene, sizeof(float4) * (size_t) ints[0], cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  *scene_size = ints[0];
  if (counters) counters[0] = ints[1], counters[1] = ints[2], counters[2] = ints[3];
  return LS2D_OK;
}

// ---- RawDataPreprocessorProjective2D ------------------------------------------------
void ls2d_default_scan_params(ls2d_scan_params* p) {
  if (!p) return;
  p->angle_min             = -2.34747f;
  p->angle_max             = 2.35619f;
  p->msg_range_min         = 0.f;
  p->msg_range_max         = 30.f;
  p->range_min             = 0.f;
  p->range_max             = 1000.f;
  p->voxelize_resolution   = 0.02f;
  p->normal_point_distance = 0.3f;
  p->normal_min_points     = 5;
}

// ranges_dev -> strided points + counts on the device (asynchronous)
static int preprocess_dev(ls2d_handle* h, const ls2d_scan_params* sp, const float* ranges_dev, int32_t n_beams,
                          int32_t n_scans, float4* out_dev, int* counts_dev, int* off_dev, int continues = 0) {
  if (n_beams < 1 || n_beams > 8192 || !(sp->angle_max > sp->angle_min)) return LS2D_ERR_INVALID;
  scan_dev_params P;
  // _processLaserMessage, raw_data_preprocessor_projective_2d.cpp:83-90 (single binary32 operations)
  P.range_max            = sp->msg_range_max < sp->range_max ? sp->msg_range_max : sp->range_max;
  P.range_min            = sp->msg_range_min > sp->range_min ? sp->msg_range_min : sp->range_min;
  const float sensor_res = (sp->angle_max - sp->angle_min) / (float) n_beams;
  const float fx         = 1.f / sensor_res;
  P.ifx                  = 1.f / fx;
  P.cx                   = (float) n_beams / 2.f;
  P.d2                   = sp->normal_point_distance * sp->normal_point_distance;
  P.inv_res              = sp->voxelize_resolution > 0.f ? 1.f / sp->voxelize_resolution : 0.f;
  P.min_points           = sp->normal_min_points;
  P.n_beams              = n_beams;
  if (P.inv_res != 0.f) {
    // the packed voxel key holds |coordinate / res| < 2^19 (coordinates are bounded by the range limit)
    if (!(P.range_max * P.inv_res < 524288.f)) return LS2D_ERR_UNSUPPORTED;
  }
  scan_args a;
  a.ranges  = ranges_dev;
  a.beam_cs = nullptr;  // launch_preprocess: the handle's table
  a.out     = out_dev;
  a.counts  = counts_dev;
  a.n_scans = n_scans;
  a.off       = off_dev;
  a.continues = continues;
  a.ticket    = nullptr;  // launch_preprocess: the handle's counter
  return launch_preprocess(h, P, a, n_scans);
}

int ls2d_preprocess_scans(ls2d_handle* h, const ls2d_scan_params* sp, const float* ranges, int32_t n_beams,
                          int32_t n_scans, float* out_points, int32_t* out_counts) {
  if (!h || !sp || n_scans < 0 || n_beams < 1 || (n_scans > 0 && (!ranges || !out_points || !out_counts)))
    return LS2D_ERR_INVALID;
  if (n_scans == 0) return LS2D_OK;
  CU(cudaSetDevice(h->device));
  int rc;
  const size_t n_pts = (size_t) n_scans * n_beams;
  if ((rc = h2d(h, h->d_ranges, ranges, sizeof(float) * n_pts))) return rc;
  if ((rc = reserve(h->d_misc, sizeof(float4) * n_pts + sizeof(int) * (size_t) n_scans))) return rc;
  float4* d_out = (float4*) h->d_misc.p;
  int* d_cnt    = (int*) ((char*) h->d_misc.p + sizeof(float4) * n_pts);
  CU(cudaMemsetAsync(d_out, 0, sizeof(float4) * n_pts, h->stream));  // rows are copied out whole: zeros past the counts
  if ((rc = preprocess_dev(h, sp, (const float*) h->d_ranges.p, n_beams, n_scans, d_out, d_cnt, nullptr))) return rc;
  CU(cudaMemcpyAsync(out_points, d_out, sizeof(float4) * n_pts, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaMemcpyAsync(out_counts, d_cnt, sizeof(int) * (size_t) n_scans, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return LS2D_OK;
}

// an owned cloud set with room for n clouds of at most `stride` points
static int sized_set(ls2d_handle* h, int which, int32_t stride, int32_t n) {
  cloud_set& c = h->sets[which];
  if (!c.owned) c = cloud_set();
  c.owned            = true;
  const size_t n_pts = (size_t) n * stride;
  const size_t n_off = (size_t) n + 1;
  if (n_pts > c.cap_pts || !c.pts) {
    if (c.pts) cudaFree(c.pts);
    c.pts     = nullptr;
    c.cap_pts = 0;
    CU(cudaMalloc((void**) &c.pts, (n_pts + 16) * sizeof(float4)));
    c.cap_pts = n_pts + 16;
  }
  if (n_off > c.cap_off || !c.off) {
    if (c.off) cudaFree(c.off);
    c.off     = nullptr;
    c.cap_off = 0;
    CU(cudaMalloc((void**) &c.off, (n_off + 16) * sizeof(int)));
    c.cap_off = n_off + 16;
  }
  c.n_clouds   = n;
  c.max_points = stride;  // upper bound: the counts stay on the device
  if (n == 0) CU(cudaMemsetAsync(c.off, 0, sizeof(int), h->stream));
  return LS2D_OK;
}

// ... and the look-back words with which the n CTAs of the producing kernel agree on the CSR offsets
// (ls2d_service.cuh: lookback_exclusive)
static int packed_set(ls2d_handle* h, int which, int32_t stride, int32_t n, pack_target* pack) {
  int rc;
  if ((rc = sized_set(h, which, stride, n))) return rc;
  if (n == 0) return LS2D_OK;
  cloud_set& c = h->sets[which];
  const size_t words = sizeof(unsigned long long) * (size_t) n;
  const bool wrap    = h->pack_epoch >= 0x3FFFFFFEu;
  if (words > h->d_look.cap || wrap) {  // fresh words must not look like this or a later launch's
    if ((rc = reserve(h->d_look, words))) return rc;
    CU(cudaMemsetAsync(h->d_look.p, 0, h->d_look.cap, h->stream));
    if (wrap) h->pack_epoch = 0;
  }
  pack->packed = c.pts;
  pack->off    = c.off;
  pack->state  = (unsigned long long*) h->d_look.p;
  pack->epoch  = ++h->pack_epoch;
  return LS2D_OK;
}

static int scans_to_set_impl(ls2d_handle* h, int which, const ls2d_scan_params* sp, const float* ranges_dev,
                             int32_t n_beams, int32_t n_scans) {
  int rc;
  if ((rc = sized_set(h, which, n_beams, n_scans))) return rc;
  if (n_scans == 0) return LS2D_OK;
  const size_t n_pts = (size_t) n_scans * n_beams;
  if ((rc = reserve(h->d_misc, sizeof(float4) * n_pts + sizeof(int) * (size_t) n_scans))) return rc;
  float4* d_rows = (float4*) h->d_misc.p;
  int* d_cnt     = (int*) ((char*) h->d_misc.p + sizeof(float4) * n_pts);
  cloud_set& c   = h->sets[which];
  if ((rc = preprocess_dev(h, sp, ranges_dev, n_beams, n_scans, d_rows, d_cnt, c.off))) return rc;
  return launch_scan_pack(h, d_rows, c.off, n_beams, n_scans, c.pts);
}

int ls2d_preprocess_scans_to_set(ls2d_handle* h, int which, const ls2d_scan_params* sp, const float* ranges,
                                 int32_t n_beams, int32_t n_scans) {
  if (!h || which < 0 || which >= LS2D_MAX_CLOUD_SETS || !sp || n_scans < 0 || n_beams < 1 || (n_scans > 0 && !ranges))
    return LS2D_ERR_INVALID;
  CU(cudaSetDevice(h->device));
  int rc;
  if (n_scans > 0 && (rc = h2d(h, h->d_ranges, ranges, sizeof(float) * (size_t) n_scans * n_beams))) return rc;
  return scans_to_set_impl(h, which, sp, (const float*) h->d_ranges.p, n_beams, n_scans);
}

int ls2d_preprocess_scans_to_set_dev(ls2d_handle* h, int which, const ls2d_scan_params* sp, const float* ranges_dev,
                                     int32_t n_beams, int32_t n_scans) {
  if (!h || which < 0 || which >= LS2D_MAX_CLOUD_SETS || !sp || n_scans < 0 || n_beams < 1 || (n_scans > 0 && !ranges_dev))
    return LS2D_ERR_INVALID;
  CU(cudaSetDevice(h->device));
  return scans_to_set_impl(h, which, sp, ranges_dev, n_beams, n_scans);
}

int ls2d_download_clouds(ls2d_handle* h, int which, float* points, int32_t* offsets, int32_t n_clouds,
                         int64_t capacity_points) {
  if (!h || which < 0 || which >= LS2D_MAX_CLOUD_SETS || !offsets) return LS2D_ERR_INVALID;
  const cloud_set& c = h->sets[which];
  if (!c.pts || !c.off) return LS2D_ERR_NOT_READY;
  if (n_clouds != c.n_clouds) return LS2D_ERR_INVALID;
  CU(cudaSetDevice(h->device));
  CU(cudaMemcpyAsync(offsets, c.off, sizeof(int) * ((size_t) n_clouds + 1), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  const int64_t total = offsets[n_clouds];
  if (total > capacity_points || (total > 0 && !points)) return LS2D_ERR_INVALID;
  if (total > 0) {
    CU(cudaMemcpyAsync(points, c.pts, sizeof(float4) * (size_t) total, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
  }
  return LS2D_OK;
}

// clip_kernel into scratch `tmp` (strided rows of canvas_cols points + counts); ids / poses already on the device
// (strided rows of canvas_cols points + counts), or straight into a packed set when `pack` names one
static int clip_dev(ls2d_handle* h, const cloud_set& c, const int* ids_dev, const float* robot_dev,
                    const float* sensor_xyt, int32_t n, scratch& tmp, float4** out, int** counts,
                    float voxelize_resolution = 0.f, const pack_target* pack = nullptr) {
  const int C = h->dp.cam.cols;
  int rc;
  const size_t out_bytes = pack ? 0 : sizeof(float4) * (size_t) n * C;
  if (!pack && (rc = reserve(tmp, out_bytes + sizeof(int) * (size_t) n))) return rc;
  clip_args a;
  a.pts       = c.pts;
  a.off       = c.off;
  a.cloud_ids = ids_dev;
  a.robot_pose  = robot_dev;
  a.pose_stride = pose_stride(h);
  memset(a.sensor_pose, 0, sizeof(a.sensor_pose));
  memcpy(a.sensor_pose, sensor_xyt, sizeof(float) * a.pose_stride);
  a.out    = pack ? nullptr : (float4*) tmp.p;
  a.counts = pack ? nullptr : (int*) ((char*) tmp.p + out_bytes);
  if (!pack) CU(cudaMemsetAsync(tmp.p, 0, out_bytes, h->stream));  // rows are copied out whole: zeros past the counts
  a.base   = 0;
  a.pack   = pack ? *pack : pack_target{};
  if (voxelize_resolution > 0.f) {  // scene_clipper_projective_2d.cpp:36-48
    const float inv_res = 1.f / voxelize_resolution;
    // the packed voxel key holds |coordinate / res| < 2^19 (coordinates are bounded by the range limit)
    if (!(h->dp.range_max * inv_res < 524288.f)) return LS2D_ERR_UNSUPPORTED;
    if ((rc = launch_clip_voxel(h, a, n, inv_res))) return rc;
  } else {
    if ((rc = launch_clip(h, a, n))) return rc;
  }
  *out    = a.out;
  *counts = a.counts;
  return LS2D_OK;
}

int ls2d_clip_scenes_voxelized(ls2d_handle* h, int which, const int32_t* cloud_ids, const float* robot_xyt,
                               const float* sensor_xyt, int32_t n, float voxelize_resolution, float* out_points,
                               int32_t* out_counts) {
  if (!h || which < 0 || which >= LS2D_MAX_CLOUD_SETS || !cloud_ids || !robot_xyt || !sensor_xyt || !out_points || !out_counts || n < 0)
    return LS2D_ERR_INVALID;
  const cloud_set& c = h->sets[which];
  if (!c.pts || !c.off) return LS2D_ERR_NOT_READY;
  if (n == 0) return LS2D_OK;
  for (int i = 0; i < n; ++i)
    if (cloud_ids[i] < 0 || cloud_ids[i] >= c.n_clouds) return LS2D_ERR_INVALID;
  CU(cudaSetDevice(h->device));
  int rc;
  if ((rc = h2d(h, h->d_mid, cloud_ids, sizeof(int) * (size_t) n))) return rc;
  if ((rc = h2d(h, h->d_init, robot_xyt, sizeof(float) * pose_stride(h) * (size_t) n))) return rc;
  float4* d_out = nullptr;
  int* d_cnt    = nullptr;
  if ((rc = clip_dev(h, c, (const int*) h->d_mid.p, (const float*) h->d_init.p, sensor_xyt, n, h->d_clip, &d_out, &d_cnt,
                     voxelize_resolution)))
    return rc;
  CU(cudaMemcpyAsync(out_points, d_out, sizeof(float4) * (size_t) n * h->dp.cam.cols, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaMemcpyAsync(out_counts, d_cnt, sizeof(int) * (size_t) n, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return LS2D_OK;
}

int ls2d_clip_scenes(ls2d_handle* h, int which, const int32_t* cloud_ids, const float* robot_pose,
                     const float* sensor_pose, int32_t n, float* out_points, int32_t* out_counts) {
  return ls2d_clip_scenes_voxelized(h, which, cloud_ids, robot_pose, sensor_pose, n, 0.f, out_points, out_counts);
}

int ls2d_clip_scenes_to_set(ls2d_handle* h, int scene_set, const int32_t* cloud_ids, const float* robot_xyt,
                            const float* sensor_xyt, int32_t n, int out_set) {
  if (!h || scene_set < 0 || scene_set >= LS2D_MAX_CLOUD_SETS || out_set < 0 || out_set >= LS2D_MAX_CLOUD_SETS ||
      out_set == scene_set || !cloud_ids || !robot_xyt || !sensor_xyt || n < 0)
    return LS2D_ERR_INVALID;
  const cloud_set& c = h->sets[scene_set];
  if (!c.pts || !c.off) return LS2D_ERR_NOT_READY;
  for (int i = 0; i < n; ++i)
    if (cloud_ids[i] < 0 || cloud_ids[i] >= c.n_clouds) return LS2D_ERR_INVALID;
  CU(cudaSetDevice(h->device));
  int rc;
  float4* d_out = nullptr;
  int* d_cnt    = nullptr;
  const float4* src_pts = c.pts;  // packed_set() may touch h->sets[out_set] only: out_set != scene_set
  pack_target pack = {};
  if ((rc = packed_set(h, out_set, h->dp.cam.cols, n, &pack))) return rc;
  if (n == 0 || !src_pts) return LS2D_OK;
  if ((rc = h2d(h, h->d_mid, cloud_ids, sizeof(int) * (size_t) n))) return rc;
  if ((rc = h2d(h, h->d_init, robot_xyt, sizeof(float) * pose_stride(h) * (size_t) n))) return rc;
  return clip_dev(h, h->sets[scene_set], (const int*) h->d_mid.p, (const float*) h->d_init.p, sensor_xyt, n, h->d_clip,
                  &d_out, &d_cnt, 0.f, &pack);
}

// MultiTracker2D's frame step, batched: raw scan -> measurement cloud (fixed), local map seen from the predicted
// pose -> clipped scene in the robot frame (moving), MultiAligner2D; only ranges, ids and poses cross the bus.
int ls2d_track_batch(ls2d_handle* h, const ls2d_scan_params* sp, const float* ranges, int32_t n_beams, int32_t n,
                     int scene_set, const int32_t* scene_ids, const float* robot_xyt, const float* init_xyt,
                     ls2d_result* out) {
  if (!h || !sp || n < 0 || n_beams < 1 || scene_set < 2 || scene_set >= LS2D_MAX_CLOUD_SETS || !out ||
      (n > 0 && (!ranges || !scene_ids || !robot_xyt)))
    return LS2D_ERR_INVALID;
  if (n == 0) return LS2D_OK;
  int rc;
  // sensor_in_robot in the handle's pose format (the parameter record may hold either form)
  const int stride = pose_stride(h);
  const iso S      = h->prm.with_sensor ? sensor_iso(h->prm) : iso_identity();
  float sensor[4]  = {S.tx, S.ty, S.c, S.s};
  if (stride == 3) {
    if (h->prm.with_sensor == 2) return LS2D_ERR_INVALID;  // an isometry cannot be handed on as (x, y, theta) bit for bit
    sensor[2] = h->prm.with_sensor ? h->prm.sensor_in_robot[2] : 0.f;
  }
  const cloud_set& scene = h->sets[scene_set];
  if (!scene.pts || !scene.off) return LS2D_ERR_NOT_READY;
  for (int i = 0; i < n; ++i)
    if (scene_ids[i] < 0 || scene_ids[i] >= scene.n_clouds) return LS2D_ERR_INVALID;
  CU(cudaSetDevice(h->device));
  // The batch is cut into chunks of whole frames: chunk k + 1's ranges cross PCIe on the copy stream while chunk k
  // runs pre-processor -> clipper -> aligner on the compute stream.  The two cloud sets end up holding the whole
  // batch: the chunks' CSR offsets continue each other (scan_args::continues, the clipper's look-back words).
  const int C = h->dp.cam.cols;
  pack_target pack = {};
  if ((rc = sized_set(h, LS2D_FIXED, n_beams, n))) return rc;
  if ((rc = packed_set(h, LS2D_MOVING, C, n, &pack))) return rc;
  cloud_set& F = h->sets[LS2D_FIXED];
  const int wave    = 8 * (h->sm_count > 0 ? h->sm_count : 148);  // two rounds of 4 CTAs per SM (measured: 3 and 7 chunks lose)
  int n_chunks      = (n + wave - 1) / wave;
  n_chunks          = n_chunks > 8 ? 8 : n_chunks;
  const int chunk   = (n + n_chunks - 1) / n_chunks;
  const size_t pose_bytes = sizeof(float) * stride * (size_t) n;
  if ((rc = reserve(h->d_ranges, sizeof(float) * (size_t) n * n_beams))) return rc;
  if ((rc = reserve(h->d_mid, sizeof(int) * (size_t) n))) return rc;
  if ((rc = reserve(h->d_clip, pose_bytes))) return rc;
  if ((rc = reserve(h->d_init, pose_bytes))) return rc;
  if ((rc = reserve(h->d_out, sizeof(ls2d_result) * (size_t) n))) return rc;
  if ((rc = reserve(h->d_misc, sizeof(float4) * (size_t) chunk * n_beams + sizeof(int) * (size_t) n))) return rc;
  float4* d_rows = (float4*) h->d_misc.p;
  int* d_cnt     = (int*) ((char*) h->d_misc.p + sizeof(float4) * (size_t) chunk * n_beams);
  if (!h->copy_stream) {
    CU(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&h->ev_ready, cudaEventDisableTiming));
    for (cudaEvent_t& e : h->ev_chunk) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  }
  if (!h->aux_stream) {
    CU(cudaStreamCreateWithFlags(&h->aux_stream, cudaStreamNonBlocking));
    for (cudaEvent_t& e : h->ev_packed) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  }
  // Two compute lanes: the pre-processor of chunk k + 1 (aux stream) shares the GPU with the clipper and aligner of
  // chunk k (the handle's stream), so the partly filled last wave of one kernel is topped up by the other's CTAs.
  // Earlier work on the handle's stream may still read the buffers: uploads and the aux lane wait for it.
  cudaStream_t const main_stream = h->stream;
  CU(cudaEventRecord(h->ev_ready, main_stream));
  CU(cudaStreamWaitEvent(h->copy_stream, h->ev_ready, 0));
  CU(cudaStreamWaitEvent(h->aux_stream, h->ev_ready, 0));
  CU(cudaMemcpyAsync(h->d_mid.p, scene_ids, sizeof(int) * (size_t) n, cudaMemcpyHostToDevice, h->copy_stream));
  CU(cudaMemcpyAsync(h->d_clip.p, robot_xyt, pose_bytes, cudaMemcpyHostToDevice, h->copy_stream));
  if (init_xyt) {
    CU(cudaMemcpyAsync(h->d_init.p, init_xyt, pose_bytes, cudaMemcpyHostToDevice, h->copy_stream));
  } else {
    h->h_ident.assign((size_t) n * stride, 0.f);
    if (stride == 4)
      for (int i = 0; i < n; ++i) h->h_ident[(size_t) i * 4 + 2] = 1.f;
    CU(cudaMemcpyAsync(h->d_init.p, h->h_ident.data(), pose_bytes, cudaMemcpyHostToDevice, h->copy_stream));
  }
  for (int k = 0; k < n_chunks; ++k) {
    const int p0 = k * chunk, p1 = (k + 1) * chunk < n ? (k + 1) * chunk : n;
    if (p1 <= p0) break;
    CU(h->stage.upload((float*) h->d_ranges.p + (size_t) p0 * n_beams, ranges + (size_t) p0 * n_beams,
                       sizeof(float) * (size_t) (p1 - p0) * n_beams, h->copy_stream));
    CU(cudaEventRecord(h->ev_chunk[k], h->copy_stream));
    CU(cudaStreamWaitEvent(h->aux_stream, h->ev_chunk[k], 0));
    if (k == 0) CU(cudaStreamWaitEvent(main_stream, h->ev_chunk[0], 0));  // ids and poses
    h->stream = h->aux_stream;  // the launchers work on the handle's stream
    rc        = preprocess_dev(h, sp, (const float*) h->d_ranges.p + (size_t) p0 * n_beams, n_beams, p1 - p0, d_rows,
                               d_cnt + p0, F.off + p0, p0 > 0);
    if (!rc) rc = launch_scan_pack(h, d_rows, F.off + p0, n_beams, p1 - p0, F.pts);
    h->stream = main_stream;
    if (rc) return rc;
    CU(cudaEventRecord(h->ev_packed[k], h->aux_stream));
    clip_args c;
    c.pts         = scene.pts;
    c.off         = scene.off;
    c.cloud_ids   = (const int*) h->d_mid.p;
    c.robot_pose  = (const float*) h->d_clip.p;
    c.pose_stride = stride;
    memset(c.sensor_pose, 0, sizeof(c.sensor_pose));
    memcpy(c.sensor_pose, sensor, sizeof(float) * stride);
    c.out    = nullptr;
    c.counts = nullptr;
    c.base   = p0;
    c.pack   = pack;
    if ((rc = launch_clip(h, c, p1 - p0))) return rc;
    CU(cudaStreamWaitEvent(main_stream, h->ev_packed[k], 0));
    align_args a = base_args(h);
    a.init_pose  = (const float*) h->d_init.p;
    a.out        = (ls2d_result*) h->d_out.p;
    a.n_pairs    = p1 - p0;
    a.pair_base  = p0;
    if ((rc = launch_icp(h, a))) return rc;
  }
  CU(cudaMemcpyAsync(out, h->d_out.p, sizeof(ls2d_result) * (size_t) n, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return LS2D_OK;
}

int ls2d_multi_reduction_threads(void) { return multi_reduction_threads(); }

int ls2d_multi_reduction_shape(const ls2d_params* slices, int32_t n_slices, int32_t max_fixed_points,
                               int32_t max_moving_points, int32_t shared_moving) {
  if (!slices || n_slices < 1 || n_slices > LS2D_MAX_SLICES || max_fixed_points < 0 || max_moving_points < 0)
    return LS2D_ERR_INVALID;
  multi_args a;
  memset(&a, 0, sizeof(a));
  static const polar_edge some_table = {1.0, 0.0};  // the kernel choice only asks whether a slice HAS an edge table
  for (int s = 0; s < n_slices; ++s) {
    if (!params_valid(slices[s])) return LS2D_ERR_INVALID;
    a.sl[s].P          = translate(slices[s]);
    a.sl[s].P.cam.edge = &some_table;
    if (slices[s].canvas_cols > a.max_cols) a.max_cols = slices[s].canvas_cols;
  }
  a.n_slices         = n_slices;
  a.fused            = 1;
  for (int s = 0; s < n_slices; ++s)
    if (slices[s].single_rounding_accumulation) a.fused = 0;
  a.max_points       = max_fixed_points > max_moving_points ? max_fixed_points : max_moving_points;
  a.max_fixed_points = max_fixed_points;
  if (!shared_moving)
    for (int s = 1; s < n_slices; ++s) a.sl[s].moving_off = reinterpret_cast<const int*>(&some_table);  // "another set"
  return multi_reduction_shape(a);
}

int ls2d_reduction_shape(const ls2d_params* p, int32_t max_points) {
  if (!p || !params_valid(*p) || max_points < 0) return LS2D_ERR_INVALID;
  return icp_reduction_shape(translate(*p), p->single_rounding_accumulation != 0, max_points);
}

int ls2d_score_reduction_shape(const ls2d_params* p, int32_t max_points) {
  if (!p || !params_valid(*p) || max_points < 0) return LS2D_ERR_INVALID;
  const dev_params dp = translate(*p);
  const int shape     = score_reduction_shape(nullptr, dp, p->single_rounding_accumulation != 0, max_points);
  if (shape >= 0) return shape;
  dev_params gn = dp;  // the scoring pass never runs the general kernel: one linearisation of the size's aligner kernel
  gn.algorithm = LS2D_ALGORITHM_GN, gn.inlier_only_runs = 0, gn.termination_epsilon = 0.f;
  return icp_reduction_shape(gn, p->single_rounding_accumulation != 0, max_points);
}

int ls2d_selftest_gated_sqrt(ls2d_handle* h, float lo, float hi, int64_t* n_checked, int64_t* n_mismatch) {
  if (!h || !n_checked || !n_mismatch || !(lo > 0x1p-100f) || !(hi >= lo) || !(hi <= 3.0e38f)) return LS2D_ERR_INVALID;
  CU(cudaSetDevice(h->device));
  int rc;
  if ((rc = reserve(h->d_misc, 8))) return rc;
  CU(cudaMemsetAsync(h->d_misc.p, 0, 8, h->stream));
  unsigned lo_bits, hi_bits;
  memcpy(&lo_bits, &lo, 4), memcpy(&hi_bits, &hi, 4);
  if ((rc = launch_selftest_sqrt(h, lo_bits, hi_bits, (unsigned long long*) h->d_misc.p))) return rc;
  unsigned long long bad = 0;
  CU(cudaMemcpyAsync(&bad, h->d_misc.p, 8, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  *n_checked  = (int64_t) hi_bits - (int64_t) lo_bits + 1;
  *n_mismatch = (int64_t) bad;
  return LS2D_OK;
}

int64_t ls2d_launch_count(const ls2d_handle* h) { return h ? h->launches : 0; }

}  // extern "C"
