// ls2d_service.cuh -- kernels behind the drop-in projector / finder / clipper / merger APIs and the verification
// gates (latency-bound by construction; not on the headline bench path).
//
//  project_kernel     PointNormal2fProjectorPolar::compute for the drop-in projector API / parity.
//  correspond_kernel  CorrespondenceFinderProjective2f::compute for the drop-in finder API / parity.
//  clip_kernel        SceneClipperProjective2D::compute (voxelize_resolution == 0).
//  merge_kernel       MergerProjective2D::compute.
//  best_of_kernel / best_of_groups_kernel   acceptance gates + deterministic arg-best of a verification shard.
//
// Reference paths: R/ = /root/reference/srrg2_laser_slam_2d/src/srrg2_laser_slam_2d/.
#pragma once

#include "ls2d_common.cuh"

namespace ls2d {

// ordered block-wide compaction step: every thread calls it with its flag for column k0 + threadIdx.x; returns the
// output slot of flagged threads (ascending column order) and advances *base.  warp_tot: 32 ints of shared memory.
__device__ __forceinline__ int ordered_slot(bool ok, int* warp_tot, int* base) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const unsigned ballot = __ballot_sync(0xffffffffu, ok);
  if (lane == 0) warp_tot[warp] = __popc(ballot);
  __syncthreads();
  int before = *base;
  for (int w = 0; w < warp; ++w) before += warp_tot[w];
  const int dst = before + __popc(ballot & ((1u << lane) - 1u));
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = *base;
    for (int w = 0; w < nwarp; ++w) t += warp_tot[w];
    *base = t;
  }
  __syncthreads();
  return dst;
}

// Ordered block-wide compaction in two barriers.  Thread `tid` owns element c * T + tid of chunk c and passes its
// flags as a bit mask (bit c); slots come back through compact_slot().  cnt: n_chunks * (T / 32) ints of shared
// memory, total: one int.  Element order = chunk, warp, lane = ascending element index.
__device__ __forceinline__ void compact_count(unsigned mask, int n_chunks, int* cnt, int* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  for (int c = 0; c < n_chunks; ++c) {
    const unsigned ballot = __ballot_sync(0xffffffffu, (mask >> c) & 1u);
    if (lane == 0) cnt[c * nwarp + warp] = __popc(ballot);
  }
  __syncthreads();
  if (warp == 0) {  // exclusive prefix over the n_chunks * nwarp counts, in place
    const int n = n_chunks * nwarp, per = (n + 31) >> 5;
    int sum = 0;
    for (int k = lane * per; k < min(n, (lane + 1) * per); ++k) sum += cnt[k];
    int incl = sum;
    for (int o = 1; o < 32; o <<= 1) {
      const int u = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += u;
    }
    int run = incl - sum;
    for (int k = lane * per; k < min(n, (lane + 1) * per); ++k) {
      const int v = cnt[k];
      cnt[k]      = run;
      run += v;
    }
    if (lane == 31) *total = incl;
  }
  __syncthreads();
}
__device__ __forceinline__ int compact_slot(unsigned mask, int c, const int* cnt) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const unsigned ballot = __ballot_sync(0xffffffffu, (mask >> c) & 1u);
  return cnt[c * nwarp + warp] + __popc(ballot & ((1u << lane) - 1u));
}


// ---- ordered packing without a second pass: decoupled look-back over one 64-bit word per CTA --------------------
// word = (epoch << 2 | flag) << 32 | value; flag 1: value = this CTA's count, flag 2: value = the inclusive prefix up
// to and including this CTA.  The epoch (one per launch, kept by the handle) makes words of earlier launches read
// as "not there yet", so the array is never cleared.  CTAs wait only for lower block indices, which the hardware
// dispatches first.  Called by warp 0; returns the exclusive prefix of CTA b to all its lanes.

__device__ __forceinline__ int lookback_exclusive(unsigned long long* state, unsigned epoch, int b, int count) {
  const int lane = threadIdx.x & 31;
  const unsigned long long tag_count = (unsigned long long) (epoch << 2 | 1u) << 32;
  const unsigned long long tag_incl  = (unsigned long long) (epoch << 2 | 2u) << 32;
  volatile unsigned long long* st = state;
  if (b == 0) {
    if (lane == 0) st[0] = tag_incl | (unsigned) count;
    return 0;
  }
  if (lane == 0) st[b] = tag_count | (unsigned) count;
  int excl = 0;
  for (int base = b - 1;; base -= 32) {
    const int idx        = base - lane;  // lane 0 reads the nearest predecessor
    unsigned long long v = tag_incl;     // before the first CTA: an inclusive prefix of 0
    if (idx >= 0) {
      v = st[idx];
      while ((unsigned) (v >> 34) != epoch || ((unsigned) (v >> 32) & 3u) == 0u) {
        __nanosleep(32);
        v = st[idx];
      }
    }
    const unsigned incl = __ballot_sync(0xffffffffu, ((unsigned) (v >> 32) & 3u) == 2u);
    const int last      = incl ? __ffs(incl) - 1 : 31;  // sum up to and including the nearest inclusive prefix
    excl += __reduce_add_sync(0xffffffffu, lane <= last ? (int) (unsigned) v : 0);
    if (incl) break;
  }
  if (lane == 0) st[b] = tag_incl | (unsigned) (excl + count);
  return excl;
}

// Output rows of a CTA: `count` points from slot 0.  Every thread calls it (two barriers) once the count is known;
// returns where slot 0 goes.  `base_smem`: one int of shared memory.
// b: the request's index in the whole job (a job may be cut into several launches with one epoch: the look-back
// crosses the launches).
__device__ __forceinline__ float4* output_rows(const pack_target& T, float4* strided, size_t stride, int b, int count,
                                               int* counts, int* base_smem) {
  if (threadIdx.x == 0 && counts) counts[b] = count;
  if (!T.packed) return strided + (size_t) b * stride;
  if (threadIdx.x < 32) {
    const int excl = lookback_exclusive(T.state, T.epoch, b, count);
    if (threadIdx.x == 0) {
      *base_smem = excl;
      if (b == 0) T.off[0] = 0;
      T.off[b + 1] = excl + count;
    }
  }
  __syncthreads();
  float4* rows = T.packed + *base_smem;
  __syncthreads();  // base_smem may be reused
  return rows;
}


// ---------------------------------------------------------------------------------------------------
// z-buffer projection of an arbitrary-size cloud with strided loops (API / parity kernels).
// W = world -> camera isometry.  On return (after the trailing barrier) zidx[c] holds the winner of
// column c (Z_EMPTY_IDX if none) and zdepth[c] its rho bits.
template <bool IDENTITY, bool PLAIN_LOAD = false>
__device__ __forceinline__ void zbuffer_project(const dev_params& P, const iso& W, const float4* pts, int n,
                                                unsigned* zdepth, unsigned* zidx, const iso* pre = nullptr) {
  const int C = P.cam.cols;
  for (int k = threadIdx.x; k < C; k += blockDim.x) {
    zdepth[k] = Z_EMPTY_DEPTH;
    zidx[k]   = Z_EMPTY_IDX;
  }
  __syncthreads();
  for (int pass = 0; pass < 2; ++pass) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const float4 p = PLAIN_LOAD ? pts[i] : ldg4(pts + i);  // PLAIN_LOAD: the buffer is written later in this kernel
      float px = p.x, py = p.y;
      if (pre) {  // the cloud is first moved by *pre (merger: measurement -> scene frame), then seen from the camera
        float qx, qy;
        iso_apply(*pre, p.x, p.y, qx, qy);
        iso_apply(W, qx, qy, px, py);
      } else if (!IDENTITY) {
        iso_apply(W, p.x, p.y, px, py);
      }
      const float rho = fsqrt(fadd(fmul(px, px), fmul(py, py)));
      if (rho < P.range_min || rho > P.range_max) continue;
      const int col = polar_column(P.cam, py, px);
      if (col < 0) continue;
      if (pass == 0)
        atomicMin(&zdepth[col], f2u(rho));
      else if (zdepth[col] == f2u(rho))
        atomicMin(&zidx[col], (unsigned) i);
    }
    __syncthreads();
  }
}


__global__ void project_kernel(const dev_params P, const project_args A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int C      = P.cam.cols;
  unsigned* zdepth = reinterpret_cast<unsigned*>(smem_raw);
  unsigned* zidx   = zdepth + C;
  const int p0 = A.off[A.cloud], n = A.off[A.cloud + 1] - p0;
  const iso W = iso_inverse(load_pose(A.cam_pose, 0, A.pose_stride));
  zbuffer_project<false>(P, W, A.pts + p0, n, zdepth, zidx);
  for (int k = threadIdx.x; k < C; k += blockDim.x) {
    const bool empty = zidx[k] == Z_EMPTY_IDX;
    A.source_idx[k]  = empty ? -1 : (int) zidx[k];
    A.depth[k]       = empty ? FLT_MAX : u2f(zdepth[k]);
  }
}


// CorrespondenceFinderProjective2f::compute (R/registration/correspondence_finder_projective_2d.cpp:18-77)
__global__ void correspond_kernel(const dev_params P, const correspond_args A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int C   = P.cam.cols;
  unsigned* zdf = reinterpret_cast<unsigned*>(smem_raw);
  unsigned* zif = zdf + C;
  unsigned* zdm = zif + C;
  unsigned* zim = zdm + C;
  __shared__ int warp_tot[32];
  __shared__ int base;
  const int f0 = A.fixed_off[A.fixed_cloud], nf = A.fixed_off[A.fixed_cloud + 1] - f0;
  const int m0 = A.moving_off[A.moving_cloud], nm = A.moving_off[A.moving_cloud + 1] - m0;
  const iso L = load_pose(A.lmis_pose, 0, A.pose_stride);
  const iso W = iso_inverse(iso_inverse(L));  // .cpp:47 + the projector's own inverse (decision D13)
  zbuffer_project<true>(P, W, A.fixed_pts + f0, nf, zdf, zif);
  zbuffer_project<false>(P, W, A.moving_pts + m0, nm, zdm, zim);
  if (threadIdx.x == 0) base = 0;
  __syncthreads();
  for (int k0 = 0; k0 < C; k0 += blockDim.x) {  // ascending columns, ordered compaction (.cpp:55-74)
    const int k = k0 + threadIdx.x;
    bool ok     = false;
    int fi = -1, mi = -1;
    if (k < C && zif[k] != Z_EMPTY_IDX && zim[k] != Z_EMPTY_IDX) {
      fi = (int) zif[k], mi = (int) zim[k];
      ok = !(fabsf(fsub(u2f(zdf[k]), u2f(zdm[k]))) > P.point_distance);
      if (ok) {
        const float4 F = ldg4(A.fixed_pts + f0 + fi);
        const float4 M = ldg4(A.moving_pts + m0 + mi);
        float nx, ny;
        iso_rot(W, M.z, M.w, nx, ny);
        ok = !(fadd(fmul(nx, F.z), fmul(ny, F.w)) < P.normal_cos);
      }
    }
    const int dst = ordered_slot(ok, warp_tot, &base);
    if (ok) {
      A.fixed_idx[dst]  = fi;
      A.moving_idx[dst] = mi;
    }
  }
  if (threadIdx.x == 0) *A.count = base;
}

// ---------------------------------------------------------------------------------------------------
// SceneClipperProjective2D::compute with voxelize_resolution == 0 (R/mapping/scene_clipper_projective_2d.cpp:22-62):
// the z-buffer winners of the scene seen from robot_in_local_map * sensor_in_robot, in column order, as points in
// the sensor frame, then moved into the robot frame.  One CTA per request.

// ls2d_set_clouds_dev: the largest cloud of a borrowed CSR set (an int, zeroed before the launch); a negative size
// anywhere makes it negative for good
__global__ void largest_cloud_kernel(const int* off, int n, int* out) {
  int best = 0;
  bool bad = false;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int m = off[i + 1] - off[i];
    bad |= m < 0;
    best = max(best, m);
  }
  best = __reduce_max_sync(0xffffffffu, best);
  bad  = __any_sync(0xffffffffu, bad);
  if ((threadIdx.x & 31) == 0) {
    if (bad)
      atomicMin(out, -0x7fffffff);
    else if (atomicMax(out, best) < 0)
      atomicMin(out, -0x7fffffff);
  }
}

constexpr int CLIP_T = 256;  // at most 32 column chunks => canvas_cols <= 8192

__global__ void __launch_bounds__(CLIP_T) clip_kernel(const dev_params P, const clip_args A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int C      = P.cam.cols;
  unsigned* zdepth = reinterpret_cast<unsigned*>(smem_raw);
  unsigned* zidx   = zdepth + C;
  __shared__ int cnt[32 * (CLIP_T / 32)];
  __shared__ int misc[2];
  const int T = CLIP_T, tid = threadIdx.x;
  const int r      = blockIdx.x + A.base;
  const int cloud  = A.cloud_ids[r];
  const int p0 = A.off[cloud], n = A.off[cloud + 1] - p0;
  const iso S   = load_pose(A.sensor_pose, 0, A.pose_stride);
  const iso cam = iso_compose(load_pose(A.robot_pose, (size_t) r, A.pose_stride), S);
  const iso W   = iso_inverse(cam);
  const bool move = !(S.c == 1.f && S.s == 0.f && S.tx == 0.f && S.ty == 0.f);  // .cpp:60
  zbuffer_project<false>(P, W, A.pts + p0, n, zdepth, zidx);
  const int col_chunks = (C + T - 1) / T;
  unsigned mask = 0;
  for (int c = 0; c < col_chunks; ++c) {
    const int k = c * T + tid;
    if (k < C && zidx[k] != Z_EMPTY_IDX) mask |= 1u << c;
  }
  compact_count(mask, col_chunks, cnt, &misc[0]);
  float4* out = output_rows(A.pack, A.out, (size_t) C, r, misc[0], A.counts, &misc[1]);
  for (int c = 0; c < col_chunks; ++c) {
    const int k = c * T + tid, dst = compact_slot(mask, c, cnt);
    if (!((mask >> c) & 1u)) continue;
    const float4 p = ldg4(A.pts + p0 + zidx[k]);
    float4 o;
    iso_apply(W, p.x, p.y, o.x, o.y);
    iso_rot(W, p.z, p.w, o.z, o.w);
    if (move) {
      float x, y, nx, ny;
      iso_apply(S, o.x, o.y, x, y);
      iso_rot(S, o.z, o.w, nx, ny);
      o = make_float4(x, y, nx, ny);
    }
    out[dst] = o;
  }
}

// ---------------------------------------------------------------------------------------------------
// MergerProjective2D::compute (R/mapping/merger_projective_2d.cpp:9-100): both clouds projected from
// measurement_in_scene, per-column add / average+renormalise / replace / append; the scene is updated in place
// and grows by an ordered append.  One CTA per (scene, measurement) request.

__global__ void merge_kernel(const dev_params P, const merge_args A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int C   = P.cam.cols;
  unsigned* zds = reinterpret_cast<unsigned*>(smem_raw);
  unsigned* zis = zds + C;
  unsigned* zdm = zis + C;
  unsigned* zim = zdm + C;
  __shared__ int warp_tot[32];
  __shared__ int base;
  __shared__ int cnt[4];
  const int n_scene = *A.scene_size;
  const iso M = load_pose(A.mis_pose, 0, A.pose_stride);
  const iso W = iso_inverse(M);
  zbuffer_project<false, true>(P, W, A.scene, n_scene, zds, zis, nullptr);   // .cpp:19-20 (plain loads: scene is written below)
  zbuffer_project<false, true>(P, W, A.meas, A.n_meas, zdm, zim, &M);        // .cpp:22-25
  if (threadIdx.x < 4) cnt[threadIdx.x] = 0;
  if (threadIdx.x == 0) base = 0;
  __syncthreads();
  const float far_limit = fmul(.9f, P.range_max);
  for (int k0 = 0; k0 < C; k0 += blockDim.x) {
    const int k  = k0 + threadIdx.x;
    bool append  = false;
    float4 mp    = make_float4(0.f, 0.f, 0.f, 0.f);
    if (k < C && zim[k] != Z_EMPTY_IDX && !(u2f(zdm[k]) > far_limit)) {      // .cpp:46-53
      const float4 m = A.meas[zim[k]];
      iso_apply(M, m.x, m.y, mp.x, mp.y);
      iso_rot(M, m.z, m.w, mp.z, mp.w);
      if (zis[k] == Z_EMPTY_IDX) {                                           // .cpp:57-62
        append = true;
        atomicAdd(&cnt[0], 1);
      } else {
        float4* sp     = A.scene + zis[k];
        const float dr = fsub(u2f(zdm[k]), u2f(zds[k]));                     // .cpp:66
        if (fabsf(dr) < A.merge_threshold) {                                 // .cpp:71-76
          const float4 s = *sp;
          float x = fmul(fadd(s.x, mp.x), 0.5f), y = fmul(fadd(s.y, mp.y), 0.5f);
          float nx = fmul(fadd(s.z, mp.z), 0.5f), ny = fmul(fadd(s.w, mp.w), 0.5f);
          const float z = fadd(fmul(nx, nx), fmul(ny, ny));
          if (z > 0.f) {
            const float nrm = fsqrt(z);
            nx = fdiv(nx, nrm), ny = fdiv(ny, nrm);
          }
          *sp = make_float4(x, y, nx, ny);
          atomicAdd(&cnt[1], 1);
        } else if (dr > 0.f) {                                               // .cpp:80-84
          *sp = mp;
          atomicAdd(&cnt[2], 1);
        } else {                                                             // .cpp:87-88
          append = true;
        }
      }
    }
    const int dst = ordered_slot(append, warp_tot, &base);
    if (append) {
      if (n_scene + dst < A.capacity)
        A.scene[n_scene + dst] = mp;
      else
        cnt[3] = 1;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    *A.scene_size = n_scene + base < A.capacity ? n_scene + base : A.capacity;
    if (A.counters) A.counters[0] = cnt[0], A.counters[1] = cnt[1], A.counters[2] = cnt[2], A.counters[3] = cnt[3];
  }
}

// ---------------------------------------------------------------------------------------------------
// inlier flag of given correspondences at the estimate X (MultiAligner2D.keep_only_inlier_correspondences): the
// factor's squared error, operation for operation correspondence_chi() before the robustifier, against the Cauchy
// threshold (oracle decision D6: chi < tau is an inlier)
struct classify_dev {
  const int* off_f;
  const int* off_m;
  int cloud_f, cloud_m;
};
__global__ void classify_kernel(const dev_params P, const classify_args A, const classify_dev D) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= A.n) return;
  const int f0 = D.off_f[D.cloud_f], nf = D.off_f[D.cloud_f + 1] - f0;
  const int m0 = D.off_m[D.cloud_m], nm = D.off_m[D.cloud_m + 1] - m0;
  const int fi = A.fixed_idx[k], mi = A.moving_idx[k];
  if (fi < 0 || fi >= nf || mi < 0 || mi >= nm) {
    A.is_inlier[k] = 0;
    return;
  }
  const iso X = load_pose(A.X_pose, 0, A.pose_stride);
  pose_bc bc;
  publish_pose(&bc, P, X, P.with_sensor != 0, 0);
  dev_params Q = P;
  Q.tau        = -1.f;  // the plain squared error
  const float4 F = ldg4(A.fixed_pts + f0 + fi), M = ldg4(A.moving_pts + m0 + mi);
  const float chi = P.with_sensor ? correspondence_chi<true>(Q, &bc, F, M) : correspondence_chi<false>(Q, &bc, F, M);
  A.is_inlier[k]  = !(P.tau > 0.f) || chi < P.tau;
}

// ---------------------------------------------------------------------------------------------------
// self-test of fsqrt_gated (ls2d_math.cuh) against __fsqrt_rn over every binary32 value in [lo_bits, hi_bits]
__global__ void selftest_sqrt_kernel(unsigned lo_bits, unsigned hi_bits, unsigned long long* n_mismatch) {
  unsigned long long bad = 0;
  const unsigned long long n = (unsigned long long) hi_bits - lo_bits + 1ull;
  for (unsigned long long k = (unsigned long long) blockIdx.x * blockDim.x + threadIdx.x; k < n;
       k += (unsigned long long) gridDim.x * blockDim.x) {
    const float a = u2f(lo_bits + (unsigned) k);
    bad += f2u(fsqrt_gated(a)) != f2u(__fsqrt_rn(a));
  }
  bad = __reduce_add_sync(0xffffffffu, (unsigned) bad);
  if ((threadIdx.x & 31) == 0 && bad) atomicAdd(n_mismatch, bad);
}

// ---------------------------------------------------------------------------------------------------
// acceptance gates (L0.json:627-634) + deterministic best-of, one CTA.
__device__ __forceinline__ bool accepts(const ls2d_result& r, const ls2d_gates& g) {
  if (r.status != LS2D_STATUS_SUCCESS) return false;
  if (r.n_inliers < g.min_inliers || r.n_inliers <= 0 || r.n_corr <= 0) return false;
  if (fdiv(r.chi_inliers, (float) r.n_inliers) > g.max_chi_per_inlier) return false;
  if (fdiv((float) r.n_inliers, (float) r.n_corr) < g.min_inlier_ratio) return false;
  return true;
}
// strict "a better than b": more inliers, then lower chi per inlier, then lower id
__device__ __forceinline__ bool better(int na, float ca, int ia, int nb, float cb, int ib) {
  if (ib < 0) return ia >= 0;
  if (ia < 0) return false;
  if (na != nb) return na > nb;
  if (ca != cb) return ca < cb;
  return ia < ib;
}

__global__ void best_of_kernel(const ls2d_result* res, int n, int n_guess, ls2d_gates g, int candidate_base,
                               ls2d_best* out) {
  __shared__ int s_n[32], s_i[32];
  __shared__ float s_c[32];
  int bn = 0, bi = -1;
  float bcpi = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const ls2d_result r = res[i];
    if (!accepts(r, g)) continue;
    const float c = fdiv(r.chi_inliers, (float) r.n_inliers);
    if (better(r.n_inliers, c, i, bn, bcpi, bi)) bn = r.n_inliers, bcpi = c, bi = i;
  }
  for (int off = 16; off >= 1; off >>= 1) {
    const int on   = __shfl_xor_sync(0xffffffffu, bn, off);
    const float oc = __shfl_xor_sync(0xffffffffu, bcpi, off);
    const int oi   = __shfl_xor_sync(0xffffffffu, bi, off);
    if (better(on, oc, oi, bn, bcpi, bi)) bn = on, bcpi = oc, bi = oi;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) s_n[warp] = bn, s_c[warp] = bcpi, s_i[warp] = bi;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int) (blockDim.x >> 5); ++w)
      if (better(s_n[w], s_c[w], s_i[w], bn, bcpi, bi)) bn = s_n[w], bcpi = s_c[w], bi = s_i[w];
    ls2d_best b;
    if (bi < 0) {
      b.x = b.y = b.theta = b.chi_inliers = 0.f;
      b.n_inliers = b.n_corr = 0;
      b.candidate = -1, b.guess = -1;
      b.c = 1.f, b.s = 0.f, b.iterations = 0, b.reserved = 0;
    } else {
      const ls2d_result r = res[bi];
      b.x = r.x, b.y = r.y, b.theta = r.theta, b.chi_inliers = r.chi_inliers;
      b.n_inliers = r.n_inliers, b.n_corr = r.n_corr;
      b.c = r.c, b.s = r.s, b.iterations = r.iterations, b.reserved = 0;
      b.candidate = candidate_base + bi / n_guess;
      b.guess     = bi % n_guess;
    }
    *out = b;
  }
}

// acceptance gates + best-of per GROUP of consecutive results (all-pairs search, BASELINE.json configs[4]: one group
// per query local map); one warp per group.  The record names the winning pair: candidate = moving_id[pair] (the
// pair index when moving_id is null), guess = index of the pair inside its group.  Ordering as above, ties by the
// lower pair index, so any sharding of the groups over ranks gives the same records.
__global__ void best_of_groups_kernel(const ls2d_result* res, const int* group_off, int n_groups, const int* moving_id,
                                      ls2d_gates g, ls2d_best* out) {
  const int lane = threadIdx.x & 31;
  const int grp  = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (grp >= n_groups) return;
  const int p0 = group_off[grp], p1 = group_off[grp + 1];
  int bn = 0, bi = -1;
  float bcpi = 0.f;
  for (int i = p0 + lane; i < p1; i += 32) {
    const ls2d_result r = res[i];
    if (!accepts(r, g)) continue;
    const float c = fdiv(r.chi_inliers, (float) r.n_inliers);
    if (better(r.n_inliers, c, i, bn, bcpi, bi)) bn = r.n_inliers, bcpi = c, bi = i;
  }
  for (int off = 16; off >= 1; off >>= 1) {
    const int on   = __shfl_xor_sync(0xffffffffu, bn, off);
    const float oc = __shfl_xor_sync(0xffffffffu, bcpi, off);
    const int oi   = __shfl_xor_sync(0xffffffffu, bi, off);
    if (better(on, oc, oi, bn, bcpi, bi)) bn = on, bcpi = oc, bi = oi;
  }
  if (lane == 0) {
    ls2d_best b;
    if (bi < 0) {
      b.x = b.y = b.theta = b.chi_inliers = 0.f;
      b.n_inliers = b.n_corr = 0;
      b.candidate = -1, b.guess = -1;
      b.c = 1.f, b.s = 0.f, b.iterations = 0, b.reserved = 0;
    } else {
      const ls2d_result r = res[bi];
      b.x = r.x, b.y = r.y, b.theta = r.theta, b.chi_inliers = r.chi_inliers;
      b.n_inliers = r.n_inliers, b.n_corr = r.n_corr;
      b.c = r.c, b.s = r.s, b.iterations = r.iterations, b.reserved = 0;
      b.candidate = moving_id ? moving_id[bi] : bi;
      b.guess     = bi - p0;
    }
    out[grp] = b;
  }
}

}  // namespace ls2d
