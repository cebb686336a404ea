// ls2d_scan.cuh -- RawDataPreprocessorProjective2D on the device: raw LaserMessage ranges in,
// PointNormal2f clouds out, so that 4 B/beam instead of 16 B/point cross PCIe.
//
//  preprocess_kernel   one CTA per scan: polar unprojection with an ordered compaction of the accepted beams,
//                      sliding-window normals (every point sums its own window sequentially: the reference's
//                      order), optional voxelisation = bitonic sort of the (voxel key, index) pairs in shared
//                      memory + one sequential sum per run of equal keys, ordered output.
//  scan_offsets_kernel exclusive scan of the per-scan counts -> CSR offsets
//  scan_pack_kernel    strided [n_scans][n_beams] -> packed CSR points
//
// Reference: R/sensor_processing/raw_data_preprocessor_projective_2d.cpp:13-51,77-104 (R/ =
// /root/reference/srrg2_laser_slam_2d/src/srrg2_laser_slam_2d/); the upstream pieces (unprojector, normal
// computator, voxelize) follow the decision points P1..P8 of oracle/ls2d_oracle.c operation by operation.
#pragma once

#include <cuda_runtime.h>
#include <float.h>

#include "ls2d_kernels.cuh"

namespace ls2d {

struct scan_dev_params {
  float range_min, range_max;  // the tighter of message and PARAM limits (.cpp:83-84)
  float ifx, cx;               // azimuth = ifx * (c - cx), sensor matrix [1/res, n/2] (.cpp:87-90)
  float d2;                    // normal_point_distance^2
  float inv_res;               // 1 / voxelize_resolution, 0: valid-only copy (.cpp:44-48)
  int min_points;              // normal_min_points
  int n_beams;
  int sort_cap;                // power of two >= n_beams (voxelisation on), else 0
};

struct scan_args {
  const float* ranges;  // [n_scans][n_beams]
  float4* out;          // [n_scans][n_beams]
  int* counts;          // [n_scans]
  int n_scans;
};

constexpr size_t scan_smem_bytes(int n_beams, int sort_cap) {
  return (size_t) n_beams * (8 + 8 + 4) + (size_t) sort_cap * (8 + 4) + 64;
}

// Eigen 3.3 SelfAdjointEigenSolver<Matrix2f>::computeDirect: eigenvector of the smallest eigenvalue (P5)
__device__ __forceinline__ void smallest_eigenvector_2x2(float m00, float m10, float m11, float& vx, float& vy) {
  const float shift = fdiv(fadd(m00, m11), 2.f);
  float a = fsub(m00, shift), b = m10, c = fsub(m11, shift);
  float scale = fabsf(a);
  if (fabsf(b) > scale) scale = fabsf(b);
  if (fabsf(c) > scale) scale = fabsf(c);
  if (scale > 0.f) a = fdiv(a, scale), b = fdiv(b, scale), c = fdiv(c, scale);
  const float d  = fsub(a, c);
  const float t0 = fmul(0.5f, fsqrt(fadd(fmul(d, d), fmul(4.f, fmul(b, b)))));
  const float t1 = fmul(0.5f, fadd(a, c));
  const float r0 = fsub(t1, t0), r1 = fadd(t1, t0);
  if (fsub(r1, r0) <= fmul(fabsf(r1), FLT_EPSILON)) {
    vx = 1.f, vy = 0.f;
    return;
  }
  const float a1 = fsub(a, r1), c1 = fsub(c, r1);
  const float a2 = fmul(a1, a1), c2 = fmul(c1, c1), b2 = fmul(b, b);
  float ux, uy;
  if (a2 > c2) {
    const float n = fsqrt(fadd(a2, b2));
    ux = fdiv(-b, n), uy = fdiv(a1, n);
  } else {
    const float n = fsqrt(fadd(c2, b2));
    ux = fdiv(-c1, n), uy = fdiv(b, n);
  }
  const float ox = -uy, oy = ux;
  const float z  = fadd(fmul(ox, ox), fmul(oy, oy));
  if (z > 0.f) {
    const float n = fsqrt(z);
    vx = fdiv(ox, n), vy = fdiv(oy, n);
  } else {
    vx = ox, vy = oy;
  }
}

__device__ __forceinline__ bool key_less(unsigned long long a1, unsigned a2, unsigned long long b1, unsigned b2) {
  return a1 < b1 || (a1 == b1 && a2 < b2);
}

__global__ void __launch_bounds__(256) preprocess_kernel(const scan_dev_params P, const scan_args A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int NB = P.n_beams;
  float2* xy                = reinterpret_cast<float2*>(smem_raw);                 // accepted beams, beam order
  float2* nrm               = xy + NB;                                             // their normals
  unsigned long long* key1  = reinterpret_cast<unsigned long long*>(nrm + NB);     // [sort_cap] (ix, iy), biased
  unsigned* key2            = reinterpret_cast<unsigned*>(key1 + P.sort_cap);      // [sort_cap] (inx, iny, index)
  unsigned* valid           = key2 + P.sort_cap;                                   // [NB]
  __shared__ int warp_tot[32];
  __shared__ int base;
  const int T = blockDim.x, tid = threadIdx.x;
  const float* ranges = A.ranges + (size_t) blockIdx.x * NB;
  float4* out         = A.out + (size_t) blockIdx.x * NB;

  // ---- PointNormal2fUnprojectorPolar, accepted beams only, beam order (P1)
  if (tid == 0) base = 0;
  __syncthreads();
  for (int c0 = 0; c0 < NB; c0 += T) {
    const int c = c0 + tid;
    bool ok     = false;
    float x = 0.f, y = 0.f;
    if (c < NB) {
      const float r = __ldg(ranges + c);
      ok            = !(r < P.range_min || r > P.range_max);
      if (ok) {
        const float az = fmul(P.ifx, fsub((float) c, P.cx));
        x              = fmul(r, cosf_glibc(az));
        y              = fmul(r, sinf_glibc(az));
      }
    }
    const int dst = ordered_slot(ok, warp_tot, &base);
    if (ok) xy[dst] = make_float2(x, y);
  }
  const int n = base;
  __syncthreads();

  // ---- NormalComputator1DSlidingWindow (P2..P6): every point scans its own window
  for (int i = tid; i < n; i += T) {
    const float2 p = xy[i];
    int lo = i, hi = i;
    while (lo > 0) {
      const float2 q = xy[lo - 1];
      const float dx = fsub(q.x, p.x), dy = fsub(q.y, p.y);
      if (!(fadd(fmul(dx, dx), fmul(dy, dy)) < P.d2)) break;
      --lo;
    }
    while (hi + 1 < n) {
      const float2 q = xy[hi + 1];
      const float dx = fsub(q.x, p.x), dy = fsub(q.y, p.y);
      if (!(fadd(fmul(dx, dx), fmul(dy, dy)) < P.d2)) break;
      ++hi;
    }
    const int cnt = hi - lo + 1;
    const bool ok = cnt >= P.min_points;
    float nx = 0.f, ny = 0.f;
    if (ok) {
      float sx = 0.f, sy = 0.f;
      for (int j = lo; j <= hi; ++j) {
        const float2 q = xy[j];
        sx = fadd(sx, q.x), sy = fadd(sy, q.y);
      }
      const float fc = (float) cnt;
      const float mx = fdiv(sx, fc), my = fdiv(sy, fc);
      float cxx = 0.f, cxy = 0.f, cyy = 0.f;
      for (int j = lo; j <= hi; ++j) {
        const float2 q = xy[j];
        const float dx = fsub(q.x, mx), dy = fsub(q.y, my);
        cxx = fadd(cxx, fmul(dx, dx)), cxy = fadd(cxy, fmul(dx, dy)), cyy = fadd(cyy, fmul(dy, dy));
      }
      cxx = fdiv(cxx, fc), cxy = fdiv(cxy, fc), cyy = fdiv(cyy, fc);
      smallest_eigenvector_2x2(cxx, cxy, cyy, nx, ny);
      if (fadd(fmul(nx, p.x), fmul(ny, p.y)) > 0.f) nx = -nx, ny = -ny;
    }
    nrm[i]   = make_float2(nx, ny);
    valid[i] = ok;
  }
  if (tid == 0) base = 0;
  __syncthreads();

  if (P.inv_res == 0.f) {  // ---- valid points in cloud order (.cpp:44-48, P8)
    for (int i0 = 0; i0 < n; i0 += T) {
      const int i   = i0 + tid;
      const bool ok = i < n && valid[i];
      const int dst = ordered_slot(ok, warp_tot, &base);
      if (ok) out[dst] = make_float4(xy[i].x, xy[i].y, nrm[i].x, nrm[i].y);
    }
    if (tid == 0) A.counts[blockIdx.x] = base;
    return;
  }

  // ---- voxelize (.cpp:38-42, P7): keys of the valid points, compacted
  for (int i0 = 0; i0 < n; i0 += T) {
    const int i   = i0 + tid;
    const bool ok = i < n && valid[i];
    const int dst = ordered_slot(ok, warp_tot, &base);
    if (ok) {
      const int ix = __float2int_rz(fmul(xy[i].x, P.inv_res)), iy = __float2int_rz(fmul(xy[i].y, P.inv_res));
      const int inx = __float2int_rz(nrm[i].x), iny = __float2int_rz(nrm[i].y);  // in {-1, 0, 1}
      key1[dst] = ((unsigned long long) ((unsigned) ix ^ 0x80000000u) << 32) | ((unsigned) iy ^ 0x80000000u);
      key2[dst] = ((unsigned) ((inx + 1) * 3 + (iny + 1)) << 16) | (unsigned) i;
    }
  }
  const int m = base;
  int cap     = 1;
  while (cap < m) cap <<= 1;
  for (int i = m + tid; i < cap; i += T) key1[i] = ~0ull, key2[i] = ~0u;
  __syncthreads();
  // bitonic sort, ascending by (key1, key2); key2 carries the cloud index, so equal voxels keep cloud order
  for (int k = 2; k <= cap; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = tid; t < (cap >> 1); t += T) {
        const int i   = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        const int l   = i | j;
        const bool up = (i & k) == 0;
        const unsigned long long a1 = key1[i], b1 = key1[l];
        const unsigned a2 = key2[i], b2 = key2[l];
        if (key_less(b1, b2, a1, a2) == up) {
          key1[i] = b1, key2[i] = b2;
          key1[l] = a1, key2[l] = a2;
        }
      }
      __syncthreads();
    }
  // one output point per run of equal keys: the run's head sums its members in sorted (= cloud) order
  if (tid == 0) base = 0;
  __syncthreads();
  for (int t0 = 0; t0 < m; t0 += T) {
    const int t     = t0 + tid;
    const bool head = t < m && (t == 0 || key1[t] != key1[t - 1] || (key2[t] >> 16) != (key2[t - 1] >> 16));
    float4 o        = make_float4(0.f, 0.f, 0.f, 0.f);
    if (head) {
      const unsigned long long k1 = key1[t];
      const unsigned k2           = key2[t] >> 16;
      int e = t;
      for (; e < m && key1[e] == k1 && (key2[e] >> 16) == k2; ++e) {
        const int i = key2[e] & 0xFFFF;
        o.x = fadd(o.x, xy[i].x), o.y = fadd(o.y, xy[i].y), o.z = fadd(o.z, nrm[i].x), o.w = fadd(o.w, nrm[i].y);
      }
      const float w = fdiv(1.f, (float) (e - t));
      o.x = fmul(o.x, w), o.y = fmul(o.y, w), o.z = fmul(o.z, w), o.w = fmul(o.w, w);
      const float z = fadd(fmul(o.z, o.z), fmul(o.w, o.w));
      if (z > 0.f) {
        const float nn = fsqrt(z);
        o.z = fdiv(o.z, nn), o.w = fdiv(o.w, nn);
      }
    }
    const int dst = ordered_slot(head, warp_tot, &base);
    if (head) out[dst] = o;
  }
  if (tid == 0) A.counts[blockIdx.x] = base;
}

// exclusive scan of counts[n] -> off[n + 1]; one CTA, chunks of blockDim.x
__global__ void scan_offsets_kernel(const int* counts, int n, int* off) {
  __shared__ int warp_tot[32];
  __shared__ int carry;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
  if (tid == 0) carry = 0;
  __syncthreads();
  for (int i0 = 0; i0 < n; i0 += blockDim.x) {
    const int i = i0 + tid;
    const int v = i < n ? counts[i] : 0;
    int incl    = v;
    for (int o = 1; o < 32; o <<= 1) {
      const int u = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += u;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    int before = carry;
    for (int w = 0; w < warp; ++w) before += warp_tot[w];
    if (i < n) off[i] = before + incl - v;
    __syncthreads();
    if (tid == 0) {
      int t = carry;
      for (int w = 0; w < nwarp; ++w) t += warp_tot[w];
      carry = t;
    }
    __syncthreads();
  }
  if (tid == 0) off[n] = carry;
}

// strided [n_scans][n_beams] -> packed CSR; one CTA per scan
__global__ void scan_pack_kernel(const float4* strided, const int* off, int n_beams, float4* packed) {
  const int s = blockIdx.x;
  const int o = off[s], n = off[s + 1] - o;
  for (int i = threadIdx.x; i < n; i += blockDim.x) packed[o + i] = strided[(size_t) s * n_beams + i];
}

}  // namespace ls2d
