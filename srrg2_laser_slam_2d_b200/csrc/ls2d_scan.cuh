// ls2d_scan.cuh -- RawDataPreprocessorProjective2D on the device: raw LaserMessage ranges in,
// PointNormal2f clouds out, so that 4 B/beam instead of 16 B/point cross PCIe.
//
//  preprocess_kernel   one CTA per scan: polar unprojection with an ordered compaction of the accepted beams,
//                      sliding-window normals (every point sums its own window sequentially: the reference's
//                      order), optional voxelisation = runs of consecutive points in one voxel ("segments"),
//                      bitonic sort of the few hundred segments in shared memory, one sequential sum per voxel in
//                      cloud order, ordered output.
//  scan_offsets_kernel exclusive scan of the per-scan counts -> CSR offsets
//  scan_pack_kernel    strided [n_scans][n_beams] -> packed CSR points
//
// Reference: R/sensor_processing/raw_data_preprocessor_projective_2d.cpp:13-51,77-104 (R/ =
// /root/reference/srrg2_laser_slam_2d/src/srrg2_laser_slam_2d/); the upstream pieces (unprojector, normal
// computator, voxelize) follow the decision points P1..P8 of oracle/ls2d_oracle.c operation by operation.
#pragma once

#include <cuda_runtime.h>
#include <float.h>

#include "ls2d_service.cuh"

namespace ls2d {



constexpr size_t scan_smem_bytes(int n_beams, int sort_cap) {
  return (size_t) n_beams * (8 + 8) + (size_t) ((n_beams + 3) & ~3) + (size_t) sort_cap * 8 + 64;
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

// voxelize(res_coeffs) parameters: inverse scales of (x, y) and of the normal, nb = bound of |trunc(n * inv_n)|
struct voxel_params {
  float inv_res, inv_n;
  int nb;
};

// voxel key of a point (P7): trunc-toward-zero of (x, y, nx, ny) * (1/res, 1/res, inv_n, inv_n), packed so that unsigned
// order = the oracle's lexicographic order: ix (20 bits, biased) | iy (20 bits, biased) | normal code (10 bits); the
// low 14 bits are left for the first point of a segment.  |coordinate / res| < 2^19 is checked on the host; normal
// cells are clamped to +-nb (unit normals never leave that range).
__device__ __forceinline__ unsigned long long voxel_key(const voxel_params& V, float2 p, float2 nv) {
  const int ix = __float2int_rz(fmul(p.x, V.inv_res)), iy = __float2int_rz(fmul(p.y, V.inv_res));
  const int inx = max(-V.nb, min(V.nb, __float2int_rz(fmul(nv.x, V.inv_n))));
  const int iny = max(-V.nb, min(V.nb, __float2int_rz(fmul(nv.y, V.inv_n))));
  const unsigned long long bx = (unsigned) (ix + (1 << 19)) & 0xFFFFFu, by = (unsigned) (iy + (1 << 19)) & 0xFFFFFu;
  return (bx << 44) | (by << 24) | ((unsigned long long) ((inx + V.nb) * (2 * V.nb + 1) + (iny + V.nb)) << 14);
}
constexpr unsigned long long VOXEL_MASK = ~0x3FFFull;  // everything but the point index

// Block-wide voxelize of the n points (xy, nrm) with flags valid (nullptr: all valid), in place in shared memory:
// consecutive valid points that share a voxel form a SEGMENT (neighbouring beams hit neighbouring places); only the
// segments are sorted, by (voxel key, first point), and a voxel's sum walks its segments in that order = cloud
// order.  emit(slot, point) receives the output points in sorted order; returns their number (to every thread).
// key: n slots of shared memory, cnt / total: the compaction scratch.  Needs n <= 32 * blockDim.x and n < 2^14.
template <typename Emit>
__device__ __forceinline__ int block_voxelize(const voxel_params& V, const float2* xy, const float2* nrm,
                                              const unsigned char* valid, int n, unsigned long long* key, int* cnt,
                                              int* total, Emit emit) {
  const int T = blockDim.x, tid = threadIdx.x;
  const int pt_chunks = (n + T - 1) / T;
  unsigned mask = 0;
  for (int c = 0; c < pt_chunks; ++c) {  // segment heads: valid points whose previous valid point has another key
    const int i = c * T + tid;
    if (i < n && (!valid || valid[i])) {
      int prev = i - 1;
      while (valid && prev >= 0 && !valid[prev]) --prev;
      const bool head = prev < 0 || voxel_key(V, xy[i], nrm[i]) != voxel_key(V, xy[prev], nrm[prev]);
      if (head) mask |= 1u << c;
    }
  }
  compact_count(mask, pt_chunks, cnt, total);
  const int S = *total;
  int cap     = 1;
  while (cap < S) cap <<= 1;
  for (int c = 0; c < pt_chunks; ++c) {
    const int i = c * T + tid, dst = compact_slot(mask, c, cnt);
    if ((mask >> c) & 1u) key[dst] = voxel_key(V, xy[i], nrm[i]) | (unsigned long long) i;
  }
  __syncthreads();
  // bitonic sort of the S segments, ascending, in the all-ascending formulation (every merge starts with a "flip"
  // stage, then half-cleaners): elements beyond S are virtual +infinity that never move, so only S slots exist.
  // A thread always owns the same pair slots, and for spans of at most 64 both elements of a pair live in its
  // warp's 64-element block: those stages need no block barrier.
  for (int k = 2; k <= cap; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      const int lx = j == (k >> 1) ? k - 1 : j;  // flip: partner = i ^ (k - 1); half-cleaner: partner = i | j
      for (int t = tid; t < (cap >> 1); t += T) {
        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1)), l = i ^ lx;
        if (l >= S) continue;
        const unsigned long long a = key[i], b2 = key[l];
        if (b2 < a) key[i] = b2, key[l] = a;
      }
      // next stage: the flip of merge 2k (span 2k) after j == 1, else a half-cleaner of span j
      const int span_now = lx == j ? 2 * j : k, span_next = j == 1 ? 2 * k : j;
      if (span_now > 64 || span_next > 64)
        __syncthreads();
      else
        __syncwarp();
    }
  __syncthreads();
  // one output point per run of equal voxel keys: the run's first segment sums all members in sorted (= cloud) order;
  // a segment's members are the valid points from its first point on that still carry its voxel key
  const int seg_chunks = (S + T - 1) / T;
  mask = 0;
  for (int c = 0; c < seg_chunks; ++c) {
    const int t = c * T + tid;
    if (t < S && (t == 0 || ((key[t] ^ key[t - 1]) & VOXEL_MASK) != 0)) mask |= 1u << c;
  }
  compact_count(mask, seg_chunks, cnt, total);
  for (int c = 0; c < seg_chunks; ++c) {
    const int t = c * T + tid, dst = compact_slot(mask, c, cnt);
    if (!((mask >> c) & 1u)) continue;
    const unsigned long long vk = key[t] & VOXEL_MASK;
    float4 o  = make_float4(0.f, 0.f, 0.f, 0.f);
    int count = 0;
    for (int e = t; e < S && (key[e] & VOXEL_MASK) == vk; ++e) {
      const int first = (int) (key[e] & 0x3FFF);
      for (int i = first; i < n; ++i) {
        if (valid && !valid[i]) continue;
        if (i != first && (voxel_key(V, xy[i], nrm[i]) & VOXEL_MASK) != vk) break;
        o.x = fadd(o.x, xy[i].x), o.y = fadd(o.y, xy[i].y), o.z = fadd(o.z, nrm[i].x), o.w = fadd(o.w, nrm[i].y);
        ++count;
      }
    }
    const float w = fdiv(1.f, (float) count);
    o.x = fmul(o.x, w), o.y = fmul(o.y, w), o.z = fmul(o.z, w), o.w = fmul(o.w, w);
    const float z = fadd(fmul(o.z, o.z), fmul(o.w, o.w));
    if (z > 0.f) {
      const float nn = fsqrt(z);
      o.z = fdiv(o.z, nn), o.w = fdiv(o.w, nn);
    }
    emit(dst, o);
  }
  return *total;
}

constexpr int SCAN_T = 384;  // threads per scan; at most 32 chunks => n_beams <= 12288 (and < 2^14: key layout)

__global__ void __launch_bounds__(SCAN_T) preprocess_kernel(const scan_dev_params P, const scan_args A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int NB = P.n_beams;
  float2* xy                = reinterpret_cast<float2*>(smem_raw);                 // accepted beams, beam order
  float2* nrm               = xy + NB;                                             // their normals
  unsigned long long* key   = reinterpret_cast<unsigned long long*>(nrm + NB);     // [sort_cap] segment: voxel key | first point
  unsigned char* valid      = reinterpret_cast<unsigned char*>(key + P.sort_cap);  // [NB]
  __shared__ int cnt[32 * (SCAN_T / 32)];
  __shared__ int total;
  const int T = SCAN_T, tid = threadIdx.x;
  const float* ranges = A.ranges + (size_t) blockIdx.x * NB;
  float4* out         = A.out + (size_t) blockIdx.x * NB;
  const int beam_chunks = (NB + T - 1) / T;

  // ---- PointNormal2fUnprojectorPolar, accepted beams only, beam order (P1); staged through `nrm`
  unsigned mask = 0;
  for (int c = 0; c < beam_chunks; ++c) {
    const int b = c * T + tid;
    if (b < NB) {
      const float r = __ldg(ranges + b);
      if (!(r < P.range_min || r > P.range_max)) {
        const float az = fmul(P.ifx, fsub((float) b, P.cx));
        nrm[b]         = make_float2(fmul(r, cosf_glibc(az)), fmul(r, sinf_glibc(az)));
        mask |= 1u << c;
      }
    }
  }
  compact_count(mask, beam_chunks, cnt, &total);
  for (int c = 0; c < beam_chunks; ++c) {
    const int dst = compact_slot(mask, c, cnt);
    if ((mask >> c) & 1u) xy[dst] = nrm[c * T + tid];
  }
  const int n = total;
  __syncthreads();

  // ---- NormalComputator1DSlidingWindow (P2..P6): every point walks its own window
  for (int i = tid; i < n; i += T) {
    const float2 p = xy[i];
    const f2 np    = mk2(-p.x, -p.y);
    // one walk over the window (down, then up): moments of d = q - p in walking order (P2, P4); the independent
    // operations go out in pairs (FADD2 / FMUL2: same roundings)
    f2 s1 = mk2(0.f, 0.f);  // sum dx, sum dy
    float sxx = 0.f, sxy = 0.f, syy = 0.f;
    int wn = 1;
    for (int j = i - 1; j >= 0; --j) {
      const float2 q = xy[j];
      const f2 d  = add2(mk2(q.x, q.y), np);
      const f2 dd = mul2(d, d);
      if (!(fadd(dd.x, dd.y) < P.d2)) break;
      s1  = add2(s1, d);
      sxx = fadd(sxx, dd.x), sxy = fadd(sxy, fmul(d.x, d.y)), syy = fadd(syy, dd.y);
      ++wn;
    }
    for (int j = i + 1; j < n; ++j) {
      const float2 q = xy[j];
      const f2 d  = add2(mk2(q.x, q.y), np);
      const f2 dd = mul2(d, d);
      if (!(fadd(dd.x, dd.y) < P.d2)) break;
      s1  = add2(s1, d);
      sxx = fadd(sxx, dd.x), sxy = fadd(sxy, fmul(d.x, d.y)), syy = fadd(syy, dd.y);
      ++wn;
    }
    const bool ok = wn >= P.min_points;
    float nx = 0.f, ny = 0.f;
    if (ok) {
      const float fc = (float) wn;
      const float mx = fdiv(s1.x, fc), my = fdiv(s1.y, fc);
      const float cxx = fsub(fdiv(sxx, fc), fmul(mx, mx)), cxy = fsub(fdiv(sxy, fc), fmul(mx, my)),
                  cyy = fsub(fdiv(syy, fc), fmul(my, my));
      smallest_eigenvector_2x2(cxx, cxy, cyy, nx, ny);
      if (fadd(fmul(nx, p.x), fmul(ny, p.y)) > 0.f) nx = -nx, ny = -ny;
    }
    nrm[i]   = make_float2(nx, ny);
    valid[i] = ok;
  }
  __syncthreads();
  const int pt_chunks = (n + T - 1) / T;

  if (P.inv_res == 0.f) {  // ---- valid points in cloud order (.cpp:44-48, P8)
    mask = 0;
    for (int c = 0; c < pt_chunks; ++c) {
      const int i = c * T + tid;
      if (i < n && valid[i]) mask |= 1u << c;
    }
    compact_count(mask, pt_chunks, cnt, &total);
    for (int c = 0; c < pt_chunks; ++c) {
      const int i = c * T + tid, dst = compact_slot(mask, c, cnt);
      if ((mask >> c) & 1u) out[dst] = make_float4(xy[i].x, xy[i].y, nrm[i].x, nrm[i].y);
    }
    if (tid == 0) A.counts[blockIdx.x] = total;
    return;
  }

  // ---- voxelize(res, res, 1, 1) (.cpp:38-42, P7)
  voxel_params V;
  V.inv_res = P.inv_res, V.inv_n = 1.f, V.nb = 1;
  const int n_out = block_voxelize(V, xy, nrm, valid, n, key, cnt, &total, [&](int slot, float4 o) { out[slot] = o; });
  if (tid == 0) A.counts[blockIdx.x] = n_out;
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

// SceneClipperProjective2D::compute with voxelize_resolution > 0 (R/mapping/scene_clipper_projective_2d.cpp:36-48):
// the z-buffer winners in column order, as points in the sensor frame, voxelized with res_coeffs (res, res, 0.1,
// 0.1), then moved into the robot frame.  One CTA per request; shared memory: clip_voxel_smem_bytes(canvas_cols).
constexpr size_t clip_voxel_smem_bytes(int cols) { return (size_t) cols * (4 + 4 + 8 + 8 + 8) + 64; }

__global__ void __launch_bounds__(SCAN_T) clip_voxel_kernel(const dev_params P, const clip_args A, float inv_res) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int C             = P.cam.cols;
  float2* xy              = reinterpret_cast<float2*>(smem_raw);
  float2* nrm             = xy + C;
  unsigned long long* key = reinterpret_cast<unsigned long long*>(nrm + C);
  unsigned* zdepth        = reinterpret_cast<unsigned*>(key + C);
  unsigned* zidx          = zdepth + C;
  __shared__ int cnt[32 * (SCAN_T / 32)];
  __shared__ int total;
  const int T = SCAN_T, tid = threadIdx.x;
  const int r     = blockIdx.x;
  const int cloud = A.cloud_ids[r];
  const int p0 = A.off[cloud], n = A.off[cloud + 1] - p0;
  const iso S   = load_pose(A.sensor_pose, 0, A.pose_stride);
  const iso cam = iso_compose(load_pose(A.robot_pose, (size_t) r, A.pose_stride), S);
  const iso W   = iso_inverse(cam);
  const bool move = !(S.c == 1.f && S.s == 0.f && S.tx == 0.f && S.ty == 0.f);  // .cpp:60
  zbuffer_project<false>(P, W, A.pts + p0, n, zdepth, zidx);
  // winners in column order (.cpp:38-43), in the sensor frame
  const int col_chunks = (C + T - 1) / T;
  unsigned mask = 0;
  for (int c = 0; c < col_chunks; ++c) {
    const int k = c * T + tid;
    if (k < C && zidx[k] != Z_EMPTY_IDX) mask |= 1u << c;
  }
  compact_count(mask, col_chunks, cnt, &total);
  for (int c = 0; c < col_chunks; ++c) {
    const int k = c * T + tid, dst = compact_slot(mask, c, cnt);
    if ((mask >> c) & 1u) {
      const float4 p = ldg4(A.pts + p0 + zidx[k]);
      float2 q, nq;
      iso_apply(W, p.x, p.y, q.x, q.y);
      iso_rot(W, p.z, p.w, nq.x, nq.y);
      xy[dst] = q, nrm[dst] = nq;
    }
  }
  const int m = total;
  __syncthreads();
  voxel_params V;
  V.inv_res = inv_res, V.inv_n = fdiv(1.f, 0.1f), V.nb = 10;
  float4* out     = A.out + (size_t) r * C;
  const int n_out = block_voxelize(V, xy, nrm, nullptr, m, key, cnt, &total, [&](int slot, float4 o) {
    if (move) {  // .cpp:60-62
      float x, y, nx, ny;
      iso_apply(S, o.x, o.y, x, y);
      iso_rot(S, o.z, o.w, nx, ny);
      o = make_float4(x, y, nx, ny);
    }
    out[slot] = o;
  });
  if (tid == 0) A.counts[r] = n_out;
}

}  // namespace ls2d
