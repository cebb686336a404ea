// ls2d_scan.cuh -- RawDataPreprocessorProjective2D on the device: raw LaserMessage ranges in,
// PointNormal2f clouds out, so that 4 B/beam instead of 16 B/point cross PCIe.
//
//  beam_table_kernel   (cos, sin) of every beam's azimuth: the same for all scans of a sensor, computed once per
//                      (n_beams, sensor matrix) and kept by the handle
//  preprocess_kernel   one CTA per scan: polar unprojection with an ordered compaction of the accepted beams,
//                      sliding-window normals (every point sums its own window sequentially: the reference's
//                      order), optional voxelisation = runs of consecutive points in one voxel ("segments"),
//                      segments ordered by a counting sort over 2048 order-preserving buckets of (ix, iy) plus a
//                      rank inside the bucket (bitonic sort in shared memory when the buckets are too uneven), one
//                      sequential sum per voxel in cloud order, ordered output into the scan's strided row.  The
//                      last CTA to finish turns the counts into CSR offsets (no offsets pass); a look-back inside
//                      this kernel was measured and dropped: scans differ 3x in run time, and CTAs that wait for a
//                      slow predecessor hold their SM slots (437 vs 294 us per 4096 scans)
//  scan_pack_kernel    strided rows -> packed CSR points
//  clip_voxel_kernel   the clipper's voxelize branch (same block_voxelize), rows placed by look-back
//
// Reference: R/sensor_processing/raw_data_preprocessor_projective_2d.cpp:13-51,77-104 (R/ =
// /root/reference/srrg2_laser_slam_2d/src/srrg2_laser_slam_2d/); the upstream pieces (unprojector, normal
// computator, voxelize) follow the decision points P1..P8 of oracle/ls2d_oracle.c operation by operation.
#pragma once

#include <cuda_runtime.h>
#include <float.h>

#include "ls2d_service.cuh"

namespace ls2d {

// Eigen 3.3 SelfAdjointEigenSolver<Matrix2f>::computeDirect: eigenvector of the smallest eigenvalue (P5)
__device__ __forceinline__ void smallest_eigenvector_2x2(float m00, float m10, float m11, float& vx, float& vy) {
  const float shift = fmul(fadd(m00, m11), 0.5f);  // = the division by 2: both are exact up to the same rounding
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
constexpr unsigned long long VOXEL_MASK = ~0x3FFFull;  // everything but the segment index
constexpr unsigned long long NO_KEY     = ~0ull;      // per-point key of an invalid point

constexpr int SCAN_T      = 512;  // threads per scan; at most 32 chunks => n_beams <= 12288 (and < 2^14: key layout)
constexpr int VOX_BUCKETS = 2048;
// the rank inside the buckets costs sum(m_b^2) comparisons, the bitonic sort about S * log2(S)^2 / 2 compare-exchanges
// of twice the price: above this many comparisons per segment the buckets are too uneven (a long wall at x = const)
constexpr int VOX_RANK_LIMIT = 64;

// ---- voxelize ------------------------------------------------------------------------------------------------------
// shared-memory scratch of block_voxelize for at most n points
struct voxel_scratch {
  unsigned long long* keys;  // [n] per-point voxel keys, later the segments grouped by bucket
  unsigned long long* seg;   // [n] segments in cloud order (voxel key | segment index), later in sorted order
  unsigned short* first;     // [n + 1] first point of every segment, first[S] = n
  unsigned short* slot;      // [n] arrival order of a segment inside its bucket
  int* hist;                 // [VOX_BUCKETS + 1]
  int* cnt;                  // [32 * warps] compaction scratch
  int* misc;                 // [8]: total, min ix, max ix, min iy, max iy
};
constexpr size_t voxel_scratch_bytes(int n) {
  return (size_t) n * 16 + (((size_t) (2 * n + 1) * 2 + 15) & ~(size_t) 15) + (size_t) (VOX_BUCKETS + 1) * 4 + 16;
}
__device__ __forceinline__ unsigned char* carve_voxel_scratch(voxel_scratch& s, unsigned char* p, int n, int* cnt,
                                                              int* misc) {
  s.keys  = reinterpret_cast<unsigned long long*>(p);
  s.seg   = s.keys + n;
  s.first = reinterpret_cast<unsigned short*>(s.seg + n);
  s.slot  = s.first + n + 1;
  p       = reinterpret_cast<unsigned char*>(s.seg + n) + (((size_t) (2 * n + 1) * 2 + 15) & ~(size_t) 15);
  s.hist  = reinterpret_cast<int*>(p);
  s.cnt   = cnt;
  s.misc  = misc;
  return p + (size_t) (VOX_BUCKETS + 1) * 4;
}

// order-preserving bucket of a voxel key: the high bits of ((ix - min ix) << bits_y | (iy - min iy))
struct bucket_map {
  unsigned min_x, min_y;
  int bits_y, shift;
  __device__ __forceinline__ int of(unsigned long long key) const {
    const unsigned long long x = (unsigned) (key >> 44) - min_x, y = ((unsigned) (key >> 24) & 0xFFFFFu) - min_y;
    return (int) (((x << bits_y) | y) >> shift);
  }
};

// exclusive prefix of hist[0 .. VOX_BUCKETS) in place, hist[VOX_BUCKETS] = the total; returns sum(hist[k]^2) to every
// thread.  wsum: 64 ints of shared memory.  Ends with a barrier.
__device__ __forceinline__ unsigned bucket_scan(int* hist, int* wsum) {
  const int T = blockDim.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = T >> 5;
  const int per = (VOX_BUCKETS + T - 1) / T;
  const int k0 = min(tid * per, VOX_BUCKETS), k1 = min(k0 + per, VOX_BUCKETS);
  int sum     = 0;
  unsigned sq = 0;
  for (int k = k0; k < k1; ++k) {
    const int v = hist[k];
    sum += v, sq += (unsigned) (v * v);
  }
  int incl = sum;
  for (int o = 1; o < 32; o <<= 1) {
    const int u = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += u;
  }
  sq = __reduce_add_sync(0xffffffffu, sq);
  if (lane == 31) wsum[warp] = incl;
  if (lane == 0) wsum[32 + warp] = (int) sq;
  __syncthreads();
  const int wv        = lane < nwarp ? wsum[lane] : 0;
  const unsigned wq   = lane < nwarp ? (unsigned) wsum[32 + lane] : 0u;
  const int before    = __reduce_add_sync(0xffffffffu, lane < warp ? wv : 0);
  const unsigned totq = __reduce_add_sync(0xffffffffu, wq);
  int run = before + incl - sum;
  for (int k = k0; k < k1; ++k) {
    const int v = hist[k];
    hist[k]     = run;
    run += v;
  }
  if (tid == T - 1) hist[VOX_BUCKETS] = run;
  __syncthreads();
  return totq;
}

// in-place bitonic sort of key[0 .. S), ascending, in the all-ascending formulation (every merge starts with a "flip"
// stage, then half-cleaners): elements beyond S are virtual +infinity that never move, so only S slots exist.
// A thread always owns the same pair slots, and for spans of at most 64 both elements of a pair live in its
// warp's 64-element block: those stages need no block barrier.  Ends with a barrier.
static __device__ __noinline__ void bitonic_sort(unsigned long long* key, int S) {
  const int T = blockDim.x, tid = threadIdx.x;
  int cap = 1;
  while (cap < S) cap <<= 1;
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
}

// Block-wide voxelize of the n points (xy, nrm) with flags valid (nullptr: all valid): consecutive valid points that
// share a voxel form a SEGMENT (neighbouring beams hit neighbouring places); only the segments are sorted, by (voxel
// key, segment index), and a voxel's sum walks its segments in that order = cloud order.  rows(n_out) is called by
// every thread once the number of output points is known and returns where they go; the points are written in
// sorted order through finish(point).  Returns n_out to every thread.  Needs n <= 32 * blockDim.x and n < 2^14.
template <typename Rows, typename Finish>
__device__ __forceinline__ int block_voxelize(const voxel_params& V, const float2* xy, const float2* nrm,
                                              const unsigned char* valid, int n, const voxel_scratch& W, Rows rows,
                                              Finish finish) {
  const int T = blockDim.x, tid = threadIdx.x;
  const int pt_chunks = (n + T - 1) / T;
  int* total          = W.misc;
  // ---- per-point keys, the histogram cleared
  for (int k = tid; k < VOX_BUCKETS; k += T) W.hist[k] = 0;
  if (tid == 0) W.misc[1] = W.misc[3] = 0x7fffffff, W.misc[2] = W.misc[4] = 0;
  for (int c = 0; c < pt_chunks; ++c) {
    const int i = c * T + tid;
    if (i < n) W.keys[i] = (!valid || valid[i]) ? voxel_key(V, xy[i], nrm[i]) : NO_KEY;
  }
  __syncthreads();
  // ---- segment heads: valid points whose previous valid point has another key; extent of (ix, iy)
  unsigned mask = 0;
  int lo_x = 0x7fffffff, hi_x = 0, lo_y = 0x7fffffff, hi_y = 0;
  for (int c = 0; c < pt_chunks; ++c) {
    const int i = c * T + tid;
    if (i >= n) continue;
    const unsigned long long k = W.keys[i];
    if (k == NO_KEY) continue;
    int prev = i - 1;
    while (prev >= 0 && W.keys[prev] == NO_KEY) --prev;
    if (prev < 0 || W.keys[prev] != k) mask |= 1u << c;
    const int ix = (int) (k >> 44), iy = (int) ((unsigned) (k >> 24) & 0xFFFFFu);
    lo_x = min(lo_x, ix), hi_x = max(hi_x, ix), lo_y = min(lo_y, iy), hi_y = max(hi_y, iy);
  }
  lo_x = __reduce_min_sync(0xffffffffu, lo_x), hi_x = __reduce_max_sync(0xffffffffu, hi_x);
  lo_y = __reduce_min_sync(0xffffffffu, lo_y), hi_y = __reduce_max_sync(0xffffffffu, hi_y);
  if ((tid & 31) == 0) {
    atomicMin(&W.misc[1], lo_x), atomicMax(&W.misc[2], hi_x);
    atomicMin(&W.misc[3], lo_y), atomicMax(&W.misc[4], hi_y);
  }
  compact_count(mask, pt_chunks, W.cnt, total);
  const int S = *total;
  for (int c = 0; c < pt_chunks; ++c) {
    const int i = c * T + tid, dst = compact_slot(mask, c, W.cnt);
    if ((mask >> c) & 1u) {
      W.seg[dst]   = W.keys[i] | (unsigned long long) dst;
      W.first[dst] = (unsigned short) i;
    }
  }
  if (tid == 0) W.first[S] = (unsigned short) n;
  bucket_map B;
  B.min_x = (unsigned) W.misc[1], B.min_y = (unsigned) W.misc[3];
  {
    const int span_x = max(W.misc[2] - W.misc[1], 0), span_y = max(W.misc[4] - W.misc[3], 0);
    B.bits_y         = 32 - __clz(span_y);
    B.shift          = max(0, 32 - __clz(span_x) + B.bits_y - 11);  // 2048 buckets
  }
  __syncthreads();
  // ---- counting sort over the buckets: histogram, prefix, scatter
  const int seg_chunks = (S + T - 1) / T;
  for (int c = 0; c < seg_chunks; ++c) {
    const int t = c * T + tid;
    if (t < S) W.slot[t] = (unsigned short) atomicAdd(&W.hist[B.of(W.seg[t])], 1);
  }
  __syncthreads();
  const unsigned work = bucket_scan(W.hist, W.cnt);
  if (work > (unsigned) VOX_RANK_LIMIT * (unsigned) S) {
    bitonic_sort(W.seg, S);
  } else {
    for (int c = 0; c < seg_chunks; ++c) {
      const int t = c * T + tid;
      if (t < S) {
        const unsigned long long e = W.seg[t];
        W.keys[W.hist[B.of(e)] + W.slot[t]] = e;
      }
    }
    __syncthreads();
    // rank inside the bucket = the number of smaller bucket mates (keys are unique: they end in the segment index)
    for (int c = 0; c < seg_chunks; ++c) {
      const int t = c * T + tid;
      if (t < S) {
        const unsigned long long e = W.keys[t];
        const int b = B.of(e), b0 = W.hist[b], b1 = W.hist[b + 1];
        int r = 0;
        for (int q = b0; q < b1; ++q) r += W.keys[q] < e;
        W.seg[b0 + r] = e;
      }
    }
    __syncthreads();
  }
  // ---- one output point per run of equal voxel keys: the run's first segment sums all members in sorted (= cloud)
  // order; a segment's members are the valid points from its first point up to the next segment's
  mask = 0;
  for (int c = 0; c < seg_chunks; ++c) {
    const int t = c * T + tid;
    if (t < S && (t == 0 || ((W.seg[t] ^ W.seg[t - 1]) & VOXEL_MASK) != 0)) mask |= 1u << c;
  }
  compact_count(mask, seg_chunks, W.cnt, total);
  const int n_out = *total;
  float4* out     = rows(n_out);
  for (int c = 0; c < seg_chunks; ++c) {
    const int t = c * T + tid, dst = compact_slot(mask, c, W.cnt);
    if (!((mask >> c) & 1u)) continue;
    const unsigned long long vk = W.seg[t] & VOXEL_MASK;
    float4 o  = make_float4(0.f, 0.f, 0.f, 0.f);
    int count = 0;
    for (int e = t; e < S && (W.seg[e] & VOXEL_MASK) == vk; ++e) {
      const int s = (int) (W.seg[e] & 0x3FFF);
      for (int i = W.first[s], i1 = W.first[s + 1]; i < i1; ++i) {
        if (valid && !valid[i]) continue;
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
    out[dst] = finish(o);
  }
  return n_out;
}

__device__ __forceinline__ float2 lds_float2(unsigned a) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a));
  return v;
}

// (cos, sin) of azimuth = ifx * (b - cx) for every beam b (P1): the unprojection's only transcendental work
__global__ void beam_table_kernel(float ifx, float cx, int n_beams, float2* table) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n_beams) return;
  const float az = fmul(ifx, fsub((float) b, cx));
  table[b]       = make_float2(cosf_glibc(az), sinf_glibc(az));
}

// exclusive scan of counts[n] -> off[n + 1] by one CTA, chunks of blockDim.x; counts were written by other CTAs of this
// launch (read past L1).  warp_tot: 32 ints, carry: one int of shared memory.
// first: offset of item 0 (the end of the job's previous launch).
__device__ __forceinline__ void block_offsets(const int* counts, int n, int first, int* off, int* warp_tot, int* carry) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
  if (tid == 0) *carry = first;
  __syncthreads();
  for (int i0 = 0; i0 < n; i0 += blockDim.x) {
    const int i = i0 + tid;
    const int v = i < n ? __ldcg(counts + i) : 0;
    int incl    = v;
    for (int o = 1; o < 32; o <<= 1) {
      const int u = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += u;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    const int wv     = lane < nwarp ? warp_tot[lane] : 0;
    const int before = *carry + __reduce_add_sync(0xffffffffu, lane < warp ? wv : 0);
    const int all    = __reduce_add_sync(0xffffffffu, wv);
    if (i < n) off[i] = before + incl - v;
    __syncthreads();
    if (tid == 0) *carry += all;
    __syncthreads();
  }
  if (tid == 0) off[n] = *carry;
}

constexpr size_t scan_smem_bytes(int n_beams, bool voxelize) {
  return (size_t) (n_beams + 2) * 8 + (size_t) n_beams * 8 + (size_t) ((n_beams + 15) & ~15) +
         (voxelize ? voxel_scratch_bytes(n_beams) : 0);
}

__global__ void __launch_bounds__(SCAN_T) preprocess_kernel(const scan_dev_params P, const scan_args A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int NB = P.n_beams;
  float2* xy           = reinterpret_cast<float2*>(smem_raw) + 1;  // accepted beams, beam order; xy[-1], xy[n]: far away
  float2* nrm          = xy + NB + 1;                              // their normals
  unsigned char* valid = reinterpret_cast<unsigned char*>(nrm + NB);  // [NB]
  __shared__ int cnt[32 * (SCAN_T / 32)];
  __shared__ int misc[8];
  const int T = SCAN_T, tid = threadIdx.x;
  const int beam_chunks  = (NB + T - 1) / T;
  const unsigned xy_addr = (unsigned) __cvta_generic_to_shared(xy);
  const int scan      = blockIdx.x;
  const float* ranges = A.ranges + (size_t) scan * NB;
  // where the scan's `count` rows go; called by every thread once the count is known
  auto rows = [&](int count) -> float4* {
    if (tid == 0) A.counts[scan] = count;
    return A.out + (size_t) scan * NB;
  };
  // the CSR offsets of a packed output: the last CTA to finish scans the counts (scan_pack_kernel then moves the rows)
  auto finish = [&]() {
    if (!A.off) return;
    __syncthreads();
    if (tid == 0) {
      __threadfence();
      misc[6] = atomicAdd(A.ticket, 1);
    }
    __syncthreads();
    if (misc[6] != (int) gridDim.x - 1) return;
    __threadfence();
    block_offsets(A.counts, A.n_scans, A.continues ? __ldcg(A.off) : 0, A.off, cnt, &misc[7]);
    if (tid == 0) *A.ticket = 0;  // armed for the next launch
  };

  // ---- PointNormal2fUnprojectorPolar, accepted beams only, beam order (P1); staged through `nrm`
  unsigned mask = 0;
  for (int c = 0; c < beam_chunks; ++c) {
    const int b = c * T + tid;
    if (b < NB) {
      const float r = __ldg(ranges + b);
      if (!(r < P.range_min || r > P.range_max)) {
        const float2 cs = __ldg(A.beam_cs + b);
        nrm[b]          = make_float2(fmul(r, cs.x), fmul(r, cs.y));
        mask |= 1u << c;
      }
    }
  }
  compact_count(mask, beam_chunks, cnt, &misc[0]);
  for (int c = 0; c < beam_chunks; ++c) {
    const int dst = compact_slot(mask, c, cnt);
    if ((mask >> c) & 1u) xy[dst] = nrm[c * T + tid];
  }
  const int n = misc[0];
  if (tid == 0) xy[-1] = xy[n] = make_float2(__int_as_float(0x7f800000), __int_as_float(0x7f800000));
  __syncthreads();

  // ---- NormalComputator1DSlidingWindow (P2..P6): every point walks its own window, down to the first point that
  // is too far (the sentinels are), then up
  for (int i = tid; i < n; i += T) {
    const float2 p = xy[i];
    const f2 np    = mk2(-p.x, -p.y);
    // moments of d = q - p in walking order (P2, P4); the independent operations go out in pairs (FADD2 / FMUL2:
    // same roundings)
    f2 s1 = mk2(0.f, 0.f), sq = mk2(0.f, 0.f);  // (sum dx, sum dy), (sum dx^2, sum dy^2)
    float sxy = 0.f;
    // one step of the walk: false at the first point that is too far.  dd feeds both the test and the sums, which
    // keeps ptxas from contracting mul2 + add2 into one FFMA2 (tests/test_abi_symbols.py checks the SASS)
    auto take = [&](unsigned a) {
      const float2 v = lds_float2(a);
      const f2 d  = add2(mk2(v.x, v.y), np);
      const f2 dd = mul2(d, d);
      if (!(fadd(dd.x, dd.y) < P.d2)) return false;
      s1 = add2(s1, d), sq = add2(sq, dd);
      sxy = fadd(sxy, fmul(d.x, d.y));
      return true;
    };
    const unsigned a_i = xy_addr + 8u * (unsigned) i;  // shared-space byte addresses: 32-bit pointer arithmetic
    unsigned lo = a_i - 8u, hi = a_i + 8u;
    for (;; lo -= 16u) {
      if (!take(lo)) break;
      if (!take(lo - 8u)) {
        lo -= 8u;
        break;
      }
    }
    for (;; hi += 16u) {
      if (!take(hi)) break;
      if (!take(hi + 8u)) {
        hi += 8u;
        break;
      }
    }
    const int wn      = (int) ((hi - lo) >> 3) - 1;  // the points strictly between the two that were too far
    const float sxx = sq.x, syy = sq.y;
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

  if (P.inv_res == 0.f) {  // ---- valid points in cloud order (.cpp:44-48, P8)
    const int pt_chunks = (n + T - 1) / T;
    mask = 0;
    for (int c = 0; c < pt_chunks; ++c) {
      const int i = c * T + tid;
      if (i < n && valid[i]) mask |= 1u << c;
    }
    compact_count(mask, pt_chunks, cnt, &misc[0]);
    float4* out = rows(misc[0]);
    for (int c = 0; c < pt_chunks; ++c) {
      const int i = c * T + tid, dst = compact_slot(mask, c, cnt);
      if ((mask >> c) & 1u) out[dst] = make_float4(xy[i].x, xy[i].y, nrm[i].x, nrm[i].y);
    }
    finish();
    return;
  }

  // ---- voxelize(res, res, 1, 1) (.cpp:38-42, P7)
  voxel_params V;
  V.inv_res = P.inv_res, V.inv_n = 1.f, V.nb = 1;
  voxel_scratch W;
  carve_voxel_scratch(W, valid + ((NB + 15) & ~15), NB, cnt, misc);
  block_voxelize(V, xy, nrm, valid, n, W, rows, [](float4 o) { return o; });
  finish();
}

// strided [n_scans][n_beams] -> packed CSR; one CTA per scan
__global__ void scan_pack_kernel(const float4* strided, const int* off, int n_beams, float4* packed) {
  const int s = blockIdx.x;
  const int o = off[s], n = off[s + 1] - o;
  for (int i = threadIdx.x; i < n; i += blockDim.x) packed[o + i] = __ldcs(strided + (size_t) s * n_beams + i);
}

// SceneClipperProjective2D::compute with voxelize_resolution > 0 (R/mapping/scene_clipper_projective_2d.cpp:36-48):
// the z-buffer winners in column order, as points in the sensor frame, voxelized with res_coeffs (res, res, 0.1,
// 0.1), then moved into the robot frame.  One CTA per request; shared memory: clip_voxel_smem_bytes(canvas_cols).
constexpr size_t clip_voxel_smem_bytes(int cols) {
  return (size_t) cols * 16 + (size_t) ((cols + 3) & ~3) * 8 + voxel_scratch_bytes(cols);
}

__global__ void __launch_bounds__(SCAN_T) clip_voxel_kernel(const dev_params P, const clip_args A, float inv_res) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int C      = P.cam.cols;
  float2* xy       = reinterpret_cast<float2*>(smem_raw);
  float2* nrm      = xy + C;
  unsigned* zdepth = reinterpret_cast<unsigned*>(nrm + C);
  unsigned* zidx   = zdepth + ((C + 3) & ~3);
  __shared__ int cnt[32 * (SCAN_T / 32)];
  __shared__ int misc[8];
  voxel_scratch W;
  carve_voxel_scratch(W, reinterpret_cast<unsigned char*>(zidx + ((C + 3) & ~3)), C, cnt, misc);
  const int T = SCAN_T, tid = threadIdx.x;
  const int r     = blockIdx.x + A.base;
  const int cloud = A.cloud_ids[r];
  const int p0 = A.off[cloud], n = A.off[cloud + 1] - p0;
  const iso S   = load_pose(A.sensor_pose, 0, A.pose_stride);
  const iso cam = iso_compose(load_pose(A.robot_pose, (size_t) r, A.pose_stride), S);
  const iso Wc  = iso_inverse(cam);
  const bool move = !(S.c == 1.f && S.s == 0.f && S.tx == 0.f && S.ty == 0.f);  // .cpp:60
  zbuffer_project<false>(P, Wc, A.pts + p0, n, zdepth, zidx);
  // winners in column order (.cpp:38-43), in the sensor frame
  const int col_chunks = (C + T - 1) / T;
  unsigned mask = 0;
  for (int c = 0; c < col_chunks; ++c) {
    const int k = c * T + tid;
    if (k < C && zidx[k] != Z_EMPTY_IDX) mask |= 1u << c;
  }
  compact_count(mask, col_chunks, cnt, &misc[0]);
  for (int c = 0; c < col_chunks; ++c) {
    const int k = c * T + tid, dst = compact_slot(mask, c, cnt);
    if ((mask >> c) & 1u) {
      const float4 p = ldg4(A.pts + p0 + zidx[k]);
      float2 q, nq;
      iso_apply(Wc, p.x, p.y, q.x, q.y);
      iso_rot(Wc, p.z, p.w, nq.x, nq.y);
      xy[dst] = q, nrm[dst] = nq;
    }
  }
  const int m = misc[0];
  __syncthreads();
  voxel_params V;
  V.inv_res = inv_res, V.inv_n = fdiv(1.f, 0.1f), V.nb = 10;
  block_voxelize(
    V, xy, nrm, nullptr, m, W,
    [&](int count) { return output_rows(A.pack, A.out, (size_t) C, r, count, A.counts, &misc[5]); },
    [&](float4 o) {
      if (move) {  // .cpp:60-62
        float x, y, nx, ny;
        iso_apply(S, o.x, o.y, x, y);
        iso_rot(S, o.z, o.w, nx, ny);
        o = make_float4(x, y, nx, ny);
      }
      return o;
    });
}

}  // namespace ls2d
