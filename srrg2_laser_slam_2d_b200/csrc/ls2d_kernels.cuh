// ls2d_kernels.cuh -- sm_100a kernels of the projective 2D registration path.
//
//  icp_fused_kernel   one CTA per scan pair; the whole MultiAligner2D::compute() loop on chip:
//                     fixed range image built once in shared memory, the moving cloud lives in
//                     REGISTERS for all iterations (owner-computes: the thread that owns a moving point
//                     projects it, fights for its column with two 32-bit shared-memory atomicMin passes,
//                     and -- if it won -- evaluates the correspondence itself), 11 sums + 2 counts reduced
//                     by a recursive-halving warp shuffle + one shared-memory stage, 3x3 solve and SE(2)
//                     update by thread 0.  HBM traffic = the compulsory bytes (each cloud read once,
//                     64 B written).
//  project_kernel     PointNormal2fProjectorPolar::compute for the drop-in projector API / parity.
//  correspond_kernel  CorrespondenceFinderProjective2f::compute for the drop-in finder API / parity.
//  best_of_kernel     acceptance gates + deterministic arg-best of a verification shard.
//
// Reference paths: R/ = /root/reference/srrg2_laser_slam_2d/src/srrg2_laser_slam_2d/.
#pragma once

#include <cuda_runtime.h>
#include <float.h>

#include "../../include/ls2d.h"
#include "ls2d_math.cuh"

namespace ls2d {

struct dev_params {
  polar_cam cam;
  float range_min, range_max;
  float point_distance, normal_cos;
  float tau, inv_tau;  // Cauchy threshold (<= 0: none) and 1/tau
  float damping;
  int max_iterations, min_num_correspondences, min_num_inliers;
  int with_sensor;
  iso Sinv;  // sensor_in_robot^-1
};

struct align_args {
  const float4* fixed_pts;
  const int* fixed_off;
  const float4* moving_pts;
  const int* moving_off;
  const int* fixed_id;   // nullable: pair index
  const int* moving_id;  // nullable: pair index
  int moving_div;        // moving_id == nullptr: moving cloud = pair / moving_div (verification: guesses)
  int fixed_const;       // >= 0: every pair uses this fixed cloud (verification: the query)
  const float* init_xyt;
  ls2d_result* out;
  ls2d_iter_stats* iters;  // nullable
  int n_pairs;
  int score_only;  // 1: one linearisation, no update
  int pair_base;   // first pair of this launch (chunked host pipeline); grid = n_pairs CTAs
};

constexpr unsigned Z_EMPTY_DEPTH = 0xFFFFFFFFu;
constexpr unsigned Z_EMPTY_IDX   = 0x7FFFFFFFu;
constexpr int NSUM               = 11;  // H00 H01 H02 H11 H12 H22 b0 b1 b2 chi_inliers chi_kernelized
constexpr int RED_STRIDE         = 12;  // + packed counts

__device__ __forceinline__ float4 ldg4(const float4* p) { return __ldg(p); }

// ---------------------------------------------------------------------------------------------------
// recursive-halving warp reduction of 11 floats: the value of slot s ends up on lane 2*s.  Every slot is
// summed along the xor-butterfly tree (offsets 16, 8, 4, 2, 1), so the result is the one a full butterfly
// gives (binary32 addition is commutative) at 16 shuffles instead of 55.
__device__ __forceinline__ float warp_reduce_slots(float (&v)[16], int lane) {
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const bool up    = lane & 16;
    const float send = up ? v[k] : v[k + 8];
    const float keep = up ? v[k + 8] : v[k];
    v[k]             = fadd(keep, __shfl_xor_sync(0xffffffffu, send, 16));
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const bool up    = lane & 8;
    const float send = up ? v[k] : v[k + 4];
    const float keep = up ? v[k + 4] : v[k];
    v[k]             = fadd(keep, __shfl_xor_sync(0xffffffffu, send, 8));
  }
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const bool up    = lane & 4;
    const float send = up ? v[k] : v[k + 2];
    const float keep = up ? v[k + 2] : v[k];
    v[k]             = fadd(keep, __shfl_xor_sync(0xffffffffu, send, 4));
  }
  {
    const bool up    = lane & 2;
    const float send = up ? v[0] : v[1];
    const float keep = up ? v[1] : v[0];
    v[0]             = fadd(keep, __shfl_xor_sync(0xffffffffu, send, 2));
  }
  v[0] = fadd(v[0], __shfl_xor_sync(0xffffffffu, v[0], 1));
  return v[0];
}

// 3x3 LDL^T in binary64, operation order of oracle/ls2d_oracle.c solve3() (decision D11)
__device__ __forceinline__ bool solve3(const float* v, float damping, float* dx) {
  const double H00 = dadd((double) v[0], (double) damping), H01 = v[1], H02 = v[2];
  const double H11 = dadd((double) v[3], (double) damping), H12 = v[4];
  const double H22 = dadd((double) v[5], (double) damping);
  const double r0 = -(double) v[6], r1 = -(double) v[7], r2 = -(double) v[8];
  if (!(H00 > 0.0)) return false;
  const double i0  = drcp(H00);
  const double l10 = dmul(H01, i0), l20 = dmul(H02, i0);
  const double d1  = dsub(H11, dmul(l10, H01));
  if (!(d1 > 0.0)) return false;
  const double i1  = drcp(d1);
  const double t21 = dsub(H12, dmul(l20, H01));
  const double l21 = dmul(t21, i1);
  const double d2  = dsub(dsub(H22, dmul(l20, H02)), dmul(l21, t21));
  if (!(d2 > 0.0)) return false;
  const double i2 = drcp(d2);
  const double z1 = dsub(r1, dmul(l10, r0));
  const double z2 = dsub(dsub(r2, dmul(l20, r0)), dmul(l21, z1));
  const double y0 = dmul(r0, i0), y1 = dmul(z1, i1), x2 = dmul(z2, i2);
  const double x1 = dsub(y1, dmul(l21, x2));
  const double x0 = dsub(dsub(y0, dmul(l10, x1)), dmul(l20, x2));
  dx[0]           = (float) x0;
  dx[1]           = (float) x1;
  dx[2]           = (float) x2;
  return isfinite(dx[0]) && isfinite(dx[1]) && isfinite(dx[2]);
}

// pose block broadcast through shared memory once per iteration
struct pose_bc {
  float Xtx, Xty, Xc, Xs;  // estimate X = moving_in_fixed
  float Lc, Ls;            // rotation of local_map_in_sensor (= X, or Sinv * X)
  float Wtx, Wty;          // translation of inverse(inverse(local_map_in_sensor))  (decision D13)
  int stop;
  int tie;                 // set by a failed optimistic z-buffer claim
};

__device__ __forceinline__ void publish_pose(pose_bc* bc, const dev_params& P, const iso& X,
                                             bool with_sensor, int stop) {
  const iso L = with_sensor ? iso_compose(P.Sinv, X) : X;
  const iso W = iso_inverse(iso_inverse(L));
  bc->Xtx = X.tx, bc->Xty = X.ty, bc->Xc = X.c, bc->Xs = X.s;
  bc->Lc = W.c, bc->Ls = W.s;
  bc->Wtx = W.tx, bc->Wty = W.ty;
  bc->stop = stop;
}

// gates of CorrespondenceFinderProjective2f (.cpp:61-73) + SE2Plane2PlaneErrorFactor + Cauchy + H/b terms for the
// winner M of one column against the fixed cell (F, fd); adds into the thread's partial sums.
template <bool SENSOR>
__device__ __forceinline__ void linearize_point(const dev_params& P, const pose_bc* bc, float fd, const float4 F,
                                                const float4 M, float rho, float Xtx, float Xty, float Lc, float Ls,
                                                float (&acc)[16], unsigned& cnt) {
  // every product / sum below is one binary32 rounding, in the oracle's order; independent ones are issued in
  // pairs (mul2 / add2, see ls2d_math.cuh for the no-contraction rule)
  do {
      if (fd < 0.f) break;  // fcell.source_idx < 0
      if (fabsf(fsub(fd, rho)) > P.point_distance) break;
      const f2 rc1 = mk2(Lc, Ls), rc2 = mk2(-Ls, Lc);  // columns of R(local_map_in_sensor)
      const f2 na = mul2s(rc1, M.z), nb = mul2s(rc2, M.w);
      const float nx = fadd(na.x, nb.x);  // transformed normal
      const float ny = fadd(na.y, nb.y);
      const f2 fn = mk2(F.z, F.w);
      const f2 nd = mul2(mk2(nx, ny), fn);
      if (fadd(nd.x, nd.y) < P.normal_cos) break;
      // SE2Plane2PlaneErrorFactor (R/registration/aligner_slice_processor_laser_2d.h:8,23)
      f2 p;
      if (SENSOR) {
        const f2 qa = mul2s(mk2(bc->Xc, bc->Xs), M.x), qb = mul2s(mk2(-bc->Xs, bc->Xc), M.y);
        const float qx = fadd(fadd(qa.x, qb.x), Xtx);
        const float qy = fadd(fadd(qa.y, qb.y), Xty);
        iso_apply(P.Sinv, qx, qy, p.x, p.y);
      } else {
        const f2 pa = mul2s(rc1, M.x), pb = mul2s(rc2, M.y);
        p = add2(mk2(fadd(pa.x, pb.x), fadd(pa.y, pb.y)), mk2(Xtx, Xty));
      }
      const f2 d  = add2(p, mk2(-F.x, -F.y));
      const f2 de = mul2(d, fn);
      const float e0 = fadd(de.x, de.y);
      const f2 e12 = add2(mk2(nx, ny), mk2(-F.z, -F.w));  // e1, e2
      const f2 ja = mul2s(mk2(Lc, -Ls), F.z), jb = mul2s(mk2(Ls, Lc), F.w);
      const float Ja = fadd(ja.x, jb.x);
      const float Jb = fadd(ja.y, jb.y);
      const f2 jc = mul2(mk2(Ja, Jb), mk2(-M.y, M.x));
      const float Jc = fadd(jc.x, jc.y);
      const float d0 = -ny, d1 = nx;  // R * (-n.y, n.x)^T, exact in binary32
      const f2 ee = mul2(e12, e12);
      const float chi = fadd(fadd(fmul(e0, e0), ee.x), ee.y);
      float w = 1.f, chi_in = chi, chi_k = 0.f;
      if (P.tau > 0.f && !(chi < P.tau)) {  // RobustifierCauchy (L0.json:76-81)
        const float aux = fadd(fmul(chi, P.inv_tau), 1.f);
        chi_k           = fmul(P.tau, __logf(aux));  // statistics only (tolerance parity)
        w               = frcp(aux);
        chi_in          = 0.f;
        cnt += 1u << 16;
      } else {
        cnt += 1u;
      }
      const f2 wab = mul2s(mk2(Ja, Jb), w);  // wa, wb
      const float wc = fmul(Jc, w);
      const f2 wd = mul2s(mk2(d0, d1), w);   // wd0, wd1
      const f2 h01 = mul2s(mk2(Ja, Jb), wab.x);            // wa*Ja, wa*Jb
      const f2 h23 = mul2(wab, mk2(Jc, Jb));               // wa*Jc, wb*Jb
      const f2 h4c = mul2(mk2(wab.y, wc), mk2(Jc, Jc));    // wb*Jc, wc*Jc
      const f2 hdd = mul2(wd, mk2(d0, d1));                // wd0*d0, wd1*d1
      const f2 b01 = mul2s(wab, e0);                       // wa*e0, wb*e0
      const f2 bde = mul2(wd, e12);                        // wd0*e1, wd1*e2
      acc[0]  = fadd(acc[0], h01.x);
      acc[1]  = fadd(acc[1], h01.y);
      acc[2]  = fadd(acc[2], h23.x);
      acc[3]  = fadd(acc[3], h23.y);
      acc[4]  = fadd(acc[4], h4c.x);
      acc[5]  = fadd(acc[5], fadd(fadd(h4c.y, hdd.x), hdd.y));
      acc[6]  = fadd(acc[6], b01.x);
      acc[7]  = fadd(acc[7], b01.y);
      acc[8]  = fadd(acc[8], fadd(fadd(fmul(wc, e0), bde.x), bde.y));
      acc[9]  = fadd(acc[9], chi_in);
      acc[10] = fadd(acc[10], chi_k);
  } while (false);
}

// reduction stage 1: warp shuffle tree, one partial row per warp in shared memory
__device__ __forceinline__ void store_partials(float (&acc)[16], unsigned cnt, float* red, int lane, int warp) {
  const float wsum    = warp_reduce_slots(acc, lane);
  const unsigned wcnt = __reduce_add_sync(0xffffffffu, cnt);
  if (!(lane & 1) && lane < 2 * NSUM) red[warp * RED_STRIDE + (lane >> 1)] = wsum;
  if (lane == 0) red[warp * RED_STRIDE + NSUM] = __uint_as_float(wcnt);
}

// reduction stage 2 + Gauss-Newton step, executed by warp 0 after the barrier: lanes 0..10 add the warps' partials in
// warp order, lane 0 checks the correspondence gate, solves the 3x3 system in binary64, applies X <- X * v2t(dx)
// (VariableSE2Right) and publishes the next pose.
// CANON: add +0.0f to the totals (a kernel whose threads ASSIGN their first contribution can produce -0 where the
// oracle's 0 + x gives +0; the sum with +0 maps both to the same bits and changes nothing else).
template <int T, bool SENSOR, bool CANON = false>
__device__ __forceinline__ void warp0_update(const dev_params& P, const align_args& A, pose_bc* bc, const float* red,
                                             int pair, int it, int lane, float& tot, unsigned& tot_cnt) {
  if (lane < NSUM) {
    tot = red[lane];
#pragma unroll
    for (int w = 1; w < T / 32; ++w) tot = fadd(tot, red[w * RED_STRIDE + lane]);
    if (CANON) tot = fadd(tot, 0.f);
  } else if (lane == NSUM) {
    tot_cnt = 0;
#pragma unroll
    for (int w = 0; w < T / 32; ++w) tot_cnt += __float_as_uint(red[w * RED_STRIDE + NSUM]);
  }
  float v[NSUM];
#pragma unroll
  for (int s = 0; s < NSUM; ++s) v[s] = __shfl_sync(0xffffffffu, tot, s);
  const unsigned c2 = __shfl_sync(0xffffffffu, tot_cnt, NSUM);
  if (lane == 0) {
    const int n_in = c2 & 0xffff, n_k = c2 >> 16, n_corr = n_in + n_k;
    iso X;
    X.tx = bc->Xtx, X.ty = bc->Xty, X.c = bc->Xc, X.s = bc->Xs;
    int stop = 0;
    float dx[3];
    if (n_corr <= P.min_num_correspondences) {
      stop = 1 + LS2D_STATUS_NOT_ENOUGH_CORRESPONDENCES;
    } else if (!A.score_only) {
      if (!solve3(v, P.damping, dx)) {
        stop = 1 + LS2D_STATUS_SINGULAR;
      } else {
        X = iso_compose(X, iso_v2t(dx[0], dx[1], dx[2]));
        if (A.iters) {
          ls2d_iter_stats st;
          st.x = X.tx, st.y = X.ty, st.theta = atan2f_fdlibm(X.s, X.c);
          st.chi_inliers = v[9], st.chi_kernelized = v[10];
          st.n_inliers = n_in, st.n_kernelized = n_k, st.n_corr = n_corr;
          A.iters[(size_t) pair * P.max_iterations + it] = st;
        }
      }
    }
    publish_pose(bc, P, X, SENSOR, stop);
  }
}

// final record of a pair, written by thread 0 (called by the whole warp 0)
__device__ __forceinline__ void write_result(const dev_params& P, const align_args& A, const pose_bc* bc, int pair,
                                             int it, int status, float tot, unsigned tot_cnt, int writer_tid = 0) {
  float v[NSUM];
#pragma unroll
  for (int s = 0; s < NSUM; ++s) v[s] = __shfl_sync(0xffffffffu, tot, s);
  const unsigned c2 = __shfl_sync(0xffffffffu, tot_cnt, NSUM);
  if ((int) threadIdx.x == writer_tid) {
    int n_in = c2 & 0xffff, n_k = c2 >> 16;
    const int n_corr = n_in + n_k;
    if (status == LS2D_STATUS_NOT_ENOUGH_CORRESPONDENCES) {  // the oracle reports empty sums here
#pragma unroll
      for (int s = 0; s < NSUM; ++s) v[s] = 0.f;
      n_in = n_k = 0;
    }
    if (status < 0) status = n_in < P.min_num_inliers ? LS2D_STATUS_NOT_ENOUGH_INLIERS : LS2D_STATUS_SUCCESS;
    ls2d_result r;
    r.x = bc->Xtx, r.y = bc->Xty, r.theta = atan2f_fdlibm(bc->Xs, bc->Xc);
    r.chi_inliers = v[9], r.chi_kernelized = v[10];
    r.n_inliers = n_in, r.n_kernelized = n_k, r.n_corr = n_corr;
    r.status = status, r.iterations = it;
#pragma unroll
    for (int s = 0; s < 6; ++s) r.H[s] = v[s];
    A.out[pair] = r;
  }
}

constexpr size_t icp_smem_bytes(int cols, int threads, int ppt) {
  return (size_t) cols * (16 + 4 + 4 + 4) + (size_t) threads * ppt * 8 + (size_t)(threads / 32) * RED_STRIDE * 4 +
         sizeof(pose_bc) + 16;
}

template <int T, int PPT, bool SENSOR, int MINB>
__global__ void __launch_bounds__(T, MINB) icp_fused_kernel(const dev_params P, const align_args A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int C      = P.cam.cols;
  float4* fimg     = reinterpret_cast<float4*>(smem_raw);           // fixed image: x y nx ny per column
  float* fdepth    = reinterpret_cast<float*>(fimg + C);            // fixed image: rho, < 0 = empty
  unsigned* zdepth = reinterpret_cast<unsigned*>(fdepth + C);       // z-buffer pass 1: min rho bits
  unsigned* zidx   = zdepth + C;                                    // z-buffer pass 2: min index among ties
  float2* mnrm     = reinterpret_cast<float2*>(zidx + C + (C & 1)); // [T * PPT] moving normals (phase 2 only)
  float* red       = reinterpret_cast<float*>(mnrm + T * PPT);      // [T/32][RED_STRIDE]
  pose_bc* bc      = reinterpret_cast<pose_bc*>(red + (T / 32) * RED_STRIDE);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int pair = blockIdx.x + A.pair_base;
  const int fcl  = A.fixed_const >= 0 ? A.fixed_const : (A.fixed_id ? A.fixed_id[pair] : pair);
  const int mcl  = A.moving_id ? A.moving_id[pair / A.moving_div] : pair / A.moving_div;
  const int f0 = A.fixed_off[fcl], nf = A.fixed_off[fcl + 1] - f0;
  const int m0 = A.moving_off[mcl], nm = A.moving_off[mcl + 1] - m0;

  for (int k = tid; k < C; k += T) {
    fdepth[k] = -1.f;
    zdepth[k] = Z_EMPTY_DEPTH;
    zidx[k]   = Z_EMPTY_IDX;
  }
  // moving cloud (issued early; consumed after the fixed image is built): coordinates -> registers for all
  // iterations, normals -> shared memory (only winners read them; the register allocator would spill them to
  // local memory otherwise)
  float2 mp[PPT];
#pragma unroll
  for (int j = 0; j < PPT; ++j) {
    const int i    = tid + j * T;
    const float4 m = i < nm ? ldg4(A.moving_pts + m0 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    mp[j]          = make_float2(m.x, m.y);
    mnrm[i]        = make_float2(m.z, m.w);
  }
  if (tid == 0) {
    const iso X = iso_v2t(A.init_xyt[3 * pair], A.init_xyt[3 * pair + 1], A.init_xyt[3 * pair + 2]);
    publish_pose(bc, P, X, SENSOR, 0);
    bc->tie = 0;
  }
  __syncthreads();

  // ---- fixed range image: identity camera (R/registration/correspondence_finder_projective_2d.cpp:37-44)
  {
    float4 fp[PPT];
    int col[PPT];
    unsigned rb[PPT];
#pragma unroll
    for (int j = 0; j < PPT; ++j) {
      const int i = tid + j * T;
      col[j]      = -1;
      rb[j]       = 0;
      if (i < nf) {
        fp[j]           = ldg4(A.fixed_pts + f0 + i);
        const float rho = fsqrt(fadd(fmul(fp[j].x, fp[j].x), fmul(fp[j].y, fp[j].y)));
        if (!(rho < P.range_min || rho > P.range_max)) {
          col[j] = polar_column(P.cam, fp[j].y, fp[j].x);
          rb[j]  = f2u(rho);
          if (col[j] >= 0) atomicMin(&zdepth[col[j]], rb[j]);
        }
      }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < PPT; ++j)
      if (col[j] >= 0 && zdepth[col[j]] == rb[j]) atomicMin(&zidx[col[j]], (unsigned) (tid + j * T));
    __syncthreads();
#pragma unroll
    for (int j = 0; j < PPT; ++j)
      if (col[j] >= 0 && zidx[col[j]] == (unsigned) (tid + j * T)) {
        fimg[col[j]]   = fp[j];
        fdepth[col[j]] = u2f(rb[j]);
      }
    __syncthreads();
    // hand the z-buffer back empty for the moving cloud (a separate pass: the losers of a column were still
    // reading zidx above; every toucher writes the same EMPTY values)
#pragma unroll
    for (int j = 0; j < PPT; ++j)
      if (col[j] >= 0) {
        zdepth[col[j]] = Z_EMPTY_DEPTH;
        zidx[col[j]]   = Z_EMPTY_IDX;
      }
    __syncthreads();
  }

  // ---- ICP loop (MultiAligner2D::compute; L0.json:487-517)
  const int max_it = A.score_only ? 1 : P.max_iterations;
  int it           = 0;
  int status       = -1;
  float tot        = 0.f;  // lane s of warp 0: total of slot s for the last linearisation
  unsigned tot_cnt = 0;
  // z-buffer winner = lowest index among the points of minimal rho (decision D3).  Equal rho bits in one
  // column are rare, so an iteration first runs OPTIMISTIC: one atomicMin pass, then every minimal-rho point
  // claims its cell with a CAS; a failed claim means a tie, and the whole iteration is redone EXACT with the
  // index tie-break pass (one more barrier).  `exact` is uniform over the CTA.
  bool exact = false;
  for (; it < max_it; ++it) {
    if (exact) __syncthreads();  // redo pass: tie flag cleared and all cells handed back
    int col[PPT];
    unsigned rb[PPT];
    // phase 1: project the moving cloud (camera = local_map_in_sensor^-1, .cpp:47-48), z-buffer pass 1
    {
    const float Lc = bc->Lc, Ls = bc->Ls, Wtx = bc->Wtx, Wty = bc->Wty;  // phase-1 copies die at the barrier
    // the three points of a thread run as independent straight-line chains (transform, rho, fast column); the rare
    // points whose column only the exact atan2 may decide are visited afterwards
    f2 pc[PPT];
    bool near[PPT];
#pragma unroll
    for (int j = 0; j < PPT; ++j) {
      const f2 ra = mul2s(mk2(Lc, Ls), mp[j].x), rb2 = mul2s(mk2(-Ls, Lc), mp[j].y);
      pc[j]       = add2(mk2(fadd(ra.x, rb2.x), fadd(ra.y, rb2.y)), mk2(Wtx, Wty));
      const f2 pq = mul2(pc[j], pc[j]);
      const float rho = fsqrt(fadd(pq.x, pq.y));
      rb[j]       = f2u(rho);
      col[j]      = polar_column_fast(P.cam, pc[j].y, pc[j].x, near[j]);
      near[j]     = near[j] && !(rho < P.range_min || rho > P.range_max);
    }
    bool any_near = false;
#pragma unroll
    for (int j = 0; j < PPT; ++j) any_near |= near[j];
    if (any_near) {
#pragma unroll
      for (int j = 0; j < PPT; ++j)
        if (near[j]) col[j] = polar_column_exact(P.cam, pc[j].y, pc[j].x);
    }
#pragma unroll
    for (int j = 0; j < PPT; ++j) {
      const float rho = u2f(rb[j]);
      const bool ok   = tid + j * T < nm && !(rho < P.range_min || rho > P.range_max) && col[j] >= 0 && col[j] < C;
      col[j]          = ok ? col[j] : -1;
      if (ok) atomicMin(&zdepth[col[j]], rb[j]);
    }
    }
    __syncthreads();
    const float Xtx = bc->Xtx, Xty = bc->Xty, Lc = bc->Lc, Ls = bc->Ls;  // re-read: shorter live ranges than 6 registers
    if (exact) {  // z-buffer pass 2: lowest index among equal depths
#pragma unroll
      for (int j = 0; j < PPT; ++j)
        if (col[j] >= 0 && zdepth[col[j]] == rb[j]) atomicMin(&zidx[col[j]], (unsigned) (tid + j * T));
      __syncthreads();
    }
    // phase 2: winners gate against the fixed column (.cpp:61-73) and linearise their correspondence
    float acc[16];
#pragma unroll
    for (int s = 0; s < 16; ++s) acc[s] = 0.f;
    unsigned cnt = 0;  // n_inliers | n_kernelized << 16
#pragma unroll
    for (int j = 0; j < PPT; ++j) {
      if (col[j] < 0 || zdepth[col[j]] != rb[j]) continue;
      const int c = col[j];
      if (exact) {
        if (zidx[c] != (unsigned) (tid + j * T)) continue;
      } else if (atomicCAS(&zidx[c], Z_EMPTY_IDX, (unsigned) (tid + j * T)) != Z_EMPTY_IDX) {
        bc->tie = 1;  // two points of equal minimal rho in one column: redo this iteration exactly
        continue;
      }
      const float2 mn = mnrm[tid + j * T];
      linearize_point<SENSOR>(P, bc, fdepth[c], fimg[c], make_float4(mp[j].x, mp[j].y, mn.x, mn.y), u2f(rb[j]), Xtx,
                              Xty, Lc, Ls, acc, cnt);
    }
    store_partials(acc, cnt, red, lane, warp);
    __syncthreads();
    // hand the touched cells back for the next pass (every toucher writes the same EMPTY values)
#pragma unroll
    for (int j = 0; j < PPT; ++j)
      if (col[j] >= 0) {
        zdepth[col[j]] = Z_EMPTY_DEPTH;
        zidx[col[j]]   = Z_EMPTY_IDX;
      }
    if (!exact && bc->tie) {  // uniform: bc->tie was written before the barrier above
      __syncthreads();        // every thread has read the flag and handed its cells back
      if (tid == 0) bc->tie = 0;
      exact = true;
      --it;
      continue;               // the next pass starts after the barrier at the loop head
    }
    exact = false;
    if (warp == 0) warp0_update<T, SENSOR>(P, A, bc, red, pair, it, lane, tot, tot_cnt);
    __syncthreads();
    if (bc->stop) {
      status = bc->stop - 1;
      break;
    }
  }

  if (tid < 32) write_result(P, A, bc, pair, it, status, tot, tot_cnt);
}

// Phase 1 of the streaming kernels: every moving point is seen from the camera (W = inverse(inverse(
// local_map_in_sensor)), decision D13), its rho and column are stashed in shared memory and its rho fights for the
// column's z-buffer cell.  U points of a thread at a time run as straight-line chains; the rare points whose
// column only the exact atan2 may decide are visited afterwards.
template <int T, int U, typename Load>
__device__ __forceinline__ void project_and_stash(const dev_params& P, const pose_bc* bc, int nm, Load load,
                                                  unsigned short* scol, unsigned* srho, unsigned* zdepth) {
  const float Lc = bc->Lc, Ls = bc->Ls, Wtx = bc->Wtx, Wty = bc->Wty;
  const int C = P.cam.cols;
  for (int i0 = threadIdx.x; i0 < nm; i0 += U * T) {
    f2 pc[U];
    unsigned rbv[U];
    int colv[U];
    bool near[U], up[U];
    bool any_near = false;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i    = i0 + u * T;
      const float4 M = i < nm ? load(i) : make_float4(0.f, 0.f, 0.f, 0.f);
      const f2 ra = mul2s(mk2(Lc, Ls), M.x), rb2 = mul2s(mk2(-Ls, Lc), M.y);
      pc[u]       = add2(mk2(fadd(ra.x, rb2.x), fadd(ra.y, rb2.y)), mk2(Wtx, Wty));
      const f2 pq = mul2(pc[u], pc[u]);
      const float rho = fsqrt(fadd(pq.x, pq.y));
      rbv[u]      = f2u(rho);
      colv[u]     = polar_column_fast2(P.cam, pc[u].y, pc[u].x, near[u], up[u]);
      near[u]     = near[u] && i < nm && !(rho < P.range_min || rho > P.range_max);
      any_near |= near[u];
    }
    if (any_near) {  // rare: side of the rounding edge's ray (when the camera carries an edge table), then exact atan2f
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (near[u]) {
          bool undecided;
          const int c2 = polar_column_edge(P.cam, pc[u].y, pc[u].x, u2f(rbv[u]), colv[u] + (up[u] ? 1 : 0), undecided);
          colv[u]      = undecided ? polar_column_exact(P.cam, pc[u].y, pc[u].x) : c2;
        }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = i0 + u * T;
      if (i < nm) {
        const float rho = u2f(rbv[u]);
        const bool ok   = !(rho < P.range_min || rho > P.range_max) && colv[u] >= 0 && colv[u] < C;
        scol[i]         = (unsigned short) (ok ? colv[u] : 0xFFFF);
        srho[i]         = rbv[u];
        if (ok) atomicMin(&zdepth[colv[u]], rbv[u]);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// icp_stream_kernel: same algorithm and the same reduction shape (thread t owns points t, t+T, ...; ascending) for
// clouds of any size.  Per-point state does not live in registers: column and rho of every point are stashed in
// shared memory (6 B/point) by the projection pass and re-read by the later passes; the points themselves are
// either staged in shared memory once (MP_SMEM, 16 B/point) or re-read from global memory / L2 every pass.
constexpr size_t icp_stream_smem_bytes(int cols, int threads, int max_points, bool mp_smem) {
  return (size_t) cols * (16 + 4 + 4 + 4) + (size_t)(threads / 32) * RED_STRIDE * 4 + sizeof(pose_bc) + 16 +
         (size_t) max_points * 4 + (size_t)((max_points + 1) / 2) * 4 + 16 + (mp_smem ? (size_t) max_points * 16 : 0);
}

template <int T, bool SENSOR, bool MP_SMEM, int MINB>
__global__ void __launch_bounds__(T, MINB) icp_stream_kernel(const dev_params P, const align_args A, int max_points) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int C          = P.cam.cols;
  float4* fimg         = reinterpret_cast<float4*>(smem_raw);
  float4* smp          = fimg + C;                                            // MP_SMEM: the moving cloud
  float* fdepth        = reinterpret_cast<float*>(smp + (MP_SMEM ? max_points : 0));
  unsigned* zdepth     = reinterpret_cast<unsigned*>(fdepth + C);
  unsigned* zidx       = zdepth + C;
  unsigned* srho       = zidx + C;                                            // [max_points] rho bits
  unsigned short* scol = reinterpret_cast<unsigned short*>(srho + max_points);  // [max_points] column, 0xFFFF = none
  float* red           = reinterpret_cast<float*>(scol + 2 * ((max_points + 1) / 2));
  pose_bc* bc          = reinterpret_cast<pose_bc*>(red + (T / 32) * RED_STRIDE);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int pair = blockIdx.x + A.pair_base;
  const int fcl  = A.fixed_const >= 0 ? A.fixed_const : (A.fixed_id ? A.fixed_id[pair] : pair);
  const int mcl  = A.moving_id ? A.moving_id[pair / A.moving_div] : pair / A.moving_div;
  const int f0 = A.fixed_off[fcl], nf = A.fixed_off[fcl + 1] - f0;
  const int m0 = A.moving_off[mcl], nm = A.moving_off[mcl + 1] - m0;
  const float4* fpts = A.fixed_pts + f0;
  const float4* mpts = A.moving_pts + m0;

  for (int k = tid; k < C; k += T) {
    fdepth[k] = -1.f;
    zdepth[k] = Z_EMPTY_DEPTH;
    zidx[k]   = Z_EMPTY_IDX;
  }
  if (MP_SMEM)
    for (int i = tid; i < nm; i += T) smp[i] = ldg4(mpts + i);
  if (tid == 0) {
    const iso X = iso_v2t(A.init_xyt[3 * pair], A.init_xyt[3 * pair + 1], A.init_xyt[3 * pair + 2]);
    publish_pose(bc, P, X, SENSOR, 0);
    bc->tie = 0;
  }
  __syncthreads();

  // ---- fixed range image (identity camera), exact two-pass z-buffer
  for (int i = tid; i < nf; i += T) {
    const float4 p  = ldg4(fpts + i);
    const float rho = fsqrt(fadd(fmul(p.x, p.x), fmul(p.y, p.y)));
    int col         = -1;
    if (!(rho < P.range_min || rho > P.range_max)) col = polar_column(P.cam, p.y, p.x);
    scol[i] = (unsigned short) (col < 0 ? 0xFFFF : col);
    srho[i] = f2u(rho);
    if (col >= 0) atomicMin(&zdepth[col], f2u(rho));
  }
  __syncthreads();
  for (int i = tid; i < nf; i += T) {
    const unsigned c = scol[i];
    if (c != 0xFFFF && zdepth[c] == srho[i]) atomicMin(&zidx[c], (unsigned) i);
  }
  __syncthreads();
  for (int i = tid; i < nf; i += T) {
    const unsigned c = scol[i];
    if (c != 0xFFFF && zidx[c] == (unsigned) i) {
      fimg[c]   = ldg4(fpts + i);
      fdepth[c] = u2f(srho[i]);
    }
  }
  __syncthreads();
  for (int i = tid; i < nf; i += T) {
    const unsigned c = scol[i];
    if (c != 0xFFFF) zdepth[c] = Z_EMPTY_DEPTH, zidx[c] = Z_EMPTY_IDX;
  }
  __syncthreads();

  const int max_it = A.score_only ? 1 : P.max_iterations;
  int it           = 0;
  int status       = -1;
  float tot        = 0.f;
  unsigned tot_cnt = 0;
  bool exact       = false;
  for (; it < max_it; ++it) {
    if (exact) __syncthreads();
    project_and_stash<T, 4>(P, bc, nm, [&](int i) { return MP_SMEM ? smp[i] : ldg4(mpts + i); }, scol, srho, zdepth);
    __syncthreads();
    if (exact) {
      for (int i = tid; i < nm; i += T) {
        const unsigned c = scol[i];
        if (c != 0xFFFF && zdepth[c] == srho[i]) atomicMin(&zidx[c], (unsigned) i);
      }
      __syncthreads();
    }
    const float Xtx = bc->Xtx, Xty = bc->Xty, Lc = bc->Lc, Ls = bc->Ls;
    float acc[16];
#pragma unroll
    for (int s = 0; s < 16; ++s) acc[s] = 0.f;
    unsigned cnt = 0;
    // phase 2: a thread first finds the z-buffer winners among its next 32 points (cheap scan), then linearises only
    // those, in ascending order: with far more points than columns most points lose, and a warp now runs the heavy
    // path max-winners-per-lane times instead of once per scanned point
    for (int k0 = 0; tid + k0 * T < nm; k0 += 32) {
      unsigned wmask = 0;
#pragma unroll 4
      for (int k = 0; k < 32; ++k) {
        const int i = tid + (k0 + k) * T;
        if (i >= nm) break;
        const unsigned c = scol[i];
        if (c == 0xFFFF || zdepth[c] != srho[i]) continue;
        if (exact) {
          if (zidx[c] != (unsigned) i) continue;
        } else if (atomicCAS(&zidx[c], Z_EMPTY_IDX, (unsigned) i) != Z_EMPTY_IDX) {
          bc->tie = 1;
          continue;
        }
        wmask |= 1u << k;
      }
      while (wmask) {
        const int k = __ffs(wmask) - 1;
        wmask &= wmask - 1;
        const int i      = tid + (k0 + k) * T;
        const unsigned c = scol[i];
        const float4 M   = MP_SMEM ? smp[i] : ldg4(mpts + i);
        linearize_point<SENSOR>(P, bc, fdepth[c], fimg[c], M, u2f(srho[i]), Xtx, Xty, Lc, Ls, acc, cnt);
      }
    }
    store_partials(acc, cnt, red, lane, warp);
    __syncthreads();
    for (int k = tid; k < C; k += T) zdepth[k] = Z_EMPTY_DEPTH, zidx[k] = Z_EMPTY_IDX;  // wholesale: C <= points
    if (!exact && bc->tie) {
      __syncthreads();
      if (tid == 0) bc->tie = 0;
      exact = true;
      --it;
      continue;
    }
    exact = false;
    if (warp == 0) warp0_update<T, SENSOR>(P, A, bc, red, pair, it, lane, tot, tot_cnt);
    __syncthreads();
    if (bc->stop) {
      status = bc->stop - 1;
      break;
    }
  }
  if (tid < 32) write_result(P, A, bc, pair, it, status, tot, tot_cnt);
}

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

struct project_args {
  const float4* pts;
  const int* off;
  int cloud;
  float cam_xyt[3];
  int* source_idx;  // [C]
  float* depth;     // [C]
};

__global__ void project_kernel(const dev_params P, const project_args A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int C      = P.cam.cols;
  unsigned* zdepth = reinterpret_cast<unsigned*>(smem_raw);
  unsigned* zidx   = zdepth + C;
  const int p0 = A.off[A.cloud], n = A.off[A.cloud + 1] - p0;
  const iso W = iso_inverse(iso_v2t(A.cam_xyt[0], A.cam_xyt[1], A.cam_xyt[2]));
  zbuffer_project<false>(P, W, A.pts + p0, n, zdepth, zidx);
  for (int k = threadIdx.x; k < C; k += blockDim.x) {
    const bool empty = zidx[k] == Z_EMPTY_IDX;
    A.source_idx[k]  = empty ? -1 : (int) zidx[k];
    A.depth[k]       = empty ? FLT_MAX : u2f(zdepth[k]);
  }
}

struct correspond_args {
  const float4* fixed_pts;
  const int* fixed_off;
  const float4* moving_pts;
  const int* moving_off;
  int fixed_cloud, moving_cloud;
  float lmis_xyt[3];  // local_map_in_sensor
  int* fixed_idx;     // [C]
  int* moving_idx;    // [C]
  int* count;
};

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
  const iso L = iso_v2t(A.lmis_xyt[0], A.lmis_xyt[1], A.lmis_xyt[2]);
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
struct clip_args {
  const float4* pts;
  const int* off;
  const int* cloud_ids;     // [n]
  const float* robot_xyt;   // [n * 3] robot_in_local_map
  float sensor_xyt[3];      // sensor_in_robot
  float4* out;              // [n * C]
  int* counts;              // [n]
};

__global__ void clip_kernel(const dev_params P, const clip_args A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int C      = P.cam.cols;
  unsigned* zdepth = reinterpret_cast<unsigned*>(smem_raw);
  unsigned* zidx   = zdepth + C;
  __shared__ int warp_tot[32];
  __shared__ int base;
  const int r      = blockIdx.x;
  const int cloud  = A.cloud_ids[r];
  const int p0 = A.off[cloud], n = A.off[cloud + 1] - p0;
  const iso S   = iso_v2t(A.sensor_xyt[0], A.sensor_xyt[1], A.sensor_xyt[2]);
  const iso cam = iso_compose(iso_v2t(A.robot_xyt[3 * r], A.robot_xyt[3 * r + 1], A.robot_xyt[3 * r + 2]), S);
  const iso W   = iso_inverse(cam);
  const bool move = !(S.c == 1.f && S.s == 0.f && S.tx == 0.f && S.ty == 0.f);  // .cpp:60
  zbuffer_project<false>(P, W, A.pts + p0, n, zdepth, zidx);
  if (threadIdx.x == 0) base = 0;
  __syncthreads();
  for (int k0 = 0; k0 < C; k0 += blockDim.x) {
    const int k   = k0 + threadIdx.x;
    const bool ok = k < C && zidx[k] != Z_EMPTY_IDX;
    float4 o      = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ok) {
      const float4 p = ldg4(A.pts + p0 + zidx[k]);
      iso_apply(W, p.x, p.y, o.x, o.y);
      iso_rot(W, p.z, p.w, o.z, o.w);
      if (move) {
        float x, y, nx, ny;
        iso_apply(S, o.x, o.y, x, y);
        iso_rot(S, o.z, o.w, nx, ny);
        o = make_float4(x, y, nx, ny);
      }
    }
    const int dst = ordered_slot(ok, warp_tot, &base);
    if (ok) A.out[(size_t) r * C + dst] = o;
  }
  if (threadIdx.x == 0) A.counts[r] = base;
}

// ---------------------------------------------------------------------------------------------------
// MergerProjective2D::compute (R/mapping/merger_projective_2d.cpp:9-100): both clouds projected from
// measurement_in_scene, per-column add / average+renormalise / replace / append; the scene is updated in place
// and grows by an ordered append.  One CTA per (scene, measurement) request.
struct merge_args {
  float4* scene;        // in/out, `capacity` points
  int* scene_size;      // in/out
  int capacity;
  const float4* meas;
  int n_meas;
  float mis_xyt[3];     // measurement_in_scene
  float merge_threshold;
  int* counters;        // [4]: new, merged, replaced, overflow flag
};

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
  const iso M = iso_v2t(A.mis_xyt[0], A.mis_xyt[1], A.mis_xyt[2]);
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
    } else {
      const ls2d_result r = res[bi];
      b.x = r.x, b.y = r.y, b.theta = r.theta, b.chi_inliers = r.chi_inliers;
      b.n_inliers = r.n_inliers, b.n_corr = r.n_corr;
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
    } else {
      const ls2d_result r = res[bi];
      b.x = r.x, b.y = r.y, b.theta = r.theta, b.chi_inliers = r.chi_inliers;
      b.n_inliers = r.n_inliers, b.n_corr = r.n_corr;
      b.candidate = moving_id ? moving_id[bi] : bi;
      b.guess     = bi - p0;
    }
    out[grp] = b;
  }
}

}  // namespace ls2d
