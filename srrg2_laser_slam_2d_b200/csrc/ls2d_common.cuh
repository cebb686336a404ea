// ls2d_common.cuh -- pieces shared by every aligner kernel: device-side parameter / argument records, the pose block a
// CTA broadcasts once per iteration, the gates + factor + robustifier + H/b terms of ONE correspondence
// (linearize_point), the fixed-shape reductions, the binary64 3x3 solve and the result writers.
//
// Reference paths: R/ = /root/reference/srrg2_laser_slam_2d/src/srrg2_laser_slam_2d/.
#pragma once

#include <cuda_runtime.h>
#include <float.h>

#include "../../include/ls2d.h"
#include "ls2d_args.h"
#include "ls2d_math.cuh"

namespace ls2d {



constexpr unsigned Z_EMPTY_DEPTH = 0xFFFFFFFFu;
constexpr unsigned Z_EMPTY_IDX   = 0x7FFFFFFFu;
constexpr int NSUM               = 11;  // H00 H01 H02 H11 H12 H22 b0 b1 b2 chi_inliers chi_kernelized
constexpr int RED_STRIDE         = 12;  // + packed counts

__device__ __forceinline__ float4 ldg4(const float4* p) { return __ldg(p); }
// a point that is read exactly once by the kernel: do not let the stream of clouds evict the small tables (rounding
// edges of the projector) that live in the few KB of L1 left beside the shared-memory carve-out
__device__ __forceinline__ float4 ldg4_once(const float4* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}

// pose k of an array in the handle's pose format (include/ls2d.h): stride 3 = (x, y, theta) rebuilt with the glibc
// cosf / sinf copies, stride 4 = the caller's Isometry2f content, used verbatim
__device__ __forceinline__ iso load_pose(const float* poses, size_t k, int stride) {
  const float* p = poses + k * (size_t) stride;
  if (stride == 4) {
    iso T;
    T.tx = p[0], T.ty = p[1], T.c = p[2], T.s = p[3];
    return T;
  }
  return iso_v2t(p[0], p[1], p[2]);
}

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
// D[j] is added to the diagonal: Gauss-Newton's damping, or lambda * D_j of a Levenberg-Marquardt trial
__device__ __forceinline__ bool solve3d(const float* v, double D0, double D1, double D2, float* dx) {
  const double H00 = dadd((double) v[0], D0), H01 = v[1], H02 = v[2];
  const double H11 = dadd((double) v[3], D1), H12 = v[4];
  const double H22 = dadd((double) v[5], D2);
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
__device__ __forceinline__ bool solve3(const float* v, float damping, float* dx) {
  return solve3d(v, (double) damping, (double) damping, (double) damping, dx);
}

// pose block broadcast through shared memory once per iteration
struct pose_bc {
  float Xtx, Xty, Xc, Xs;  // estimate X = moving_in_fixed
  float Lc, Ls;            // rotation of local_map_in_sensor (= X, or Sinv * X)
  float Wtx, Wty;          // translation of inverse(inverse(local_map_in_sensor))  (decision D13)
  int stop;
  int tie;                 // set by a failed optimistic z-buffer claim
};

// the pose fields only (a Levenberg-Marquardt trial pose: the other threads may still be reading bc->stop)
__device__ __forceinline__ void publish_trial_pose(pose_bc* bc, const dev_params& P, const iso& X, bool with_sensor) {
  const iso L = with_sensor ? iso_compose(P.Sinv, X) : X;
  const iso W = iso_inverse(iso_inverse(L));
  bc->Xtx = X.tx, bc->Xty = X.ty, bc->Xc = X.c, bc->Xs = X.s;
  bc->Lc = W.c, bc->Ls = W.s;
  bc->Wtx = W.tx, bc->Wty = W.ty;
}
__device__ __forceinline__ void publish_pose(pose_bc* bc, const dev_params& P, const iso& X,
                                             bool with_sensor, int stop) {
  publish_trial_pose(bc, P, X, with_sensor);
  bc->stop = stop;
}

// RobustifierCauchy (L0.json:76-81; oracle decision D6): weight and statistics of one factor of squared error chi
// EXACT_LOG: the general kernel's Levenberg-Marquardt rounds decide on the kernelized chi2 and need glibc's logf bit
// for bit; the Gauss-Newton kernels only report it (tolerance parity) and keep the fast __logf.
template <bool EXACT_LOG>
__device__ __forceinline__ void cauchy(const dev_params& P, float chi, float& w, float& chi_in, float& chi_k,
                                       unsigned& cnt) {
  w = 1.f, chi_in = chi, chi_k = 0.f;
  if (P.tau > 0.f && !(chi < P.tau)) {
    const float aux = fadd(fmul(chi, P.inv_tau), 1.f);
    chi_k           = fmul(P.tau, EXACT_LOG ? logf_glibc(aux) : __logf(aux));
    w               = frcp(aux);
    chi_in          = 0.f;
    cnt += 1u << 16;
  } else {
    cnt += 1u;
  }
}

// gates of CorrespondenceFinderProjective2f (.cpp:61-73) + the slice's factor + Cauchy + H/b terms for the
// winner M of one column against the fixed cell (F, fd); adds into the thread's partial sums.  Returns true when the
// pair (F, M) is a correspondence (passed the finder's gates).
// GENERAL (icp_general_kernel): exact logf, and in an inlier-only round (decision I1) a kernelized factor keeps its
// statistics but adds nothing to H and b.
template <bool SENSOR, bool GENERAL = false>
__device__ __forceinline__ bool linearize_point(const dev_params& P, const pose_bc* bc, float fd, const float4 F,
                                                const float4 M, float rho, float Xtx, float Xty, float Lc, float Ls,
                                                float (&acc)[16], unsigned& cnt, bool inlier_only = false) {
  // every product / sum below is one binary32 rounding, in the oracle's order; independent ones are issued in
  // pairs (mul2 / add2, see ls2d_math.cuh for the no-contraction rule)
  if (fd < 0.f) return false;  // fcell.source_idx < 0
  if (fabsf(fsub(fd, rho)) > P.point_distance) return false;
  const f2 rc1 = mk2(Lc, Ls), rc2 = mk2(-Ls, Lc);  // columns of R(local_map_in_sensor)
  const f2 na = mul2s(rc1, M.z), nb = mul2s(rc2, M.w);
  const float nx = fadd(na.x, nb.x);  // transformed normal
  const float ny = fadd(na.y, nb.y);
  const f2 fn = mk2(F.z, F.w);
  const f2 nd = mul2(mk2(nx, ny), fn);
  if (fadd(nd.x, nd.y) < P.normal_cos) return false;
  f2 p;  // predicted point: X * M, then into the sensor frame (WithSensor)
  if (SENSOR) {
    const f2 qa = mul2s(mk2(bc->Xc, bc->Xs), M.x), qb = mul2s(mk2(-bc->Xs, bc->Xc), M.y);
    const float qx = fadd(fadd(qa.x, qb.x), Xtx);
    const float qy = fadd(fadd(qa.y, qb.y), Xty);
    iso_apply(P.Sinv, qx, qy, p.x, p.y);
  } else {
    const f2 pa = mul2s(rc1, M.x), pb = mul2s(rc2, M.y);
    p = add2(mk2(fadd(pa.x, pb.x), fadd(pa.y, pb.y)), mk2(Xtx, Xty));
  }
  const f2 d = add2(p, mk2(-F.x, -F.y));
  float w, chi_in, chi_k;
  if (P.factor == LS2D_FACTOR_POINT2POINT) {
    // SE2Point2PointErrorFactor[WithSensor] (oracle decision D19): e = p - p_fixed, J = [R | R (-y, x)^T], Omega = I2
    const float jc0 = fadd(fmul(Lc, -M.y), fmul(-Ls, M.x));
    const float jc1 = fadd(fmul(Ls, -M.y), fmul(Lc, M.x));
    const float chi = fadd(fmul(d.x, d.x), fmul(d.y, d.y));
    cauchy<GENERAL>(P, chi, w, chi_in, chi_k, cnt);
    if (GENERAL && inlier_only && w != 1.f) {
      acc[9] = fadd(acc[9], chi_in), acc[10] = fadd(acc[10], chi_k);
      return true;
    }
    const float wc = fmul(Lc, w), ws = fmul(Ls, w), wms = fmul(-Ls, w), wj0 = fmul(jc0, w), wj1 = fmul(jc1, w);
    acc[0]  = fadd(acc[0], fadd(fmul(wc, Lc), fmul(ws, Ls)));
    acc[1]  = fadd(acc[1], fadd(fmul(wc, -Ls), fmul(ws, Lc)));
    acc[2]  = fadd(acc[2], fadd(fmul(wc, jc0), fmul(ws, jc1)));
    acc[3]  = fadd(acc[3], fadd(fmul(wms, -Ls), fmul(wc, Lc)));
    acc[4]  = fadd(acc[4], fadd(fmul(wms, jc0), fmul(wc, jc1)));
    acc[5]  = fadd(acc[5], fadd(fmul(wj0, jc0), fmul(wj1, jc1)));
    acc[6]  = fadd(acc[6], fadd(fmul(wc, d.x), fmul(ws, d.y)));
    acc[7]  = fadd(acc[7], fadd(fmul(wms, d.x), fmul(wc, d.y)));
    acc[8]  = fadd(acc[8], fadd(fmul(wj0, d.x), fmul(wj1, d.y)));
    acc[9]  = fadd(acc[9], chi_in);
    acc[10] = fadd(acc[10], chi_k);
    return true;
  }
  // SE2Plane2PlaneErrorFactor (R/registration/aligner_slice_processor_laser_2d.h:8,23)
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
  cauchy<GENERAL>(P, chi, w, chi_in, chi_k, cnt);
  if (GENERAL && inlier_only && w != 1.f) {
    acc[9] = fadd(acc[9], chi_in), acc[10] = fadd(acc[10], chi_k);
    return true;
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
  return true;
}

// robustified squared error of the correspondence (F, M) at the pose in *bc -- the error half of linearize_point(),
// for the trial steps of Levenberg-Marquardt (computeActiveErrors: same correspondences, no Jacobians)
template <bool SENSOR>
__device__ __forceinline__ float correspondence_chi(const dev_params& P, const pose_bc* bc, const float4 F,
                                                    const float4 M) {
  const float Lc = bc->Lc, Ls = bc->Ls;
  const f2 rc1 = mk2(Lc, Ls), rc2 = mk2(-Ls, Lc);
  f2 p;
  if (SENSOR) {
    const f2 qa = mul2s(mk2(bc->Xc, bc->Xs), M.x), qb = mul2s(mk2(-bc->Xs, bc->Xc), M.y);
    const float qx = fadd(fadd(qa.x, qb.x), bc->Xtx);
    const float qy = fadd(fadd(qa.y, qb.y), bc->Xty);
    iso_apply(P.Sinv, qx, qy, p.x, p.y);
  } else {
    const f2 pa = mul2s(rc1, M.x), pb = mul2s(rc2, M.y);
    p = add2(mk2(fadd(pa.x, pb.x), fadd(pa.y, pb.y)), mk2(bc->Xtx, bc->Xty));
  }
  const f2 d = add2(p, mk2(-F.x, -F.y));
  float chi;
  if (P.factor == LS2D_FACTOR_POINT2POINT) {
    chi = fadd(fmul(d.x, d.x), fmul(d.y, d.y));
  } else {
    const f2 na = mul2s(rc1, M.z), nb = mul2s(rc2, M.w);
    const float nx = fadd(na.x, nb.x), ny = fadd(na.y, nb.y);
    const float e0 = fadd(fmul(d.x, F.z), fmul(d.y, F.w));
    const float e1 = fsub(nx, F.z), e2 = fsub(ny, F.w);
    chi = fadd(fadd(fmul(e0, e0), fmul(e1, e1)), fmul(e2, e2));
  }
  if (P.tau > 0.f && !(chi < P.tau)) chi = fmul(P.tau, logf_glibc(fadd(fmul(chi, P.inv_tau), 1.f)));
  return chi;
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
          st.c = X.c, st.s = X.s;
          A.iters[(size_t) pair * A.iters_stride + it] = st;
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
    r.c = bc->Xc, r.s = bc->Xs;
    r.lm_rejected = 0, r.reserved = 0;
    A.out[pair] = r;
  }
}

constexpr size_t icp_smem_bytes(int cols, int threads, int ppt) {
  return (size_t) cols * (16 + 4 + 4 + 4) + (size_t) threads * ppt * 8 + (size_t)(threads / 32) * RED_STRIDE * 4 +
         sizeof(pose_bc) + 16;
}

}  // namespace ls2d
