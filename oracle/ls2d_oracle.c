/*
 * ls2d_oracle.c -- CPU restatement of the reference's projective 2D registration path.
 * TEST INFRASTRUCTURE ONLY (see ls2d_oracle.h).  PARITY UNPINNED (see ls2d_oracle.h).
 *
 * Reference paths are relative to /root/reference/srrg2_laser_slam_2d/ ; R/ abbreviates
 * src/srrg2_laser_slam_2d/ ; L0.json = /root/reference/configurations/
 * stage_segway_double_config_LASER_0.json.
 *
 * Decision points (result-affecting choices the in-repo sources do not pin):
 *  D1  column index = lrintf(u) (round to nearest, ties to even) -- the unprojector puts beam i
 *      exactly at u = i (R/sensor_processing/raw_data_preprocessor_projective_2d.cpp:87-90).
 *  D2  K00 = (float)C / (angle_col_max - angle_col_min), K01 = (float)C * 0.5f, all binary32
 *      (camera-matrix layout: apps/synthetic_scene_generator.cpp:60-66).
 *  D2b range gate: reject rho < range_min || rho > range_max.
 *  D3  z-buffer: strict "rho < cell.depth" => nearest wins, first in iteration order wins ties.
 *  D4  empty cell: source_idx = -1, depth = FLT_MAX.
 *  D5  all points are Valid (no status channel; invalid beams are encoded as far points by the
 *      generators).
 *  D6  Cauchy: chi < tau => inlier (w = 1); else kernelized: a = chi/tau + 1, rho = tau*ln a,
 *      w = 1/a.  tau <= 0 => no robustifier.
 *  D7  error order [point-to-line ; dn.x ; dn.y].
 *  D8  information matrix Omega = I3.
 *  D9  WithSensor: finder gets S^-1 * X; factor predicts S^-1 * (X * p_m), rotations S^-1 * R.
 *  D10 sum order: sequential in correspondence order (reference) or the CUDA tree (debug).
 *  D11 3x3 solve in binary64 LDL^T with reciprocal pivots (Cholmod is double; its simplicial
 *      factorisation is LDL^T); a pivot <= 0 means "not positive definite" => status SINGULAR.
 *  D12 pose state = Isometry2f content (tx, ty, c, s); X <- X * v2t(dx) by plain products,
 *      no re-orthonormalisation; theta reported as atan2f(s, c).
 *  D13 the camera pose goes through TWO Isometry2f::inverse() calls before touching a point
 *      (finder: setCameraPose(local_map_in_sensor.inverse()),
 *      R/registration/correspondence_finder_projective_2d.cpp:47; projector applies the inverse
 *      of its camera pose) -- restated literally, since inverse(inverse(T)) != T in binary32.
 *  Multi-slice aligner (MultiAligner2D with several slice processors; MULTI.json:700-730):
 *  D14 every laser slice finds its correspondences with its own projector / finder / sensor_in_robot
 *      and linearises with its own robustifier; a slice with n_corr <= min_num_correspondences is
 *      skipped in that iteration (no H/b/statistics contribution); if no laser slice contributes the
 *      aligner stops with NotEnoughCorrespondences.  n_corr reports the contributing slices' total
 *      (on that failure: everything the finders produced).
 *  D15 AlignerSliceOdom2DPrior = SE2PriorErrorFactor on VariableSE2Right: e = t2v(Z^-1 * X),
 *      J = blockdiag(R(Z^-1 * X), 1) (exact derivative of e for X <- X * v2t(dx)), Omega as given;
 *      chi = e^T (Omega e); Cauchy as D6 when the slice has a robustifier (the configurations: none).
 *      The factor counts as ONE inlier (or one kernelized factor) in the statistics, not as a
 *      correspondence.
 *  D16 H and b are the binary32 sums of the slices' totals in slice order, the prior last.
 *  D17 max_iterations, min_num_inliers and damping are aligner-level (read from slice 0's record).
 *  D18 fused accumulation arithmetic of the compile-time-stride kernels (see contribution_fused below).
 *  Options north_star names but the shipped configurations never switch on (all UNPINNED restatements):
 *  D19 factor POINT2POINT = SE2Point2PointErrorFactor[WithSensor] on VariableSE2Right: e = X p_m - p_f
 *      (S^-1 (X p_m) - p_f), J = [R | R (-y, x)^T] with R the rotation of X (of S^-1 X), Omega = I2,
 *      chi = e0^2 + e1^2; the finder and its normal gate are unchanged.
 *  L1  algorithm LM = IterationAlgorithmLM (the Levenberg-Marquardt of g2o / srrg2_solver, Nielsen's update):
 *      per round, H, b and chi0 = chi_inliers + chi_kernelized are built at X with the round's correspondences.
 *  L2  lambda is initialised in the first round (user_lambda_init if > 0, else tau * max diag H) and carried
 *      across the rounds of one compute(); nu = 2 at every round's start.
 *  L3  trial: solve (H + lambda D) dx = -b, D = diag H (variable_damping) or I, binary64 LDL^T as D11; a system
 *      that is not positive definite counts as a rejected trial.
 *  L4  chi1 = robustified chi of the SAME correspondences at X * v2t(dx) (computeActiveErrors: no new association;
 *      a factor's inlier / kernelized decision is re-taken at the trial pose), binary32, summed in the selected
 *      order.
 *  L5  scale = sum_j dx_j (lambda D_j dx_j - b_j) + 1e-3, rho = (chi0 - chi1) / scale, binary64.
 *  L6  rho > 0 and chi1 finite: accept (X <- X * v2t(dx)), lambda *= max(step_low, min(1 - (2 rho - 1)^3,
 *      step_high)), round done.  Else lambda *= nu, nu *= 2, next trial (at most lm_iterations_max per round).
 *  L7  no accepted trial: X stays; the round still counts as an iteration.  `damping` is not used by LM.
 *  L8  lambda, nu, rho, scale live in binary64.
 *  I1  enable_inlier_only_runs: after the main rounds, if they did not fail and n_inliers >= min_num_inliers, up
 *      to max_iterations more rounds run in which kernelized factors get weight 0 (they still count in the
 *      statistics); their iteration records follow the main ones.
 *  T1  termination_epsilon > 0: after a round (not the first of its phase) with
 *      chi_prev - chi < epsilon * chi_prev, chi = chi_inliers + chi_kernelized of the round's linearisation, the
 *      phase stops.
 */
#include "ls2d_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

void orc_default_params(orc_params* p) {
  memset(p, 0, sizeof(*p));
  p->canvas_cols             = 721;      /* L0.json:328 */
  p->angle_col_min           = -3.14159f; /* L0.json:319 */
  p->angle_col_max           = 3.14159f;  /* L0.json:316 */
  p->range_min               = 0.3f;     /* L0.json:337 */
  p->range_max               = 20.f;     /* L0.json:334 */
  p->point_distance          = 0.5f;     /* correspondence_finder_projective_2d.h:19 */
  p->normal_cos              = 0.8f;     /* correspondence_finder_projective_2d.h:21 */
  p->cauchy_chi_threshold    = 0.01f;    /* L0.json:80 */
  p->damping                 = 0.f;      /* L0.json:87 */
  p->max_iterations          = 10;       /* L0.json:498 */
  p->min_num_correspondences = 0;        /* L0.json:134 */
  p->min_num_inliers         = 10;       /* L0.json:501 */
  p->with_sensor             = 0;
  p->factor                  = ORC_FACTOR_PLANE2PLANE;
  p->algorithm               = ORC_ALGORITHM_GN;
  p->lm_user_lambda_init     = 0.f;
  p->lm_tau                  = 1e-5f;
  p->lm_step_low             = 1.f / 3.f;
  p->lm_step_high            = 2.f / 3.f;
  p->lm_iterations_max       = 10;
  p->lm_variable_damping     = 1;
}

/* sensor_in_robot of a WithSensor slice in either form (with_sensor 1: x, y, theta; 2: the isometry itself) */
static orc_iso orc_v2t_fwd(float x, float y, float theta);
static orc_iso sensor_iso(const orc_params* prm) {
  if (prm->with_sensor == 2) {
    orc_iso S;
    S.tx = prm->sensor_in_robot[0], S.ty = prm->sensor_in_robot[1];
    S.c = prm->sensor_in_robot_cs[0], S.s = prm->sensor_in_robot_cs[1];
    return S;
  }
  return orc_v2t_fwd(prm->sensor_in_robot[0], prm->sensor_in_robot[1], prm->sensor_in_robot[2]);
}

static orc_iso pose_at(const float* poses, size_t k, int32_t stride) {
  const float* p = poses + k * (size_t) stride;
  if (stride == 4) {
    orc_iso T;
    T.tx = p[0], T.ty = p[1], T.c = p[2], T.s = p[3];
    return T;
  }
  return orc_v2t_fwd(p[0], p[1], p[2]);
}

/* ---------------------------------------------------------------- A.0 geometry helpers */

/* geometry2d::v2t [srrg2_core; used at apps/visual_test_correspondence_finder_projective_2d.cpp:71] */
orc_iso orc_v2t(float x, float y, float theta) {
  orc_iso T;
  T.tx = x;
  T.ty = y;
  T.c  = cosf(theta);
  T.s  = sinf(theta);
  return T;
}
static orc_iso orc_v2t_fwd(float x, float y, float theta) { return orc_v2t(x, y, theta); }

/* geometry2d::t2v [used at apps/visual_test_aligner_2d.cpp:145]: theta = atan2(R10, R00) */
void orc_t2v(orc_iso T, float* xyt) {
  xyt[0] = T.tx;
  xyt[1] = T.ty;
  xyt[2] = atan2f(T.s, T.c);
}

/* R * v for R = [c -s; s c]: Eigen evaluates row0 = R00*v0 + R01*v1 with R01 = -s */
static inline void rot(orc_iso T, float vx, float vy, float* ox, float* oy) {
  *ox = T.c * vx + (-T.s) * vy;
  *oy = T.s * vx + T.c * vy;
}

/* Isometry2f * Vector2f = linear()*v + translation() */
static inline void apply(orc_iso T, float vx, float vy, float* ox, float* oy) {
  float rx, ry;
  rot(T, vx, vy, &rx, &ry);
  *ox = rx + T.tx;
  *oy = ry + T.ty;
}

/* Eigen Transform::inverse(Isometry): linear = R^T, translation = -(R^T * t) */
orc_iso orc_inverse(orc_iso T) {
  orc_iso I;
  I.c = T.c;
  I.s = -T.s;
  float rx, ry;
  rot(I, T.tx, T.ty, &rx, &ry);
  I.tx = -rx;
  I.ty = -ry;
  return I;
}

/* Isometry2f * Isometry2f: linear = Ra*Rb (first column kept; the second is its exact
 * rotation by 90 degrees in binary32, see D12), translation = Ra*tb + ta */
orc_iso orc_compose(orc_iso A, orc_iso B) {
  orc_iso C;
  C.c = A.c * B.c + (-A.s) * B.s;
  C.s = A.s * B.c + A.c * B.s;
  apply(A, B.tx, B.ty, &C.tx, &C.ty);
  return C;
}

/* ---------------------------------------------------------------- A.1 polar projector */

/* PointNormal2fProjectorPolar::compute [srrg2_core srrg_pcl/point_projector*.h; call sites
 * R/registration/correspondence_finder_projective_2d.cpp:40-41,47-48] */
void orc_project(const orc_params* prm, orc_iso camera_pose, const orc_point* pts, int32_t n,
                 orc_cell* image) {
  const int32_t C = prm->canvas_cols;
  const float K00 = (float) C / (prm->angle_col_max - prm->angle_col_min); /* D2 */
  const float K01 = (float) C * 0.5f;
  const orc_iso W = orc_inverse(camera_pose); /* world -> camera (D13) */
  for (int32_t k = 0; k < C; ++k) {
    image[k].source_idx = -1; /* D4 */
    image[k].depth      = FLT_MAX;
    image[k].px = image[k].py = image[k].nx = image[k].ny = 0.f;
  }
  for (int32_t i = 0; i < n; ++i) {
    float px, py, nx, ny;
    apply(W, pts[i].x, pts[i].y, &px, &py);
    rot(W, pts[i].nx, pts[i].ny, &nx, &ny);
    const float rho = sqrtf(px * px + py * py);
    if (rho < prm->range_min || rho > prm->range_max) { /* D2b */
      continue;
    }
    const float theta = atan2f(py, px);
    const float u     = K00 * theta + K01;
    const long col    = lrintf(u); /* D1 */
    if (col < 0 || col >= C) {
      continue;
    }
    orc_cell* cell = &image[col];
    if (rho < cell->depth) { /* D3 */
      cell->source_idx = i;
      cell->depth      = rho;
      cell->px         = px;
      cell->py         = py;
      cell->nx         = nx;
      cell->ny         = ny;
    }
  }
}

/* ---------------------------------------------------------------- A.2 correspondence finder */

/* R/registration/correspondence_finder_projective_2d.cpp:46-76, line by line */
int32_t orc_find_correspondences(const orc_params* prm, const orc_cell* fixed_image,
                                 const orc_point* moving, int32_t n_moving,
                                 orc_iso local_map_in_sensor, orc_cell* moving_image,
                                 int32_t* fixed_idx, int32_t* moving_idx) {
  /* :47-48  the camera sits on the fixed */
  orc_project(prm, orc_inverse(local_map_in_sensor), moving, n_moving, moving_image);
  int32_t k = 0;
  for (int32_t col = 0; col < prm->canvas_cols; ++col) { /* :55-59 lock-step walk */
    const orc_cell* m = &moving_image[col];
    const orc_cell* f = &fixed_image[col];
    if (m->source_idx < 0 || f->source_idx < 0) { /* :61 */
      continue;
    }
    if (fabsf(f->depth - m->depth) > prm->point_distance) { /* :65 */
      continue;
    }
    if (m->nx * f->nx + m->ny * f->ny < prm->normal_cos) { /* :69 */
      continue;
    }
    fixed_idx[k]  = f->source_idx; /* :73 Correspondence(fixed, moving) */
    moving_idx[k] = m->source_idx;
    ++k;
  }
  return k; /* :76 */
}

/* ---------------------------------------------------------------- A.3 factor */

typedef struct {
  orc_iso X;      /* estimate (moving_in_fixed, robot frame) */
  orc_iso Sinv;   /* sensor_in_robot^-1 (WithSensor) */
  orc_iso RX;     /* rotation used by the Jacobian: R (plain) or Rs^-1 * R (WithSensor) */
  int with_sensor;
} factor_ctx;

/* e, J entries of SE2Plane2PlaneErrorFactor [srrg2_solver types_2d/se2_plane2plane_error_factor;
 * type bound at R/registration/aligner_slice_processor_laser_2d.h:8,23]; 2D reduction of
 * octave/solver/nicp_post.m:13-25 with the post-multiplied increment of nicp_post.m:96.
 * J = [ Ja Jb Jc ; 0 0 d0 ; 0 0 d1 ]. */
static inline void factor_eval(const factor_ctx* f, orc_point pf, orc_point pm, float* e, float* Ja,
                               float* Jb, float* Jc, float* d0, float* d1) {
  float px, py, nx, ny;
  apply(f->X, pm.x, pm.y, &px, &py); /* p_pred = X * p_moving */
  if (f->with_sensor) {              /* D9 */
    float qx, qy;
    apply(f->Sinv, px, py, &qx, &qy);
    px = qx;
    py = qy;
  }
  rot(f->RX, pm.nx, pm.ny, &nx, &ny); /* n_pred = R * n_moving */
  const float dx = px - pf.x;
  const float dy = py - pf.y;
  e[0] = dx * pf.nx + dy * pf.ny; /* nt * (p_pred - p_fixed)   nicp_post.m:20 */
  e[1] = nx - pf.nx;              /* n_pred - n_fixed          nicp_post.m:21 */
  e[2] = ny - pf.ny;
  *Ja = pf.nx * f->RX.c + pf.ny * f->RX.s;    /* nt * R        nicp_post.m:23 */
  *Jb = pf.nx * (-f->RX.s) + pf.ny * f->RX.c;
  *Jc = *Ja * (-pm.y) + *Jb * pm.x;           /* nt * R * (-y, x)^T   nicp_post.m:24 */
  rot(f->RX, -pm.ny, pm.nx, d0, d1);          /* R * (-n.y, n.x)^T    nicp_post.m:25 */
}

static factor_ctx make_factor(const orc_params* prm, orc_iso X) {
  factor_ctx f;
  f.X           = X;
  f.with_sensor = prm->with_sensor;
  f.RX          = X;
  f.Sinv        = orc_v2t(0.f, 0.f, 0.f);
  if (prm->with_sensor) {
    f.Sinv = orc_inverse(sensor_iso(prm));
    f.RX = orc_compose(f.Sinv, X); /* only its rotation is used */
  }
  return f;
}

void orc_error_and_jacobian(const orc_params* prm, orc_iso X, orc_point fixed, orc_point moving,
                            float* e, float* J) {
  factor_ctx f = make_factor(prm, X);
  float Ja, Jb, Jc, d0, d1;
  if (prm->factor == ORC_FACTOR_POINT2POINT) { /* D19: e = p_pred - p_fixed (2 rows, the third is zero), J = [R | R (-y, x)^T] */
    float px, py, jc0, jc1;
    apply(f.X, moving.x, moving.y, &px, &py);
    if (f.with_sensor) {
      float qx, qy;
      apply(f.Sinv, px, py, &qx, &qy);
      px = qx;
      py = qy;
    }
    rot(f.RX, -moving.y, moving.x, &jc0, &jc1);
    e[0] = px - fixed.x, e[1] = py - fixed.y, e[2] = 0.f;
    J[0] = f.RX.c, J[1] = -f.RX.s, J[2] = jc0;
    J[3] = f.RX.s, J[4] = f.RX.c, J[5] = jc1;
    J[6] = J[7] = J[8] = 0.f;
    return;
  }
  factor_eval(&f, fixed, moving, e, &Ja, &Jb, &Jc, &d0, &d1);
  J[0] = Ja;
  J[1] = Jb;
  J[2] = Jc;
  J[3] = 0.f;
  J[4] = 0.f;
  J[5] = d0;
  J[6] = 0.f;
  J[7] = 0.f;
  J[8] = d1;
}

/* ---------------------------------------------------------------- A.4 / A.5 linearisation */

/* slots: 0..5 = H00 H01 H02 H11 H12 H22, 6..8 = b, 9 = chi_inliers, 10 = chi_kernelized */
#define NSLOT 11

/* one correspondence's contribution: robustifier (A.4) + J^T (w Omega) J, J^T (w Omega) e (A.5).
 * returns 1 if inlier, 0 if kernelized */
static inline int contribution(const orc_params* prm, const factor_ctx* f, orc_point pf,
                               orc_point pm, float* v) {
  float e[3], Ja, Jb, Jc, d0, d1;
  factor_eval(f, pf, pm, e, &Ja, &Jb, &Jc, &d0, &d1);
  const float chi = (e[0] * e[0] + e[1] * e[1]) + e[2] * e[2]; /* e^T Omega e, Omega = I (D8) */
  float w         = 1.f;
  int inlier      = 1;
  float chi_in = chi, chi_k = 0.f;
  const float tau = prm->cauchy_chi_threshold;
  if (tau > 0.f && !(chi < tau)) { /* D6 */
    const float inv_tau = 1.f / tau;
    const float aux     = chi * inv_tau + 1.f;
    chi_k               = tau * logf(aux);
    w                   = 1.f / aux;
    chi_in              = 0.f;
    inlier              = 0;
  }
  const float wa = Ja * w, wb = Jb * w, wc = Jc * w, wd0 = d0 * w, wd1 = d1 * w;
  v[0]  = wa * Ja;
  v[1]  = wa * Jb;
  v[2]  = wa * Jc;
  v[3]  = wb * Jb;
  v[4]  = wb * Jc;
  v[5]  = (wc * Jc + wd0 * d0) + wd1 * d1;
  v[6]  = wa * e[0];
  v[7]  = wb * e[0];
  v[8]  = (wc * e[0] + wd0 * e[1]) + wd1 * e[2];
  v[9]  = chi_in;
  v[10] = chi_k;
  return inlier;
}

/* D18 -- the same contribution with the kernel's FUSED arithmetic (ORC_SUM_TREE with bit 17 of tree_threads set;
 * icp_fused2_kernel): every gate and everything that decides a pixel index stays single binary32 operations, but the
 * error / Jacobian entries and the accumulation into the thread's partial sums p[] use fused multiply-adds, in
 * exactly this association.  (The reference's own builds are not uniquely defined here either: GCC's default
 * -ffp-contract=fast contracts Eigen's expressions wherever the target has FMA.)  Returns 1 if inlier. */
static inline int contribution_fused(const orc_params* prm, const factor_ctx* f, orc_point pf, orc_point pm, float* p) {
  const float c = f->RX.c, s = f->RX.s;
  float px, py, nx, ny;
  if (f->with_sensor) { /* D9: unfused, as factor_eval */
    float qx, qy;
    apply(f->X, pm.x, pm.y, &qx, &qy);
    apply(f->Sinv, qx, qy, &px, &py);
  } else {
    px = fmaf(c, pm.x, fmaf(-s, pm.y, f->X.tx));
    py = fmaf(s, pm.x, fmaf(c, pm.y, f->X.ty));
  }
  rot(f->RX, pm.nx, pm.ny, &nx, &ny); /* the finder's gate used this very value */
  const float dx = px - pf.x, dy = py - pf.y;
  const float e0 = fmaf(dx, pf.nx, dy * pf.ny);
  const float e1 = nx - pf.nx, e2 = ny - pf.ny;
  const float Ja = fmaf(pf.nx, c, pf.ny * s);
  const float Jb = fmaf(pf.ny, c, pf.nx * (-s));
  const float Jc = fmaf(Ja, -pm.y, Jb * pm.x);
  const float d0 = -ny, d1 = nx;
  const float chi = fmaf(e2, e2, fmaf(e1, e1, e0 * e0));
  float w = 1.f, chi_in = chi, chi_k = 0.f;
  int inlier      = 1;
  const float tau = prm->cauchy_chi_threshold;
  if (tau > 0.f && !(chi < tau)) { /* D6 */
    const float inv_tau = 1.f / tau;
    const float aux     = chi * inv_tau + 1.f;
    chi_k               = tau * logf(aux);
    w                   = 1.f / aux;
    chi_in              = 0.f;
    inlier              = 0;
  }
  const float wa = Ja * w, wb = Jb * w, wc = Jc * w, wd0 = d0 * w, wd1 = d1 * w;
  p[0]  = fmaf(wa, Ja, p[0]);
  p[1]  = fmaf(wa, Jb, p[1]);
  p[2]  = fmaf(wa, Jc, p[2]);
  p[3]  = fmaf(wb, Jb, p[3]);
  p[4]  = fmaf(wb, Jc, p[4]);
  p[5]  = fmaf(wd1, d1, fmaf(wd0, d0, fmaf(wc, Jc, p[5])));
  p[6]  = fmaf(wa, e0, p[6]);
  p[7]  = fmaf(wb, e0, p[7]);
  p[8]  = fmaf(wd1, e2, fmaf(wd0, e1, fmaf(wc, e0, p[8])));
  p[9]  = p[9] + chi_in;
  p[10] = p[10] + chi_k;
  return inlier;
}

/* D19 -- SE2Point2PointErrorFactor[WithSensor]: e = p_pred - p_fixed, J = [R | R (-y, x)^T], Omega = I2.
 * returns 1 if inlier, 0 if kernelized */
static inline int contribution_p2p(const orc_params* prm, const factor_ctx* f, orc_point pf, orc_point pm, float* v) {
  float px, py, jc0, jc1;
  apply(f->X, pm.x, pm.y, &px, &py);
  if (f->with_sensor) { /* D9 */
    float qx, qy;
    apply(f->Sinv, px, py, &qx, &qy);
    px = qx;
    py = qy;
  }
  const float e0 = px - pf.x, e1 = py - pf.y;
  const float c = f->RX.c, s = f->RX.s;
  rot(f->RX, -pm.y, pm.x, &jc0, &jc1);
  const float chi = e0 * e0 + e1 * e1;
  float w         = 1.f;
  int inlier      = 1;
  float chi_in = chi, chi_k = 0.f;
  const float tau = prm->cauchy_chi_threshold;
  if (tau > 0.f && !(chi < tau)) { /* D6 */
    const float inv_tau = 1.f / tau;
    const float aux     = chi * inv_tau + 1.f;
    chi_k               = tau * logf(aux);
    w                   = 1.f / aux;
    chi_in              = 0.f;
    inlier              = 0;
  }
  const float wc = c * w, ws = s * w, wms = (-s) * w, wj0 = jc0 * w, wj1 = jc1 * w;
  v[0]  = wc * c + ws * s;
  v[1]  = wc * (-s) + ws * c;
  v[2]  = wc * jc0 + ws * jc1;
  v[3]  = wms * (-s) + wc * c;
  v[4]  = wms * jc0 + wc * jc1;
  v[5]  = wj0 * jc0 + wj1 * jc1;
  v[6]  = wc * e0 + ws * e1;
  v[7]  = wms * e0 + wc * e1;
  v[8]  = wj0 * e0 + wj1 * e1;
  v[9]  = chi_in;
  v[10] = chi_k;
  return inlier;
}

/* robustified squared error of one correspondence at the factor's pose (L4): the error half of contribution() */
static inline float correspondence_chi(const orc_params* prm, const factor_ctx* f, orc_point pf, orc_point pm) {
  float px, py;
  apply(f->X, pm.x, pm.y, &px, &py);
  if (f->with_sensor) {
    float qx, qy;
    apply(f->Sinv, px, py, &qx, &qy);
    px = qx;
    py = qy;
  }
  const float dx = px - pf.x, dy = py - pf.y;
  float chi;
  if (prm->factor == ORC_FACTOR_POINT2POINT) {
    chi = dx * dx + dy * dy;
  } else {
    float nx, ny;
    rot(f->RX, pm.nx, pm.ny, &nx, &ny);
    const float e0 = dx * pf.nx + dy * pf.ny, e1 = nx - pf.nx, e2 = ny - pf.ny;
    chi = (e0 * e0 + e1 * e1) + e2 * e2;
  }
  const float tau = prm->cauchy_chi_threshold;
  if (tau > 0.f && !(chi < tau)) {
    chi = tau * logf(chi * (1.f / tau) + 1.f);
  }
  return chi;
}

typedef struct {
  float v[NSLOT];
  int32_t n_inliers, n_kernelized;
} lin_sums;

/* one correspondence through the slice's factor (D19); inlier_only (I1): a kernelized factor keeps its statistics
 * but adds nothing to H and b */
static inline int contribution_any(const orc_params* prm, const factor_ctx* f, orc_point pf, orc_point pm,
                                   int inlier_only, float* v, int* n_slots) {
  const int inl = prm->factor == ORC_FACTOR_POINT2POINT ? contribution_p2p(prm, f, pf, pm, v)
                                                        : contribution(prm, f, pf, pm, v);
  *n_slots = NSLOT;
  if (inlier_only && !inl) {
    v[0] = v[9], v[1] = v[10]; /* caller adds them to slots 9, 10 only */
    *n_slots = 0;
  }
  return inl;
}

static void linearize(const orc_params* prm, orc_iso X, const orc_point* fixed,
                      const orc_point* moving, int32_t n_moving, const int32_t* fixed_idx,
                      const int32_t* moving_idx, int32_t n_corr, int32_t sum_mode,
                      int32_t tree_threads, int inlier_only, lin_sums* out) {
  const factor_ctx f = make_factor(prm, X);
  memset(out, 0, sizeof(*out));
  if (sum_mode == ORC_SUM_SEQUENTIAL) {
    for (int32_t k = 0; k < n_corr; ++k) {
      float v[NSLOT];
      int ns;
      const int inl = contribution_any(prm, &f, fixed[fixed_idx[k]], moving[moving_idx[k]], inlier_only, v, &ns);
      if (ns) {
        for (int s = 0; s < NSLOT; ++s) {
          out->v[s] = out->v[s] + v[s];
        }
      } else {
        out->v[9] = out->v[9] + v[0], out->v[10] = out->v[10] + v[1];
      }
      out->n_inliers += inl;
      out->n_kernelized += !inl;
    }
    return;
  }
  /* ORC_SUM_TREE: the CUDA kernels' fixed reduction shapes (D10).  tree_threads = threads per pair (bits 0..15);
   * bit 16 selects how a warp's 32 lane partials are combined:
   *   0  xor-butterfly (offsets 16, 8, 4, 2, 1)                      -- icp_fused_kernel / stream / multi
   *   1  lanes 0..15 and 16..31 summed in ascending order, then added -- icp_fused2_kernel (transposed tile)
   * bit 17: the contributions and their accumulation into the thread's partial use the fused arithmetic of D18 */
  const int32_t T        = tree_threads & 0xFFFF;
  const int32_t half_seq = (tree_threads >> 16) & 1;
  const int32_t fused    = (tree_threads >> 17) & 1;
  float* part     = (float*) calloc((size_t) T * NSLOT, sizeof(float));
  int32_t* fx_of  = (int32_t*) malloc(sizeof(int32_t) * (size_t)(n_moving > 0 ? n_moving : 1));
  for (int32_t i = 0; i < n_moving; ++i) {
    fx_of[i] = -1;
  }
  for (int32_t k = 0; k < n_corr; ++k) {
    fx_of[moving_idx[k]] = fixed_idx[k]; /* a moving point wins at most one column */
  }
  for (int32_t i = 0; i < n_moving; ++i) { /* ascending i == ascending slot within a thread */
    if (fx_of[i] < 0) {
      continue;
    }
    float* p = part + (size_t)(i % T) * NSLOT;
    int inl;
    if (fused) { /* D18: the plane-to-plane factor of the compile-time-stride kernels only */
      inl = contribution_fused(prm, &f, fixed[fx_of[i]], moving[i], p);
    } else {
      float v[NSLOT];
      int ns;
      inl = contribution_any(prm, &f, fixed[fx_of[i]], moving[i], inlier_only, v, &ns);
      if (ns) {
        for (int s = 0; s < NSLOT; ++s) {
          p[s] = p[s] + v[s];
        }
      } else {
        p[9] = p[9] + v[0], p[10] = p[10] + v[1];
      }
    }
    out->n_inliers += inl;
    out->n_kernelized += !inl;
  }
  for (int32_t w = 0; w < T / 32; ++w) {
    float* lane = part + (size_t) w * 32 * NSLOT;
    if (half_seq) {
      for (int s = 0; s < NSLOT; ++s) {
        float half[2];
        for (int h = 0; h < 2; ++h) {
          half[h] = lane[(h * 16) * NSLOT + s];
          for (int l = 1; l < 16; ++l) {
            half[h] = half[h] + lane[(h * 16 + l) * NSLOT + s];
          }
        }
        lane[s] = half[0] + half[1];
      }
    } else {
      for (int off = 16; off >= 1; off >>= 1) { /* v[l] = v[l] + v[l ^ off] on all lanes */
        for (int l = 0; l < 32; ++l) {
          if (l & off) {
            continue;
          }
          for (int s = 0; s < NSLOT; ++s) {
            const float a             = lane[l * NSLOT + s];
            const float b             = lane[(l ^ off) * NSLOT + s];
            lane[l * NSLOT + s]         = a + b;
            lane[(l ^ off) * NSLOT + s] = b + a;
          }
        }
      }
    }
    for (int s = 0; s < NSLOT; ++s) { /* sequential over warps */
      out->v[s] = (w == 0) ? lane[s] : out->v[s] + lane[s];
    }
  }
  free(part);
  free(fx_of);
}

/* ---------------------------------------------------------------- A.6 GN step */

/* L4: chi1 of the trial pose over the round's correspondences, summed sequentially in correspondence order or in
 * the general kernel's shape (thread t owns moving points t, t + T, ...; xor-butterfly inside a warp; warps in
 * order) */
static float total_chi(const orc_params* prm, orc_iso X, const orc_point* fixed, const orc_point* moving,
                       int32_t n_moving, const int32_t* fixed_idx, const int32_t* moving_idx, int32_t n_corr,
                       int32_t sum_mode, int32_t tree_threads) {
  const factor_ctx f = make_factor(prm, X);
  if (sum_mode == ORC_SUM_SEQUENTIAL) {
    float t = 0.f;
    for (int32_t k = 0; k < n_corr; ++k) {
      t = t + correspondence_chi(prm, &f, fixed[fixed_idx[k]], moving[moving_idx[k]]);
    }
    return t;
  }
  const int32_t T = tree_threads & 0xFFFF;
  float* part     = (float*) calloc((size_t) T, sizeof(float));
  int32_t* fx_of  = (int32_t*) malloc(sizeof(int32_t) * (size_t)(n_moving > 0 ? n_moving : 1));
  for (int32_t i = 0; i < n_moving; ++i) {
    fx_of[i] = -1;
  }
  for (int32_t k = 0; k < n_corr; ++k) {
    fx_of[moving_idx[k]] = fixed_idx[k];
  }
  for (int32_t i = 0; i < n_moving; ++i) {
    if (fx_of[i] >= 0) {
      part[i % T] = part[i % T] + correspondence_chi(prm, &f, fixed[fx_of[i]], moving[i]);
    }
  }
  float total = 0.f;
  for (int32_t w = 0; w < T / 32; ++w) {
    float* lane = part + (size_t) w * 32;
    for (int off = 16; off >= 1; off >>= 1) {
      for (int l = 0; l < 32; ++l) {
        if (!(l & off)) {
          const float a = lane[l], b = lane[l ^ off];
          lane[l] = a + b, lane[l ^ off] = b + a;
        }
      }
    }
    total = (w == 0) ? lane[0] : total + lane[0];
  }
  free(part);
  free(fx_of);
  return total;
}

/* IterationAlgorithmGN + 3x3 LDL^T (D11), D[j] added to the diagonal (GN: the damping; LM: lambda * D_j).
 * returns 0 on success, -1 if not positive definite */
static int solve3d(const float* v, const double* D, float* dx) {
  const double H00 = (double) v[0] + D[0], H01 = v[1], H02 = v[2];
  const double H11 = (double) v[3] + D[1], H12 = v[4];
  const double H22 = (double) v[5] + D[2];
  const double r0 = -(double) v[6], r1 = -(double) v[7], r2 = -(double) v[8];
  if (!(H00 > 0.0)) { /* d0 */
    return -1;
  }
  const double i0  = 1.0 / H00;
  const double l10 = H01 * i0, l20 = H02 * i0;
  const double d1  = H11 - l10 * H01;
  if (!(d1 > 0.0)) {
    return -1;
  }
  const double i1  = 1.0 / d1;
  const double t21 = H12 - l20 * H01; /* = l21 * d1 */
  const double l21 = t21 * i1;
  const double d2  = (H22 - l20 * H02) - l21 * t21;
  if (!(d2 > 0.0)) {
    return -1;
  }
  const double i2 = 1.0 / d2;
  /* L z = r, D y = z, L^T x = y */
  const double z1 = r1 - l10 * r0;
  const double z2 = (r2 - l20 * r0) - l21 * z1;
  const double y0 = r0 * i0, y1 = z1 * i1, x2 = z2 * i2;
  const double x1 = y1 - l21 * x2;
  const double x0 = (y0 - l10 * x1) - l20 * x2;
  dx[0]           = (float) x0;
  dx[1]           = (float) x1;
  dx[2]           = (float) x2;
  if (!isfinite(dx[0]) || !isfinite(dx[1]) || !isfinite(dx[2])) {
    return -1;
  }
  return 0;
}

static int solve3(const float* v, float damping, float* dx) {
  const double D[3] = {(double) damping, (double) damping, (double) damping};
  return solve3d(v, D, dx);
}

/* Levenberg-Marquardt state of one compute() (L2, L8) */
typedef struct {
  double lambda;
  int started;
  int32_t rejected;
} lm_state;

/* one LM round (L3..L7) on the quadratic form `sums` built at X; returns the new estimate */
static orc_iso lm_round(const orc_params* prm, lm_state* lm, const lin_sums* sums, orc_iso X, const orc_point* fixed,
                        const orc_point* moving, int32_t n_moving, const int32_t* fidx, const int32_t* midx,
                        int32_t n_corr, int32_t sum_mode, int32_t tree_threads) {
  const float* v       = sums->v;
  const double diag[3] = {(double) v[0], (double) v[3], (double) v[5]};
  if (!lm->started) { /* L2 */
    double mx = diag[0] > diag[1] ? diag[0] : diag[1];
    mx        = mx > diag[2] ? mx : diag[2];
    lm->lambda  = prm->lm_user_lambda_init > 0.f ? (double) prm->lm_user_lambda_init : (double) prm->lm_tau * mx;
    lm->started = 1;
  }
  const float chi0 = v[9] + v[10];
  double nu        = 2.0;
  for (int32_t t = 0; t < prm->lm_iterations_max; ++t) {
    double D[3];
    for (int j = 0; j < 3; ++j) {
      D[j] = prm->lm_variable_damping ? lm->lambda * diag[j] : lm->lambda;
    }
    float dx[3];
    if (solve3d(v, D, dx) == 0) {
      const orc_iso Xt = orc_compose(X, orc_v2t(dx[0], dx[1], dx[2]));
      const float chi1 = total_chi(prm, Xt, fixed, moving, n_moving, fidx, midx, n_corr, sum_mode, tree_threads);
      double scale     = 0.0;
      for (int j = 0; j < 3; ++j) { /* L5 */
        scale = scale + (double) dx[j] * (D[j] * (double) dx[j] - (double) v[6 + j]);
      }
      scale            = scale + 1e-3;
      const double rho = ((double) chi0 - (double) chi1) / scale;
      if (rho > 0.0 && isfinite(chi1)) { /* L6 */
        const double q = 2.0 * rho - 1.0;
        double alpha   = 1.0 - (q * q) * q;
        alpha          = alpha < (double) prm->lm_step_high ? alpha : (double) prm->lm_step_high;
        const double g = alpha > (double) prm->lm_step_low ? alpha : (double) prm->lm_step_low;
        lm->lambda     = lm->lambda * g;
        return Xt;
      }
    }
    lm->lambda = lm->lambda * nu;
    nu         = nu * 2.0;
    lm->rejected++;
  }
  return X; /* L7 */
}

/* ---------------------------------------------------------------- A.7 outer loop */

static void fill_iter(orc_iter_stats* st, orc_iso X, const lin_sums* sums, int32_t n_corr) {
  float xyt[3];
  orc_t2v(X, xyt);
  st->x              = xyt[0];
  st->y              = xyt[1];
  st->theta          = xyt[2];
  st->c              = X.c;
  st->s              = X.s;
  st->chi_inliers    = sums->v[9];
  st->chi_kernelized = sums->v[10];
  st->n_inliers      = sums->n_inliers;
  st->n_kernelized   = sums->n_kernelized;
  st->n_corr         = n_corr;
}

void orc_align_iso(const orc_params* prm, const orc_point* fixed, int32_t n_fixed, const orc_point* moving,
                   int32_t n_moving, orc_iso init, int32_t sum_mode, int32_t tree_threads, orc_result* out,
                   orc_iter_stats* iter_stats) {
  const int32_t C      = prm->canvas_cols;
  orc_cell* fixed_img  = (orc_cell*) malloc(sizeof(orc_cell) * (size_t) C);
  orc_cell* moving_img = (orc_cell*) malloc(sizeof(orc_cell) * (size_t) C);
  int32_t* fidx        = (int32_t*) malloc(sizeof(int32_t) * (size_t) C);
  int32_t* midx        = (int32_t*) malloc(sizeof(int32_t) * (size_t) C);

  /* correspondence_finder_projective_2d.cpp:37-44: fixed projected once, identity camera */
  orc_project(prm, orc_v2t(0.f, 0.f, 0.f), fixed, n_fixed, fixed_img);

  orc_iso X    = init; /* setMovingInFixed */
  orc_iso Sinv = orc_v2t(0.f, 0.f, 0.f);
  if (prm->with_sensor) {
    Sinv = orc_inverse(sensor_iso(prm));
  }
  const int32_t n_phases = prm->enable_inlier_only_runs ? 2 : 1;
  memset(out, 0, sizeof(*out));
  if (iter_stats) {
    memset(iter_stats, 0, sizeof(orc_iter_stats) * (size_t) prm->max_iterations * (size_t) n_phases);
  }
  lin_sums sums;
  memset(&sums, 0, sizeof(sums));
  lm_state lm;
  memset(&lm, 0, sizeof(lm));
  int32_t n_corr = 0;
  int32_t status = -1;
  int32_t it     = 0; /* rounds executed over both phases */
  for (int32_t phase = 0; phase < n_phases && status < 0; ++phase) {
    if (phase == 1 && sums.n_inliers < prm->min_num_inliers) { /* I1: "if sufficient inliers are available" */
      break;
    }
    float chi_prev = 0.f;
    for (int32_t k = 0; k < prm->max_iterations; ++k) {
      /* slice->findCorrespondences(): local_map_in_sensor = sensor_in_robot^-1 * moving_in_fixed */
      const orc_iso L = prm->with_sensor ? orc_compose(Sinv, X) : X;
      n_corr = orc_find_correspondences(prm, fixed_img, moving, n_moving, L, moving_img, fidx, midx);
      memset(&sums, 0, sizeof(sums));
      if (n_corr <= prm->min_num_correspondences) {
        status = ORC_STATUS_NOT_ENOUGH_CORRESPONDENCES;
        break;
      }
      linearize(prm, X, fixed, moving, n_moving, fidx, midx, n_corr, sum_mode, tree_threads, phase == 1, &sums);
      if (prm->algorithm == ORC_ALGORITHM_LM) {
        X = lm_round(prm, &lm, &sums, X, fixed, moving, n_moving, fidx, midx, n_corr, sum_mode, tree_threads);
      } else {
        float dx[3];
        if (solve3(sums.v, prm->damping, dx) != 0) {
          status = ORC_STATUS_SINGULAR;
          break;
        }
        X = orc_compose(X, orc_v2t(dx[0], dx[1], dx[2])); /* VariableSE2Right: X <- X * v2t(dx) */
      }
      if (iter_stats) {
        fill_iter(&iter_stats[it], X, &sums, n_corr);
      }
      ++it;
      const float chi = sums.v[9] + sums.v[10];
      if (prm->termination_epsilon > 0.f && k > 0 && chi_prev - chi < prm->termination_epsilon * chi_prev) { /* T1 */
        break;
      }
      chi_prev = chi;
    }
  }
  if (status < 0) {
    status = (sums.n_inliers < prm->min_num_inliers) ? ORC_STATUS_NOT_ENOUGH_INLIERS
                                                     : ORC_STATUS_SUCCESS;
  }
  float xyt[3];
  orc_t2v(X, xyt);
  out->x              = xyt[0];
  out->y              = xyt[1];
  out->theta          = xyt[2];
  out->c              = X.c;
  out->s              = X.s;
  out->chi_inliers    = sums.v[9];
  out->chi_kernelized = sums.v[10];
  out->n_inliers      = sums.n_inliers;
  out->n_kernelized   = sums.n_kernelized;
  out->n_corr         = n_corr;
  out->status         = status;
  out->iterations     = it;
  out->lm_rejected    = lm.rejected;
  for (int s = 0; s < 6; ++s) {
    out->H[s] = sums.v[s];
  }
  free(fixed_img);
  free(moving_img);
  free(fidx);
  free(midx);
}

void orc_align(const orc_params* prm, const orc_point* fixed, int32_t n_fixed,
               const orc_point* moving, int32_t n_moving, const float* init_xyt,
               int32_t sum_mode, int32_t tree_threads, orc_result* out,
               orc_iter_stats* iter_stats) {
  orc_align_iso(prm, fixed, n_fixed, moving, n_moving, orc_v2t(init_xyt[0], init_xyt[1], init_xyt[2]), sum_mode,
                tree_threads, out, iter_stats);
}

void orc_align_batch(const orc_params* prm, const orc_point* fixed_pts, const int32_t* fixed_off,
                     const orc_point* moving_pts, const int32_t* moving_off,
                     const int32_t* fixed_id, const int32_t* moving_id, const float* init_pose,
                     int32_t pose_stride, int32_t n_pairs, int32_t sum_mode, int32_t tree_threads,
                     int32_t n_threads, orc_result* out, orc_iter_stats* iter_stats) {
  const size_t n_iter = (size_t) prm->max_iterations * (prm->enable_inlier_only_runs ? 2 : 1);
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 8) num_threads(n_threads > 1 ? n_threads : 1)
#endif
  for (int32_t p = 0; p < n_pairs; ++p) {
    const int32_t f = fixed_id ? fixed_id[p] : p;
    const int32_t m = moving_id ? moving_id[p] : p;
    orc_align_iso(prm, fixed_pts + fixed_off[f], fixed_off[f + 1] - fixed_off[f], moving_pts + moving_off[m],
                  moving_off[m + 1] - moving_off[m], pose_at(init_pose, (size_t) p, pose_stride), sum_mode,
                  tree_threads, out + p, iter_stats ? iter_stats + (size_t) p * n_iter : NULL);
  }
}


/* ---------------------------------------------------------------- multi-slice aligner (D14-D17) */

/* SE2PriorErrorFactor [srrg2_solver types_2d/se2_prior_error_factor; slice class named at
 * L0.json:291-310, MULTI.json:400-422] -- D15 */
static void prior_eval(const orc_prior* pr, orc_iso X, float* e, orc_iso* P_out) {
  const orc_iso Zinv = orc_inverse(pose_at(pr->z, 0, pr->z_is_iso ? 4 : 3));
  const orc_iso P    = orc_compose(Zinv, X);
  e[0]               = P.tx;
  e[1]               = P.ty;
  e[2]               = atan2f(P.s, P.c);
  *P_out             = P;
}

void orc_prior_error_and_jacobian(const orc_prior* prior, orc_iso X, float* e, float* J) {
  orc_iso P;
  prior_eval(prior, X, e, &P);
  J[0] = P.c, J[1] = -P.s, J[2] = 0.f;
  J[3] = P.s, J[4] = P.c, J[5] = 0.f;
  J[6] = 0.f, J[7] = 0.f, J[8] = 1.f;
}

/* the prior's H/b/chi contribution; returns 1 if inlier, 0 if kernelized */
static int prior_contribution(const orc_prior* pr, orc_iso X, float* v) {
  float e[3];
  orc_iso P;
  prior_eval(pr, X, e, &P);
  const float* O = pr->information; /* O00 O01 O02 O11 O12 O22 */
  const float Om[3][3] = {{O[0], O[1], O[2]}, {O[1], O[3], O[4]}, {O[2], O[4], O[5]}};
  float Oe[3];
  for (int i = 0; i < 3; ++i) {
    Oe[i] = (Om[i][0] * e[0] + Om[i][1] * e[1]) + Om[i][2] * e[2];
  }
  const float chi = (e[0] * Oe[0] + e[1] * Oe[1]) + e[2] * Oe[2];
  float w = 1.f, chi_in = chi, chi_k = 0.f;
  int inlier      = 1;
  const float tau = pr->cauchy_chi_threshold;
  if (tau > 0.f && !(chi < tau)) { /* D6 */
    const float aux = chi * (1.f / tau) + 1.f;
    chi_k           = tau * logf(aux);
    w               = 1.f / aux;
    chi_in          = 0.f;
    inlier          = 0;
  }
  /* A = J^T Omega (zeros of J skipped), scaled by w */
  float A[3][3];
  for (int j = 0; j < 3; ++j) {
    A[0][j] = (P.c * Om[0][j] + P.s * Om[1][j]) * w;
    A[1][j] = ((-P.s) * Om[0][j] + P.c * Om[1][j]) * w;
    A[2][j] = Om[2][j] * w;
  }
  float H[3][3], b[3];
  for (int i = 0; i < 3; ++i) {
    H[i][0] = A[i][0] * P.c + A[i][1] * P.s;
    H[i][1] = A[i][0] * (-P.s) + A[i][1] * P.c;
    H[i][2] = A[i][2];
    b[i]    = (A[i][0] * e[0] + A[i][1] * e[1]) + A[i][2] * e[2];
  }
  v[0] = H[0][0], v[1] = H[0][1], v[2] = H[0][2], v[3] = H[1][1], v[4] = H[1][2], v[5] = H[2][2];
  v[6] = b[0], v[7] = b[1], v[8] = b[2];
  v[9]  = chi_in;
  v[10] = chi_k;
  return inlier;
}

void orc_align_multi(const orc_params* slices, int32_t n_slices, const orc_point* const* fixed,
                     const int32_t* n_fixed, const orc_point* const* moving, const int32_t* n_moving,
                     const orc_prior* prior, orc_iso init, int32_t sum_mode, int32_t tree_threads,
                     orc_result* out, orc_iter_stats* iter_stats) {
  const int32_t max_it = slices[0].max_iterations; /* D17 */
  orc_cell* fixed_img[ORC_MAX_SLICES];
  orc_iso Sinv[ORC_MAX_SLICES];
  int32_t Cmax = 1;
  for (int32_t s = 0; s < n_slices; ++s) {
    const int32_t C = slices[s].canvas_cols;
    Cmax            = C > Cmax ? C : Cmax;
    fixed_img[s]    = (orc_cell*) malloc(sizeof(orc_cell) * (size_t) C);
    orc_project(&slices[s], orc_v2t(0.f, 0.f, 0.f), fixed[s], n_fixed[s], fixed_img[s]);
    Sinv[s] = orc_v2t(0.f, 0.f, 0.f);
    if (slices[s].with_sensor) {
      Sinv[s] = orc_inverse(sensor_iso(&slices[s]));
    }
  }
  orc_cell* moving_img = (orc_cell*) malloc(sizeof(orc_cell) * (size_t) Cmax);
  int32_t* fidx        = (int32_t*) malloc(sizeof(int32_t) * (size_t) Cmax);
  int32_t* midx        = (int32_t*) malloc(sizeof(int32_t) * (size_t) Cmax);
  orc_iso X            = init;
  memset(out, 0, sizeof(*out));
  if (iter_stats) {
    memset(iter_stats, 0, sizeof(orc_iter_stats) * (size_t) max_it);
  }
  lin_sums tot;
  memset(&tot, 0, sizeof(tot));
  int32_t n_corr = 0, status = -1, it = 0;
  for (; it < max_it; ++it) {
    memset(&tot, 0, sizeof(tot));
    n_corr          = 0;
    int contributed = 0;
    int32_t n_found = 0;
    for (int32_t s = 0; s < n_slices; ++s) { /* D14 */
      const orc_iso L  = slices[s].with_sensor ? orc_compose(Sinv[s], X) : X;
      const int32_t nc = orc_find_correspondences(&slices[s], fixed_img[s], moving[s], n_moving[s], L, moving_img,
                                                  fidx, midx);
      n_found += nc;
      if (nc <= slices[s].min_num_correspondences) {
        continue;
      }
      lin_sums ss;
      linearize(&slices[s], X, fixed[s], moving[s], n_moving[s], fidx, midx, nc, sum_mode, tree_threads, 0, &ss);
      for (int k = 0; k < NSLOT; ++k) { /* D16 */
        tot.v[k] = contributed ? tot.v[k] + ss.v[k] : ss.v[k];
      }
      tot.n_inliers += ss.n_inliers;
      tot.n_kernelized += ss.n_kernelized;
      n_corr += nc;
      contributed = 1;
    }
    if (!contributed) {
      memset(&tot, 0, sizeof(tot));
      n_corr = n_found; /* what the finders produced, as orc_align reports it */
      status = ORC_STATUS_NOT_ENOUGH_CORRESPONDENCES;
      break;
    }
    if (prior) {
      float pv[NSLOT];
      const int inl = prior_contribution(prior, X, pv);
      for (int k = 0; k < NSLOT; ++k) {
        tot.v[k] = tot.v[k] + pv[k];
      }
      tot.n_inliers += inl;
      tot.n_kernelized += !inl;
    }
    float dx[3];
    if (solve3(tot.v, slices[0].damping, dx) != 0) {
      status = ORC_STATUS_SINGULAR;
      break;
    }
    X = orc_compose(X, orc_v2t(dx[0], dx[1], dx[2]));
    if (iter_stats) {
      fill_iter(&iter_stats[it], X, &tot, n_corr);
    }
  }
  if (status < 0) {
    status = (tot.n_inliers < slices[0].min_num_inliers) ? ORC_STATUS_NOT_ENOUGH_INLIERS : ORC_STATUS_SUCCESS;
  }
  float xyt[3];
  orc_t2v(X, xyt);
  out->x = xyt[0], out->y = xyt[1], out->theta = xyt[2];
  out->c = X.c, out->s = X.s;
  out->chi_inliers    = tot.v[9];
  out->chi_kernelized = tot.v[10];
  out->n_inliers      = tot.n_inliers;
  out->n_kernelized   = tot.n_kernelized;
  out->n_corr         = n_corr;
  out->status         = status;
  out->iterations     = it;
  for (int k = 0; k < 6; ++k) {
    out->H[k] = tot.v[k];
  }
  for (int32_t s = 0; s < n_slices; ++s) {
    free(fixed_img[s]);
  }
  free(moving_img);
  free(fidx);
  free(midx);
}

void orc_align_multi_batch(const orc_params* slices, int32_t n_slices, const orc_point* const* fixed_pts,
                           const int32_t* const* fixed_off, const orc_point* const* moving_pts,
                           const int32_t* const* moving_off, const int32_t* fixed_id, const int32_t* moving_id,
                           const orc_prior* prior, const float* prior_z, const float* init_pose, int32_t pose_stride,
                           int32_t n_pairs, int32_t sum_mode, int32_t tree_threads, int32_t n_threads,
                           orc_result* out, orc_iter_stats* iter_stats) {
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 8) num_threads(n_threads > 1 ? n_threads : 1)
#endif
  for (int32_t p = 0; p < n_pairs; ++p) {
    const int32_t f = fixed_id ? fixed_id[p] : p;
    const int32_t m = moving_id ? moving_id[p] : p;
    const orc_point* fx[ORC_MAX_SLICES];
    const orc_point* mv[ORC_MAX_SLICES];
    int32_t nf[ORC_MAX_SLICES], nm[ORC_MAX_SLICES];
    for (int32_t s = 0; s < n_slices; ++s) {
      fx[s] = fixed_pts[s] + fixed_off[s][f];
      nf[s] = fixed_off[s][f + 1] - fixed_off[s][f];
      mv[s] = moving_pts[s] + moving_off[s][m];
      nm[s] = moving_off[s][m + 1] - moving_off[s][m];
    }
    orc_prior pr;
    if (prior && prior_z) {
      pr = *prior;
      memcpy(pr.z, prior_z + (size_t) pose_stride * (size_t) p, sizeof(float) * (size_t) pose_stride);
      pr.z_is_iso = pose_stride == 4;
    }
    orc_align_multi(slices, n_slices, fx, nf, mv, nm, (prior && prior_z) ? &pr : NULL,
                    pose_at(init_pose, (size_t) p, pose_stride), sum_mode, tree_threads, out + p,
                    iter_stats ? iter_stats + (size_t) p * slices[0].max_iterations : NULL);
  }
}

/* ---------------------------------------------------------------- A.8 verification gates */

int32_t orc_accept(const orc_result* r, int32_t min_inliers, float max_chi_per_inlier,
                   float min_inlier_ratio) {
  if (r->status != ORC_STATUS_SUCCESS) {
    return 0;
  }
  if (r->n_inliers < min_inliers) { /* relocalize_min_inliers  L0.json:627-634 */
    return 0;
  }
  if (r->n_inliers <= 0 || r->n_corr <= 0) {
    return 0;
  }
  if (r->chi_inliers / (float) r->n_inliers > max_chi_per_inlier) {
    return 0;
  }
  if ((float) r->n_inliers / (float) r->n_corr < min_inlier_ratio) {
    return 0;
  }
  return 1;
}

int32_t orc_best_of(const orc_result* r, int32_t n, int32_t min_inliers, float max_chi_per_inlier,
                    float min_inlier_ratio) {
  int32_t best = -1;
  for (int32_t i = 0; i < n; ++i) {
    if (!orc_accept(&r[i], min_inliers, max_chi_per_inlier, min_inlier_ratio)) {
      continue;
    }
    if (best < 0) {
      best = i;
      continue;
    }
    const float ci = r[i].chi_inliers / (float) r[i].n_inliers;
    const float cb = r[best].chi_inliers / (float) r[best].n_inliers;
    if (r[i].n_inliers > r[best].n_inliers ||
        (r[i].n_inliers == r[best].n_inliers && ci < cb)) {
      best = i; /* ties keep the lowest id */
    }
  }
  return best;
}

int32_t orc_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* ---------------------------------------------------------------- host libm bulk drivers
 * (tests/test_math_host.py compares the kernels' own libm copies against these; numpy's float32
 * ufuncs may dispatch to SIMD implementations that are NOT the libm the reference links) */
void orc_libm_atan2f_n(const float* y, const float* x, float* out, long n) {
  for (long i = 0; i < n; ++i) {
    out[i] = atan2f(y[i], x[i]);
  }
}

void orc_libm_logf_n(const float* x, float* out, long n) {
  for (long i = 0; i < n; ++i) {
    out[i] = logf(x[i]);
  }
}

void orc_libm_sincosf_n(const float* x, float* s, float* c, long n) {
  for (long i = 0; i < n; ++i) {
    s[i] = sinf(x[i]);
    c[i] = cosf(x[i]);
  }
}

/* lrintf(K00 * atan2f(y, x) + K01) with the projector's gating: -1 outside [0, C) */
void orc_column_n(const orc_params* prm, const float* y, const float* x, int32_t* col, long n) {
  const int32_t C = prm->canvas_cols;
  const float K00 = (float) C / (prm->angle_col_max - prm->angle_col_min);
  const float K01 = (float) C * 0.5f;
  for (long i = 0; i < n; ++i) {
    const float u = K00 * atan2f(y[i], x[i]) + K01;
    const long c  = lrintf(u);
    col[i]        = (c < 0 || c >= C) ? -1 : (int32_t) c;
  }
}

/* ---------------------------------------------------------------- next rows
 *
 * D14 PointNormal2f arithmetic in the merger (merger_projective_2d.cpp:72-74): "+=" and "*= 0.5f" act on all
 *     four fields, normalize() renormalises the normal only, with Eigen's rule (divide by sqrt(squaredNorm)
 *     when squaredNorm > 0). */

/* voxelize_cloud is defined with the pre-processor below */
static int32_t voxelize_cloud(const orc_point* pts, const uint8_t* valid, int32_t n, const float* inv, orc_point* out);

static int iso_is_identity(orc_iso T) {
  return T.c == 1.f && T.s == 0.f && T.tx == 0.f && T.ty == 0.f;
}

/* SceneClipperProjective2D::compute, R/mapping/scene_clipper_projective_2d.cpp:22-62 (both shipped configurations
 * set voxelize_resolution 0).  out holds canvas_cols points; returns the count. */
int32_t orc_clip_scene_voxelized(const orc_params* prm, const orc_point* scene, int32_t n_scene,
                                 orc_iso robot_in_local_map, orc_iso sensor_in_robot, float voxelize_resolution,
                                 orc_point* out) {
  const int32_t C = prm->canvas_cols;
  orc_cell* img   = (orc_cell*) malloc(sizeof(orc_cell) * (size_t) C);
  const orc_iso sensor_in_local_map = orc_compose(robot_in_local_map, sensor_in_robot); /* :29 */
  orc_project(prm, sensor_in_local_map, scene, n_scene, img);                             /* :31-32 */
  int32_t k = 0;
  for (int32_t c = 0; c < C; ++c) { /* :38-43 / :50-56 */
    if (img[c].source_idx < 0) {
      continue;
    }
    out[k].x  = img[c].px;
    out[k].y  = img[c].py;
    out[k].nx = img[c].nx;
    out[k].ny = img[c].ny;
    ++k;
  }
  if (voxelize_resolution > 0.f) { /* :36-48: voxelize(res_coeffs = (res, res, 0.1, 0.1)) of the points in the sensor */
    orc_point* tmp = (orc_point*) malloc(sizeof(orc_point) * (size_t)(k > 0 ? k : 1));
    memcpy(tmp, out, sizeof(orc_point) * (size_t) k);
    const float inv[4] = {1.f / voxelize_resolution, 1.f / voxelize_resolution, 1.f / 0.1f, 1.f / 0.1f};
    k = voxelize_cloud(tmp, NULL, k, inv, out);
    free(tmp);
  }
  if (!iso_is_identity(sensor_in_robot)) { /* :60-62 move the local scene in robot's coords */
    for (int32_t i = 0; i < k; ++i) {
      float x, y, nx, ny;
      apply(sensor_in_robot, out[i].x, out[i].y, &x, &y);
      rot(sensor_in_robot, out[i].nx, out[i].ny, &nx, &ny);
      out[i].x = x, out[i].y = y, out[i].nx = nx, out[i].ny = ny;
    }
  }
  free(img);
  return k;
}

int32_t orc_clip_scene(const orc_params* prm, const orc_point* scene, int32_t n_scene,
                       orc_iso robot_in_local_map, orc_iso sensor_in_robot, orc_point* out) {
  return orc_clip_scene_voxelized(prm, scene, n_scene, robot_in_local_map, sensor_in_robot, 0.f, out);
}

/* MergerProjective2D::compute, R/mapping/merger_projective_2d.cpp:9-100.  `scene` must have room for
 * n_scene + canvas_cols points.  counters = {new, merged, replaced}.  Returns the new scene size. */
int32_t orc_merge(const orc_params* prm, float merge_threshold, orc_point* scene, int32_t n_scene,
                  const orc_point* measurement, int32_t n_measurement, orc_iso measurement_in_scene,
                  int32_t* counters) {
  const int32_t C  = prm->canvas_cols;
  orc_cell* simg   = (orc_cell*) malloc(sizeof(orc_cell) * (size_t) C);
  orc_cell* mimg   = (orc_cell*) malloc(sizeof(orc_cell) * (size_t) C);
  orc_point* tmeas = (orc_point*) malloc(sizeof(orc_point) * (size_t)(n_measurement > 0 ? n_measurement : 1));
  orc_project(prm, measurement_in_scene, scene, n_scene, simg); /* :19-20 */
  for (int32_t i = 0; i < n_measurement; ++i) {                 /* :22-23 transformInPlace */
    apply(measurement_in_scene, measurement[i].x, measurement[i].y, &tmeas[i].x, &tmeas[i].y);
    rot(measurement_in_scene, measurement[i].nx, measurement[i].ny, &tmeas[i].nx, &tmeas[i].ny);
  }
  orc_project(prm, measurement_in_scene, tmeas, n_measurement, mimg); /* :24-25 */
  int32_t scene_size = n_scene;
  int32_t n_new = 0, n_merged = 0, n_replaced = 0;
  for (int32_t c = 0; c < C; ++c) { /* :39-89 */
    orc_cell* s = &simg[c];
    orc_cell* m = &mimg[c];
    if (m->depth > .9f * prm->range_max) { /* :46 */
      m->source_idx = -1;
    }
    if (m->source_idx < 0) { /* :51 */
      continue;
    }
    const orc_point mp = tmeas[m->source_idx];
    if (s->source_idx < 0) { /* :57 */
      scene[scene_size++] = mp;
      ++n_new;
      continue;
    }
    orc_point* sp      = &scene[s->source_idx];
    const float dr     = m->depth - s->depth; /* :66 */
    const float abs_dr = fabsf(dr);
    if (abs_dr < merge_threshold) { /* :71-76, D14 */
      float x = sp->x + mp.x, y = sp->y + mp.y, nx = sp->nx + mp.nx, ny = sp->ny + mp.ny;
      x *= 0.5f, y *= 0.5f, nx *= 0.5f, ny *= 0.5f;
      const float z = nx * nx + ny * ny;
      if (z > 0.f) {
        const float nrm = sqrtf(z);
        nx = nx / nrm;
        ny = ny / nrm;
      }
      sp->x = x, sp->y = y, sp->nx = nx, sp->ny = ny;
      ++n_merged;
      continue;
    }
    if (dr > 0) { /* :80-84 measure is behind: replace */
      *sp = mp;
      ++n_replaced;
      continue;
    }
    scene[scene_size++] = mp; /* :87-88 */
  }
  if (counters) {
    counters[0] = n_new, counters[1] = n_merged, counters[2] = n_replaced;
  }
  free(simg);
  free(mimg);
  free(tmeas);
  return scene_size;
}

/* ---------------------------------------------------------------- next row
 * RawDataPreprocessorProjective2D: LaserMessage ranges -> PointNormal2fVectorCloud
 * (R/sensor_processing/raw_data_preprocessor_projective_2d.cpp:13-51, 77-104).  The in-repo part (range limits,
 * sensor matrix, voxelize / valid-only branch) is followed line by line; the unprojector, the sliding-window normal
 * computator and PointCloud::voxelize live in srrg2_core (un-vendored) and are restated with decision points:
 *  P1 PointNormal2fUnprojectorPolar reads only fx = K(0,0) and cx = K(0,1) of the sensor matrix (its second row is
 *     zero, .cpp:89-90, so no matrix inverse can be involved): azimuth = (1/fx) * (c - cx); a beam is skipped when
 *     range < range_min || range > range_max; point = (range * cosf(azimuth), range * sinf(azimuth)), normal 0.
 *     The cloud handed to the normal computator holds the accepted beams only, in beam order (.cpp:29-31 uses the
 *     unorganised back-inserter overload).
 *  P2 NormalComputator1DSlidingWindow: the window of point i is the maximal run of consecutive points j around i
 *     with |p_j - p_i|^2 < normal_point_distance^2 (scan outwards from i, stop at the first violation).
 *  P3 fewer than normal_min_points points in the window (i included) => the point is Invalid (dropped by both
 *     branches of .cpp:37-48).
 *  P4 covariance from the first and second moments of d_j = p_j - p_i (relative to the query point, which keeps
 *     binary32 accurate far from the origin): m = sum d / n, C = sum d d^T / n - m m^T; the sums accumulate in the
 *     order the window walk visits the points (i-1 down to the window's start, then i+1 up to its end), so one
 *     pass finds the window and the moments -- the "sliding window" accumulates as it slides.
 *  P5 normal = eigenvector of the smallest eigenvalue by Eigen's closed-form 2x2
 *     SelfAdjointEigenSolver::computeDirect (shift by trace/2, scale by max|coeff|, roots t1 -/+ t0, eigenvector of
 *     the larger root from the better-conditioned row, the other by unitOrthogonal()); degenerate (equal roots):
 *     identity => normal (1, 0).
 *  P6 the normal is flipped to face the sensor: n <- -n when n . p > 0.
 *  P7 voxelize(res_coeffs = (res, res, 1, 1)) (.cpp:40-42): key = trunc-toward-zero of the plain vector
 *     (x, y, nx, ny) * (1/res, 1/res, 1, 1) as int; entries sorted by key (lexicographic), equal keys keep cloud
 *     order; each run of equal keys emits ONE point: the sequential sum of the run times (1 / count), normal
 *     renormalised with Eigen's rule; only Valid points take part; output in sorted order.
 *  P8 voxelize_resolution <= 0: the Valid points in cloud order (.cpp:44-48).
 */

typedef struct {
  int32_t k[4];
  int32_t idx;
} vox_entry;

static int vox_cmp(const void* a, const void* b) {
  const vox_entry* x = (const vox_entry*) a;
  const vox_entry* y = (const vox_entry*) b;
  for (int d = 0; d < 4; ++d) {
    if (x->k[d] != y->k[d]) {
      return x->k[d] < y->k[d] ? -1 : 1;
    }
  }
  return x->idx < y->idx ? -1 : (x->idx > y->idx ? 1 : 0);
}

/* PointCloud::voxelize(out, res_coeffs) (P7): valid points only; `valid` may be NULL (all valid).  inv[4] = the
 * inverse scales of the plain vector (x, y, nx, ny).  Returns the number of points written to out. */
static int32_t voxelize_cloud(const orc_point* pts, const uint8_t* valid, int32_t n, const float* inv, orc_point* out) {
  vox_entry* e = (vox_entry*) malloc(sizeof(vox_entry) * (size_t)(n > 0 ? n : 1));
  int32_t m = 0, k = 0;
  for (int32_t i = 0; i < n; ++i) {
    if (valid && !valid[i]) {
      continue;
    }
    e[m].k[0] = (int32_t)(pts[i].x * inv[0]);
    e[m].k[1] = (int32_t)(pts[i].y * inv[1]);
    e[m].k[2] = (int32_t)(pts[i].nx * inv[2]);
    e[m].k[3] = (int32_t)(pts[i].ny * inv[3]);
    e[m].idx  = i;
    ++m;
  }
  qsort(e, (size_t) m, sizeof(vox_entry), vox_cmp);
  int32_t s = 0;
  while (s < m) {
    int32_t t = s;
    float ax = 0.f, ay = 0.f, anx = 0.f, any = 0.f;
    while (t < m && e[t].k[0] == e[s].k[0] && e[t].k[1] == e[s].k[1] && e[t].k[2] == e[s].k[2] &&
           e[t].k[3] == e[s].k[3]) {
      const orc_point* p = &pts[e[t].idx];
      ax = ax + p->x, ay = ay + p->y, anx = anx + p->nx, any = any + p->ny;
      ++t;
    }
    const float w = 1.f / (float) (t - s);
    ax = ax * w, ay = ay * w, anx = anx * w, any = any * w;
    const float z = anx * anx + any * any;
    if (z > 0.f) {
      const float nrm = sqrtf(z);
      anx = anx / nrm, any = any / nrm;
    }
    out[k].x = ax, out[k].y = ay, out[k].nx = anx, out[k].ny = any;
    ++k;
    s = t;
  }
  free(e);
  return k;
}

void orc_default_scan_params(orc_scan_params* p) {
  p->angle_min             = -2.34747f; /* L0.json:411 (the message carries the sensor's own values) */
  p->angle_max             = 2.35619f;  /* L0.json:408 */
  p->msg_range_min         = 0.f;
  p->msg_range_max         = 30.f;      /* L0.json:429 */
  p->range_min             = 0.f;       /* raw_data_preprocessor_projective_2d.h:39 */
  p->range_max             = 1000.f;    /* raw_data_preprocessor_projective_2d.h:40 */
  p->voxelize_resolution   = 0.02f;     /* raw_data_preprocessor_projective_2d.h:41-45 */
  p->normal_point_distance = 0.3f;      /* L0.json:718 */
  p->normal_min_points     = 5;         /* L0.json:715 */
}

/* Eigen 3.3 SelfAdjointEigenSolver<Matrix2f>::computeDirect, eigenvector of the smallest eigenvalue (P5) */
static void smallest_eigenvector_2x2(float m00, float m10, float m11, float* vx, float* vy) {
  const float shift = (m00 + m11) / 2.f;
  float a = m00 - shift, b = m10, c = m11 - shift;
  float scale = fabsf(a);
  if (fabsf(b) > scale) {
    scale = fabsf(b);
  }
  if (fabsf(c) > scale) {
    scale = fabsf(c);
  }
  if (scale > 0.f) {
    a = a / scale, b = b / scale, c = c / scale;
  }
  const float d  = a - c;
  const float t0 = 0.5f * sqrtf(d * d + 4.f * (b * b));
  const float t1 = 0.5f * (a + c);
  const float r0 = t1 - t0, r1 = t1 + t0;
  if ((r1 - r0) <= fabsf(r1) * FLT_EPSILON) {
    *vx = 1.f, *vy = 0.f; /* eivecs.setIdentity(): column 0 */
    return;
  }
  const float a1 = a - r1, c1 = c - r1;
  const float a2 = a1 * a1, c2 = c1 * c1, b2 = b * b;
  float ux, uy; /* eigenvector of the larger root */
  if (a2 > c2) {
    const float n = sqrtf(a2 + b2);
    ux = -b / n, uy = a1 / n;
  } else {
    const float n = sqrtf(c2 + b2);
    ux = -c1 / n, uy = b / n;
  }
  /* unitOrthogonal(): (-y, x).normalized() */
  const float ox = -uy, oy = ux;
  const float z  = ox * ox + oy * oy;
  if (z > 0.f) {
    const float n = sqrtf(z);
    *vx = ox / n, *vy = oy / n;
  } else {
    *vx = ox, *vy = oy;
  }
}

/* the three upstream stages of the pre-processor, separately callable so that the reference's own in-repo source
 * (compiled against oracle/ref_shim/) can run on top of the very same restatement (tests/test_oracle_vs_reference_sources.py) */

/* PointNormal2fUnprojectorPolar::compute (P1): fx = K(0,0), cx = K(0,1) of the sensor matrix; returns the number of
 * accepted beams, written to pts in beam order with zero normals */
int32_t orc_unproject(float range_min, float range_max, float fx, float cx, const float* ranges, int32_t n_beams,
                      orc_point* pts) {
  const float ifx = 1.f / fx; /* P1 */
  int32_t n = 0;
  for (int32_t c = 0; c < n_beams; ++c) {
    const float r = ranges[c];
    if (r < range_min || r > range_max) {
      continue;
    }
    const float az = ifx * ((float) c - cx);
    pts[n].x  = r * cosf(az);
    pts[n].y  = r * sinf(az);
    pts[n].nx = 0.f;
    pts[n].ny = 0.f;
    ++n;
  }
  return n;
}

/* NormalComputator1DSlidingWindow::computeNormals (P2..P6): fills the normals of pts in place and valid[i] = 0 for
 * the points that become Invalid (P3) */
void orc_sliding_window_normals(orc_point* pts, int32_t n, float normal_point_distance, int32_t normal_min_points,
                                uint8_t* valid) {
  const float d2 = normal_point_distance * normal_point_distance;
  for (int32_t i = 0; i < n; ++i) {
    /* one walk over the window, first towards lower indices, then towards higher ones (P2); first and second
     * moments of d_j = p_j - p_i accumulate in walking order (P4); p_i itself contributes d = 0 */
    float sx = 0.f, sy = 0.f, sxx = 0.f, sxy = 0.f, syy = 0.f;
    int32_t cnt = 1;
    for (int32_t j = i - 1; j >= 0; --j) {
      const float dx = pts[j].x - pts[i].x, dy = pts[j].y - pts[i].y;
      const float dxx = dx * dx, dyy = dy * dy;
      if (!(dxx + dyy < d2)) {
        break;
      }
      sx = sx + dx, sy = sy + dy;
      sxx = sxx + dxx, sxy = sxy + dx * dy, syy = syy + dyy;
      ++cnt;
    }
    for (int32_t j = i + 1; j < n; ++j) {
      const float dx = pts[j].x - pts[i].x, dy = pts[j].y - pts[i].y;
      const float dxx = dx * dx, dyy = dy * dy;
      if (!(dxx + dyy < d2)) {
        break;
      }
      sx = sx + dx, sy = sy + dy;
      sxx = sxx + dxx, sxy = sxy + dx * dy, syy = syy + dyy;
      ++cnt;
    }
    valid[i] = cnt >= normal_min_points; /* P3 */
    if (!valid[i]) {
      continue;
    }
    const float fc = (float) cnt;
    const float mx = sx / fc, my = sy / fc;
    const float cxx = sxx / fc - mx * mx, cxy = sxy / fc - mx * my, cyy = syy / fc - my * my;
    float nx, ny;
    smallest_eigenvector_2x2(cxx, cxy, cyy, &nx, &ny); /* P5 */
    if (nx * pts[i].x + ny * pts[i].y > 0.f) {          /* P6 */
      nx = -nx, ny = -ny;
    }
    pts[i].nx = nx, pts[i].ny = ny;
  }
}

/* PointCloud::voxelize(out, res_coeffs) (P7): res_coeffs = the four resolutions as the reference's sources write them
 * ((res, res, 1, 1) in the pre-processor, (res, res, 0.1, 0.1) in the clipper); valid == NULL: all points valid */
int32_t orc_voxelize(const orc_point* pts, const uint8_t* valid, int32_t n, const float* res_coeffs, orc_point* out) {
  const float inv[4] = {1.f / res_coeffs[0], 1.f / res_coeffs[1], 1.f / res_coeffs[2], 1.f / res_coeffs[3]};
  return voxelize_cloud(pts, valid, n, inv, out);
}

int32_t orc_preprocess_scan(const orc_scan_params* sp, const float* ranges, int32_t n_beams, orc_point* out) {
  if (n_beams <= 0) {
    return 0;
  }
  /* _processLaserMessage, .cpp:83-90 */
  const float range_max  = sp->msg_range_max < sp->range_max ? sp->msg_range_max : sp->range_max;
  const float range_min  = sp->msg_range_min > sp->range_min ? sp->msg_range_min : sp->range_min;
  const float sensor_res = (sp->angle_max - sp->angle_min) / (float) n_beams;
  const float fx = 1.f / sensor_res, cx = (float) n_beams / 2.f;
  orc_point* pts   = (orc_point*) malloc(sizeof(orc_point) * (size_t) n_beams);
  uint8_t* valid   = (uint8_t*) malloc((size_t) n_beams);
  const int32_t n  = orc_unproject(range_min, range_max, fx, cx, ranges, n_beams, pts);               /* .cpp:29-31 */
  orc_sliding_window_normals(pts, n, sp->normal_point_distance, sp->normal_min_points, valid);       /* .cpp:32 */
  int32_t k = 0;
  if (sp->voxelize_resolution > 0.f) { /* .cpp:38-42, P7: res_coeffs = (res, res, 1, 1) */
    const float res_coeffs[4] = {sp->voxelize_resolution, sp->voxelize_resolution, 1.f, 1.f};
    k = orc_voxelize(pts, valid, n, res_coeffs, out);
  } else { /* .cpp:44-48, P8 */
    for (int32_t i = 0; i < n; ++i) {
      if (valid[i]) {
        out[k++] = pts[i];
      }
    }
  }
  free(pts);
  free(valid);
  return k;
}

void orc_preprocess_scans(const orc_scan_params* sp, const float* ranges, int32_t n_beams, int32_t n_scans,
                          int32_t n_threads, orc_point* out, int32_t* counts) {
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 16) num_threads(n_threads > 1 ? n_threads : 1)
#endif
  for (int32_t s = 0; s < n_scans; ++s) {
    counts[s] = orc_preprocess_scan(sp, ranges + (size_t) s * n_beams, n_beams, out + (size_t) s * n_beams);
  }
}
