/*
 * ls2d_oracle.h -- CPU restatement ("oracle") of the srrg2_laser_slam_2d projective
 * 2D scan-to-local-map registration path.
 *
 * THIS IS TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it.  The product library
 * (libls2d.so) never links, loads or calls anything in this directory.
 *
 * PARITY UNPINNED: the reference cannot be built here (it needs Eigen3, srrg2_core,
 * srrg2_solver, srrg2_slam_interfaces, catkin -- none present, none version-pinned:
 * srrg2_laser_slam_2d/package.xml:14-21) and its own tests hold no golden vector for
 * this path.  What is in-repo is followed line by line
 * (correspondence_finder_projective_2d.cpp:18-77); what lives in the un-vendored
 * dependencies is restated from their published behaviour, and every result-affecting
 * choice is a numbered decision point (D1..D19 aligner, L1..L8 Levenberg-Marquardt, I1 / T1 aligner options, P1..P8 raw-scan pre-processor) documented in
 * ls2d_oracle.c.  The one number the reference's own tests pin on a restated row -- the Synthetic
 * fixture pre-processes into exactly 100 points (tests/test_measurement_adaptor.cpp:36) -- is checked
 * in tests/test_oracle_preprocess.py.
 *
 * PARTLY PINNED AGAINST THE REFERENCE'S OWN SOURCES: four in-repo files -- correspondence_finder_projective_2d.cpp,
 * merger_projective_2d.cpp, scene_clipper_projective_2d.cpp, raw_data_preprocessor_projective_2d.cpp -- are compiled where they lie under /root/reference,
 * unmodified, against stand-in headers for the absent libraries (oracle/ref_shim/, oracle/ref_harness.cpp ->
 * oracle/_ref/libls2d_ref.so, built by oracle/Makefile) and orc_find_correspondences / orc_merge / orc_clip_scene /
 * orc_preprocess_scan must agree with them bit for bit (tests/test_oracle_vs_reference_sources.py).  That pins the control flow the
 * reference itself owns (gates, strictness, ordering, caching, the merger's decision tree); the upstream arithmetic
 * (projector, factor, robustifier, solver) is the oracle's own in both arms and stays unpinned.
 *
 * Arithmetic contract: every floating-point operation below is ONE IEEE-754 binary32
 * operation (no FMA contraction; build with -ffp-contract=off), in the order Eigen
 * evaluates the reference's expressions; transcendental functions are glibc's
 * (atan2f, sinf, cosf, logf, sqrtf).  The 3x3 solve is binary64 (Cholmod is double).
 */
#ifndef LS2D_ORACLE_H
#define LS2D_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* PointNormal2f [srrg2_core, used at correspondence_finder_normal_2f.h:9-12] */
typedef struct {
  float x, y, nx, ny;
} orc_point;

/* SE(2) isometry kept as Eigen's Isometry2f content: R = [c -s; s c], t = (tx, ty) */
typedef struct {
  float tx, ty, c, s;
} orc_iso;

/* one cell of PointNormal2fProjectorPolar::TargetMatrixType
 * (fields read at correspondence_finder_projective_2d.cpp:61-73) */
typedef struct {
  int32_t source_idx; /* -1 = empty */
  float depth;        /* range rho [m]; FLT_MAX when empty (D4) */
  float px, py, nx, ny; /* "transformed" point in the camera frame */
} orc_cell;

typedef struct {
  /* PointNormal2fProjectorPolar params (LASER_0.json:312-338) */
  int32_t canvas_cols;
  float angle_col_min, angle_col_max;
  float range_min, range_max;
  /* CorrespondenceFinderProjective2f params (correspondence_finder_projective_2d.h:16-21) */
  float point_distance, normal_cos;
  /* RobustifierCauchy chi_threshold (LASER_0.json:76-81); <= 0 means "no robustifier" */
  float cauchy_chi_threshold;
  /* IterationAlgorithmGN damping (LASER_0.json:83-88) */
  float damping;
  /* MultiAligner2D (LASER_0.json:9-37) / slice (LASER_0.json:115-143) */
  int32_t max_iterations;
  int32_t min_num_correspondences;
  int32_t min_num_inliers;
  /* AlignerSliceProcessorLaser2DWithSensor: 1 = sensor_in_robot as (x, y, theta); 2 = as the isometry
   * (tx, ty) = sensor_in_robot[0..1], (c, s) = sensor_in_robot_cs */
  int32_t with_sensor;
  float sensor_in_robot[3];
  float sensor_in_robot_cs[2];
  /* the rest mirrors ls2d_params (include/ls2d.h) field for field */
  int32_t factor;                  /* ORC_FACTOR_PLANE2PLANE | ORC_FACTOR_POINT2POINT (D19) */
  int32_t algorithm;               /* ORC_ALGORITHM_GN | ORC_ALGORITHM_LM (L1..L8) */
  float lm_user_lambda_init, lm_tau, lm_step_low, lm_step_high;
  int32_t lm_iterations_max, lm_variable_damping;
  int32_t single_rounding_accumulation; /* kernel selection only; the oracle's arithmetic follows tree_threads */
  int32_t enable_inlier_only_runs;      /* I1 */
  int32_t keep_only_inlier_correspondences;
  float termination_epsilon;            /* T1; <= 0: none */
} orc_params;

enum { ORC_FACTOR_PLANE2PLANE = 0, ORC_FACTOR_POINT2POINT = 1 };
enum { ORC_ALGORITHM_GN = 0, ORC_ALGORITHM_LM = 1 };

enum {
  ORC_STATUS_SUCCESS = 0,
  ORC_STATUS_NOT_ENOUGH_CORRESPONDENCES = 1,
  ORC_STATUS_NOT_ENOUGH_INLIERS = 2,
  ORC_STATUS_SINGULAR = 3
};

/* 80-byte result record; same field order as ls2d_result in include/ls2d.h */
typedef struct {
  float x, y, theta;     /* movingInFixed() as t2v */
  float chi_inliers;     /* last iteration's stats (linearisation point of that iteration) */
  float chi_kernelized;
  int32_t n_inliers;
  int32_t n_kernelized;
  int32_t n_corr;
  int32_t status;
  int32_t iterations;    /* iterations actually run */
  float H[6];            /* H00 H01 H02 H11 H12 H22 of the last linearisation */
  float c, s;            /* rotation of movingInFixed() as the state holds it (D12) */
  int32_t lm_rejected;   /* rejected Levenberg-Marquardt trial steps, all rounds */
  int32_t reserved;
} orc_result;

/* per-iteration record (40 B); pose is the estimate AFTER that iteration's update */
typedef struct {
  float x, y, theta;
  float chi_inliers, chi_kernelized;
  int32_t n_inliers, n_kernelized, n_corr;
  float c, s;
} orc_iter_stats;

/* accumulation order (D10): sequential in correspondence order (the reference's), or the
 * fixed-shape tree of the CUDA kernel: thread t owns moving points t, t+T, ...; per-thread
 * sequential, xor-butterfly inside each 32-lane warp, then sequential over warps. */
enum { ORC_SUM_SEQUENTIAL = 0, ORC_SUM_TREE = 1 };
/* ORC_SUM_TREE reads tree_threads as: bits 0..15 threads per pair; bit 16 how a warp combines its lanes (0 xor-
 * butterfly, 1 two ascending 16-lane halves); bit 17 fused accumulation arithmetic (decision D18: gates and pixel
 * indices single-rounding, error / Jacobian entries and the sums with fmaf in the kernel's association) */

void orc_default_params(orc_params* p);

/* geometry2d::v2t / t2v, Isometry2f::inverse, operator* */
orc_iso orc_v2t(float x, float y, float theta);
void orc_t2v(orc_iso T, float* xyt);
orc_iso orc_inverse(orc_iso T);
orc_iso orc_compose(orc_iso A, orc_iso B);

/* PointNormal2fProjectorPolar: setCameraPose(camera_pose) + compute(image, pts[0..n)) */
void orc_project(const orc_params* prm, orc_iso camera_pose, const orc_point* pts, int32_t n,
                 orc_cell* image /* canvas_cols cells */);

/* CorrespondenceFinderProjective2f::compute() with an already projected fixed image.
 * Returns the number of correspondences; fills moving_image (canvas_cols cells). */
int32_t orc_find_correspondences(const orc_params* prm, const orc_cell* fixed_image,
                                 const orc_point* moving, int32_t n_moving,
                                 orc_iso local_map_in_sensor, orc_cell* moving_image,
                                 int32_t* fixed_idx, int32_t* moving_idx);

/* SE2Plane2PlaneErrorFactor::errorAndJacobian for one correspondence: e[3], J[9] row-major */
void orc_error_and_jacobian(const orc_params* prm, orc_iso X, orc_point fixed, orc_point moving,
                            float* e, float* J);

/* MultiAligner2D::compute() for one pair; iter_stats holds max_iterations records (2 * max_iterations with
 * enable_inlier_only_runs).  orc_align takes the initial guess as (x, y, theta), orc_align_iso as the isometry. */
void orc_align_iso(const orc_params* prm, const orc_point* fixed, int32_t n_fixed, const orc_point* moving,
                   int32_t n_moving, orc_iso init, int32_t sum_mode, int32_t tree_threads, orc_result* out,
                   orc_iter_stats* iter_stats);
void orc_align(const orc_params* prm, const orc_point* fixed, int32_t n_fixed,
               const orc_point* moving, int32_t n_moving, const float* init_xyt,
               int32_t sum_mode, int32_t tree_threads, orc_result* out,
               orc_iter_stats* iter_stats /* max_iterations records or NULL */);

/* batch over CSR clouds; n_threads <= 1 runs the reference's single-threaded model,
 * otherwise an OpenMP parallel-for over the independent pairs.  init_pose: pose_stride floats per pair
 * (3: x, y, theta; 4: tx, ty, c, s) */
void orc_align_batch(const orc_params* prm, const orc_point* fixed_pts, const int32_t* fixed_off,
                     const orc_point* moving_pts, const int32_t* moving_off,
                     const int32_t* fixed_id, const int32_t* moving_id, const float* init_pose,
                     int32_t pose_stride, int32_t n_pairs, int32_t sum_mode, int32_t tree_threads,
                     int32_t n_threads, orc_result* out, orc_iter_stats* iter_stats);


/* ---- multi-slice aligner (MULTI.json:700-730: laser_0 + odometry prior + laser_1 in one 3x3 system) ----
 * A laser slice is described by an orc_params (its projector, finder, robustifier, min_num_correspondences and
 * sensor_in_robot); max_iterations / min_num_inliers / damping are read from slice 0. */
#define ORC_MAX_SLICES 4

/* AlignerSliceOdom2DPrior -> SE2PriorErrorFactor (LASER_0.json:291-310, MULTI.json:400-422): the odometry's
 * prediction z of moving_in_fixed with information matrix Omega (upper triangle O00 O01 O02 O11 O12 O22);
 * cauchy_chi_threshold <= 0: no robustifier (both configurations: "#pointer" -1) */
typedef struct {
  float z[4];          /* z_is_iso == 0: (x, y, theta); 1: (tx, ty, c, s) */
  int32_t z_is_iso;
  float information[6];
  float cauchy_chi_threshold;
} orc_prior;

/* e[3] and J[9] (row-major) of the prior factor at estimate X (decision D15) */
void orc_prior_error_and_jacobian(const orc_prior* prior, orc_iso X, float* e, float* J);

/* MultiAligner2D::compute() with n_slices laser slices (slice s aligns moving[s] onto fixed[s]) and an
 * optional prior (NULL: none) */
void orc_align_multi(const orc_params* slices, int32_t n_slices, const orc_point* const* fixed,
                     const int32_t* n_fixed, const orc_point* const* moving, const int32_t* n_moving,
                     const orc_prior* prior, orc_iso init, int32_t sum_mode, int32_t tree_threads,
                     orc_result* out, orc_iter_stats* iter_stats);

/* batch: slice s reads clouds fixed_id[p] / moving_id[p] (NULL: p) of its own CSR sets; prior_z [n_pairs * pose_stride]
 * (NULL: no prior) shares information / threshold of `prior`; init_pose / prior_z: pose_stride floats per pair */
void orc_align_multi_batch(const orc_params* slices, int32_t n_slices, const orc_point* const* fixed_pts,
                           const int32_t* const* fixed_off, const orc_point* const* moving_pts,
                           const int32_t* const* moving_off, const int32_t* fixed_id, const int32_t* moving_id,
                           const orc_prior* prior, const float* prior_z, const float* init_pose, int32_t pose_stride,
                           int32_t n_pairs, int32_t sum_mode, int32_t tree_threads, int32_t n_threads,
                           orc_result* out, orc_iter_stats* iter_stats);

/* loop-closure acceptance gates (MultiLoopDetectorBruteForce2D, LASER_0.json:627-634) and the
 * deterministic best-of rule. Returns index of the best accepted result or -1. */
int32_t orc_accept(const orc_result* r, int32_t min_inliers, float max_chi_per_inlier,
                   float min_inlier_ratio);
int32_t orc_best_of(const orc_result* r, int32_t n, int32_t min_inliers, float max_chi_per_inlier,
                    float min_inlier_ratio);

int32_t orc_max_threads(void);

/* SceneClipperProjective2D::compute (voxelize_resolution == 0): visibility clip of a scene seen from
 * robot_in_local_map * sensor_in_robot; `out` holds canvas_cols points; returns the count */
int32_t orc_clip_scene(const orc_params* prm, const orc_point* scene, int32_t n_scene,
                       orc_iso robot_in_local_map, orc_iso sensor_in_robot, orc_point* out);

/* the same with the clipper's voxelize branch (.cpp:36-48: res_coeffs (res, res, 0.1, 0.1), applied to the
 * points in the sensor frame, before the move to the robot frame); voxelize_resolution <= 0: orc_clip_scene */
int32_t orc_clip_scene_voxelized(const orc_params* prm, const orc_point* scene, int32_t n_scene,
                                 orc_iso robot_in_local_map, orc_iso sensor_in_robot, float voxelize_resolution,
                                 orc_point* out);

/* MergerProjective2D::compute: merges `measurement` into `scene` (room for n_scene + canvas_cols points
 * required); counters = {new, merged, replaced} (nullable); returns the new scene size */
int32_t orc_merge(const orc_params* prm, float merge_threshold, orc_point* scene, int32_t n_scene,
                  const orc_point* measurement, int32_t n_measurement, orc_iso measurement_in_scene,
                  int32_t* counters);

/* ---- RawDataPreprocessorProjective2D: LaserMessage ranges -> PointNormal2f cloud
 * R/sensor_processing/raw_data_preprocessor_projective_2d.cpp:13-51,77-104; decision points P1..P8 in the .c */
typedef struct {
  float angle_min, angle_max;         /* LaserMessage angle_min / angle_max (.cpp:85-86) */
  float msg_range_min, msg_range_max; /* LaserMessage range_min / range_max (.cpp:83-84) */
  float range_min, range_max;         /* PARAMs range_min / range_max (.h:39-40) */
  float voxelize_resolution;          /* PARAM (.h:41-45); <= 0: valid-only copy */
  float normal_point_distance;        /* NormalComputator1DSlidingWindow (L0.json:711-719) */
  int32_t normal_min_points;
} orc_scan_params;

void orc_default_scan_params(orc_scan_params* p);

/* the pre-processor's three upstream stages (P1, P2..P6, P7), separately callable: orc_preprocess_scan is their
 * composition by the in-repo logic of raw_data_preprocessor_projective_2d.cpp */
int32_t orc_unproject(float range_min, float range_max, float fx, float cx, const float* ranges, int32_t n_beams,
                      orc_point* pts);
void orc_sliding_window_normals(orc_point* pts, int32_t n, float normal_point_distance, int32_t normal_min_points,
                                uint8_t* valid);
int32_t orc_voxelize(const orc_point* pts, const uint8_t* valid, int32_t n, const float* res_coeffs, orc_point* out);

/* one scan; `out` holds n_beams points; returns the number of points produced */
int32_t orc_preprocess_scan(const orc_scan_params* sp, const float* ranges, int32_t n_beams, orc_point* out);

/* batch: scan s reads ranges[s * n_beams ..], writes out[s * n_beams ..] and counts[s] */
void orc_preprocess_scans(const orc_scan_params* sp, const float* ranges, int32_t n_beams, int32_t n_scans,
                          int32_t n_threads, orc_point* out, int32_t* counts);

/* host libm bulk drivers for tests/test_math_host.py */
void orc_libm_atan2f_n(const float* y, const float* x, float* out, long n);
void orc_libm_sincosf_n(const float* x, float* s, float* c, long n);
void orc_libm_logf_n(const float* x, float* out, long n);
void orc_column_n(const orc_params* prm, const float* y, const float* x, int32_t* col, long n);

#ifdef __cplusplus
}
#endif
#endif
