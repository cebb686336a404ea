/*
 * ls2d.h -- C ABI of the B200-native projective 2D scan-to-local-map registration path.
 *
 * This is the drop-in boundary for ONE hot path of rvp-group/srrg2_laser_slam_2d: everything that runs
 * inside one MultiAligner2D::compute() call with laser slices.  Each entry point
 * names the reference interface it replaces.  Reference paths are relative to
 * /root/reference/srrg2_laser_slam_2d/ ; R/ = src/srrg2_laser_slam_2d/ ;
 * L0.json = /root/reference/configurations/stage_segway_double_config_LASER_0.json.
 *
 * Conventions
 *  - plain C: opaque handle, plain pointers and sizes, no C++/torch types.
 *  - every function returns LS2D_OK (0) or a negative ls2d_error; nothing throws.
 *  - pointers are HOST pointers unless the parameter name ends in _dev.
 *  - a point is 4 floats (x, y, nx, ny) == PointNormal2f coordinates() + normal()
 *    (R/registration/correspondence_finder_normal_2f.h:9-12); a cloud set is a CSR batch:
 *    points [total, 4] + offsets [n_clouds + 1].
 *  - poses cross the boundary in the handle's POSE FORMAT (ls2d_set_pose_format): LS2D_POSE_XYT = 3 floats
 *    (x, y, theta) == geometry2d::t2v(Isometry2f), rebuilt on the device with cosf/sinf; LS2D_POSE_ISO = 4 floats
 *    (tx, ty, c, s) == the Isometry2f itself (translation() and the first column of linear(), R = [c -s; s c]),
 *    used verbatim -- the format a caller that HOLDS an Isometry2f must use to get the reference's bits, because
 *    v2t(t2v(T)) != T in binary32 (the reference hands matrices to finder / clipper / merger / aligner:
 *    R/registration/correspondence_finder_projective_2d.cpp:40,47, R/mapping/scene_clipper_projective_2d.cpp:22-32,
 *    R/mapping/merger_projective_2d.cpp:19-22, apps/visual_test_aligner_2d.cpp:123-128).  Every parameter named
 *    *_pose below holds `stride` floats per pose, stride = 3 or 4 by that format.  Results always carry both forms.
 *  - one handle == one device + one CUDA stream; a handle is not re-entrant (the reference's modules
 *    are not thread-safe either), different handles may be used from different threads.
 *  - there is NO CPU fallback: without a CUDA device every compute call returns LS2D_ERR_CUDA.
 */
#ifndef LS2D_H
#define LS2D_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LS2D_VERSION 200

typedef struct ls2d_handle ls2d_handle;

typedef enum {
  LS2D_OK               = 0,
  LS2D_ERR_INVALID      = -1, /* bad argument (mis-wiring; the reference throws std::runtime_error) */
  LS2D_ERR_CUDA         = -2, /* CUDA runtime/driver error, or no device */
  LS2D_ERR_NOT_READY    = -3, /* clouds / params not set (the reference: "Missing fixed!") */
  LS2D_ERR_UNSUPPORTED  = -4, /* size outside the compiled kernel table */
  LS2D_ERR_NCCL         = -5
} ls2d_error;

/* All parameters of the path, with the reference's names and defaults. */
typedef struct {
  /* PointNormal2fProjectorPolar (L0.json:312-338) */
  int32_t canvas_cols;   /* 721 */
  float angle_col_min;   /* -3.14159 */
  float angle_col_max;   /*  3.14159 */
  float range_min;       /* 0.3 */
  float range_max;       /* 20 */
  /* CorrespondenceFinderProjective2f (R/registration/correspondence_finder_projective_2d.h:16-21) */
  float point_distance;  /* 0.5 */
  float normal_cos;      /* 0.8 */
  /* RobustifierCauchy.chi_threshold (L0.json:76-81); <= 0: slice has no robustifier (#pointer -1) */
  float cauchy_chi_threshold; /* 0.01 */
  /* IterationAlgorithmGN.damping (L0.json:83-88) */
  float damping;         /* 0 */
  /* MultiAligner2D.max_iterations / min_num_inliers (L0.json:9-37, 487-517) */
  int32_t max_iterations;          /* 10 */
  int32_t min_num_correspondences; /* AlignerSliceProcessor*.min_num_correspondences, 0 (L0.json:134) */
  int32_t min_num_inliers;         /* 10 */
  /* AlignerSliceProcessorLaser2DWithSensor (R/registration/aligner_slice_processor_laser_2d.h:21-42):
   * sensor_in_robot looked up from the tf tree by setupFactor()
   * (R/registration/aligner_slice_processor_laser_2d_impl.cpp:7-10) */
  int32_t with_sensor;             /* 0 = AlignerSliceProcessorLaser2D, 1 = WithSensor, sensor_in_robot as
                                    * (x, y, theta), 2 = WithSensor, sensor_in_robot as the isometry
                                    * (tx, ty) = sensor_in_robot[0..1], (c, s) = sensor_in_robot_cs */
  float sensor_in_robot[3];
  float sensor_in_robot_cs[2];
  /* factor bound by the slice (the reference binds SE2Plane2PlaneErrorFactor,
   * R/registration/aligner_slice_processor_laser_2d.h:8,23; BASELINE.json's north_star also names point-to-point) */
  int32_t factor;                  /* LS2D_FACTOR_PLANE2PLANE (0) | LS2D_FACTOR_POINT2POINT */
  /* Solver.algorithm (L0.json:193-215 names IterationAlgorithmGN; north_star: "Gauss-Newton/LM update") */
  int32_t algorithm;               /* LS2D_ALGORITHM_GN (0) | LS2D_ALGORITHM_LM */
  /* IterationAlgorithmLM: user_lambda_init (<= 0: tau * max diag H), tau, step_low / step_high (clamps of the
   * lambda update after an accepted step), lm_iterations_max (trials per round), variable_damping (lambda * diag H
   * instead of lambda * I) */
  float lm_user_lambda_init;       /* 0 */
  float lm_tau;                    /* 1e-5 */
  float lm_step_low, lm_step_high; /* 1/3, 2/3 */
  int32_t lm_iterations_max;       /* 10 */
  int32_t lm_variable_damping;     /* 1 */
  /* arithmetic of the H/b accumulation (oracle decision D18): 0 = fused multiply-adds where
   * the kernel for the cloud size has them (default), 1 = single-rounding everywhere (the reference built without
   * FMA contraction).  Gates, pixel indices and z-buffer winners are single-rounding in both. */
  int32_t single_rounding_accumulation;
  /* MultiAligner2D.enable_inlier_only_runs / keep_only_inlier_correspondences / termination criteria
   * (L0.json:14-17,34-36; off in both shipped configurations) */
  int32_t enable_inlier_only_runs;          /* 0 */
  int32_t keep_only_inlier_correspondences; /* 0 */
  float termination_epsilon;                /* <= 0: none (run max_iterations) */
} ls2d_params;

enum { LS2D_FACTOR_PLANE2PLANE = 0, LS2D_FACTOR_POINT2POINT = 1 };
enum { LS2D_ALGORITHM_GN = 0, LS2D_ALGORITHM_LM = 1 };
enum { LS2D_POSE_XYT = 0, LS2D_POSE_ISO = 1 };

/* MultiAligner2D status (+ SINGULAR for a non positive definite H) */
typedef enum {
  LS2D_STATUS_SUCCESS                    = 0,
  LS2D_STATUS_NOT_ENOUGH_CORRESPONDENCES = 1,
  LS2D_STATUS_NOT_ENOUGH_INLIERS         = 2,
  LS2D_STATUS_SINGULAR                   = 3
} ls2d_status;

/* One alignment's outcome (80 B): movingInFixed() -- as t2v (x, y, theta) AND as the isometry itself
 * (x, y, c, s: R = [c -s; s c]) --, the last iterationStats() entry, the information matrix H
 * (apps/visual_test_aligner_2d.cpp:145-156). */
typedef struct {
  float x, y, theta;
  float chi_inliers;
  float chi_kernelized;
  int32_t n_inliers;
  int32_t n_kernelized;
  int32_t n_corr;
  int32_t status;
  int32_t iterations;
  float H[6]; /* H00 H01 H02 H11 H12 H22 */
  float c, s; /* rotation of movingInFixed(), the bits the kernel holds (theta = atan2f(s, c)) */
  int32_t lm_rejected; /* LM: rejected trial steps over all rounds (0 with GN) */
  int32_t reserved;
} ls2d_result;

/* iterationStats() record (40 B); the pose is the estimate after that iteration's update */
typedef struct {
  float x, y, theta;
  float chi_inliers, chi_kernelized;
  int32_t n_inliers, n_kernelized, n_corr;
  float c, s;
} ls2d_iter_stats;

/* loop-closure acceptance gates: MultiLoopDetectorBruteForce2D relocalize_min_inliers /
 * relocalize_max_chi_inliers / relocalize_min_inliers_ratio (L0.json:627-634) */
typedef struct {
  int32_t min_inliers;        /* 300 */
  float max_chi_per_inlier;   /* 0.1 */
  float min_inlier_ratio;     /* 0.8 */
} ls2d_gates;

/* best accepted candidate of a verification shard (48 B) -- the record the ranks all-gather */
typedef struct {
  float x, y, theta;
  float chi_inliers;
  int32_t n_inliers;
  int32_t n_corr;
  int32_t candidate;  /* global candidate id, -1 = nothing accepted */
  int32_t guess;      /* index of the winning initial guess */
  float c, s;         /* rotation of the winning pose as the kernel holds it */
  int32_t iterations; /* iterations that alignment ran */
  int32_t reserved;
} ls2d_best;

/* cloud sets of a handle: ids 0 .. LS2D_MAX_CLOUD_SETS-1.  The single-slice entry points read LS2D_FIXED and
 * LS2D_MOVING; the multi-slice aligner names a (fixed, moving) set per slice. */
enum { LS2D_FIXED = 0, LS2D_MOVING = 1 };
#define LS2D_MAX_CLOUD_SETS 8
#define LS2D_MAX_SLICES 4

/* AlignerSliceOdom2DPrior (L0.json:291-310, MULTI.json:400-422) -> SE2PriorErrorFactor: information matrix
 * (upper triangle O00 O01 O02 O11 O12 O22) of the odometry's prediction of moving_in_fixed; the prediction
 * itself is per pair.  cauchy_chi_threshold <= 0: the slice has no robustifier (both configurations). */
typedef struct {
  float information[6];
  float cauchy_chi_threshold;
} ls2d_prior;

/* ---- lifetime ------------------------------------------------------------------------------------- */
/* replaces: construction of the BOSS-registered modules (R/instances.cpp:27-36) */
int ls2d_create(ls2d_handle** h, int device);
int ls2d_destroy(ls2d_handle* h);
/* run on a caller-owned cudaStream_t (e.g. torch's current stream); NULL restores the handle's own */
int ls2d_set_stream(ls2d_handle* h, void* cuda_stream);
int ls2d_sync(ls2d_handle* h);
const char* ls2d_strerror(int err);
int ls2d_version(void);

/* ---- configuration ---------------------------------------------------------------------------------
 * replaces: PARAM() properties of CorrespondenceFinderProjective2f, PointNormal2fProjectorPolar,
 * AlignerSliceProcessorLaser2D[WithSensor], MultiAligner2D, RobustifierCauchy, IterationAlgorithmGN */
void ls2d_default_params(ls2d_params* p);
int ls2d_set_params(ls2d_handle* h, const ls2d_params* p);
int ls2d_get_params(const ls2d_handle* h, ls2d_params* p);
/* pose format of every *_pose argument of this handle: LS2D_POSE_XYT (default; 3 floats per pose) or LS2D_POSE_ISO
 * (4 floats per pose: tx, ty, c, s -- the caller's Isometry2f, used verbatim).  replaces: nothing -- the reference
 * passes Isometry2f objects; this is how they cross a C boundary without a t2v / v2t round trip */
int ls2d_set_pose_format(ls2d_handle* h, int format);
int ls2d_get_pose_format(const ls2d_handle* h);

/* ---- clouds ----------------------------------------------------------------------------------------
 * replaces: MultiAligner2D::setFixed / setMoving(PropertyContainer*) and
 * CorrespondenceFinder_::setFixed / setMoving(const PointNormal2fVectorCloud*)
 * (apps/visual_test_aligner_2d.cpp:108-126, apps/visual_test_correspondence_finder_projective_2d.cpp:73-79) */
int ls2d_upload_clouds(ls2d_handle* h, int which /* set id */, const float* points_xynn, const int32_t* offsets,
                       int32_t n_clouds);
/* borrow device-resident clouds (no copy); max_points >= the largest cloud in the set: it picks the kernel and its
 * capacity, so the offsets -- which must be complete on the device when this is called -- are checked against it
 * here (one small kernel and a 4-byte read back); LS2D_ERR_INVALID when a cloud is larger or the offsets descend */
int ls2d_set_clouds_dev(ls2d_handle* h, int which, const void* points_dev, const int32_t* offsets_dev,
                        int32_t n_clouds, int32_t max_points);

/* ---- registration ----------------------------------------------------------------------------------
 * replaces: MultiAligner2D::setMovingInFixed + compute + movingInFixed + iterationStats
 * (apps/visual_test_aligner_2d.cpp:123-156), batched: pair p aligns moving cloud moving_id[p] onto fixed
 * cloud fixed_id[p] from init_pose[p].  NULL ids mean id == p.  iter_stats (nullable) holds
 * n_pairs * max_iterations records. */
int ls2d_align_batch(ls2d_handle* h, const int32_t* fixed_id, const int32_t* moving_id,
                     const float* init_pose, int32_t n_pairs, ls2d_result* out,
                     ls2d_iter_stats* iter_stats);
/* same, everything device-resident, asynchronous on the handle's stream */
int ls2d_align_batch_dev(ls2d_handle* h, const int32_t* fixed_id_dev, const int32_t* moving_id_dev,
                         const float* init_pose_dev, int32_t n_pairs, ls2d_result* out_dev,
                         ls2d_iter_stats* iter_stats_dev);
/* one call from host buffers: upload both cloud sets, align pair p = (fixed p, moving p), download.  The batch is
 * cut into chunks whose uploads overlap the previous chunk's kernel; pinned (or registered) buffers go to the copy
 * engine directly, pageable ones through the handle's pinned ring filled by a few copy threads */
int ls2d_align_pairs_host(ls2d_handle* h, const float* fixed_points, const int32_t* fixed_offsets,
                          const float* moving_points, const int32_t* moving_offsets,
                          const float* init_pose, int32_t n_pairs, ls2d_result* out);
/* linearise once at init_pose without updating the pose (chi / inliers / H of a guess) */
int ls2d_score_batch(ls2d_handle* h, const int32_t* fixed_id, const int32_t* moving_id,
                     const float* pose, int32_t n_pairs, ls2d_result* out);
int ls2d_score_batch_dev(ls2d_handle* h, const int32_t* fixed_id_dev, const int32_t* moving_id_dev,
                         const float* pose_dev, int32_t n_pairs, ls2d_result* out_dev);

/* ---- multi-slice registration ------------------------------------------------------------------------
 * replaces: MultiAligner2D::compute with several slice processors (MULTI.json:700-730: al_sl_laser_0 +
 * ad_sl_odom + al_sl_laser_1; LASER_0.json:502-506: laser + odom): n_slices laser slices
 * (AlignerSliceProcessorLaser2D[WithSensor], R/registration/aligner_slice_processor_laser_2d.h:7-42), slice s
 * described by slices[s] (its projector, finder, robustifier, min_num_correspondences, sensor_in_robot) and
 * aligning cloud moving_id[p] of set moving_set[s] onto cloud fixed_id[p] of set fixed_set[s]; their H and b
 * are summed with the odometry prior's (prior / prior_z_pose [n_pairs * stride], both NULL: no prior slice) and
 * solved once per iteration.  max_iterations / min_num_inliers / damping are read from slices[0].  A slice
 * with n_corr <= min_num_correspondences is skipped in that iteration. */
int ls2d_align_multi(ls2d_handle* h, const ls2d_params* slices, const int32_t* fixed_set,
                     const int32_t* moving_set, int32_t n_slices, const ls2d_prior* prior,
                     const float* prior_z_pose, const int32_t* fixed_id, const int32_t* moving_id,
                     const float* init_pose, int32_t n_pairs, ls2d_result* out, ls2d_iter_stats* iter_stats);
/* same, per-pair arrays device-resident, asynchronous on the handle's stream */
int ls2d_align_multi_dev(ls2d_handle* h, const ls2d_params* slices, const int32_t* fixed_set,
                         const int32_t* moving_set, int32_t n_slices, const ls2d_prior* prior,
                         const float* prior_z_pose_dev, const int32_t* fixed_id_dev, const int32_t* moving_id_dev,
                         const float* init_pose_dev, int32_t n_pairs, ls2d_result* out_dev,
                         ls2d_iter_stats* iter_stats_dev);

/* ---- finder / projector (drop-in + parity) ---------------------------------------------------------
 * replaces: CorrespondenceFinderProjective2f::compute()
 * (R/registration/correspondence_finder_projective_2d.cpp:18-77): ordered (ascending column) list of
 * Correspondence(fixed_idx, moving_idx); arrays must hold canvas_cols entries. */
int ls2d_find_correspondences(ls2d_handle* h, int32_t fixed_id, int32_t moving_id,
                              const float* local_map_in_sensor_pose, int32_t* fixed_idx,
                              int32_t* moving_idx, int32_t* n_correspondences);
/* same, between cloud fixed_id of set fixed_set and cloud moving_id of set moving_set (a slice of the
 * multi-slice aligner), with the handle's current parameters */
int ls2d_find_correspondences_in(ls2d_handle* h, int32_t fixed_set, int32_t moving_set, int32_t fixed_id,
                                 int32_t moving_id, const float* local_map_in_sensor_pose, int32_t* fixed_idx,
                                 int32_t* moving_idx, int32_t* n_correspondences);
/* replaces: MultiAligner2D.keep_only_inlier_correspondences (L0.json:17: "toggles removal of correspondences which
 * factors are not inliers in the last iteration"): for the n correspondences (fixed_idx[k], moving_idx[k]) between cloud
 * fixed_id of set fixed_set and cloud moving_id of set moving_set, is_inlier[k] = 1 when the slice's factor at the
 * estimate moving_in_fixed_pose has chi < cauchy_chi_threshold (every factor of a slice without robustifier is an
 * inlier), with the handle's current parameters (factor, sensor_in_robot, threshold) */
int ls2d_classify_correspondences(ls2d_handle* h, int32_t fixed_set, int32_t moving_set, int32_t fixed_id,
                                  int32_t moving_id, const float* moving_in_fixed_pose, const int32_t* fixed_idx,
                                  const int32_t* moving_idx, int32_t n, uint8_t* is_inlier);
/* replaces: PointNormal2fProjectorPolar::setCameraPose + compute
 * (R/registration/correspondence_finder_projective_2d.cpp:40-41,47-48): per column the winning
 * source_idx (-1 empty) and its depth (FLT_MAX empty); arrays hold canvas_cols entries. */
int ls2d_project(ls2d_handle* h, int which, int32_t cloud_id, const float* camera_pose,
                 int32_t* source_idx, float* depth);

/* ---- loop-closure verification ---------------------------------------------------------------------
 * replaces: the per-candidate loop of MultiLoopDetectorBruteForce2D::compute (config L0.json:613-635):
 * fixed cloud query_id against moving clouds candidate_ids[0..n_cand) from guesses_pose
 * [n_cand * n_guess * stride], acceptance gates, deterministic best-of (most inliers, then lowest chi per
 * inlier, then lowest candidate/guess).  candidate_base is added to the local candidate index in the
 * reported record so that shards report global ids.  all_results (nullable) receives every alignment. */
int ls2d_verify(ls2d_handle* h, int32_t query_id, const int32_t* candidate_ids, int32_t n_cand,
                const float* guesses_pose, int32_t n_guess, const ls2d_gates* gates,
                int32_t candidate_base, ls2d_best* best, ls2d_result* all_results);
/* device-resident variant: best_dev receives the shard's record (ready for an NCCL all-gather) */
int ls2d_verify_dev(ls2d_handle* h, int32_t query_id, const int32_t* candidate_ids_dev, int32_t n_cand,
                    const float* guesses_pose_dev, int32_t n_guess, const ls2d_gates* gates,
                    int32_t candidate_base, ls2d_best* best_dev, ls2d_result* all_results_dev);
/* all-pairs loop-closure search (BASELINE.json configs[4]): pair p aligns moving cloud moving_ids[p] onto fixed
 * cloud fixed_ids[p] from guesses_pose[p]; the pairs are grouped by group_offsets (CSR over pairs, n_groups + 1
 * entries; one group per query local map) and best[g] is the best accepted pair of group g under the same gates
 * and ordering (candidate = its moving cloud id, guess = its index inside the group; candidate -1: none accepted).
 * Ranks own disjoint runs of groups; the records all-gather like ls2d_verify's. */
int ls2d_verify_pairs(ls2d_handle* h, const int32_t* fixed_ids, const int32_t* moving_ids, const float* guesses_pose,
                      int32_t n_pairs, const int32_t* group_offsets, int32_t n_groups, const ls2d_gates* gates,
                      ls2d_best* best, ls2d_result* all_results);
int ls2d_verify_pairs_dev(ls2d_handle* h, const int32_t* fixed_ids_dev, const int32_t* moving_ids_dev,
                          const float* guesses_pose_dev, int32_t n_pairs, const int32_t* group_offsets_dev,
                          int32_t n_groups, const ls2d_gates* gates, ls2d_best* best_dev,
                          ls2d_result* all_results_dev);
/* best-of over gathered shard records (host), same ordering rule */
int ls2d_reduce_best(const ls2d_best* records, int32_t n, ls2d_best* out);
/* ls2d_verify_dev + all-gather of the 48-byte records over an existing NCCL communicator
 * (ncclComm_t passed as void*; libnccl is resolved at run time -- the copy already loaded into the process, e.g.
 * torch's, else libnccl.so.2) + ls2d_reduce_best on every rank.  n_ranks must equal the communicator's size
 * (checked with ncclCommCount), else LS2D_ERR_INVALID. */
int ls2d_verify_sharded_nccl(ls2d_handle* h, int32_t query_id, const int32_t* candidate_ids_dev,
                             int32_t n_cand, const float* guesses_pose_dev, int32_t n_guess,
                             const ls2d_gates* gates, int32_t candidate_base, void* nccl_comm,
                             int32_t n_ranks, ls2d_best* best);

/* ---- local-map maintenance around the aligner -------------------------------
 * replaces: SceneClipperProjective2D::compute (R/mapping/scene_clipper_projective_2d.cpp:11-65) with
 * voxelize_resolution == 0, the value both shipped configurations use: for request r the scene cloud
 * cloud_ids[r] of set `which` is seen from robot_in_local_map[r] * sensor_in_robot; the z-buffer winners, in
 * column order, come back as points in the ROBOT frame: out_points [n * canvas_cols * 4], out_counts [n]. */
int ls2d_clip_scenes(ls2d_handle* h, int which, const int32_t* cloud_ids, const float* robot_in_local_map_pose,
                     const float* sensor_in_robot_pose, int32_t n, float* out_points, int32_t* out_counts);
/* replaces: MergerProjective2D::compute (R/mapping/merger_projective_2d.cpp:9-100): merges a measurement cloud
 * into a scene IN PLACE (add / average+renormalise / replace / ordered append per column, merge_threshold
 * as the reference's PARAM).  scene_points holds `capacity` points of which *scene_size are valid; the call
 * needs capacity >= *scene_size + canvas_cols, else LS2D_ERR_INVALID.  counters (nullable) = {new, merged,
 * replaced}. */
int ls2d_merge_scene(ls2d_handle* h, float* scene_points, int32_t* scene_size, int32_t capacity,
                     const float* measurement_points, int32_t n_measurement,
                     const float* measurement_in_scene_pose, float merge_threshold, int32_t* counters);
/* device-resident variant, asynchronous: scene_size_dev and counters_dev (4 ints: new, merged, replaced,
 * overflow) live on the device */
int ls2d_merge_scene_dev(ls2d_handle* h, void* scene_points_dev, int32_t* scene_size_dev, int32_t capacity,
                         const void* measurement_points_dev, int32_t n_measurement,
                         const float* measurement_in_scene_pose, float merge_threshold, int32_t* counters_dev);

/* ---- raw scans as the wire format --------------------------------------------------
 * replaces: RawDataPreprocessorProjective2D::setRawData + compute
 * (R/sensor_processing/raw_data_preprocessor_projective_2d.cpp:13-51, 53-104): LaserMessage ranges -> polar
 * unprojection with the sensor matrix [1/res, n/2] (.cpp:87-90) -> sliding-window normals
 * (NormalComputator1DSlidingWindow, L0.json:711-719) -> voxelize(res, res, 1, 1) (.cpp:38-42) or the valid points
 * in beam order (.cpp:44-48).  4 B per beam cross the bus instead of 16 B per point. */
typedef struct ls2d_scan_params {
  float angle_min, angle_max;         /* LaserMessage::angle_min / angle_max (.cpp:85-86) */
  float msg_range_min, msg_range_max; /* LaserMessage::range_min / range_max (.cpp:83-84) */
  float range_min, range_max;         /* PARAM range_min / range_max (raw_data_preprocessor_projective_2d.h:39-40) */
  float voxelize_resolution;          /* PARAM voxelize_resolution (.h:41-45, default 0.02); <= 0: valid-only copy */
  float normal_point_distance;        /* NormalComputator1DSlidingWindow::normal_point_distance (L0.json:718) */
  int32_t normal_min_points;          /* NormalComputator1DSlidingWindow::normal_min_points (L0.json:715) */
} ls2d_scan_params;
void ls2d_default_scan_params(ls2d_scan_params* p);
/* n_scans scans of n_beams ranges each (host, row-major) -> out_points [n_scans * n_beams * 4] (scan s starts at
 * s * n_beams * 4; out_counts[s] points are valid), the cloud each RawDataPreprocessorProjective2D::compute would
 * hand to setMeas().  n_beams <= 8192; with voxelize_resolution > 0 a scan's working set must fit one CTA's shared
 * memory: n_beams <= 6000 (LS2D_ERR_UNSUPPORTED beyond). */
int ls2d_preprocess_scans(ls2d_handle* h, const ls2d_scan_params* sp, const float* ranges, int32_t n_beams,
                          int32_t n_scans, float* out_points, int32_t* out_counts);
/* same, but the clouds stay on the device as cloud set `which` (packed CSR), ready for ls2d_align_batch & co.;
 * asynchronous on the handle's stream */
int ls2d_preprocess_scans_to_set(ls2d_handle* h, int which, const ls2d_scan_params* sp, const float* ranges,
                                 int32_t n_beams, int32_t n_scans);
int ls2d_preprocess_scans_to_set_dev(ls2d_handle* h, int which, const ls2d_scan_params* sp, const float* ranges_dev,
                                     int32_t n_beams, int32_t n_scans);
/* copies cloud set `which` back: offsets [n_clouds + 1], points [offsets[n_clouds] * 4] (capacity_points is the
 * room in `points`); n_clouds must match the set */
int ls2d_download_clouds(ls2d_handle* h, int which, float* points, int32_t* offsets, int32_t n_clouds,
                         int64_t capacity_points);

/* the clipper's voxelize branch (.cpp:36-48): the winners, as points in the sensor frame, are voxelized with
 * res_coeffs (voxelize_resolution, voxelize_resolution, 0.1, 0.1) before the move to the robot frame;
 * voxelize_resolution <= 0 behaves like ls2d_clip_scenes */
int ls2d_clip_scenes_voxelized(ls2d_handle* h, int which, const int32_t* cloud_ids,
                               const float* robot_in_local_map_pose, const float* sensor_in_robot_pose, int32_t n,
                               float voxelize_resolution, float* out_points, int32_t* out_counts);
/* device-resident clipper: same clip as ls2d_clip_scenes, but the clipped clouds become cloud set `out_set`
 * (packed CSR, out_set != scene_set) without leaving the device -- the tracker's moving clouds */
int ls2d_clip_scenes_to_set(ls2d_handle* h, int scene_set, const int32_t* cloud_ids,
                            const float* robot_in_local_map_pose, const float* sensor_in_robot_pose, int32_t n,
                            int out_set);
/* replaces: one MultiTracker2D frame step, preprocessRawData -> clip -> align
 * (apps/visual_test_tracker_2d.cpp:167-179), batched over n frames.  Frame f: ranges[f] is
 * pre-processed into the measurement cloud (fixed), local map scene_ids[f] of resident set `scene_set` (>= 2) is
 * clipped from robot_in_local_map[f] * sensor_in_robot (sensor_in_robot = the handle's params when with_sensor,
 * else identity) into the moving cloud, and the aligner runs from init_pose[f] (NULL: identity).  Only 4 B/beam,
 * ids and poses go to the device, 80 B/frame come back.  Overwrites sets LS2D_FIXED and LS2D_MOVING (they hold the
 * whole batch afterwards).  Large batches are pipelined in chunks: the next chunk's ranges upload while the current one
 * runs, the pre-processor shares the GPU with the previous chunk's clipper and aligner.  Pageable `ranges` are staged
 * through the handle's pinned ring. */
int ls2d_track_batch(ls2d_handle* h, const ls2d_scan_params* sp, const float* ranges, int32_t n_beams, int32_t n,
                     int scene_set, const int32_t* scene_ids, const float* robot_in_local_map_pose,
                     const float* init_pose, ls2d_result* out);

/* ---- introspection ---------------------------------------------------------------------------------*/
/* shape of the H/b reduction the aligner runs with parameters *p on clouds of up to max_points points (the kernel
 * is picked by cloud size, canvas width and the options in *p): bits 0..15 = threads per pair, bit 16 = how a warp
 * combines its 32 lane partials (0: xor-butterfly, 1: lanes 0..15 and 16..31 in ascending order, then the two
 * halves), bit 17 = the contributions are accumulated with fused multiply-adds (oracle decision D18).  This is
 * the value ORC_SUM_TREE takes as tree_threads.  Negative: ls2d_error. */
int ls2d_reduction_shape(const ls2d_params* p, int32_t max_points);
/* the same for the scoring pass (ls2d_score_batch), which has a kernel of its own for clouds of up to 1152 points */
int ls2d_score_reduction_shape(const ls2d_params* p, int32_t max_points);
/* the same for the multi-slice aligner (ls2d_align_multi): the general kernel's shape, whatever the cloud sizes ... */
int ls2d_multi_reduction_threads(void);
/* ... and the shape ls2d_align_multi runs for these slices: two slices whose fixed clouds hold up to 768 points, on
 * canvases below 768 columns, that share ONE moving cloud set of up to 1536 points (the MULTI.json aligner: fixed
 * "points_0" / "points_1", moving "points") have a register-resident kernel with a shape of its own */
int ls2d_multi_reduction_shape(const ls2d_params* slices, int32_t n_slices, int32_t max_fixed_points,
                               int32_t max_moving_points, int32_t shared_moving);
/* device self-test: the kernels evaluate sqrtf on range-gated operands with the five-instruction core of __fsqrt_rn
 * (no operand-class test; csrc/ls2d_math.cuh fsqrt_gated); compares it with __fsqrt_rn on EVERY binary32 value in
 * [lo, hi] (lo > 2^-100) and reports how many were checked and how many differ (must be 0) */
int ls2d_selftest_gated_sqrt(ls2d_handle* h, float lo, float hi, int64_t* n_checked, int64_t* n_mismatch);
/* kernels launched by this handle since creation */
int64_t ls2d_launch_count(const ls2d_handle* h);

#ifdef __cplusplus
}
#endif
#endif
