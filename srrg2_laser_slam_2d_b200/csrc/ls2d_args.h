// ls2d_args.h -- plain records that cross the host / device boundary of libls2d.so: kernel parameters and launch
// arguments.  No device code: the API translation unit includes this without seeing any kernel.
#pragma once

#include <cuda_runtime.h>

#include "../../include/ls2d.h"
#include "ls2d_math.cuh"

namespace ls2d {

constexpr int MAX_SLICES = LS2D_MAX_SLICES;

struct dev_params {
  polar_cam cam;
  float range_min, range_max;
  range_gate2 gate2;  // the same gate on the squared range (ls2d_math.cuh)
  float point_distance, normal_cos;
  float tau, inv_tau;  // Cauchy threshold (<= 0: none) and 1/tau
  float damping;
  int max_iterations, min_num_correspondences, min_num_inliers;
  int with_sensor;
  iso Sinv;  // sensor_in_robot^-1
  int factor;  // LS2D_FACTOR_PLANE2PLANE | LS2D_FACTOR_POINT2POINT
  // options only the general kernel (icp_stream_kernel) implements
  int algorithm;  // LS2D_ALGORITHM_GN | LS2D_ALGORITHM_LM
  float lm_user_lambda_init, lm_tau, lm_step_low, lm_step_high;
  int lm_iterations_max, lm_variable_damping;
  int inlier_only_runs;
  float termination_epsilon;
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
  const float* init_pose;  // pose_stride floats per pair
  int pose_stride;         // 3: (x, y, theta) -> v2t on the device; 4: (tx, ty, c, s) used verbatim
  ls2d_result* out;
  ls2d_iter_stats* iters;  // nullable
  int iters_stride;        // iteration records per pair (max_iterations; twice that with inlier-only runs)
  int n_pairs;
  int score_only;  // 1: one linearisation, no update
  int pair_base;   // first pair of this launch (chunked host pipeline); grid = n_pairs CTAs
};

// ---- service kernels (ls2d_service.cuh)
struct project_args {
  const float4* pts;
  const int* off;
  int cloud;
  float cam_pose[4];  // pose_stride floats are valid (include/ls2d.h: pose formats)
  int pose_stride;
  int* source_idx;  // [C]
  float* depth;     // [C]
};

struct correspond_args {
  const float4* fixed_pts;
  const int* fixed_off;
  const float4* moving_pts;
  const int* moving_off;
  int fixed_cloud, moving_cloud;
  float lmis_pose[4];  // local_map_in_sensor, pose_stride floats valid
  int pose_stride;
  int* fixed_idx;     // [C]
  int* moving_idx;    // [C]
  int* count;
};

// where a one-CTA-per-request kernel puts its rows: packed CSR (offsets agreed by a decoupled look-back, see
// lookback_exclusive in ls2d_service.cuh) or strided
struct pack_target {
  float4* packed;              // CSR points, nullptr: strided rows (out + b * stride) and no look-back
  int* off;                    // [n + 1] CSR offsets
  unsigned long long* state;   // [n] look-back words
  unsigned epoch;              // 1 .. 2^30 - 1
};

struct clip_args {
  const float4* pts;
  const int* off;
  const int* cloud_ids;     // [n]
  const float* robot_pose;  // [n * pose_stride] robot_in_local_map
  float sensor_pose[4];     // sensor_in_robot, pose_stride floats valid
  int pose_stride;
  float4* out;              // [n * C] strided rows (pack.packed == nullptr)
  int* counts;              // [n], may be nullptr with packed rows
  int base;                 // index of this launch's first request in cloud_ids / robot_pose / out / counts / pack
  pack_target pack;
};

struct merge_args {
  float4* scene;        // in/out, `capacity` points
  int* scene_size;      // in/out
  int capacity;
  const float4* meas;
  int n_meas;
  float mis_pose[4];    // measurement_in_scene, pose_stride floats valid
  int pose_stride;
  float merge_threshold;
  int* counters;        // [4]: new, merged, replaced, overflow flag
};

struct classify_args {
  const float4* fixed_pts;   // the two clouds (already offset to their first point)
  const float4* moving_pts;
  int n_fixed, n_moving;
  float X_pose[4];           // moving_in_fixed, pose_stride floats valid
  int pose_stride;
  const int* fixed_idx;      // [n]
  const int* moving_idx;     // [n]
  int n;
  unsigned char* is_inlier;  // [n]
};

// ---- multi-slice aligner (ls2d_multi.cuh)
struct dev_slice {
  dev_params P;
  const float4* fixed_pts;
  const int* fixed_off;
  const float4* moving_pts;
  const int* moving_off;
};

struct multi_args {
  dev_slice sl[MAX_SLICES];
  int n_slices;
  int max_cols, max_points;  // capacity of the shared z-buffer / stash
  int max_fixed_points;      // the largest fixed cloud alone (kernel selection)
  int fused;                 // icp_multi2_kernel: the fused accumulation arithmetic (D18) unless a slice asks for
                             // single-rounding sums
  const int* fixed_id;       // nullable: pair index
  const int* moving_id;      // nullable: pair index
  const float* init_pose;    // pose_stride floats per pair
  int pose_stride;           // 3: (x, y, theta); 4: (tx, ty, c, s)
  const float* prior_z;      // nullable: [n_pairs * pose_stride], the prior slice's measurement
  float prior_info[6];       // O00 O01 O02 O11 O12 O22
  float prior_tau, prior_inv_tau;
  ls2d_result* out;
  ls2d_iter_stats* iters;  // nullable
  int n_pairs;
  int score_only;
};

// ---- raw-scan pre-processor (ls2d_scan.cuh)
struct scan_dev_params {
  float range_min, range_max;  // the tighter of message and PARAM limits (.cpp:83-84)
  float ifx, cx;               // azimuth = ifx * (c - cx), sensor matrix [1/res, n/2] (.cpp:87-90)
  float d2;                    // normal_point_distance^2
  float inv_res;               // 1 / voxelize_resolution, 0: valid-only copy (.cpp:44-48)
  int min_points;              // normal_min_points
  int n_beams;
};

struct scan_args {
  const float* ranges;  // [n_scans][n_beams]
  const float2* beam_cs;  // [n_beams] (cos, sin) of the beams' azimuths (beam_table_kernel)
  float4* out;          // [n_scans][n_beams] strided rows
  int* counts;          // [n_scans]
  int n_scans;
  int* off;             // [n_scans + 1] CSR offsets of the packed copy, written by the last CTA to finish; or nullptr
  int continues;        // off[0] already holds the offset of scan 0 (a job cut into several launches), else it is 0
  int* ticket;          // zero between launches: CTAs that have finished
};

}  // namespace ls2d
