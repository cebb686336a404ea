// ref_harness.cpp -- TEST INFRASTRUCTURE.  C entry points over the reference's OWN classes, compiled from the
// reference's own sources where they lie under /root/reference (see oracle/Makefile, target _ref/libls2d_ref.so) against
// the stand-in headers of oracle/ref_shim/.  Drives them the way the reference's apps do
// (apps/visual_test_correspondence_finder_projective_2d.cpp:73-79, apps/visual_test_merger_projective_2d.cpp:100-125).
#include "srrg2_laser_slam_2d/mapping/merger_projective_2d.h"
#include "srrg2_laser_slam_2d/mapping/scene_clipper_projective_2d.h"
#include "srrg2_laser_slam_2d/registration/correspondence_finder_projective_2d.h"
#include "srrg2_laser_slam_2d/sensor_processing/raw_data_preprocessor_projective_2d.h"

using namespace srrg2_core;
using namespace srrg2_laser_slam_2d;

namespace {
  void to_cloud(const orc_point* p, int32_t n, PointNormal2fVectorCloud& c) {
    c.resize((size_t) n);
    for (int32_t i = 0; i < n; ++i) {
      c[(size_t) i].coordinates() = Vector2f(p[i].x, p[i].y);
      c[(size_t) i].normal()      = Vector2f(p[i].nx, p[i].ny);
    }
  }
  void from_cloud(const PointNormal2fVectorCloud& c, orc_point* p) {
    for (size_t i = 0; i < c.size(); ++i) {
      p[i].x = c[i].coordinates().x(), p[i].y = c[i].coordinates().y();
      p[i].nx = c[i].normal().x(), p[i].ny = c[i].normal().y();
    }
  }
  void configure(PointNormal2fProjectorPolar& pr, const orc_params* prm) {
    pr.param_canvas_cols.setValue(prm->canvas_cols);
    pr.param_angle_col_min.setValue(prm->angle_col_min);
    pr.param_angle_col_max.setValue(prm->angle_col_max);
    pr.param_range_min.setValue(prm->range_min);
    pr.param_range_max.setValue(prm->range_max);
  }
  // a pose in either wire format: stride 3 = (x, y, theta) through geometry2d::v2t, stride 4 = the Isometry2f content
  // (tx, ty, c, s) itself -- what a caller of the reference holds (e.g. an accumulated product of increments)
  Isometry2f iso(const float* p, int32_t stride = 3) {
    if (stride == 4) {
      orc_iso T;
      T.tx = p[0], T.ty = p[1], T.c = p[2], T.s = p[3];
      return Isometry2f(T);
    }
    return Isometry2f(orc_v2t(p[0], p[1], p[2]));
  }
}  // namespace

extern "C" {

// CorrespondenceFinderProjective2f::compute(); n_calls > 1 repeats compute() on the same object with the poses
// xyt[3 * k] (exercises the _fixed_changed_flag caching, .cpp:37-44); the LAST call's list is returned
static int32_t find_impl(const orc_params* prm, const orc_point* fixed, int32_t n_fixed, const orc_point* moving,
                         int32_t n_moving, const float* xyt, int32_t stride, int32_t n_calls, int32_t* fixed_idx,
                         int32_t* moving_idx) {
  PointNormal2fVectorCloud f, m;
  to_cloud(fixed, n_fixed, f);
  to_cloud(moving, n_moving, m);
  CorrespondenceFinderProjective2f cf;
  configure(*cf.param_projector.value(), prm);
  cf.param_point_distance.setValue(prm->point_distance);
  cf.param_normal_cos.setValue(prm->normal_cos);
  CorrespondenceVector corr;
  cf.setFixed(&f);
  cf.setMoving(&m);
  cf.setCorrespondences(&corr);
  for (int32_t k = 0; k < n_calls; ++k) {
    cf.setLocalMapInSensor(iso(xyt + stride * k, stride));
    cf.compute();
  }
  for (size_t i = 0; i < corr.size(); ++i) {
    fixed_idx[i]  = corr[i].fixed_idx;
    moving_idx[i] = corr[i].moving_idx;
  }
  return (int32_t) corr.size();
}

int32_t ref_find_correspondences(const orc_params* prm, const orc_point* fixed, int32_t n_fixed, const orc_point* moving,
                                 int32_t n_moving, const float* xyt, int32_t n_calls, int32_t* fixed_idx,
                                 int32_t* moving_idx) {
  return find_impl(prm, fixed, n_fixed, moving, n_moving, xyt, 3, n_calls, fixed_idx, moving_idx);
}
// the same with local_map_in_sensor handed over as the Isometry2f itself (4 floats per pose: tx, ty, c, s)
int32_t ref_find_correspondences_iso(const orc_params* prm, const orc_point* fixed, int32_t n_fixed,
                                     const orc_point* moving, int32_t n_moving, const float* iso4, int32_t n_calls,
                                     int32_t* fixed_idx, int32_t* moving_idx) {
  return find_impl(prm, fixed, n_fixed, moving, n_moving, iso4, 4, n_calls, fixed_idx, moving_idx);
}

// MergerProjective2D::compute(): scene must have room for n_scene + canvas_cols points; returns the new size
static int32_t merge_impl(const orc_params* prm, float merge_threshold, orc_point* scene, int32_t n_scene,
                          const orc_point* measurement, int32_t n_measurement, const float* measurement_in_scene_xyt,
                          int32_t stride) {
  PointNormal2fVectorCloud s, m;
  to_cloud(scene, n_scene, s);
  to_cloud(measurement, n_measurement, m);
  MergerProjective2D mg;
  configure(*mg.param_projector.value(), prm);
  mg.param_merge_threshold.setValue(merge_threshold);
  mg.setScene(&s);
  mg.setMeasurement(&m);
  mg.setMeasurementInScene(iso(measurement_in_scene_xyt, stride));
  mg.compute();
  from_cloud(s, scene);
  return (int32_t) s.size();
}
int32_t ref_merge(const orc_params* prm, float merge_threshold, orc_point* scene, int32_t n_scene,
                  const orc_point* measurement, int32_t n_measurement, const float* measurement_in_scene_xyt) {
  return merge_impl(prm, merge_threshold, scene, n_scene, measurement, n_measurement, measurement_in_scene_xyt, 3);
}
int32_t ref_merge_iso(const orc_params* prm, float merge_threshold, orc_point* scene, int32_t n_scene,
                      const orc_point* measurement, int32_t n_measurement, const float* measurement_in_scene_iso4) {
  return merge_impl(prm, merge_threshold, scene, n_scene, measurement, n_measurement, measurement_in_scene_iso4, 4);
}

// SceneClipperProjective2D::compute() (voxelize_resolution 0 in both shipped configurations; > 0 takes the
// voxelize branch, .cpp:36-48); out holds canvas_cols points; returns the count
static int32_t clip_impl(const orc_params* prm, const orc_point* scene, int32_t n_scene, const float* robot_in_local_map_xyt,
                         const float* sensor_in_robot_xyt, int32_t stride, float voxelize_resolution, orc_point* out) {
  PointNormal2fVectorCloud full, clipped;
  to_cloud(scene, n_scene, full);
  SceneClipperProjective2D cl;
  configure(*cl.param_projector.value(), prm);
  cl.param_voxelize_resolution.setValue(voxelize_resolution);
  cl.setFullScene(&full);
  cl.setClippedSceneInRobot(&clipped);
  cl.setRobotInLocalMap(iso(robot_in_local_map_xyt, stride));
  cl.setSensorInRobot(iso(sensor_in_robot_xyt, stride));
  cl.compute();
  from_cloud(clipped, out);
  return (int32_t) clipped.size();
}
int32_t ref_clip(const orc_params* prm, const orc_point* scene, int32_t n_scene, const float* robot_in_local_map_xyt,
                 const float* sensor_in_robot_xyt, float voxelize_resolution, orc_point* out) {
  return clip_impl(prm, scene, n_scene, robot_in_local_map_xyt, sensor_in_robot_xyt, 3, voxelize_resolution, out);
}
int32_t ref_clip_iso(const orc_params* prm, const orc_point* scene, int32_t n_scene, const float* robot_in_local_map_iso4,
                     const float* sensor_in_robot_iso4, float voxelize_resolution, orc_point* out) {
  return clip_impl(prm, scene, n_scene, robot_in_local_map_iso4, sensor_in_robot_iso4, 4, voxelize_resolution, out);
}

// RawDataPreprocessorProjective2D: setRawData(LaserMessage) + compute(), driven as
// apps/visual_test_correspondence_finder_projective_2d.cpp:62-66 does; out holds n_beams points; returns the count
int32_t ref_preprocess_scan(const orc_scan_params* sp, const float* ranges, int32_t n_beams, orc_point* out) {
  LaserMessagePtr msg(new LaserMessage);
  msg->topic = "/scan";
  msg->ranges.value().assign(ranges, ranges + n_beams);
  msg->range_min.setValue(sp->msg_range_min);
  msg->range_max.setValue(sp->msg_range_max);
  msg->angle_min.setValue(sp->angle_min);
  msg->angle_max.setValue(sp->angle_max);
  RawDataPreprocessorProjective2D pre;
  pre.param_range_min.setValue(sp->range_min);
  pre.param_range_max.setValue(sp->range_max);
  pre.param_voxelize_resolution.setValue(sp->voxelize_resolution);
  pre.param_normal_computator_sliding->param_normal_point_distance.setValue(sp->normal_point_distance);
  pre.param_normal_computator_sliding->param_normal_min_points.setValue(sp->normal_min_points);
  RawDataPreprocessorProjective2D::MeasurementType cloud;
  pre.setMeas(&cloud);
  if (!pre.setRawData(msg)) {
    return -1;
  }
  pre.compute();
  from_cloud(cloud, out);
  return (int32_t) cloud.size();
}

// mis-wiring must throw std::runtime_error (correspondence_finder_projective_2d.cpp:20-31): 1 = it did
int32_t ref_finder_throws_without_inputs(void) {
  CorrespondenceFinderProjective2f cf;
  try {
    cf.compute();
  } catch (const std::runtime_error&) {
    return 1;
  }
  return 0;
}
}
