// ls2d_modules.cpp -- host logic of the CUDA-backed drop-in modules (see ls2d_modules.h).
// Parameter gathering, argument checks with the reference's exception messages, staging of clouds, and
// calls into the C ABI.  No arithmetic of the path is done here.
#include "ls2d_modules.h"

#include <cfloat>
#include <cstdlib>
#include <cstring>
#include <iostream>

namespace srrg2_core {

  Ls2dDevice::~Ls2dDevice() {
    if (_h) ls2d_destroy(_h);
  }

  void Ls2dDevice::check(int rc, const char* where) {
    if (rc != LS2D_OK) throw std::runtime_error(std::string(where) + "| " + ls2d_strerror(rc));
  }

  ls2d_handle* Ls2dDevice::handle() {
    if (!_h) {
      int device = 0;
      if (const char* d = std::getenv("LS2D_DEVICE")) device = std::atoi(d);
      check(ls2d_create(&_h, device), "Ls2dDevice::handle");
      // every pose crosses the C ABI as the Isometry2f itself (tx, ty, c, s): no t2v / v2t round trip, the device
      // sees the bits the caller holds (the reference hands Isometry2f objects to its modules)
      check(ls2d_set_pose_format(_h, LS2D_POSE_ISO), "Ls2dDevice::handle");
    }
    return _h;
  }

  void isoFloats(const Isometry2f& T, float* out) {
    out[0] = T.raw().tx, out[1] = T.raw().ty, out[2] = T.raw().c, out[3] = T.raw().s;
  }

  void flattenCloud(const PointNormal2fVectorCloud& cloud, std::vector<float>& out) {
    out.resize(cloud.size() * 4);
    for (size_t i = 0; i < cloud.size(); ++i) {
      const PointNormal2f& p = cloud[i];
      if (p.status == Valid) {
        out[4 * i + 0] = p.coordinates().x();
        out[4 * i + 1] = p.coordinates().y();
        out[4 * i + 2] = p.normal().x();
        out[4 * i + 3] = p.normal().y();
      } else {  // non-Valid points never project: encode them beyond any range_max
        out[4 * i + 0] = 1.0e6f;
        out[4 * i + 1] = out[4 * i + 2] = out[4 * i + 3] = 0.f;
      }
    }
  }

  void PointNormal2fProjectorPolar::cameraMatrix(float& K00, float& K01) const {
    const ls2d::polar_cam k = ls2d::make_polar_cam((int) param_canvas_cols.value(), param_angle_col_min.value(),
                                                   param_angle_col_max.value());
    K00 = k.K00, K01 = k.K01;
  }

  void PointNormal2fProjectorPolar::fillParams(ls2d_params& p) const {
    p.canvas_cols   = (int32_t) param_canvas_cols.value();
    p.angle_col_min = param_angle_col_min.value();
    p.angle_col_max = param_angle_col_max.value();
    p.range_min     = param_range_min.value();
    p.range_max     = param_range_max.value();
  }

  size_t PointNormal2fProjectorPolar::compute(TargetMatrixType& target, const PointNormal2f* begin,
                                              const PointNormal2f* end) {
    const size_t n    = (size_t)(end - begin);
    const size_t cols = param_canvas_cols.value();
    ls2d_params p;
    ls2d_default_params(&p);
    fillParams(p);
    ls2d_handle* h = _device.handle();
    Ls2dDevice::check(ls2d_set_params(h, &p), "PointNormal2fProjectorPolar::compute");
    PointNormal2fVectorCloud cloud(begin, end);
    std::vector<float> flat;
    flattenCloud(cloud, flat);
    const int32_t off[2] = {0, (int32_t) n};
    Ls2dDevice::check(ls2d_upload_clouds(h, LS2D_FIXED, flat.data(), off, 1), "PointNormal2fProjectorPolar::compute");
    float cam[4];
    isoFloats(_camera_pose, cam);
    std::vector<int32_t> idx(cols);
    std::vector<float> depth(cols);
    Ls2dDevice::check(ls2d_project(h, LS2D_FIXED, 0, cam, idx.data(), depth.data()),
                      "PointNormal2fProjectorPolar::compute");
    target.resize(1, cols);
    // `transformed` of the winners: same isometry, same single-rounding arithmetic as the device
    const Isometry2f W = _camera_pose.inverse();
    size_t filled      = 0;
    for (size_t c = 0; c < cols; ++c) {
      ProjectedEntry& e = target.at(0, c);
      e.source_idx      = idx[c];
      e.depth           = depth[c];
      if (idx[c] >= 0) {
        e.transformed.coordinates() = W * cloud[idx[c]].coordinates();
        e.transformed.normal()      = W.rotate(cloud[idx[c]].normal());
        ++filled;
      }
    }
    return filled;
  }

}  // namespace srrg2_core

namespace srrg2_solver {
  std::ostream& operator<<(std::ostream& os, const IterationStatsVector& stats) {
    for (const auto& s : stats)
      os << "it= " << s.iteration << "; chi_in= " << s.chi_inliers << "; chi_k= " << s.chi_kernelized
         << "; #in= " << s.num_inliers << "; #out= " << s.num_outliers << "; #corr= " << s.num_correspondences
         << "\n";
    return os;
  }
}  // namespace srrg2_solver

namespace srrg2_laser_slam_2d {

  void CorrespondenceFinderProjective2f::fillParams(ls2d_params& p) const {
    if (param_projector.value()) param_projector->fillParams(p);
    p.point_distance = param_point_distance.value();
    p.normal_cos     = param_normal_cos.value();
  }

  // same checks, same messages, same caching and the same side effect on the shared projector as
  // R/registration/correspondence_finder_projective_2d.cpp:18-77
  void CorrespondenceFinderProjective2f::compute() {
    PointNormal2fProjectorPolarPtr projector = param_projector.value();
    if (!projector) throw std::runtime_error("CorrespondenceFinderProjective2f::compute| Missing Projector");
    if (!_fixed) throw std::runtime_error("CorrespondenceFinderProjective2f::compute| Missing fixed!");
    if (!_moving) throw std::runtime_error("CorrespondenceFinderProjective2f::compute| Missing moving!");
    if (!_correspondences)
      throw std::runtime_error("CorrespondenceFinderProjective2f::compute| Missing correspondences!");
    const int num_beams = (int) projector->param_canvas_cols.value();
    ls2d_handle* h      = _device.handle();
    ls2d_params p;
    ls2d_default_params(&p);
    fillParams(p);
    Ls2dDevice::check(ls2d_set_params(h, &p), "CorrespondenceFinderProjective2f::compute");
    if (_fixed_changed_flag || _projector_changed_flag) {  // .cpp:37-44: the fixed image is cached
      flattenCloud(*_fixed, _staging);
      const int32_t off[2] = {0, (int32_t) _fixed->size()};
      Ls2dDevice::check(ls2d_upload_clouds(h, LS2D_FIXED, _staging.data(), off, 1),
                        "CorrespondenceFinderProjective2f::compute");
      _fixed_changed_flag     = false;
      _projector_changed_flag = false;
    }
    projector->setCameraPose(_local_map_in_sensor.inverse());  // .cpp:47 (visible to whoever shares the projector)
    flattenCloud(*_moving, _staging);
    const int32_t off[2] = {0, (int32_t) _moving->size()};
    Ls2dDevice::check(ls2d_upload_clouds(h, LS2D_MOVING, _staging.data(), off, 1),
                      "CorrespondenceFinderProjective2f::compute");
    std::vector<int32_t> fi(num_beams), mi(num_beams);
    int32_t n = 0;
    float lmis[4];
    isoFloats(_local_map_in_sensor, lmis);  // .cpp:47: the isometry as the caller set it, bit for bit
    Ls2dDevice::check(ls2d_find_correspondences(h, 0, 0, lmis, fi.data(), mi.data(), &n),
                      "CorrespondenceFinderProjective2f::compute");
    _correspondences->resize(n);
    for (int k = 0; k < n; ++k) (*_correspondences)[k] = Correspondence(fi[k], mi[k]);
  }

  static void unflatten(const float* flat, size_t n, PointNormal2fVectorCloud& cloud) {
    cloud.resize(n);
    for (size_t i = 0; i < n; ++i) {
      cloud[i].coordinates() = Vector2f(flat[4 * i], flat[4 * i + 1]);
      cloud[i].normal()      = Vector2f(flat[4 * i + 2], flat[4 * i + 3]);
      cloud[i].status        = Valid;
    }
  }

  // R/mapping/scene_clipper_projective_2d.cpp:11-65
  void SceneClipperProjective2D::compute() {
    if (!_clipped_scene_in_robot || !_full_scene) {
      _status = Error;
      std::cerr << "SceneClipperProjective2D::compute| missing local OR global scene" << std::endl;
      return;
    }
    if (!param_projector.value()) throw std::runtime_error("SceneClipperProjective2D::compute| Missing Projector");
    ls2d_params p;
    ls2d_default_params(&p);
    param_projector->fillParams(p);
    ls2d_handle* h = _device.handle();
    Ls2dDevice::check(ls2d_set_params(h, &p), "SceneClipperProjective2D::compute");
    std::vector<float> flat;
    flattenCloud(*_full_scene, flat);
    const int32_t off[2] = {0, (int32_t) _full_scene->size()};
    Ls2dDevice::check(ls2d_upload_clouds(h, LS2D_FIXED, flat.data(), off, 1), "SceneClipperProjective2D::compute");
    param_projector->setCameraPose(_robot_in_local_map * _sensor_in_robot);  // .cpp:29-31, visible to sharers
    float robot[4], sensor[4];
    isoFloats(_robot_in_local_map, robot);  // .cpp:22-32: both isometries verbatim
    isoFloats(_sensor_in_robot, sensor);
    std::vector<float> out((size_t) p.canvas_cols * 4);
    int32_t n        = 0;
    const int32_t id = 0;
    Ls2dDevice::check(ls2d_clip_scenes_voxelized(h, LS2D_FIXED, &id, robot, sensor, 1,
                                                 param_voxelize_resolution.value(), out.data(), &n),
                      "SceneClipperProjective2D::compute");  // <= 0: the plain clip (.cpp:49-57)
    unflatten(out.data(), (size_t) n, *_clipped_scene_in_robot);
    _status = Successful;
  }

  // R/mapping/merger_projective_2d.cpp:9-100
  void MergerProjective2D::compute() {
    if (!param_projector.value()) throw std::runtime_error("MergerProjective2D::compute| Missing Projector");
    if (!_scene || !_measurement) throw std::runtime_error("MergerProjective2D::compute| Missing scene or measurement");
    ls2d_params p;
    ls2d_default_params(&p);
    param_projector->fillParams(p);
    ls2d_handle* h = _device.handle();
    Ls2dDevice::check(ls2d_set_params(h, &p), "MergerProjective2D::compute");
    param_projector->setCameraPose(_measurement_in_scene);  // .cpp:19
    // only Valid scene points take part (projector and merger skip the others by status): they are compacted for the
    // device and written back through the index map, so a non-Valid point keeps its status, coordinates and slot
    std::vector<size_t> slot;
    std::vector<float> scene, meas;
    slot.reserve(_scene->size());
    scene.reserve((_scene->size() + p.canvas_cols) * 4);
    for (size_t i = 0; i < _scene->size(); ++i) {
      const PointNormal2f& q = (*_scene)[i];
      if (q.status != Valid) continue;
      slot.push_back(i);
      scene.insert(scene.end(), {q.coordinates().x(), q.coordinates().y(), q.normal().x(), q.normal().y()});
    }
    flattenCloud(*_measurement, meas);
    const int32_t n_valid  = (int32_t) slot.size();
    int32_t size           = n_valid;
    const int32_t capacity = size + p.canvas_cols;  // .cpp:31
    scene.resize((size_t) capacity * 4);
    float mis[4];
    isoFloats(_measurement_in_scene, mis);  // .cpp:19-22: the isometry verbatim
    Ls2dDevice::check(ls2d_merge_scene(h, scene.data(), &size, capacity, meas.data(), (int32_t) _measurement->size(),
                                       mis, param_merge_threshold.value(), nullptr),
                      "MergerProjective2D::compute");
    for (int32_t k = 0; k < size; ++k) {
      PointNormal2f q;
      q.coordinates() = Vector2f(scene[4 * (size_t) k], scene[4 * (size_t) k + 1]);
      q.normal()      = Vector2f(scene[4 * (size_t) k + 2], scene[4 * (size_t) k + 3]);
      q.status        = Valid;
      if (k < n_valid)
        (*_scene)[slot[k]] = q;   // merged / replaced in place (.cpp:72-84)
      else
        _scene->push_back(q);     // ordered append (.cpp:57-62, 87-88)
    }
    _status = Success;
  }

  void AlignerSliceProcessorLaser2DWithSensor::setupFactor() {
    (void) sensorInRobot();  // throws when the tf lookup fails, like setupFactorWithSensor
  }

  // ---- RawDataPreprocessorProjective2D (R/sensor_processing/raw_data_preprocessor_projective_2d.cpp)
  bool RawDataPreprocessorProjective2D::setRawData(BaseSensorMessagePtr msg) {
    if (!msg) throw std::runtime_error("RawDataPreprocessorProjective2D::setMeasurement|measurement is not set");  // .cpp:54-57
    _raw_data = msg;
    _status   = Error;
    // extractMessage<LaserMessage>(msg, scan_topic): the message itself when type and topic match (.cpp:62-63)
    LaserMessagePtr laser = std::dynamic_pointer_cast<LaserMessage>(msg);
    if (laser && laser->topic.value() != param_scan_topic.value()) laser.reset();
    if (!laser) {
      std::cerr << "RawDataPreprocessorProjective2D::setMeasurement|measurement does not contain a laser message"
                << std::endl;  // .cpp:64-68
      return false;
    }
    _processLaserMessage(laser);
    _status = Ready;
    return true;
  }

  void RawDataPreprocessorProjective2D::_processLaserMessage(LaserMessagePtr message) {
    _laser  = message;
    _ranges = &message->ranges.value();
    if (!param_unprojector.value())
      throw std::runtime_error("RawDataPreprocessorProjective2D::_processLaserMessage|missing unprojector");  // .cpp:94-97
    ls2d_scan_params sp;
    fillScanParams(sp);
    // .cpp:83-102: the unprojector takes the message's limits and the sensor matrix [1/res, n/2]
    PointNormal2fUnprojectorPolarPtr unprojector = param_unprojector.value();
    unprojector->param_range_min.setValue(sp.msg_range_min > sp.range_min ? sp.msg_range_min : sp.range_min);
    unprojector->param_range_max.setValue(sp.msg_range_max < sp.range_max ? sp.msg_range_max : sp.range_max);
    unprojector->param_angle_max.setValue(sp.angle_max);
    unprojector->param_angle_min.setValue(sp.angle_min);
    const float n = (float) _ranges->size();
    if (n > 0.f) unprojector->setCameraMatrix(1.f / ((sp.angle_max - sp.angle_min) / n), n / 2.f);
  }

  void RawDataPreprocessorProjective2D::fillScanParams(ls2d_scan_params& sp) const {
    ls2d_default_scan_params(&sp);
    if (_laser) {
      sp.angle_min     = _laser->angle_min.value();
      sp.angle_max     = _laser->angle_max.value();
      sp.msg_range_min = _laser->range_min.value();
      sp.msg_range_max = _laser->range_max.value();
    }
    sp.range_min           = param_range_min.value();
    sp.range_max           = param_range_max.value();
    sp.voxelize_resolution = param_voxelize_resolution.value();
    if (param_normal_computator_sliding.value()) {
      sp.normal_point_distance = param_normal_computator_sliding->param_normal_point_distance.value();
      sp.normal_min_points     = param_normal_computator_sliding->param_normal_min_points.value();
    }
  }

  void RawDataPreprocessorProjective2D::compute() {
    if (!_meas || !_raw_data) {  // .cpp:14-17
      _status = Error;
      return;
    }
    if (!param_unprojector.value())
      throw std::runtime_error("RawDataPreprocessorProjective2D::compute| missing unprojector");  // .cpp:19-21
    if (!param_normal_computator_sliding.value())
      throw std::runtime_error("RawDataPreprocessorProjective2D::compute| missing normal computator");
    if (!_ranges) {
      _status = Error;
      return;
    }
    _meas->clear();
    const int32_t n_beams = (int32_t) _ranges->size();
    if (n_beams > 0) {
      ls2d_scan_params sp;
      fillScanParams(sp);
      std::vector<float> out((size_t) n_beams * 4);
      int32_t n = 0;
      Ls2dDevice::check(ls2d_preprocess_scans(_device.handle(), &sp, _ranges->data(), n_beams, 1, out.data(), &n),
                        "RawDataPreprocessorProjective2D::compute");
      unflatten(out.data(), (size_t) n, *_meas);
    }
    _status = Ready;
  }

  void srrg2_laser_slam_2d_registerTypes() {
    using namespace srrg2_core;
    using namespace srrg2_solver;
    using namespace srrg2_slam_interfaces;
    // upstream *_registerTypes() of R/instances.cpp:21-25, reduced to the classes on the hot path
    BOSS_REGISTER_CLASS(PointNormal2fProjectorPolar);
    BOSS_REGISTER_CLASS(RobustifierCauchy);
    BOSS_REGISTER_CLASS(IterationAlgorithmGN);
    BOSS_REGISTER_CLASS(IterationAlgorithmLM);
    BOSS_REGISTER_CLASS(SparseBlockLinearSolverCholmodFull);
    BOSS_REGISTER_CLASS(SparseBlockLinearSolverCholeskyCSparse);
    BOSS_REGISTER_CLASS(SimpleTerminationCriteria);
    BOSS_REGISTER_CLASS(Solver);
    BOSS_REGISTER_CLASS(AlignerSliceOdom2DPrior);
    BOSS_REGISTER_CLASS(MultiAligner2D);
    BOSS_REGISTER_CLASS(MultiLoopDetectorBruteForce2D);
    // R/instances.cpp:27,34,35
    BOSS_REGISTER_CLASS(CorrespondenceFinderProjective2f);
    BOSS_REGISTER_CLASS(AlignerSliceProcessorLaser2D);
    BOSS_REGISTER_CLASS(AlignerSliceProcessorLaser2DWithSensor);
    // R/instances.cpp:30,32
    BOSS_REGISTER_CLASS(MergerProjective2D);
    BOSS_REGISTER_CLASS(SceneClipperProjective2D);
    // R/instances.cpp:28 and the srrg2_core classes its parameters point at
    BOSS_REGISTER_CLASS(RawDataPreprocessorProjective2D);
    BOSS_REGISTER_CLASS(PointNormal2fUnprojectorPolar);
    BOSS_REGISTER_CLASS(NormalComputator1DSlidingWindowNormal);
    // explicit CUDA names, for configurations that want to say so
    BOSS_REGISTER_CLASS_AS(CorrespondenceFinderProjective2f, "CorrespondenceFinderProjective2fCUDA");
    BOSS_REGISTER_CLASS_AS(AlignerSliceProcessorLaser2D, "AlignerSliceProcessorLaser2DCUDA");
    BOSS_REGISTER_CLASS_AS(AlignerSliceProcessorLaser2DWithSensor, "AlignerSliceProcessorLaser2DWithSensorCUDA");
    BOSS_REGISTER_CLASS_AS(MultiAligner2D, "MultiAligner2DCUDA");
    // slices that bind the point-to-point factor (not in the reference, which binds plane-to-plane only)
    BOSS_REGISTER_CLASS(AlignerSliceProcessorLaser2DPoint2Point);
    BOSS_REGISTER_CLASS(AlignerSliceProcessorLaser2DPoint2PointWithSensor);
  }

}  // namespace srrg2_laser_slam_2d

namespace srrg2_slam_interfaces {
  using srrg2_laser_slam_2d::CorrespondenceFinderProjective2f;

  Isometry2f AlignerSliceProcessorLaserBase::sensorInRobot() const {
    if (!withSensor()) return Isometry2f::Identity();
    if (!_platform) throw std::runtime_error("AlignerSliceProcessor::setupFactorWithSensor| no platform set");
    Isometry2f S;
    if (!_platform->getTransform(S, param_frame_id.value(), param_base_frame_id.value()))
      throw std::runtime_error("AlignerSliceProcessor::setupFactorWithSensor| unable to find transform [" +
                               param_frame_id.value() + "] in [" + param_base_frame_id.value() + "]");
    return S;
  }

  std::vector<std::shared_ptr<AlignerSliceProcessorLaserBase>> MultiAligner2D::laserSlices() {
    std::vector<std::shared_ptr<AlignerSliceProcessorLaserBase>> out;
    for (size_t i = 0; i < param_slice_processors.size(); ++i) {
      const AlignerSliceProcessorBasePtr& s = param_slice_processors.value(i);
      if (!s) continue;
      if (s->isLaserSlice()) {
        out.push_back(std::dynamic_pointer_cast<AlignerSliceProcessorLaserBase>(s));
      } else if (!std::dynamic_pointer_cast<AlignerSliceOdom2DPrior>(s)) {
        throw std::runtime_error("MultiAligner2D::compute| slice \"" + s->className() +
                                 "\" is not supported by the CUDA aligner");
      }
    }
    if (out.empty()) throw std::runtime_error("MultiAligner2D::compute| no slice processor set");
    if (out.size() > LS2D_MAX_SLICES)
      throw std::runtime_error("MultiAligner2D::compute| more than " + std::to_string(LS2D_MAX_SLICES) +
                               " laser slices");
    return out;
  }

  std::shared_ptr<AlignerSliceProcessorLaserBase> MultiAligner2D::laserSlice() { return laserSlices().front(); }

  // the odometry prior takes part only if its pose slice is present in both scenes (a scene without odometry,
  // e.g. the loop detector's local maps, simply drops the factor)
  std::shared_ptr<AlignerSliceOdom2DPrior> MultiAligner2D::boundPrior() {
    std::shared_ptr<AlignerSliceOdom2DPrior> prior;
    for (size_t i = 0; i < param_slice_processors.size(); ++i) {
      auto p = std::dynamic_pointer_cast<AlignerSliceOdom2DPrior>(param_slice_processors.value(i));
      if (!p || !p->bound(_fixed_scene, _moving_scene)) continue;
      if (prior) throw std::runtime_error("MultiAligner2D::compute| more than one bound prior slice");
      prior = p;
    }
    return prior;
  }

  void MultiAligner2D::fillOne(ls2d_params& p, const std::shared_ptr<AlignerSliceProcessorLaserBase>& slice) {
    ls2d_default_params(&p);
    auto finder = std::dynamic_pointer_cast<CorrespondenceFinderProjective2f>(slice->param_finder.value());
    if (!finder)
      throw std::runtime_error("MultiAligner2D::compute| the CUDA aligner needs a CorrespondenceFinderProjective2f "
                               "finder on its laser slice");
    if (!finder->param_projector.value())
      throw std::runtime_error("CorrespondenceFinderProjective2f::compute| Missing Projector");
    finder->fillParams(p);
    auto cauchy = std::dynamic_pointer_cast<srrg2_solver::RobustifierCauchy>(slice->param_robustifier.value());
    if (slice->param_robustifier.value() && !cauchy)
      throw std::runtime_error("MultiAligner2D::compute| only RobustifierCauchy is supported");
    p.cauchy_chi_threshold    = cauchy ? cauchy->param_chi_threshold.value() : -1.f;
    p.min_num_correspondences = slice->param_min_num_correspondences.value();
    p.max_iterations          = param_max_iterations.value();
    p.min_num_inliers         = param_min_num_inliers.value();
    p.damping                 = 0.f;
    p.factor                  = slice->pointToPoint() ? LS2D_FACTOR_POINT2POINT : LS2D_FACTOR_PLANE2PLANE;
    if (auto solver = param_solver.value()) {
      if (solver->param_max_iterations.size() && solver->param_max_iterations.value(0) != 1)
        throw std::runtime_error("MultiAligner2D::compute| the inner solver must run 1 iteration per ICP round "
                                 "(both shipped configurations do)");
      if (auto gn = std::dynamic_pointer_cast<srrg2_solver::IterationAlgorithmGN>(solver->param_algorithm.value())) {
        p.damping = gn->param_damping.value();
      } else if (auto lm = std::dynamic_pointer_cast<srrg2_solver::IterationAlgorithmLM>(solver->param_algorithm.value())) {
        p.algorithm           = LS2D_ALGORITHM_LM;
        p.lm_user_lambda_init = lm->param_user_lambda_init.value();
        p.lm_tau              = lm->param_tau.value();
        p.lm_step_low         = lm->param_step_low.value();
        p.lm_step_high        = lm->param_step_high.value();
        p.lm_iterations_max   = lm->param_lm_iterations_max.value();
        p.lm_variable_damping = lm->param_variable_damping.value() ? 1 : 0;
      } else if (solver->param_algorithm.value()) {
        throw std::runtime_error("MultiAligner2D::compute| only IterationAlgorithmGN / IterationAlgorithmLM are supported");
      }
    }
    p.enable_inlier_only_runs          = param_enable_inlier_only_runs.value() ? 1 : 0;
    p.keep_only_inlier_correspondences = param_keep_only_inlier_correspondences.value() ? 1 : 0;
    if (auto tc = param_termination_criteria.value()) {
      auto simple = std::dynamic_pointer_cast<srrg2_solver::SimpleTerminationCriteria>(tc);
      if (!simple) throw std::runtime_error("MultiAligner2D::compute| only SimpleTerminationCriteria is supported");
      p.termination_epsilon = simple->param_epsilon.value();
    }
    p.with_sensor = slice->withSensor() ? 2 : 0;  // 2: sensor_in_robot handed over as the isometry itself
    if (p.with_sensor) {
      slice->setupFactor();
      const Isometry2f S = slice->sensorInRobot();
      p.sensor_in_robot[0] = S.raw().tx, p.sensor_in_robot[1] = S.raw().ty, p.sensor_in_robot[2] = 0.f;
      p.sensor_in_robot_cs[0] = S.raw().c, p.sensor_in_robot_cs[1] = S.raw().s;
    }
  }

  void MultiAligner2D::fillParams(ls2d_params& p) { fillOne(p, laserSlice()); }

  void MultiAligner2D::fillSliceParams(std::vector<ls2d_params>& p) {
    const auto slices = laserSlices();
    p.resize(slices.size());
    for (size_t i = 0; i < slices.size(); ++i) fillOne(p[i], slices[i]);
  }

  static Isometry2f isoOf(float tx, float ty, float c, float s) {
    ls2d::iso T;
    T.tx = tx, T.ty = ty, T.c = c, T.s = s;
    return Isometry2f(T);
  }

  static AlignmentResult toResult(const ls2d_result& r) {
    AlignmentResult a;
    a.estimate        = Vector3f(r.x, r.y, r.theta);
    a.moving_in_fixed = isoOf(r.x, r.y, r.c, r.s);  // the isometry the kernel holds, not v2t(t2v(.))
    a.status          = r.status;
    a.iterations      = r.iterations;
    a.last.iteration           = r.iterations - 1;
    a.last.chi_inliers         = r.chi_inliers;
    a.last.chi_kernelized      = r.chi_kernelized;
    a.last.num_inliers         = r.n_inliers;
    a.last.num_outliers        = r.n_kernelized;
    a.last.num_correspondences = r.n_corr;
    a.last.estimate            = a.estimate;
    const int map[3][3] = {{0, 1, 2}, {1, 3, 4}, {2, 4, 5}};
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) a.information_matrix.m[i][j] = r.H[map[i][j]];
    return a;
  }

  static void uploadSet(ls2d_handle* h, int which, const std::vector<const PointNormal2fVectorCloud*>& clouds,
                        const char* where) {
    std::vector<int32_t> off(clouds.size() + 1, 0);
    for (size_t i = 0; i < clouds.size(); ++i) {
      if (!clouds[i]) throw std::runtime_error(std::string(where) + "| null cloud in batch");
      off[i + 1] = off[i] + (int32_t) clouds[i]->size();
    }
    std::vector<float> flat((size_t) off.back() * 4), one;
    for (size_t i = 0; i < clouds.size(); ++i) {
      flattenCloud(*clouds[i], one);
      if (!one.empty()) std::memcpy(flat.data() + (size_t) off[i] * 4, one.data(), one.size() * sizeof(float));
    }
    Ls2dDevice::check(ls2d_upload_clouds(h, which, flat.data(), off.data(), (int32_t) clouds.size()), where);
  }

  void MultiAligner2D::storeOutcome(const ls2d_result& r, const std::vector<ls2d_iter_stats>& its) {
    const AlignmentResult a = toResult(r);
    _moving_in_fixed        = a.moving_in_fixed;
    _information_matrix     = a.information_matrix;
    _status                 = r.status == LS2D_STATUS_SUCCESS ? Success
                              : r.status == LS2D_STATUS_NOT_ENOUGH_CORRESPONDENCES ? NotEnoughCorrespondences
                              : r.status == LS2D_STATUS_NOT_ENOUGH_INLIERS ? NotEnoughInliers : Fail;
    for (int i = 0; i < r.iterations; ++i) {
      srrg2_solver::IterationStats s;
      s.iteration           = i;
      s.chi_inliers         = its[i].chi_inliers;
      s.chi_kernelized      = its[i].chi_kernelized;
      s.num_inliers         = its[i].n_inliers;
      s.num_outliers        = its[i].n_kernelized;
      s.num_correspondences = its[i].n_corr;
      s.estimate            = Vector3f(its[i].x, its[i].y, its[i].theta);
      _iteration_stats.push_back(s);
    }
  }

  // the estimate the LAST executed finder pass ran at: a completed run linearised its last round at the estimate
  // after round iterations - 2; a run that stopped early in round k (NotEnoughCorrespondences / singular system)
  // completed k = iterations rounds and its failing finder pass ran at the estimate after round k - 1
  static Isometry2f lastFinderEstimate(const ls2d_result& r, const std::vector<ls2d_iter_stats>& its,
                                       const Isometry2f& init) {
    const bool early = r.status == LS2D_STATUS_NOT_ENOUGH_CORRESPONDENCES || r.status == LS2D_STATUS_SINGULAR;
    const int k      = early ? r.iterations - 1 : r.iterations - 2;
    if (k < 0) return init;
    return isoOf(its[k].x, its[k].y, its[k].c, its[k].s);
  }

  // slice->correspondences() after compute(): the last finder pass, re-run at the very isometry the kernel used
  // (sensor_in_robot^-1 * estimate composed with the same binary32 sequence); keep_only_inlier_correspondences
  // drops the pairs whose factor was kernelized
  void MultiAligner2D::exportCorrespondences(ls2d_handle* h, const ls2d_params& p, int fixed_set, int moving_set,
                                             const std::shared_ptr<AlignerSliceProcessorLaserBase>& slice,
                                             const Isometry2f& estimate) {
    Isometry2f lmis = estimate;
    if (p.with_sensor) lmis = slice->sensorInRobot().inverse() * lmis;
    AlignerSliceProcessorLaserBase* s = slice.get();  // the closure lives in the slice: no owning pointer back to it
    s->_correspondences.clear();
    const std::weak_ptr<int> alive = _alive;  // the handle is this aligner's: a fetch after its death finds nothing
    s->_fetch_correspondences = [h, p, fixed_set, moving_set, s, lmis, estimate, alive]() {
      if (alive.expired()) return;
      float lv[4], ev[4];
      isoFloats(lmis, lv);
      isoFloats(estimate, ev);
      Ls2dDevice::check(ls2d_set_params(h, &p), "MultiAligner2D::compute");  // a multi-slice aligner leaves the last slice's
      std::vector<int32_t> fi(p.canvas_cols), mi(p.canvas_cols);
      int32_t n = 0;
      Ls2dDevice::check(ls2d_find_correspondences_in(h, fixed_set, moving_set, 0, 0, lv, fi.data(), mi.data(), &n),
                        "MultiAligner2D::compute");
      std::vector<uint8_t> inlier((size_t) (n > 0 ? n : 1), 1);
      if (p.keep_only_inlier_correspondences && n > 0)
        Ls2dDevice::check(ls2d_classify_correspondences(h, fixed_set, moving_set, 0, 0, ev, fi.data(), mi.data(), n,
                                                        inlier.data()),
                          "MultiAligner2D::compute");
      for (int k = 0; k < n; ++k)
        if (inlier[k]) s->_correspondences.push_back(Correspondence(fi[k], mi[k]));
    };
  }

  static PointNormal2fVectorCloud* sliceCloud(PropertyContainerDynamic* scene, const std::string& name, const char* which) {
    PointNormal2fVectorCloud* c = scene->cloud(name);
    if (!c) throw std::runtime_error(std::string("MultiAligner2D::compute| ") + which + " scene has no slice \"" + name + "\"");
    return c;
  }

  void MultiAligner2D::compute() {
    _status = Fail;
    _iteration_stats.clear();
    if (!_fixed_scene) throw std::runtime_error("MultiAligner2D::compute| Missing fixed!");
    if (!_moving_scene) throw std::runtime_error("MultiAligner2D::compute| Missing moving!");
    const auto slices = laserSlices();
    const auto prior  = boundPrior();
    if (slices.size() > 1 || prior) return computeMulti(slices, prior);
    ls2d_params p;
    fillParams(p);
    std::shared_ptr<AlignerSliceProcessorLaserBase> slice = slices.front();
    PointNormal2fVectorCloud* fixed  = sliceCloud(_fixed_scene, slice->param_fixed_slice_name.value(), "fixed");
    PointNormal2fVectorCloud* moving = sliceCloud(_moving_scene, slice->param_moving_slice_name.value(), "moving");
    slice->_fixed  = fixed;
    slice->_moving = moving;
    ls2d_handle* h = _device.handle();
    Ls2dDevice::check(ls2d_set_params(h, &p), "MultiAligner2D::compute");
    uploadSet(h, LS2D_FIXED, {fixed}, "MultiAligner2D::compute");
    uploadSet(h, LS2D_MOVING, {moving}, "MultiAligner2D::compute");
    const Isometry2f init = _moving_in_fixed;
    float iv[4];
    isoFloats(init, iv);  // apps/visual_test_aligner_2d.cpp:123-128: setMovingInFixed(Isometry2f), verbatim
    ls2d_result r;
    const int n_its = p.max_iterations * (p.enable_inlier_only_runs ? 2 : 1);
    std::vector<ls2d_iter_stats> its((size_t) (n_its > 0 ? n_its : 1));
    Ls2dDevice::check(ls2d_align_batch(h, nullptr, nullptr, iv, 1, &r, its.data()), "MultiAligner2D::compute");
    storeOutcome(r, its);
    exportCorrespondences(h, p, LS2D_FIXED, LS2D_MOVING, slice, lastFinderEstimate(r, its, init));
  }

  // several laser slices and / or a bound odometry prior: one fused 3x3 system per iteration (MULTI.json:700-730,
  // LASER_0.json:502-506).  Fixed cloud of slice s -> cloud set s, its moving cloud -> set LS2D_MAX_SLICES + s
  // (slices that share a moving cloud share the set).
  void MultiAligner2D::computeMulti(const std::vector<std::shared_ptr<AlignerSliceProcessorLaserBase>>& slices,
                                    const std::shared_ptr<AlignerSliceOdom2DPrior>& prior) {
    std::vector<ls2d_params> p;
    fillSliceParams(p);
    ls2d_handle* h = _device.handle();
    std::vector<int32_t> fset(slices.size()), mset(slices.size());
    for (size_t s = 0; s < slices.size(); ++s) {
      slices[s]->_fixed  = sliceCloud(_fixed_scene, slices[s]->param_fixed_slice_name.value(), "fixed");
      slices[s]->_moving = sliceCloud(_moving_scene, slices[s]->param_moving_slice_name.value(), "moving");
      fset[s] = (int32_t) s;
      mset[s] = LS2D_MAX_SLICES + (int32_t) s;
      uploadSet(h, fset[s], {slices[s]->_fixed}, "MultiAligner2D::compute");
      bool shared = false;
      for (size_t t = 0; t < s && !shared; ++t)
        if (slices[t]->_moving == slices[s]->_moving) mset[s] = mset[t], shared = true;
      if (!shared) uploadSet(h, mset[s], {slices[s]->_moving}, "MultiAligner2D::compute");
    }
    ls2d_prior pr;
    float z[4];
    if (prior) {
      const Matrix3f& O = prior->informationMatrix();
      const float info[6] = {O.m[0][0], O.m[0][1], O.m[0][2], O.m[1][1], O.m[1][2], O.m[2][2]};
      std::memcpy(pr.information, info, sizeof(info));
      auto cauchy = std::dynamic_pointer_cast<srrg2_solver::RobustifierCauchy>(prior->param_robustifier.value());
      if (prior->param_robustifier.value() && !cauchy)
        throw std::runtime_error("MultiAligner2D::compute| only RobustifierCauchy is supported");
      pr.cauchy_chi_threshold = cauchy ? cauchy->param_chi_threshold.value() : -1.f;
      isoFloats(prior->measurement(_fixed_scene, _moving_scene), z);
    }
    const Isometry2f init = _moving_in_fixed;
    float iv[4];
    isoFloats(init, iv);
    ls2d_result r;
    std::vector<ls2d_iter_stats> its((size_t) (p[0].max_iterations > 0 ? p[0].max_iterations : 1));
    Ls2dDevice::check(ls2d_align_multi(h, p.data(), fset.data(), mset.data(), (int32_t) slices.size(), prior ? &pr : nullptr,
                                       prior ? z : nullptr, nullptr, nullptr, iv, 1, &r, its.data()),
                      "MultiAligner2D::compute");
    storeOutcome(r, its);
    // every slice's correspondences(): the last finder pass
    const Isometry2f before = lastFinderEstimate(r, its, init);
    for (size_t s = 0; s < slices.size(); ++s) exportCorrespondences(h, p[s], fset[s], mset[s], slices[s], before);
  }

  void MultiAligner2D::computeBatch(const std::vector<const PointNormal2fVectorCloud*>& fixed,
                                    const std::vector<const PointNormal2fVectorCloud*>& moving,
                                    const std::vector<Isometry2f>& guesses, std::vector<AlignmentResult>& results) {
    if (fixed.size() != moving.size() || fixed.size() != guesses.size())
      throw std::runtime_error("MultiAligner2D::computeBatch| fixed, moving and guesses differ in size");
    ls2d_params p;
    fillParams(p);
    ls2d_handle* h = _device.handle();
    Ls2dDevice::check(ls2d_set_params(h, &p), "MultiAligner2D::computeBatch");
    uploadSet(h, LS2D_FIXED, fixed, "MultiAligner2D::computeBatch");
    uploadSet(h, LS2D_MOVING, moving, "MultiAligner2D::computeBatch");
    std::vector<float> init(guesses.size() * 4);
    for (size_t i = 0; i < guesses.size(); ++i) isoFloats(guesses[i], &init[4 * i]);
    std::vector<ls2d_result> out(guesses.size());
    Ls2dDevice::check(ls2d_align_batch(h, nullptr, nullptr, init.data(), (int32_t) guesses.size(), out.data(), nullptr),
                      "MultiAligner2D::computeBatch");
    results.clear();
    for (const auto& r : out) results.push_back(toResult(r));
  }

  LoopClosure2D MultiLoopDetectorBruteForce2D::verify(const PointNormal2fVectorCloud& query,
                                                      const std::vector<const PointNormal2fVectorCloud*>& candidates,
                                                      const std::vector<std::vector<Isometry2f>>& guesses,
                                                      std::vector<AlignmentResult>* all) {
    MultiAligner2DPtr aligner = param_relocalize_aligner.value();
    if (!aligner) throw std::runtime_error("MultiLoopDetectorBruteForce2D::compute| relocalize_aligner not set");
    if (candidates.size() != guesses.size() || candidates.empty())
      throw std::runtime_error("MultiLoopDetectorBruteForce2D::compute| candidates and guesses differ in size");
    const size_t n_guess = guesses[0].size();
    for (const auto& g : guesses)
      if (g.size() != n_guess || n_guess == 0)
        throw std::runtime_error("MultiLoopDetectorBruteForce2D::compute| every candidate needs the same, non-zero, "
                                 "number of initial guesses");
    ls2d_params p;
    aligner->fillParams(p);
    ls2d_handle* h = aligner->device().handle();
    Ls2dDevice::check(ls2d_set_params(h, &p), "MultiLoopDetectorBruteForce2D::compute");
    uploadSet(h, LS2D_FIXED, {&query}, "MultiLoopDetectorBruteForce2D::compute");
    uploadSet(h, LS2D_MOVING, candidates, "MultiLoopDetectorBruteForce2D::compute");
    std::vector<float> init(candidates.size() * n_guess * 4);
    for (size_t c = 0; c < candidates.size(); ++c)
      for (size_t g = 0; g < n_guess; ++g) isoFloats(guesses[c][g], &init[4 * (c * n_guess + g)]);
    ls2d_gates gates;
    gates.min_inliers        = param_relocalize_min_inliers.value();
    gates.max_chi_per_inlier = param_relocalize_max_chi_inliers.value();
    gates.min_inlier_ratio   = param_relocalize_min_inliers_ratio.value();
    ls2d_best best;
    std::vector<ls2d_result> out(all ? candidates.size() * n_guess : 0);
    Ls2dDevice::check(ls2d_verify(h, 0, nullptr, (int32_t) candidates.size(), init.data(), (int32_t) n_guess, &gates, 0,
                                  &best, all ? out.data() : nullptr),
                      "MultiLoopDetectorBruteForce2D::compute");
    if (all) {
      all->clear();
      for (const auto& r : out) all->push_back(toResult(r));
    }
    LoopClosure2D lc;
    lc.candidate = best.candidate;
    lc.guess     = best.guess;
    if (best.candidate >= 0) {
      lc.moving_in_fixed     = isoOf(best.x, best.y, best.c, best.s);
      lc.chi_inliers         = best.chi_inliers;
      lc.num_inliers         = best.n_inliers;
      lc.num_correspondences = best.n_corr;
    }
    return lc;
  }

}  // namespace srrg2_slam_interfaces
