// ls2d_modules.h -- CUDA-backed drop-ins for the reference's hot-path modules, with the reference's class
// names, PARAM names, method names, ownership and error behaviour, so that the reference's configuration
// files instantiate THESE classes unchanged.  Every compute() forwards to the C ABI of include/ls2d.h;
// there is no host implementation of the path in here.
//
// Reference interfaces mirrored (R/ = /root/reference/srrg2_laser_slam_2d/src/srrg2_laser_slam_2d/):
//   PointNormal2fProjectorPolar            params L0.json:312-338; use R/registration/correspondence_finder_projective_2d.cpp:35-48
//   CorrespondenceFinderProjective2f       R/registration/correspondence_finder_projective_2d.{h,cpp}
//   AlignerSliceProcessorLaser2D[WithSensor]  R/registration/aligner_slice_processor_laser_2d.h:7-45, _impl.cpp:7-10
//   MultiAligner2D                         driven at apps/visual_test_aligner_2d.cpp:123-156; params L0.json:9-37
//   RobustifierCauchy / IterationAlgorithmGN / Solver / ...   parameter holders, L0.json:76-88,193-289
//   MultiLoopDetectorBruteForce2D          gates L0.json:613-635
#pragma once

#include <functional>
#include <memory>
#include <string>
#include <vector>

#include "../../include/ls2d.h"
#include "ls2d_boss.h"
#include "ls2d_types.h"

namespace srrg2_core {

  // owns one ls2d_handle; created on first use (device from $LS2D_DEVICE, default 0); throws
  // std::runtime_error when the CUDA library reports an error -- there is no CPU fallback
  class Ls2dDevice {
  public:
    Ls2dDevice() {}
    ~Ls2dDevice();
    Ls2dDevice(const Ls2dDevice&)            = delete;
    Ls2dDevice& operator=(const Ls2dDevice&) = delete;
    ls2d_handle* handle();
    static void check(int rc, const char* where);

  private:
    ls2d_handle* _h = nullptr;
  };

  // flat (x, y, nx, ny) staging of a cloud for ls2d_upload_clouds
  void flattenCloud(const PointNormal2fVectorCloud& cloud, std::vector<float>& out);
  // (tx, ty, c, s): an Isometry2f in the LS2D_POSE_ISO wire format of include/ls2d.h -- the bits the caller holds
  void isoFloats(const Isometry2f& T, float* out);

  // 1 x cols image of {source_idx, depth, transformed}  (PointNormal2fProjectorPolar::TargetMatrixType)
  struct ProjectedEntry {
    int source_idx = -1;
    float depth    = 0.f;
    PointNormal2f transformed;
  };
  class ProjectedMatrix {
  public:
    void resize(size_t rows, size_t cols) {
      _rows = rows, _cols = cols;
      _data.assign(rows * cols, ProjectedEntry());
    }
    void clear() { _data.clear(), _rows = _cols = 0; }
    size_t rows() const { return _rows; }
    size_t cols() const { return _cols; }
    ProjectedEntry& at(size_t r, size_t c) { return _data.at(r * _cols + c); }
    const ProjectedEntry& at(size_t r, size_t c) const { return _data.at(r * _cols + c); }
    std::vector<ProjectedEntry>::iterator begin() { return _data.begin(); }
    std::vector<ProjectedEntry>::iterator end() { return _data.end(); }

  private:
    size_t _rows = 0, _cols = 0;
    std::vector<ProjectedEntry> _data;
  };

  class PointNormal2fProjectorPolar : public Configurable {
  public:
    using TargetMatrixType = ProjectedMatrix;
    PARAM(PropertyFloat, angle_col_max, "end col angle    [rad]", 3.14159f, &_config_changed);
    PARAM(PropertyFloat, angle_col_min, "start col angle  [rad]", -3.14159f, &_config_changed);
    PARAM(PropertyFloat, angle_row_max, "end row angle    [rad]", 1.5708f, 0);
    PARAM(PropertyFloat, angle_row_min, "start row angle  [rad]", -1.5708f, 0);
    PARAM(PropertyUnsignedInt, canvas_cols, "cols of the canvas", 721, &_config_changed);
    PARAM(PropertyUnsignedInt, canvas_rows, "rows of the canvas", 1, 0);
    PARAM(PropertyFloat, range_max, "maximum range [m]", 20.f, &_config_changed);
    PARAM(PropertyFloat, range_min, "minimum range [m]", 0.3f, &_config_changed);
    PointNormal2fProjectorPolar() { _class_name = "PointNormal2fProjectorPolar"; }

    void setCameraPose(const Isometry2f& camera_pose) { _camera_pose = camera_pose; }
    const Isometry2f& cameraPose() const { return _camera_pose; }
    void initCameraMatrix() {}
    // [[K00, K01], [0, 1]] with K00 = cols / (angle_col_max - angle_col_min), K01 = cols / 2
    void cameraMatrix(float& K00, float& K01) const;
    // project [begin, end) into `target` (resized to 1 x canvas_cols); returns the number of filled cells
    size_t compute(TargetMatrixType& target, const PointNormal2f* begin, const PointNormal2f* end);
    size_t compute(TargetMatrixType& target, PointNormal2fVectorCloud::const_iterator begin,
                   PointNormal2fVectorCloud::const_iterator end) {
      return compute(target, begin == end ? nullptr : &*begin, begin == end ? nullptr : &*begin + (end - begin));
    }
    void fillParams(ls2d_params& p) const;

  protected:
    bool _config_changed = true;
    Isometry2f _camera_pose;
    Ls2dDevice _device;
  };
  using PointNormal2fProjectorPolarPtr = std::shared_ptr<PointNormal2fProjectorPolar>;

  // srrg2_core::LaserMessage, reduced to the fields RawDataPreprocessorProjective2D reads
  // (R/sensor_processing/raw_data_preprocessor_projective_2d.cpp:78-86; built in tests/fixtures.hpp:25-35)
  template <typename T>
  struct MessageField {
    T _v = T();
    void setValue(const T& v) { _v = v; }
    const T& value() const { return _v; }
    T& value() { return _v; }
  };
  struct BaseSensorMessage {
    explicit BaseSensorMessage(const std::string& topic_ = "") { topic.setValue(topic_); }
    virtual ~BaseSensorMessage() {}
    MessageField<std::string> topic;
  };
  using BaseSensorMessagePtr = std::shared_ptr<BaseSensorMessage>;
  struct LaserMessage : public BaseSensorMessage {
    using BaseSensorMessage::BaseSensorMessage;
    MessageField<float> angle_min, angle_max, angle_increment, time_increment, scan_time, range_min, range_max;
    MessageField<std::vector<float>> ranges, intensities;
  };
  using LaserMessagePtr = std::shared_ptr<LaserMessage>;

  // parameter holder of PointNormal2fUnprojectorPolar (L0.json:405-434); the pre-processor overwrites its
  // range / angle limits from every message (raw_data_preprocessor_projective_2d.cpp:98-102)
  class PointNormal2fUnprojectorPolar : public Configurable {
  public:
    PARAM(PropertyFloat, angle_max, "end angle    [rad]", 3.14159f, 0);
    PARAM(PropertyFloat, angle_min, "start angle  [rad]", -3.14159f, 0);
    PARAM(PropertyUnsignedInt, canvas_cols, "cols of the canvas", 721, 0);
    PARAM(PropertyUnsignedInt, canvas_rows, "rows of the canvas", 1, 0);
    PARAM(PropertyInt, normal_min_points, "minimum number of points in ball when computing a valid normal", 5, 0);
    PARAM(PropertyFloat, normal_point_distance, "range of points considered while computing normal", 0.2f, 0);
    PARAM(PropertyInt, num_ranges, "number of laser beams", 721, 0);
    PARAM(PropertyFloat, range_max, "max laser range [m]", 20.f, 0);
    PARAM(PropertyFloat, range_min, "min laser range [m]", 0.3f, 0);
    PointNormal2fUnprojectorPolar() { _class_name = "PointNormal2fUnprojectorPolar"; }
    // sensor matrix [[fx, cx], [0, 0]] (.cpp:89-90)
    void setCameraMatrix(float fx, float cx) { _fx = fx, _cx = cx; }
    float fx() const { return _fx; }
    float cx() const { return _cx; }

  protected:
    float _fx = 1.f, _cx = 0.f;
  };
  using PointNormal2fUnprojectorPolarPtr = std::shared_ptr<PointNormal2fUnprojectorPolar>;

  // parameter holder of NormalComputator1DSlidingWindow<PointNormal2fVectorCloud, 1> (L0.json:711-719)
  class NormalComputator1DSlidingWindowNormal : public Configurable {
  public:
    PARAM(PropertyInt, normal_min_points, "min number of points to compute a normal", 5, 0);
    PARAM(PropertyFloat, normal_point_distance, "max normal point distance", 0.3f, 0);
    NormalComputator1DSlidingWindowNormal() { _class_name = "NormalComputator1DSlidingWindowNormal"; }
  };
  using NormalComputator1DSlidingWindowNormalPtr = std::shared_ptr<NormalComputator1DSlidingWindowNormal>;

}  // namespace srrg2_core

namespace srrg2_solver {
  using namespace srrg2_core;

  class RobustifierBase : public Configurable {};
  class RobustifierCauchy : public RobustifierBase {
  public:
    PARAM(PropertyFloat, chi_threshold, "threshold of chi after which the kernel is active", 1.f, 0);
    RobustifierCauchy() { _class_name = "RobustifierCauchy"; }
  };
  class IterationAlgorithmBase : public Configurable {};
  class IterationAlgorithmGN : public IterationAlgorithmBase {
  public:
    PARAM(PropertyFloat, damping, "damping factor, the higher the closer to gradient descend. Default:0", 0.f, 0);
    IterationAlgorithmGN() { _class_name = "IterationAlgorithmGN"; }
  };
  // IterationAlgorithmLM of srrg2_solver (not instantiated by the shipped configurations; BASELINE.json's north_star
  // names the LM update): parameter holder, run by icp_general_kernel (oracle decisions L1..L8)
  class IterationAlgorithmLM : public IterationAlgorithmBase {
  public:
    PARAM(PropertyFloat, user_lambda_init, "initial lm lambda, if 0 is computed by system", 0.f, 0);
    PARAM(PropertyFloat, step_high, "upper clamp for lambda if things go well", 2.f / 3.f, 0);
    PARAM(PropertyFloat, step_low, "lower clamp for lambda if things go well", 1.f / 3.f, 0);
    PARAM(PropertyInt, lm_iterations_max, "max lm iterations", 10, 0);
    PARAM(PropertyBool, variable_damping, "set to true uses lambda*diag(H), otherwise uses lambda*I", true, 0);
    PARAM(PropertyFloat, tau, "scale factor for the lambda computed by the system", 1e-5f, 0);
    IterationAlgorithmLM() { _class_name = "IterationAlgorithmLM"; }
  };
  class SparseBlockLinearSolver : public Configurable {};
  class SparseBlockLinearSolverCholmodFull : public SparseBlockLinearSolver {
  public:
    SparseBlockLinearSolverCholmodFull() { _class_name = "SparseBlockLinearSolverCholmodFull"; }
  };
  class SparseBlockLinearSolverCholeskyCSparse : public SparseBlockLinearSolver {
  public:
    SparseBlockLinearSolverCholeskyCSparse() { _class_name = "SparseBlockLinearSolverCholeskyCSparse"; }
  };
  class TerminationCriteria : public Configurable {};
  class SimpleTerminationCriteria : public TerminationCriteria {
  public:
    PARAM(PropertyFloat, epsilon, "ratio of decay of chi2 between iteration", 1e-3f, 0);
    SimpleTerminationCriteria() { _class_name = "SimpleTerminationCriteria"; }
  };
  class Solver : public Configurable {
  public:
    PARAM(PropertyConfigurable_<IterationAlgorithmBase>, algorithm,
          "pointer to the optimization algorithm (GN/LM or others)", nullptr, 0);
    PARAM(PropertyConfigurable_<SparseBlockLinearSolver>, linear_solver,
          "pointer to linear solver used to compute Hx=b", nullptr, 0);
    PARAM_VECTOR(PropertyVector_<int>, max_iterations, "maximum iterations if no stopping criteria is set", 0);
    PARAM(PropertyFloat, mse_threshold, "Minimum mean square error variation to perform global optimization", -1.f, 0);
    PARAM(PropertyConfigurable_<TerminationCriteria>, termination_criteria,
          "term criteria ptr, if 0 solver will do max iterations", nullptr, 0);
    Solver() { _class_name = "Solver"; }
  };

  // one entry of MultiAligner2D::iterationStats() (apps/visual_test_aligner_2d.cpp:156)
  struct IterationStats {
    int iteration        = 0;
    float chi_inliers    = 0.f;
    float chi_kernelized = 0.f;
    int num_inliers      = 0;
    int num_outliers     = 0;  // kernelized factors
    int num_correspondences = 0;
    Vector3f estimate;         // t2v of the estimate after this iteration
  };
  using IterationStatsVector = std::vector<IterationStats>;
  std::ostream& operator<<(std::ostream& os, const IterationStatsVector& stats);

}  // namespace srrg2_solver

namespace srrg2_laser_slam_2d {
  using namespace srrg2_core;

  // CorrespondenceFinder_<Isometry2f, PointNormal2fVectorCloud, PointNormal2fVectorCloud>
  // (R/registration/correspondence_finder_normal_2f.h:9-12)
  class CorrespondenceFinderNormal2f : public Configurable {
  public:
    void setFixed(const PointNormal2fVectorCloud* fixed) {
      _fixed              = fixed;
      _fixed_changed_flag = true;
    }
    void setMoving(const PointNormal2fVectorCloud* moving) { _moving = moving; }
    void setLocalMapInSensor(const Isometry2f& T) { _local_map_in_sensor = T; }
    void setCorrespondences(CorrespondenceVector* c) { _correspondences = c; }
    virtual void compute() = 0;

  protected:
    const PointNormal2fVectorCloud* _fixed  = nullptr;
    const PointNormal2fVectorCloud* _moving = nullptr;
    Isometry2f _local_map_in_sensor;
    CorrespondenceVector* _correspondences = nullptr;
    bool _fixed_changed_flag               = true;
  };
  using CorrespondenceFinderNormal2fPtr = std::shared_ptr<CorrespondenceFinderNormal2f>;

  class CorrespondenceFinderProjective2f : public CorrespondenceFinderNormal2f {
  public:
    PARAM(PropertyFloat, point_distance, "max distance between corresponding points", 0.5f, 0);
    PARAM(PropertyFloat, normal_cos, "min cosinus between normals", 0.8f, 0);
    PARAM(PropertyConfigurable_<PointNormal2fProjectorPolar>, projector, "projector to compute correspondences",
          PointNormal2fProjectorPolarPtr(new PointNormal2fProjectorPolar), &_projector_changed_flag);
    CorrespondenceFinderProjective2f() { _class_name = "CorrespondenceFinderProjective2f"; }
    void compute() override;
    void fillParams(ls2d_params& p) const;

  protected:
    bool _projector_changed_flag = true;
    Ls2dDevice _device;
    std::vector<float> _staging;
  };
  using CorrespondenceFinderProjective2DPtr = std::shared_ptr<CorrespondenceFinderProjective2f>;

}  // namespace srrg2_laser_slam_2d

namespace srrg2_slam_interfaces {
  using namespace srrg2_core;

  // common part of AlignerSliceProcessor_<Factor, Fixed, Moving> (params L0.json:115-143)
  class AlignerSliceProcessorBase : public Configurable {
  public:
    PARAM(PropertyString, base_frame_id, "name of the base frame in the tf tree", "", 0);
    PARAM(PropertyString, fixed_slice_name, "name of the slice in the fixed scene", "", 0);
    PARAM(PropertyString, frame_id, "name of the sensor's frame in the tf tree", "", 0);
    PARAM(PropertyString, moving_slice_name, "name of the slice in the moving scene", "", 0);
    PARAM(PropertyConfigurable_<srrg2_solver::RobustifierBase>, robustifier, "robustifier used on this slice",
          nullptr, 0);
    void setPlatform(const PlatformPtr& platform) { _platform = platform; }
    virtual bool isLaserSlice() const { return false; }

  protected:
    PlatformPtr _platform;
  };
  using AlignerSliceProcessorBasePtr = std::shared_ptr<AlignerSliceProcessorBase>;

  // AlignerSliceProcessorPrior_<SE2PriorErrorFactor, Isometry2f, Isometry2f> (L0.json:291-310,
  // apps/visual_test_tracker_2d.cpp:88-93): both scenes carry a pose-valued slice ("odom"); the factor's
  // measurement is the odometry's prediction of moving_in_fixed, Z = fixed_pose^-1 * moving_pose (both poses
  // live in the odometry frame), its information matrix the SE2PriorErrorFactor default (identity) unless
  // setInformationMatrix() is called (oracle decision D15).
  class AlignerSliceOdom2DPrior : public AlignerSliceProcessorBase {
  public:
    AlignerSliceOdom2DPrior() {
      _class_name = "AlignerSliceOdom2DPrior";
      for (int i = 0; i < 3; ++i) _information.m[i][i] = 1.f;
    }
    void setInformationMatrix(const Matrix3f& omega) { _information = omega; }
    const Matrix3f& informationMatrix() const { return _information; }
    // true if both scenes hold this slice's pose
    bool bound(const PropertyContainerDynamic* fixed, const PropertyContainerDynamic* moving) const {
      return fixed && moving && fixed->pose(param_fixed_slice_name.value()) &&
             moving->pose(param_moving_slice_name.value());
    }
    Isometry2f measurement(const PropertyContainerDynamic* fixed, const PropertyContainerDynamic* moving) const {
      return fixed->pose(param_fixed_slice_name.value())->inverse() * *moving->pose(param_moving_slice_name.value());
    }

  protected:
    Matrix3f _information;
  };

  class AlignerSliceProcessorLaserBase : public AlignerSliceProcessorBase {
  public:
    PARAM(PropertyConfigurable_<srrg2_laser_slam_2d::CorrespondenceFinderNormal2f>, finder,
          "correspondence finder used in this cue", nullptr, 0);
    PARAM(PropertyInt, min_num_correspondences, "minimum number of correspondences in this slice", 0, 0);
    bool isLaserSlice() const override { return true; }
    virtual bool withSensor() const { return false; }
    virtual bool pointToPoint() const { return false; }  // SE2Point2PointErrorFactor instead of plane-to-plane
    // sensor_in_robot of this slice (identity unless WithSensor); throws if the tf lookup fails
    Isometry2f sensorInRobot() const;
    // the last finder pass of the aligner's compute().  The list is fetched from the device on first access (the
    // clouds are still resident there): a tracker that never looks at it does not pay a launch and a download per frame
    const CorrespondenceVector& correspondences() const {
      if (_fetch_correspondences) {
        std::function<void()> fetch;
        fetch.swap(_fetch_correspondences);
        fetch();
      }
      return _correspondences;
    }
    const PointNormal2fVectorCloud* fixed() const { return _fixed; }
    const PointNormal2fVectorCloud* moving() const { return _moving; }

  protected:
    friend class MultiAligner2D;
    virtual void setupFactor() {}
    mutable CorrespondenceVector _correspondences;
    mutable std::function<void()> _fetch_correspondences;  // set by MultiAligner2D::compute
    const PointNormal2fVectorCloud* _fixed  = nullptr;
    const PointNormal2fVectorCloud* _moving = nullptr;
  };

  struct AlignmentResult {
    Isometry2f moving_in_fixed;
    Vector3f estimate;  // t2v(moving_in_fixed)
    int status = 0;
    srrg2_solver::IterationStats last;
    Matrix3f information_matrix;
    int iterations = 0;
  };

  class MultiAligner2D : public Configurable {
  public:
    enum Status { Success = 0, NotEnoughCorrespondences = 1, NotEnoughInliers = 2, Fail = 3 };
    PARAM(PropertyBool, enable_inlier_only_runs,
          "toggles additional inlier only runs if sufficient inliers are available", false, 0);
    PARAM(PropertyBool, keep_only_inlier_correspondences,
          "toggles removal of correspondences which factors are not inliers in the last iteration", false, 0);
    PARAM(PropertyInt, max_iterations, "maximum number of iterations", 10, 0);
    PARAM(PropertyInt, min_num_inliers, "minimum number ofinliers", 10, 0);
    PARAM_VECTOR(PropertyConfigurableVector_<AlignerSliceProcessorBase>, slice_processors, "slices", 0);
    PARAM(PropertyConfigurable_<srrg2_solver::Solver>, solver, "this solver", nullptr, 0);
    PARAM(PropertyConfigurable_<srrg2_solver::TerminationCriteria>, termination_criteria,
          "termination criteria, not set=max iterations", nullptr, 0);
    MultiAligner2D() { _class_name = "MultiAligner2D"; }

    void setFixed(PropertyContainerDynamic* fixed) { _fixed_scene = fixed; }
    void setMoving(PropertyContainerDynamic* moving) { _moving_scene = moving; }
    void setMovingInFixed(const Isometry2f& T) { _moving_in_fixed = T; }
    void compute();
    const Isometry2f& movingInFixed() const { return _moving_in_fixed; }
    Status status() const { return _status; }
    int numInliers() const { return _iteration_stats.empty() ? 0 : _iteration_stats.back().num_inliers; }
    const srrg2_solver::IterationStatsVector& iterationStats() const { return _iteration_stats; }
    const Matrix3f& informationMatrix() const { return _information_matrix; }

    // ---- batched extension: pair p aligns *moving[p] onto *fixed[p] from guesses[p] in ONE launch
    void computeBatch(const std::vector<const PointNormal2fVectorCloud*>& fixed,
                      const std::vector<const PointNormal2fVectorCloud*>& moving,
                      const std::vector<Isometry2f>& guesses, std::vector<AlignmentResult>& results);

    // the complete parameter set this aligner hands to the C ABI (public for tests / tools): slice 0's record
    void fillParams(ls2d_params& p);
    // every laser slice's record (aligner-level fields repeated), in slice_processors order
    void fillSliceParams(std::vector<ls2d_params>& p);
    Ls2dDevice& device() { return _device; }

  protected:
    std::shared_ptr<AlignerSliceProcessorLaserBase> laserSlice();
    std::vector<std::shared_ptr<AlignerSliceProcessorLaserBase>> laserSlices();
    std::shared_ptr<AlignerSliceOdom2DPrior> boundPrior();
    void fillOne(ls2d_params& p, const std::shared_ptr<AlignerSliceProcessorLaserBase>& slice);
    void computeMulti(const std::vector<std::shared_ptr<AlignerSliceProcessorLaserBase>>& slices,
                      const std::shared_ptr<AlignerSliceOdom2DPrior>& prior);
    void storeOutcome(const ls2d_result& r, const std::vector<ls2d_iter_stats>& its);
    void exportCorrespondences(ls2d_handle* h, const ls2d_params& p, int fixed_set, int moving_set,
                               const std::shared_ptr<AlignerSliceProcessorLaserBase>& slice, const Isometry2f& estimate);
    PropertyContainerDynamic* _fixed_scene  = nullptr;
    PropertyContainerDynamic* _moving_scene = nullptr;
    Isometry2f _moving_in_fixed;
    Status _status = Fail;
    srrg2_solver::IterationStatsVector _iteration_stats;
    Matrix3f _information_matrix;
    Ls2dDevice _device;
    std::shared_ptr<int> _alive = std::make_shared<int>(0);  // what the slices' pending correspondence fetches watch
  };
  using MultiAligner2DPtr = std::shared_ptr<MultiAligner2D>;

  // result of a verified loop closure
  struct LoopClosure2D {
    int candidate = -1;  // index into the candidate list, -1: nothing accepted
    int guess     = -1;
    Isometry2f moving_in_fixed;
    float chi_inliers = 0.f;
    int num_inliers = 0, num_correspondences = 0;
  };

  // the candidate-verification part of MultiLoopDetectorBruteForce2D (L0.json:613-635): aligns the query
  // local map against every candidate from every initial guess in one batched launch and applies the
  // relocalize_* acceptance gates.  (The pose-graph search that proposes candidates is out of scope.)
  class MultiLoopDetectorBruteForce2D : public Configurable {
  public:
    PARAM(PropertyConfigurable_<Configurable>, local_map_selector, "module used to figure out which local maps should be checked", nullptr, 0);
    PARAM(PropertyConfigurable_<MultiAligner2D>, relocalize_aligner, "aligner used to register loop closures", nullptr, 0);
    PARAM(PropertyFloat, relocalize_max_chi_inliers, "maximum chi per inlier for success [chi]", 0.05f, 0);
    PARAM(PropertyInt, relocalize_min_inliers, "minimum number of inliers for success [int]", 500, 0);
    PARAM(PropertyFloat, relocalize_min_inliers_ratio, "minimum fraction of inliers over total correspondences [num_inliers/num_correspondences]", 0.7f, 0);
    MultiLoopDetectorBruteForce2D() { _class_name = "MultiLoopDetectorBruteForce2D"; }

    LoopClosure2D verify(const PointNormal2fVectorCloud& query,
                         const std::vector<const PointNormal2fVectorCloud*>& candidates,
                         const std::vector<std::vector<Isometry2f>>& guesses,
                         std::vector<AlignmentResult>* all = nullptr);
  };

}  // namespace srrg2_slam_interfaces

namespace srrg2_laser_slam_2d {

  // R/registration/aligner_slice_processor_laser_2d.h:7-17
  class AlignerSliceProcessorLaser2D : public srrg2_slam_interfaces::AlignerSliceProcessorLaserBase {
  public:
    AlignerSliceProcessorLaser2D() {
      _class_name = "AlignerSliceProcessorLaser2D";
      param_finder.setValue(CorrespondenceFinderProjective2DPtr(new CorrespondenceFinderProjective2f));
    }
  };
  using AlignerSliceProcessorLaser2DPtr = std::shared_ptr<AlignerSliceProcessorLaser2D>;

  // R/registration/aligner_slice_processor_laser_2d.h:21-42
  class AlignerSliceProcessorLaser2DWithSensor : public srrg2_slam_interfaces::AlignerSliceProcessorLaserBase {
  public:
    AlignerSliceProcessorLaser2DWithSensor() {
      _class_name = "AlignerSliceProcessorLaser2DWithSensor";
      param_finder.setValue(CorrespondenceFinderProjective2DPtr(new CorrespondenceFinderProjective2f));
    }
    bool withSensor() const override { return true; }

  protected:
    void setupFactor() override;  // R/registration/aligner_slice_processor_laser_2d_impl.cpp:7-10
  };
  using AlignerSliceProcessorLaser2DWithSensorPtr = std::shared_ptr<AlignerSliceProcessorLaser2DWithSensor>;

  // the same slices bound to SE2Point2PointErrorFactor[WithSensor] (BASELINE.json north_star: "point-to-line (and
  // point-to-point)"); the reference itself binds only the plane-to-plane factor (aligner_slice_processor_laser_2d.h:8,23)
  class AlignerSliceProcessorLaser2DPoint2Point : public AlignerSliceProcessorLaser2D {
  public:
    AlignerSliceProcessorLaser2DPoint2Point() { _class_name = "AlignerSliceProcessorLaser2DPoint2Point"; }
    bool pointToPoint() const override { return true; }
  };
  class AlignerSliceProcessorLaser2DPoint2PointWithSensor : public AlignerSliceProcessorLaser2DWithSensor {
  public:
    AlignerSliceProcessorLaser2DPoint2PointWithSensor() { _class_name = "AlignerSliceProcessorLaser2DPoint2PointWithSensor"; }
    bool pointToPoint() const override { return true; }
  };

  // R/mapping/scene_clipper_projective_2d.{h,cpp}: visibility clip of the local map (the tracker's moving cloud).
  // Driven as apps/visual_test_scene_clipper_projective_2d.cpp:103-114.
  class SceneClipperProjective2D : public Configurable {
  public:
    enum Status { Error = 0, Ready = 1, Successful = 2 };
    PARAM(PropertyConfigurable_<PointNormal2fProjectorPolar>, projector, "projector used to remap the points",
          PointNormal2fProjectorPolarPtr(new PointNormal2fProjectorPolar), 0);
    PARAM(PropertyFloat, voxelize_resolution, "resolution used to decimate the points in the scan on a grid [meters]",
          0.1f, 0);
    SceneClipperProjective2D() { _class_name = "SceneClipperProjective2D"; }
    void setFullScene(PointNormal2fVectorCloud* scene) { _full_scene = scene; }
    void setClippedSceneInRobot(PointNormal2fVectorCloud* scene) { _clipped_scene_in_robot = scene; }
    void setRobotInLocalMap(const Isometry2f& T) { _robot_in_local_map = T; }
    void setSensorInRobot(const Isometry2f& T) { _sensor_in_robot = T; }
    void compute();
    Status status() const { return _status; }

  protected:
    PointNormal2fVectorCloud* _full_scene             = nullptr;
    PointNormal2fVectorCloud* _clipped_scene_in_robot = nullptr;
    Isometry2f _robot_in_local_map, _sensor_in_robot;
    Status _status = Error;
    Ls2dDevice _device;
  };
  using SceneClipperProjective2DPtr = std::shared_ptr<SceneClipperProjective2D>;

  // R/mapping/merger_projective_2d.{h,cpp}: merges an aligned measurement into the local map, in place.
  // Driven as apps/visual_test_merger_projective_2d.cpp:120-123.
  class MergerProjective2D : public Configurable {
  public:
    enum Status { Error = 0, Success = 1 };
    PARAM(PropertyFloat, merge_threshold, "max distance for merging the points in the scene and the moving", 0.2f, 0);
    PARAM(PropertyConfigurable_<PointNormal2fProjectorPolar>, projector, "projector to compute correspondences",
          PointNormal2fProjectorPolarPtr(new PointNormal2fProjectorPolar), 0);
    MergerProjective2D() { _class_name = "MergerProjective2D"; }
    void setScene(PointNormal2fVectorCloud* scene) { _scene = scene; }
    void setMeasurement(const PointNormal2fVectorCloud* m) { _measurement = m; }
    void setMeasurementInScene(const Isometry2f& T) { _measurement_in_scene = T; }
    void compute();
    Status status() const { return _status; }

  protected:
    PointNormal2fVectorCloud* _scene             = nullptr;
    const PointNormal2fVectorCloud* _measurement = nullptr;
    Isometry2f _measurement_in_scene;
    Status _status = Error;
    Ls2dDevice _device;
  };
  using MergerProjective2DPtr = std::shared_ptr<MergerProjective2D>;

  // R/sensor_processing/raw_data_preprocessor_projective_2d.{h,cpp}: LaserMessage -> PointNormal2fVectorCloud on
  // the device (unprojection, sliding-window normals, voxelisation).  Driven as tests/test_measurement_adaptor.cpp:12-33.
  class RawDataPreprocessorProjective2D : public Configurable {
  public:
    enum Status { Error = 0, Ready = 1 };
    using MeasurementType      = PointNormal2fVectorCloud;
    using NormalComputatorType = NormalComputator1DSlidingWindowNormal;
    PARAM(PropertyConfigurable_<PointNormal2fUnprojectorPolar>, unprojector,
          "un-projector used to compute the scan from the cloud",
          PointNormal2fUnprojectorPolarPtr(new PointNormal2fUnprojectorPolar), 0);
    PARAM(PropertyConfigurable_<NormalComputatorType>, normal_computator_sliding, "normal computator object",
          NormalComputator1DSlidingWindowNormalPtr(new NormalComputator1DSlidingWindowNormal), 0);
    PARAM(PropertyFloat, range_min, "range_min [meters]", 0.f, 0);
    PARAM(PropertyFloat, range_max, "range_max [meters]", 1000.f, 0);
    PARAM(PropertyFloat, voxelize_resolution, "unproject voxelization resolution", 0.02f, 0);
    PARAM(PropertyString, scan_topic, "topic of the scan", "/scan", 0);
    RawDataPreprocessorProjective2D() { _class_name = "RawDataPreprocessorProjective2D"; }
    void setMeas(MeasurementType* meas) { _meas = meas; }
    bool setRawData(BaseSensorMessagePtr msg);
    void compute();
    Status status() const { return _status; }
    // the record handed to the C ABI for the current message (public for tests)
    void fillScanParams(ls2d_scan_params& sp) const;

  protected:
    void _processLaserMessage(LaserMessagePtr message);
    MeasurementType* _meas = nullptr;
    BaseSensorMessagePtr _raw_data;
    LaserMessagePtr _laser;
    std::vector<float>* _ranges = nullptr;
    Status _status              = Error;
    Ls2dDevice _device;
  };
  using RawDataPreprocessorProjective2DPtr = std::shared_ptr<RawDataPreprocessorProjective2D>;

  // ELF-constructor registration, like R/instances.h:14
  void srrg2_laser_slam_2d_registerTypes() __attribute__((constructor));

}  // namespace srrg2_laser_slam_2d
