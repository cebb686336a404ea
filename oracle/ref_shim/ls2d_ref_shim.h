// ls2d_ref_shim.h -- TEST INFRASTRUCTURE.  A stand-in for the un-vendored libraries the reference's in-repo sources
// include (Eigen3, srrg2_core, srrg2_slam_interfaces), just wide enough to compile THREE reference files where they
// lie under /root/reference, unmodified:
//     src/srrg2_laser_slam_2d/registration/correspondence_finder_projective_2d.cpp
//     src/srrg2_laser_slam_2d/mapping/merger_projective_2d.cpp
//     src/srrg2_laser_slam_2d/mapping/scene_clipper_projective_2d.cpp
//     src/srrg2_laser_slam_2d/sensor_processing/raw_data_preprocessor_projective_2d.cpp
// so that the oracle's line-by-line restatement of those files (ls2d_oracle.c: orc_find_correspondences, orc_merge,
// orc_clip_scene[_voxelized], orc_preprocess_scan) can be checked against the reference's own compiled control flow (oracle/_ref/libls2d_ref.so,
// tests/test_oracle_vs_reference_sources.py).
//
// What this does NOT pin: everything that lives upstream.  The polar projector, the isometry algebra and the point
// arithmetic below are supplied by the oracle's own restatement (orc_project, orc_inverse, orc_compose, decision
// points D1-D4, D13, D14 of ls2d_oracle.c; the unprojector, the sliding-window normals and PointCloud::voxelize by
// orc_unproject, orc_sliding_window_normals, orc_voxelize, P1-P7), because the real ones are absent.  Names, members and call signatures
// follow the way the reference's sources and apps use them (cited per item); nothing here is copied from upstream.
#pragma once

#include <algorithm>
#include <cmath>
#include <cstddef>
#include <iostream>
#include <iterator>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../ls2d_oracle.h"

#define EIGEN_MAKE_ALIGNED_OPERATOR_NEW  // merger_projective_2d.h:10, correspondence_finder_projective_2d.h:11
#define FG_RED(s) std::string(s)          // srrg_system_utils/shell_colors.h, used at scene_clipper_projective_2d.cpp:14

namespace srrg2_core {

  // ---- the slice of Eigen the three files touch ----------------------------------------------------------------
  struct Vector2f {
    float v[2] = {0.f, 0.f};
    Vector2f() {}
    Vector2f(float x_, float y_) { v[0] = x_, v[1] = y_; }
    float& x() { return v[0]; }
    float& y() { return v[1]; }
    float x() const { return v[0]; }
    float y() const { return v[1]; }
    float dot(const Vector2f& o) const { return v[0] * o.v[0] + v[1] * o.v[1]; }  // .cpp:69 of the finder
    float squaredNorm() const { return v[0] * v[0] + v[1] * v[1]; }
  };

  struct Matrix2f {
    float m[4] = {1.f, 0.f, 0.f, 1.f};  // row-major
    static Matrix2f Identity() { return Matrix2f(); }
    float operator()(int r, int c) const { return m[2 * r + c]; }
    struct Comma {  // "sensor_matrix << a, b, c, d;"  raw_data_preprocessor_projective_2d.cpp:89-90
      Matrix2f* t;
      int k;
      Comma operator,(double x) {
        t->m[k] = (float) x;
        return Comma{t, k + 1};
      }
    };
    Comma operator<<(double x) {
      m[0] = (float) x;
      return Comma{this, 1};
    }
  };
  inline std::ostream& operator<<(std::ostream& os, const Matrix2f& M) {  // printInfo(), finder .cpp:13-14
    return os << M.m[0] << " " << M.m[1] << "\n" << M.m[2] << " " << M.m[3];
  }

  struct Matrix3f {
    float m[9];
    bool operator!=(const Matrix3f& o) const {  // scene_clipper_projective_2d.cpp:60
      for (int i = 0; i < 9; ++i) {
        if (m[i] != o.m[i]) {
          return true;
        }
      }
      return false;
    }
  };

  // Isometry2f: rotation + translation; the algebra is the oracle's (orc_inverse / orc_compose)
  struct Isometry2f {
    orc_iso T;
    Isometry2f() { T = orc_v2t(0.f, 0.f, 0.f); }
    explicit Isometry2f(orc_iso t_) : T(t_) {}
    static Isometry2f Identity() { return Isometry2f(); }
    Isometry2f inverse() const { return Isometry2f(orc_inverse(T)); }                         // finder .cpp:47
    Isometry2f operator*(const Isometry2f& o) const { return Isometry2f(orc_compose(T, o.T)); }  // clipper .cpp:29
    Matrix3f matrix() const {
      Matrix3f M;
      M.m[0] = T.c, M.m[1] = -T.s, M.m[2] = T.tx;
      M.m[3] = T.s, M.m[4] = T.c, M.m[5] = T.ty;
      M.m[6] = 0.f, M.m[7] = 0.f, M.m[8] = 1.f;
      return M;
    }
  };

  // ---- srrg_config: Configurable, PARAM, the two property kinds the files declare ----------------------------------
  struct Configurable {
    virtual ~Configurable() {}
  };

  template <typename V>
  struct Property_ {
    Property_(const char*, const char*, V def, bool* changed_flag) : _value(def), _flag(changed_flag) {}
    const V& value() const { return _value; }
    V& value() { return _value; }  // "&laser_message->ranges.value()", raw_data_preprocessor_projective_2d.cpp:78
    void setValue(const V& v_) {
      _value = v_;
      if (_flag) {
        *_flag = true;
      }
    }
    V _value;
    bool* _flag;
  };
  using PropertyFloat = Property_<float>;
  using PropertyInt   = Property_<int>;
  using PropertyString = Property_<std::string>;

  template <typename C>
  struct PropertyConfigurable_ {
    PropertyConfigurable_(const char*, const char*, std::shared_ptr<C> def, bool* changed_flag) :
      _value(def), _flag(changed_flag) {}
    std::shared_ptr<C> value() const { return _value; }
    C* operator->() const { return _value.get(); }
    void setValue(std::shared_ptr<C> v_) {
      _value = v_;
      if (_flag) {
        *_flag = true;
      }
    }
    std::shared_ptr<C> _value;
    bool* _flag;
  };

// PARAM(type, name, description, default, changed-flag pointer) -> member param_<name>
#define PARAM(TYPE, NAME, DESC, DEFAULT, FLAG) TYPE param_##NAME = TYPE(#NAME, DESC, DEFAULT, FLAG)

  // ---- srrg_pcl: points, clouds, the projector's output matrix ---------------------------------------------------
  enum POINT_STATUS { Valid = 0, Invalid = 1 };
  enum TRANSFORM_CLASS { Isometry = 0 };  // transformInPlace<Isometry>(...), scene_clipper_projective_2d.cpp:61

  struct PointNormal2f {
    Vector2f _coordinates, _normal;
    POINT_STATUS status = Valid;
    Vector2f& coordinates() { return _coordinates; }
    const Vector2f& coordinates() const { return _coordinates; }
    Vector2f& normal() { return _normal; }
    const Vector2f& normal() const { return _normal; }
    // merger_projective_2d.cpp:72-74; semantics = decision D14 of ls2d_oracle.c
    PointNormal2f& operator+=(const PointNormal2f& o) {
      _coordinates.v[0] = _coordinates.v[0] + o._coordinates.v[0];
      _coordinates.v[1] = _coordinates.v[1] + o._coordinates.v[1];
      _normal.v[0]      = _normal.v[0] + o._normal.v[0];
      _normal.v[1]      = _normal.v[1] + o._normal.v[1];
      return *this;
    }
    PointNormal2f& operator*=(float s) {
      _coordinates.v[0] *= s, _coordinates.v[1] *= s, _normal.v[0] *= s, _normal.v[1] *= s;
      return *this;
    }
    void normalize() {
      const float z = _normal.squaredNorm();
      if (z > 0.f) {
        const float n = std::sqrt(z);
        _normal.v[0]  = _normal.v[0] / n;
        _normal.v[1]  = _normal.v[1] / n;
      }
    }
  };

  struct Vector4f {  // PlainVectorType of the cloud (voxelize coefficients, scene_clipper_projective_2d.cpp:45-46)
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    struct Comma {
      Vector4f* t;
      int k;
      Comma operator,(double x) {
        t->v[k] = (float) x;
        return Comma{t, k + 1};
      }
    };
    Comma operator<<(double x) {
      v[0] = (float) x;
      return Comma{this, 1};
    }
  };

  struct PointNormal2fVectorCloud : public std::vector<PointNormal2f> {
    using PlainVectorType = Vector4f;
    // R * p + t on the coordinates, R * n on the normal, each product/sum one binary32 operation in Eigen's order
    template <TRANSFORM_CLASS tc = Isometry>
    void transformInPlace(const Isometry2f& iso) {
      const orc_iso& T = iso.T;
      for (PointNormal2f& p : *this) {
        const float x = p._coordinates.v[0], y = p._coordinates.v[1], nx = p._normal.v[0], ny = p._normal.v[1];
        p._coordinates.v[0] = (T.c * x + (-T.s) * y) + T.tx;
        p._coordinates.v[1] = (T.s * x + T.c * y) + T.ty;
        p._normal.v[0]      = T.c * nx + (-T.s) * ny;
        p._normal.v[1]      = T.s * nx + T.c * ny;
      }
    }
    // PointCloud::voxelize(out, res_coeffs): the oracle's restatement (orc_voxelize, P7); Invalid points are skipped
    template <typename OutIt>
    void voxelize(OutIt out, const PlainVectorType& res_coeffs) const {
      std::vector<orc_point> in(size()), res(size() ? size() : 1);
      std::vector<uint8_t> valid(size() ? size() : 1);
      for (size_t i = 0; i < size(); ++i) {
        const PointNormal2f& p = (*this)[i];
        in[i].x = p._coordinates.v[0], in[i].y = p._coordinates.v[1], in[i].nx = p._normal.v[0], in[i].ny = p._normal.v[1];
        valid[i] = p.status == Valid;
      }
      const int32_t k = orc_voxelize(in.data(), valid.data(), (int32_t) size(), res_coeffs.v, res.data());
      for (int32_t i = 0; i < k; ++i) {
        PointNormal2f p;
        p._coordinates = Vector2f(res[(size_t) i].x, res[(size_t) i].y);
        p._normal      = Vector2f(res[(size_t) i].nx, res[(size_t) i].ny);
        *out++ = p;
      }
    }
  };

  template <typename E>
  struct Matrix_ {
    void resize(size_t rows, size_t cols) { _cols = cols, _d.assign(rows * cols, E()); }
    size_t size() const { return _d.size(); }
    size_t cols() const { return _cols; }
    E& at(size_t r, size_t c) { return _d[r * _cols + c]; }  // raw_data_preprocessor_projective_2d.cpp:26
    const E& at(size_t r, size_t c) const { return _d[r * _cols + c]; }
    size_t _cols = 0;
    typename std::vector<E>::iterator begin() { return _d.begin(); }
    typename std::vector<E>::iterator end() { return _d.end(); }
    typename std::vector<E>::const_iterator begin() const { return _d.begin(); }
    typename std::vector<E>::const_iterator end() const { return _d.end(); }
    std::vector<E> _d;
  };

  // PointNormal2fProjectorPolar: parameters as the configurations name them (LASER_0.json:312-338); compute() is the
  // ORACLE's projector (orc_project, decisions D1-D4) -- the real one is upstream and absent.
  struct PointNormal2fProjectorPolar : public Configurable {
    struct Entry {
      int source_idx = -1;
      float depth    = 0.f;
      PointNormal2f transformed;
    };
    using TargetMatrixType = Matrix_<Entry>;
    PARAM(PropertyInt, canvas_cols, "cols of the canvas", 721, nullptr);
    PARAM(PropertyInt, canvas_rows, "rows of the canvas", 1, nullptr);
    PARAM(PropertyFloat, angle_col_min, "start col angle [rad]", -3.14159f, nullptr);
    PARAM(PropertyFloat, angle_col_max, "end col angle [rad]", 3.14159f, nullptr);
    PARAM(PropertyFloat, range_min, "min laser range [m]", 0.3f, nullptr);
    PARAM(PropertyFloat, range_max, "max laser range [m]", 20.f, nullptr);
    void setCameraPose(const Isometry2f& pose) { _camera_pose = pose; }
    Matrix2f cameraMatrix() const {
      Matrix2f K;
      const float C = (float) param_canvas_cols.value();
      K.m[0] = C / (param_angle_col_max.value() - param_angle_col_min.value()), K.m[1] = C * 0.5f, K.m[2] = 0.f,
      K.m[3] = 0.f;
      return K;
    }
    template <typename It>
    void compute(TargetMatrixType& target, It begin, It end) {
      orc_params prm;
      orc_default_params(&prm);
      prm.canvas_cols   = param_canvas_cols.value();
      prm.angle_col_min = param_angle_col_min.value();
      prm.angle_col_max = param_angle_col_max.value();
      prm.range_min     = param_range_min.value();
      prm.range_max     = param_range_max.value();
      std::vector<orc_point> pts;
      for (It it = begin; it != end; ++it) {
        orc_point p;
        p.x = it->coordinates().x(), p.y = it->coordinates().y(), p.nx = it->normal().x(), p.ny = it->normal().y();
        pts.push_back(p);
      }
      std::vector<orc_cell> img((size_t) prm.canvas_cols);
      orc_project(&prm, _camera_pose.T, pts.data(), (int32_t) pts.size(), img.data());
      target.resize(1, (size_t) prm.canvas_cols);
      size_t c = 0;
      for (Entry& e : target) {
        e.source_idx                 = img[c].source_idx;
        e.depth                      = img[c].depth;
        e.transformed.coordinates() = Vector2f(img[c].px, img[c].py);
        e.transformed.normal()      = Vector2f(img[c].nx, img[c].ny);
        ++c;
      }
    }
    Isometry2f _camera_pose;
  };
  using PointNormal2fProjectorPolarPtr = std::shared_ptr<PointNormal2fProjectorPolar>;

  // PointNormal2fUnprojectorPolar (parameters as LASER_0.json:405-430 names them): compute<WithNormals>() is the
  // oracle's orc_unproject (P1: reads K(0,0) and K(0,1) of the sensor matrix the in-repo code builds)
  enum { WithNormals = 1, WithoutNormals = 0 };  // raw_data_preprocessor_projective_2d.cpp:31
  struct PointNormal2fUnprojectorPolar : public Configurable {
    PARAM(PropertyFloat, range_min, "min laser range [m]", 0.f, nullptr);
    PARAM(PropertyFloat, range_max, "max laser range [m]", 1000.f, nullptr);
    PARAM(PropertyFloat, angle_min, "start angle [rad]", -3.14159f, nullptr);
    PARAM(PropertyFloat, angle_max, "end angle [rad]", 3.14159f, nullptr);
    void setCameraMatrix(const Matrix2f& K) { _K = K; }
    template <int Mode, typename OutIt>
    void compute(OutIt out, const Matrix_<float>& ranges) {
      const int32_t n = (int32_t) ranges.size();
      std::vector<orc_point> pts((size_t)(n > 0 ? n : 1));
      const int32_t k = orc_unproject(param_range_min.value(), param_range_max.value(), _K(0, 0), _K(0, 1),
                                      ranges._d.data(), n, pts.data());
      for (int32_t i = 0; i < k; ++i) {
        PointNormal2f p;
        p._coordinates = Vector2f(pts[(size_t) i].x, pts[(size_t) i].y);
        p._normal      = Vector2f(pts[(size_t) i].nx, pts[(size_t) i].ny);
        *out++ = p;
      }
    }
    Matrix2f _K;
  };
  using PointNormal2fUnprojectorPolarPtr = std::shared_ptr<PointNormal2fUnprojectorPolar>;

  // NormalComputator1DSlidingWindow (LASER_0.json:711-719): the oracle's orc_sliding_window_normals (P2-P6)
  template <typename CloudType_, int idx_>
  struct NormalComputator1DSlidingWindow : public Configurable {
    PARAM(PropertyInt, normal_min_points, "min number of points to compute a normal", 5, nullptr);
    PARAM(PropertyFloat, normal_point_distance, "max normal point distance", 0.3f, nullptr);
    void computeNormals(CloudType_& cloud) {
      const int32_t n = (int32_t) cloud.size();
      std::vector<orc_point> pts((size_t)(n > 0 ? n : 1));
      std::vector<uint8_t> valid((size_t)(n > 0 ? n : 1));
      for (int32_t i = 0; i < n; ++i) {
        const PointNormal2f& p = cloud[(size_t) i];
        pts[(size_t) i].x = p._coordinates.v[0], pts[(size_t) i].y = p._coordinates.v[1];
        pts[(size_t) i].nx = p._normal.v[0], pts[(size_t) i].ny = p._normal.v[1];
      }
      orc_sliding_window_normals(pts.data(), n, param_normal_point_distance.value(), param_normal_min_points.value(),
                                 valid.data());
      for (int32_t i = 0; i < n; ++i) {
        PointNormal2f& p = cloud[(size_t) i];
        p._normal        = Vector2f(pts[(size_t) i].nx, pts[(size_t) i].ny);
        p.status         = valid[(size_t) i] ? Valid : Invalid;
      }
    }
  };

  // srrg_messages: the fields raw_data_preprocessor_projective_2d.cpp:78-87 reads
  struct BaseSensorMessage {
    virtual ~BaseSensorMessage() {}
    std::string topic;
  };
  using BaseSensorMessagePtr = std::shared_ptr<BaseSensorMessage>;
  struct LaserMessage : public BaseSensorMessage {
    Property_<std::vector<float>> ranges{"ranges", "", std::vector<float>(), nullptr};
    PropertyFloat range_min{"range_min", "", 0.f, nullptr}, range_max{"range_max", "", 0.f, nullptr};
    PropertyFloat angle_min{"angle_min", "", 0.f, nullptr}, angle_max{"angle_max", "", 0.f, nullptr};
  };
  using LaserMessagePtr = std::shared_ptr<LaserMessage>;

  // srrg_data_structures/correspondence.h
  struct Correspondence {
    int fixed_idx = -1, moving_idx = -1;
    float response = 0.f;
    Correspondence() {}
    Correspondence(int f, int m, float r = 0.f) : fixed_idx(f), moving_idx(m), response(r) {}
  };
  using CorrespondenceVector = std::vector<Correspondence>;

}  // namespace srrg2_core

namespace srrg2_slam_interfaces {

  // setters as the reference's apps call them (apps/visual_test_correspondence_finder_projective_2d.cpp:73-79),
  // members as the in-repo compute() reads them (correspondence_finder_projective_2d.cpp:24-48)
  template <typename EstimateType_, typename FixedType_, typename MovingType_>
  struct CorrespondenceFinder_ : public srrg2_core::Configurable {
    void setFixed(const FixedType_* f) { _fixed = f, _fixed_changed_flag = true; }
    void setMoving(const MovingType_* m) { _moving = m; }
    void setLocalMapInSensor(const EstimateType_& e) { _local_map_in_sensor = e; }
    void setCorrespondences(srrg2_core::CorrespondenceVector* c) { _correspondences = c; }
    virtual void compute() = 0;
    const FixedType_* _fixed   = nullptr;
    const MovingType_* _moving = nullptr;
    EstimateType_ _local_map_in_sensor;
    srrg2_core::CorrespondenceVector* _correspondences = nullptr;
    bool _fixed_changed_flag                           = true;
  };

  // apps/visual_test_merger_projective_2d.cpp:120-123; merger_projective_2d.cpp:19-99
  struct MergerBase : public srrg2_core::Configurable {
    enum Status { Error = 0, Initializing = 1, Success = 2 };
    Status _status = Error;
  };
  template <typename EstimateType_, typename SceneType_, typename MeasurementType_>
  struct Merger_ : public MergerBase {
    using MovingMeasurementType = MeasurementType_;
    void setScene(SceneType_* s) { _scene = s; }
    void setMeasurement(const MeasurementType_* m) { _measurement = m; }
    void setMeasurementInScene(const EstimateType_& e) { _measurement_in_scene = e; }
    virtual void compute() = 0;
    SceneType_* _scene                   = nullptr;
    const MeasurementType_* _measurement = nullptr;
    EstimateType_ _measurement_in_scene;
  };

  // apps/visual_test_merger_projective_2d.cpp:105-108; scene_clipper_projective_2d.cpp:12-64
  template <typename EstimateType_, typename SceneType_>
  struct SceneClipper_ : public srrg2_core::Configurable {
    using EstimateType = EstimateType_;
    using SceneType    = SceneType_;
    enum Status { Error = 0, Ready = 1, Successful = 2 };
    void setFullScene(SceneType_* s) { _full_scene = s; }
    void setClippedSceneInRobot(SceneType_* s) { _clipped_scene_in_robot = s; }
    void setRobotInLocalMap(const EstimateType_& e) { _robot_in_local_map = e; }
    void setSensorInRobot(const EstimateType_& e) { _sensor_in_robot = e; }
    virtual void compute() = 0;
    SceneType_* _full_scene              = nullptr;
    SceneType_* _clipped_scene_in_robot = nullptr;
    EstimateType_ _robot_in_local_map, _sensor_in_robot;
    Status _status = Error;
  };

  // srrg2_slam_interfaces/raw_data_preprocessors/raw_data_preprocessor.h: members and calls as the in-repo
  // source and apps/visual_test_correspondence_finder_projective_2d.cpp:62-66 use them
  template <typename MessageType_>
  std::shared_ptr<MessageType_> extractMessage(srrg2_core::BaseSensorMessagePtr msg, const std::string& topic) {
    std::shared_ptr<MessageType_> m = std::dynamic_pointer_cast<MessageType_>(msg);
    if (m && !topic.empty() && !m->topic.empty() && m->topic != topic) {
      return nullptr;
    }
    return m;
  }
  template <typename MeasurementType_>
  struct RawDataPreprocessor_ : public srrg2_core::Configurable {
    using MeasurementType = MeasurementType_;
    enum Status { Error = 0, Ready = 1 };
    void setMeas(MeasurementType_* m) { _meas = m; }
    virtual bool setRawData(srrg2_core::BaseSensorMessagePtr msg) {
      _raw_data = msg;
      return true;
    }
    virtual void compute() = 0;
    MeasurementType_* _meas = nullptr;
    srrg2_core::BaseSensorMessagePtr _raw_data;
    Status _status = Error;
  };

}  // namespace srrg2_slam_interfaces
