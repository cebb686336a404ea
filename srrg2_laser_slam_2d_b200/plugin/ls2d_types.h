// ls2d_types.h -- the data types the reference passes through the hot path's plugin interface, without
// Eigen: PointNormal2f / PointNormal2fVectorCloud, Isometry2f with geometry2d::v2t / t2v, Correspondence,
// the dynamic property container clouds are handed over in, and the (tf-tree) Platform the WithSensor
// slice reads sensor_in_robot from.  Use sites in the reference (R/ = /root/reference/srrg2_laser_slam_2d/
// src/srrg2_laser_slam_2d/):  R/registration/correspondence_finder_normal_2f.h:9-12,
// R/registration/correspondence_finder_projective_2d.cpp:69-73, apps/visual_test_aligner_2d.cpp:97-127,
// apps/visual_test_correspondence_finder_projective_2d.cpp:71-79.
#pragma once

#include <cmath>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../csrc/ls2d_math.cuh"

namespace srrg2_core {

  struct Vector2f {
    float v[2] = {0.f, 0.f};
    Vector2f() {}
    Vector2f(float x_, float y_) { v[0] = x_, v[1] = y_; }
    float& x() { return v[0]; }
    float& y() { return v[1]; }
    const float& x() const { return v[0]; }
    const float& y() const { return v[1]; }
    float dot(const Vector2f& o) const { return v[0] * o.v[0] + v[1] * o.v[1]; }
    void setZero() { v[0] = v[1] = 0.f; }
  };

  struct Vector3f {
    float v[3] = {0.f, 0.f, 0.f};
    Vector3f() {}
    Vector3f(float x_, float y_, float z_) { v[0] = x_, v[1] = y_, v[2] = z_; }
    float& x() { return v[0]; }
    float& y() { return v[1]; }
    float& z() { return v[2]; }
    const float& x() const { return v[0]; }
    const float& y() const { return v[1]; }
    const float& z() const { return v[2]; }
    float operator()(int i) const { return v[i]; }
  };

  // symmetric 3x3 (the aligner's information matrix = H of the last linearisation)
  struct Matrix3f {
    float m[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    float operator()(int r, int c) const { return m[r][c]; }
  };

  // Eigen::Isometry2f content: R = [c -s; s c], t.  All arithmetic is the single-rounding binary32 sequence
  // the kernels and the oracle use (csrc/ls2d_math.cuh), so host-side compositions match the device's.
  class Isometry2f {
  public:
    Isometry2f() : _t(ls2d::iso_identity()) {}
    explicit Isometry2f(const ls2d::iso& t) : _t(t) {}
    static Isometry2f Identity() { return Isometry2f(); }
    Isometry2f inverse() const { return Isometry2f(ls2d::iso_inverse(_t)); }
    Isometry2f operator*(const Isometry2f& o) const { return Isometry2f(ls2d::iso_compose(_t, o._t)); }
    Vector2f operator*(const Vector2f& p) const {
      Vector2f r;
      ls2d::iso_apply(_t, p.x(), p.y(), r.x(), r.y());
      return r;
    }
    Vector2f rotate(const Vector2f& n) const {
      Vector2f r;
      ls2d::iso_rot(_t, n.x(), n.y(), r.x(), r.y());
      return r;
    }
    Vector2f translation() const { return Vector2f(_t.tx, _t.ty); }
    const ls2d::iso& raw() const { return _t; }

  private:
    ls2d::iso _t;
  };

  namespace geometry2d {
    inline Isometry2f v2t(const Vector3f& v) { return Isometry2f(ls2d::iso_v2t(v.x(), v.y(), v.z())); }
    inline Vector3f t2v(const Isometry2f& T) {
      return Vector3f(T.raw().tx, T.raw().ty, ls2d::atan2f_fdlibm(T.raw().s, T.raw().c));
    }
  }  // namespace geometry2d

  enum POINT_STATUS { Valid = 0, Invalid = 1 };

  class PointNormal2f {
  public:
    Vector2f& coordinates() { return _coordinates; }
    const Vector2f& coordinates() const { return _coordinates; }
    Vector2f& normal() { return _normal; }
    const Vector2f& normal() const { return _normal; }
    POINT_STATUS status = Valid;

  private:
    Vector2f _coordinates, _normal;
  };

  class PointNormal2fVectorCloud : public std::vector<PointNormal2f> {
  public:
    using std::vector<PointNormal2f>::vector;
    void transformInPlace(const Isometry2f& T) {
      for (auto& p : *this) {
        p.coordinates() = T * p.coordinates();
        p.normal()      = T.rotate(p.normal());
      }
    }
  };

  struct Correspondence {
    int fixed_idx  = -1;
    int moving_idx = -1;
    float response = 0.f;
    Correspondence() {}
    Correspondence(int f, int m, float r = 0.f) : fixed_idx(f), moving_idx(m), response(r) {}
  };
  using CorrespondenceVector = std::vector<Correspondence>;

  // named dynamic properties: how clouds reach the aligner (apps/visual_test_aligner_2d.cpp:108-118)
  class PropertyContainerDynamic {
  public:
    void setCloud(const std::string& name, PointNormal2fVectorCloud* cloud) { _clouds[name] = cloud; }
    PointNormal2fVectorCloud* cloud(const std::string& name) const {
      auto it = _clouds.find(name);
      return it == _clouds.end() ? nullptr : it->second;
    }
    // pose-valued slices (the "odom" slice of AlignerSliceOdom2DPrior, L0.json:291-310)
    void setPose(const std::string& name, Isometry2f* pose) { _poses[name] = pose; }
    Isometry2f* pose(const std::string& name) const {
      auto it = _poses.find(name);
      return it == _poses.end() ? nullptr : it->second;
    }

  private:
    std::map<std::string, PointNormal2fVectorCloud*> _clouds;
    std::map<std::string, Isometry2f*> _poses;
  };

  // the part of srrg2_core::Platform the WithSensor slice uses: a static transform per sensor frame
  class Platform {
  public:
    void addTransform(const std::string& frame_id, const std::string& base_frame_id, const Isometry2f& T) {
      _tf[frame_id + "<-" + base_frame_id] = T;
    }
    bool getTransform(Isometry2f& T, const std::string& frame_id, const std::string& base_frame_id) const {
      auto it = _tf.find(frame_id + "<-" + base_frame_id);
      if (it == _tf.end()) return false;
      T = it->second;
      return true;
    }

  private:
    std::map<std::string, Isometry2f> _tf;
  };
  using PlatformPtr = std::shared_ptr<Platform>;

}  // namespace srrg2_core
