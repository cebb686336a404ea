// plugin_test.cpp -- driver used by tests/test_plugin_*.py.
//   plugin_test selftest                          config system + mis-wiring behaviour (no GPU needed)
//   plugin_test parse <config>                    load a BOSS-text configuration, print the hot-path objects
//   plugin_test align <config> <aligner> <in.bin> <out.bin>     (GPU) single-pair compute(), computeBatch(), finder
//   plugin_test verify <config> <detector> <in.bin> <out.bin>   (GPU) loop-closure candidate verification
//   plugin_test multi <config> <aligner> <in.bin> <out.bin>     (GPU) two laser slices + odometry prior, compute()
//   plugin_test scan <voxel_res> <in.bin> <out.bin>             (GPU) RawDataPreprocessorProjective2D; <in.bin>: int32
//                                                 n_scans, n_beams; float angle_min, angle_max; float ranges[]
// Binary layout of <in.bin>: int32 n_pairs, int32 n_guess, float sensor_in_robot[3]; then per pair:
//   int32 n_fixed, n_moving; float fixed[n_fixed*4]; float moving[n_moving*4]; float init[n_guess*3].
#include <chrono>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>

#include "ls2d_modules.h"

using namespace srrg2_core;
using namespace srrg2_solver;
using namespace srrg2_slam_interfaces;
using namespace srrg2_laser_slam_2d;

#define REQUIRE(cond)                                                             \
  do {                                                                            \
    if (!(cond)) {                                                                \
      std::fprintf(stderr, "REQUIRE failed at %s:%d: %s\n", __FILE__, __LINE__, #cond); \
      std::exit(1);                                                               \
    }                                                                             \
  } while (0)

template <typename F>
static std::string thrown(F f) {
  try {
    f();
  } catch (const std::runtime_error& e) {
    return e.what();
  }
  return "";
}

static const char* kEmbedded = R"BOSS(
// two finders sharing one projector, an aligner, an unknown class and a null link
"PointNormal2fProjectorPolar" { "#id" : 31, "angle_col_max" : 3.14159, "angle_col_min" : -3.14159,
  "canvas_cols" : 721, "canvas_rows" : 1, "range_max" : 20, "range_min" : 0.3 }
"CorrespondenceFinderProjective2f" { "#id" : 17, "normal_cos" : 0.9, "point_distance" : 0.5,
  "projector" : { "#pointer" : 31 } }
"CorrespondenceFinderProjective2f" { "#id" : 42, "name" : "ld_finder", "normal_cos" : 0.8, "point_distance" : 1.414,
  "projector" : { "#pointer" : 31 } }
"RobustifierCauchy" { "#id" : 24, // threshold of chi after which the kernel is active
  "chi_threshold" : 0.05 }
"AlignerSliceProcessorLaser2D" { "#id" : 3, "name" : "al_sl_ld_laser_0", "base_frame_id" : "b", "frame_id" : "f",
  "finder" : { "#pointer" : 42 }, "fixed_slice_name" : "points", "moving_slice_name" : "points",
  "min_num_correspondences" : 0, "robustifier" : { "#pointer" : 24 } }
"SomethingOutsideTheHotPath" { "#id" : 99, "name" : "slam", "whatever" : [ 1, 2, 3 ], "link" : { "#pointer" : 3 } }
"MultiAligner2D" { "#id" : 2, "name" : "multi_aligner_ld", "enable_inlier_only_runs" : 0,
  "keep_only_inlier_correspondences" : 0, "max_iterations" : 30, "min_num_inliers" : 10,
  "slice_processors" : [ { "#pointer" : 3 } ], "solver" : { "#pointer" : -1 },
  "termination_criteria" : { "#pointer" : -1 } }
)BOSS";

static int selftest() {
  // registration happened in the ELF constructor (R/instances.h:14)
  for (const char* c : {"CorrespondenceFinderProjective2f", "AlignerSliceProcessorLaser2D",
                        "AlignerSliceProcessorLaser2DWithSensor", "MultiAligner2D", "PointNormal2fProjectorPolar",
                        "RobustifierCauchy", "IterationAlgorithmGN", "Solver", "MultiLoopDetectorBruteForce2D",
                        "SceneClipperProjective2D", "MergerProjective2D", "RawDataPreprocessorProjective2D",
                        "PointNormal2fUnprojectorPolar", "NormalComputator1DSlidingWindowNormal"})
    REQUIRE(ClassRegistry::instance().has(c));

  ConfigurableManager m;
  m.readString(kEmbedded);
  REQUIRE(m.objects().size() == 7);
  auto aligner = m.getByName<MultiAligner2D>("multi_aligner_ld");
  REQUIRE(aligner && aligner->param_max_iterations.value() == 30 && aligner->param_min_num_inliers.value() == 10);
  REQUIRE(aligner->param_slice_processors.size() == 1 && !aligner->param_solver.value());
  auto slice = std::dynamic_pointer_cast<AlignerSliceProcessorLaser2D>(aligner->param_slice_processors.value(0));
  REQUIRE(slice && slice->name() == "al_sl_ld_laser_0");
  auto f42 = m.getByName<CorrespondenceFinderProjective2f>("ld_finder");
  auto f17 = m.getById<CorrespondenceFinderProjective2f>(17);
  REQUIRE(f42 && f17 && slice->param_finder.value() == f42);
  REQUIRE(f42->param_projector.value() == f17->param_projector.value());  // nested modules are shared
  REQUIRE(f42->param_point_distance.value() == 1.414f && f17->param_normal_cos.value() == 0.9f);
  REQUIRE(m.getByName<GenericConfigurable>("slam") && m.getByName<GenericConfigurable>("slam")->className() ==
                                                        "SomethingOutsideTheHotPath");
  ls2d_params p;
  aligner->fillParams(p);
  REQUIRE(p.canvas_cols == 721 && p.max_iterations == 30 && p.cauchy_chi_threshold == 0.05f && p.with_sensor == 0);
  REQUIRE(p.point_distance == 1.414f && p.normal_cos == 0.8f && p.range_min == 0.3f && p.range_max == 20.f);

  // write -> read round trip keeps parameters, links and unknown objects
  ConfigurableManager m2;
  m2.readString(m.writeString());
  REQUIRE(m2.objects().size() == 7);
  ls2d_params p2;
  m2.getByName<MultiAligner2D>("multi_aligner_ld")->fillParams(p2);
  REQUIRE(std::memcmp(&p, &p2, sizeof(p)) == 0);
  REQUIRE(m2.getByName<GenericConfigurable>("slam")->unknown_fields.size() == 2);

  // programmatic construction as apps/slam_app.cpp:87-167 does
  ConfigurableManager m3;
  auto a3 = m3.create<MultiAligner2D>("tracker_aligner");
  auto s3 = m3.create<AlignerSliceProcessorLaser2DWithSensor>("slice");
  s3->param_fixed_slice_name.setValue("points");
  s3->param_moving_slice_name.setValue("points");
  s3->param_frame_id.setValue("scan");
  s3->param_base_frame_id.setValue("base_frame");
  a3->param_slice_processors.pushBack(s3);
  a3->param_max_iterations.setValue(10);
  REQUIRE(s3->param_finder.value());  // default finder set by the slice constructor (aligner_slice_processor_laser_2d.h:13-16)
  REQUIRE(thrown([&] { ls2d_params q; a3->fillParams(q); }).find("no platform set") != std::string::npos);
  PlatformPtr platform(new Platform);
  s3->setPlatform(platform);
  REQUIRE(thrown([&] { ls2d_params q; a3->fillParams(q); }).find("unable to find transform") != std::string::npos);
  platform->addTransform("scan", "base_frame", geometry2d::v2t(Vector3f(0.2f, 0.2f, 0.1f)));
  ls2d_params q;
  a3->fillParams(q);
  // sensor_in_robot crosses the ABI as the isometry the platform holds (tx, ty, c, s), not as t2v of it
  const Isometry2f S3 = geometry2d::v2t(Vector3f(0.2f, 0.2f, 0.1f));
  REQUIRE(q.with_sensor == 2 && q.sensor_in_robot[0] == 0.2f && q.sensor_in_robot_cs[0] == S3.raw().c &&
          q.sensor_in_robot_cs[1] == S3.raw().s);
  REQUIRE(q.factor == LS2D_FACTOR_PLANE2PLANE && q.algorithm == LS2D_ALGORITHM_GN && q.enable_inlier_only_runs == 0);
  REQUIRE(q.cauchy_chi_threshold < 0.f);  // no robustifier on the slice

  // mis-wiring throws std::runtime_error with the reference's messages (correspondence_finder_projective_2d.cpp:20-31)
  CorrespondenceFinderProjective2f finder;
  CorrespondenceVector corr;
  PointNormal2fVectorCloud cloud(3);
  REQUIRE(thrown([&] { finder.compute(); }) == "CorrespondenceFinderProjective2f::compute| Missing fixed!");
  finder.setFixed(&cloud);
  REQUIRE(thrown([&] { finder.compute(); }) == "CorrespondenceFinderProjective2f::compute| Missing moving!");
  finder.param_projector.setValue(nullptr);
  REQUIRE(thrown([&] { finder.compute(); }) == "CorrespondenceFinderProjective2f::compute| Missing Projector");
  MultiAligner2D bare;
  REQUIRE(thrown([&] { bare.compute(); }) == "MultiAligner2D::compute| Missing fixed!");
  PropertyContainerDynamic scene;
  bare.setFixed(&scene);
  bare.setMoving(&scene);
  REQUIRE(thrown([&] { bare.compute(); }) == "MultiAligner2D::compute| no slice processor set");
  REQUIRE(thrown([&] { ConfigurableManager bad; bad.readString("\"Solver\" { \"#id\" : 1, \"algorithm\" : { \"#pointer\" : 5 } }"); })
            .find("dangling #pointer 5") != std::string::npos);
  REQUIRE(thrown([&] { ConfigurableManager bad; bad.readString("\"MultiAligner2D\" { \"#id\" : 1, \"solver\" : { \"#pointer\" : 2 } }\n"
                                                               "\"RobustifierCauchy\" { \"#id\" : 2 }"); })
            .find("wrong class") != std::string::npos);

  // RawDataPreprocessorProjective2D: the reference's error behaviour (raw_data_preprocessor_projective_2d.cpp:14-21,54-69)
  RawDataPreprocessorProjective2D adaptor;
  REQUIRE(adaptor.param_normal_computator_sliding.value() && adaptor.param_unprojector.value());  // test_measurement_adaptor.cpp:13-14
  REQUIRE(adaptor.param_voxelize_resolution.value() == 0.02f && adaptor.param_scan_topic.value() == "/scan");
  REQUIRE(thrown([&] { adaptor.setRawData(nullptr); }) ==
          "RawDataPreprocessorProjective2D::setMeasurement|measurement is not set");
  LaserMessagePtr other(new LaserMessage("/other_scan"));
  REQUIRE(!adaptor.setRawData(other) && adaptor.status() == RawDataPreprocessorProjective2D::Error);
  adaptor.compute();  // no measurement cloud set: status Error, no throw
  REQUIRE(adaptor.status() == RawDataPreprocessorProjective2D::Error);
  LaserMessagePtr scan(new LaserMessage("/scan"));
  scan->angle_min.setValue(-1.f), scan->angle_max.setValue(1.f), scan->range_min.setValue(0.f), scan->range_max.setValue(30.f);
  scan->ranges.setValue(std::vector<float>(100, 1.f));
  REQUIRE(adaptor.setRawData(scan) && adaptor.status() == RawDataPreprocessorProjective2D::Ready);
  REQUIRE(adaptor.param_unprojector->param_range_max.value() == 30.f);  // the tighter of message and PARAM (.cpp:83)
  REQUIRE(std::fabs(adaptor.param_unprojector->fx() - 50.f) < 1e-3f && adaptor.param_unprojector->cx() == 50.f);  // .cpp:89-90
  adaptor.param_unprojector.setValue(nullptr);
  PointNormal2fVectorCloud meas;
  adaptor.setMeas(&meas);
  REQUIRE(thrown([&] { adaptor.compute(); }) == "RawDataPreprocessorProjective2D::compute| missing unprojector");

  // geometry helpers agree with their definition
  const Isometry2f T = geometry2d::v2t(Vector3f(1.f, -2.f, 0.5f));
  const Vector3f back = geometry2d::t2v(T * T.inverse());
  REQUIRE(std::fabs(back.x()) < 1e-6f && std::fabs(back.y()) < 1e-6f && std::fabs(back.z()) < 1e-6f);
  std::printf("SELFTEST OK\n");
  return 0;
}

static PlatformPtr identityPlatformFor(ConfigurableManager& m, const Isometry2f& S) {
  PlatformPtr platform(new Platform);
  for (auto& s : m.getAll<AlignerSliceProcessorBase>()) {
    platform->addTransform(s->param_frame_id.value(), s->param_base_frame_id.value(), S);
    s->setPlatform(platform);
  }
  return platform;
}

static int parse(const std::string& file) {
  ConfigurableManager m;
  m.read(file);
  identityPlatformFor(m, Isometry2f::Identity());
  std::printf("{\"objects\": %zu, \"aligners\": [", m.objects().size());
  bool first = true;
  for (auto& a : m.getAll<MultiAligner2D>()) {
    ls2d_params p;
    const std::string why = thrown([&] { a->fillParams(p); });
    if (!why.empty()) {
      std::printf("%s{\"id\": %d, \"name\": \"%s\", \"slices\": %zu, \"unsupported\": \"%s\"}", first ? "" : ", ",
                  m.idOf(a.get()), a->name().c_str(), a->param_slice_processors.size(), why.c_str());
      first = false;
      continue;
    }
    std::printf("%s{\"id\": %d, \"name\": \"%s\", \"slices\": %zu, \"canvas_cols\": %d, \"angle_col_min\": %.6f, "
                "\"angle_col_max\": %.6f, \"range_min\": %.4f, \"range_max\": %.4f, \"point_distance\": %.4f, "
                "\"normal_cos\": %.4f, \"cauchy_chi_threshold\": %.4f, \"damping\": %.4f, \"max_iterations\": %d, "
                "\"min_num_correspondences\": %d, \"min_num_inliers\": %d, \"with_sensor\": %d, \"laser_slices\": [",
                first ? "" : ", ", m.idOf(a.get()), a->name().c_str(), a->param_slice_processors.size(), p.canvas_cols,
                p.angle_col_min, p.angle_col_max, p.range_min, p.range_max, p.point_distance, p.normal_cos,
                p.cauchy_chi_threshold, p.damping, p.max_iterations, p.min_num_correspondences, p.min_num_inliers,
                p.with_sensor ? 1 : 0);
    std::vector<ls2d_params> all;
    a->fillSliceParams(all);  // every laser slice of the aligner (MULTI.json: two rangefinders)
    for (size_t k = 0; k < all.size(); ++k)
      std::printf("%s{\"point_distance\": %.4f, \"normal_cos\": %.4f, \"cauchy_chi_threshold\": %.4f, "
                  "\"min_num_correspondences\": %d, \"with_sensor\": %d, \"canvas_cols\": %d}",
                  k ? ", " : "", all[k].point_distance, all[k].normal_cos, all[k].cauchy_chi_threshold,
                  all[k].min_num_correspondences, all[k].with_sensor ? 1 : 0, all[k].canvas_cols);
    int priors = 0;
    for (size_t k = 0; k < a->param_slice_processors.size(); ++k)
      if (std::dynamic_pointer_cast<AlignerSliceOdom2DPrior>(a->param_slice_processors.value(k))) ++priors;
    std::printf("], \"prior_slices\": %d}", priors);
    first = false;
  }
  std::printf("], \"loop_detectors\": [");
  first = true;
  for (auto& d : m.getAll<MultiLoopDetectorBruteForce2D>()) {
    std::printf("%s{\"name\": \"%s\", \"min_inliers\": %d, \"max_chi\": %.4f, \"min_ratio\": %.4f, \"aligner\": %d}",
                first ? "" : ", ", d->name().c_str(), d->param_relocalize_min_inliers.value(),
                d->param_relocalize_max_chi_inliers.value(), d->param_relocalize_min_inliers_ratio.value(),
                d->param_relocalize_aligner.value() ? m.idOf(d->param_relocalize_aligner.value().get()) : -1);
    first = false;
  }
  std::printf("], \"preprocessors\": [");
  first = true;
  for (auto& r : m.getAll<RawDataPreprocessorProjective2D>()) {
    ls2d_scan_params sp;
    r->fillScanParams(sp);
    std::printf("%s{\"name\": \"%s\", \"scan_topic\": \"%s\", \"voxelize_resolution\": %.4f, \"range_min\": %.4f, "
                "\"range_max\": %.4f, \"normal_point_distance\": %.4f, \"normal_min_points\": %d, \"num_ranges\": %d}",
                first ? "" : ", ", r->name().c_str(), r->param_scan_topic.value().c_str(), sp.voxelize_resolution,
                sp.range_min, sp.range_max, sp.normal_point_distance, sp.normal_min_points,
                r->param_unprojector.value() ? r->param_unprojector->param_num_ranges.value() : -1);
    first = false;
  }
  std::printf("], \"finders\": %zu, \"projectors\": %zu}\n", m.getAll<CorrespondenceFinderProjective2f>().size(),
              m.getAll<PointNormal2fProjectorPolar>().size());
  return 0;
}

struct PairsFile {
  int32_t n_pairs = 0, n_guess = 0;
  float sensor[3] = {0, 0, 0};
  std::vector<PointNormal2fVectorCloud> fixed, moving;
  std::vector<std::vector<Isometry2f>> guesses;
};

static PointNormal2fVectorCloud readCloud(std::ifstream& is, int n) {
  std::vector<float> buf((size_t) n * 4);
  is.read((char*) buf.data(), (std::streamsize)(buf.size() * sizeof(float)));
  PointNormal2fVectorCloud c((size_t) n);
  for (int i = 0; i < n; ++i) {
    c[i].coordinates() = Vector2f(buf[4 * i], buf[4 * i + 1]);
    c[i].normal()      = Vector2f(buf[4 * i + 2], buf[4 * i + 3]);
  }
  return c;
}

static PairsFile readPairs(const std::string& file) {
  std::ifstream is(file, std::ios::binary);
  if (!is.good()) throw std::runtime_error("cannot open " + file);
  PairsFile f;
  is.read((char*) &f.n_pairs, 4);
  is.read((char*) &f.n_guess, 4);
  is.read((char*) f.sensor, 12);
  for (int p = 0; p < f.n_pairs; ++p) {
    int32_t nf, nm;
    is.read((char*) &nf, 4);
    is.read((char*) &nm, 4);
    f.fixed.push_back(readCloud(is, nf));
    f.moving.push_back(readCloud(is, nm));
    std::vector<float> g((size_t) f.n_guess * 3);
    is.read((char*) g.data(), (std::streamsize)(g.size() * sizeof(float)));
    std::vector<Isometry2f> gs;
    for (int k = 0; k < f.n_guess; ++k) gs.push_back(geometry2d::v2t(Vector3f(g[3 * k], g[3 * k + 1], g[3 * k + 2])));
    f.guesses.push_back(gs);
  }
  if (!is.good()) throw std::runtime_error("short read on " + file);
  return f;
}

static void writeResult(std::ofstream& os, const Vector3f& est, int status, float chi_in, float chi_k, int n_in, int n_out,
                        int n_corr, int iterations) {
  const float f[5]   = {est.x(), est.y(), est.z(), chi_in, chi_k};
  const int32_t i[5] = {status, n_in, n_out, n_corr, iterations};
  os.write((const char*) f, sizeof(f));
  os.write((const char*) i, sizeof(i));
}

// (GPU) single-pair latency of MultiAligner2D::compute() -- the reference's real tracker use: one compute() per frame
// (apps/visual_test_tracker_2d.cpp:167-179).  Every call stages the two clouds, uploads them, runs the aligner and
// reads the result and the last correspondence list back; prints one JSON line (microseconds per call).
static int latency(const std::string& config, const std::string& name, const std::string& in, int reps) {
  ConfigurableManager m;
  m.read(config);
  PairsFile f = readPairs(in);
  identityPlatformFor(m, geometry2d::v2t(Vector3f(f.sensor[0], f.sensor[1], f.sensor[2])));
  MultiAligner2DPtr aligner = m.getByName<MultiAligner2D>(name);
  if (!aligner) throw std::runtime_error("no MultiAligner2D named " + name);
  std::vector<double> us;
  int ok = 0;
  for (int r = -3; r < reps; ++r) {  // three untimed warm-up calls (context, kernel load, buffer growth)
    const int p = ((r % f.n_pairs) + f.n_pairs) % f.n_pairs;
    PropertyContainerDynamic fixed_scene, moving_scene;
    fixed_scene.setCloud("points", &f.fixed[p]);
    moving_scene.setCloud("points", &f.moving[p]);
    const auto t0 = std::chrono::steady_clock::now();
    aligner->setFixed(&fixed_scene);
    aligner->setMoving(&moving_scene);
    aligner->setMovingInFixed(f.guesses[p][0]);
    aligner->compute();
    const auto t1 = std::chrono::steady_clock::now();
    if (r >= 0) {
      us.push_back(std::chrono::duration<double, std::micro>(t1 - t0).count());
      ok += aligner->status() == MultiAligner2D::Success;
    }
  }
  std::sort(us.begin(), us.end());
  double mean = 0;
  for (double v : us) mean += v;
  mean /= (double) us.size();
  std::printf("{\"calls\": %d, \"success\": %d, \"us_median\": %.1f, \"us_mean\": %.1f, \"us_min\": %.1f, \"us_p90\": %.1f, "
              "\"points\": %zu}\n",
              reps, ok, us[us.size() / 2], mean, us.front(), us[us.size() * 9 / 10], f.fixed[0].size());
  return 0;
}

static int align(const std::string& config, const std::string& name, const std::string& in, const std::string& out) {
  ConfigurableManager m;
  m.read(config);
  PairsFile f = readPairs(in);
  identityPlatformFor(m, geometry2d::v2t(Vector3f(f.sensor[0], f.sensor[1], f.sensor[2])));
  MultiAligner2DPtr aligner = m.getByName<MultiAligner2D>(name);
  if (!aligner) throw std::runtime_error("no MultiAligner2D named " + name);
  std::ofstream os(out, std::ios::binary);
  // (1) the reference's single-pair call sequence (apps/visual_test_aligner_2d.cpp:108-156)
  for (int p = 0; p < f.n_pairs; ++p) {
    PropertyContainerDynamic fixed_scene, moving_scene;
    fixed_scene.setCloud("points", &f.fixed[p]);
    moving_scene.setCloud("points", &f.moving[p]);
    aligner->setFixed(&fixed_scene);
    aligner->setMoving(&moving_scene);
    aligner->setMovingInFixed(f.guesses[p][0]);
    aligner->compute();
    const auto& st = aligner->iterationStats();
    const IterationStats last = st.empty() ? IterationStats() : st.back();
    writeResult(os, geometry2d::t2v(aligner->movingInFixed()), (int) aligner->status(), last.chi_inliers,
                last.chi_kernelized, last.num_inliers, last.num_outliers, last.num_correspondences, (int) st.size());
    if (p == 0) {  // slice->correspondences() after compute()
      auto slice = std::dynamic_pointer_cast<AlignerSliceProcessorLaserBase>(aligner->param_slice_processors.value(0));
      const int32_t n = (int32_t) slice->correspondences().size();
      os.write((const char*) &n, 4);
      for (const auto& c : slice->correspondences()) {
        os.write((const char*) &c.fixed_idx, 4);
        os.write((const char*) &c.moving_idx, 4);
      }
    }
  }
  // (2) the batched extension: one launch for all pairs
  std::vector<const PointNormal2fVectorCloud*> fx, mv;
  std::vector<Isometry2f> gs;
  for (int p = 0; p < f.n_pairs; ++p) fx.push_back(&f.fixed[p]), mv.push_back(&f.moving[p]), gs.push_back(f.guesses[p][0]);
  std::vector<AlignmentResult> res;
  aligner->computeBatch(fx, mv, gs, res);
  for (const auto& r : res)
    writeResult(os, r.estimate, r.status, r.last.chi_inliers, r.last.chi_kernelized, r.last.num_inliers,
                r.last.num_outliers, r.last.num_correspondences, r.iterations);
  // (3) the finder on its own (apps/visual_test_correspondence_finder_projective_2d.cpp:73-79), pair 0 at its guess
  auto slice  = std::dynamic_pointer_cast<AlignerSliceProcessorLaserBase>(aligner->param_slice_processors.value(0));
  auto finder = slice->param_finder.value();
  CorrespondenceVector corr;
  finder->setFixed(&f.fixed[0]);
  finder->setMoving(&f.moving[0]);
  finder->setLocalMapInSensor(f.guesses[0][0]);
  finder->setCorrespondences(&corr);
  finder->compute();
  finder->compute();  // second call reuses the cached fixed image
  const int32_t n = (int32_t) corr.size();
  os.write((const char*) &n, 4);
  for (const auto& c : corr) {
    os.write((const char*) &c.fixed_idx, 4);
    os.write((const char*) &c.moving_idx, 4);
  }
  std::printf("ALIGN OK %d pairs\n", f.n_pairs);
  return 0;
}

static int verify(const std::string& config, const std::string& name, const std::string& in, const std::string& out) {
  ConfigurableManager m;
  m.read(config);
  PairsFile f = readPairs(in);
  identityPlatformFor(m, Isometry2f::Identity());
  auto det = m.getByName<MultiLoopDetectorBruteForce2D>(name);
  if (!det) throw std::runtime_error("no MultiLoopDetectorBruteForce2D named " + name);
  std::vector<const PointNormal2fVectorCloud*> cands;
  for (int p = 0; p < f.n_pairs; ++p) cands.push_back(&f.moving[p]);
  std::vector<AlignmentResult> all;
  const LoopClosure2D lc = det->verify(f.fixed[0], cands, f.guesses, &all);
  std::ofstream os(out, std::ios::binary);
  const int32_t head[4] = {lc.candidate, lc.guess, lc.num_inliers, lc.num_correspondences};
  os.write((const char*) head, sizeof(head));
  const Vector3f est = geometry2d::t2v(lc.moving_in_fixed);
  os.write((const char*) est.v, 12);
  for (const auto& r : all)
    writeResult(os, r.estimate, r.status, r.last.chi_inliers, r.last.chi_kernelized, r.last.num_inliers,
                r.last.num_outliers, r.last.num_correspondences, r.iterations);
  std::printf("VERIFY OK candidate %d guess %d\n", lc.candidate, lc.guess);
  return 0;
}

// Binary layout of the multi-slice input: int32 n_pairs; float sensor0[3], sensor1[3]; float information[6]; then per
// pair: int32 n_fixed0, n_fixed1, n_moving; the three clouds; float init[3], odom_fixed[3], odom_moving[3].
static int multi(const std::string& config, const std::string& name, const std::string& in, const std::string& out) {
  ConfigurableManager m;
  m.read(config);
  MultiAligner2DPtr aligner = m.getByName<MultiAligner2D>(name);
  if (!aligner) throw std::runtime_error("no MultiAligner2D named " + name);
  std::ifstream is(in, std::ios::binary);
  if (!is.good()) throw std::runtime_error("cannot open " + in);
  int32_t n_pairs = 0;
  float sensors[2][3], info[6];
  is.read((char*) &n_pairs, 4);
  is.read((char*) sensors, sizeof(sensors));
  is.read((char*) info, sizeof(info));
  // tf tree: every laser slice finds its own sensor_in_robot by frame_id (aligner_slice_processor_laser_2d_impl.cpp:7-10)
  PlatformPtr platform(new Platform);
  std::vector<std::shared_ptr<AlignerSliceProcessorLaserBase>> lasers;
  std::shared_ptr<AlignerSliceOdom2DPrior> prior;
  for (size_t k = 0; k < aligner->param_slice_processors.size(); ++k) {
    auto s = aligner->param_slice_processors.value(k);
    s->setPlatform(platform);
    if (auto l = std::dynamic_pointer_cast<AlignerSliceProcessorLaserBase>(s)) {
      const float* v = sensors[lasers.size() < 2 ? lasers.size() : 1];
      platform->addTransform(l->param_frame_id.value(), l->param_base_frame_id.value(),
                             geometry2d::v2t(Vector3f(v[0], v[1], v[2])));
      lasers.push_back(l);
    } else if (auto p = std::dynamic_pointer_cast<AlignerSliceOdom2DPrior>(s)) {
      prior = p;
    }
  }
  REQUIRE(lasers.size() == 2 && prior);
  Matrix3f O;
  const int map[3][3] = {{0, 1, 2}, {1, 3, 4}, {2, 4, 5}};
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) O.m[i][j] = info[map[i][j]];
  prior->setInformationMatrix(O);
  std::ofstream os(out, std::ios::binary);
  for (int p = 0; p < n_pairs; ++p) {
    int32_t n[3];
    is.read((char*) n, sizeof(n));
    PointNormal2fVectorCloud f0 = readCloud(is, n[0]), f1 = readCloud(is, n[1]), mv = readCloud(is, n[2]);
    float v[9];
    is.read((char*) v, sizeof(v));
    if (!is.good()) throw std::runtime_error("short read on " + in);
    Isometry2f odom_fixed  = geometry2d::v2t(Vector3f(v[3], v[4], v[5]));
    Isometry2f odom_moving = geometry2d::v2t(Vector3f(v[6], v[7], v[8]));
    PropertyContainerDynamic fixed_scene, moving_scene;
    fixed_scene.setCloud(lasers[0]->param_fixed_slice_name.value(), &f0);
    fixed_scene.setCloud(lasers[1]->param_fixed_slice_name.value(), &f1);
    moving_scene.setCloud(lasers[0]->param_moving_slice_name.value(), &mv);
    for (int pass = 0; pass < 2; ++pass) {  // pass 0: with the odometry slice bound, pass 1: scenes without odometry
      if (pass == 0) {
        fixed_scene.setPose(prior->param_fixed_slice_name.value(), &odom_fixed);
        moving_scene.setPose(prior->param_moving_slice_name.value(), &odom_moving);
      } else {
        fixed_scene.setPose(prior->param_fixed_slice_name.value(), nullptr);
      }
      aligner->setFixed(&fixed_scene);
      aligner->setMoving(&moving_scene);
      aligner->setMovingInFixed(geometry2d::v2t(Vector3f(v[0], v[1], v[2])));
      aligner->compute();
      const auto& st = aligner->iterationStats();
      const IterationStats last = st.empty() ? IterationStats() : st.back();
      writeResult(os, geometry2d::t2v(aligner->movingInFixed()), (int) aligner->status(), last.chi_inliers,
                  last.chi_kernelized, last.num_inliers, last.num_outliers, last.num_correspondences, (int) st.size());
      const int32_t nc[2] = {(int32_t) lasers[0]->correspondences().size(), (int32_t) lasers[1]->correspondences().size()};
      os.write((const char*) nc, sizeof(nc));
      const Matrix3f& H = aligner->informationMatrix();
      const float h6[6] = {H.m[0][0], H.m[0][1], H.m[0][2], H.m[1][1], H.m[1][2], H.m[2][2]};
      os.write((const char*) h6, sizeof(h6));
    }
  }
  std::printf("MULTI OK %d pairs\n", n_pairs);
  return 0;
}

static void writeCloud(std::ofstream& os, const PointNormal2fVectorCloud& c) {
  const int32_t n = (int32_t) c.size();
  os.write((const char*) &n, 4);
  for (const auto& p : c) {
    const float f[4] = {p.coordinates().x(), p.coordinates().y(), p.normal().x(), p.normal().y()};
    os.write((const char*) f, sizeof(f));
  }
}

// (GPU) the reference's clipper / merger call sequences (apps/visual_test_merger_projective_2d.cpp:103-123)
// through the plugin classes, canvas = cols columns over the full circle
static int map(int cols, const std::string& in, const std::string& out) {
  PairsFile f = readPairs(in);
  SceneClipperProjective2DPtr clipper(new SceneClipperProjective2D);
  clipper->param_projector->param_canvas_cols.setValue((unsigned) cols);
  clipper->param_voxelize_resolution.setValue(0.f);
  MergerProjective2DPtr merger(new MergerProjective2D);
  merger->param_projector->param_canvas_cols.setValue((unsigned) cols);
  const Isometry2f sensor = geometry2d::v2t(Vector3f(f.sensor[0], f.sensor[1], f.sensor[2]));
  std::ofstream os(out, std::ios::binary);
  for (int p = 0; p < f.n_pairs; ++p) {
    PointNormal2fVectorCloud clipped;
    clipper->setFullScene(&f.fixed[p]);
    clipper->setClippedSceneInRobot(&clipped);
    clipper->setRobotInLocalMap(f.guesses[p][0]);
    clipper->setSensorInRobot(sensor);
    clipper->compute();
    if (clipper->status() != SceneClipperProjective2D::Successful) return 1;
    writeCloud(os, clipped);
    PointNormal2fVectorCloud scene = f.fixed[p];
    merger->setScene(&scene);
    merger->setMeasurement(&f.moving[p]);
    merger->setMeasurementInScene(f.guesses[p][0]);
    merger->compute();
    writeCloud(os, scene);
  }
  SceneClipperProjective2D unwired;
  if (thrown([&] { unwired.param_projector.setValue(nullptr); PointNormal2fVectorCloud a, b; unwired.setFullScene(&a);
                   unwired.setClippedSceneInRobot(&b); unwired.compute(); }) !=
      "SceneClipperProjective2D::compute| Missing Projector")
    return 1;
  std::printf("MAP OK %d\n", f.n_pairs);
  return 0;
}

// (GPU) the reference's adaptor call sequence (tests/test_measurement_adaptor.cpp:12-33) for every scan of the file
static int scan(float voxel_res, const std::string& in, const std::string& out) {
  std::ifstream is(in, std::ios::binary);
  if (!is.good()) throw std::runtime_error("cannot open " + in);
  int32_t n_scans = 0, n_beams = 0;
  float amin = 0.f, amax = 0.f;
  is.read((char*) &n_scans, 4), is.read((char*) &n_beams, 4), is.read((char*) &amin, 4), is.read((char*) &amax, 4);
  RawDataPreprocessorProjective2D adaptor;
  adaptor.param_voxelize_resolution.setValue(voxel_res);
  std::ofstream os(out, std::ios::binary);
  for (int s = 0; s < n_scans; ++s) {
    LaserMessagePtr msg(new LaserMessage("/scan"));
    msg->angle_min.setValue(amin), msg->angle_max.setValue(amax);
    msg->range_min.setValue(0.f), msg->range_max.setValue(30.f);
    msg->ranges.value().resize((size_t) n_beams);
    is.read((char*) msg->ranges.value().data(), (std::streamsize)(sizeof(float) * (size_t) n_beams));
    PointNormal2fVectorCloud points;
    adaptor.setMeas(&points);
    if (!adaptor.setRawData(msg)) return 1;
    adaptor.compute();
    if (adaptor.status() != RawDataPreprocessorProjective2D::Ready) return 1;
    writeCloud(os, points);
  }
  std::printf("SCAN OK %d\n", n_scans);
  return 0;
}

int main(int argc, char** argv) {
  try {
    const std::string cmd = argc > 1 ? argv[1] : "";
    if (cmd == "selftest") return selftest();
    if (cmd == "latency" && argc == 6) return latency(argv[2], argv[3], argv[4], std::atoi(argv[5]));
    if (cmd == "parse" && argc == 3) return parse(argv[2]);
    if (cmd == "align" && argc == 6) return align(argv[2], argv[3], argv[4], argv[5]);
    if (cmd == "verify" && argc == 6) return verify(argv[2], argv[3], argv[4], argv[5]);
    if (cmd == "multi" && argc == 6) return multi(argv[2], argv[3], argv[4], argv[5]);
    if (cmd == "map" && argc == 5) return map(std::atoi(argv[2]), argv[3], argv[4]);
    if (cmd == "scan" && argc == 5) return scan((float) std::atof(argv[2]), argv[3], argv[4]);
    std::fprintf(stderr, "usage: plugin_test selftest | parse <config> | align|verify <config> <name> <in> <out> | latency <config> <name> <in> <reps>\n");
    return 2;
  } catch (const std::exception& e) {
    std::fprintf(stderr, "plugin_test: %s\n", e.what());
    return 1;
  }
}
