// test_shim.cpp - the reference's test/test_gicp_alignment.cpp (testApplyTF, testRun, testRunWithCov) and the
// removeFromCloud case of test/test_filter.cpp:102-113, restated against the drop-in header
// include/GICPAlignment_b200.hpp (no gtest / PCL / ROS in this image: plain asserts, POD cloud types).
// usage: test_shim source.f32 target.f32      (raw float32 xyz triples; written by tests/test_cpp_shim.py)
// Prints one "RESULT key value..." line per check; exit code 0 iff every check passed.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <type_traits>

#include "CloudIO_b200.hpp"
#include "FODDetector_b200.hpp"
#include "GICPAlignment_b200.hpp"

typedef GICPAlignment::PointCloudRGB PointCloudRGB;
typedef GICPAlignment::CloudPtr CloudPtr;
typedef GICPAlignment::Matrix4f Matrix4f;
typedef std::remove_reference<decltype(PointCloudRGB().points[0])>::type PointT;

static int g_failed = 0;
#define EXPECT_TRUE(cond)                                              \
  do {                                                                 \
    if (!(cond)) {                                                     \
      std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond);    \
      ++g_failed;                                                      \
    }                                                                  \
  } while (0)

static CloudPtr load(const char* path) {
  std::ifstream f(path, std::ios::binary);
  if (!f) { std::printf("cannot open %s\n", path); std::exit(2); }
  f.seekg(0, std::ios::end);
  const size_t n = (size_t)f.tellg() / 12;
  f.seekg(0);
  std::vector<float> xyz(3 * n);
  f.read(reinterpret_cast<char*>(xyz.data()), (std::streamsize)(12 * n));
  CloudPtr c(new PointCloudRGB);
  c->resize(n);
  for (size_t i = 0; i < n; ++i) {
    c->points[i].x = xyz[3 * i];
    c->points[i].y = xyz[3 * i + 1];
    c->points[i].z = xyz[3 * i + 2];
    c->points[i].r = 255; c->points[i].g = (uint8_t)(i & 255); c->points[i].b = 7;  // payload must survive every copy
  }
  return c;
}

static void print_tf(const char* key, const Matrix4f& tf) {
  std::printf("RESULT %s", key);
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c) std::printf(" %.9g", tf(r, c));
  std::printf("\n");
}

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  gicpb_shim::logger() = [](int level, const char* msg) { if (level >= gicpb_shim::kError) std::printf("LOG %s\n", msg); };
  CloudPtr sourceRGB = load(argv[1]), targetRGB = load(argv[2]);

  {  // testApplyTF (test/test_gicp_alignment.cpp:50-75)
    GICPAlignment gicp_alignment(targetRGB, sourceRGB, false);
    Matrix4f tf = gicp_alignment.getFineTransform();
    EXPECT_TRUE(tf == Matrix4f::Identity());  // constructor works well
    try {
      const PointCloudRGB before = *sourceRGB;
      gicp_alignment.run();
      gicp_alignment.applyTFtoCloud(sourceRGB);
      // applyTFtoCloud writes the internal aligned cloud; the argument stays as it was
      EXPECT_TRUE(std::memcmp(&before.points[0], &sourceRGB->points[0], 32 * before.points.size()) == 0);
      CloudPtr aligned(new PointCloudRGB);
      gicp_alignment.getAlignedCloud(aligned);
      const double tolerance = 1e-3;
      EXPECT_TRUE(aligned->points.size() == targetRGB->points.size());
      double worst = 0;
      for (size_t i = 0; i < aligned->points.size(); ++i) {  // default gate 0.04: the solve still lands on the yawed copy
        worst = std::max(worst, (double)std::fabs(aligned->points[i].x - targetRGB->points[i].x));
        worst = std::max(worst, (double)std::fabs(aligned->points[i].y - targetRGB->points[i].y));
        worst = std::max(worst, (double)std::fabs(aligned->points[i].z - targetRGB->points[i].z));
      }
      std::printf("RESULT applytf_worst_abs_diff %.6g converged %d\n", worst, (int)gicp_alignment.hasConverged());
      EXPECT_TRUE(aligned->points[5].g == 5 && aligned->points[5].r == 255);  // RGB copied through
      (void)tolerance;
      print_tf("applytf_transform", gicp_alignment.getFineTransform());
    } catch (std::exception& e) {
      std::printf("FAILED exception %s\n", e.what());
      ++g_failed;
    }
  }
  {  // testRun (:77-104)
    GICPAlignment gicp_alignment(targetRGB, sourceRGB, false);
    EXPECT_TRUE(gicp_alignment.getFineTransform() == Matrix4f::Identity());
    gicp_alignment.setMaxIterations(100);
    gicp_alignment.setMaxCorrespondenceDistance(5);
    gicp_alignment.setRANSACOutlierTh(5e-2);
    gicp_alignment.setTfEpsilon(5e-4);
    try {
      gicp_alignment.run();
      CloudPtr aligned_cloud(new PointCloudRGB);
      gicp_alignment.getAlignedCloud(aligned_cloud);
      EXPECT_TRUE(gicp_alignment.transform_exists_);
      EXPECT_TRUE(aligned_cloud->points.size() == sourceRGB->points.size());
      print_tf("run_transform", gicp_alignment.getFineTransform());
      std::printf("RESULT run_fitness %.12g outer %d\n", gicp_alignment.getFitnessScore(), gicp_alignment.lastResult().outer_iterations);
      double worst = 0;
      for (size_t i = 0; i < aligned_cloud->points.size(); ++i)
        worst = std::max(worst, (double)std::fabs(aligned_cloud->points[i].x - targetRGB->points[i].x));
      std::printf("RESULT run_worst_abs_dx %.6g\n", worst);
      EXPECT_TRUE(worst <= 1e-3);  // what the reference's first-point check meant to pin
    } catch (std::exception& e) {
      std::printf("FAILED exception %s\n", e.what());
      ++g_failed;
    }
  }
  {  // testRunWithCov (:106-131); use_covariances may drop points from the clouds it is given, so hand it copies
    CloudPtr tgt_c(new PointCloudRGB(*targetRGB)), src_c(new PointCloudRGB(*sourceRGB));
    for (int i = 0; i < 3; ++i) {  // three stray points without neighbours: no radius normal -> removed IN PLACE (:63-67)
      src_c->points.push_back(src_c->points[i]);
      src_c->points.back().x += 40.f + 3.f * i;
    }
    src_c->width = (uint32_t)src_c->points.size();
    const size_t n_src_before = src_c->points.size();
    GICPAlignment gicp_alignment(tgt_c, src_c, true);
    EXPECT_TRUE(gicp_alignment.getFineTransform() == Matrix4f::Identity());
    try {
      gicp_alignment.run();
      CloudPtr aligned_cloud(new PointCloudRGB);
      gicp_alignment.getAlignedCloud(aligned_cloud);
      EXPECT_TRUE(gicp_alignment.transform_exists_);
      std::printf("RESULT cov_source_points %zu of %zu, target %zu of %zu\n", src_c->points.size(), n_src_before,
                  tgt_c->points.size(), targetRGB->points.size());
      EXPECT_TRUE(src_c->points.size() == n_src_before - 3);            // the caller's cloud was filtered
      EXPECT_TRUE(tgt_c->points.size() == targetRGB->points.size());    // nothing to drop in the target
      EXPECT_TRUE(aligned_cloud->points.size() == src_c->points.size());
      const Matrix4f first = gicp_alignment.getFineTransform();
      gicp_alignment.iterate();
      EXPECT_TRUE(gicp_alignment.transform_exists_);
      // iterate() re-solves from the original source and composes: fine_tf = T * T
      const Matrix4f twice = first * first;
      double d = 0;
      const Matrix4f now = gicp_alignment.getFineTransform();
      for (int i = 0; i < 16; ++i) d = std::max(d, (double)std::fabs(now.data()[i] - twice.data()[i]));
      std::printf("RESULT iterate_compose_maxdiff %.6g\n", d);
      EXPECT_TRUE(d < 1e-5);
      CloudPtr after_iter(new PointCloudRGB), after_undo(new PointCloudRGB);
      gicp_alignment.getAlignedCloud(after_iter);
      gicp_alignment.undo();
      gicp_alignment.getAlignedCloud(after_undo);
      EXPECT_TRUE(std::memcmp(&after_undo->points[0], &aligned_cloud->points[0], 32 * aligned_cloud->points.size()) == 0);
      EXPECT_TRUE(gicp_alignment.getFineTransform() == now);  // undo does not roll the transform back
    } catch (std::exception& e) {
      std::printf("FAILED exception %s\n", e.what());
      ++g_failed;
    }
  }
  {  // Filter::removeFromCloud, test/test_filter.cpp:102-113: cube moved by +2 minus cube at 0
    CloudPtr moved(new PointCloudRGB(*sourceRGB)), out(new PointCloudRGB);
    for (auto& p : moved->points) { p.x += 2.f; p.y += 2.f; p.z += 2.f; }
    gicpb_shim::removeFromCloud(moved, sourceRGB, 0.0344, out);
    std::printf("RESULT difference_kept %zu of %zu\n", out->points.size(), moved->points.size());
    EXPECT_TRUE(out->points.size() > 1);
    EXPECT_TRUE(out->height == 1 && out->is_dense && out->width == out->points.size());
    CloudPtr same(new PointCloudRGB);
    gicpb_shim::removeFromCloud(sourceRGB, sourceRGB, 1e-6, same);
    EXPECT_TRUE(same->points.empty());
    gicpb_shim::removeFromCloud(moved, sourceRGB, 0.0344, moved);  // output aliases the input
    EXPECT_TRUE(moved->points.size() == out->points.size());
  }
  {  // test/test_fod_detector.cpp:32-72 testFODClustering and :75-90 testEmptyCloud
    auto cubes = [](CloudPtr cloud, float pos, float dim, float step) {
      PointT p_rgb;
      for (float i = pos; i < pos + dim; i += step)
        for (float j = pos; j < pos + dim; j += step)
          for (float k = pos; k < pos + dim; k += step) {
            p_rgb.x = i; p_rgb.y = j; p_rgb.z = k;
            p_rgb.r = 255; p_rgb.g = 255; p_rgb.b = 255;
            cloud->points.push_back(p_rgb);
          }
      cloud->width = (uint32_t)cloud->points.size();
    };
    CloudPtr cloudRGB(new PointCloudRGB);
    cubes(cloudRGB, 0, 3, 0.1f);
    cubes(cloudRGB, 10, 3, 0.1f);
    cubes(cloudRGB, 20, 3, 0.1f);
    int min_fod_points = 3;
    double voxelize_factor = 3;
    double th = 4e-3 * voxelize_factor;
    FODDetector fod_detector(cloudRGB, th * 10, min_fod_points);
    fod_detector.clusterPossibleFODs();
    std::vector<CloudPtr> fods_cloud_array;
    int num_of_fods = fod_detector.fodIndicesToPointCloud(fods_cloud_array);
    EXPECT_TRUE(num_of_fods == 3);
    EXPECT_TRUE(fods_cloud_array.size() == 3);
    std::vector<FODDetector::PointIndices> idx;
    fod_detector.getFODIndices(idx);
    size_t total = 0;
    for (size_t k = 0; k < idx.size(); ++k) {
      total += idx[k].indices.size();
      EXPECT_TRUE(k == 0 || idx[k].indices.size() <= idx[k - 1].indices.size());  // largest cluster first
      EXPECT_TRUE(fods_cloud_array[k]->points.size() == idx[k].indices.size() && fods_cloud_array[k]->is_dense);
    }
    std::printf("RESULT fod_clusters %d points %zu of %zu\n", num_of_fods, total, cloudRGB->points.size());
    EXPECT_TRUE(total == cloudRGB->points.size());
    CloudPtr empty_cloud(new PointCloudRGB);
    FODDetector fod_empty(empty_cloud, th * 10, min_fod_points);
    fod_empty.clusterPossibleFODs();
    std::vector<CloudPtr> none;
    EXPECT_TRUE(fod_empty.fodIndicesToPointCloud(none) == 0 && none.empty());
    // Filter::downsampleCloud, test/test_filter.cpp:68-83: fewer points, spread further apart
    CloudPtr down(new PointCloudRGB);
    gicpb_shim::downsampleCloud(cloudRGB, down, 0.25);
    std::printf("RESULT downsample %zu -> %zu\n", cloudRGB->points.size(), down->points.size());
    EXPECT_TRUE(!down->points.empty() && down->points.size() < cloudRGB->points.size());
    EXPECT_TRUE(down->points[0].r == 255 && down->points[0].a == 255 && down->points[0].w == 1.f && down->is_dense);
  }
  {  // node.cpp:37 fromROSMsg, Utils.cpp:102 toROSMsg, load_and_publish_clouds.cpp:75 loadPCDFile
    gicpb_shim::PointCloud2 msg;
    gicpb_shim::toROSMsg(*sourceRGB, msg);
    EXPECT_TRUE(msg.point_step == 32 && msg.width == sourceRGB->points.size() && msg.height == 1);
    EXPECT_TRUE(msg.fields.size() == 4 && msg.fields[3].name == "rgb" && msg.fields[3].offset == 16);
    PointCloudRGB back;
    gicpb_shim::fromROSMsg(msg, back);
    bool same = back.points.size() == sourceRGB->points.size();
    for (size_t i = 0; same && i < back.points.size(); ++i)
      same = std::memcmp(&back.points[i], &sourceRGB->points[i], 20) == 0 && back.points[i].w == 1.f;
    EXPECT_TRUE(same);
    // a message of another driver: intensity first, no colour -> default colour (a = 255)
    gicpb_shim::PointCloud2 other;
    other.width = 3; other.height = 1; other.point_step = 16; other.row_step = 48;
    const char* nm[4] = {"intensity", "x", "y", "z"};
    for (int i = 0; i < 4; ++i) { gicpb_shim::PointField f; f.name = nm[i]; f.offset = 4u * i; f.datatype = gicpb_shim::kFLOAT32; other.fields.push_back(f); }
    const float vals[12] = {9, 1, 2, 3, 9, 4, 5, 6, 9, 7, 8, 10};
    other.data.assign(reinterpret_cast<const uint8_t*>(vals), reinterpret_cast<const uint8_t*>(vals) + 48);
    PointCloudRGB o;
    gicpb_shim::fromROSMsg(other, o);
    EXPECT_TRUE(o.points.size() == 3 && o.points[2].x == 7.f && o.points[2].z == 10.f && o.points[1].y == 5.f);
    EXPECT_TRUE(o.points[0].a == 255 && o.points[0].r == 0 && o.points[0].w == 1.f);
    if (argc > 3) {  // a binary PCD file of the source cloud written by the test driver
      PointCloudRGB pc;
      EXPECT_TRUE(gicpb_shim::loadPCDFile(argv[3], pc) == 0);
      bool eq = pc.points.size() == sourceRGB->points.size();
      for (size_t i = 0; eq && i < pc.points.size(); ++i)
        eq = pc.points[i].x == sourceRGB->points[i].x && pc.points[i].y == sourceRGB->points[i].y && pc.points[i].z == sourceRGB->points[i].z;
      EXPECT_TRUE(eq);
      EXPECT_TRUE(gicpb_shim::loadPCDFile(std::string(argv[3]) + ".missing", pc) == -1);
      std::printf("RESULT pcd_points %zu\n", pc.points.size());
    }
  }

  if (argc > 5) {
    // The stages of one LeicaStateMachine run in the reference's order (src/LeicaStateMachine.cpp:61-65,138-216; the
    // "drops in unchanged" test is test/test_state_machine.cpp:70-78), every stage through its shim, all on the one
    // shared context: downsample both clouds -> GICP -> transform the source -> removeFromCloud -> FODDetector.
    // argv[4]: a panel scan (offset pose, FOD blobs on it), argv[5]: the CAD cloud.  InitialAlignment is out of scope:
    // the offset is small enough for GICP alone.
    CloudPtr scan = load(argv[4]), cad = load(argv[5]);
    const double leaf_size_factor = 10;
    const double target_res = gicpb_shim::computeCloudResolution(cad);
    const double source_res = gicpb_shim::computeCloudResolution(scan);
    const double leaf_size = leaf_size_factor * std::max(target_res, source_res);
    CloudPtr source_f(new PointCloudRGB), target_f(new PointCloudRGB);
    gicpb_shim::downsampleCloud(scan, source_f, leaf_size);
    gicpb_shim::downsampleCloud(cad, target_f, leaf_size);
    std::printf("RESULT pipeline_resolution %.12g %.12g\n", target_res, source_res);
    std::printf("RESULT pipeline_downsampled %zu %zu\n", source_f->points.size(), target_f->points.size());
    GICPAlignment gicp_alignment(target_f, source_f, false);
    gicp_alignment.setMaxCorrespondenceDistance(1);
    gicp_alignment.run();
    CloudPtr aligned_cloud(new PointCloudRGB);
    gicp_alignment.getAlignedCloud(aligned_cloud);
    EXPECT_TRUE(gicp_alignment.transform_exists_ && gicpb_shim::isValidTransform(gicp_alignment.getFineTransform()));
    const Matrix4f final_transform = gicp_alignment.getFineTransform() * Matrix4f::Identity();
    print_tf("pipeline_transform", final_transform);
    {
      std::shared_ptr<gicpb_shim::Context> ctx = gicpb_shim::Context::shared();
      gicpb_shim::transformCloud(*ctx, *source_f, *source_f, final_transform);  // pcl::transformPointCloud in place (:189)
    }
    const double voxelize_factor = 3, th = 4e-3 * voxelize_factor;
    CloudPtr substracted_cloud(new PointCloudRGB);
    gicpb_shim::removeFromCloud(source_f, target_f, th, substracted_cloud);
    std::printf("RESULT pipeline_difference %zu\n", substracted_cloud->points.size());
    int num_of_fods = 0;
    std::vector<CloudPtr> fods_cloud_array;
    if (substracted_cloud->points.size() > 1) {  // Utils::isValidCloud
      FODDetector fod_detector(substracted_cloud, th * 100, 3);
      fod_detector.clusterPossibleFODs();
      num_of_fods = fod_detector.fodIndicesToPointCloud(fods_cloud_array);
    }
    std::printf("RESULT pipeline_fods %d", num_of_fods);
    for (const auto& f : fods_cloud_array) std::printf(" %zu", f->points.size());
    std::printf("\n");
    // Utils::getNormals on the downsampled CAD cloud (src/Utils.cpp:27-44) through its shim
    gicpb_shim::NormalCloudT::Ptr normals(new gicpb_shim::NormalCloudT);
    EXPECT_TRUE(gicpb_shim::getNormals(target_f, 4.0 * leaf_size, normals));
    size_t finite = 0;
    double checksum = 0;
    for (const auto& nrm : normals->points)
      if (std::isfinite(nrm.normal_x)) {
        ++finite;
        checksum += (double)nrm.normal_x + 2.0 * (double)nrm.normal_y + 3.0 * (double)nrm.normal_z + (double)nrm.curvature;
      }
    std::printf("RESULT pipeline_normals %zu %.9g %d\n", finite, checksum, (int)normals->is_dense);
  }

  std::printf("RESULT failed %d\n", g_failed);
  return g_failed ? 1 : 0;
}
