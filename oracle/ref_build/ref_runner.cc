// ref_runner.cc — drives the UNMODIFIED reference classes on a binary fixture and prints what they computed as JSON.
// TEST INFRASTRUCTURE (oracle/ref_build): pins oracle/sicp_oracle.cpp to the reference itself wherever PCL, Eigen, Sophus
// and Ceres are installed (they are not in the build image).  Uses the reference exactly like its own drivers do:
//   GICP<PointXYZ>::setSourceCloud/setTargetCloud/align/getFinalTransFormation/getOuterIter   exec/kitti_eval.cc:208-221
//   EmIterativeClosestPoint<N>::set*Cloud/setConfusionMatrix/align                            exec/kitti_eval.cc:182-197
//   pcl_2_semantic + SemanticIterativeClosestPoint::setInput*/align                           exec/test_icp.cc:41-100
// align() reports its per-pass state only on stdout ("MSE:", "Transform:", "Itteration:", the Ceres report:
// impl/em_icp.hpp:177-185, impl/gicp.hpp:152-160), so stdout is captured and parsed for the per-pass poses and iterations.
//
// Fixture (little endian, written by make_ref_fixture.py):
//   int32 magic 0x53494350, int32 ns, int32 nt, int32 N, float64 init7[7] (qx qy qz qw tx ty tz), float64 cm[N*N] row-major,
//   float32 src_xyz[ns*3], uint32 src_lab[ns], float32 tgt_xyz[nt*3], uint32 tgt_lab[nt]
// usage: sicp_ref_runner fixture.bin {gicp|em|semantic} > out.json        (N must equal SICP_REF_N, default 11)
#include <cstdint>
#include <cstdio>
#include <fstream>
#include <iostream>
#include <memory>
#include <regex>
#include <sstream>
#include <string>
#include <vector>

#include <pcl/point_types.h>
#include <pcl/point_cloud.h>

#include <semantic_point_cloud.h>
#include <pcl_2_semantic.h>
#include <gicp.h>
#include <semantic_icp.h>
#include <em_icp.h>

#ifndef SICP_REF_N
#define SICP_REF_N 11
#endif

struct Fixture {
  int ns = 0, nt = 0, N = 0;
  double init7[7];
  std::vector<double> cm;
  std::vector<float> sxyz, txyz;
  std::vector<uint32_t> slab, tlab;
};

static bool load(const char* path, Fixture* f) {
  std::ifstream in(path, std::ios::binary);
  int32_t hdr[4];
  if (!in.read(reinterpret_cast<char*>(hdr), sizeof hdr) || hdr[0] != 0x53494350) return false;
  f->ns = hdr[1]; f->nt = hdr[2]; f->N = hdr[3];
  f->cm.resize((size_t)f->N * f->N); f->sxyz.resize(3 * (size_t)f->ns); f->slab.resize(f->ns); f->txyz.resize(3 * (size_t)f->nt); f->tlab.resize(f->nt);
  in.read(reinterpret_cast<char*>(f->init7), sizeof f->init7);
  in.read(reinterpret_cast<char*>(f->cm.data()), sizeof(double) * f->cm.size());
  in.read(reinterpret_cast<char*>(f->sxyz.data()), sizeof(float) * f->sxyz.size());
  in.read(reinterpret_cast<char*>(f->slab.data()), sizeof(uint32_t) * f->slab.size());
  in.read(reinterpret_cast<char*>(f->txyz.data()), sizeof(float) * f->txyz.size());
  in.read(reinterpret_cast<char*>(f->tlab.data()), sizeof(uint32_t) * f->tlab.size());
  return static_cast<bool>(in);
}

template <class PointT>
static typename pcl::PointCloud<PointT>::Ptr cloud_xyz(const std::vector<float>& xyz, const std::vector<uint32_t>* lab);
template <>
pcl::PointCloud<pcl::PointXYZ>::Ptr cloud_xyz<pcl::PointXYZ>(const std::vector<float>& xyz, const std::vector<uint32_t>*) {
  pcl::PointCloud<pcl::PointXYZ>::Ptr c(new pcl::PointCloud<pcl::PointXYZ>);
  for (size_t i = 0; i < xyz.size() / 3; i++) c->push_back(pcl::PointXYZ(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]));
  return c;
}
template <>
pcl::PointCloud<pcl::PointXYZL>::Ptr cloud_xyz<pcl::PointXYZL>(const std::vector<float>& xyz, const std::vector<uint32_t>* lab) {
  pcl::PointCloud<pcl::PointXYZL>::Ptr c(new pcl::PointCloud<pcl::PointXYZL>);
  for (size_t i = 0; i < xyz.size() / 3; i++) {
    pcl::PointXYZL p;
    p.x = xyz[3 * i]; p.y = xyz[3 * i + 1]; p.z = xyz[3 * i + 2]; p.label = (*lab)[i];
    c->push_back(p);
  }
  return c;
}

static void print_pose(const char* key, const Sophus::SE3d& T) {
  const double* d = T.data();  // [qx qy qz qw tx ty tz] (gicp_cost_function.h:64-70)
  std::printf("\"%s\": [%.17g, %.17g, %.17g, %.17g, %.17g, %.17g, %.17g]", key, d[0], d[1], d[2], d[3], d[4], d[5], d[6]);
}

// per-pass 4x4 transforms ("Transform:" + 4 rows) and Ceres iteration counts ("Minimizer iterations  N") from the log
static void print_passes(const std::string& log) {
  std::printf("\"pass_matrix\": [");
  std::istringstream in(log);
  std::string line;
  bool first = true;
  std::vector<int> iters;
  const std::regex it_full("Minimizer iterations\\s+(\\d+)"), it_brief("iterations: (\\d+)");
  while (std::getline(in, line)) {
    std::smatch m;
    if (std::regex_search(line, m, it_full) || std::regex_search(line, m, it_brief)) iters.push_back(std::stoi(m[1]));
    if (line.find("Transform:") != std::string::npos) {
      double v[16];
      bool ok = true;
      for (int r = 0; r < 4 && ok; r++) {
        if (!std::getline(in, line)) { ok = false; break; }
        std::istringstream row(line);
        for (int c = 0; c < 4; c++) ok = ok && static_cast<bool>(row >> v[4 * r + c]);
      }
      if (!ok) continue;
      std::printf("%s[", first ? "" : ", ");
      for (int i = 0; i < 16; i++) std::printf("%s%.9g", i ? ", " : "", v[i]);  // std::cout prints 6 significant digits: a coarse cross-check only
      std::printf("]");
      first = false;
    }
  }
  std::printf("], \"pass_lm_iters\": [");
  for (size_t i = 0; i < iters.size(); i++) std::printf("%s%d", i ? ", " : "", iters[i]);
  std::printf("]");
}

int main(int argc, char** argv) {
  if (argc < 3) { std::fprintf(stderr, "usage: %s fixture.bin {gicp|em|semantic}\n", argv[0]); return 2; }
  Fixture f;
  if (!load(argv[1], &f)) { std::fprintf(stderr, "bad fixture\n"); return 2; }
  const std::string algo = argv[2];
  Eigen::Quaterniond q(f.init7[3], f.init7[0], f.init7[1], f.init7[2]);
  Sophus::SE3d init(q, Eigen::Vector3d(f.init7[4], f.init7[5], f.init7[6]));
  std::ostringstream log;
  std::streambuf* old = std::cout.rdbuf(log.rdbuf());  // align() prints its per-pass state; keep it for parsing
  Sophus::SE3d result;
  int outer = -1;
  if (algo == "gicp") {
    semanticicp::GICP<pcl::PointXYZ> icp;
    pcl::PointCloud<pcl::PointXYZ>::Ptr fin(new pcl::PointCloud<pcl::PointXYZ>);
    icp.setSourceCloud(cloud_xyz<pcl::PointXYZ>(f.sxyz, nullptr));
    icp.setTargetCloud(cloud_xyz<pcl::PointXYZ>(f.txyz, nullptr));
    icp.align(fin, init);
    result = icp.getFinalTransFormation();
    outer = icp.getOuterIter();
  } else if (algo == "em") {
    if (f.N != SICP_REF_N) { std::cout.rdbuf(old); std::fprintf(stderr, "fixture has N=%d, runner built for N=%d\n", f.N, SICP_REF_N); return 2; }
    semanticicp::EmIterativeClosestPoint<SICP_REF_N> icp;
    Eigen::Matrix<double, SICP_REF_N, SICP_REF_N> cm;
    for (int r = 0; r < f.N; r++) for (int c = 0; c < f.N; c++) cm(r, c) = f.cm[(size_t)r * f.N + c];
    pcl::PointCloud<pcl::PointXYZL>::Ptr fin(new pcl::PointCloud<pcl::PointXYZL>);
    icp.setSourceCloud(cloud_xyz<pcl::PointXYZL>(f.sxyz, &f.slab));
    icp.setTargetCloud(cloud_xyz<pcl::PointXYZL>(f.txyz, &f.tlab));
    icp.setConfusionMatrix(cm);
    icp.align(fin, init);
    result = icp.getFinalTransFormation();
    outer = icp.getOuterIter();
  } else if (algo == "semantic") {
    typedef semanticicp::SemanticPointCloud<pcl::PointXYZ, uint32_t> SCloud;
    std::shared_ptr<SCloud> s(new SCloud()), t(new SCloud()), fin(new SCloud());
    semanticicp::pcl_2_semantic(cloud_xyz<pcl::PointXYZL>(f.sxyz, &f.slab), s);
    semanticicp::pcl_2_semantic(cloud_xyz<pcl::PointXYZL>(f.txyz, &f.tlab), t);
    semanticicp::pcl_2_semantic(cloud_xyz<pcl::PointXYZL>(f.sxyz, &f.slab), fin);
    semanticicp::SemanticIterativeClosestPoint<pcl::PointXYZ, uint32_t> icp;
    icp.setInputSource(s);
    icp.setInputTarget(t);
    icp.align(fin, init);
    result = icp.getFinalTransFormation();
  } else {
    std::cout.rdbuf(old);
    std::fprintf(stderr, "unknown algorithm %s\n", algo.c_str());
    return 2;
  }
  std::cout.rdbuf(old);
  std::printf("{\"algo\": \"%s\", \"ns\": %d, \"nt\": %d, \"N\": %d, ", algo.c_str(), f.ns, f.nt, f.N);
  print_pose("pose7", result);
  std::printf(", \"outer_iter\": %d, ", outer);
  print_passes(log.str());
  std::printf("}\n");
  return 0;
}
