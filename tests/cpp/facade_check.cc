// facade_check.cc — drives the reference-compatible C++ facade (semantic-icp_b200/facade/*.h) the way the reference's
// exec/test_icp.cc:22-124 and exec/kitti_eval.cc:184-231 drive the reference: pcl_2_semantic -> SemanticICP,
// GICP, EmIterativeClosestPoint<11>, getFusedLabels; prints one JSON object that tests/test_facade.py compares with the
// oracle.  Input file (written by the test): int32 ns, nt, N(=11); float xyz_s[ns][3]; uint32 lab_s[ns];
// float xyz_t[nt][3]; uint32 lab_t[nt]; double cm[N][N]; double init7[7].
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <fstream>

#include <em_icp.h>
#include <gicp.h>
#include <pcl_2_semantic.h>
#include <semantic_icp.h>

static void print_pose(const char* name, Sophus::SE3d T, int outer, bool last = false) {
  const double* p = T.data();
  std::printf("\"%s\": {\"pose\": [%.17g, %.17g, %.17g, %.17g, %.17g, %.17g, %.17g], \"outer_iter\": %d}%s\n", name, p[0], p[1], p[2], p[3], p[4],
              p[5], p[6], outer, last ? "" : ",");
}

int main(int argc, char** argv) {
  if (argc < 2) { std::fprintf(stderr, "usage: facade_check <input.bin>\n"); return 2; }
  std::ifstream f(argv[1], std::ios::binary);
  int32_t hdr[3];
  f.read((char*)hdr, sizeof hdr);
  const int ns = hdr[0], nt = hdr[1];
  constexpr size_t N = 11;
  if (hdr[2] != (int)N) { std::fprintf(stderr, "this check is compiled for N=11\n"); return 2; }
  std::vector<float> xs(3 * ns), xt(3 * nt);
  std::vector<uint32_t> ls(ns), lt(nt);
  std::vector<double> cm(N * N);
  double init7[7];
  f.read((char*)xs.data(), 12 * ns); f.read((char*)ls.data(), 4 * ns);
  f.read((char*)xt.data(), 12 * nt); f.read((char*)lt.data(), 4 * nt);
  f.read((char*)cm.data(), 8 * N * N); f.read((char*)init7, 56);
  if (!f) { std::fprintf(stderr, "short input file\n"); return 2; }

  typedef pcl::PointCloud<pcl::PointXYZL> CloudL;
  typedef pcl::PointCloud<pcl::PointXYZ> Cloud;
  CloudL::Ptr cloudA(new CloudL()), cloudB(new CloudL());
  Cloud::Ptr xyzA(new Cloud()), xyzB(new Cloud());
  for (int i = 0; i < ns; i++) { pcl::PointXYZL p; p.x = xs[3 * i]; p.y = xs[3 * i + 1]; p.z = xs[3 * i + 2]; p.label = ls[i]; cloudA->push_back(p); xyzA->push_back(pcl::PointXYZ(p.x, p.y, p.z)); }
  for (int i = 0; i < nt; i++) { pcl::PointXYZL p; p.x = xt[3 * i]; p.y = xt[3 * i + 1]; p.z = xt[3 * i + 2]; p.label = lt[i]; cloudB->push_back(p); xyzB->push_back(pcl::PointXYZ(p.x, p.y, p.z)); }
  Sophus::SE3d init;
  std::memcpy(init.data(), init7, 56);

  try {
    std::printf("{\n");
    // --- SemanticICP (exec/test_icp.cc:45-85)
    std::shared_ptr<semanticicp::SemanticPointCloud<pcl::PointXYZ, uint32_t>> semA(new semanticicp::SemanticPointCloud<pcl::PointXYZ, uint32_t>());
    std::shared_ptr<semanticicp::SemanticPointCloud<pcl::PointXYZ, uint32_t>> semB(new semanticicp::SemanticPointCloud<pcl::PointXYZ, uint32_t>());
    semanticicp::pcl_2_semantic(cloudA, semA);
    semanticicp::pcl_2_semantic(cloudB, semB);
    std::printf("\"labels_first_appearance\": [");
    for (size_t i = 0; i < semA->semanticLabels.size(); i++) std::printf("%s%u", i ? ", " : "", semA->semanticLabels[i]);
    std::printf("],\n");
    {  // public covariance map: first class, first point
      const uint32_t l0 = semA->semanticLabels[0];
      const Eigen::Matrix3d& c0 = semA->labeledCovariances[l0]->at(0);
      std::printf("\"class0_cov0\": [%.17g, %.17g, %.17g, %.17g, %.17g, %.17g, %.17g, %.17g, %.17g],\n", c0(0, 0), c0(0, 1), c0(0, 2), c0(1, 0), c0(1, 1),
                  c0(1, 2), c0(2, 0), c0(2, 1), c0(2, 2));
      std::vector<int> idx; std::vector<float> d2;
      const int found = semA->labeledKdTrees[l0]->nearestKSearch(semA->labeledPointClouds[l0]->points[0], 3, idx, d2);
      std::printf("\"class0_knn3\": {\"found\": %d, \"idx\": [%d, %d, %d], \"d2\": [%.9g, %.9g, %.9g]},\n", found, idx[0], idx[1], idx[2], d2[0], d2[1], d2[2]);
    }
    std::shared_ptr<semanticicp::SemanticPointCloud<pcl::PointXYZ, uint32_t>> semFinal(new semanticicp::SemanticPointCloud<pcl::PointXYZ, uint32_t>());
    semanticicp::pcl_2_semantic(cloudA, semFinal);
    semanticicp::SemanticIterativeClosestPoint<pcl::PointXYZ, uint32_t> sicp;
    sicp.setInputSource(semA);
    sicp.setInputTarget(semB);
    sicp.align(semFinal, init);
    print_pose("semantic", sicp.getFinalTransFormation(), sicp.getOuterIter());

    // --- GICP (exec/test_icp.cc:94-103)
    Cloud::Ptr finalGicp(new Cloud());
    semanticicp::GICP<pcl::PointXYZ> gicp;
    gicp.setSourceCloud(xyzA);
    gicp.setTargetCloud(xyzB);
    gicp.align(finalGicp, init);
    print_pose("gicp", gicp.getFinalTransFormation(), gicp.getOuterIter());
    std::printf("\"gicp_final0\": [%.9g, %.9g, %.9g],\n", finalGicp->points[0].x, finalGicp->points[0].y, finalGicp->points[0].z);
    std::printf("\"gicp_src_cov_n\": %zu,\n", gicp.getSourceCovariances()->size());

    // --- EM-ICP (exec/kitti_eval.cc:184-193) + fused labels (exec/scenenet_eval.cc:193-198)
    Eigen::Matrix<double, N, N> CM;
    for (size_t b = 0; b < N; b++) for (size_t s = 0; s < N; s++) CM((int)b, (int)s) = cm[b * N + s];
    semanticicp::EmIterativeClosestPoint<N> em;
    em.setSourceCloud(cloudA);
    em.setTargetCloud(cloudB);
    em.setConfusionMatrix(CM);
    CloudL::Ptr finalEm(new CloudL());
    em.align(finalEm, init);
    CloudL::Ptr fused(new CloudL());
    em.getFusedLabels(fused, em.getFinalTransFormation());
    size_t same = 0;
    for (int i = 0; i < ns; i++) same += fused->points[i].label == cloudA->points[i].label;
    std::printf("\"fused_n\": %zu, \"fused_same_as_input\": %zu,\n", fused->size(), same);
    print_pose("em", em.getFinalTransFormation(), em.getOuterIter(), true);
    std::printf("}\n");
  } catch (const std::exception& e) {
    std::fprintf(stderr, "facade_check failed: %s\n", e.what());
    return 1;
  }
  return 0;
}
