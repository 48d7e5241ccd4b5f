// pair_eval.cc — what exec/kitti_eval.cc does for ONE scan pair, written against the drop-in facade only:
//   load two labelled PCD scans -> range filter (exec/kitti_eval.cc:124-127) -> EM-ICP and GICP (:184-231) -> SE(3) error
//   against a ground-truth pose (exec/kitti_metrics.h:31-37) -> label agreement of the aligned scan (exec/roc_metrics.h:21-41).
// Build:  g++ -std=c++11 -O2 -Isemantic-icp_b200/facade -Iinclude examples/pair_eval.cc -Lsemantic-icp_b200/lib -lsicp_b200
// Usage:  pair_eval source.pcd target.pcd confusion.txt [range_m] [gt: qx qy qz qw tx ty tz]
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>

#include <em_icp.h>
#include <gicp.h>

typedef pcl::PointCloud<pcl::PointXYZL> CloudL;
typedef pcl::PointCloud<pcl::PointXYZ> Cloud;
static const size_t N = 11;  // exec/kitti_eval.cc:67

// filterRange (exec/filter_range.h:6-18) through the device stream compaction instead of the O(n^2) erase loop
static void filter_range(CloudL::Ptr cloud, double range) {
  std::vector<uint32_t> keep(cloud->size());
  size_t n_keep = 0;
  semanticicp::detail::check(sicp_filter_range(cloud->points.empty() ? nullptr : &cloud->points[0].x, sizeof(pcl::PointXYZL), cloud->size(), range, 0,
                                               keep.data(), &n_keep), "filterRange");
  CloudL out;
  for (size_t i = 0; i < n_keep; i++) out.push_back(cloud->points[keep[i]]);
  *cloud = out;
}

int main(int argc, char** argv) {
  if (argc < 4) { std::fprintf(stderr, "usage: pair_eval source.pcd target.pcd confusion.txt [range_m] [gt pose7]\n"); return 2; }
  CloudL::Ptr cloudA(new CloudL()), cloudB(new CloudL());
  if (pcl::io::loadPCDFile(argv[1], *cloudA) < 0 || pcl::io::loadPCDFile(argv[2], *cloudB) < 0) { std::fprintf(stderr, "cannot read the PCD files\n"); return 1; }
  Eigen::Matrix<double, N, N> cm;
  { std::ifstream f(argv[3]); for (size_t r = 0; r < N; r++) for (size_t c = 0; c < N; c++) { double v = 0; f >> v; cm((int)r, (int)c) = v; } }  // read_confusion_matrix.h:6-19
  const double range = argc > 4 ? std::atof(argv[4]) : 40.0;
  try {
    filter_range(cloudA, range);
    filter_range(cloudB, range);
    std::printf("{\"n_source\": %zu, \"n_target\": %zu,\n", cloudA->size(), cloudB->size());

    semanticicp::EmIterativeClosestPoint<N> emicp;           // exec/kitti_eval.cc:184-193
    emicp.setSourceCloud(cloudA);
    emicp.setTargetCloud(cloudB);
    emicp.setConfusionMatrix(cm);
    CloudL::Ptr finalEm(new CloudL());
    Sophus::SE3d init;
    emicp.align(finalEm, init);
    const Sophus::SE3d Tem = emicp.getFinalTransFormation();

    Cloud::Ptr xyzA(new Cloud()), xyzB(new Cloud()), finalGicp(new Cloud());   // exec/kitti_eval.cc:205-231
    for (const pcl::PointXYZL& p : cloudA->points) xyzA->push_back(pcl::PointXYZ(p.x, p.y, p.z));
    for (const pcl::PointXYZL& p : cloudB->points) xyzB->push_back(pcl::PointXYZ(p.x, p.y, p.z));
    semanticicp::GICP<pcl::PointXYZ> gicp;
    gicp.setSourceCloud(xyzA);
    gicp.setTargetCloud(xyzB);
    gicp.align(finalGicp);
    const Sophus::SE3d Tg = gicp.getFinalTransFormation();

    const double* pe = Tem.data();
    const double* pg = Tg.data();
    std::printf("\"em\": {\"pose\": [%.17g, %.17g, %.17g, %.17g, %.17g, %.17g, %.17g], \"outer_iter\": %d},\n", pe[0], pe[1], pe[2], pe[3], pe[4], pe[5], pe[6], emicp.getOuterIter());
    std::printf("\"gicp\": {\"pose\": [%.17g, %.17g, %.17g, %.17g, %.17g, %.17g, %.17g], \"outer_iter\": %d},\n", pg[0], pg[1], pg[2], pg[3], pg[4], pg[5], pg[6], gicp.getOuterIter());
    if (argc >= 12) {                                      // KittiMetrics::evaluate error terms
      double gt[7], err[6];
      for (int i = 0; i < 7; i++) gt[i] = std::atof(argv[5 + i]);
      const double est[14] = {pe[0], pe[1], pe[2], pe[3], pe[4], pe[5], pe[6], pg[0], pg[1], pg[2], pg[3], pg[4], pg[5], pg[6]};
      const double gts[14] = {gt[0], gt[1], gt[2], gt[3], gt[4], gt[5], gt[6], gt[0], gt[1], gt[2], gt[3], gt[4], gt[5], gt[6]};
      semanticicp::detail::check(sicp_pose_errors(2, gts, est, err), "pose errors");
      std::printf("\"em_error\": [%.6e, %.6e, %.6e], \"gicp_error\": [%.6e, %.6e, %.6e],\n", err[0], err[1], err[2], err[3], err[4], err[5]);
    }
    // ROCMetrics::evaluate on the aligned source (finalEm) against the target
    semanticicp::DeviceKdTree<pcl::PointXYZL> aligned, target;
    aligned.setInputCloud(finalEm);
    target.setInputCloud(cloudB);
    std::vector<int64_t> confusion((N + 1) * (N + 1));
    double stats[3];
    semanticicp::detail::check(sicp_label_agreement(aligned.handle().get(), target.handle().get(), nullptr, 25.0, (int)N + 1, confusion.data(), stats, nullptr), "label agreement");
    std::printf("\"label_agreement\": {\"accuracy\": %.6f, \"pairs\": %.0f, \"mean_distance\": %.6f}}\n", stats[0] / stats[1], stats[1], stats[2] / stats[1]);
  } catch (const std::exception& e) {
    std::fprintf(stderr, "pair_eval failed: %s\n", e.what());
    return 1;
  }
  return 0;
}
