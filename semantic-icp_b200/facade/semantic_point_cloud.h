// semantic_point_cloud.h — semanticicp::SemanticPointCloud<PointT,SemanticT>, source-compatible with the reference's
// semantic_icp/semantic_point_cloud.h:15-63.  The public maps keep their host-side meaning (callers read them
// directly: exec/make_semantic.cc:37-39); the per-class kd-trees and k-neighbour PCA covariances are computed by
// libsicp_b200 on the device:
//   * one PER_CLASS device cloud holds every class (points sorted by (class, Morton code)), built lazily and cached —
//     this is what SemanticIterativeClosestPoint::align consumes;
//   * labeledCovariances[label] is filled from that cloud's covariances (neighbours restricted to the class, divisor
//     k even when the class is smaller: impl/semantic_point_cloud.hpp:41,60-64);
//   * labeledKdTrees[label] is a DeviceKdTree over the class's points.
#ifndef SICP_FACADE_SEMANTIC_POINT_CLOUD_H_
#define SICP_FACADE_SEMANTIC_POINT_CLOUD_H_
#include <algorithm>
#include <map>
#include "device_kdtree.h"

namespace semanticicp {

template <typename PointT, typename SemanticT>
class SemanticPointCloud {
 public:
  typedef std::shared_ptr<SemanticPointCloud<PointT, SemanticT>> Ptr;
  typedef std::shared_ptr<const SemanticPointCloud<PointT, SemanticT>> ConstPtr;
  typedef pcl::PointCloud<PointT> PointCloud;
  typedef typename PointCloud::Ptr PointCloudPtr;
  typedef DeviceKdTree<PointT> KdTree;
  typedef typename KdTree::Ptr KdTreePtr;
  typedef detail::MatricesVector MatricesVector;
  typedef std::shared_ptr<MatricesVector> MatricesVectorPtr;

  SemanticPointCloud(int k = 20, double epsilon = 0.001) : k_correspondences_(k), epsilon_(epsilon) {}  // semantic_point_cloud.h:31

  std::vector<SemanticT> semanticLabels;
  std::map<SemanticT, PointCloudPtr> labeledPointClouds;
  std::map<SemanticT, MatricesVectorPtr> labeledCovariances;
  std::map<SemanticT, KdTreePtr> labeledKdTrees;

  // impl/semantic_point_cloud.hpp:12-87
  void addSemanticCloud(SemanticT label, PointCloudPtr cloud_ptr, bool computeKd = true, bool computeCov = true) {
    semanticLabels.push_back(label);
    labeledPointClouds[label] = cloud_ptr;
    device_.reset();
    if (computeKd) {
      KdTreePtr tree(new KdTree());
      tree->setInputCloud(cloud_ptr);
      labeledKdTrees[label] = tree;
      if (computeCov) {  // the class tree is exactly the neighbour set the reference uses (hpp:41)
        const std::size_t n = cloud_ptr->points.size();
        detail::check(sicp_cloud_precompute(tree->handle().get(), k_correspondences_, epsilon_, 0, nullptr), "class covariances");
        std::vector<double> rows(9 * n);
        detail::check(sicp_cloud_get_covariances(tree->handle().get(), rows.data()), "covariance download");
        MatricesVectorPtr covs(new MatricesVector());
        detail::fill_matrices(rows, n, 0, covs.get());
        labeledCovariances[label] = covs;
      }
    }
  }
  // Extension used by pcl_2_semantic: all classes at once — ONE device cloud, one covariance pass, one download.
  void addSemanticClouds(const std::vector<SemanticT>& labels, const std::vector<PointCloudPtr>& clouds) {
    for (std::size_t i = 0; i < labels.size(); i++) {
      semanticLabels.push_back(labels[i]);
      labeledPointClouds[labels[i]] = clouds[i];
    }
    device_.reset();
    detail::CloudHandle h = device();
    std::size_t n = 0;
    detail::check(sicp_cloud_size(h.get(), &n), "cloud size");
    std::vector<double> rows(9 * n);
    detail::check(sicp_cloud_get_covariances(h.get(), rows.data()), "covariance download");
    std::size_t first = 0;
    for (SemanticT s : semanticLabels) {  // rows are in upload order = semanticLabels order, class by class
      const std::size_t nc = labeledPointClouds[s]->points.size();
      MatricesVectorPtr covs(new MatricesVector());
      detail::fill_matrices(rows, nc, first, covs.get());
      labeledCovariances[s] = covs;
      first += nc;
      // the public per-class tree (semantic_point_cloud.h:39): registered LAZILY — its device upload happens on the first
      // nearestKSearch, so a cloud that is only ever aligned (the PER_CLASS device cloud above already holds every class)
      // is not uploaded a second time class by class
      KdTreePtr tree(new KdTree());
      tree->setInputCloudLazy(labeledPointClouds[s]);
      labeledKdTrees[s] = tree;
    }
  }
  void removeSemanticClass(SemanticT label) {  // semantic_point_cloud.h:44-52
    auto it = std::find(semanticLabels.begin(), semanticLabels.end(), label);
    if (it != semanticLabels.end()) {
      semanticLabels.erase(it);
      labeledPointClouds.erase(label);
      labeledCovariances.erase(label);
      labeledKdTrees.erase(label);
      device_.reset();
    }
  }
  typename pcl::PointCloud<pcl::PointXYZL>::Ptr getpclPointCloud() {  // impl/semantic_point_cloud.hpp:89-103
    typename pcl::PointCloud<pcl::PointXYZL>::Ptr out(new pcl::PointCloud<pcl::PointXYZL>());
    for (SemanticT s : semanticLabels)
      for (const PointT& p : labeledPointClouds[s]->points) {
        pcl::PointXYZL q;
        q.x = p.x; q.y = p.y; q.z = p.z; q.label = std::uint32_t(s);
        out->push_back(q);
      }
    return out;
  }
  // impl/semantic_point_cloud.hpp:105-111 — moves the points only; like the reference, trees and covariances in the
  // public maps are NOT recomputed.  The cached device cloud is dropped, so the next align() sees
  // the moved points.
  void transform(Eigen::Matrix4f trans) {
    for (SemanticT s : semanticLabels) pcl::transformPointCloud(*labeledPointClouds[s], *labeledPointClouds[s], trans);
    device_.reset();
  }

  // The PER_CLASS device cloud over all current classes (class order = semanticLabels), covariances precomputed.
  detail::CloudHandle device() {
    if (device_) return device_;
    std::size_t n = 0;
    for (SemanticT s : semanticLabels) n += labeledPointClouds[s]->points.size();
    std::vector<float> xyz(3 * n);
    std::vector<std::uint32_t> lab(n);
    std::size_t i = 0;
    for (SemanticT s : semanticLabels)
      for (const PointT& p : labeledPointClouds[s]->points) {
        xyz[3 * i] = p.x; xyz[3 * i + 1] = p.y; xyz[3 * i + 2] = p.z;
        lab[i++] = std::uint32_t(s);
      }
    sicp_cloud* c = nullptr;
    static const float zero3[3] = {0.f, 0.f, 0.f};
    static const std::uint32_t zero1 = 0;
    detail::check(sicp_cloud_create(n ? xyz.data() : zero3, 12, n ? lab.data() : &zero1, 4, n, SICP_CLOUD_PER_CLASS, detail::device_index(), &c), "semantic cloud upload");
    device_ = detail::make_handle(c);
    detail::check(sicp_cloud_precompute(c, k_correspondences_, epsilon_, 0, nullptr), "semantic cloud covariances");
    return device_;
  }

 private:
  int k_correspondences_;
  double epsilon_;
  detail::CloudHandle device_;
};

}  // namespace semanticicp
#endif  // SICP_FACADE_SEMANTIC_POINT_CLOUD_H_
