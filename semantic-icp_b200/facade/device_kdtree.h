// device_kdtree.h — semanticicp::DeviceKdTree<PointT>: the B200 counterpart of pcl::KdTreeFLANN<PointT> at the
// places the reference's API exposes a kd-tree (GICP::getSourceKdTree / setSourceCloud(cloud, tree, covs), gicp.h:50-90;
// SemanticPointCloud::labeledKdTrees, semantic_point_cloud.h:39).  setInputCloud uploads the points into device SoA
// buffers and builds the Morton-sorted box tree (sicp_cloud_create); nearestKSearch is the exact kNN of the C ABI
// (sicp_knn): k results ordered by (d2_f32, index), FLANN L2_Simple<float> distance arithmetic.
#ifndef SICP_FACADE_DEVICE_KDTREE_H_
#define SICP_FACADE_DEVICE_KDTREE_H_
#include "sicp_compat.h"

namespace semanticicp {

template <typename PointT>
class DeviceKdTree {
 public:
  typedef std::shared_ptr<DeviceKdTree<PointT>> Ptr;
  typedef std::shared_ptr<const DeviceKdTree<PointT>> ConstPtr;
  typedef typename pcl::PointCloud<PointT>::Ptr PointCloudPtr;

  DeviceKdTree() {}
  // pcl::KdTreeFLANN::setInputCloud: builds the search structure NOW (gicp.h:45-46, em_icp.h:53-54)
  void setInputCloud(const PointCloudPtr& cloud) {
    cloud_ = cloud;
    handle_ = detail::upload_whole(*cloud);
  }
  // Same, but the device upload / build is postponed to the first search (SemanticPointCloud::addSemanticClouds)
  void setInputCloudLazy(const PointCloudPtr& cloud) {
    cloud_ = cloud;
    handle_.reset();
  }
  PointCloudPtr getInputCloud() const { return cloud_; }
  // pcl::KdTreeFLANN::nearestKSearch: returns the number of neighbours found (k clamped to the cloud size)
  int nearestKSearch(const PointT& p, int k, std::vector<int>& k_indices, std::vector<float>& k_sqr_distances) const {
    ensure();
    const float q[3] = {p.x, p.y, p.z};
    std::vector<std::int32_t> idx(k);
    std::vector<float> d2(k);
    detail::check(sicp_knn(handle_.get(), q, nullptr, 1, nullptr, k, idx.data(), d2.data()), "nearestKSearch");
    int found = 0;
    while (found < k && idx[found] >= 0) found++;
    k_indices.assign(idx.begin(), idx.begin() + found);
    k_sqr_distances.assign(d2.begin(), d2.begin() + found);
    return found;
  }
  // Batched form (one launch for all queries): idx / d2 are nq x k, row-major, -1 / +inf where fewer than k exist.
  void nearestKSearchBatch(const pcl::PointCloud<PointT>& queries, int k, std::vector<int>& idx, std::vector<float>& d2) const {
    ensure();
    const std::size_t nq = queries.points.size();
    std::vector<float> q(3 * nq);
    for (std::size_t i = 0; i < nq; i++) { q[3 * i] = queries.points[i].x; q[3 * i + 1] = queries.points[i].y; q[3 * i + 2] = queries.points[i].z; }
    std::vector<std::int32_t> tmp(nq * (std::size_t)k);
    d2.resize(nq * (std::size_t)k);
    detail::check(sicp_knn(handle_.get(), q.data(), nullptr, nq, nullptr, k, tmp.data(), d2.data()), "nearestKSearchBatch");
    idx.assign(tmp.begin(), tmp.end());
  }
  const detail::CloudHandle& handle() const { ensure(); return handle_; }

 private:
  void ensure() const {  // lazy upload (setInputCloudLazy) on the first use
    if (handle_) return;
    if (!cloud_) throw std::runtime_error("semanticicp (B200): DeviceKdTree used before setInputCloud");
    handle_ = detail::upload_whole(*cloud_);
  }
  PointCloudPtr cloud_;
  mutable detail::CloudHandle handle_;
};

}  // namespace semanticicp
#endif  // SICP_FACADE_DEVICE_KDTREE_H_
