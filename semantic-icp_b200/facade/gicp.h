// gicp.h — semanticicp::GICP<PointT>, source-compatible with the reference's semantic_icp/gicp.h:14-132 and
// implemented over the C ABI of libsicp_b200 (include/sicp_b200.h).  Same names, argument meaning and call order;
// align() is synchronous and void, results are fetched with the getters.  Differences:
//   * KdTree is semanticicp::DeviceKdTree<PointT> (device Morton/box tree) instead of pcl::KdTreeFLANN<PointT>;
//   * a failure of the CUDA library throws std::runtime_error (the reference has no error path);
//   * nothing is printed (the reference prints the Ceres report every pass, impl/gicp.hpp:65,151,156-159).
#ifndef SICP_FACADE_GICP_H_
#define SICP_FACADE_GICP_H_
#include "device_kdtree.h"

namespace semanticicp {

template <typename PointT>
class GICP {
 public:
  typedef pcl::PointCloud<PointT> PointCloud;
  typedef typename PointCloud::Ptr PointCloudPtr;
  typedef detail::MatricesVector MatricesVector;
  typedef std::shared_ptr<MatricesVector> MatricesVectorPtr;
  typedef std::shared_ptr<const MatricesVector> MatricesVectorConstPtr;
  typedef DeviceKdTree<PointT> KdTree;
  typedef typename KdTree::Ptr KdTreePtr;
  typedef Eigen::Matrix<double, 6, 1> Vector6d;

  GICP(int k = 20, double epsilon = 0.001) : kCorrespondences_(k), epsilon_(epsilon), outer_iter(0) {}  // gicp.h:34

  // gicp.h:42-48 — stores the shared pointer (no copy of the caller's cloud), builds the search tree immediately
  inline void setSourceCloud(const PointCloudPtr& cloud) {
    sourceCloud_ = cloud;
    sourceKdTree_ = KdTreePtr(new KdTree());
    sourceKdTree_->setInputCloud(sourceCloud_);
    sourceCovariances_ = MatricesVectorPtr(new MatricesVector());
    sourceCovStale_ = false;
  }
  // gicp.h:50-55 — injected covariances are overwritten by align() exactly like the reference (impl/gicp.hpp:33-34)
  inline void setSourceCloud(const PointCloudPtr& cloud, const KdTreePtr& tree, const MatricesVectorPtr& covs) {
    sourceCloud_ = cloud; sourceKdTree_ = tree; sourceCovariances_ = covs; sourceCovStale_ = false;
  }
  inline void setTargetCloud(const PointCloudPtr& cloud) {  // gicp.h:57-63
    targetCloud_ = cloud;
    targetKdTree_ = KdTreePtr(new KdTree());
    targetKdTree_->setInputCloud(targetCloud_);
    targetCovariances_ = MatricesVectorPtr(new MatricesVector());
    targetCovStale_ = false;
  }
  inline void setTargetCloud(const PointCloudPtr& cloud, const KdTreePtr& tree, const MatricesVectorPtr& covs) {  // gicp.h:65-70
    targetCloud_ = cloud; targetKdTree_ = tree; targetCovariances_ = covs; targetCovStale_ = false;
  }
  inline KdTreePtr getSourceKdTree() { return sourceKdTree_; }
  inline KdTreePtr getTargetKdTree() { return targetKdTree_; }
  // The reference fills these vectors inside align() (computeCovariances, impl/gicp.hpp:177-239).  Here the
  // covariances stay on the device and are downloaded the first time a getter asks for them after an align().
  inline MatricesVectorPtr getSourceCovariances() { refresh(sourceKdTree_, sourceCovariances_, &sourceCovStale_); return sourceCovariances_; }
  inline MatricesVectorPtr getTargetCovariances() { refresh(targetKdTree_, targetCovariances_, &targetCovStale_); return targetCovariances_; }

  void align(PointCloudPtr finalCloud) {  // impl/gicp.hpp:22-27: identity start
    Sophus::SE3d init;
    align(finalCloud, init);
  }
  void align(PointCloudPtr finalCloud, Sophus::SE3d& initTransform) {  // impl/gicp.hpp:29-175
    if (!sourceKdTree_ || !targetKdTree_) throw std::runtime_error("semanticicp (B200): GICP::align before setSourceCloud/setTargetCloud");
    sicp_options opts;
    sicp_options_default(SICP_ALGO_GICP, &opts);
    opts.k_cov = kCorrespondences_;
    opts.epsilon = epsilon_;
    double init7[7];
    detail::se3_to_pose7(initTransform, init7);
    sicp_result res;
    detail::check(sicp_register(SICP_ALGO_GICP, sourceKdTree_->handle().get(), targetKdTree_->handle().get(), &opts, init7, &res), "GICP::align");
    finalTransformation_ = detail::pose7_to_se3(res.pose7);
    outer_iter = res.outer_iter;
    sourceCovStale_ = targetCovStale_ = true;
    if (finalCloud != nullptr) {  // impl/gicp.hpp:166-172: Matrix4f (float) transform of the source
      // the float-matrix transform runs on the device (sicp_cloud_transform_f32); every other field of the points
      // (labels, padding) is carried over from the source, like pcl::transformPointCloud does
      if (finalCloud.get() != sourceCloud_.get()) *finalCloud = *sourceCloud_;
      if (!finalCloud->points.empty())
        detail::check(sicp_cloud_transform_f32(sourceKdTree_->handle().get(), res.pose7, &finalCloud->points[0].x, sizeof(PointT)), "final cloud transform");
    }
  }
  Sophus::SE3d getFinalTransFormation() { return finalTransformation_; }  // gicp.h:98 (sic)
  int getOuterIter() { return outer_iter; }                               // gicp.h:104

 protected:
  int kCorrespondences_;
  double epsilon_;
  int outer_iter;
  Sophus::SE3d finalTransformation_;
  PointCloudPtr sourceCloud_, targetCloud_;
  KdTreePtr sourceKdTree_, targetKdTree_;
  MatricesVectorPtr sourceCovariances_, targetCovariances_;
  bool sourceCovStale_ = false, targetCovStale_ = false;

  static void refresh(const KdTreePtr& tree, const MatricesVectorPtr& covs, bool* stale) {
    if (!*stale || !tree || !covs) return;
    std::size_t n = 0;
    detail::check(sicp_cloud_size(tree->handle().get(), &n), "cloud size");
    std::vector<double> rows(9 * n);
    detail::check(sicp_cloud_get_covariances(tree->handle().get(), rows.data()), "covariance download");
    detail::fill_matrices(rows, n, 0, covs.get());
    *stale = false;
  }
};

}  // namespace semanticicp
#endif  // SICP_FACADE_GICP_H_
