// em_icp.h — semanticicp::EmIterativeClosestPoint<N>, source-compatible with the reference's semantic_icp/em_icp.h:16-122
// and implemented over the C ABI of libsicp_b200.  N stays a template parameter here (the reference's signature);
// the ABI takes it at run time (sicp_options.n_classes).  Labels must be 1..N (impl/em_icp.hpp:301 indexes label-1):
// the library rejects label 0 / label > N with an error instead of the reference's out-of-bounds write.
#ifndef SICP_FACADE_EM_ICP_H_
#define SICP_FACADE_EM_ICP_H_
#include "device_kdtree.h"

namespace semanticicp {

template <size_t N>
class EmIterativeClosestPoint {
 public:
  typedef pcl::PointXYZL PointT;
  typedef typename pcl::PointCloud<PointT> PointCloud;
  typedef typename PointCloud::Ptr PointCloudPtr;
  typedef detail::MatricesVector MatricesVector;
  typedef std::vector<Eigen::Matrix<double, N, 1>, Eigen::aligned_allocator<Eigen::Matrix<double, N, 1>>> DistVector;
  typedef std::shared_ptr<MatricesVector> MatricesVectorPtr;
  typedef std::shared_ptr<const MatricesVector> MatricesVectorConstPtr;
  typedef std::shared_ptr<DistVector> DistVectorPtr;
  typedef DeviceKdTree<PointT> KdTree;
  typedef typename KdTree::Ptr KdTreePtr;
  typedef Eigen::Matrix<double, 6, 1> Vector6d;

  EmIterativeClosestPoint(int k = 20, double epsilon = 0.001) : kCorrespondences_(k), kEpsilon_(epsilon), outer_iter(0) {  // em_icp.h:42
    for (size_t i = 0; i < N * N; i++) cm_[i] = 0.0;
  }
  inline void setSourceCloud(const PointCloudPtr& cloud) {  // em_icp.h:50-57: tree built now, xyz + labels uploaded
    source_cloud_ = cloud;
    source_kd_tree_ = KdTreePtr(new KdTree());
    source_kd_tree_->setInputCloud(source_cloud_);
  }
  inline void setTargetCloud(const PointCloudPtr& cloud) {  // em_icp.h:59-66
    target_cloud_ = cloud;
    target_kd_tree_ = KdTreePtr(new KdTree());
    target_kd_tree_->setInputCloud(target_cloud_);
  }
  inline void setConfusionMatrix(const Eigen::Matrix<double, N, N>& in) {  // em_icp.h:68
    for (size_t b = 0; b < N; b++)
      for (size_t s = 0; s < N; s++) cm_[b * N + s] = in((int)b, (int)s);  // ABI: row-major, CM(b, s) = in.col(s)[b]
  }
  // declared but never defined in the reference (em_icp.h:73-74): provided here with an identity start
  void align(PointCloudPtr finalCloud) {
    Sophus::SE3d init;
    align(finalCloud, init);
  }
  void align(PointCloudPtr finalCloud, const Sophus::SE3d& initTransform) {  // impl/em_icp.hpp:24-200
    require_clouds("align");
    sicp_options opts = options();
    double init7[7];
    detail::se3_to_pose7(initTransform, init7);
    sicp_result res;
    detail::check(sicp_register(SICP_ALGO_EM, source_kd_tree_->handle().get(), target_kd_tree_->handle().get(), &opts, init7, &res), "EmIterativeClosestPoint::align");
    final_transformation_ = detail::pose7_to_se3(res.pose7);
    outer_iter = res.outer_iter;
    if (finalCloud != nullptr) {  // impl/em_icp.hpp:192-198
      // the float-matrix transform runs on the device (sicp_cloud_transform_f32); every other field of the points
      // (labels, padding) is carried over from the source, like pcl::transformPointCloud does
      if (finalCloud.get() != source_cloud_.get()) *finalCloud = *source_cloud_;
      if (!finalCloud->points.empty())
        detail::check(sicp_cloud_transform_f32(source_kd_tree_->handle().get(), res.pose7, &finalCloud->points[0].x, sizeof(PointT)), "final cloud transform");
    }
  }
  // impl/em_icp.hpp:202-268: appends the source points, relabelled with the arg-max fused class, to labeledCloud
  void getFusedLabels(PointCloudPtr labeledCloud, const Sophus::SE3d& transformation) {
    require_clouds("getFusedLabels");
    sicp_options opts = options();
    double pose7[7];
    detail::se3_to_pose7(transformation, pose7);
    std::vector<std::uint32_t> labels(source_cloud_->points.size());
    detail::check(sicp_fused_labels(source_kd_tree_->handle().get(), target_kd_tree_->handle().get(), &opts, pose7, labels.data()), "getFusedLabels");
    for (std::size_t i = 0; i < labels.size(); i++) {
      PointT p = source_cloud_->points[i];
      p.label = labels[i];
      labeledCloud->push_back(p);
    }
  }
  Sophus::SE3d getFinalTransFormation() { return final_transformation_; }  // em_icp.h:82 (sic)
  int getOuterIter() { return outer_iter; }                                // em_icp.h:88

 protected:
  int kCorrespondences_;
  double kEpsilon_;
  int outer_iter;
  double cm_[N * N];
  Sophus::SE3d final_transformation_;
  PointCloudPtr source_cloud_, target_cloud_;
  KdTreePtr source_kd_tree_, target_kd_tree_;

  sicp_options options() const {
    sicp_options o;
    sicp_options_default(SICP_ALGO_EM, &o);
    o.k_cov = kCorrespondences_;
    o.epsilon = kEpsilon_;
    o.n_classes = (int)N;
    o.confusion = cm_;
    return o;
  }
  void require_clouds(const char* what) const {
    if (!source_kd_tree_ || !target_kd_tree_)
      throw std::runtime_error(std::string("semanticicp (B200): EmIterativeClosestPoint::") + what + " before setSourceCloud/setTargetCloud");
  }
};

}  // namespace semanticicp
#endif  // SICP_FACADE_EM_ICP_H_
