// semantic_icp.h — semanticicp::SemanticIterativeClosestPoint<PointT,SemanticT>, source-compatible with the
// reference's semantic_icp/semantic_icp.h:14-83 (setInputSource / setInputTarget / align / getFinalTransFormation)
// and implemented over the C ABI of libsicp_b200: per-class 1-NN in the same-label target tree, classes used only
// when present in the target and the source class has > 400 points, CauchyLoss(1.5), stop at mse < 1e-3 or 36 passes
// (impl/semantic_icp.hpp:27-166).
#ifndef SICP_FACADE_SEMANTIC_ICP_H_
#define SICP_FACADE_SEMANTIC_ICP_H_
#include <iostream>
#include "semantic_point_cloud.h"

namespace semanticicp {

template <typename PointT, typename SemanticT>
class SemanticIterativeClosestPoint {
 public:
  typedef SemanticPointCloud<PointT, SemanticT> SemanticCloud;
  typedef typename std::shared_ptr<SemanticCloud> SemanticCloudPtr;
  typedef typename std::shared_ptr<const SemanticCloud> SemanticCloudConstPtr;
  typedef DeviceKdTree<PointT> KdTree;
  typedef typename KdTree::Ptr KdTreePtr;
  typedef Eigen::Matrix<double, 6, 1> Vector6d;

  SemanticIterativeClosestPoint() : outer_iter(0) {}
  inline void setInputSource(const SemanticCloudPtr& cloud) { sourceCloud_ = cloud; }  // semantic_icp.h:41-44
  inline void setInputTarget(const SemanticCloudPtr& cloud) { targetCloud_ = cloud; }  // semantic_icp.h:46-49

  void align(SemanticCloudPtr finalCloud) {  // impl/semantic_icp.hpp:20-25
    Sophus::SE3d init;
    align(finalCloud, init);
  }
  void align(SemanticCloudPtr finalCloud, Sophus::SE3d& initTransform) {  // impl/semantic_icp.hpp:27-166
    if (!sourceCloud_ || !targetCloud_) throw std::runtime_error("semanticicp (B200): SemanticIterativeClosestPoint::align before setInputSource/setInputTarget");
    if (!finalCloud) throw std::runtime_error("semanticicp (B200): SemanticIterativeClosestPoint::align needs a non-null finalCloud (the reference dereferences it, impl/semantic_icp.hpp:165)");
    sicp_options opts;
    sicp_options_default(SICP_ALGO_SEMANTIC, &opts);
    double init7[7];
    detail::se3_to_pose7(initTransform, init7);
    sicp_result res;
    detail::CloudHandle src = sourceCloud_->device(), tgt = targetCloud_->device();
    detail::check(sicp_register(SICP_ALGO_SEMANTIC, src.get(), tgt.get(), &opts, init7, &res), "SemanticIterativeClosestPoint::align");
    finalTransformation_ = detail::pose7_to_se3(res.pose7);
    outer_iter = res.outer_iter;
    Eigen::Matrix4f mat = finalTransformation_.matrix().template cast<float>();
    finalCloud->transform(mat);  // in place, like the reference (hpp:161-165)
  }
  Sophus::SE3d getFinalTransFormation() { return finalTransformation_; }  // semantic_icp.h:58 (sic)
  int getOuterIter() { return outer_iter; }                               // extension (GICP / EM have it)

  // semantic_icp.h:78-81 (protected there and never called by align(): dead code in the reference; public here so that a
  // caller who wants per-class pose fusion can reach them).  Both run on the device (sicp_iterative_mean, sicp_pose_fusion).
  typedef std::vector<Eigen::Matrix<double, 6, 6>, Eigen::aligned_allocator<Eigen::Matrix<double, 6, 6>>> CovarianceVector;
  Sophus::SE3d iterativeMean(std::vector<Sophus::SE3d> const& in, size_t maxIterations) {  // impl/semantic_icp.hpp:169-191
    std::vector<double> p(7 * in.size());
    for (std::size_t i = 0; i < in.size(); i++) detail::se3_to_pose7(in[i], &p[7 * i]);
    double out[7];
    int converged = 0;
    detail::check(sicp_iterative_mean(in.size(), p.data(), (int)maxIterations, out, &converged), "iterativeMean");
    if (!converged) std::cout << "Iterative Mean Failed";  // impl/semantic_icp.hpp:189
    return detail::pose7_to_se3(out);
  }
  Sophus::SE3d poseFusion(std::vector<Sophus::SE3d> const& poses, CovarianceVector const& covs, Sophus::SE3d const& initTransform) {  // hpp:213-265
    std::vector<double> p(7 * poses.size()), c(36 * covs.size());
    for (std::size_t i = 0; i < poses.size(); i++) detail::se3_to_pose7(poses[i], &p[7 * i]);
    for (std::size_t i = 0; i < covs.size(); i++)
      for (int r = 0; r < 6; r++)
        for (int q = 0; q < 6; q++) c[36 * i + 6 * r + q] = covs[i](r, q);
    double init7[7], out[7];
    detail::se3_to_pose7(initTransform, init7);
    int iters = 0;
    if (covs.size() != poses.size()) throw std::runtime_error("semanticicp (B200): poseFusion needs one covariance per pose");
    detail::check(sicp_pose_fusion(poses.size(), p.data(), c.data(), init7, out, &iters), "poseFusion");
    return detail::pose7_to_se3(out);
  }

 protected:
  int outer_iter;
  Sophus::SE3d finalTransformation_;
  SemanticCloudPtr sourceCloud_, targetCloud_;
};

}  // namespace semanticicp
#endif  // SICP_FACADE_SEMANTIC_ICP_H_
