// pcl_2_semantic.h — semanticicp::pcl_2_semantic, same signature as the reference's semantic_icp/pcl_2_semantic.h:14-16
// (marked inline here: the reference defines a non-inline function in a header, an ODR hazard — SURVEY Appendix D).
// Labels keep first-appearance order and points keep their original order inside a class (pcl_2_semantic.h:24-35);
// the per-class kd-trees and covariances are then built on the device in one pass over all classes.
#ifndef SICP_FACADE_PCL_2_SEMANTIC_H_
#define SICP_FACADE_PCL_2_SEMANTIC_H_
#include "semantic_point_cloud.h"

namespace semanticicp {

inline void pcl_2_semantic(const pcl::PointCloud<pcl::PointXYZL>::Ptr pclCloud,
                           std::shared_ptr<SemanticPointCloud<pcl::PointXYZ, uint32_t>> semanticCloud) {
  typedef pcl::PointCloud<pcl::PointXYZ> PointCloud;
  typedef PointCloud::Ptr PointCloudPtr;
  std::vector<uint32_t> labels;
  std::vector<PointCloudPtr> clouds;
  std::map<uint32_t, std::size_t> slot;
  for (const pcl::PointXYZL& p : pclCloud->points) {
    auto it = slot.find(p.label);
    if (it == slot.end()) {
      it = slot.emplace(p.label, clouds.size()).first;
      labels.push_back(p.label);
      clouds.push_back(PointCloudPtr(new PointCloud()));
    }
    clouds[it->second]->push_back(pcl::PointXYZ(p.x, p.y, p.z));
  }
  semanticCloud->addSemanticClouds(labels, clouds);
}

}  // namespace semanticicp
#endif  // SICP_FACADE_PCL_2_SEMANTIC_H_
