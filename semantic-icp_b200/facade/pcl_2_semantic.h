// pcl_2_semantic.h — semanticicp::pcl_2_semantic with the signature of the reference's semantic_icp/pcl_2_semantic.h:14-16
// (inline here; the reference defines a non-inline function in a header — an ODR hazard).
// Semantics kept: classes appear in first-appearance order of their label and points keep their original order inside a
// class (pcl_2_semantic.h:24-35).  Two passes over the labels (count, then fill pre-sized per-class clouds); the
// per-class search trees and covariances are then built on the device for all classes at once.
#ifndef SICP_FACADE_PCL_2_SEMANTIC_H_
#define SICP_FACADE_PCL_2_SEMANTIC_H_
#include "semantic_point_cloud.h"

namespace semanticicp {

inline void pcl_2_semantic(const pcl::PointCloud<pcl::PointXYZL>::Ptr pclCloud,
                           std::shared_ptr<SemanticPointCloud<pcl::PointXYZ, uint32_t>> semanticCloud) {
  typedef pcl::PointCloud<pcl::PointXYZ> ClassCloud;
  const std::size_t n = pclCloud->points.size();
  // pass 1: class index of every point + class sizes, classes numbered by first appearance
  std::vector<uint32_t> order;                 // label of class c
  std::vector<std::size_t> sizes;              // points in class c
  std::vector<uint32_t> cls(n);
  std::map<uint32_t, uint32_t> index_of;
  for (std::size_t i = 0; i < n; i++) {
    const uint32_t l = pclCloud->points[i].label;
    std::map<uint32_t, uint32_t>::iterator it = index_of.find(l);
    if (it == index_of.end()) {
      it = index_of.insert(std::make_pair(l, (uint32_t)order.size())).first;
      order.push_back(l);
      sizes.push_back(0);
    }
    cls[i] = it->second;
    sizes[it->second]++;
  }
  // pass 2: fill the per-class clouds in the original point order
  std::vector<ClassCloud::Ptr> parts(order.size());
  std::vector<std::size_t> cursor(order.size(), 0);
  for (std::size_t c = 0; c < order.size(); c++) { parts[c] = ClassCloud::Ptr(new ClassCloud()); parts[c]->resize(sizes[c]); }
  for (std::size_t i = 0; i < n; i++) {
    const pcl::PointXYZL& p = pclCloud->points[i];
    parts[cls[i]]->points[cursor[cls[i]]++] = pcl::PointXYZ(p.x, p.y, p.z);
  }
  semanticCloud->addSemanticClouds(order, parts);
}

}  // namespace semanticicp
#endif  // SICP_FACADE_PCL_2_SEMANTIC_H_
