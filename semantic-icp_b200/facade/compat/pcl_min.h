// pcl_min.h — layout-compatible stand-ins for the PCL types in the reference's public API, for builds where PCL is
// not installed: pcl::PointXYZ (16 B), pcl::PointXYZL (32 B, label at byte 16), pcl::PointCloud<T> with
// points/width/height/is_dense and the std-container conveniences the reference's drivers use, and the float / double
// pcl::transformPointCloud overloads.  Used only when <pcl/point_types.h> is absent (see ../sicp_compat.h).
#ifndef SICP_FACADE_PCL_MIN_H_
#define SICP_FACADE_PCL_MIN_H_
#include <cstdint>
#include <memory>
#include <vector>

namespace pcl {

struct alignas(16) PointXYZ {
  float x, y, z, pad_;
  PointXYZ() : x(0.f), y(0.f), z(0.f), pad_(1.f) {}
  PointXYZ(float x_, float y_, float z_) : x(x_), y(y_), z(z_), pad_(1.f) {}
};
struct alignas(16) PointXYZL {
  float x, y, z, pad_;
  std::uint32_t label;
  std::uint32_t pad2_[3];
  PointXYZL() : x(0.f), y(0.f), z(0.f), pad_(1.f), label(0) { pad2_[0] = pad2_[1] = pad2_[2] = 0; }
};
static_assert(sizeof(PointXYZ) == 16 && sizeof(PointXYZL) == 32, "PCL point layouts");

template <typename PointT>
class PointCloud {
 public:
  typedef std::shared_ptr<PointCloud<PointT>> Ptr;
  typedef std::shared_ptr<const PointCloud<PointT>> ConstPtr;
  typedef typename std::vector<PointT>::iterator iterator;
  typedef typename std::vector<PointT>::const_iterator const_iterator;
  std::vector<PointT> points;
  std::uint32_t width = 0, height = 0;
  bool is_dense = true;
  void push_back(const PointT& p) { points.push_back(p); width = (std::uint32_t)points.size(); height = 1; }
  std::size_t size() const { return points.size(); }
  bool empty() const { return points.empty(); }
  void clear() { points.clear(); width = height = 0; }
  void resize(std::size_t n) { points.resize(n); width = (std::uint32_t)n; height = 1; }
  PointT& operator[](std::size_t i) { return points[i]; }
  const PointT& operator[](std::size_t i) const { return points[i]; }
  PointT& at(std::size_t i) { return points.at(i); }
  const PointT& at(std::size_t i) const { return points.at(i); }
  iterator begin() { return points.begin(); }
  iterator end() { return points.end(); }
  const_iterator begin() const { return points.begin(); }
  const_iterator end() const { return points.end(); }
};

// out = M * in per point, arithmetic in Scalar (float for Matrix4f, double for Matrix4d), stored as float;
// every non-coordinate field (label) is carried over.  in and out may alias.
template <typename PointT, typename Scalar>
void transformPointCloud(const PointCloud<PointT>& in, PointCloud<PointT>& out, const Eigen::Matrix<Scalar, 4, 4>& m) {
  if (&in != &out) { out.points = in.points; out.width = in.width; out.height = in.height; out.is_dense = in.is_dense; }
  for (PointT& p : out.points) {
    const Scalar x = p.x, y = p.y, z = p.z;
    p.x = static_cast<float>(m(0, 0) * x + m(0, 1) * y + m(0, 2) * z + m(0, 3));
    p.y = static_cast<float>(m(1, 0) * x + m(1, 1) * y + m(1, 2) * z + m(1, 3));
    p.z = static_cast<float>(m(2, 0) * x + m(2, 1) * y + m(2, 2) * z + m(2, 3));
  }
}

}  // namespace pcl
#endif  // SICP_FACADE_PCL_MIN_H_
