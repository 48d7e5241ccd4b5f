// sicp_compat.h — picks the real Eigen / Sophus / PCL headers when they are installed and the minimal stand-ins in
// compat/ otherwise, and holds the helpers the facade classes share (C-ABI status -> exception, pose conversion,
// device handles).  The facade never computes registration results on the host: every number comes from
// libsicp_b200.so (include/sicp_b200.h); if the library reports an error the facade throws std::runtime_error.
#ifndef SICP_FACADE_COMPAT_H_
#define SICP_FACADE_COMPAT_H_

#if defined(__has_include)
#if __has_include(<Eigen/Core>) && !defined(SICP_FACADE_FORCE_SHIMS)
#define SICP_HAVE_EIGEN 1
#endif
#if __has_include(<sophus/se3.hpp>) && !defined(SICP_FACADE_FORCE_SHIMS)
#define SICP_HAVE_SOPHUS 1
#endif
#if __has_include(<pcl/point_types.h>) && !defined(SICP_FACADE_FORCE_SHIMS)
#define SICP_HAVE_PCL 1
#endif
#endif

#ifdef SICP_HAVE_EIGEN
#include <Eigen/Core>
#include <Eigen/Geometry>
#include <Eigen/StdVector>
#else
#include "compat/eigen_min.h"
#endif
#if defined(SICP_HAVE_SOPHUS) && defined(SICP_HAVE_EIGEN)
#include <sophus/se3.hpp>
#else
#include "compat/sophus_min.h"
#endif
#if defined(SICP_HAVE_PCL) && defined(SICP_HAVE_EIGEN)
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#include <pcl/common/transforms.h>
#if __has_include(<pcl/io/pcd_io.h>)
#include <pcl/io/pcd_io.h>
#endif
#else
#include "compat/pcl_min.h"
#include "compat/pcl_io_min.h"
#endif

#include <cstddef>
#include <cstdint>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "sicp_b200.h"

namespace semanticicp {
namespace detail {

inline void check(sicp_status st, const char* what) {
  if (st != SICP_OK) throw std::runtime_error(std::string("semanticicp (B200): ") + what + ": " + sicp_last_error());
}

// Sophus::SE3d <-> the ABI's 7 doubles [qx,qy,qz,qw,tx,ty,tz] (Sophus::SE3d::data() order)
inline void se3_to_pose7(const Sophus::SE3d& T, double* p7) { std::memcpy(p7, T.data(), 7 * sizeof(double)); }
inline Sophus::SE3d pose7_to_se3(const double* p7) {
  Sophus::SE3d T;
  std::memcpy(T.data(), p7, 7 * sizeof(double));
  return T;
}

typedef std::shared_ptr<sicp_cloud> CloudHandle;
inline CloudHandle make_handle(sicp_cloud* c) { return CloudHandle(c, [](sicp_cloud* p) { sicp_cloud_destroy(p); }); }

// byte offset of the label inside a labelled point type (pcl::PointXYZL: 16)
template <typename PointT>
struct LabelOffset {
  static const void* get(const PointT*) { return nullptr; }
};
template <>
struct LabelOffset<pcl::PointXYZL> {
  static const void* get(const pcl::PointXYZL* p) { return p ? &p->label : nullptr; }
};

// GPU the facade objects of this process put their clouds on (one process per GPU: LOCAL_RANK, see shard.py).  The
// reference has no notion of a device; semanticicp::setDevice() is the one extension a multi-GPU host needs.
inline int& device_index() {
  static int dev = 0;
  return dev;
}

// Upload a PCL cloud (WHOLE layout): xyz read in place with the point stride, label (if the type has one) likewise.
template <typename PointT>
inline CloudHandle upload_whole(const pcl::PointCloud<PointT>& cloud, int device = -1) {
  if (device < 0) device = device_index();
  sicp_cloud* c = nullptr;
  const PointT* p0 = cloud.points.empty() ? nullptr : cloud.points.data();
  static PointT dummy;
  const void* xyz = p0 ? (const void*)&p0->x : (const void*)&dummy.x;
  const void* lab = LabelOffset<PointT>::get(p0 ? p0 : &dummy);
  check(sicp_cloud_create(xyz, sizeof(PointT), lab, sizeof(PointT), cloud.points.size(), SICP_CLOUD_WHOLE, device, &c), "cloud upload");
  return make_handle(c);
}

typedef std::vector<Eigen::Matrix3d, Eigen::aligned_allocator<Eigen::Matrix3d>> MatricesVector;

// n x 9 row-major doubles -> vector<Matrix3d> (the matrices are exactly symmetric)
inline void fill_matrices(const std::vector<double>& rows, std::size_t n, std::size_t first, MatricesVector* out) {
  out->resize(n);
  for (std::size_t i = 0; i < n; i++)
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++) (*out)[i](r, c) = rows[(first + i) * 9 + 3 * r + c];
}

}  // namespace detail

// Select the GPU for every facade object created afterwards (default 0).
inline void setDevice(int device) { detail::device_index() = device; }
inline int getDevice() { return detail::device_index(); }

}  // namespace semanticicp
#endif  // SICP_FACADE_COMPAT_H_
