// pcl_io_min.h — pcl::io::loadPCDFile / savePCDFileASCII / savePCDFileBinary for the stand-in point types, so that the
// reference's drivers (exec/test_icp.cc:35-43, exec/kitti_eval.cc:132-159 load PCD scans) run without PCL installed.
// Supports PCD v0.7 headers with any field list (x y z are required, `label` is read when the point type has one,
// other fields are skipped), SIZE 1/2/4/8, TYPE F/U/I, COUNT n, DATA ascii | binary (binary_compressed is rejected).
// Used only when <pcl/io/pcd_io.h> is absent.  Return convention as in PCL: 0 on success, -1 on failure.
#ifndef SICP_FACADE_PCL_IO_MIN_H_
#define SICP_FACADE_PCL_IO_MIN_H_
#include <cstdint>
#include <cstring>
#include <exception>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

namespace pcl {
namespace io {
namespace detail {

struct Field { std::string name; int size = 4; char type = 'F'; int count = 1; int offset = 0; };

template <typename PointT> struct HasLabel { enum { value = 0 }; static void set(PointT&, std::uint32_t) {} static std::uint32_t get(const PointT&) { return 0; } };
template <> struct HasLabel<pcl::PointXYZL> {
  enum { value = 1 };
  static void set(pcl::PointXYZL& p, std::uint32_t l) { p.label = l; }
  static std::uint32_t get(const pcl::PointXYZL& p) { return p.label; }
};

inline double read_scalar(const char* p, const Field& f) {
  switch (f.type) {
    case 'F': if (f.size == 4) { float v; std::memcpy(&v, p, 4); return v; } else { double v; std::memcpy(&v, p, 8); return v; }
    case 'U': switch (f.size) { case 1: { std::uint8_t v; std::memcpy(&v, p, 1); return v; } case 2: { std::uint16_t v; std::memcpy(&v, p, 2); return v; }
                                case 4: { std::uint32_t v; std::memcpy(&v, p, 4); return v; } default: { std::uint64_t v; std::memcpy(&v, p, 8); return (double)v; } }
    default:  switch (f.size) { case 1: { std::int8_t v; std::memcpy(&v, p, 1); return v; } case 2: { std::int16_t v; std::memcpy(&v, p, 2); return v; }
                                case 4: { std::int32_t v; std::memcpy(&v, p, 4); return v; } default: { std::int64_t v; std::memcpy(&v, p, 8); return (double)v; } }
  }
}

}  // namespace detail

template <typename PointT>
int loadPCDFile(const std::string& file_name, pcl::PointCloud<PointT>& cloud) {
  std::ifstream in(file_name.c_str(), std::ios::binary);
  if (!in) return -1;
  std::vector<detail::Field> fields;
  std::size_t points = 0, width = 0, height = 1;
  std::string data_mode, line;
  while (std::getline(in, line)) {
    if (!line.empty() && line[line.size() - 1] == '\r') line.erase(line.size() - 1);
    if (line.empty() || line[0] == '#') continue;
    std::istringstream ls(line);
    std::string key;
    ls >> key;
    if (key == "FIELDS") { std::string n; while (ls >> n) { detail::Field f; f.name = n; fields.push_back(f); } }
    else if (key == "SIZE") { for (std::size_t i = 0; i < fields.size() && (ls >> fields[i].size); i++) {} }
    else if (key == "TYPE") { for (std::size_t i = 0; i < fields.size() && (ls >> fields[i].type); i++) {} }
    else if (key == "COUNT") { for (std::size_t i = 0; i < fields.size() && (ls >> fields[i].count); i++) {} }
    else if (key == "WIDTH") ls >> width;
    else if (key == "HEIGHT") ls >> height;
    else if (key == "POINTS") ls >> points;
    else if (key == "DATA") { ls >> data_mode; break; }
  }
  if (fields.empty() || data_mode.empty()) return -1;
  if (points == 0) points = width * height;
  // The header is untrusted input: only the SIZE / TYPE / COUNT combinations of the format are accepted, and the point
  // count must fit the bytes that are actually left in the file (a record is at least one byte, in ascii two per value).
  const std::streampos data_pos = in.tellg();
  in.seekg(0, std::ios::end);
  const std::streamoff remaining = in.tellg() - data_pos;
  in.seekg(data_pos);
  if (remaining < 0) return -1;
  int ix = -1, iy = -1, iz = -1, il = -1;
  long long stride = 0;
  for (std::size_t i = 0; i < fields.size(); i++) {
    const detail::Field& f = fields[i];
    const bool size_ok = f.type == 'F' ? (f.size == 4 || f.size == 8) : ((f.type == 'U' || f.type == 'I') && (f.size == 1 || f.size == 2 || f.size == 4 || f.size == 8));
    if (!size_ok || f.count < 1 || f.count > (1 << 20)) return -1;
    fields[i].offset = (int)stride;
    stride += (long long)f.size * f.count;
    if (stride > (1 << 24)) return -1;
    if (fields[i].name == "x") ix = (int)i; else if (fields[i].name == "y") iy = (int)i; else if (fields[i].name == "z") iz = (int)i;
    else if (fields[i].name == "label") il = (int)i;
  }
  if (ix < 0 || iy < 0 || iz < 0 || stride <= 0) return -1;
  if (data_mode == "binary" ? (unsigned long long)points > (unsigned long long)remaining / (unsigned long long)stride
                            : (unsigned long long)points > (unsigned long long)remaining) return -1;
  cloud.points.clear();
  try {
    cloud.points.reserve(points);
  } catch (const std::exception&) {
    return -1;
  }
  if (data_mode == "ascii") {
    for (std::size_t n = 0; n < points; n++) {
      PointT p;
      for (std::size_t i = 0; i < fields.size(); i++)
        for (int c = 0; c < fields[i].count; c++) {
          double v;
          if (!(in >> v)) return -1;
          if (c) continue;
          if ((int)i == ix) p.x = (float)v; else if ((int)i == iy) p.y = (float)v; else if ((int)i == iz) p.z = (float)v;
          else if ((int)i == il) detail::HasLabel<PointT>::set(p, (std::uint32_t)v);
        }
      cloud.points.push_back(p);
    }
  } else if (data_mode == "binary") {
    std::vector<char> buf;
    try {
      buf.resize((std::size_t)stride * points);
    } catch (const std::exception&) {
      return -1;
    }
    in.read(buf.data(), (std::streamsize)buf.size());
    if ((std::size_t)in.gcount() != buf.size()) return -1;
    for (std::size_t n = 0; n < points; n++) {
      const char* r = buf.data() + n * (std::size_t)stride;
      PointT p;
      p.x = (float)detail::read_scalar(r + fields[ix].offset, fields[ix]);
      p.y = (float)detail::read_scalar(r + fields[iy].offset, fields[iy]);
      p.z = (float)detail::read_scalar(r + fields[iz].offset, fields[iz]);
      if (il >= 0) detail::HasLabel<PointT>::set(p, (std::uint32_t)detail::read_scalar(r + fields[il].offset, fields[il]));
      cloud.points.push_back(p);
    }
  } else {
    return -1;  // binary_compressed (LZF) is not supported by the stand-in
  }
  cloud.width = (std::uint32_t)(height > 1 ? width : cloud.points.size());
  cloud.height = (std::uint32_t)(height > 1 ? height : 1);
  cloud.is_dense = true;
  return 0;
}

template <typename PointT>
int savePCDFile(const std::string& file_name, const pcl::PointCloud<PointT>& cloud, bool binary_mode = false) {
  std::ofstream out(file_name.c_str(), std::ios::binary);
  if (!out) return -1;
  const bool lab = detail::HasLabel<PointT>::value != 0;
  const std::size_t n = cloud.points.size();
  out << "# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\n";
  out << (lab ? "FIELDS x y z label\nSIZE 4 4 4 4\nTYPE F F F U\nCOUNT 1 1 1 1\n" : "FIELDS x y z\nSIZE 4 4 4\nTYPE F F F\nCOUNT 1 1 1\n");
  out << "WIDTH " << n << "\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS " << n << "\nDATA " << (binary_mode ? "binary" : "ascii") << "\n";
  if (binary_mode) {
    for (std::size_t i = 0; i < n; i++) {
      const PointT& p = cloud.points[i];
      const float xyz[3] = {p.x, p.y, p.z};
      out.write((const char*)xyz, 12);
      if (lab) { const std::uint32_t l = detail::HasLabel<PointT>::get(p); out.write((const char*)&l, 4); }
    }
  } else {
    out.precision(9);  // round-trips every float
    for (std::size_t i = 0; i < n; i++) {
      const PointT& p = cloud.points[i];
      out << p.x << " " << p.y << " " << p.z;
      if (lab) out << " " << detail::HasLabel<PointT>::get(p);
      out << "\n";
    }
  }
  return out ? 0 : -1;
}
template <typename PointT>
int savePCDFileASCII(const std::string& file_name, const pcl::PointCloud<PointT>& cloud) { return savePCDFile(file_name, cloud, false); }
template <typename PointT>
int savePCDFileBinary(const std::string& file_name, const pcl::PointCloud<PointT>& cloud) { return savePCDFile(file_name, cloud, true); }

}  // namespace io
}  // namespace pcl
#endif  // SICP_FACADE_PCL_IO_MIN_H_
