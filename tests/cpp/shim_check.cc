// shim_check.cc — CPU-only checks of the Eigen / Sophus / PCL stand-ins in semantic-icp_b200/facade/compat against the
// golden SE(3) vectors of tests/golden/se3.json (passed as a flat text file: delta[6] pose7[7] delta_b[6] pose7_ab[7]
// pose7_inv[7] per line) and against the layouts the reference's API relies on.  Exit code 0 = all checks pass.
#include <cmath>
#include <cstdio>
#include <fstream>
#include <sicp_compat.h>

static int fails = 0;
#define CHECK(cond, what) do { if (!(cond)) { std::fprintf(stderr, "FAIL %s (line %d)\n", what, __LINE__); fails++; } } while (0)

static bool same_pose(const Sophus::SE3d& a, const double* b, double tol) {
  const double* p = a.data();
  double dot = 0;
  for (int i = 0; i < 4; i++) dot += p[i] * b[i];
  const double s = dot < 0 ? -1.0 : 1.0;  // q and -q are the same rotation
  for (int i = 0; i < 4; i++) if (std::fabs(p[i] - s * b[i]) > tol) return false;
  for (int i = 4; i < 7; i++) if (std::fabs(p[i] - b[i]) > tol) return false;
  return true;
}

int main(int argc, char** argv) {
  // ---- layouts (pcl::PointXYZ 16 B, pcl::PointXYZL 32 B with the label at byte 16; column-major matrices)
  CHECK(sizeof(pcl::PointXYZ) == 16 && sizeof(pcl::PointXYZL) == 32, "point sizes");
  pcl::PointXYZL pl;
  CHECK((char*)&pl.label - (char*)&pl == 16, "label offset");
  Eigen::Matrix4d M = Eigen::Matrix4d::Identity();
  M(0, 3) = 5.0; M(1, 0) = 2.0;
  CHECK(M.data()[12] == 5.0 && M.data()[1] == 2.0, "column-major storage");
  Eigen::Matrix4f Mf = M.cast<float>();
  CHECK(Mf(0, 3) == 5.0f && Mf(1, 0) == 2.0f && Mf(3, 3) == 1.0f, "cast<float>");
  // ---- pcl::PointCloud conveniences + transformPointCloud (float and double overloads, in place)
  pcl::PointCloud<pcl::PointXYZL> c;
  pl.x = 1; pl.y = 2; pl.z = 3; pl.label = 7;
  c.push_back(pl);
  CHECK(c.size() == 1 && c.width == 1 && c.height == 1 && c.at(0).label == 7, "PointCloud push_back");
  Eigen::Matrix4f T = Eigen::Matrix4f::Identity();
  T(0, 3) = 10.f; T(1, 1) = 2.f;
  pcl::transformPointCloud(c, c, T);
  CHECK(c[0].x == 11.f && c[0].y == 4.f && c[0].z == 3.f && c[0].label == 7, "transformPointCloud keeps the label");
  // ---- SE(3): golden vectors
  if (argc > 1) {
    std::ifstream f(argv[1]);
    double v[33];
    int rows = 0;
    while (true) {
      for (int i = 0; i < 33; i++) if (!(f >> v[i])) goto done;
      rows++;
      Sophus::SE3d::Tangent d, db;
      for (int i = 0; i < 6; i++) { d(i) = v[i]; db(i) = v[13 + i]; }
      const Sophus::SE3d A = Sophus::SE3d::exp(d), B = Sophus::SE3d::exp(db);
      CHECK(same_pose(A, v + 6, 1e-12), "exp");
      const Sophus::SE3d::Tangent lg = A.log();
      double e = 0;
      for (int i = 0; i < 6; i++) e = std::fmax(e, std::fabs(lg(i) - v[i]));
      CHECK(e < 1e-10, "log(exp(d)) == d");
      CHECK(same_pose(A * B, v + 19, 1e-12), "composition");
      CHECK(same_pose(A.inverse(), v + 26, 1e-12), "inverse");
      CHECK(same_pose(Sophus::SE3d(A.matrix()), v + 6, 1e-12), "SE3d(Matrix4d) round trip");
      const Sophus::SE3d I = A * A.inverse();
      const double id[7] = {0, 0, 0, 1, 0, 0, 0};
      CHECK(same_pose(I, id, 1e-12), "A * A^-1 == I");
    }
  done:
    CHECK(rows >= 5, "golden rows read");
    std::printf("%d golden SE(3) rows checked\n", rows);
  }
  // ---- PCD I/O stand-in: ASCII and binary round trips (labelled and plain), and a foreign header with extra fields
  {
    const std::string dir = argc > 2 ? argv[2] : ".";
    pcl::PointCloud<pcl::PointXYZL> a, b;
    for (int i = 0; i < 50; i++) { pcl::PointXYZL p; p.x = 0.1f * i + 1e-7f; p.y = -3.25f * i; p.z = 1.0f / (i + 1); p.label = (std::uint32_t)(i % 7 + 1); a.push_back(p); }
    for (int mode = 0; mode < 2; mode++) {
      const std::string fn = dir + (mode ? "/l_bin.pcd" : "/l_ascii.pcd");
      CHECK((mode ? pcl::io::savePCDFileBinary(fn, a) : pcl::io::savePCDFileASCII(fn, a)) == 0, "save labelled");
      CHECK(pcl::io::loadPCDFile(fn, b) == 0 && b.size() == a.size(), "load labelled");
      bool same = b.size() == a.size();
      for (std::size_t i = 0; same && i < a.size(); i++) same = a[i].x == b[i].x && a[i].y == b[i].y && a[i].z == b[i].z && a[i].label == b[i].label;
      CHECK(same, "labelled round trip is exact");
      pcl::PointCloud<pcl::PointXYZ> plain;  // a labelled file read into an unlabelled type: label skipped
      CHECK(pcl::io::loadPCDFile(fn, plain) == 0 && plain.size() == a.size() && plain[3].y == a[3].y, "labelled file into PointXYZ");
    }
    const std::string fn = dir + "/foreign.pcd";
    { std::ofstream o(fn.c_str()); o << "# comment\nVERSION .7\nFIELDS x y z intensity label rgb\nSIZE 4 4 4 4 4 4\nTYPE F F F F U F\nCOUNT 1 1 1 1 1 1\nWIDTH 2\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS 2\nDATA ascii\n1 2 3 0.5 9 4.2e-38\n-4 5.5 6 0.25 11 0\n"; }
    CHECK(pcl::io::loadPCDFile(fn, b) == 0 && b.size() == 2 && b[1].x == -4.f && b[1].y == 5.5f && b[0].label == 9 && b[1].label == 11, "foreign header with extra fields");
    CHECK(pcl::io::loadPCDFile(dir + "/does_not_exist.pcd", b) == -1, "missing file");
    // malformed / hostile headers must fail with -1: no out-of-bounds read, no giant allocation, no exception
    const char* bad_headers[] = {
        "FIELDS x y z\nSIZE 2 4 4\nTYPE F F F\nCOUNT 1 1 1\nWIDTH 1\nHEIGHT 1\nPOINTS 1\nDATA binary\n0123456789",          // TYPE F with SIZE 2
        "FIELDS x y z\nSIZE 4 4 3\nTYPE F F U\nCOUNT 1 1 1\nWIDTH 1\nHEIGHT 1\nPOINTS 1\nDATA binary\n01234567890123",      // TYPE U with SIZE 3
        "FIELDS x y z\nSIZE 4 4 4\nTYPE F F F\nCOUNT 1 -3 1\nWIDTH 1\nHEIGHT 1\nPOINTS 1\nDATA binary\n012345678901",       // negative COUNT
        "FIELDS x y z\nSIZE 4 4 4\nTYPE F F F\nCOUNT 1 0 1\nWIDTH 1\nHEIGHT 1\nPOINTS 1\nDATA binary\n012345678901",        // zero COUNT
        "FIELDS x y z\nSIZE 4 4 4\nTYPE F F F\nCOUNT 1 1 1\nWIDTH 1\nHEIGHT 1\nPOINTS 4000000000000\nDATA binary\n0123456", // POINTS beyond the file
        "FIELDS x y z\nSIZE 4 4 4\nTYPE F F F\nCOUNT 1 1 1\nWIDTH 1\nHEIGHT 1\nPOINTS 900000000000\nDATA ascii\n1 2 3\n",   // ... in ascii too
        "FIELDS x y z\nSIZE 4 4 4\nTYPE F F F\nCOUNT 1 1 1\nWIDTH 3\nHEIGHT 1\nPOINTS 3\nDATA binary\n0123456789012345678",  // truncated data
    };
    for (std::size_t i = 0; i < sizeof bad_headers / sizeof bad_headers[0]; i++) {
      const std::string bf = dir + "/bad.pcd";
      { std::ofstream o(bf.c_str(), std::ios::binary); o << bad_headers[i]; }
      CHECK(pcl::io::loadPCDFile(bf, b) == -1, "malformed header rejected");
    }
  }
  // default-constructed pose is the identity (GICP::align(finalCloud) relies on it)
  const double id[7] = {0, 0, 0, 1, 0, 0, 0};
  CHECK(same_pose(Sophus::SE3d(), id, 0.0), "default SE3d is identity");
  if (fails) { std::fprintf(stderr, "%d check(s) failed\n", fails); return 1; }
  std::printf("shim_check OK\n");
  return 0;
}
