// apps/io.hpp -- the reference's on-disk formats, restated:
//   cloud  "x y z nx ny nz" per line (include/common.h:224-239 loadXYZ).  The reference's `while(file){...push_back}`
//          appends one extra copy of the last point at end-of-file; that quirk is reproduced only with ref_eof_quirk.
//   pose   16 numbers, 4x4 row-major (common.h:172-187 loadMatrix4d)
//   folder files whose name starts with a prefix and ends in .txt/.xyz, sorted by length then lexicographically
//          (common.h:106-166 getAllFilesFromFolder)
#pragma once
#include <dirent.h>
#include <algorithm>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>
#include "mini_types.hpp"

namespace io {
inline std::vector<std::string> files_with_prefix(const std::string& dir, const std::string& prefix) {
  std::vector<std::string> out;
  DIR* d = opendir(dir.c_str());
  if (!d) { std::cerr << "Could not open directory " << dir << std::endl; return out; }
  while (dirent* e = readdir(d)) {
    const std::string n(e->d_name);
    const bool suffix = n.size() > 4 && (n.compare(n.size() - 4, 4, ".txt") == 0 || n.compare(n.size() - 4, 4, ".xyz") == 0);
    if (suffix && n.compare(0, prefix.size(), prefix) == 0) out.push_back(dir + "/" + n);
  }
  closedir(d);
  std::sort(out.begin(), out.end(), [](const std::string& a, const std::string& b) { return a.size() != b.size() ? a.size() < b.size() : a < b; });
  return out;
}
inline bool load_xyz(const std::string& fn, std::vector<Eigen::Vector3d>& pts, std::vector<Eigen::Vector3d>& nor, bool ref_eof_quirk) {
  std::ifstream f(fn.c_str());
  if (f.fail()) { std::cerr << fn << " could not be opened" << std::endl; return false; }
  Eigen::Vector3d p{}, n{};
  while (f >> p.v[0] >> p.v[1] >> p.v[2] >> n.v[0] >> n.v[1] >> n.v[2]) { pts.push_back(p); nor.push_back(n); }
  if (ref_eof_quirk && !pts.empty()) { pts.push_back(pts.back()); nor.push_back(nor.back()); }
  return true;
}
inline bool load_pose(const std::string& fn, Eigen::Isometry3d& P) {
  std::ifstream f(fn.c_str());
  if (f.fail()) { std::cerr << fn << " could not be opened" << std::endl; return false; }
  double a[16] = {0}; a[15] = 1; int i = 0;
  while (i < 16 && (f >> a[i])) ++i;
  for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) P(r, c) = a[4 * r + c];
  return true;
}
inline void save_pose(const std::string& fn, const Eigen::Isometry3d& P) {
  std::ofstream f(fn.c_str()); f.precision(17);
  for (int r = 0; r < 4; ++r) { for (int c = 0; c < 4; ++c) f << P(r, c) << (c == 3 ? "\n" : " "); }
}
}  // namespace io
