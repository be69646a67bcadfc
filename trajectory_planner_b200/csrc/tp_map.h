// Host-side occupancy map: the occMap contract of include/tp_b200.h (stand-in for the external
// mapManager::occMap the reference's ViGO path queries; call sites bsplineTraj.h:197,199,312,319,
// bsplineTraj.cpp:292,412,435,736-783,841, astarOcc.h:58).  Byte grids for insertion, bit-packed
// (z fastest, ceil(nz/32) uint32 words per (x,y) column) for the GPU.
#pragma once
#include <cmath>
#include <cstdint>
#include <string>
#include <vector>

struct tp_map {
  double res = 0.1;
  double origin[3] = {0, 0, 0};
  int dims[3] = {0, 0, 0};
  int inflate[3] = {0, 0, 0};
  std::vector<uint8_t> occ, known, inflated;

  size_t cells() const { return (size_t)dims[0] * dims[1] * dims[2]; }
  size_t addr(int ix, int iy, int iz) const { return ((size_t)ix * dims[1] + iy) * dims[2] + iz; }
  int wz() const { return (dims[2] + 31) / 32; }

  void init(double r, const double o[3], const int d[3], const int inf[3]);
  bool index_of(double x, double y, double z, int idx[3]) const;
  void add_occupied_cell(int ix, int iy, int iz);
  void add_free_cell(int ix, int iy, int iz);
  void add_point(double x, double y, double z);
  bool is_inflated_occupied(double x, double y, double z) const;
  bool is_unknown(double x, double y, double z) const;
  bool is_inflated_occupied_line(const double a[3], const double b[3]) const;
  // which: 0 occ, 1 known, 2 inflated
  void pack(int which, std::vector<uint32_t>& words) const;
};

// parsers (return 0 on success, negative TP_ERR_* otherwise; message in tp_set_error)
int tp_read_pcd_ascii(const std::string& path, std::vector<float>& xyz);
struct tp_bt_leaf {
  int kx, ky, kz, size, occupied;
};
int tp_read_bt(const std::string& path, double& res, std::vector<tp_bt_leaf>& leaves);

void tp_set_error(const char* fmt, ...);
