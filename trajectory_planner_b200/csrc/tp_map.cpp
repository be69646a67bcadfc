// Host-side occupancy map + loaders for the reference's map files (map/square_static_map.pcd,
// map/*.bt; formats in SURVEY.md §5.8) and this repo's compact .tpm raster format.
#include "tp_map.h"

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>

#include "../../include/tp_b200.h"

static thread_local char g_err[512] = "";
void tp_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
extern "C" const char* tp_last_error(void) { return g_err; }

void tp_map::init(double r, const double o[3], const int d[3], const int inf[3]) {
  res = r;
  for (int a = 0; a < 3; ++a) {
    origin[a] = o[a];
    dims[a] = d[a];
    inflate[a] = inf[a];
  }
  occ.assign(cells(), 0);
  known.assign(cells(), 0);
  inflated.assign(cells(), 0);
}

bool tp_map::index_of(double x, double y, double z, int idx[3]) const {
  const double q[3] = {x, y, z};
  bool in = true;
  for (int a = 0; a < 3; ++a) {
    double f = std::floor((q[a] - origin[a]) / res);
    if (!(f >= -1.0)) {
      idx[a] = -1;
      in = false;
    } else if (f > 2147483000.0) {
      idx[a] = 2147483000;
      in = false;
    } else {
      idx[a] = (int)f;
      if (idx[a] < 0 || idx[a] >= dims[a]) in = false;
    }
  }
  return in;
}

void tp_map::add_occupied_cell(int ix, int iy, int iz) {
  if (ix < 0 || iy < 0 || iz < 0 || ix >= dims[0] || iy >= dims[1] || iz >= dims[2]) return;
  occ[addr(ix, iy, iz)] = 1;
  known[addr(ix, iy, iz)] = 1;
  const int x0 = std::max(0, ix - inflate[0]), x1 = std::min(dims[0] - 1, ix + inflate[0]);
  const int y0 = std::max(0, iy - inflate[1]), y1 = std::min(dims[1] - 1, iy + inflate[1]);
  const int z0 = std::max(0, iz - inflate[2]), z1 = std::min(dims[2] - 1, iz + inflate[2]);
  for (int jx = x0; jx <= x1; ++jx)
    for (int jy = y0; jy <= y1; ++jy) memset(&inflated[addr(jx, jy, z0)], 1, (size_t)(z1 - z0 + 1));
}
void tp_map::add_free_cell(int ix, int iy, int iz) {
  if (ix < 0 || iy < 0 || iz < 0 || ix >= dims[0] || iy >= dims[1] || iz >= dims[2]) return;
  known[addr(ix, iy, iz)] = 1;
}
void tp_map::add_point(double x, double y, double z) {
  int idx[3];
  if (!index_of(x, y, z, idx)) return;
  add_occupied_cell(idx[0], idx[1], idx[2]);
}
bool tp_map::is_inflated_occupied(double x, double y, double z) const {
  int idx[3];
  if (!index_of(x, y, z, idx)) return true;
  return inflated[addr(idx[0], idx[1], idx[2])] != 0;
}
bool tp_map::is_unknown(double x, double y, double z) const {
  int idx[3];
  if (!index_of(x, y, z, idx)) return true;
  return known[addr(idx[0], idx[1], idx[2])] == 0;
}
bool tp_map::is_inflated_occupied_line(const double a[3], const double b[3]) const {
  if (is_inflated_occupied(a[0], a[1], a[2]) || is_inflated_occupied(b[0], b[1], b[2])) return true;
  const double dx = b[0] - a[0], dy = b[1] - a[1], dz = b[2] - a[2];
  const double dist = std::sqrt((dx * dx + dy * dy) + dz * dz);
  const double ux = dx / dist, uy = dy / dist, uz = dz / dist;
  const int steps = (int)(dist / res);
  const double sx = ux * res, sy = uy * res, sz = uz * res;
  for (int i = 1; i < steps; ++i) {
    const double t = (double)i;
    if (is_inflated_occupied(a[0] + t * sx, a[1] + t * sy, a[2] + t * sz)) return true;
  }
  return false;
}
void tp_map::pack(int which, std::vector<uint32_t>& words) const {
  const std::vector<uint8_t>& g = which == 0 ? occ : (which == 1 ? known : inflated);
  const int w = wz();
  words.assign((size_t)dims[0] * dims[1] * w, 0u);
  for (int ix = 0; ix < dims[0]; ++ix)
    for (int iy = 0; iy < dims[1]; ++iy) {
      const uint8_t* col = &g[addr(ix, iy, 0)];
      uint32_t* out = &words[((size_t)ix * dims[1] + iy) * w];
      for (int iz = 0; iz < dims[2]; ++iz)
        if (col[iz]) out[iz >> 5] |= (1u << (iz & 31));
    }
}

// ------------------------------------------------------------------------------ parsers
int tp_read_pcd_ascii(const std::string& path, std::vector<float>& xyz) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) {
    tp_set_error("cannot open %s", path.c_str());
    return TP_ERR_IO;
  }
  std::string buf;
  {
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    buf.resize((size_t)n);
    if (n > 0 && fread(&buf[0], 1, (size_t)n, f) != (size_t)n) {
      fclose(f);
      tp_set_error("short read on %s", path.c_str());
      return TP_ERR_IO;
    }
    fclose(f);
  }
  size_t pos = buf.find("DATA ascii");
  if (pos == std::string::npos) {
    tp_set_error("%s: only ASCII .pcd is supported (no 'DATA ascii' line)", path.c_str());
    return TP_ERR_IO;
  }
  long declared = -1;
  {
    size_t pp = buf.find("POINTS");
    if (pp != std::string::npos && pp < pos) declared = atol(buf.c_str() + pp + 6);
  }
  pos = buf.find('\n', pos);
  if (pos == std::string::npos) pos = buf.size();
  const char* s = buf.c_str() + pos;
  xyz.clear();
  if (declared > 0) xyz.reserve((size_t)declared * 3);
  for (;;) {
    char* end;
    float v = strtof(s, &end);
    if (end == s) break;
    xyz.push_back(v);
    s = end;
  }
  if (xyz.size() % 3 != 0 || (declared >= 0 && (long)(xyz.size() / 3) != declared)) {
    tp_set_error("%s: malformed point list (%zu floats, POINTS %ld)", path.c_str(), xyz.size(), declared);
    return TP_ERR_IO;
  }
  return TP_OK;
}

// OctoMap binary tree: per inner node two bytes = 8 children x 2 bits, LSB first; child index
// bit0->x, bit1->y, bit2->z; 0 unknown, 1 free leaf, 2 occupied leaf, 3 inner; depth-first.
int tp_read_bt(const std::string& path, double& res, std::vector<tp_bt_leaf>& leaves) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) {
    tp_set_error("cannot open %s", path.c_str());
    return TP_ERR_IO;
  }
  std::vector<unsigned char> buf;
  fseek(f, 0, SEEK_END);
  long n = ftell(f);
  fseek(f, 0, SEEK_SET);
  buf.resize((size_t)n);
  if (n > 0 && fread(buf.data(), 1, (size_t)n, f) != (size_t)n) {
    fclose(f);
    tp_set_error("short read on %s", path.c_str());
    return TP_ERR_IO;
  }
  fclose(f);
  const char* marker = "data\n";
  size_t pos = std::string::npos;
  for (size_t i = 0; i + 5 <= buf.size() && i < 4096; ++i)
    if (memcmp(&buf[i], marker, 5) == 0) {
      pos = i + 5;
      break;
    }
  if (pos == std::string::npos) {
    tp_set_error("%s: no 'data' line in .bt header", path.c_str());
    return TP_ERR_IO;
  }
  std::string header((const char*)buf.data(), pos);
  res = -1;
  long declared_nodes = -1;
  {
    std::istringstream hs(header);
    std::string line;
    while (std::getline(hs, line)) {
      if (line.compare(0, 4, "res ") == 0) res = atof(line.c_str() + 4);
      if (line.compare(0, 5, "size ") == 0) declared_nodes = atol(line.c_str() + 5);
    }
  }
  if (!(res > 0)) {
    tp_set_error("%s: missing res in .bt header", path.c_str());
    return TP_ERR_IO;
  }
  struct Item {
    int kx, ky, kz, size;
  };
  std::vector<Item> stack;
  stack.push_back({0, 0, 0, 1 << 16});
  size_t cur = pos;
  long nodes = 1;
  leaves.clear();
  while (!stack.empty()) {
    Item it = stack.back();
    stack.pop_back();
    if (cur + 2 > buf.size()) {
      tp_set_error("%s: truncated .bt stream", path.c_str());
      return TP_ERR_IO;
    }
    unsigned bits = buf[cur] | (buf[cur + 1] << 8);
    cur += 2;
    const int h = it.size >> 1;
    Item inner[8];
    int ninner = 0;
    for (int c = 0; c < 8; ++c) {
      unsigned code = (bits >> (2 * c)) & 3u;
      if (code == 0) continue;
      Item ch = {it.kx + ((c & 1) ? h : 0), it.ky + ((c & 2) ? h : 0), it.kz + ((c & 4) ? h : 0), h};
      ++nodes;
      if (code == 3) {
        if (h < 1) {
          tp_set_error("%s: inner node below the finest level", path.c_str());
          return TP_ERR_IO;
        }
        inner[ninner++] = ch;
      } else
        leaves.push_back({ch.kx, ch.ky, ch.kz, h < 1 ? 1 : h, code == 2 ? 1 : 0});
    }
    for (int i = ninner - 1; i >= 0; --i) stack.push_back(inner[i]);
  }
  if (cur != buf.size() || (declared_nodes >= 0 && declared_nodes != nodes)) {
    tp_set_error("%s: .bt stream/node count mismatch (%zu of %zu bytes, %ld of %ld nodes)", path.c_str(), cur,
                 buf.size(), nodes, declared_nodes);
    return TP_ERR_IO;
  }
  return TP_OK;
}

// ------------------------------------------------------------------------------ C ABI: maps
extern "C" {

tp_map_t* tp_map_create(double res, const double origin[3], const int32_t dims[3], const int32_t inflate[3]) {
  if (!(res > 0) || !origin || !dims || !inflate || dims[0] <= 0 || dims[1] <= 0 || dims[2] <= 0 ||
      (double)dims[0] * dims[1] * dims[2] > 4e9) {
    tp_set_error("tp_map_create: invalid geometry");
    return nullptr;
  }
  tp_map* m = new tp_map();
  m->init(res, origin, dims, inflate);
  return m;
}
void tp_map_destroy(tp_map_t* m) { delete m; }

int tp_map_add_points(tp_map_t* m, const double* xyz, int64_t n) {
  if (!m || (!xyz && n > 0)) return TP_ERR_INVALID_ARG;
  for (int64_t i = 0; i < n; ++i) m->add_point(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
  return TP_OK;
}
int tp_map_add_cells(tp_map_t* m, const int32_t* ijk, int64_t n, int occupied) {
  if (!m || (!ijk && n > 0)) return TP_ERR_INVALID_ARG;
  for (int64_t i = 0; i < n; ++i) {
    if (occupied)
      m->add_occupied_cell(ijk[3 * i], ijk[3 * i + 1], ijk[3 * i + 2]);
    else
      m->add_free_cell(ijk[3 * i], ijk[3 * i + 1], ijk[3 * i + 2]);
  }
  return TP_OK;
}
int tp_map_load_pcd(tp_map_t* m, const char* path) {
  if (!m || !path) return TP_ERR_INVALID_ARG;
  std::vector<float> xyz;
  int rc = tp_read_pcd_ascii(path, xyz);
  if (rc != TP_OK) return rc;
  for (size_t i = 0; i < xyz.size(); i += 3) m->add_point((double)xyz[i], (double)xyz[i + 1], (double)xyz[i + 2]);
  return TP_OK;
}
int tp_map_load_bt(tp_map_t* m, const char* path) {
  if (!m || !path) return TP_ERR_INVALID_ARG;
  double res;
  std::vector<tp_bt_leaf> leaves;
  int rc = tp_read_bt(path, res, leaves);
  if (rc != TP_OK) return rc;
  if (std::fabs(res - m->res) > 1e-12) {
    tp_set_error("%s: tree res %.6f != map res %.6f", path, res, m->res);
    return TP_ERR_INVALID_ARG;
  }
  for (const tp_bt_leaf& l : leaves) {
    // clip the leaf's cell range to the map before expanding it
    for (int dx = 0; dx < l.size; ++dx) {
      const double cx = ((double)(l.kx + dx - 32768) + 0.5) * res;
      const double fx = std::floor((cx - m->origin[0]) / m->res);
      if (fx < 0 || fx >= m->dims[0]) continue;
      for (int dy = 0; dy < l.size; ++dy) {
        const double cy = ((double)(l.ky + dy - 32768) + 0.5) * res;
        const double fy = std::floor((cy - m->origin[1]) / m->res);
        if (fy < 0 || fy >= m->dims[1]) continue;
        for (int dz = 0; dz < l.size; ++dz) {
          const double cz = ((double)(l.kz + dz - 32768) + 0.5) * res;
          const double fz = std::floor((cz - m->origin[2]) / m->res);
          if (fz < 0 || fz >= m->dims[2]) continue;
          if (l.occupied)
            m->add_occupied_cell((int)fx, (int)fy, (int)fz);
          else
            m->add_free_cell((int)fx, (int)fy, (int)fz);
        }
      }
    }
  }
  return TP_OK;
}
int tp_bt_bbox(const char* path, double* res_out, double mn[3], double mx[3], int occupied_only) {
  double res;
  std::vector<tp_bt_leaf> leaves;
  int rc = tp_read_bt(path, res, leaves);
  if (rc != TP_OK) return rc;
  long lo[3] = {1 << 30, 1 << 30, 1 << 30}, hi[3] = {-(1 << 30), -(1 << 30), -(1 << 30)};
  for (const tp_bt_leaf& l : leaves) {
    if (occupied_only && !l.occupied) continue;
    const long k[3] = {l.kx, l.ky, l.kz};
    for (int a = 0; a < 3; ++a) {
      lo[a] = std::min(lo[a], k[a]);
      hi[a] = std::max(hi[a], k[a] + l.size);
    }
  }
  for (int a = 0; a < 3; ++a) {
    mn[a] = (double)(lo[a] - 32768) * res;
    mx[a] = (double)(hi[a] - 32768) * res;
  }
  if (res_out) *res_out = res;
  return TP_OK;
}

// .tpm: "TPM1" | res f64 | origin 3 x f64 | dims 3 x i32 | pad i32 | two RLE streams (occ, known):
// each = n_pairs u64 then (count u32, word u32) pairs over the packed words.
static void rle_write(FILE* f, const std::vector<uint32_t>& w) {
  std::vector<uint32_t> pairs;
  size_t i = 0;
  while (i < w.size()) {
    size_t j = i;
    while (j < w.size() && w[j] == w[i] && j - i < 0xFFFFFFFFu) ++j;
    pairs.push_back((uint32_t)(j - i));
    pairs.push_back(w[i]);
    i = j;
  }
  uint64_t np = pairs.size() / 2;
  fwrite(&np, 8, 1, f);
  fwrite(pairs.data(), 4, pairs.size(), f);
}
static bool rle_read(FILE* f, std::vector<uint32_t>& w, size_t expect) {
  uint64_t np;
  if (fread(&np, 8, 1, f) != 1 || np > (1ull << 32)) return false;
  std::vector<uint32_t> pairs(np * 2);
  if (np && fread(pairs.data(), 4, pairs.size(), f) != pairs.size()) return false;
  w.clear();
  w.reserve(expect);
  for (uint64_t i = 0; i < np; ++i) {
    if (w.size() + pairs[2 * i] > expect) return false;
    w.insert(w.end(), pairs[2 * i], pairs[2 * i + 1]);
  }
  return w.size() == expect;
}
int tp_map_save_tpm(const tp_map_t* m, const char* path) {
  if (!m || !path) return TP_ERR_INVALID_ARG;
  FILE* f = fopen(path, "wb");
  if (!f) {
    tp_set_error("cannot write %s", path);
    return TP_ERR_IO;
  }
  fwrite("TPM1", 1, 4, f);
  int32_t pad = 0;
  fwrite(&pad, 4, 1, f);
  fwrite(&m->res, 8, 1, f);
  fwrite(m->origin, 8, 3, f);
  int32_t d[4] = {m->dims[0], m->dims[1], m->dims[2], 0};
  fwrite(d, 4, 4, f);
  std::vector<uint32_t> w;
  m->pack(0, w);
  rle_write(f, w);
  m->pack(1, w);
  rle_write(f, w);
  fclose(f);
  return TP_OK;
}
tp_map_t* tp_map_load_tpm(const char* path, const int32_t inflate[3]) {
  if (!path || !inflate) return nullptr;
  FILE* f = fopen(path, "rb");
  if (!f) {
    tp_set_error("cannot open %s", path);
    return nullptr;
  }
  char magic[4];
  int32_t pad, d[4];
  double res, origin[3];
  bool ok = fread(magic, 1, 4, f) == 4 && memcmp(magic, "TPM1", 4) == 0 && fread(&pad, 4, 1, f) == 1 &&
            fread(&res, 8, 1, f) == 1 && fread(origin, 8, 3, f) == 3 && fread(d, 4, 4, f) == 4;
  tp_map* m = nullptr;
  if (ok) {
    const int32_t dims[3] = {d[0], d[1], d[2]};
    m = tp_map_create(res, origin, dims, inflate);
  }
  if (m) {
    const int w = m->wz();
    const size_t nw = (size_t)m->dims[0] * m->dims[1] * w;
    std::vector<uint32_t> wo, wk;
    ok = rle_read(f, wo, nw) && rle_read(f, wk, nw);
    if (ok) {
      for (int ix = 0; ix < m->dims[0]; ++ix)
        for (int iy = 0; iy < m->dims[1]; ++iy)
          for (int iz = 0; iz < m->dims[2]; ++iz) {
            const size_t wi = ((size_t)ix * m->dims[1] + iy) * w + (iz >> 5);
            if ((wo[wi] >> (iz & 31)) & 1u)
              m->add_occupied_cell(ix, iy, iz);
            else if ((wk[wi] >> (iz & 31)) & 1u)
              m->add_free_cell(ix, iy, iz);
          }
    }
  }
  fclose(f);
  if (!ok || !m) {
    tp_set_error("%s: not a valid .tpm file", path);
    delete m;
    return nullptr;
  }
  return m;
}
int tp_map_info_get(const tp_map_t* m, tp_map_info* info) {
  if (!m || !info) return TP_ERR_INVALID_ARG;
  info->res = m->res;
  for (int a = 0; a < 3; ++a) {
    info->origin[a] = m->origin[a];
    info->dims[a] = m->dims[a];
    info->inflate[a] = m->inflate[a];
  }
  int64_t no = 0, ni = 0, nk = 0;
  for (size_t i = 0; i < m->cells(); ++i) {
    no += m->occ[i];
    ni += m->inflated[i];
    nk += m->known[i];
  }
  info->n_occupied = no;
  info->n_inflated = ni;
  info->n_known = nk;
  info->packed_bytes = (int64_t)m->dims[0] * m->dims[1] * m->wz() * 4;
  return TP_OK;
}
int tp_map_get_grid(const tp_map_t* m, int which, uint8_t* out) {
  if (!m || !out || which < 0 || which > 2) return TP_ERR_INVALID_ARG;
  const std::vector<uint8_t>& g = which == 0 ? m->occ : (which == 1 ? m->known : m->inflated);
  memcpy(out, g.data(), g.size());
  return TP_OK;
}

}  // extern "C"
