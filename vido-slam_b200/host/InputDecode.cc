// InputDecode.cc -- see InputDecode.h.  PNG per the W3C PNG specification (second edition): signature, IHDR, the concatenated
// IDAT stream inflated with zlib, the five scan-line filters (None, Sub, Up, Average, Paeth), 16-bit samples big-endian in the file.
#include "InputDecode.h"

#include <zlib.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>

namespace VIDO_SLAM {
namespace io {
namespace {
bool fail(std::string* err, const std::string& msg) {
  if (err) *err = msg;
  return false;
}
uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
bool slurp(const std::string& path, std::vector<uint8_t>& buf) {
  FILE* fh = fopen(path.c_str(), "rb");
  if (!fh) return false;
  fseek(fh, 0, SEEK_END);
  const long n = ftell(fh);
  fseek(fh, 0, SEEK_SET);
  buf.resize(n > 0 ? (size_t)n : 0);
  const bool ok = n >= 0 && fread(buf.data(), 1, buf.size(), fh) == buf.size();
  fclose(fh);
  return ok;
}
inline int paeth(int a, int b, int c) {
  const int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
  return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}
}  // namespace

bool read_png(const std::string& path, Image& out, std::string* err) {
  std::vector<uint8_t> f;
  if (!slurp(path, f)) return fail(err, "cannot read " + path);
  static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
  if (f.size() < 8 + 25 || memcmp(f.data(), sig, 8) != 0) return fail(err, path + ": not a PNG file");
  size_t pos = 8;
  int w = 0, h = 0, depth = 0, ctype = -1;
  std::vector<uint8_t> z;
  bool end = false;
  while (!end && pos + 12 <= f.size()) {
    const uint32_t len = be32(&f[pos]);
    const uint8_t* type = &f[pos + 4];
    if (pos + 12 + (size_t)len > f.size()) return fail(err, path + ": truncated chunk");
    const uint8_t* d = &f[pos + 8];
    if (!memcmp(type, "IHDR", 4)) {
      if (len < 13) return fail(err, path + ": bad IHDR");
      w = (int)be32(d); h = (int)be32(d + 4); depth = d[8]; ctype = d[9];
      if (d[10] != 0 || d[11] != 0) return fail(err, path + ": unknown compression / filter method");
      if (d[12] != 0) return fail(err, path + ": interlaced PNG is not supported");
    } else if (!memcmp(type, "IDAT", 4)) {
      z.insert(z.end(), d, d + len);
    } else if (!memcmp(type, "IEND", 4)) {
      end = true;
    }
    pos += 12 + (size_t)len;
  }
  int ch = 0;
  switch (ctype) {
    case 0: ch = 1; break;
    case 2: ch = 3; break;
    case 4: ch = 2; break;
    case 6: ch = 4; break;
    default: return fail(err, path + ": colour type " + std::to_string(ctype) + " is not supported");
  }
  if (w <= 0 || h <= 0 || (depth != 8 && depth != 16)) return fail(err, path + ": unsupported size / bit depth");
  const size_t bpp = (size_t)ch * depth / 8, row = bpp * w;
  std::vector<uint8_t> raw((row + 1) * (size_t)h);
  uLongf got = (uLongf)raw.size();
  if (uncompress(raw.data(), &got, z.data(), (uLong)z.size()) != Z_OK || got != raw.size()) return fail(err, path + ": inflate failed");
  out.width = w; out.height = h; out.channels = ch; out.bit_depth = depth;
  out.data.assign(row * h, 0);
  const std::vector<uint8_t> zero(row, 0);
  for (int y = 0; y < h; y++) {
    const uint8_t* src = &raw[(row + 1) * (size_t)y];
    const int ft = src[0];
    src++;
    uint8_t* cur = &out.data[row * (size_t)y];
    const uint8_t* up = y ? cur - row : zero.data();
    switch (ft) {
      case 0: memcpy(cur, src, row); break;
      case 1: for (size_t i = 0; i < row; i++) cur[i] = (uint8_t)(src[i] + (i >= bpp ? cur[i - bpp] : 0)); break;
      case 2: for (size_t i = 0; i < row; i++) cur[i] = (uint8_t)(src[i] + up[i]); break;
      case 3: for (size_t i = 0; i < row; i++) cur[i] = (uint8_t)(src[i] + (((i >= bpp ? cur[i - bpp] : 0) + up[i]) >> 1)); break;
      case 4: for (size_t i = 0; i < row; i++) cur[i] = (uint8_t)(src[i] + paeth(i >= bpp ? cur[i - bpp] : 0, up[i], i >= bpp ? up[i - bpp] : 0)); break;
      default: return fail(err, path + ": unknown filter type");
    }
  }
  if (depth == 16)   // file order is big-endian
    for (size_t i = 0; i + 1 < out.data.size(); i += 2) { const uint8_t t = out.data[i]; out.data[i] = out.data[i + 1]; out.data[i + 1] = t; }
  return true;
}

bool read_flo(const std::string& path, int& width, int& height, std::vector<float>& uv, std::string* err) {
  std::vector<uint8_t> f;
  if (!slurp(path, f)) return fail(err, "cannot read " + path);
  if (f.size() < 12 || memcmp(f.data(), "PIEH", 4) != 0) return fail(err, path + ": not a .flo file");
  int32_t w, h;
  memcpy(&w, &f[4], 4); memcpy(&h, &f[8], 4);
  if (w <= 0 || h <= 0 || f.size() < 12 + (size_t)w * h * 8) return fail(err, path + ": truncated .flo file");
  width = w; height = h;
  uv.resize((size_t)w * h * 2);
  memcpy(uv.data(), &f[12], uv.size() * 4);
  return true;
}

bool load_kaist_imu(const std::string& path, std::vector<ImuSample>& out, std::string* err) {
  std::ifstream in(path);
  if (!in.is_open()) return fail(err, "cannot read " + path);
  std::string s;
  while (std::getline(in, s)) {
    if (s.empty() || s[0] == '#') continue;
    std::vector<double> v;
    std::istringstream line(s);
    std::string field;
    while (std::getline(line, field, ',')) v.push_back(atof(field.c_str()));
    if (v.size() < 14) return fail(err, path + ": a line has fewer than 14 columns");
    out.push_back({v[0] / 1e9, (float)v[11], (float)v[12], (float)v[13], (float)v[8], (float)v[9], (float)v[10]});
  }
  return true;
}

bool load_kaist_timestamps(const std::string& image_dir, std::vector<std::string>& names, std::vector<double>& times, std::string* err) {
  const std::string path = image_dir + "/../vTimestampsImage.txt";
  FILE* fh = fopen(path.c_str(), "r");
  if (!fh) return fail(err, "cannot read " + path);
  char line[256];
  if (fgets(line, sizeof line, fh)) {   // header
    while (fgets(line, sizeof line, fh) && line[0] != '\n' && line[0] != '\0') {
      const long double s = strtold(line, nullptr);   // `long double s; ss >> s;`
      char printed[64];
      snprintf(printed, sizeof printed, "%Lf", s);    // std::to_string(long double)
      names.push_back(std::string(printed).substr(0, 19) + ".png");
      times.push_back((double)(s / 1e9));
    }
  }
  fclose(fh);
  return true;
}

std::vector<ImuSample> imu_between(const std::vector<ImuSample>& all, double t_last, double t_cur) {
  std::vector<ImuSample> r;
  for (const ImuSample& s : all)
    if (t_last <= s.t && s.t <= t_cur) r.push_back(s);
  return r;
}

}  // namespace io
}  // namespace VIDO_SLAM

using namespace VIDO_SLAM::io;

extern "C" int vido_io_read_png(const char* path, int32_t* width, int32_t* height, int32_t* channels, int32_t* bit_depth, uint8_t* buf, size_t cap) {
  Image im;
  if (!path || !read_png(path, im)) return -1;
  if (width) *width = im.width;
  if (height) *height = im.height;
  if (channels) *channels = im.channels;
  if (bit_depth) *bit_depth = im.bit_depth;
  if (buf) {
    if (cap < im.data.size()) return -2;
    memcpy(buf, im.data.data(), im.data.size());
  }
  return 0;
}
extern "C" int vido_io_read_flo(const char* path, int32_t* width, int32_t* height, float* uv, size_t cap_floats) {
  int w = 0, h = 0;
  std::vector<float> v;
  if (!path || !read_flo(path, w, h, v)) return -1;
  if (width) *width = w;
  if (height) *height = h;
  if (uv) {
    if (cap_floats < v.size()) return -2;
    memcpy(uv, v.data(), v.size() * 4);
  }
  return 0;
}
extern "C" int vido_io_load_kaist_imu(const char* path, double* rows7, int cap_rows) {
  std::vector<ImuSample> v;
  if (!path || !load_kaist_imu(path, v)) return -1;
  for (int i = 0; i < (int)v.size() && i < cap_rows && rows7; i++) {
    double* r = rows7 + 7 * (size_t)i;
    r[0] = v[i].t; r[1] = v[i].ax; r[2] = v[i].ay; r[3] = v[i].az; r[4] = v[i].wx; r[5] = v[i].wy; r[6] = v[i].wz;
  }
  return (int)v.size();
}
extern "C" int vido_io_load_kaist_timestamps(const char* image_dir, double* times, char* names20, int cap) {
  std::vector<std::string> names;
  std::vector<double> t;
  if (!image_dir || !load_kaist_timestamps(image_dir, names, t)) return -1;
  for (int i = 0; i < (int)t.size() && i < cap; i++) {
    if (times) times[i] = t[i];
    if (names20) { memset(names20 + 20 * (size_t)i, 0, 20); strncpy(names20 + 20 * (size_t)i, names[i].c_str(), 19); }
  }
  return (int)t.size();
}
