// cloud_io.cu - the wire / on-disk formats either side of the registration path.
//
//   launch_pc2_unpack   pcl::fromROSMsg(sensor_msgs::PointCloud2, PointCloud<PointXYZRGB>) (reference src/node.cpp:37,41):
//                       gathers the FLOAT32 fields x, y, z and the 4-byte rgb / rgba field of every point of a
//                       PointCloud2 payload into the 32-byte pcl::PointXYZRGB rows the rest of the library takes
//                       (x, y, z, 1.0f, rgba, 12 zero bytes).  One pass, HBM-bound: point_step + 32 bytes per point.
//   PcdFile             pcl::io::loadPCDFile<pcl::PointXYZRGB> (reference src/load_and_publish_clouds.cpp:75): PCL
//                       1.8.1's PCDReader reads the file into a PCLPointCloud2 blob (header -> field table; body ascii,
//                       binary or binary_compressed = LZF over a field-major copy) and fromPCLPointCloud2 then maps
//                       the fields by name - which is the unpack kernel above.  The reader is host code (file I/O and
//                       text parsing); the field mapping runs on the GPU.
#include <cerrno>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>

#include "cloud_io.hpp"

namespace gicpb {

namespace {

// 4 bytes at any alignment
__device__ __forceinline__ uint32_t load_u32(const unsigned char* p, bool aligned) {
  if (aligned) return *reinterpret_cast<const uint32_t*>(p);
  return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}

__global__ void __launch_bounds__(256) pc2_unpack_kernel(const unsigned char* __restrict__ data, int64_t n, int64_t width,
                                                          int64_t point_step, int64_t row_step, int off_x, int off_y,
                                                          int off_z, int off_rgb, bool aligned, float4* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t row = i / width, col = i - row * width;
  const unsigned char* p = data + row * row_step + col * point_step;
  float4 a, b;
  a.x = __uint_as_float(load_u32(p + off_x, aligned));
  a.y = __uint_as_float(load_u32(p + off_y, aligned));
  a.z = __uint_as_float(load_u32(p + off_z, aligned));
  a.w = 1.0f;                                                             // PointXYZRGB(): data[3] = 1
  b.x = __uint_as_float(off_rgb >= 0 ? load_u32(p + off_rgb, aligned) : 0xff000000u);  // PointXYZRGB(): r = g = b = 0, a = 255
  b.y = b.z = b.w = 0.f;
  out[2 * i] = a;
  out[2 * i + 1] = b;
}

}  // namespace

void launch_pc2_unpack(const unsigned char* data, int64_t n, int64_t width, int64_t point_step, int64_t row_step, int off_x,
                       int off_y, int off_z, int off_rgb, float4* out, cudaStream_t stream) {
  if (n <= 0) return;
  const bool aligned = (reinterpret_cast<uintptr_t>(data) % 4 == 0) && point_step % 4 == 0 && row_step % 4 == 0 &&
                       off_x % 4 == 0 && off_y % 4 == 0 && off_z % 4 == 0 && (off_rgb < 0 || off_rgb % 4 == 0);
  pc2_unpack_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(data, n, width, point_step, row_step, off_x, off_y,
                                                                     off_z, off_rgb, aligned, out);
  GICPB_LAUNCHED();
}

// ---- PCD files ------------------------------------------------------------------------------------------------------------
namespace {

std::vector<std::string> split_ws(const std::string& line) {  // boost::split(st, line, is_any_of("\t\r "), token_compress_on)
  std::vector<std::string> out;
  size_t i = 0;
  while (i < line.size()) {
    while (i < line.size() && (line[i] == ' ' || line[i] == '\t' || line[i] == '\r')) ++i;
    size_t j = i;
    while (j < line.size() && line[j] != ' ' && line[j] != '\t' && line[j] != '\r') ++j;
    if (j > i) out.push_back(line.substr(i, j - i));
    i = j;
  }
  return out;
}

int64_t to_int(const std::string& s, const char* what) {
  errno = 0;
  char* end = nullptr;
  const long long v = std::strtoll(s.c_str(), &end, 10);
  if (errno || end == s.c_str()) throw ArgError(std::string("PCD header: bad ") + what + " '" + s + "'");
  return v;
}

// liblzf's lzf_decompress (the codec PCL bundles for DATA binary_compressed); returns the number of bytes written, 0 on error
size_t lzf_decompress(const unsigned char* in, size_t in_len, unsigned char* out, size_t out_len) {
  const unsigned char* ip = in;
  const unsigned char* const in_end = in + in_len;
  unsigned char* op = out;
  unsigned char* const out_end = out + out_len;
  while (ip < in_end) {
    unsigned ctrl = *ip++;
    if (ctrl < 32) {  // literal run of ctrl + 1 bytes
      ++ctrl;
      if (op + ctrl > out_end || ip + ctrl > in_end) return 0;
      std::memcpy(op, ip, ctrl);
      op += ctrl;
      ip += ctrl;
    } else {  // back reference
      unsigned len = ctrl >> 5;
      if (ip >= in_end) return 0;
      if (len == 7) {
        len += *ip++;
        if (ip >= in_end) return 0;
      }
      const size_t dist = ((size_t)(ctrl & 0x1f) << 8) + *ip++ + 1;
      len += 2;
      if (op + len > out_end || dist > (size_t)(op - out)) return 0;
      const unsigned char* ref = op - dist;
      for (unsigned k = 0; k < len; ++k) *op++ = *ref++;  // may overlap
    }
  }
  return (size_t)(op - out);
}

template <typename T>
void store(unsigned char* dst, T v) {
  std::memcpy(dst, &v, sizeof(T));
}

// PCL's copyStringValue<T>: "nan" -> quiet_NaN (0 for integer types) and the cloud stops being dense
void parse_ascii(const std::string& tok, const PcdField& f, unsigned char* dst, bool* dense) {
  const bool nan = tok == "nan";
  if (nan) *dense = false;
  const char* s = tok.c_str();
  if (f.type == 'F') {
    if (f.size == 4) store<float>(dst, nan ? std::nanf("") : std::strtof(s, nullptr));
    else store<double>(dst, nan ? std::nan("") : std::strtod(s, nullptr));
  } else if (f.type == 'I') {
    const long long v = nan ? 0 : std::strtoll(s, nullptr, 10);
    if (f.size == 1) store<int8_t>(dst, (int8_t)v);
    else if (f.size == 2) store<int16_t>(dst, (int16_t)v);
    else if (f.size == 4) store<int32_t>(dst, (int32_t)v);
    else store<int64_t>(dst, (int64_t)v);
  } else {
    const unsigned long long v = nan ? 0 : std::strtoull(s, nullptr, 10);
    if (f.size == 1) store<uint8_t>(dst, (uint8_t)v);
    else if (f.size == 2) store<uint16_t>(dst, (uint16_t)v);
    else if (f.size == 4) store<uint32_t>(dst, (uint32_t)v);
    else store<uint64_t>(dst, (uint64_t)v);
  }
}

}  // namespace

void PcdFile::read_header(const std::string& path) {
  std::ifstream fs(path, std::ios::binary);
  if (!fs) throw ArgError("could not open PCD file " + path);
  fields.clear();
  width = height = points = 0;
  point_step = 0;
  data_kind = -1;
  bool have_height = false, have_points = false;
  std::string line;
  while (std::getline(fs, line)) {
    const std::vector<std::string> st = split_ws(line);
    if (st.empty() || st[0][0] == '#') continue;
    const std::string& key = st[0];
    if (key == "VERSION" || key == "VIEWPOINT") continue;
    if (key == "FIELDS" || key == "COLUMNS") {
      fields.resize(st.size() - 1);
      int off = 0;
      for (size_t i = 0; i < fields.size(); ++i) {  // until SIZE / TYPE / COUNT say otherwise: float32, count 1
        fields[i] = PcdField{st[i + 1], off, 4, 'F', 1};
        off += 4;
      }
      point_step = off;
      continue;
    }
    if (key == "SIZE" || key == "TYPE" || key == "COUNT") {
      if (st.size() - 1 != fields.size()) throw ArgError("PCD header: " + key + " does not match FIELDS");
      int off = 0;
      for (size_t i = 0; i < fields.size(); ++i) {
        PcdField& f = fields[i];
        if (key == "SIZE") {
          f.size = (int)to_int(st[i + 1], "SIZE");
          if (f.size != 1 && f.size != 2 && f.size != 4 && f.size != 8) throw ArgError("PCD header: unsupported SIZE " + st[i + 1]);
        } else if (key == "TYPE") {
          f.type = st[i + 1][0];
          if (f.type != 'I' && f.type != 'U' && f.type != 'F') throw ArgError("PCD header: unsupported TYPE " + st[i + 1]);
        } else {
          f.count = (int)to_int(st[i + 1], "COUNT");
          if (f.count < 1) throw ArgError("PCD header: COUNT must be >= 1");
        }
        f.offset = off;
        off += f.size * f.count;
      }
      point_step = off;
      continue;
    }
    if (key == "WIDTH" && st.size() > 1) { width = to_int(st[1], "WIDTH"); continue; }
    if (key == "HEIGHT" && st.size() > 1) { height = to_int(st[1], "HEIGHT"); have_height = true; continue; }
    if (key == "POINTS" && st.size() > 1) { points = to_int(st[1], "POINTS"); have_points = true; continue; }
    if (key == "DATA" && st.size() > 1) {
      if (st[1] == "ascii") data_kind = 0;
      else if (st[1] == "binary") data_kind = 1;
      else if (st[1] == "binary_compressed") data_kind = 2;
      else throw ArgError("PCD header: unknown DATA kind " + st[1]);
      data_offset = (int64_t)fs.tellg();
      break;
    }
    throw ArgError("PCD header: unknown entry " + key);
  }
  if (data_kind < 0) throw ArgError("PCD header: no DATA entry in " + path);
  if (fields.empty() || point_step <= 0) throw ArgError("PCD header: no FIELDS");
  if (!have_height) {  // unorganised: one row
    height = 1;
    if (width == 0 && have_points) width = points;
  }
  if (!have_points) points = width * height;
  if (width < 0 || height < 0 || width * height != points) throw ArgError("PCD header: WIDTH * HEIGHT != POINTS");
  off_x = off_y = off_z = off_rgb = -1;
  for (const PcdField& f : fields) {  // pcl::FieldMatches: same name, datatype and count; rgb also takes rgba (UINT32)
    const bool f32 = f.type == 'F' && f.size == 4 && f.count == 1;
    if (f.name == "x" && f32) off_x = f.offset;
    if (f.name == "y" && f32) off_y = f.offset;
    if (f.name == "z" && f32) off_z = f.offset;
    if ((f.name == "rgb" && f32) || (f.name == "rgba" && f.type == 'U' && f.size == 4 && f.count == 1)) off_rgb = f.offset;
  }
}

void PcdFile::read_body(const std::string& path, unsigned char* blob) {
  const size_t total = (size_t)points * point_step;
  is_dense = true;
  std::ifstream fs(path, std::ios::binary);
  if (!fs) throw ArgError("could not open PCD file " + path);
  fs.seekg(data_offset);
  if (data_kind == 0) {
    std::memset(blob, 0, total);
    std::string line;
    int64_t idx = 0;
    while (idx < points && std::getline(fs, line)) {
      const std::vector<std::string> st = split_ws(line);
      if (st.empty()) continue;
      size_t tok = 0;
      for (const PcdField& f : fields) {
        if (f.name == "_") {  // padding: its tokens are skipped
          tok += f.count;
          continue;
        }
        for (int c = 0; c < f.count; ++c, ++tok)
          if (tok < st.size()) parse_ascii(st[tok], f, blob + (size_t)idx * point_step + f.offset + c * f.size, &is_dense);
      }
      ++idx;
    }
    if (idx != points) throw ArgError("PCD file: fewer data lines than POINTS in " + path);
    return;
  }
  if (data_kind == 1) {
    fs.read(reinterpret_cast<char*>(blob), (std::streamsize)total);
    if ((size_t)fs.gcount() != total) throw ArgError("PCD file: binary body is shorter than POINTS * point size in " + path);
  } else {
    for (const PcdField& f : fields)
      if (f.name == "_") throw ArgError("PCD file: padding fields in a binary_compressed body are not supported");
    uint32_t sizes[2] = {0, 0};  // compressed size, uncompressed size
    fs.read(reinterpret_cast<char*>(sizes), 8);
    if (fs.gcount() != 8) throw ArgError("PCD file: truncated binary_compressed body in " + path);
    if (sizes[1] != total) throw ArgError("PCD file: uncompressed size does not match the header in " + path);
    std::vector<unsigned char> comp(sizes[0]), soa(total);
    fs.read(reinterpret_cast<char*>(comp.data()), (std::streamsize)comp.size());
    if ((size_t)fs.gcount() != comp.size()) throw ArgError("PCD file: truncated binary_compressed body in " + path);
    if (total && lzf_decompress(comp.data(), comp.size(), soa.data(), total) != total)
      throw ArgError("PCD file: LZF stream is corrupt in " + path);
    // field-major ("xxyyzz") -> point-major ("xyz")
    size_t toff = 0;
    for (const PcdField& f : fields) {
      const size_t fs_bytes = (size_t)f.size * f.count;
      const unsigned char* srcp = soa.data() + toff;
      for (int64_t i = 0; i < points; ++i) std::memcpy(blob + (size_t)i * point_step + f.offset, srcp + (size_t)i * fs_bytes, fs_bytes);
      toff += fs_bytes * (size_t)points;
    }
  }
  // PCDReader::read: a non-finite value in any FLOAT32 / FLOAT64 field makes the cloud not dense
  for (const PcdField& f : fields) {
    if (f.name == "_" || f.type != 'F' || f.size < 4) continue;
    for (int64_t i = 0; i < points && is_dense; ++i)
      for (int c = 0; c < f.count; ++c) {
        const unsigned char* src = blob + (size_t)i * point_step + f.offset + f.size * c;
        if (f.size == 4) {
          float v;
          std::memcpy(&v, src, 4);
          if (!std::isfinite(v)) is_dense = false;
        } else {
          double v;
          std::memcpy(&v, src, 8);
          if (!std::isfinite(v)) is_dense = false;
        }
      }
  }
}

}  // namespace gicpb
