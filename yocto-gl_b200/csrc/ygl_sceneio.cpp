// ygl_sceneio.cpp — scene ingestion: Yocto/GL JSON scene files (formats 4.0 and 4.2 / 5.0) with
// PLY shapes and PNG / Radiance-HDR textures, read straight into the flat host arrays a ygl_scene_desc views, ready
// for ygl_scene_create / ygl_bvh_build / ygl_lights_create. Host C++ only.
//
// Behavioural contract (what the arrays must equal, bit for bit, for the renders to match the reference's):
//   load_json_scene     libs/yocto/yocto_sceneio.cpp:3618-3860   keys, defaults (yocto_scene.h:83-160), lookat, fix-ups
//   load_shape (.ply)   libs/yocto/yocto_sceneio.cpp:1008-1035   + the ply getters of yocto_modelio.h:618-815
//   load_texture        libs/yocto/yocto_sceneio.cpp:1796-1837   .hdr -> float rgba (stb's RGBE rule), .png -> byte rgba
//   add_missing_camera / add_missing_radius          yocto_sceneio.cpp:2119-2148
// Numbers in JSON are read as doubles and narrowed to float, as nlohmann::json does for the reference.
// Subdivs (.obj control meshes) are read and tesselated at load (tesselate_subdivs, yocto_scene.cpp:739-813).
// Shapes: .ply, .obj, binary .stl; textures .png, .jpg, .hdr, .exr; scenes .json, .ply, .gltf / .glb, .obj. Not built: pbrt /
// mitsuba scenes, tiled EXR: refused.
#include <sched.h>
#include <sys/stat.h>
#include <zlib.h>

#include <algorithm>
#include <array>
#include <atomic>
#include <cerrno>
#include <charconv>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <string>
#include <thread>
#include <unordered_map>
#include <utility>
#include <vector>

#include "../../include/ygl_b200.h"
#include "ygl_hostio.h"

namespace ygl_io {
bool read_file(const std::string& filename, std::vector<uint8_t>& data, std::string& error) {
  FILE* f = fopen(filename.c_str(), "rb");
  if (!f) return error = "cannot open " + filename, false;
  fseek(f, 0, SEEK_END);
  long size = ftell(f);
  fseek(f, 0, SEEK_SET);
  data.resize(size > 0 ? (size_t)size : 0);
  size_t got = data.empty() ? 0 : fread(data.data(), 1, data.size(), f);
  fclose(f);
  if (got != data.size()) return error = "cannot read " + filename, false;
  return true;
}
std::string path_extension(const std::string& path) {
  auto pos = path.find_last_of('.');
  if (pos == std::string::npos) return {};
  auto ext = path.substr(pos);
  for (auto& c : ext) c = (char)tolower((unsigned char)c);
  return ext;
}
}  // namespace ygl_io
using namespace ygl_io;

namespace {

// Cores this process may really use: the smaller of the affinity mask and the cgroup CPU quota (a container lease of
// 16 CPUs on a 128-thread host reports 128 from hardware_concurrency; starting 128 threads there only adds switching)
inline int host_parallelism() {
  static const int cores = []() {
    int n = (int)std::thread::hardware_concurrency();
    cpu_set_t set;
    if (sched_getaffinity(0, sizeof(set), &set) == 0 && CPU_COUNT(&set) > 0) n = std::min(n > 0 ? n : CPU_COUNT(&set), CPU_COUNT(&set));
    if (FILE* f = fopen("/sys/fs/cgroup/cpu.max", "r")) {
      long long quota = 0, period = 0;
      if (fscanf(f, "%lld %lld", &quota, &period) == 2 && quota > 0 && period > 0)
        n = std::min<long long>(n > 0 ? n : 1, (quota + period - 1) / period);
      fclose(f);
    }
    return std::max(1, n);
  }();
  return cores;
}

// the library's per-thread error string lives in ygl_api.cpp

// ---------------------------------------------------------------------------------------------------------------
// files
// ---------------------------------------------------------------------------------------------------------------
std::string path_dirname(const std::string& path) {
  auto pos = path.find_last_of("/\\");
  return pos == std::string::npos ? std::string{} : path.substr(0, pos);
}
std::string path_join(const std::string& a, const std::string& b) { return a.empty() ? b : a + "/" + b; }

// ---------------------------------------------------------------------------------------------------------------
// JSON: a small document model with the acceptance rules and number model of nlohmann::ordered_json, the reference's
// parser: objects keep insertion order and a repeated key overwrites the first one's value in place; a number without
// fraction or exponent is an integer (int64 if negative, uint64 otherwise - so "-0" is +0 - and a double only when it
// fits neither), everything else a double that must be finite; strict number grammar (no "+1", ".5", "1.", "01", "inf");
// strings must be well-formed UTF-8 without raw control characters, \u escapes pair up surrogates; one optional BOM.
// ---------------------------------------------------------------------------------------------------------------
struct JValue {
  enum Type { Null, Bool, Number, String, Array, Object } type = Null;
  enum NumberKind : uint8_t { Float, Signed, Unsigned };
  bool                                           boolean = false;
  NumberKind                                     kind    = Float;
  double                                         number  = 0;  // Float
  int64_t                                        inumber = 0;  // Signed
  uint64_t                                       unumber = 0;  // Unsigned
  std::string                                    string;
  std::vector<JValue>                            array;
  std::vector<std::pair<std::string, JValue>>    object;
  const JValue* find(const char* key) const {
    if (type != Object) return nullptr;
    for (auto& kv : object)
      if (kv.first == key) return &kv.second;
    return nullptr;
  }
  // static_cast<T>(stored value), what nlohmann's get<T>() does for a number
  float as_float() const { return kind == Float ? (float)number : kind == Signed ? (float)inumber : (float)unumber; }
  int   as_int() const { return kind == Float ? (int)number : kind == Signed ? (int)inumber : (int)unumber; }
};

struct JParser {
  const char* p;
  const char* end;
  bool        ok = true;
  void skip() {
    while (p < end && (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r')) p++;
  }
  bool fail() { return ok = false; }
  static void append_utf8(std::string& out, unsigned cp) {
    if (cp < 0x80) out += (char)cp;
    else if (cp < 0x800) out += (char)(0xC0 | (cp >> 6)), out += (char)(0x80 | (cp & 0x3F));
    else if (cp < 0x10000) out += (char)(0xE0 | (cp >> 12)), out += (char)(0x80 | ((cp >> 6) & 0x3F)), out += (char)(0x80 | (cp & 0x3F));
    else
      out += (char)(0xF0 | (cp >> 18)), out += (char)(0x80 | ((cp >> 12) & 0x3F)), out += (char)(0x80 | ((cp >> 6) & 0x3F)),
          out += (char)(0x80 | (cp & 0x3F));
  }
  bool hex4(unsigned& cp) {  // p at the 'u'
    if (end - p < 5) return false;
    cp = 0;
    for (int k = 1; k <= 4; k++) {
      const char c = p[k];
      cp <<= 4;
      if (c >= '0' && c <= '9') cp |= (unsigned)(c - '0');
      else if (c >= 'a' && c <= 'f') cp |= (unsigned)(c - 'a' + 10);
      else if (c >= 'A' && c <= 'F') cp |= (unsigned)(c - 'A' + 10);
      else return false;
    }
    p += 4;
    return true;
  }
  bool parse_string(std::string& out) {
    if (p >= end || *p != '"') return fail();
    p++;
    while (p < end && *p != '"') {
      const unsigned char c = (unsigned char)*p;
      if (c == '\\') {
        if (++p >= end) return fail();
        switch (*p) {
          case '"': out += '"'; break;
          case '\\': out += '\\'; break;
          case '/': out += '/'; break;
          case 'b': out += '\b'; break;
          case 'f': out += '\f'; break;
          case 'n': out += '\n'; break;
          case 'r': out += '\r'; break;
          case 't': out += '\t'; break;
          case 'u': {
            unsigned cp = 0;
            if (!hex4(cp)) return fail();
            if (cp >= 0xD800 && cp <= 0xDBFF) {  // a high surrogate needs its low half right behind
              unsigned low = 0;
              if (end - p < 3 || p[1] != '\\' || p[2] != 'u') return fail();
              p += 2;
              if (!hex4(low) || low < 0xDC00 || low > 0xDFFF) return fail();
              cp = 0x10000 + ((cp - 0xD800) << 10) + (low - 0xDC00);
            } else if (cp >= 0xDC00 && cp <= 0xDFFF) {
              return fail();
            }
            append_utf8(out, cp);
          } break;
          default: return fail();
        }
        p++;
      } else if (c < 0x20) {
        return fail();
      } else if (c < 0x80) {
        out += *p++;
      } else {  // one well-formed UTF-8 sequence (Unicode table 3-7)
        int      more = 0;
        unsigned lo = 0x80, hi = 0xBF;
        if (c >= 0xC2 && c <= 0xDF) more = 1;
        else if (c == 0xE0) more = 2, lo = 0xA0;
        else if ((c >= 0xE1 && c <= 0xEC) || c == 0xEE || c == 0xEF) more = 2;
        else if (c == 0xED) more = 2, hi = 0x9F;
        else if (c == 0xF0) more = 3, lo = 0x90;
        else if (c >= 0xF1 && c <= 0xF3) more = 3;
        else if (c == 0xF4) more = 3, hi = 0x8F;
        else return fail();
        if (end - p <= more) return fail();
        for (int k = 1; k <= more; k++) {
          const unsigned char t = (unsigned char)p[k];
          if (t < lo || t > hi) return fail();
          lo = 0x80, hi = 0xBF;
        }
        out.append(p, (size_t)more + 1);
        p += more + 1;
      }
    }
    if (p >= end) return fail();
    p++;
    return true;
  }
  bool parse_number(JValue& v) {
    const char* start = p;
    const char* q     = p;
    auto digit = [&](const char* c) { return c < end && *c >= '0' && *c <= '9'; };
    if (q < end && *q == '-') q++;
    if (!digit(q)) return fail();
    if (*q == '0') q++;
    else
      while (digit(q)) q++;
    bool integer = true;
    if (q < end && *q == '.') {
      integer = false;
      if (!digit(++q)) return fail();
      while (digit(q)) q++;
    }
    if (q < end && (*q == 'e' || *q == 'E')) {
      integer = false;
      q++;
      if (q < end && (*q == '+' || *q == '-')) q++;
      if (!digit(q)) return fail();
      while (digit(q)) q++;
    }
    const std::string token(start, q);
    v.type = JValue::Number, v.kind = JValue::Float;
    p      = q;
    if (integer) {
      errno = 0;
      char* stop = nullptr;
      if (token[0] == '-') {
        const long long x = strtoll(token.c_str(), &stop, 10);
        if (errno == 0) return v.kind = JValue::Signed, v.inumber = x, true;
      } else {
        const unsigned long long x = strtoull(token.c_str(), &stop, 10);
        if (errno == 0) return v.kind = JValue::Unsigned, v.unumber = x, true;
      }
    }
    v.number = strtod(token.c_str(), nullptr);
    if (!std::isfinite(v.number)) return fail();  // "number overflow parsing"
    return true;
  }
  bool parse_document(JValue& v) {
    if (end - p >= 3 && (unsigned char)p[0] == 0xEF && (unsigned char)p[1] == 0xBB && (unsigned char)p[2] == 0xBF) p += 3;
    if (!parse(v)) return false;
    skip();
    return p == end ? true : fail();
  }
  bool parse(JValue& v, int depth = 0) {
    if (depth > 256) return fail();
    skip();
    if (p >= end) return fail();
    if (*p == '{') {
      v.type = JValue::Object;
      p++;
      skip();
      if (p < end && *p == '}') return p++, true;
      std::unordered_map<std::string, size_t> index;  // only kept up once the object is large
      while (true) {
        skip();
        std::string key;
        if (!parse_string(key)) return false;
        skip();
        if (p >= end || *p != ':') return fail();
        p++;
        size_t slot = v.object.size();
        if (v.object.size() < 32) {
          for (size_t k = 0; k < v.object.size(); k++)
            if (v.object[k].first == key) slot = k;
        } else {
          if (index.empty())
            for (size_t k = 0; k < v.object.size(); k++) index.emplace(v.object[k].first, k);
          auto it = index.find(key);
          if (it != index.end()) slot = it->second;
          else index.emplace(key, slot);
        }
        if (slot == v.object.size()) v.object.emplace_back(std::move(key), JValue{});
        else v.object[slot].second = JValue{};
        {
          JValue value;  // parsed aside: the vector may not grow under a reference into it
          if (!parse(value, depth + 1)) return false;
          v.object[slot].second = std::move(value);
        }
        skip();
        if (p < end && *p == ',') {
          p++;
          continue;
        }
        if (p < end && *p == '}') return p++, true;
        return fail();
      }
    }
    if (*p == '[') {
      v.type = JValue::Array;
      p++;
      skip();
      if (p < end && *p == ']') return p++, true;
      while (true) {
        v.array.emplace_back();
        if (!parse(v.array.back(), depth + 1)) return false;
        skip();
        if (p < end && *p == ',') {
          p++;
          continue;
        }
        if (p < end && *p == ']') return p++, true;
        return fail();
      }
    }
    if (*p == '"') {
      v.type = JValue::String;
      return parse_string(v.string);
    }
    if (end - p >= 4 && !strncmp(p, "true", 4)) return v.type = JValue::Bool, v.boolean = true, p += 4, true;
    if (end - p >= 5 && !strncmp(p, "false", 5)) return v.type = JValue::Bool, v.boolean = false, p += 5, true;
    if (end - p >= 4 && !strncmp(p, "null", 4)) return v.type = JValue::Null, p += 4, true;
    return parse_number(v);
  }
};

// get_opt of the reference: a present key of the wrong type is a parse error, a missing key keeps the default
struct JReader {
  bool ok = true;
  // (nlohmann converts a JSON boolean to any arithmetic type but double / int64: `"lens": true` reads as 1)
  void get(const JValue& e, const char* key, float& value) {
    if (auto v = e.find(key)) {
      if (v->type == JValue::Bool) value = v->boolean ? 1.0f : 0.0f;
      else if (v->type != JValue::Number) ok = false;
      else value = v->as_float();
    }
  }
  void get(const JValue& e, const char* key, int& value) {
    if (auto v = e.find(key)) {
      if (v->type == JValue::Bool) value = v->boolean ? 1 : 0;
      else if (v->type != JValue::Number) ok = false;
      else value = v->as_int();
    }
  }
  void get_bool(const JValue& e, const char* key, int& value) {
    if (auto v = e.find(key)) {
      if (v->type != JValue::Bool) ok = false;
      else value = v->boolean ? 1 : 0;
    }
  }
  void get(const JValue& e, const char* key, std::string& value) {
    if (auto v = e.find(key)) {
      if (v->type != JValue::String) ok = false;
      else value = v->string;
    }
  }
  void get_floats(const JValue& e, const char* key, float* value, size_t n) {
    if (auto v = e.find(key)) {
      // std::array from_json reads j.at(i) for i < n: a longer array passes, a shorter one throws
      if (v->type != JValue::Array || v->array.size() < n) return void(ok = false);
      for (size_t i = 0; i < n; i++) {
        if (v->array[i].type == JValue::Bool) value[i] = v->array[i].boolean ? 1.0f : 0.0f;
        else if (v->array[i].type != JValue::Number) return void(ok = false);
        else value[i] = v->array[i].as_float();
      }
    }
  }
};

// ---------------------------------------------------------------------------------------------------------------
// small vector algebra with the reference's operation order (yocto_math.h): only for lookat / missing camera
// ---------------------------------------------------------------------------------------------------------------
struct v3 {
  float x, y, z;
};
v3    operator-(const v3& a, const v3& b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
v3    operator+(const v3& a, const v3& b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
v3    operator*(const v3& a, float b) { return {a.x * b, a.y * b, a.z * b}; }
v3    operator/(const v3& a, float b) { return {a.x / b, a.y / b, a.z / b}; }
v3    operator-(const v3& a) { return {-a.x, -a.y, -a.z}; }
float dot(const v3& a, const v3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
v3    cross(const v3& a, const v3& b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
float length(const v3& a) { return std::sqrt(dot(a, a)); }
v3    normalize(const v3& a) {
  auto l = length(a);
  return (l != 0) ? a / l : a;
}
// lookat_frame, yocto_math.h:2348-2358
void lookat_frame(float* frame, const v3& eye, const v3& center, const v3& up, bool inv_xz) {
  auto w = normalize(eye - center);
  auto u = normalize(cross(up, w));
  auto v = normalize(cross(w, u));
  if (inv_xz) w = -w, u = -u;
  float out[12] = {u.x, u.y, u.z, v.x, v.y, v.z, w.x, w.y, w.z, eye.x, eye.y, eye.z};
  memcpy(frame, out, sizeof(out));
}

// ---------------------------------------------------------------------------------------------------------------
// PLY (load_ply, yocto_modelio.cpp; element / property model of yocto_modelio.h:66-107)
// ---------------------------------------------------------------------------------------------------------------
enum PlyType { I8, I16, I32, I64, U8, U16, U32, U64, F32, F64, PlyBad };
int ply_size(PlyType t) {
  static const int size[] = {1, 2, 4, 8, 1, 2, 4, 8, 4, 8, 0};
  return size[t];
}
PlyType ply_type(const std::string& s) {
  if (s == "char" || s == "int8") return I8;
  if (s == "short" || s == "int16") return I16;
  if (s == "int" || s == "int32") return I32;
  if (s == "long" || s == "int64") return I64;
  if (s == "uchar" || s == "uint8") return U8;
  if (s == "ushort" || s == "uint16") return U16;
  if (s == "uint" || s == "uint32") return U32;
  if (s == "ulong" || s == "uint64") return U64;
  if (s == "float" || s == "float32") return F32;
  if (s == "double" || s == "float64") return F64;
  return PlyBad;
}
struct PlyProperty {
  std::string          name;
  bool                 is_list = false;
  PlyType              type = F32, count_type = U8;
  std::vector<double>  values;  // every i8..i32 / u8..u32 / f32 / f64 value is exact in a double
  std::vector<uint8_t> sizes;   // list sizes, kept as uint8 like the reference's ldata_u8
};
struct PlyElement {
  std::string              name;
  size_t                   count = 0;
  std::vector<PlyProperty> properties;
};
double ply_read_binary(const uint8_t*& p, PlyType t, bool big_endian) {
  uint8_t   b[8];
  const int n = ply_size(t);
  if (big_endian)
    for (int i = 0; i < n; i++) b[i] = p[n - 1 - i];
  else memcpy(b, p, n);
  p += n;
  switch (t) {
    case I8: return (double)(int8_t)b[0];
    case U8: return (double)b[0];
    case I16: { int16_t v; memcpy(&v, b, 2); return v; }
    case U16: { uint16_t v; memcpy(&v, b, 2); return v; }
    case I32: { int32_t v; memcpy(&v, b, 4); return v; }
    case U32: { uint32_t v; memcpy(&v, b, 4); return v; }
    case I64: { int64_t v; memcpy(&v, b, 8); return (double)v; }
    case U64: { uint64_t v; memcpy(&v, b, 8); return (double)v; }
    case F32: { float v; memcpy(&v, b, 4); return v; }
    case F64: { double v; memcpy(&v, b, 8); return v; }
    default: return 0;
  }
}
bool load_ply(const std::string& filename, std::vector<PlyElement>& elements, std::string& error) {
  std::vector<uint8_t> data;
  if (!read_file(filename, data, error)) return false;
  auto parse_error = [&]() { return error = "cannot parse " + filename, false; };
  data.push_back(0);
  const char* p   = (const char*)data.data();
  const char* end = p + data.size() - 1;
  auto next_line = [&](std::string& line) {
    if (p >= end) return false;
    const char* e = (const char*)memchr(p, '\n', end - p);
    if (!e) e = end;
    line.assign(p, e);
    while (!line.empty() && (line.back() == '\r' || line.back() == ' ')) line.pop_back();
    p = e < end ? e + 1 : end;
    return true;
  };
  auto split = [](const std::string& line) {
    std::vector<std::string> tokens;
    size_t                   i = 0;
    while (i < line.size()) {
      while (i < line.size() && isspace((unsigned char)line[i])) i++;
      size_t j = i;
      while (j < line.size() && !isspace((unsigned char)line[j])) j++;
      if (j > i) tokens.push_back(line.substr(i, j - i));
      i = j;
    }
    return tokens;
  };
  std::string line;
  int  format = -1;  // 0 ascii, 1 little endian, 2 big endian
  bool header_done = false, first_line = true;
  while (next_line(line)) {
    if (auto hash = line.find('#'); hash != std::string::npos) line.resize(hash);  // remove_comment, yocto_modelio.cpp:336
    auto tok = split(line);
    if (tok.empty()) continue;
    if (first_line) {
      if (tok[0] != "ply") return parse_error();
      first_line = false;
    } else if (tok[0] == "ply") {
    } else if (tok[0] == "format") {
      if (tok.size() < 2) return parse_error();
      format = tok[1] == "ascii" ? 0 : tok[1] == "binary_little_endian" ? 1 : tok[1] == "binary_big_endian" ? 2 : -1;
      if (format < 0) return parse_error();
    } else if (tok[0] == "comment" || tok[0] == "obj_info") {
    } else if (tok[0] == "element") {
      if (tok.size() < 3) return parse_error();
      size_t count = 0;
      if (std::from_chars(tok[2].data(), tok[2].data() + tok[2].size(), count).ptr == tok[2].data()) return parse_error();
      elements.emplace_back();
      elements.back().name  = tok[1];
      elements.back().count = count;
    } else if (tok[0] == "property") {
      if (elements.empty() || tok.size() < 3) return parse_error();
      PlyProperty prop;
      if (tok[1] == "list") {
        if (tok.size() < 5) return parse_error();
        prop.is_list = true, prop.count_type = ply_type(tok[2]), prop.type = ply_type(tok[3]), prop.name = tok[4];
        if (prop.count_type != U8) return parse_error();  // the reference keeps list sizes in bytes and refuses the rest (:570)
      } else {
        prop.type = ply_type(tok[1]), prop.name = tok[2];
      }
      if (prop.type == PlyBad) return parse_error();
      elements.back().properties.push_back(std::move(prop));
    } else if (tok[0] == "end_header") {
      header_done = true;
      break;
    } else {
      return parse_error();
    }
  }
  if (!header_done || format < 0) return parse_error();
  for (auto& elem : elements)
    for (auto& prop : elem.properties) {
      if (prop.is_list) prop.sizes.reserve(elem.count), prop.values.reserve(elem.count * 3);
      else prop.values.reserve(elem.count);
    }
  if (format == 0) {
    // one line per element row, every value read with the parser of its declared type (parse_value, yocto_modelio.cpp:
    // 405-418): from_chars - integers stop at the first character that is not a digit, a value that does not fit leaves
    // the zero it started from, floats are rounded once to their own width
    const char* line_end = p;
    auto read_row = [&]() {
      if (p >= end) return false;
      const char* e = (const char*)memchr(p, '\n', end - p);
      line_end      = e ? e + 1 : end;
      return true;
    };
    auto skip_space = [&]() {
      while (p < line_end && (*p == ' ' || *p == '\t' || *p == '\r' || *p == '\n')) p++;
    };
    auto parse_as = [&](PlyType t, double& v) {
      skip_space();
      const char* stop = p;
      auto integer = [&](auto x) {
        stop = std::from_chars(p, line_end, x).ptr;
        v    = (double)x;
      };
      switch (t) {
        case I8: integer((int8_t)0); break;
        case I16: integer((int16_t)0); break;
        case I32: integer((int32_t)0); break;
        case I64: integer((int64_t)0); break;
        case U8: integer((uint8_t)0); break;
        case U16: integer((uint16_t)0); break;
        case U32: integer((uint32_t)0); break;
        case U64: integer((uint64_t)0); break;
        case F32: {
          float x = 0;
          stop    = std::from_chars(p, line_end, x).ptr;
          v       = x;
        } break;
        default: {
          double x = 0;
          stop     = std::from_chars(p, line_end, x).ptr;
          v        = x;
        } break;
      }
      if (stop == p) return false;
      p = stop;
      return true;
    };
    for (auto& elem : elements)
      for (size_t row = 0; row < elem.count; row++) {
        if (!read_row()) return error = "cannot read " + filename, false;
        for (auto& prop : elem.properties) {
          double v;
          if (!prop.is_list) {
            if (!parse_as(prop.type, v)) return parse_error();
            prop.values.push_back(v);
          } else {
            if (!parse_as(U8, v)) return parse_error();
            const uint8_t n = (uint8_t)v;
            prop.sizes.push_back(n);
            for (int k = 0; k < n; k++) {
              if (!parse_as(prop.type, v)) return parse_error();
              prop.values.push_back(v);
            }
          }
        }
        p = line_end;
      }
  } else {
    const uint8_t* q    = (const uint8_t*)p;
    const uint8_t* qend = (const uint8_t*)end;
    const bool     be   = format == 2;
    // a column of `count` values of one type, `stride` bytes apart, converted with the type decided outside the loop
    auto read_column = [&](const uint8_t* src, size_t stride, size_t count, PlyType t, double* dst) {
      auto run = [&](auto sample) {
        using T = decltype(sample);
        for (size_t i = 0; i < count; i++, src += stride) {
          T v;
          if (be && sizeof(T) > 1) {
            uint8_t b[sizeof(T)];
            for (size_t k = 0; k < sizeof(T); k++) b[k] = src[sizeof(T) - 1 - k];
            memcpy(&v, b, sizeof(T));
          } else {
            memcpy(&v, src, sizeof(T));
          }
          dst[i] = (double)v;
        }
      };
      switch (t) {
        case I8: run(int8_t{}); break;
        case U8: run(uint8_t{}); break;
        case I16: run(int16_t{}); break;
        case U16: run(uint16_t{}); break;
        case I32: run(int32_t{}); break;
        case U32: run(uint32_t{}); break;
        case I64: run(int64_t{}); break;
        case U64: run(uint64_t{}); break;
        case F32: run(float{}); break;
        default: run(double{}); break;
      }
    };
    for (auto& elem : elements) {
      bool has_list = false;
      size_t row_bytes = 0;
      for (auto& prop : elem.properties) has_list |= prop.is_list, row_bytes += ply_size(prop.type);
      if (!has_list) {  // fixed-size rows: one strided pass per property
        if (row_bytes && (size_t)(qend - q) / row_bytes < elem.count) return parse_error();
        size_t offset = 0;
        for (auto& prop : elem.properties) {
          prop.values.resize(elem.count);
          read_column(q + offset, row_bytes, elem.count, prop.type, prop.values.data());
          offset += ply_size(prop.type);
        }
        q += row_bytes * elem.count;
        continue;
      }
      for (size_t row = 0; row < elem.count; row++)
        for (auto& prop : elem.properties) {
          if (!prop.is_list) {
            if (qend - q < ply_size(prop.type)) return parse_error();
            prop.values.push_back(ply_read_binary(q, prop.type, be));
          } else {
            if (qend - q < 1) return parse_error();
            const uint8_t n = *q++;  // list sizes are bytes (checked in the header)
            prop.sizes.push_back(n);
            const size_t item = ply_size(prop.type);
            if ((size_t)(qend - q) < n * item) return parse_error();
            const size_t at = prop.values.size();
            prop.values.resize(at + n);
            read_column(q, item, n, prop.type, prop.values.data() + at);
            q += n * item;
          }
        }
    }
  }
  return true;
}
const PlyProperty* ply_find(const std::vector<PlyElement>& elements, const char* element, const char* property) {
  for (auto& elem : elements) {
    if (elem.name != element) continue;
    for (auto& prop : elem.properties)
      if (prop.name == property) return &prop;
  }
  return nullptr;
}

struct HostShape {
  std::vector<int32_t> points, lines, triangles, quads;
  std::vector<float>   positions, normals, texcoords, colors, radius;
};
// get_values(ply, "vertex", {...}), yocto_modelio.h:548-566: all properties present and scalar, else nothing
template <size_t N>
void ply_vertex_values(const std::vector<PlyElement>& ply, const char* const (&names)[N], std::vector<float>& out) {
  out.clear();
  const PlyProperty* props[N];
  for (size_t k = 0; k < N; k++) {
    props[k] = ply_find(ply, "vertex", names[k]);
    if (!props[k] || props[k]->is_list) return;
  }
  const size_t n = props[0]->values.size();
  out.resize(n * N);
  for (size_t k = 0; k < N; k++)
    for (size_t i = 0; i < n && i < props[k]->values.size(); i++) out[i * N + k] = (float)props[k]->values[i];
}
// load_shape for .ply, yocto_sceneio.cpp:1018-1035 (flip_texcoord = true as load_json_scene passes)
bool load_obj_shape(const std::string& filename, HostShape& shape, std::string& error);  // .obj, further down
// load_shape for .stl (yocto_sceneio.cpp:1053-1062 over load_stl, yocto_modelio.cpp:2164-2331): the binary format only
// - 80 bytes of header, then per solid a count and 50 bytes per triangle; the file is binary unless it starts with
// "solid" and its length disagrees with its first count; exactly one solid. (The reference's ascii branch fails on
// every "outer loop" line, :2303, so ascii files are refused there as here.) Vertices are merged when their three
// floats compare equal (-0 == +0, a NaN equals nothing), numbered in order of first use.
struct StlVertexHash {
  size_t operator()(const std::array<float, 3>& v) const {
    size_t h = 0;
    for (float x : v) {
      uint32_t bits = 0;
      if (x != 0) memcpy(&bits, &x, 4);  // both zeros hash alike
      h ^= std::hash<uint32_t>()(bits) + 0x9e3779b9 + (h << 6) + (h >> 2);
    }
    return h;
  }
};
bool load_stl_shape(const std::string& filename, HostShape& shape, std::string& error) {
  std::vector<uint8_t> data;
  if (!read_file(filename, data, error)) return false;
  auto read_error = [&]() { return error = "cannot read " + filename, false; };
  if (data.size() < 80) return read_error();
  bool binary = memcmp(data.data(), "solid", 5) != 0;
  if (!binary) {
    if (data.size() < 84) return read_error();
    uint32_t count;
    memcpy(&count, &data[80], 4);
    binary = data.size() == 80 + 4 + (size_t)50 * count;
  }
  if (!binary) return error = "cannot parse " + filename, false;
  size_t pos = 80, solids = 0;
  std::vector<std::array<float, 3>> corners;
  while (pos < data.size()) {
    if (data.size() - pos < 4) return read_error();
    uint32_t count;
    memcpy(&count, &data[pos], 4);
    pos += 4;
    if ((data.size() - pos) / 50 < count) return read_error();
    if (solids++ == 0) {
      corners.resize((size_t)count * 3);
      for (size_t t = 0; t < count; t++) memcpy(corners[3 * t].data(), &data[pos + 50 * t + 12], 36);
    }
    pos += (size_t)50 * count;
  }
  if (solids == 0) return read_error();
  if (solids != 1) return error = "empty shape " + filename, false;  // "shape_error": one solid per file
  std::unordered_map<std::array<float, 3>, int, StlVertexHash> index;
  shape.triangles.reserve(corners.size());
  for (auto& corner : corners) {
    auto it = index.find(corner);
    if (it == index.end()) {
      it = index.insert({corner, (int)(shape.positions.size() / 3)}).first;
      shape.positions.insert(shape.positions.end(), corner.begin(), corner.end());
    }
    shape.triangles.push_back(it->second);
  }
  return true;
}
bool load_shape(const std::string& filename, HostShape& shape, std::string& error) {
  const auto ext = path_extension(filename);
  if (ext == ".obj" || ext == ".OBJ") return load_obj_shape(filename, shape, error);
  if (ext == ".stl") return load_stl_shape(filename, shape, error);
  if (ext != ".ply") return error = "unsupported format " + filename, false;
  std::vector<PlyElement> ply;
  if (!load_ply(filename, ply, error)) return false;
  ply_vertex_values(ply, {"x", "y", "z"}, shape.positions);
  ply_vertex_values(ply, {"nx", "ny", "nz"}, shape.normals);
  if (ply_find(ply, "vertex", "u")) ply_vertex_values(ply, {"u", "v"}, shape.texcoords);
  else ply_vertex_values(ply, {"s", "t"}, shape.texcoords);
  for (size_t i = 0; i + 1 < shape.texcoords.size(); i += 2) shape.texcoords[i + 1] = 1 - shape.texcoords[i + 1];
  if (ply_find(ply, "vertex", "alpha")) {
    ply_vertex_values(ply, {"red", "green", "blue", "alpha"}, shape.colors);
  } else {
    std::vector<float> rgb;
    ply_vertex_values(ply, {"red", "green", "blue"}, rgb);
    shape.colors.resize(rgb.size() / 3 * 4);
    for (size_t i = 0; i < rgb.size() / 3; i++)
      shape.colors[4 * i] = rgb[3 * i], shape.colors[4 * i + 1] = rgb[3 * i + 1], shape.colors[4 * i + 2] = rgb[3 * i + 2],
                     shape.colors[4 * i + 3] = 1;
  }
  if (auto r = ply_find(ply, "vertex", "radius"); r && !r->is_list) {
    shape.radius.resize(r->values.size());
    for (size_t i = 0; i < r->values.size(); i++) shape.radius[i] = (float)r->values[i];
  }
  // faces: quads as soon as one face has four corners, else triangles (get_faces, yocto_modelio.h:700-708);
  // polygons become fans, short faces are padded with -1 exactly like get_triangles / get_quads (:618-688)
  if (auto f = ply_find(ply, "face", "vertex_indices"); f && f->is_list) {
    bool has_quads = false;
    for (auto n : f->sizes) has_quads |= n == 4;
    auto   at      = [&](size_t i) { return (int32_t)f->values[i]; };
    size_t current = 0;
    for (auto n : f->sizes) {
      if (has_quads) {
        auto& q = shape.quads;
        if (n == 0) q.insert(q.end(), {-1, -1, -1, -1});
        else if (n == 1) q.insert(q.end(), {at(current), -1, -1, -1});
        else if (n == 2) q.insert(q.end(), {at(current), at(current + 1), -1, -1});
        else if (n == 3) q.insert(q.end(), {at(current), at(current + 1), at(current + 2), at(current + 2)});
        else if (n == 4) q.insert(q.end(), {at(current), at(current + 1), at(current + 2), at(current + 3)});
        else
          for (size_t item = 2; item < n; item++)
            q.insert(q.end(), {at(current), at(current + item - 1), at(current + item), at(current + item)});
      } else {
        auto& t = shape.triangles;
        if (n == 0) t.insert(t.end(), {-1, -1, -1});
        else if (n == 1) t.insert(t.end(), {at(current), -1, -1});
        else if (n == 2) t.insert(t.end(), {at(current), at(current + 1), -1});
        else if (n == 3) t.insert(t.end(), {at(current), at(current + 1), at(current + 2)});
        else
          for (size_t item = 2; item < n; item++) t.insert(t.end(), {at(current), at(current + item - 1), at(current + item)});
      }
      current += n;
    }
  }
  if (auto l = ply_find(ply, "line", "vertex_indices"); l && l->is_list) {  // polylines -> segments, :710-737
    auto   at      = [&](size_t i) { return (int32_t)l->values[i]; };
    size_t current = 0;
    for (auto n : l->sizes) {
      auto& s = shape.lines;
      if (n == 0) s.insert(s.end(), {-1, -1});
      else if (n == 1) s.insert(s.end(), {at(current), -1});
      else if (n == 2) s.insert(s.end(), {at(current), at(current + 1)});
      else
        for (size_t item = 1; item < n; item++) s.insert(s.end(), {at(current + item - 1), at(current + item)});
      current += n;
    }
  }
  if (auto pt = ply_find(ply, "point", "vertex_indices"); pt && pt->is_list) {  // get_list_values, :605-616
    shape.points.resize(pt->values.size());
    for (size_t i = 0; i < pt->values.size(); i++) shape.points[i] = (int32_t)pt->values[i];
  }
  if (shape.points.empty() && shape.lines.empty() && shape.triangles.empty() && shape.quads.empty())
    return error = "empty shape " + filename, false;
  return true;
}

// textures: ygl_imageio.cpp (PNG, JPEG, Radiance HDR, OpenEXR) through load_texture of ygl_hostio.h


// ---------------------------------------------------------------------------------------------------------------
// the loaded scene
// ---------------------------------------------------------------------------------------------------------------
int material_type_from_name(const std::string& name) {
  static const char* names[] = {"matte", "glossy", "reflective", "transparent", "refractive", "subsurface", "volumetric", "gltfpbr"};
  for (int i = 0; i < 8; i++)
    if (name == names[i]) return i;
  return 0;  // NLOHMANN_JSON_SERIALIZE_ENUM maps an unknown label to the first entry
}

const float kIdentityFrame[12] = {1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0};

// load resources on the host cores; the first error (lowest index) wins, like the reference's parallel_for
// (`order`, if given, is the sequence in which items are handed out - largest first keeps the cores even)
bool parallel_load(size_t n, std::string& error, const std::function<bool(size_t, std::string&)>& fn,
    const std::vector<size_t>* order = nullptr) {
  std::vector<std::string> errors(n);
  std::vector<char>        failed(n, 0);
  std::atomic<size_t>      next{0};
  const unsigned           nthreads = std::max(1u, std::min((unsigned)n, (unsigned)host_parallelism()));
  std::vector<std::thread> pool;
  for (unsigned t = 0; t < nthreads; t++)
    pool.emplace_back([&]() {
      for (size_t k = next++; k < n; k = next++) {
        const size_t i = order ? (*order)[k] : k;
        try {  // nothing may leave a worker thread (or, further up, the C ABI) as an exception
          if (!fn(i, errors[i])) failed[i] = 1;
        } catch (const std::exception& e) {
          errors[i] = std::string("out of memory or internal error (") + e.what() + ")", failed[i] = 1;
        }
      }
    });
  for (auto& t : pool) t.join();
  for (size_t i = 0; i < n; i++)
    if (failed[i]) return error = errors[i], false;
  return true;
}


// ---------------------------------------------------------------------------------------------------------------
// subdivs: load_subdiv (.obj, yocto_sceneio.cpp:1188-1232, :2915-2926; load_obj yocto_modelio.cpp:1393-1492) and
// tesselate_subdiv (yocto_scene.cpp:739-805) with subdivide_catmullclark / subdivide_quads (yocto_shape.cpp:2779-3064),
// quads_normals (:1495-1511), split_facevarying (:2567-2618) and the host-side eval_texture (yocto_scene.cpp:111-160).
// ---------------------------------------------------------------------------------------------------------------
struct v2 {
  float x, y;
};
v2 operator+(const v2& a, const v2& b) { return {a.x + b.x, a.y + b.y}; }
v2 operator-(const v2& a, const v2& b) { return {a.x - b.x, a.y - b.y}; }
v2 operator*(const v2& a, float b) { return {a.x * b, a.y * b}; }
v2 operator/(const v2& a, float b) { return {a.x / b, a.y / b}; }
struct i2 {
  int  x, y;
  bool operator==(const i2& o) const { return x == o.x && y == o.y; }
};
struct i3 {
  int  x, y, z;
  bool operator==(const i3& o) const { return x == o.x && y == o.y && z == o.z; }
};
struct i4 {
  int x, y, z, w;
};
// std::hash<vec2i> / <vec3i> of yocto_shape.h:376-395: with the same hash, the same insertion sequence and the same
// standard library the maps below iterate in the reference's order (get_boundary walks the edge map)
struct i2_hash {
  size_t operator()(const i2& v) const {
    const std::hash<int> hasher;
    size_t               h = 0;
    h ^= hasher(v.x) + 0x9e3779b9 + (h << 6) + (h >> 2);
    h ^= hasher(v.y) + 0x9e3779b9 + (h << 6) + (h >> 2);
    return h;
  }
};
struct i3_hash {
  size_t operator()(const i3& v) const {
    const std::hash<int> hasher;
    size_t               h = 0;
    h ^= hasher(v.x) + 0x9e3779b9 + (h << 6) + (h >> 2);
    h ^= hasher(v.y) + 0x9e3779b9 + (h << 6) + (h >> 2);
    h ^= hasher(v.z) + 0x9e3779b9 + (h << 6) + (h >> 2);
    return h;
  }
};
struct EdgeData {
  int index, nfaces;
};
using EdgeMap = std::unordered_map<i2, EdgeData, i2_hash>;
void insert_edge(EdgeMap& emap, const i2& edge) {
  auto es = edge.x < edge.y ? edge : i2{edge.y, edge.x};
  auto it = emap.find(es);
  if (it == emap.end()) emap.insert(it, {es, EdgeData{(int)emap.size(), 1}});
  else it->second.nfaces += 1;
}
EdgeMap make_edge_map(const std::vector<i4>& quads) {
  EdgeMap emap;
  for (auto& q : quads) {
    insert_edge(emap, {q.x, q.y});
    insert_edge(emap, {q.y, q.z});
    if (q.z != q.w) insert_edge(emap, {q.z, q.w});
    insert_edge(emap, {q.w, q.x});
  }
  return emap;
}
int edge_index(const EdgeMap& emap, const i2& edge) {
  auto es = edge.x < edge.y ? edge : i2{edge.y, edge.x};
  auto it = emap.find(es);
  return it == emap.end() ? -1 : it->second.index;
}
// the split shared by subdivide_quads and subdivide_catmullclark: vertices, then edge midpoints, then face centres
template <class T>
void split_quads(const std::vector<i4>& quads, const std::vector<T>& vertices, const EdgeMap& emap, std::vector<i4>& tquads,
    std::vector<T>& tvertices) {
  std::vector<i2> edges(emap.size());
  for (auto& [edge, data] : emap) edges[data.index] = edge;
  tvertices.clear();
  tvertices.reserve(vertices.size() + edges.size() + quads.size());
  for (auto& vertex : vertices) tvertices.push_back(vertex);
  for (auto& edge : edges) tvertices.push_back((vertices[edge.x] + vertices[edge.y]) / 2);
  for (auto& quad : quads) {
    if (quad.z != quad.w) tvertices.push_back((vertices[quad.x] + vertices[quad.y] + vertices[quad.z] + vertices[quad.w]) / 4);
    else tvertices.push_back((vertices[quad.x] + vertices[quad.y] + vertices[quad.z]) / 3);
  }
  const int nverts = (int)vertices.size(), nedges = (int)edges.size();
  auto edge_vertex = [&](int a, int b) { return nverts + edge_index(emap, {a, b}); };
  tquads.clear();
  tquads.reserve(quads.size() * 4);
  for (size_t quad_id = 0; quad_id < quads.size(); quad_id++) {
    const i4& quad   = quads[quad_id];
    const int centre = nverts + nedges + (int)quad_id;
    if (quad.z != quad.w) {
      tquads.push_back({quad.x, edge_vertex(quad.x, quad.y), centre, edge_vertex(quad.w, quad.x)});
      tquads.push_back({quad.y, edge_vertex(quad.y, quad.z), centre, edge_vertex(quad.x, quad.y)});
      tquads.push_back({quad.z, edge_vertex(quad.z, quad.w), centre, edge_vertex(quad.y, quad.z)});
      tquads.push_back({quad.w, edge_vertex(quad.w, quad.x), centre, edge_vertex(quad.z, quad.w)});
    } else {
      tquads.push_back({quad.x, edge_vertex(quad.x, quad.y), centre, edge_vertex(quad.z, quad.x)});
      tquads.push_back({quad.y, edge_vertex(quad.y, quad.z), centre, edge_vertex(quad.x, quad.y)});
      tquads.push_back({quad.z, edge_vertex(quad.z, quad.x), centre, edge_vertex(quad.y, quad.z)});
    }
  }
}
template <class T>
void subdivide_quads(std::vector<i4>& quads, std::vector<T>& vertices) {
  if (quads.empty() || vertices.empty()) return;
  auto            emap = make_edge_map(quads);
  std::vector<i4> tquads;
  std::vector<T>  tvertices;
  split_quads(quads, vertices, emap, tquads, tvertices);
  quads.swap(tquads), vertices.swap(tvertices);
}
template <class T>
void subdivide_catmullclark(std::vector<i4>& quads, std::vector<T>& vertices, bool lock_boundary) {
  if (quads.empty() || vertices.empty()) return;
  auto            emap = make_edge_map(quads);
  std::vector<i2> boundary;
  for (auto& [edge, data] : emap)
    if (data.nfaces < 2) boundary.push_back(edge);
  std::vector<i4> tquads;
  std::vector<T>  tvertices;
  split_quads(quads, vertices, emap, tquads, tvertices);
  const int       nverts = (int)vertices.size();
  std::vector<i2> tboundary;
  tboundary.reserve(boundary.size() * 2);
  for (auto& edge : boundary) {
    const int mid = nverts + edge_index(emap, edge);
    tboundary.push_back({edge.x, mid});
    tboundary.push_back({mid, edge.y});
  }
  // creases
  // (locked boundary: its vertices are crease vertices; free boundary: its edges are crease edges)
  std::vector<int> tcrease_verts;
  if (lock_boundary)
    for (auto& b : tboundary) tcrease_verts.push_back(b.x), tcrease_verts.push_back(b.y);
  const std::vector<i2>  no_edges;
  const std::vector<i2>& tcrease_edges = lock_boundary ? no_edges : tboundary;
  // valence
  std::vector<int> tvert_val(tvertices.size(), 2);
  for (auto& edge : tboundary) tvert_val[edge.x] = tvert_val[edge.y] = lock_boundary ? 0 : 1;
  // averaging pass
  std::vector<T>   avert(tvertices.size(), T{});
  std::vector<int> acount(tvertices.size(), 0);
  for (auto point : tcrease_verts) {
    if (tvert_val[point] != 0) continue;
    avert[point] = avert[point] + tvertices[point];
    acount[point] += 1;
  }
  for (auto& edge : tcrease_edges) {
    auto centroid = (tvertices[edge.x] + tvertices[edge.y]) / 2;
    for (int vid : {edge.x, edge.y}) {
      if (tvert_val[vid] != 1) continue;
      avert[vid] = avert[vid] + centroid;
      acount[vid] += 1;
    }
  }
  for (auto& quad : tquads) {
    auto centroid = (tvertices[quad.x] + tvertices[quad.y] + tvertices[quad.z] + tvertices[quad.w]) / 4;
    for (int vid : {quad.x, quad.y, quad.z, quad.w}) {
      if (tvert_val[vid] != 2) continue;
      avert[vid] = avert[vid] + centroid;
      acount[vid] += 1;
    }
  }
  for (size_t i = 0; i < tvertices.size(); i++) avert[i] = avert[i] / (float)acount[i];
  // correction pass: p = p + (avg_p - p) * (4 / avg_count)
  for (size_t i = 0; i < tvertices.size(); i++) {
    if (tvert_val[i] != 2) continue;
    avert[i] = tvertices[i] + (avert[i] - tvertices[i]) * (4 / (float)acount[i]);
  }
  quads.swap(tquads), vertices.swap(avert);
}
// quads_normals, yocto_shape.cpp:1495-1511 (quad_normal / quad_area, yocto_geometry.h:516-532)
std::vector<v3> quads_normals(const std::vector<i4>& quads, const std::vector<v3>& positions) {
  std::vector<v3> normals(positions.size(), v3{0, 0, 0});
  auto triangle_normal = [](const v3& p0, const v3& p1, const v3& p2) { return normalize(cross(p1 - p0, p2 - p0)); };
  auto triangle_area   = [](const v3& p0, const v3& p1, const v3& p2) { return length(cross(p1 - p0, p2 - p0)) / 2; };
  for (auto& q : quads) {
    const v3 &p0 = positions[q.x], &p1 = positions[q.y], &p2 = positions[q.z], &p3 = positions[q.w];
    auto normal = normalize(triangle_normal(p0, p1, p3) + triangle_normal(p2, p3, p1));
    auto area   = triangle_area(p0, p1, p3) + triangle_area(p2, p3, p1);
    normals[q.x] = normals[q.x] + normal * area;
    normals[q.y] = normals[q.y] + normal * area;
    normals[q.z] = normals[q.z] + normal * area;
    if (q.z != q.w) normals[q.w] = normals[q.w] + normal * area;
  }
  for (auto& normal : normals) normal = normalize(normal);
  return normals;
}
// mean(eval_texture(texture, uv, false)), yocto_scene.cpp:111-163 + yocto_math.h:1521-1522, on the host
float texture_mean(const HostTexture& tex, const v2& uv) {
  if (tex.width == 0 || tex.height == 0) return 0;
  auto fmin = [](float a, float b) { return (a < b) ? a : b; };
  auto fmax = [](float a, float b) { return (a > b) ? a : b; };
  auto imin = [](int a, int b) { return (a < b) ? a : b; };
  auto imax = [](int a, int b) { return (a > b) ? a : b; };
  const int w = tex.width, h = tex.height;
  float     s = 0.0f, t = 0.0f;
  if (tex.clamp) {
    s = fmin(fmax(uv.x, 0.0f), 1.0f) * w;
    t = fmin(fmax(uv.y, 0.0f), 1.0f) * h;
  } else {
    s = std::fmod(uv.x, 1.0f) * w;
    if (s < 0) s += w;
    t = std::fmod(uv.y, 1.0f) * h;
    if (t < 0) t += h;
  }
  const int   i = imin(imax((int)s, 0), w - 1), j = imin(imax((int)t, 0), h - 1);
  const int   ii = (i + 1) % w, jj = (j + 1) % h;
  const float u = s - i, v = t - j;
  struct c4 {
    float x, y, z, w;
  };
  auto lookup = [&](int i, int j) {
    const size_t k = ((size_t)j * w + i) * 4;
    if (!tex.pixelsf.empty()) return c4{tex.pixelsf[k], tex.pixelsf[k + 1], tex.pixelsf[k + 2], tex.pixelsf[k + 3]};
    return c4{tex.pixelsb[k] / 255.0f, tex.pixelsb[k + 1] / 255.0f, tex.pixelsb[k + 2] / 255.0f, tex.pixelsb[k + 3] / 255.0f};
  };
  auto scale = [](const c4& a, float b) { return c4{a.x * b, a.y * b, a.z * b, a.w * b}; };
  auto add   = [](const c4& a, const c4& b) { return c4{a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w}; };
  c4 color;
  if (tex.nearest) {
    color = lookup(i, j);
  } else {
    color = add(add(add(scale(scale(lookup(i, j), 1 - u), 1 - v), scale(scale(lookup(i, jj), 1 - u), v)),
                    scale(scale(lookup(ii, j), u), 1 - v)),
        scale(scale(lookup(ii, jj), u), v));
  }
  return (color.x + color.y + color.z + color.w) / 4;
}

struct HostSubdiv {
  std::vector<i4> quadspos, quadsnorm, quadstexcoord;
  std::vector<v3> positions, normals;
  std::vector<v2> texcoords;
  int   subdivisions = 0, catmullclark = 1, smooth = 1;
  float displacement     = 0;
  int   displacement_tex = -1, shape = -1;
};
// The single-shape load_obj of yocto_modelio.cpp:1393-1458 up to (not including) its vertex conversion: raw v / vn / vt
// arrays, the vertices of every f / l / p element with 1-based (already resolved) indices, element sizes and types.
struct ObjVertex {
  int  position = 0, texcoord = 0, normal = 0;
  bool operator==(const ObjVertex& o) const { return position == o.position && texcoord == o.texcoord && normal == o.normal; }
};
struct ObjVertexHash {  // std::hash<obj_vertex>, yocto_modelio.h:399-408
  size_t operator()(const ObjVertex& v) const {
    const std::hash<int> hasher;
    size_t               h = 0;
    h ^= hasher(v.position) + 0x9e3779b9 + (h << 6) + (h >> 2);
    h ^= hasher(v.normal) + 0x9e3779b9 + (h << 6) + (h >> 2);
    h ^= hasher(v.texcoord) + 0x9e3779b9 + (h << 6) + (h >> 2);
    return h;
  }
};
struct ObjData {
  std::vector<v3>        positions, normals;
  std::vector<v2>        texcoords;
  std::vector<ObjVertex> vertices;
  std::vector<int>       sizes;  // per element
  std::vector<char>      types;  // per element: 'f', 'l' or 'p'
};
bool read_obj(const std::string& filename, ObjData& obj, std::string& error) {
  std::vector<uint8_t> data;
  if (!read_file(filename, data, error)) return false;
  auto parse_error = [&]() { return error = "cannot parse " + filename, false; };
  const char*       p   = (const char*)data.data();
  const char* const end = p + data.size();
  auto is_space = [](char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\n'; };
  // lines are parsed in place; numbers with from_chars like the reference's parse_value (yocto_modelio.cpp:405-418):
  // no leading '+', an integer that does not fit leaves the 0 it started from
  while (p < end) {
    const char* eol = (const char*)memchr(p, '\n', end - p);
    if (!eol) eol = end;
    const char* s    = p;
    const char* stop = (const char*)memchr(p, '#', eol - p);  // remove_comment
    if (!stop) stop = eol;
    p = eol < end ? eol + 1 : end;
    auto skip = [&]() { while (s < stop && is_space(*s)) s++; };
    auto number = [&](float& v) {
      skip();
      auto r = std::from_chars(s, stop, v);
      if (r.ptr == s) return false;
      s = r.ptr;
      return true;
    };
    auto integer = [&](int& v) {
      skip();
      auto r = std::from_chars(s, stop, v);
      if (r.ptr == s) return false;
      s = r.ptr;
      return true;
    };
    skip();
    if (s >= stop) continue;
    const char* cmd = s;
    while (s < stop && !is_space(*s)) s++;
    const size_t cmd_len = (size_t)(s - cmd);
    auto is_cmd = [&](const char* name) { return strlen(name) == cmd_len && !memcmp(cmd, name, cmd_len); };
    if (is_cmd("v")) {
      v3 v = {0, 0, 0};
      if (!number(v.x) || !number(v.y) || !number(v.z)) return parse_error();
      obj.positions.push_back(v);
    } else if (is_cmd("vn")) {
      v3 v = {0, 0, 0};
      if (!number(v.x) || !number(v.y) || !number(v.z)) return parse_error();
      obj.normals.push_back(v);
    } else if (is_cmd("vt")) {
      v2 v = {0, 0};
      if (!number(v.x) || !number(v.y)) return parse_error();
      obj.texcoords.push_back(v);
    } else if (is_cmd("f") || is_cmd("l") || is_cmd("p")) {
      int size = 0;
      skip();
      while (s < stop) {
        ObjVertex vert;
        if (!integer(vert.position)) return parse_error();
        if (s < stop && *s == '/') {
          s++;
          if (s < stop && *s == '/') {
            s++;
            if (!integer(vert.normal)) return parse_error();
          } else {
            if (!integer(vert.texcoord)) return parse_error();
            if (s < stop && *s == '/') {
              s++;
              if (!integer(vert.normal)) return parse_error();
            }
          }
        }
        if (vert.position == 0) break;
        if (vert.position < 0) vert.position = (int)obj.positions.size() + vert.position + 1;
        if (vert.texcoord < 0) vert.texcoord = (int)obj.texcoords.size() + vert.texcoord + 1;
        if (vert.normal < 0) vert.normal = (int)obj.normals.size() + vert.normal + 1;
        obj.vertices.push_back(vert);
        size++;
        skip();
      }
      obj.sizes.push_back(size);
      obj.types.push_back(cmd[0]);
    }
  }
  return true;
}

// load_subdiv for .obj: load_obj(face_varying = true) + get_positions / get_normals / get_texcoords(flip) / get_fvquads
// (yocto_sceneio.cpp:1209-1221, yocto_modelio.cpp:1857-1912)
bool load_subdiv(const std::string& filename, HostSubdiv& subdiv, std::string& error) {
  const auto ext = path_extension(filename);
  if (ext != ".obj" && ext != ".OBJ") return error = "unsupported format " + filename + " (subdivs are read from .obj)", false;
  ObjData obj;
  if (!read_obj(filename, obj, error)) return false;
  subdiv.positions = obj.positions, subdiv.normals = obj.normals, subdiv.texcoords = obj.texcoords;
  for (auto& t : subdiv.texcoords) t.y = 1 - t.y;  // get_texcoords(obj, texcoords, flipv = true)
  // get_fvquads: quads stay quads, other faces become fans of degenerate quads; a channel exists if the FIRST vertex of
  // the shape has it. (Like the reference, non-face elements do not advance the vertex cursor.)
  const auto& vertices = obj.vertices;
  const bool  has_pos = !vertices.empty() && vertices[0].position != 0, has_norm = !vertices.empty() && vertices[0].normal != 0,
              has_tex = !vertices.empty() && vertices[0].texcoord != 0;
  size_t cur = 0;
  for (size_t e = 0; e < obj.sizes.size(); e++) {
    if (obj.types[e] != 'f') continue;
    const int size = obj.sizes[e];
    auto emit = [&](int a, int b, int c, int d) {
      const ObjVertex &va = vertices[cur + a], &vb = vertices[cur + b], &vc = vertices[cur + c], &vd = vertices[cur + d];
      if (has_pos) subdiv.quadspos.push_back({va.position - 1, vb.position - 1, vc.position - 1, vd.position - 1});
      if (has_norm) subdiv.quadsnorm.push_back({va.normal - 1, vb.normal - 1, vc.normal - 1, vd.normal - 1});
      if (has_tex) subdiv.quadstexcoord.push_back({va.texcoord - 1, vb.texcoord - 1, vc.texcoord - 1, vd.texcoord - 1});
    };
    if (size == 4) emit(0, 1, 2, 3);
    else
      for (int c = 2; c < size; c++) emit(0, c - 1, c, c);
    cur += size;
  }
  if (subdiv.quadspos.empty()) return error = "empty shape " + filename, false;
  return true;
}

// load_shape for .obj (yocto_sceneio.cpp:1036-1051): load_obj(face_varying = false) - one vertex per distinct
// (position, texcoord, normal) triple in order of first use (yocto_modelio.cpp:1460-1489) - then get_faces / get_lines /
// get_points (:1771-1856), including their cursor rule: an element of another type does not advance the vertex cursor.
// The vertices of one OBJ shape -> de-duplicated vertex arrays and element arrays (load_obj's non-face-varying
// conversion, yocto_modelio.cpp:1330-1360, and the getters :1850-2010). `material` < 0 takes every element; otherwise only
// the elements of that material count - and, as there, only those advance the vertex cursor.
bool convert_obj_shape(const std::vector<v3>& positions, const std::vector<v3>& normals, const std::vector<v2>& texcoords,
    std::vector<ObjVertex>& vertices, const std::vector<int>& sizes, const std::vector<char>& types, const std::vector<int>* materials,
    int material, HostShape& shape) {
  for (auto& vertex : vertices)  // an index past its array is read unchecked by the reference: refused here
    if (vertex.position > (int)positions.size() || vertex.normal > (int)normals.size() || vertex.texcoord > (int)texcoords.size() ||
        vertex.position < 0 || vertex.normal < 0 || vertex.texcoord < 0)
      return false;
  std::unordered_map<ObjVertex, ObjVertex, ObjVertexHash> vertex_map;
  for (auto& vertex : vertices) {
    auto it = vertex_map.find(vertex);
    if (it == vertex_map.end()) {
      ObjVertex  fresh = vertex;
      const int  index = (int)vertex_map.size();
      if (vertex.position > 0) {
        const v3& p = positions[vertex.position - 1];
        shape.positions.insert(shape.positions.end(), {p.x, p.y, p.z});
        fresh.position = index + 1;
      }
      if (vertex.normal > 0) {
        const v3& n = normals[vertex.normal - 1];
        shape.normals.insert(shape.normals.end(), {n.x, n.y, n.z});
        fresh.normal = index + 1;
      }
      if (vertex.texcoord > 0) {
        const v2& t = texcoords[vertex.texcoord - 1];
        shape.texcoords.insert(shape.texcoords.end(), {t.x, 1 - t.y});  // get_texcoords(obj, texcoords, flipv = true)
        fresh.texcoord = index + 1;
      }
      vertex_map[vertex] = fresh;
      vertex             = fresh;
    } else {
      vertex = it->second;
    }
  }
  auto counts = [&](size_t e, char type) { return types[e] == type && (material < 0 || !materials || (*materials)[e] == material); };
  bool has_quads = false;
  for (size_t e = 0; e < sizes.size(); e++) has_quads |= types[e] == 'f' && sizes[e] == 4;
  auto pos = [&](size_t i) { return vertices[i].position - 1; };
  size_t cur = 0;
  for (size_t e = 0; e < sizes.size(); e++) {  // get_faces
    if (!counts(e, 'f')) continue;
    const int size = sizes[e];
    if (has_quads) {
      if (size == 4) shape.quads.insert(shape.quads.end(), {pos(cur), pos(cur + 1), pos(cur + 2), pos(cur + 3)});
      else
        for (int c = 2; c < size; c++) shape.quads.insert(shape.quads.end(), {pos(cur), pos(cur + c - 1), pos(cur + c), pos(cur + c)});
    } else {
      for (int c = 2; c < size; c++) shape.triangles.insert(shape.triangles.end(), {pos(cur), pos(cur + c - 1), pos(cur + c)});
    }
    cur += size;
  }
  cur = 0;
  for (size_t e = 0; e < sizes.size(); e++) {  // get_lines
    if (!counts(e, 'l')) continue;
    for (int c = 1; c < sizes[e]; c++) shape.lines.insert(shape.lines.end(), {pos(cur + c - 1), pos(cur + c)});
    cur += sizes[e];
  }
  cur = 0;
  for (size_t e = 0; e < sizes.size(); e++) {  // get_points (every point of an element is its first vertex, as there)
    if (!counts(e, 'p')) continue;
    for (int c = 0; c < sizes[e]; c++) shape.points.push_back(pos(cur));
    cur += sizes[e];
  }
  return true;
}
bool load_obj_shape(const std::string& filename, HostShape& shape, std::string& error) {
  ObjData obj;
  if (!read_obj(filename, obj, error)) return false;
  if (!convert_obj_shape(obj.positions, obj.normals, obj.texcoords, obj.vertices, obj.sizes, obj.types, nullptr, -1, shape))
    return error = "cannot parse " + filename + ": vertex index out of range", false;
  if (shape.points.empty() && shape.lines.empty() && shape.triangles.empty() && shape.quads.empty())
    return error = "empty shape " + filename, false;
  return true;
}

// tesselate_subdiv, yocto_scene.cpp:739-805: the subdivided, displaced, split surface replaces the shape
bool tesselate_subdiv(HostShape& shape, HostSubdiv subdiv, const std::vector<HostTexture>& textures, std::string& error) {
  // What the reference leaves undefined is refused here: an index outside its array, per-face arrays of different
  // lengths, and - when the mesh is subdivided - position and texcoord topologies that disagree on which faces are
  // triangles (q.z == q.w): those split into three and four faces and stop lining up (the reference then reads past the
  // shorter array). The normals of the file are dead as soon as the mesh is subdivided (recomputed if smooth, dropped if
  // not, :752-767), so they are neither checked nor subdivided then: the reference's own cube-subdiv.obj names normals
  // 9..24 of 8, which it reads out of bounds and then throws away.
  const bool normals_live = subdiv.subdivisions <= 0;
  if (!normals_live) subdiv.quadsnorm.clear(), subdiv.normals.clear();
  auto in_range = [](const std::vector<i4>& quads, size_t n) {
    if (n == 0) return true;  // no data: the topology is carried along but never dereferenced
    for (auto& q : quads)
      for (int c : {q.x, q.y, q.z, q.w})
        if (c < 0 || (size_t)c >= n) return false;
    return true;
  };
  if (!in_range(subdiv.quadspos, subdiv.positions.size()) || !in_range(subdiv.quadsnorm, subdiv.normals.size()) ||
      !in_range(subdiv.quadstexcoord, subdiv.texcoords.size()))
    return error = "subdiv index out of range", false;
  for (auto* other : {&subdiv.quadsnorm, &subdiv.quadstexcoord}) {
    if (other->empty()) continue;
    if (other->size() != subdiv.quadspos.size()) return error = "subdiv topologies differ in size", false;
    if (subdiv.subdivisions > 0)
      for (size_t f = 0; f < other->size(); f++)
        if (((*other)[f].z == (*other)[f].w) != (subdiv.quadspos[f].z == subdiv.quadspos[f].w))
          return error = "subdiv topologies disagree on a triangle", false;
  }
  if (subdiv.subdivisions > 0) {
    for (int k = 0; k < subdiv.subdivisions; k++) {
      if (subdiv.catmullclark) {
        subdivide_catmullclark(subdiv.quadstexcoord, subdiv.texcoords, true);
        subdivide_catmullclark(subdiv.quadsnorm, subdiv.normals, true);
        subdivide_catmullclark(subdiv.quadspos, subdiv.positions, false);
      } else {
        subdivide_quads(subdiv.quadstexcoord, subdiv.texcoords);
        subdivide_quads(subdiv.quadsnorm, subdiv.normals);
        subdivide_quads(subdiv.quadspos, subdiv.positions);
      }
    }
    if (subdiv.smooth) {
      subdiv.normals   = quads_normals(subdiv.quadspos, subdiv.positions);
      subdiv.quadsnorm = subdiv.quadspos;
    } else {
      subdiv.normals.clear(), subdiv.quadsnorm.clear();
    }
  }
  if (subdiv.displacement != 0 && subdiv.displacement_tex != -1) {
    if (subdiv.texcoords.empty()) return error = "missing texture coordinates", false;
    if (subdiv.displacement_tex < 0 || subdiv.displacement_tex >= (int)textures.size())
      return error = "displacement texture id out of range", false;
    const HostTexture& tex = textures[subdiv.displacement_tex];
    std::vector<float> offset(subdiv.positions.size(), 0);
    std::vector<int>   count(subdiv.positions.size(), 0);
    for (size_t fid = 0; fid < subdiv.quadspos.size(); fid++) {
      const int qpos[4] = {subdiv.quadspos[fid].x, subdiv.quadspos[fid].y, subdiv.quadspos[fid].z, subdiv.quadspos[fid].w};
      const int qtxt[4] = {subdiv.quadstexcoord[fid].x, subdiv.quadstexcoord[fid].y, subdiv.quadstexcoord[fid].z,
          subdiv.quadstexcoord[fid].w};
      for (int i = 0; i < 4; i++) {
        float disp = texture_mean(tex, subdiv.texcoords[qtxt[i]]);
        if (!tex.pixelsb.empty()) disp -= 0.5f;
        offset[qpos[i]] += subdiv.displacement * disp;
        count[qpos[i]] += 1;
      }
    }
    auto normals = quads_normals(subdiv.quadspos, subdiv.positions);
    for (size_t vid = 0; vid < subdiv.positions.size(); vid++)
      subdiv.positions[vid] = subdiv.positions[vid] + normals[vid] * offset[vid] / (float)count[vid];
    if (subdiv.smooth || !subdiv.normals.empty()) {
      subdiv.quadsnorm = subdiv.quadspos;
      subdiv.normals   = quads_normals(subdiv.quadspos, subdiv.positions);
    }
  }
  // split_facevarying, yocto_shape.cpp:2567-2618: one vertex per distinct (position, normal, texcoord) triple, in order
  // of first use
  shape = HostShape{};
  std::unordered_map<i3, int, i3_hash> vert_map;
  std::vector<i3>                      verts;
  shape.quads.resize(subdiv.quadspos.size() * 4);
  for (size_t fid = 0; fid < subdiv.quadspos.size(); fid++) {
    const int qp[4] = {subdiv.quadspos[fid].x, subdiv.quadspos[fid].y, subdiv.quadspos[fid].z, subdiv.quadspos[fid].w};
    for (int c = 0; c < 4; c++) {
      const int n = subdiv.quadsnorm.empty() ? -1 : (&subdiv.quadsnorm[fid].x)[c];
      const int t = subdiv.quadstexcoord.empty() ? -1 : (&subdiv.quadstexcoord[fid].x)[c];
      const i3  v = {qp[c], n, t};
      auto      it = vert_map.find(v);
      if (it == vert_map.end()) {
        const int index = (int)vert_map.size();
        vert_map.insert(it, {v, index});
        verts.push_back(v);
        shape.quads[fid * 4 + c] = index;
      } else {
        shape.quads[fid * 4 + c] = it->second;
      }
    }
  }
  if (!subdiv.positions.empty()) {
    shape.positions.resize(verts.size() * 3);
    for (size_t k = 0; k < verts.size(); k++) memcpy(&shape.positions[k * 3], &subdiv.positions[verts[k].x], 12);
  }
  if (!subdiv.normals.empty()) {
    shape.normals.resize(verts.size() * 3);
    for (size_t k = 0; k < verts.size(); k++) memcpy(&shape.normals[k * 3], &subdiv.normals[verts[k].y], 12);
  }
  if (!subdiv.texcoords.empty()) {
    shape.texcoords.resize(verts.size() * 2);
    for (size_t k = 0; k < verts.size(); k++) memcpy(&shape.texcoords[k * 2], &subdiv.texcoords[verts[k].z], 8);
  }
  return true;
}

}  // namespace

struct ygl_loaded_scene {
  std::vector<ygl_camera>      cameras;
  std::vector<ygl_instance>    instances;
  std::vector<ygl_environment> environments;
  std::vector<ygl_material>    materials;
  std::vector<HostShape>       shape_data;
  std::vector<HostTexture>     texture_data;
  std::vector<ygl_shape>       shapes;
  std::vector<ygl_texture>     textures;
  std::vector<std::string>     names[6];  // camera, texture, material, shape, instance, environment
  std::string                  copyright;
  ygl_scene_desc               desc = {};
};

namespace {

// what a scene file names next to its own arrays: files to read, subdivs to tesselate, and (format 4.0 only) the
// instance lists that expand one JSON object into many instances
struct SceneParts {
  std::vector<std::string>     shape_files, texture_files, subdiv_files;  // relative to the scene's directory
  std::vector<HostSubdiv>      subdivs;
  std::vector<std::string>     instance_files;                           // 4.0: instances/<name>.ply
  std::unordered_map<int, int> instance_ply;                              // 4.0: instance id -> instance_files index
};

float* frame_of(ygl_frame3f& f) { return (float*)&f; }
static_assert(sizeof(ygl_frame3f) == 48, "frame layout");

bool parse_json_scene_v42(const JValue& json, ygl_loaded_scene& scene, SceneParts& parts, JReader& rd);
bool parse_json_scene_v40(const JValue& json, const std::string& dirname, ygl_loaded_scene& scene, SceneParts& parts, JReader& rd);
bool load_scene_parts(const std::string& filename, ygl_loaded_scene& scene, SceneParts& parts, std::string& error);
void add_missing_camera(ygl_loaded_scene& scene);
void add_missing_radius(ygl_loaded_scene& scene);

// load_json_scene, yocto_sceneio.cpp:3618-3631: a file without asset.version is format 4.0 (:3025-3372), "4.1" goes to a
// loader that ends in `return false` whatever it read (:3614) - the reference cannot load such files and neither do we -,
// everything else must say 4.2 or 5.0
bool load_json_scene(const std::string& filename, ygl_loaded_scene& scene, std::string& error) {
  std::vector<uint8_t> text;
  if (!read_file(filename, text, error)) return false;
  text.push_back(0);
  JValue  json;
  JParser parser{(const char*)text.data(), (const char*)text.data() + text.size() - 1};
  auto    parse_error = [&]() { return error = "cannot parse " + filename, false; };
  if (!parser.parse_document(json)) return parse_error();

  JReader      rd;
  SceneParts   parts;
  const JValue* asset       = json.find("asset");
  const JValue* version_key = asset ? asset->find("version") : nullptr;
  if (asset) rd.get(*asset, "copyright", scene.copyright);
  if (!version_key) {
    if (json.type != JValue::Object) return parse_error();
    if (!parse_json_scene_v40(json, path_dirname(filename), scene, parts, rd) || !rd.ok) return parse_error();
  } else {
    if (version_key->type == JValue::String && version_key->string == "4.1")
      return error = "cannot load " + filename + ": the reference's loader for scene format 4.1 always fails", false;
    std::string version;
    rd.get(*asset, "version", version);
    if (!rd.ok || (version != "4.2" && version != "5.0")) return parse_error();
    if (!parse_json_scene_v42(json, scene, parts, rd) || !rd.ok) return parse_error();
  }
  return load_scene_parts(filename, scene, parts, error);
}

bool parse_json_scene_v42(const JValue& json, ygl_loaded_scene& scene, SceneParts& parts, JReader& rd) {
  // `for (auto& element : group)` of nlohmann: the elements of an array, the values of an object, nothing for null, the
  // value itself otherwise; get_opt on anything but an object throws - a parse error
  auto elements = [&](const char* name) {
    std::vector<const JValue*> out;
    auto group = json.find(name);
    if (!group || group->type == JValue::Null) return out;
    if (group->type == JValue::Array)
      for (auto& e : group->array) out.push_back(&e);
    else if (group->type == JValue::Object)
      for (auto& kv : group->object) out.push_back(&kv.second);
    else rd.ok = false;
    for (auto e : out)
      if (e->type != JValue::Object) rd.ok = false;
    if (!rd.ok) out.clear();
    return out;
  };
  auto& shape_files = parts.shape_files;
  auto& texture_files = parts.texture_files;
  {
    for (auto ep : elements("cameras")) {
      auto& e = *ep;
      ygl_camera camera = {};
      memcpy(&camera.frame, kIdentityFrame, 48);
      camera.orthographic = 0, camera.lens = 0.050f, camera.film = 0.036f, camera.aspect = 1.500f, camera.focus = 10000,
      camera.aperture = 0;
      std::string name;
      rd.get(e, "name", name);
      rd.get_floats(e, "frame", frame_of(camera.frame), 12);
      rd.get_bool(e, "orthographic", camera.orthographic);
      rd.get(e, "lens", camera.lens);
      rd.get(e, "aspect", camera.aspect);
      rd.get(e, "film", camera.film);
      rd.get(e, "focus", camera.focus);
      rd.get(e, "aperture", camera.aperture);
      if (e.find("lookat")) {
        float m[9] = {};
        rd.get_floats(e, "lookat", m, 9);
        v3 eye = {m[0], m[1], m[2]}, center = {m[3], m[4], m[5]}, up = {m[6], m[7], m[8]};
        camera.focus = length(eye - center);
        lookat_frame(frame_of(camera.frame), eye, center, up, false);
      }
      scene.cameras.push_back(camera);
      scene.names[0].push_back(name);
    }
  }
  {
    for (auto ep : elements("textures")) {
      auto& e = *ep;
      HostTexture tex;
      std::string name, uri;
      rd.get(e, "name", name);
      rd.get(e, "uri", uri);
      rd.get_bool(e, "linear", tex.linear);
      rd.get_bool(e, "nearest", tex.nearest);
      rd.get_bool(e, "clamp", tex.clamp);
      scene.texture_data.push_back(std::move(tex));
      scene.names[1].push_back(name);
      texture_files.push_back(uri);
    }
  }
  {
    for (auto ep : elements("materials")) {
      auto& e = *ep;
      ygl_material m = {};
      m.type = 0, m.roughness = 0, m.metallic = 0, m.ior = 1.5f, m.scanisotropy = 0, m.trdepth = 0.01f, m.opacity = 1;
      m.emission_tex = m.color_tex = m.roughness_tex = m.scattering_tex = m.normal_tex = -1;
      std::string name, type;
      rd.get(e, "name", name);
      if (auto t = e.find("type"))  // the enum conversion never throws: anything but a known label is the first entry
        m.type = t->type == JValue::String ? material_type_from_name(t->string) : 0;
      rd.get_floats(e, "emission", m.emission, 3);
      rd.get_floats(e, "color", m.color, 3);
      rd.get(e, "metallic", m.metallic);
      rd.get(e, "roughness", m.roughness);
      rd.get(e, "ior", m.ior);
      rd.get(e, "trdepth", m.trdepth);
      rd.get_floats(e, "scattering", m.scattering, 3);
      rd.get(e, "scanisotropy", m.scanisotropy);
      rd.get(e, "opacity", m.opacity);
      rd.get(e, "emission_tex", m.emission_tex);
      rd.get(e, "color_tex", m.color_tex);
      rd.get(e, "roughness_tex", m.roughness_tex);
      rd.get(e, "scattering_tex", m.scattering_tex);
      rd.get(e, "normal_tex", m.normal_tex);
      scene.materials.push_back(m);
      scene.names[2].push_back(name);
    }
  }
  {
    for (auto ep : elements("shapes")) {
      auto& e = *ep;
      std::string name, uri;
      rd.get(e, "name", name);
      rd.get(e, "uri", uri);
      scene.names[3].push_back(name);
      shape_files.push_back(uri);
    }
  }
  auto& subdivs      = parts.subdivs;
  auto& subdiv_files = parts.subdiv_files;
  {
    for (auto ep : elements("subdivs")) {
      auto& e = *ep;
      HostSubdiv  subdiv;
      std::string name, uri;
      rd.get(e, "name", name);
      rd.get(e, "uri", uri);
      rd.get(e, "shape", subdiv.shape);
      rd.get(e, "subdivisions", subdiv.subdivisions);
      rd.get_bool(e, "catmullclark", subdiv.catmullclark);
      rd.get_bool(e, "smooth", subdiv.smooth);
      rd.get(e, "displacement", subdiv.displacement);
      rd.get(e, "displacement_tex", subdiv.displacement_tex);
      subdivs.push_back(subdiv);
      subdiv_files.push_back(uri);
    }
  }
  {
    for (auto ep : elements("instances")) {
      auto& e = *ep;
      ygl_instance inst = {};
      memcpy(&inst.frame, kIdentityFrame, 48);
      inst.shape = inst.material = -1;
      std::string name;
      rd.get(e, "name", name);
      rd.get_floats(e, "frame", frame_of(inst.frame), 12);
      rd.get(e, "shape", inst.shape);
      rd.get(e, "material", inst.material);
      if (e.find("lookat")) {
        float m[9] = {};
        rd.get_floats(e, "lookat", m, 9);
        lookat_frame(frame_of(inst.frame), {m[0], m[1], m[2]}, {m[3], m[4], m[5]}, {m[6], m[7], m[8]}, true);
      }
      scene.instances.push_back(inst);
      scene.names[4].push_back(name);
    }
  }
  {
    for (auto ep : elements("environments")) {
      auto& e = *ep;
      ygl_environment env = {};
      memcpy(&env.frame, kIdentityFrame, 48);
      env.emission_tex = -1;
      std::string name;
      rd.get(e, "name", name);
      rd.get_floats(e, "frame", frame_of(env.frame), 12);
      rd.get_floats(e, "emission", env.emission, 3);
      rd.get(e, "emission_tex", env.emission_tex);
      if (e.find("lookat")) {
        float m[9] = {};
        rd.get_floats(e, "lookat", m, 9);
        lookat_frame(frame_of(env.frame), {m[0], m[1], m[2]}, {m[3], m[4], m[5]}, {m[6], m[7], m[8]}, true);
      }
      scene.environments.push_back(env);
      scene.names[5].push_back(name);
    }
  }
  return rd.ok;
}

// Scene format 4.0 (load_json_scene_version40, yocto_sceneio.cpp:3025-3372): every group is an object keyed by element
// name, read in the fixed order cameras, environments, materials, instances, objects, subdivs; shapes and textures have
// no entries of their own - an instance / material / environment / subdiv names them and the first mention creates
// them; material labels are the older set (metallic, volume); `lookat` never flips x and z; files are found by trying
// extensions under shapes/ textures/ subdivs/ instances/.
bool path_exists(const std::string& path) {
  FILE* f = fopen(path.c_str(), "rb");
  if (f) fclose(f);
  return f != nullptr;
}
bool parse_json_scene_v40(const JValue& json, const std::string& dirname, ygl_loaded_scene& scene, SceneParts& parts, JReader& rd) {
  if (auto asset = json.find("asset"); asset && asset->type != JValue::Object) return false;
  // nlohmann's items(): object members in file order, array elements under their index, nothing for null
  using Item = std::pair<std::string, const JValue*>;
  auto items = [&](const char* name) {
    std::vector<Item> out;
    auto group = json.find(name);
    if (!group || group->type == JValue::Null) return out;
    if (group->type == JValue::Object) {
      for (auto& kv : group->object) out.push_back({kv.first, &kv.second});
    } else if (group->type == JValue::Array) {
      for (size_t i = 0; i < group->array.size(); i++) out.push_back({std::to_string(i), &group->array[i]});
    } else {
      rd.ok = false;
    }
    for (auto& item : out)
      if (item.second->type != JValue::Object) rd.ok = false;  // json.value(key, default) throws on anything else
    if (!rd.ok) out.clear();
    return out;
  };
  // a reference by name; an empty or missing name keeps the default (-1)
  std::vector<std::string>             shape_names, texture_names, subdiv_names;
  std::unordered_map<std::string, int> shape_map, texture_map, material_map, instance_map;
  auto named = [&](const JValue& e, const char* key, std::string& name) {
    name.clear();
    rd.get(e, key, name);
    return rd.ok && !name.empty();
  };
  auto get_created = [&](const JValue& e, const char* key, int& value, std::unordered_map<std::string, int>& map,
                         std::vector<std::string>& names) {
    std::string name;
    if (!named(e, key, name)) return;
    auto it = map.find(name);
    if (it == map.end()) {
      names.push_back(name);
      it = map.insert({name, (int)names.size() - 1}).first;
    }
    value = it->second;
  };
  auto lookat = [&](const JValue& e, ygl_frame3f& frame, float* focus) {
    if (!e.find("lookat")) return;
    float m[9];
    memcpy(m, &frame, sizeof(m));  // the default of the read is what the frame holds
    rd.get_floats(e, "lookat", m, 9);
    v3 eye = {m[0], m[1], m[2]}, center = {m[3], m[4], m[5]}, up = {m[6], m[7], m[8]};
    if (focus) *focus = length(eye - center);
    lookat_frame(frame_of(frame), eye, center, up, false);
  };
  for (auto& [key, e] : items("cameras")) {
    ygl_camera camera = {};
    memcpy(&camera.frame, kIdentityFrame, 48);
    camera.orthographic = 0, camera.lens = 0.050f, camera.film = 0.036f, camera.aspect = 1.500f, camera.focus = 10000,
    camera.aperture = 0;
    rd.get_floats(*e, "frame", frame_of(camera.frame), 12);
    rd.get_bool(*e, "orthographic", camera.orthographic);
    rd.get_bool(*e, "ortho", camera.orthographic);
    rd.get(*e, "lens", camera.lens);
    rd.get(*e, "aspect", camera.aspect);
    rd.get(*e, "film", camera.film);
    rd.get(*e, "focus", camera.focus);
    rd.get(*e, "aperture", camera.aperture);
    lookat(*e, camera.frame, &camera.focus);
    scene.cameras.push_back(camera);
    scene.names[0].push_back(key);
  }
  for (auto& [key, e] : items("environments")) {
    ygl_environment env = {};
    memcpy(&env.frame, kIdentityFrame, 48);
    env.emission_tex = -1;
    rd.get_floats(*e, "frame", frame_of(env.frame), 12);
    rd.get_floats(*e, "emission", env.emission, 3);
    get_created(*e, "emission_tex", env.emission_tex, texture_map, texture_names);
    lookat(*e, env.frame, nullptr);
    scene.environments.push_back(env);
    scene.names[5].push_back(key);
  }
  for (auto& [key, e] : items("materials")) {
    ygl_material m = {};
    m.type = 0, m.roughness = 0, m.metallic = 0, m.ior = 1.5f, m.scanisotropy = 0, m.trdepth = 0.01f, m.opacity = 1;
    m.emission_tex = m.color_tex = m.roughness_tex = m.scattering_tex = m.normal_tex = -1;
    material_map[key] = (int)scene.materials.size();
    if (auto t = e->find("type"); t && t->type == JValue::String) {
      static const char* labels[] = {"matte", "glossy", "metallic", "transparent", "refractive", "subsurface", "volume", "gltfpbr"};
      for (int i = 0; i < 8; i++)
        if (t->string == labels[i]) m.type = i;  // same ordinals as today's reflective / volumetric
    }
    rd.get_floats(*e, "emission", m.emission, 3);
    rd.get_floats(*e, "color", m.color, 3);
    rd.get(*e, "metallic", m.metallic);
    rd.get(*e, "roughness", m.roughness);
    rd.get(*e, "ior", m.ior);
    rd.get(*e, "trdepth", m.trdepth);
    rd.get_floats(*e, "scattering", m.scattering, 3);
    rd.get(*e, "scanisotropy", m.scanisotropy);
    rd.get(*e, "opacity", m.opacity);
    get_created(*e, "emission_tex", m.emission_tex, texture_map, texture_names);
    get_created(*e, "color_tex", m.color_tex, texture_map, texture_names);
    get_created(*e, "roughness_tex", m.roughness_tex, texture_map, texture_names);
    get_created(*e, "scattering_tex", m.scattering_tex, texture_map, texture_names);
    get_created(*e, "normal_tex", m.normal_tex, texture_map, texture_names);
    scene.materials.push_back(m);
    scene.names[2].push_back(key);
  }
  for (const char* group : {"instances", "objects"}) {
    for (auto& [key, e] : items(group)) {
      ygl_instance inst = {};
      memcpy(&inst.frame, kIdentityFrame, 48);
      inst.shape = inst.material = -1;
      rd.get_floats(*e, "frame", frame_of(inst.frame), 12);
      get_created(*e, "shape", inst.shape, shape_map, shape_names);
      std::string material;
      if (named(*e, "material", material)) {  // materials are not created on mention: an unknown name is an error
        auto it = material_map.find(material);
        if (it == material_map.end()) rd.ok = false;
        else inst.material = it->second;
      }
      lookat(*e, inst.frame, nullptr);
      if (group[0] == 'o' && e->find("instance")) {  // only "objects" read the instance list
        int list = -1;
        get_created(*e, "instance", list, instance_map, parts.instance_files);
        if (list >= 0) parts.instance_ply[(int)scene.instances.size()] = list;
      }
      scene.instances.push_back(inst);
      scene.names[4].push_back(key);
    }
  }
  for (auto& [key, e] : items("subdivs")) {
    HostSubdiv subdiv;
    get_created(*e, "shape", subdiv.shape, shape_map, shape_names);
    rd.get(*e, "subdivisions", subdiv.subdivisions);
    rd.get_bool(*e, "catmullclark", subdiv.catmullclark);
    rd.get_bool(*e, "smooth", subdiv.smooth);
    rd.get(*e, "displacement", subdiv.displacement);
    get_created(*e, "displacement_tex", subdiv.displacement_tex, texture_map, texture_names);
    parts.subdivs.push_back(subdiv);
    subdiv_names.push_back(key);
  }
  if (!rd.ok) return false;
  // find_path, yocto_sceneio.cpp:3250-3258: the first extension whose file exists, else the first extension
  auto find_path = [&](const std::string& name, const char* group, std::initializer_list<const char*> extensions) {
    for (auto ext : extensions)
      if (path_exists(path_join(dirname, path_join(group, name + ext)))) return path_join(group, name + ext);
    return path_join(group, name + *extensions.begin());
  };
  for (auto& name : shape_names) parts.shape_files.push_back(find_path(name, "shapes", {".ply", ".obj"}));
  for (auto& name : subdiv_names) parts.subdiv_files.push_back(find_path(name, "subdivs", {".ply", ".obj"}));
  for (auto& name : texture_names) parts.texture_files.push_back(find_path(name, "textures", {".hdr", ".exr", ".png", ".jpg"}));
  for (auto& name : parts.instance_files) name = find_path(name, "instances", {".ply"});
  scene.names[1] = texture_names;
  scene.names[3] = shape_names;
  return true;
}

// frame3f * frame3f, yocto_math.h:2108-2110 with mat3f * vec3f = a.x * b.x + a.y * b.y + a.z * b.z (:1942-1944)
void frame_multiply(float* out, const float* a, const float* b) {
  auto rot = [&](const float* v, float* r) {
    for (int k = 0; k < 3; k++) r[k] = (a[0 + k] * v[0] + a[3 + k] * v[1]) + a[6 + k] * v[2];
  };
  float r[12];
  rot(b + 0, r + 0), rot(b + 3, r + 3), rot(b + 6, r + 6), rot(b + 9, r + 9);
  for (int k = 0; k < 3; k++) r[9 + k] = r[9 + k] + a[9 + k];
  memcpy(out, r, sizeof(r));
}

// load_instance, yocto_sceneio.cpp:2873-2892: element "instance" with twelve scalar properties
bool load_instance_frames(const std::string& filename, std::vector<float>& frames, std::string& error) {
  if (path_extension(filename) != ".ply") return error = "unsupported format " + filename, false;
  std::vector<PlyElement> ply;
  if (!load_ply(filename, ply, error)) return false;
  static const char* names[12] = {"xx", "xy", "xz", "yx", "yy", "yz", "zx", "zy", "zz", "ox", "oy", "oz"};
  const PlyProperty* props[12];
  for (int k = 0; k < 12; k++) {
    props[k] = ply_find(ply, "instance", names[k]);
    if (!props[k] || props[k]->is_list) return error = "cannot parse " + filename, false;
  }
  const size_t n = props[0]->values.size();
  frames.assign(n * 12, 0.0f);
  for (int k = 0; k < 12; k++)
    for (size_t i = 0; i < n && i < props[k]->values.size(); i++) frames[i * 12 + k] = (float)props[k]->values[i];
  return true;
}

// the part every JSON format shares: read the files the scene names (in parallel like the reference), expand 4.0's
// instance lists, apply the fix-ups, tesselate
bool load_scene_parts(const std::string& filename, ygl_loaded_scene& scene, SceneParts& parts, std::string& error) {
  auto& shape_files = parts.shape_files;
  auto& texture_files = parts.texture_files;
  auto& subdiv_files = parts.subdiv_files;
  auto& subdivs = parts.subdivs;
  const auto dirname = path_dirname(filename);
  scene.shape_data.resize(shape_files.size());
  scene.texture_data.resize(texture_files.size());
  auto dependent_error = [&]() { return error = "cannot load " + filename + " since " + error, false; };
  // One work list for every file the scene names - shapes, subdiv control meshes, textures, instance lists - instead of
  // the reference's four passes: the largest file of each kind no longer waits for the other kinds. Errors keep the
  // reference's precedence (the first failing shape, else the first failing subdiv, ...): the list is in that order and
  // the lowest failing index wins.
  std::vector<std::vector<float>> instance_frames(parts.instance_files.size());
  const size_t n_shapes = shape_files.size(), n_subdivs = subdiv_files.size(), n_textures = texture_files.size();
  const size_t n_files = n_shapes + n_subdivs + n_textures + parts.instance_files.size();
  std::vector<size_t> order(n_files), bytes(n_files, 0);
  for (size_t i = 0; i < n_files; i++) {
    const auto& name = i < n_shapes ? shape_files[i] : i < n_shapes + n_subdivs ? subdiv_files[i - n_shapes]
                       : i < n_shapes + n_subdivs + n_textures ? texture_files[i - n_shapes - n_subdivs]
                                                               : parts.instance_files[i - n_shapes - n_subdivs - n_textures];
    struct stat st;
    if (stat(path_join(dirname, name).c_str(), &st) == 0) bytes[i] = (size_t)st.st_size;
    order[i] = i;
  }
  std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return bytes[a] > bytes[b]; });
  if (!parallel_load(n_files, error, [&](size_t i, std::string& err) {
        if (i < n_shapes) return load_shape(path_join(dirname, shape_files[i]), scene.shape_data[i], err);
        i -= n_shapes;
        if (i < n_subdivs) return load_subdiv(path_join(dirname, subdiv_files[i]), subdivs[i], err);
        i -= n_subdivs;
        if (i < n_textures) return load_texture(path_join(dirname, texture_files[i]), scene.texture_data[i], err);
        i -= n_textures;
        return load_instance_frames(path_join(dirname, parts.instance_files[i]), instance_frames[i], err);
      }, &order))
    return dependent_error();
  // "apply instances", yocto_sceneio.cpp:3335-3363: an object with an instance list becomes one instance per listed
  // frame (list frame * object frame), named <object>_<k>
  if (!parts.instance_files.empty()) {
    auto instances = std::move(scene.instances);
    auto names     = std::move(scene.names[4]);
    scene.instances.clear(), scene.names[4].clear();
    for (size_t i = 0; i < instances.size(); i++) {
      auto it = parts.instance_ply.find((int)i);
      if (it == parts.instance_ply.end()) {
        scene.instances.push_back(instances[i]);
        scene.names[4].push_back(names[i]);
        continue;
      }
      const auto& frames = instance_frames[it->second];
      for (size_t k = 0; k < frames.size() / 12; k++) {
        ygl_instance inst = instances[i];
        frame_multiply(frame_of(inst.frame), &frames[k * 12], frame_of(instances[i].frame));
        scene.instances.push_back(inst);
        scene.names[4].push_back(names[i] + "_" + std::to_string(k));
      }
    }
  }
  add_missing_camera(scene);
  add_missing_radius(scene);
  // tesselate_subdivs (yocto_scene.cpp:807-812): what every reference app does right after load_scene
  std::vector<char> targeted(scene.shape_data.size(), 0);
  bool              distinct = true;
  for (auto& subdiv : subdivs) {
    if (subdiv.shape < 0 || subdiv.shape >= (int)scene.shape_data.size())
      return error = "cannot load " + filename + ": subdiv shape id out of range", false;
    distinct = distinct && !targeted[subdiv.shape];
    targeted[subdiv.shape] = 1;
  }
  if (distinct) {  // every subdiv writes its own shape: one core each (the reference walks them one after the other)
    if (!parallel_load(subdivs.size(), error, [&](size_t i, std::string& err) {
          return tesselate_subdiv(scene.shape_data[subdivs[i].shape], subdivs[i], scene.texture_data, err);
        }))
      return dependent_error();
  } else {
    for (auto& subdiv : subdivs)
      if (!tesselate_subdiv(scene.shape_data[subdiv.shape], subdiv, scene.texture_data, error)) return dependent_error();
  }
  return true;
}

// fix-ups: add_missing_camera (yocto_sceneio.cpp:2119-2139), add_missing_radius (:2142-2148)
void add_missing_camera(ygl_loaded_scene& scene) {
  if (scene.cameras.empty()) {
    struct box {
      v3 min = {3.402823466e+38f, 3.402823466e+38f, 3.402823466e+38f}, max = {-3.402823466e+38f, -3.402823466e+38f, -3.402823466e+38f};
    };
    auto fmin = [](float a, float b) { return (a < b) ? a : b; };
    auto fmax = [](float a, float b) { return (a > b) ? a : b; };
    auto grow = [&](box& b, const v3& p) {
      b.min = {fmin(b.min.x, p.x), fmin(b.min.y, p.y), fmin(b.min.z, p.z)};
      b.max = {fmax(b.max.x, p.x), fmax(b.max.y, p.y), fmax(b.max.z, p.z)};
    };
    std::vector<box> shape_box(scene.shape_data.size());
    for (size_t s = 0; s < scene.shape_data.size(); s++) {
      auto& pos = scene.shape_data[s].positions;
      for (size_t i = 0; i + 2 < pos.size(); i += 3) grow(shape_box[s], {pos[i], pos[i + 1], pos[i + 2]});
    }
    box bbox;
    for (auto& inst : scene.instances) {
      if (inst.shape < 0 || inst.shape >= (int)shape_box.size()) continue;
      const box& sb = shape_box[inst.shape];
      const float* f = (const float*)&inst.frame;
      v3 fx = {f[0], f[1], f[2]}, fy = {f[3], f[4], f[5]}, fz = {f[6], f[7], f[8]}, fo = {f[9], f[10], f[11]};
      v3 corners[8] = {{sb.min.x, sb.min.y, sb.min.z}, {sb.min.x, sb.min.y, sb.max.z}, {sb.min.x, sb.max.y, sb.min.z},
          {sb.min.x, sb.max.y, sb.max.z}, {sb.max.x, sb.min.y, sb.min.z}, {sb.max.x, sb.min.y, sb.max.z},
          {sb.max.x, sb.max.y, sb.min.z}, {sb.max.x, sb.max.y, sb.max.z}};
      box xf;
      for (auto& c : corners) grow(xf, fx * c.x + fy * c.y + fz * c.z + fo);  // transform_point, yocto_math.h
      grow(bbox, xf.min), grow(bbox, xf.max);
    }
    ygl_camera camera = {};
    camera.orthographic = 0, camera.film = 0.036f, camera.aspect = (float)16 / (float)9, camera.aperture = 0, camera.lens = 0.050f;
    auto center      = (bbox.max + bbox.min) / 2;
    auto bbox_radius = length(bbox.max - bbox.min) / 2;
    auto camera_dist = bbox_radius * camera.lens / (camera.film / camera.aspect);
    camera_dist *= 2.0f;
    v3 from = v3{0, 0, 1} * camera_dist + center, to = center;
    lookat_frame((float*)&camera.frame, from, to, {0, 1, 0}, false);
    camera.focus = length(from - to);
    scene.cameras.push_back(camera);
    scene.names[0].push_back("camera");
  }
}
void add_missing_radius(ygl_loaded_scene& scene) {
  for (auto& shape : scene.shape_data) {
    if (shape.points.empty() && shape.lines.empty()) continue;
    if (!shape.radius.empty()) continue;
    shape.radius.assign(shape.positions.size() / 3, 0.001f);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// A single .ply file as a scene (load_ply_scene, yocto_sceneio.cpp:4364-4381): the shape under an identity frame, the
// default matte material (add_missing_material, :2151-2163), the framing camera, and - since such a scene never has a
// light - the procedural sky of add_sky (yocto_scene.cpp:645-653): make_sunsky(1024, 512, pi / 4) as a linear
// environment texture with emission 1.
// make_sunsky (yocto_image.cpp:1245-1373) is the Preetham / Perez daylight model in xyY, evaluated for the upper half of
// a lat-long image, with the lower half filled by the light the ground reflects. Float arithmetic in the reference's
// order through this host's libm (the same one the reference calls); the sun disk is off in add_sky's call.
// ---------------------------------------------------------------------------------------------------------------
struct SkyModel {
  float zenith[3];       // xyY at the zenith
  float perez[5][3];     // A..E per xyY channel
};
// cubic in the sun angle with the reference's association: ((k3 t^3 + k2 t^2) + k1 t) + k0
float sun_cubic(const float (&k)[4], float t) { return k[0] * std::pow(t, 3.f) + k[1] * std::pow(t, 2.f) + k[2] * t + k[3]; }
SkyModel make_sky_model(float theta_sun, float turbidity) {
  static const float chroma[2][3][4] = {
      {{+0.00165f, -0.00374f, +0.00208f, +0.00000f}, {-0.02902f, +0.06377f, -0.03202f, +0.00394f}, {+0.11693f, -0.21196f, +0.06052f, +0.25885f}},
      {{+0.00275f, -0.00610f, +0.00316f, +0.00000f}, {-0.04214f, +0.08970f, -0.04153f, +0.00515f}, {+0.15346f, -0.26756f, +0.06669f, +0.26688f}}};
  static const float perez_slope[5][3]  = {{-0.01925f, -0.01669f, +0.17872f}, {-0.06651f, -0.09495f, -0.35540f},
       {-0.00041f, -0.00792f, -0.02266f}, {-0.06409f, -0.04405f, +0.12064f}, {-0.00325f, -0.01092f, -0.06696f}};
  static const float perez_offset[5][3] = {{-0.25922f, -0.26078f, -1.46303f}, {+0.00081f, +0.00921f, +0.42749f},
      {+0.21247f, +0.21023f, +5.32505f}, {-0.89887f, -1.65369f, -2.57705f}, {+0.04517f, +0.05291f, +0.37027f}};
  const float pif = 3.14159265358979323846f;
  SkyModel sky;
  for (int c = 0; c < 2; c++)
    sky.zenith[c] = sun_cubic(chroma[c][0], theta_sun) * std::pow(turbidity, 2.f) + sun_cubic(chroma[c][1], theta_sun) * turbidity +
                    sun_cubic(chroma[c][2], theta_sun);
  sky.zenith[2] = 1000 * (4.0453f * turbidity - 4.9710f) * std::tan((4.0f / 9.0f - turbidity / 120.0f) * (pif - 2 * theta_sun)) -
                  .2155f * turbidity + 2.4192f;
  for (int k = 0; k < 5; k++)
    for (int c = 0; c < 3; c++) sky.perez[k][c] = perez_slope[k][c] * turbidity + perez_offset[k][c];
  return sky;
}
// one sky colour: Perez' F(theta, gamma) / F(0, theta_sun) per xyY channel, then xyY -> XYZ -> linear sRGB, / 10000
void sky_radiance(const SkyModel& sky, float theta, float gamma, float theta_sun, float rgb[3]) {
  float xyY[3];
  for (int c = 0; c < 3; c++) {
    const float A = sky.perez[0][c], B = sky.perez[1][c], C = sky.perez[2][c], D = sky.perez[3][c], E = sky.perez[4][c];
    const float num = (1 + A * std::exp(B / std::cos(theta))) * (1 + C * std::exp(D * gamma) + E * std::cos(gamma) * std::cos(gamma));
    const float den = (1 + A * std::exp(B)) * (1 + C * std::exp(D * theta_sun) + E * std::cos(theta_sun) * std::cos(theta_sun));
    xyY[c]          = sky.zenith[c] * num / den;
  }
  float xyz[3] = {0, 0, 0};  // xyY_to_xyz, yocto_color.h:402-405
  if (xyY[1] != 0) xyz[0] = xyY[0] * xyY[2] / xyY[1], xyz[1] = xyY[2], xyz[2] = (1 - xyY[0] - xyY[1]) * xyY[2] / xyY[1];
  static const float to_rgb[3][3] = {{+3.2406f, -0.9689f, +0.0557f}, {-1.5372f, +1.8758f, -0.2040f}, {-0.4986f, +0.0415f, +1.0570f}};
  for (int c = 0; c < 3; c++) rgb[c] = (to_rgb[0][c] * xyz[0] + to_rgb[1][c] * xyz[1] + to_rgb[2][c] * xyz[2]) / 10000;
}
void parallel_rows(int rows, const std::function<void(int)>& fn) {
  const unsigned nthreads = std::max(1u, std::min((unsigned)rows, (unsigned)host_parallelism()));
  std::atomic<int>         next{0};
  std::vector<std::thread> pool;
  for (unsigned t = 0; t < nthreads; t++)
    pool.emplace_back([&]() {
      for (int j = next++; j < rows; j = next++) fn(j);
    });
  for (auto& t : pool) t.join();
}
void make_sky_texture(HostTexture& tex, int width, int height, float theta_sun, float turbidity, const float (&ground_albedo)[3]) {
  const float pif = 3.14159265358979323846f, eps = 1.1920928955078125e-07f;
  auto fmin = [](float a, float b) { return (a < b) ? a : b; };
  auto fmax = [](float a, float b) { return (a > b) ? a : b; };
  const SkyModel sky          = make_sky_model(theta_sun, turbidity);
  const float    sun_dir[3]   = {0, std::cos(theta_sun), std::sin(theta_sun)};
  tex.width = width, tex.height = height, tex.linear = 1;
  tex.pixelsf.assign((size_t)width * height * 4, 0.0f);
  // rows are independent: one strip per core (the ground term below is a running sum and stays in pixel order)
  parallel_rows(height / 2, [&](int j) {
    float theta = pif * ((j + 0.5f) / height);
    theta       = fmin(fmax(theta, 0.0f), pif / 2 - eps);
    for (int i = 0; i < width; i++) {
      const float phi   = 2 * pif * (float(i + 0.5f) / width);
      const float w[3]  = {std::cos(phi) * std::sin(theta), std::cos(theta), std::sin(phi) * std::sin(theta)};
      const float gamma = std::acos(fmin(fmax(w[0] * sun_dir[0] + w[1] * sun_dir[1] + w[2] * sun_dir[2], -1.0f), 1.0f));
      float rgb[3];
      sky_radiance(sky, theta, gamma, theta_sun, rgb);
      float* px = &tex.pixelsf[((size_t)j * width + i) * 4];
      px[0] = rgb[0] + 0.0f, px[1] = rgb[1] + 0.0f, px[2] = rgb[2] + 0.0f, px[3] = 1;  // "+ sun colour", which is zero here
    }
  });
  float ground[3] = {0, 0, 0};
  if (ground_albedo[0] != 0 || ground_albedo[1] != 0 || ground_albedo[2] != 0) {
    for (int j = 0; j < height / 2; j++) {
      const float theta = pif * ((j + 0.5f) / height);
      for (int i = 0; i < width; i++) {
        const float* le    = &tex.pixelsf[((size_t)j * width + i) * 4];
        const float  angle = std::sin(theta) * 4 * pif / (width * height);
        for (int c = 0; c < 3; c++) ground[c] += le[c] * (ground_albedo[c] / pif) * std::cos(theta) * angle;
      }
    }
  }
  for (int j = height / 2; j < height; j++)
    for (int i = 0; i < width; i++) {
      float* px = &tex.pixelsf[((size_t)j * width + i) * 4];
      px[0] = ground[0], px[1] = ground[1], px[2] = ground[2], px[3] = 1;
    }
}

bool load_ply_scene(const std::string& filename, ygl_loaded_scene& scene, std::string& error) {
  scene.shape_data.resize(1);
  // the sky does not depend on the shape: it is made while the file is read. The model's parameters reach it through
  // memory the compiler cannot see through, as they reach the reference's make_sunsky from another translation unit:
  // libm is called at run time, nothing is folded at build time
  volatile float sun_angle = 3.14159265358979323846f / 4, turbidity = 3;
  HostTexture    sky;
  std::thread    sky_maker([&]() { make_sky_texture(sky, 1024, 512, sun_angle, turbidity, {0.2f, 0.2f, 0.2f}); });
  const bool     loaded = load_shape(filename, scene.shape_data[0], error);
  sky_maker.join();
  if (!loaded) return false;
  ygl_instance inst = {};
  memcpy(&inst.frame, kIdentityFrame, 48);
  inst.shape = 0, inst.material = 0;
  scene.instances.push_back(inst);
  ygl_material m = {};
  m.type = 0, m.roughness = 0, m.metallic = 0, m.ior = 1.5f, m.scanisotropy = 0, m.trdepth = 0.01f, m.opacity = 1;
  m.emission_tex = m.color_tex = m.roughness_tex = m.scattering_tex = m.normal_tex = -1;
  m.color[0] = m.color[1] = m.color[2] = 0.8f;
  scene.materials.push_back(m);
  add_missing_camera(scene);
  add_missing_radius(scene);
  scene.texture_data.push_back(std::move(sky));
  scene.names[1].push_back("sky");
  ygl_environment env = {};
  memcpy(&env.frame, kIdentityFrame, 48);
  env.emission[0] = env.emission[1] = env.emission[2] = 1;
  env.emission_tex = 0;
  scene.environments.push_back(env);
  scene.names[5].push_back("sky");
  return true;
}

// ---------------------------------------------------------------------------------------------------------------
// glTF 2.0 (.gltf with external or base64 buffers, .glb): load_gltf_scene, yocto_sceneio.cpp:4430-4767, which reads
// the file with cgltf (exts/cgltf) and converts: one shape per mesh primitive (POSITION / NORMAL / TEXCOORD_0 /
// COLOR_0 / RADIUS through cgltf_accessor_read_float: component types, normalisation, strides; indices or implied
// indices for triangles, strips, fans, lines, loops, line strips), one instance per primitive of every node with a
// mesh under the node's world transform (cgltf_node_transform_world: TRS or matrix, multiplied up the parent chain in
// cgltf's float association), one camera per node with a camera, gltfpbr materials (base colour / alpha, metallic,
// roughness, emission x KHR_materials_emissive_strength, KHR_materials_transmission -> transparent), one texture per
// image (uri, "%20" -> space), then add_missing_material / camera / radius and - a glTF file has no environment - the
// procedural sky. Numbers are read as cgltf reads them (atof, narrowed to float). Where cgltf reads unchecked
// (accessors past their buffers, indices past their arrays) the file is refused here.
// ---------------------------------------------------------------------------------------------------------------
struct GltfAccessor {
  int    view = -1, component = 0, components = 0;
  size_t offset = 0, count = 0, stride = 0;
  bool   normalized = false, sparse = false, matrix = false;
};
struct GltfView {
  int    buffer = -1;
  size_t offset = 0, size = 0, stride = 0;
};
float gltf_float(const JValue& v) { return (float)(v.kind == JValue::Float ? v.number : v.kind == JValue::Signed ? (double)v.inumber : (double)v.unumber); }
long long gltf_int(const JValue& v) { return v.kind == JValue::Float ? (long long)v.number : v.kind == JValue::Signed ? (long long)v.inumber : (long long)v.unumber; }
bool base64_decode(const char* text, size_t size, std::vector<uint8_t>& out) {  // cgltf_load_buffer_base64: `size` bytes
  out.resize(size);
  unsigned buffer = 0, bits = 0;
  for (size_t i = 0; i < size; i++) {
    while (bits < 8) {
      const char ch = *text++;
      const int  index = (unsigned)(ch - 'A') < 26 ? ch - 'A' : (unsigned)(ch - 'a') < 26 ? ch - 'a' + 26 : (unsigned)(ch - '0') < 10 ? ch - '0' + 52
                         : ch == '+' ? 62 : ch == '/' ? 63 : -1;
      if (index < 0) return false;
      buffer = (buffer << 6) | (unsigned)index;
      bits += 6;
    }
    out[i] = (uint8_t)(buffer >> (bits - 8));
    bits -= 8;
  }
  return true;
}
std::string uri_decode(const std::string& uri) {  // cgltf_decode_uri: %XX
  std::string out;
  auto hex = [](char c) { return (unsigned)(c - '0') < 10 ? c - '0' : (unsigned)(c - 'a') < 6 ? c - 'a' + 10 : (unsigned)(c - 'A') < 6 ? c - 'A' + 10 : -1; };
  for (size_t i = 0; i < uri.size(); i++) {
    if (uri[i] == '%' && i + 2 < uri.size() + 0 && hex(uri[i + 1]) >= 0 && hex(uri[i + 2]) >= 0) {
      out += (char)(hex(uri[i + 1]) * 16 + hex(uri[i + 2]));
      i += 2;
    } else {
      out += uri[i];
    }
  }
  return out;
}

bool load_gltf_scene(const std::string& filename, ygl_loaded_scene& scene, std::string& error) {
  std::vector<uint8_t> file;
  if (!read_file(filename, file, error)) return false;
  auto parse_error = [&]() { return error = "cannot parse " + filename, false; };
  auto unsupported = [&](const char* what) { return error = "cannot load " + filename + " for sunsupported " + what, false; };
  auto buffers_error = [&]() { return error = "cannot load " + filename + " since cannot load buffers", false; };
  // container: a .glb wraps the JSON and one binary chunk
  const uint8_t* json_begin = file.data();
  size_t         json_size  = file.size();
  const uint8_t* bin        = nullptr;
  size_t         bin_size   = 0;
  auto u32 = [&](size_t at) { uint32_t v; memcpy(&v, &file[at], 4); return v; };
  if (file.size() >= 4 && u32(0) == 0x46546C67u) {
    if (file.size() < 20 || u32(4) != 2 || u32(8) > file.size()) return parse_error();
    const size_t total = u32(8), json_len = u32(12);
    if (u32(16) != 0x4E4F534Au || 20 + json_len > total) return parse_error();
    json_begin = &file[20], json_size = json_len;
    const size_t next = 20 + json_len;
    if (next + 8 <= total) {
      const size_t len = u32(next);
      if (u32(next + 4) != 0x004E4942u || next + 8 + len > total) return parse_error();
      bin = &file[next + 8], bin_size = len;
    }
  }
  JValue  json;
  JParser parser{(const char*)json_begin, (const char*)json_begin + json_size};
  if (!parser.parse_document(json) || json.type != JValue::Object) return parse_error();
  auto array_of = [&](const JValue& parent, const char* key) -> const std::vector<JValue>& {
    static const std::vector<JValue> none;
    auto v = parent.find(key);
    return v && v->type == JValue::Array ? v->array : none;
  };
  bool sizes_ok = true;  // an offset / length / count / index that is negative or absurd makes the file invalid
  auto number_at = [&](const JValue& e, const char* key, double fallback) {
    auto v = e.find(key);
    if (!v || v->type != JValue::Number) return fallback;
    const double value = (double)gltf_int(*v);
    if (!(value >= -1 && value < 1e12)) sizes_ok = false;
    return value < 0 ? -1.0 : value;
  };
  auto float_at = [&](const JValue& e, const char* key, float fallback) {
    auto v = e.find(key);
    return v && v->type == JValue::Number ? gltf_float(*v) : fallback;
  };
  auto floats_at = [&](const JValue& e, const char* key, float* out, size_t n) {
    auto v = e.find(key);
    if (!v || v->type != JValue::Array) return false;
    for (size_t k = 0; k < n && k < v->array.size(); k++)
      if (v->array[k].type == JValue::Number) out[k] = gltf_float(v->array[k]);
    return true;
  };
  auto string_at = [&](const JValue& e, const char* key) -> const std::string* {
    auto v = e.find(key);
    return v && v->type == JValue::String ? &v->string : nullptr;
  };
  const auto dirname = path_dirname(filename);

  // buffers
  const auto& jbuffers = array_of(json, "buffers");
  std::vector<std::vector<uint8_t>> buffers(jbuffers.size());
  std::vector<const uint8_t*>       buffer_data(jbuffers.size(), nullptr);
  std::vector<size_t>               buffer_size(jbuffers.size(), 0);
  for (size_t i = 0; i < jbuffers.size(); i++) {
    const double declared = number_at(jbuffers[i], "byteLength", 0);
    if (declared < 0) return buffers_error();
    const size_t size = (size_t)declared;
    auto         uri  = string_at(jbuffers[i], "uri");
    if (i == 0 && !uri && bin) {
      if (bin_size < size) return buffers_error();
      buffer_data[i] = bin, buffer_size[i] = bin_size;
      continue;
    }
    if (!uri) continue;
    if (uri->compare(0, 5, "data:") == 0) {
      const auto comma = uri->find(',');
      if (comma == std::string::npos || comma < 7 || uri->compare(comma - 7, 7, ";base64") != 0) return buffers_error();
      if ((uri->size() - comma - 1) * 6 / 8 < size || !base64_decode(uri->c_str() + comma + 1, size, buffers[i])) return buffers_error();
    } else if (uri->find("://") == std::string::npos) {
      std::string ignored;
      if (!read_file(path_join(dirname, uri_decode(*uri)), buffers[i], ignored) || buffers[i].size() < size) return buffers_error();
    } else {
      return buffers_error();
    }
    buffer_data[i] = buffers[i].data(), buffer_size[i] = buffers[i].size();
  }
  std::vector<GltfView> views;
  for (auto& jv : array_of(json, "bufferViews")) {
    GltfView view;
    view.buffer = (int)number_at(jv, "buffer", -1);
    auto size_of = [&](const char* key) {
      const double v = number_at(jv, key, 0);
      if (v < 0) sizes_ok = false;
      return v < 0 ? (size_t)0 : (size_t)v;
    };
    view.offset = size_of("byteOffset"), view.size = size_of("byteLength"), view.stride = size_of("byteStride");
    views.push_back(view);
  }
  std::vector<GltfAccessor> accessors;
  for (auto& ja : array_of(json, "accessors")) {
    GltfAccessor a;
    a.view   = (int)number_at(ja, "bufferView", -1);
    auto size_of = [&](const char* key) {
      const double v = number_at(ja, key, 0);
      if (v < 0) sizes_ok = false;
      return v < 0 ? (size_t)0 : (size_t)v;
    };
    a.offset = size_of("byteOffset"), a.count = size_of("count");
    const int ctype = (int)number_at(ja, "componentType", 0);
    a.component     = ctype;  // 5120 i8, 5121 u8, 5122 i16, 5123 u16, 5125 u32, 5126 f32
    if (auto n = ja.find("normalized")) a.normalized = n->type == JValue::Bool && n->boolean;
    a.sparse = ja.find("sparse") != nullptr;
    if (auto t = string_at(ja, "type")) {
      a.components = *t == "SCALAR" ? 1 : *t == "VEC2" ? 2 : *t == "VEC3" ? 3 : *t == "VEC4" ? 4 : *t == "MAT2" ? 4 : *t == "MAT3" ? 9 : *t == "MAT4" ? 16 : 0;
      a.matrix     = (*t)[0] == 'M';
    }
    const size_t csize = ctype == 5120 || ctype == 5121 ? 1 : ctype == 5122 || ctype == 5123 ? 2 : ctype == 5125 || ctype == 5126 ? 4 : 0;
    a.stride           = a.view >= 0 && a.view < (int)views.size() ? views[a.view].stride : 0;
    if (a.stride == 0) a.stride = csize * a.components;  // (matrix alignment rules only matter for matrices, refused below)
    accessors.push_back(a);
  }
  if (!sizes_ok) return parse_error();
  // one element of an accessor, bounds-checked; nullptr = the accessor has no data (reads as zeros, like cgltf)
  auto component_size = [](int ctype) { return ctype == 5120 || ctype == 5121 ? 1 : ctype == 5122 || ctype == 5123 ? 2 : ctype == 5125 || ctype == 5126 ? 4 : 0; };
  auto element_at = [&](const GltfAccessor& a, size_t index, const uint8_t*& ptr) {
    ptr = nullptr;
    if (a.view < 0) return true;
    if (a.view >= (int)views.size()) return false;
    const GltfView& view = views[a.view];
    if (view.buffer < 0 || view.buffer >= (int)buffer_data.size() || !buffer_data[view.buffer]) return false;
    const size_t at = view.offset + a.offset + a.stride * index, need = (size_t)component_size(a.component) * a.components;
    if (at + need > buffer_size[view.buffer]) return false;
    ptr = buffer_data[view.buffer] + at;
    return true;
  };
  auto read_float = [](const uint8_t* in, int ctype, bool normalized) -> float {  // cgltf_component_read_float
    if (ctype == 5126) { float v; memcpy(&v, in, 4); return v; }
    int16_t  s16; uint16_t u16; uint32_t u32v;
    switch (ctype) {
      case 5122: memcpy(&s16, in, 2); return normalized ? s16 / (float)32767 : (float)(size_t)s16;
      case 5123: memcpy(&u16, in, 2); return normalized ? u16 / (float)65535 : (float)(size_t)u16;
      case 5120: return normalized ? (int8_t)in[0] / (float)127 : (float)(size_t)(int8_t)in[0];
      case 5121: return normalized ? in[0] / (float)255 : (float)(size_t)in[0];
      case 5125: memcpy(&u32v, in, 4); return normalized ? 0.0f : (float)(size_t)u32v;
      default: return 0;
    }
  };
  auto read_uint = [](const uint8_t* in, int ctype) -> uint32_t {  // cgltf_component_read_uint
    int16_t s16; uint16_t u16; uint32_t u32v;
    switch (ctype) {
      case 5120: return (uint32_t)(int8_t)in[0];
      case 5121: return in[0];
      case 5122: memcpy(&s16, in, 2); return (uint32_t)s16;
      case 5123: memcpy(&u16, in, 2); return u16;
      case 5125: memcpy(&u32v, in, 4); return u32v;
      default: return 0;
    }
  };

  // cameras (attached to nodes below)
  std::vector<ygl_camera> cameras;
  for (auto& jc : array_of(json, "cameras")) {
    ygl_camera camera = {};
    memcpy(&camera.frame, kIdentityFrame, 48);
    camera.orthographic = 0, camera.lens = 0.050f, camera.film = 0.036f, camera.aspect = 1.500f, camera.focus = 10000, camera.aperture = 0;
    auto type = string_at(jc, "type");
    if (type && *type == "orthographic") {
      float xmag = 0, ymag = 0;
      if (auto o = jc.find("orthographic")) xmag = float_at(*o, "xmag", 0), ymag = float_at(*o, "ymag", 0);
      camera.aspect = xmag / ymag, camera.lens = ymag, camera.film = 0.036f;
    } else if (type && *type == "perspective") {
      float aspect = 0, yfov = 0;
      if (auto o = jc.find("perspective")) aspect = float_at(*o, "aspectRatio", 0.0f), yfov = float_at(*o, "yfov", 0);
      camera.aspect = aspect;
      if (camera.aspect == 0) camera.aspect = 16.0f / 9.0f;
      camera.film = 0.036f;
      if (camera.aspect >= 1) camera.lens = (camera.film / camera.aspect) / (2 * std::tan(yfov / 2));
      else camera.lens = camera.film / (2 * std::tan(yfov / 2));
      camera.focus = 1;
    } else {
      return unsupported("camera type");
    }
    cameras.push_back(camera);
  }
  // images -> textures
  const auto& jimages   = array_of(json, "images");
  const auto& jtextures = array_of(json, "textures");
  std::vector<std::string> texture_files;
  for (auto& ji : jimages) {
    auto uri = string_at(ji, "uri");
    if (!uri) return unsupported("image without uri");  // (the reference dereferences a null uri here)
    std::string path = *uri;
    for (size_t at = 0; (at = path.find("%20", at)) != std::string::npos; at += 1) path.replace(at, 3, " ");
    texture_files.push_back(path);
  }
  auto texture_of = [&](const JValue& owner, const char* key) {  // texture view -> image index
    auto view = owner.find(key);
    if (!view || view->type != JValue::Object) return -1;
    const int t = (int)number_at(*view, "index", -1);
    if (t < 0 || t >= (int)jtextures.size()) return -1;
    const int image = (int)number_at(jtextures[t], "source", -1);
    return image >= 0 && image < (int)jimages.size() ? image : -1;
  };
  // materials
  const auto& jmaterials = array_of(json, "materials");
  for (auto& jm : jmaterials) {
    ygl_material m = {};
    m.type = 7 /* gltfpbr */, m.roughness = 0, m.metallic = 0, m.ior = 1.5f, m.scanisotropy = 0, m.trdepth = 0.01f, m.opacity = 1;
    m.emission_tex = m.color_tex = m.roughness_tex = m.scattering_tex = m.normal_tex = -1;
    floats_at(jm, "emissiveFactor", m.emission, 3);
    const JValue* extensions = jm.find("extensions");
    if (extensions) {
      if (auto strength = extensions->find("KHR_materials_emissive_strength")) {
        const float k = float_at(*strength, "emissiveStrength", 1.0f);
        for (float& e : m.emission) e *= k;
      }
    }
    m.emission_tex = texture_of(jm, "emissiveTexture");
    m.normal_tex   = texture_of(jm, "normalTexture");
    if (auto pbr = jm.find("pbrMetallicRoughness"); pbr && pbr->type == JValue::Object) {
      float base[4] = {1, 1, 1, 1};
      floats_at(*pbr, "baseColorFactor", base, 4);
      m.color[0] = base[0], m.color[1] = base[1], m.color[2] = base[2], m.opacity = base[3];
      m.metallic      = float_at(*pbr, "metallicFactor", 1.0f);
      m.roughness     = float_at(*pbr, "roughnessFactor", 1.0f);
      m.color_tex     = texture_of(*pbr, "baseColorTexture");
      m.roughness_tex = texture_of(*pbr, "metallicRoughnessTexture");
    }
    if (extensions) {
      if (auto transmission = extensions->find("KHR_materials_transmission"); transmission && transmission->type == JValue::Object) {
        const float t = float_at(*transmission, "transmissionFactor", 0.0f);
        if (t > 0) {
          m.type = 3 /* transparent */, m.color[0] = m.color[1] = m.color[2] = t;
          m.color_tex = texture_of(*transmission, "transmissionTexture");
        }
      }
    }
    scene.materials.push_back(m);
  }
  // meshes -> shapes, and per mesh the (shape, material) pairs its nodes instantiate
  std::vector<std::vector<std::pair<int, int>>> mesh_primitives;
  for (auto& jmesh : array_of(json, "meshes")) {
    mesh_primitives.emplace_back();
    for (auto& jp : array_of(jmesh, "primitives")) {
      auto attributes = jp.find("attributes");
      if (!attributes || attributes->type != JValue::Object || attributes->object.empty()) continue;
      scene.shape_data.emplace_back();
      HostShape& shape    = scene.shape_data.back();
      const int  material = (int)number_at(jp, "material", -1);
      mesh_primitives.back().push_back({(int)scene.shape_data.size() - 1, material >= 0 && material < (int)jmaterials.size() ? material : -1});
      for (auto& [name, jacc] : attributes->object) {
        const int id = jacc.type == JValue::Number ? (int)gltf_int(jacc) : -1;
        if (id < 0 || id >= (int)accessors.size()) return parse_error();
        const GltfAccessor& a = accessors[id];
        if (a.sparse) return unsupported("sparse accessor");
        std::vector<float>* target = nullptr;
        size_t              width  = (size_t)a.components;
        if (name == "POSITION") {
          if (a.components != 3) return unsupported("position components");
          target = &shape.positions;
        } else if (name == "NORMAL") {
          if (a.components != 3) return unsupported("normal components");
          target = &shape.normals;
        } else if (name == "TEXCOORD" || name == "TEXCOORD_0") {
          if (a.components != 2) return unsupported("texcoord components");
          target = &shape.texcoords;
        } else if (name == "COLOR" || name == "COLOR_0") {
          if (a.components != 3 && a.components != 4) return unsupported("color components");
          target = &shape.colors, width = 4;
        } else if (name == "TANGENT") {
          if (a.components != 4) return unsupported("tangent components");
          continue;  // tangents are not read on the path (ygl_shape has none)
        } else if (name == "RADIUS") {
          if (a.components != 1) return unsupported("radius components");
          target = &shape.radius;
        } else {
          continue;
        }
        if (a.matrix) return unsupported("accessor float conversion");
        target->assign(a.count * width, 0.0f);
        if (width != (size_t)a.components)
          for (size_t i = 0; i < a.count; i++) (*target)[i * 4 + 3] = 1;
        const size_t csize = (size_t)component_size(a.component);
        for (size_t i = 0; i < a.count; i++) {
          const uint8_t* element = nullptr;
          if (!element_at(a, i, element)) return unsupported("accessor float conversion");
          for (int c = 0; c < a.components; c++)
            (*target)[i * width + c] = element ? read_float(element + csize * c, a.component, a.normalized) : 0.0f;
        }
      }
      const int    mode      = (int)number_at(jp, "mode", 4);
      const int    nverts    = (int)(shape.positions.size() / 3);
      std::vector<int> indices;
      const bool   indexed = jp.find("indices") != nullptr;
      if (indexed) {
        const int id = (int)number_at(jp, "indices", -1);
        if (id < 0 || id >= (int)accessors.size()) return parse_error();
        const GltfAccessor& a = accessors[id];
        if (a.components != 1 || a.matrix) return unsupported("non-scalar indices");
        if (a.sparse) return unsupported("accessor uint conversion");
        indices.resize(a.count);
        for (size_t i = 0; i < a.count; i++) {
          const uint8_t* element = nullptr;
          if (!element_at(a, i, element)) return unsupported("accessor uint conversion");
          indices[i] = element ? (int)read_uint(element, a.component) : 0;
        }
      } else {
        indices.resize(nverts);
        for (int i = 0; i < nverts; i++) indices[i] = i;
      }
      const int n = (int)indices.size();
      auto&     t = shape.triangles;
      auto&     l = shape.lines;
      switch (mode) {
        case 4:
          for (int i = 0; i < n / 3; i++) t.insert(t.end(), {indices[i * 3], indices[i * 3 + 1], indices[i * 3 + 2]});
          break;
        case 6:
          if (n < 2) return unsupported("primitive type");  // (the reference resizes to a negative count)
          for (int i = 2; i < n; i++) t.insert(t.end(), {indices[0], indices[i - 1], indices[i]});
          break;
        case 5:
          if (n < 2) return unsupported("primitive type");
          for (int i = 2; i < n; i++) t.insert(t.end(), {indices[i - 2], indices[i - 1], indices[i]});
          break;
        case 1:
          for (int i = 0; i < n / 2; i++) l.insert(l.end(), {indices[i * 2], indices[i * 2 + 1]});
          break;
        case 2:
          if (n < 1) return unsupported("primitive type");
          if (indexed) {  // {indices[i], indices[i + 1] % n}: the reference reads one index past the end for the last
            return unsupported("indexed line loop");  // segment and reduces the VALUE, not the position; undefined there
          }
          for (int i = 1; i < n; i++) l.insert(l.end(), {i - 1, i});
          l.insert(l.end(), {n - 1, 0});
          break;
        case 3:
          if (n < 1) return unsupported("primitive type");
          for (int i = 1; i < n; i++) l.insert(l.end(), {indices[i - 1], indices[i]});
          break;
        case 0: return unsupported(indexed ? "points primitive" : "point primitive");
        default: return unsupported("primitive type");
      }
    }
  }
  // nodes: parents from the children lists, then cameras and instances in node order
  const auto& jnodes = array_of(json, "nodes");
  std::vector<int> parent(jnodes.size(), -1);
  for (size_t i = 0; i < jnodes.size(); i++)
    for (auto& child : array_of(jnodes[i], "children")) {
      const long long c = child.type == JValue::Number ? gltf_int(child) : -1;
      if (c < 0 || c >= (long long)jnodes.size()) return parse_error();
      parent[c] = (int)i;
    }
  auto local_matrix = [&](const JValue& node, float* lm) {  // cgltf_node_transform_local
    float matrix[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    if (floats_at(node, "matrix", matrix, 16)) {
      memcpy(lm, matrix, sizeof(matrix));
      return;
    }
    float t[3] = {0, 0, 0}, q[4] = {0, 0, 0, 1}, sc[3] = {1, 1, 1};
    floats_at(node, "translation", t, 3), floats_at(node, "rotation", q, 4), floats_at(node, "scale", sc, 3);
    const float qx = q[0], qy = q[1], qz = q[2], qw = q[3], sx = sc[0], sy = sc[1], sz = sc[2];
    lm[0] = (1 - 2 * qy * qy - 2 * qz * qz) * sx, lm[1] = (2 * qx * qy + 2 * qz * qw) * sx, lm[2] = (2 * qx * qz - 2 * qy * qw) * sx, lm[3] = 0.f;
    lm[4] = (2 * qx * qy - 2 * qz * qw) * sy, lm[5] = (1 - 2 * qx * qx - 2 * qz * qz) * sy, lm[6] = (2 * qy * qz + 2 * qx * qw) * sy, lm[7] = 0.f;
    lm[8] = (2 * qx * qz + 2 * qy * qw) * sz, lm[9] = (2 * qy * qz - 2 * qx * qw) * sz, lm[10] = (1 - 2 * qx * qx - 2 * qy * qy) * sz, lm[11] = 0.f;
    lm[12] = t[0], lm[13] = t[1], lm[14] = t[2], lm[15] = 1.f;
  };
  auto world_frame = [&](size_t node, ygl_frame3f& frame) {  // cgltf_node_transform_world, then mat_to_frame
    float lm[16];
    local_matrix(jnodes[node], lm);
    int hops = 0;
    for (int p = parent[node]; p >= 0; p = parent[p]) {
      if (++hops > (int)jnodes.size()) return false;  // a cycle
      float pm[16];
      local_matrix(jnodes[p], pm);
      for (int i = 0; i < 4; ++i) {
        const float l0 = lm[i * 4 + 0], l1 = lm[i * 4 + 1], l2 = lm[i * 4 + 2];
        const float r0 = l0 * pm[0] + l1 * pm[4] + l2 * pm[8], r1 = l0 * pm[1] + l1 * pm[5] + l2 * pm[9], r2 = l0 * pm[2] + l1 * pm[6] + l2 * pm[10];
        lm[i * 4 + 0] = r0, lm[i * 4 + 1] = r1, lm[i * 4 + 2] = r2;
      }
      lm[12] += pm[12], lm[13] += pm[13], lm[14] += pm[14];
    }
    const float f[12] = {lm[0], lm[1], lm[2], lm[4], lm[5], lm[6], lm[8], lm[9], lm[10], lm[12], lm[13], lm[14]};
    memcpy(&frame, f, sizeof(f));
    return true;
  };
  for (size_t i = 0; i < jnodes.size(); i++) {
    if (jnodes[i].find("camera")) {
      const int c = (int)number_at(jnodes[i], "camera", -1);
      if (c < 0 || c >= (int)cameras.size()) return parse_error();
      ygl_camera camera = cameras[c];
      if (!world_frame(i, camera.frame)) return parse_error();
      scene.cameras.push_back(camera);
    }
    if (jnodes[i].find("mesh")) {
      const int mesh = (int)number_at(jnodes[i], "mesh", -1);
      if (mesh < 0 || mesh >= (int)mesh_primitives.size()) return parse_error();
      for (auto& [shape, material] : mesh_primitives[mesh]) {
        ygl_instance inst = {};
        inst.shape = shape, inst.material = material;
        if (!world_frame(i, inst.frame)) return parse_error();
        scene.instances.push_back(inst);
      }
    }
  }
  // textures, fix-ups, sky
  volatile float sun_angle = 3.14159265358979323846f / 4, turbidity = 3;
  HostTexture    sky;
  std::thread    sky_maker([&]() { make_sky_texture(sky, 1024, 512, sun_angle, turbidity, {0.2f, 0.2f, 0.2f}); });
  scene.texture_data.resize(texture_files.size());
  const bool loaded = parallel_load(texture_files.size(), error, [&](size_t i, std::string& err) {
    return load_texture(path_join(dirname, texture_files[i]), scene.texture_data[i], err);
  });
  sky_maker.join();
  if (!loaded) return error = "cannot save " + filename + " since " + error, false;  // (the reference's own wording)
  int default_material = -1;  // add_missing_material, yocto_sceneio.cpp:2151-2163
  for (auto& inst : scene.instances) {
    if (inst.material >= 0) continue;
    if (default_material < 0) {
      ygl_material m = {};
      m.type = 0, m.roughness = 0, m.metallic = 0, m.ior = 1.5f, m.scanisotropy = 0, m.trdepth = 0.01f, m.opacity = 1;
      m.emission_tex = m.color_tex = m.roughness_tex = m.scattering_tex = m.normal_tex = -1;
      m.color[0] = m.color[1] = m.color[2] = 0.8f;
      scene.materials.push_back(m);
      default_material = (int)scene.materials.size() - 1;
    }
    inst.material = default_material;
  }
  add_missing_camera(scene);
  add_missing_radius(scene);
  // add_missing_lights: has_lights (yocto_scene.cpp:678-689) only counts emissive instances whose shape has triangles AND
  // quads, which no glTF shape has, and a glTF file has no environment: the sky is always added
  scene.texture_data.push_back(std::move(sky));
  scene.names[1].push_back("sky");  // (the only name the reference's scene carries: glTF images stay nameless)
  ygl_environment env = {};
  memcpy(&env.frame, kIdentityFrame, 48);
  env.emission[0] = env.emission[1] = env.emission[2] = 1;
  env.emission_tex = (int)scene.texture_data.size() - 1;
  scene.environments.push_back(env);
  scene.names[5].push_back("sky");
  return true;
}

// ---------------------------------------------------------------------------------------------------------------
// Wavefront OBJ as a scene (load_obj_scene, yocto_sceneio.cpp:4111-4243, over load_obj / load_mtl / load_obx of
// yocto_modelio.cpp:949-1376): one shape + instance per (object / group, material) with vertices de-duplicated per
// shape; materials of the .mtl files classified as transparent / reflective / glossy / matte from Kt, Ks, Kd with the
// roughness taken from the Phong exponent; textures by first mention; cameras and environments from the reference's own
// ".obx" side file; a framing camera if there is none. No light is added: such a scene renders with what it emits.
// ---------------------------------------------------------------------------------------------------------------
struct ObjLine {  // one line, comment removed, read token by token like parse_value (yocto_modelio.cpp:376-430)
  const char *s, *stop;
  void skip() { while (s < stop && (*s == ' ' || *s == '\t' || *s == '\r' || *s == '\n')) s++; }
  bool empty() { return skip(), s >= stop; }
  bool word(std::string& out) {
    skip();
    if (s >= stop) return false;
    if (*s != '"') {
      const char* b = s;
      while (s < stop && !(*s == ' ' || *s == '\t' || *s == '\r' || *s == '\n')) s++;
      out.assign(b, s);
      return true;
    }
    const char* b = ++s;
    while (s < stop && *s != '"') s++;
    if (s >= stop) return false;
    out.assign(b, s++);
    return true;
  }
  template <typename T>
  bool number(T& v) {
    skip();
    auto r = std::from_chars(s, stop, v);
    if (r.ptr == s) return false;
    s = r.ptr;
    return true;
  }
  bool numbers(float* v, int n) {
    for (int k = 0; k < n; k++)
      if (!number(v[k])) return false;
    return true;
  }
  bool vertex(ObjVertex& v) {
    v = ObjVertex{};
    if (!number(v.position)) return false;
    if (s < stop && *s == '/') {  // (blanks are skipped before an index, never before a slash)
      s++;
      if (s < stop && *s == '/') {
        s++;
        if (!number(v.normal)) return false;
      } else {
        if (!number(v.texcoord)) return false;
        if (s < stop && *s == '/') {
          s++;
          if (!number(v.normal)) return false;
        }
      }
    }
    return true;
  }
};
template <typename F>
bool for_each_obj_line(const std::vector<uint8_t>& data, F&& fn) {
  const char *p = (const char*)data.data(), *end = p + data.size();
  while (p < end) {
    const char* eol = (const char*)memchr(p, '\n', end - p);
    if (!eol) eol = end;
    const char* stop = (const char*)memchr(p, '#', eol - p);
    ObjLine     line{p, stop ? stop : eol};
    p = eol < end ? eol + 1 : end;
    if (line.empty()) continue;
    std::string cmd;
    if (!line.word(cmd)) return false;
    if (cmd.empty()) continue;
    if (!fn(cmd, line)) return false;
  }
  return true;
}
struct ObjMaterial {
  std::string name;
  float emission[3] = {0, 0, 0}, diffuse[3] = {0, 0, 0}, specular[3] = {0, 0, 0}, transmission[3] = {0, 0, 0};
  float exponent = 10, ior = 1.5f, opacity = 1;
  int   emission_tex = -1, diffuse_tex = -1, specular_tex = -1, transmission_tex = -1, normal_tex = -1;
};
struct ObjShapeData {
  std::vector<ObjVertex> vertices;
  std::vector<int>       sizes, materials;
  std::vector<char>      types;
};
struct ObjModel {
  std::vector<v3>              positions, normals;
  std::vector<v2>              texcoords;
  std::vector<ObjShapeData>    shapes;
  std::vector<ObjMaterial>     materials;
  std::vector<std::string>     textures;
  std::vector<ygl_camera>      cameras;
  std::vector<ygl_environment> environments;
  std::unordered_map<std::string, int> texture_ids;
  bool texture(ObjLine& line, int& id) {  // the last token of the line is the path; -bm / -clamp options are not used here
    std::string token, path;
    while (!line.empty()) {
      if (!line.word(token)) return false;
      path = token;
    }
    if (path.empty() && token.empty()) return false;
    for (auto& c : path)
      if (c == '\\') c = '/';
    auto it = texture_ids.find(path);
    if (it == texture_ids.end()) {
      textures.push_back(path);
      it = texture_ids.insert({path, (int)textures.size() - 1}).first;
    }
    id = it->second;
    return true;
  }
};
bool load_mtl(const std::string& filename, ObjModel& obj, std::string& error) {
  std::vector<uint8_t> data;
  if (!read_file(filename, data, error)) return false;
  const size_t first = obj.materials.size();
  obj.materials.emplace_back();  // a placeholder that swallows what comes before the first newmtl
  int  unused_tex = -1;
  float unused[3];
  const bool ok = for_each_obj_line(data, [&](const std::string& cmd, ObjLine& line) {
    ObjMaterial& m = obj.materials.back();
    if (cmd == "newmtl") {
      obj.materials.emplace_back();
      return line.word(obj.materials.back().name);
    }
    if (cmd == "illum") { int illum; return line.number(illum); }
    if (cmd == "Ke") return line.numbers(m.emission, 3);
    if (cmd == "Ka") return line.numbers(unused, 3);
    if (cmd == "Kd") return line.numbers(m.diffuse, 3);
    if (cmd == "Ks") return line.numbers(m.specular, 3);
    if (cmd == "Kt") return line.numbers(m.transmission, 3);
    if (cmd == "Tf") {
      if (!line.numbers(m.transmission, 3)) return false;
      for (float& t : m.transmission) t = std::max(1 - t, 0.0f);
      if (std::max(m.transmission[0], std::max(m.transmission[1], m.transmission[2])) < 0.001f) m.transmission[0] = m.transmission[1] = m.transmission[2] = 0;
      return true;
    }
    if (cmd == "Tr") {
      if (!line.number(m.opacity)) return false;
      m.opacity = 1 - m.opacity;
      return true;
    }
    if (cmd == "Ns") return line.number(m.exponent);
    if (cmd == "d") return line.number(m.opacity);
    if (cmd == "map_Ke") return obj.texture(line, m.emission_tex);
    if (cmd == "map_Kd") return obj.texture(line, m.diffuse_tex);
    if (cmd == "map_Ks") return obj.texture(line, m.specular_tex);
    if (cmd == "map_Tr") return obj.texture(line, m.transmission_tex);
    if (cmd == "map_norm" || cmd == "norm") return obj.texture(line, m.normal_tex);
    if (cmd == "map_Ka" || cmd == "map_d" || cmd == "map_bump" || cmd == "bump" || cmd == "map_disp" || cmd == "disp")
      return obj.texture(line, unused_tex);  // registered in the texture list like there, not used by the conversion
    return true;
  });
  if (!ok) return error = "cannot parse " + filename, false;
  // "remove placeholder material": the reference erases the FIRST material of the model, which is the placeholder only
  // when nothing was there before - after elements without usemtl (the grey default) or a second mtllib it drops that
  // earlier material and the empty placeholder stays. Kept, so that material indices come out the same.
  (void)first;
  obj.materials.erase(obj.materials.begin());
  return true;
}
bool load_obx(const std::string& filename, ObjModel& obj, std::string& error) {
  std::vector<uint8_t> data;
  if (!read_file(filename, data, error)) return false;
  auto fresh_camera = []() {
    ygl_camera c = {};
    memcpy(&c.frame, kIdentityFrame, 48);
    c.orthographic = 0, c.aspect = 16.0f / 9.0f, c.lens = 0.50f, c.film = 0.036f, c.focus = 0, c.aperture = 0;
    return c;
  };
  auto fresh_environment = []() {
    ygl_environment e = {};
    memcpy(&e.frame, kIdentityFrame, 48);
    e.emission_tex = -1;
    return e;
  };
  obj.cameras.push_back(fresh_camera()), obj.environments.push_back(fresh_environment());  // placeholders
  const bool ok = for_each_obj_line(data, [&](const std::string& cmd, ObjLine& line) {
    ygl_camera&      camera = obj.cameras.back();
    ygl_environment& env    = obj.environments.back();
    std::string      name;
    float            l[9];
    if (cmd == "newCam") return obj.cameras.push_back(fresh_camera()), line.word(name);
    if (cmd == "Co") { int v; if (!line.number(v)) return false; camera.orthographic = v != 0; return true; }
    if (cmd == "Ca") return line.number(camera.aspect);
    if (cmd == "Cl") return line.number(camera.lens);
    if (cmd == "Cs") return line.number(camera.film);
    if (cmd == "Cf") return line.number(camera.focus);
    if (cmd == "Cp") return line.number(camera.aperture);
    if (cmd == "Cx") return line.numbers(frame_of(camera.frame), 12);
    if (cmd == "Ct") {
      if (!line.numbers(l, 9)) return false;
      lookat_frame(frame_of(camera.frame), {l[0], l[1], l[2]}, {l[3], l[4], l[5]}, {l[6], l[7], l[8]}, false);
      if (camera.focus == 0) camera.focus = length(v3{l[3] - l[0], l[4] - l[1], l[5] - l[2]});
      return true;
    }
    if (cmd == "newEnv") return obj.environments.push_back(fresh_environment()), line.word(name);
    if (cmd == "Ee") return line.numbers(env.emission, 3);
    if (cmd == "map_Ee") return obj.texture(line, env.emission_tex);
    if (cmd == "Ex") return line.numbers(frame_of(env.frame), 12);
    if (cmd == "Et") {
      if (!line.numbers(l, 9)) return false;
      lookat_frame(frame_of(env.frame), {l[0], l[1], l[2]}, {l[3], l[4], l[5]}, {l[6], l[7], l[8]}, true);
      return true;
    }
    return true;
  });
  if (!ok) return error = "cannot parse " + filename, false;
  obj.cameras.erase(obj.cameras.begin()), obj.environments.erase(obj.environments.begin());
  return true;
}
bool load_obj_model(const std::string& filename, ObjModel& obj, std::string& error) {
  std::vector<uint8_t> data;
  if (!read_file(filename, data, error)) return false;
  std::string oname, gname;
  std::vector<std::string>             mtllibs;
  std::unordered_map<std::string, int> material_ids;
  int  cur_material = -1;
  std::unordered_map<int, int> cur_shapes = {{-1, 0}};  // material -> shape, within the current object / group
  obj.shapes.emplace_back();
  int         cur_shape = 0;
  std::string failure;
  const bool  ok = for_each_obj_line(data, [&](const std::string& cmd, ObjLine& line) {
    if (cmd == "v" || cmd == "vn") {
      v3 v = {0, 0, 0};
      (cmd == "v" ? obj.positions : obj.normals).push_back(v);  // (the slot exists even if the line then fails)
      return line.numbers(&(cmd == "v" ? obj.positions : obj.normals).back().x, 3);
    }
    if (cmd == "vt") {
      obj.texcoords.push_back({0, 0});
      return line.numbers(&obj.texcoords.back().x, 2);
    }
    if (cmd == "f" || cmd == "l" || cmd == "p") {
      if (cur_material < 0) {  // elements before any usemtl get a grey default material
        ObjMaterial m;
        m.name = "__default__", m.diffuse[0] = m.diffuse[1] = m.diffuse[2] = 0.8f;
        obj.materials.push_back(m);
        cur_material = 0, material_ids[m.name] = 0;
      }
      ObjShapeData& shape = obj.shapes[cur_shape];
      shape.sizes.push_back(0), shape.types.push_back(cmd[0]), shape.materials.push_back(cur_material);
      while (!line.empty()) {
        ObjVertex v;
        if (!line.vertex(v)) return false;
        if (v.position == 0) break;
        if (v.position < 0) v.position = (int)obj.positions.size() + v.position + 1;
        if (v.texcoord < 0) v.texcoord = (int)obj.texcoords.size() + v.texcoord + 1;
        if (v.normal < 0) v.normal = (int)obj.normals.size() + v.normal + 1;
        shape.vertices.push_back(v);
        shape.sizes.back()++;
      }
      return true;
    }
    if (cmd == "o" || cmd == "g") {
      std::string& name = cmd == "o" ? oname : gname;
      name.clear();
      if (!line.empty() && !line.word(name)) return false;
      obj.shapes.emplace_back();
      cur_shape  = (int)obj.shapes.size() - 1;
      cur_shapes = {{cur_material, cur_shape}};
      return true;
    }
    if (cmd == "usemtl") {
      std::string name;
      if (!line.word(name)) return false;
      auto it = material_ids.find(name);
      if (it == material_ids.end()) return false;
      if (cur_material != it->second) {
        cur_material = it->second;
        auto shape   = cur_shapes.find(cur_material);
        if (shape == cur_shapes.end()) {
          obj.shapes.emplace_back();
          cur_shape = (int)obj.shapes.size() - 1, cur_shapes[cur_material] = cur_shape;
        } else {
          cur_shape = shape->second;
        }
      }
      return true;
    }
    if (cmd == "mtllib") {
      std::string name;
      if (!line.word(name)) return false;
      if (std::find(mtllibs.begin(), mtllibs.end(), name) != mtllibs.end()) return true;
      mtllibs.push_back(name);
      if (!load_mtl(path_join(path_dirname(filename), name), obj, failure)) return false;
      for (size_t k = 0; k < obj.materials.size(); k++) material_ids[obj.materials[k].name] = (int)k;
      return true;
    }
    return true;
  });
  if (!ok) return error = failure.empty() ? "cannot parse " + filename : "cannot load " + filename + " since " + failure, false;
  obj.shapes.erase(std::remove_if(obj.shapes.begin(), obj.shapes.end(), [](const ObjShapeData& shape) { return shape.sizes.empty(); }),
      obj.shapes.end());
  auto dot = filename.find_last_of('.');
  const std::string side = filename.substr(0, dot) + ".obx";
  if (path_exists(side) && !load_obx(side, obj, failure)) return error = "cannot load " + filename + " since " + failure, false;
  return true;
}
bool load_obj_scene(const std::string& filename, ygl_loaded_scene& scene, std::string& error) {
  ObjModel obj;
  if (!load_obj_model(filename, obj, error)) return false;
  scene.cameras = obj.cameras;
  auto exponent_to_roughness = [](float exponent) {
    if (exponent >= 1000) return 0.0f;
    float roughness = exponent;
    roughness       = std::pow(2 / (roughness + 2), 1 / 4.0f);
    if (roughness < 0.01f) roughness = 0;
    if (roughness > 0.99f) roughness = 1;
    return roughness;
  };
  auto largest = [](const float* v) { return (double)std::max(v[0], std::max(v[1], v[2])); };
  for (auto& om : obj.materials) {
    ygl_material m = {};
    m.type = 7, m.roughness = 0, m.metallic = 0, m.ior = 1.5f, m.scanisotropy = 0, m.trdepth = 0.01f, m.opacity = 1;
    m.emission_tex = m.color_tex = m.roughness_tex = m.scattering_tex = m.normal_tex = -1;
    memcpy(m.emission, om.emission, 12);
    m.emission_tex = om.emission_tex;
    if (largest(om.transmission) > 0.1) m.type = 3, memcpy(m.color, om.transmission, 12), m.color_tex = om.transmission_tex;
    else if (largest(om.specular) > 0.2) m.type = 2, memcpy(m.color, om.specular, 12), m.color_tex = om.specular_tex;
    else if (largest(om.specular) > 0) m.type = 1, memcpy(m.color, om.diffuse, 12), m.color_tex = om.diffuse_tex;
    else m.type = 0, memcpy(m.color, om.diffuse, 12), m.color_tex = om.diffuse_tex;
    m.roughness = exponent_to_roughness(om.exponent), m.ior = om.ior, m.metallic = 0, m.opacity = om.opacity, m.normal_tex = om.normal_tex;
    scene.materials.push_back(m);
  }
  for (auto& shape : obj.shapes) {
    scene.shape_data.emplace_back();
    ygl_instance inst = {};
    memcpy(&inst.frame, kIdentityFrame, 48);
    inst.shape = (int)scene.shape_data.size() - 1, inst.material = shape.materials.front();
    if (!convert_obj_shape(obj.positions, obj.normals, obj.texcoords, shape.vertices, shape.sizes, shape.types, &shape.materials,
            inst.material, scene.shape_data.back()))
      return error = "cannot parse " + filename + ": vertex index out of range", false;
    scene.instances.push_back(inst);
  }
  scene.environments = obj.environments;
  auto numbered = [](size_t count, const char* prefix, std::vector<std::string>& names) {  // make_names, :2104-2116
    const std::string width = std::to_string(count);
    for (size_t k = 0; k < count; k++) {
      std::string number = std::to_string(k + 1);
      while (number.size() < width.size()) number = "0" + number;
      names.push_back(prefix + number);
    }
  };
  numbered(scene.cameras.size(), "camera", scene.names[0]), numbered(obj.textures.size(), "texture", scene.names[1]);
  numbered(scene.materials.size(), "material", scene.names[2]), numbered(scene.shape_data.size(), "shape", scene.names[3]);
  numbered(scene.instances.size(), "instance", scene.names[4]);
  scene.texture_data.resize(obj.textures.size());
  const auto dirname = path_dirname(filename);
  if (!parallel_load(obj.textures.size(), error, [&](size_t i, std::string& err) {
        return load_texture(path_join(dirname, obj.textures[i]), scene.texture_data[i], err);
      }))
    return error = "cannot load " + filename + " since " + error, false;
  add_missing_camera(scene);
  add_missing_radius(scene);
  return true;
}

void make_desc(ygl_loaded_scene& scene) {
  scene.shapes.clear(), scene.textures.clear();
  for (auto& s : scene.shape_data) {
    ygl_shape v     = {};
    v.num_points    = (int)s.points.size();
    v.num_lines     = (int)s.lines.size() / 2;
    v.num_triangles = (int)s.triangles.size() / 3;
    v.num_quads     = (int)s.quads.size() / 4;
    v.points = s.points.data(), v.lines = s.lines.data(), v.triangles = s.triangles.data(), v.quads = s.quads.data();
    v.num_positions = (int)s.positions.size() / 3, v.num_normals = (int)s.normals.size() / 3;
    v.num_texcoords = (int)s.texcoords.size() / 2, v.num_colors = (int)s.colors.size() / 4;
    v.num_radius    = (int)s.radius.size();
    v.positions = s.positions.data(), v.normals = s.normals.data(), v.texcoords = s.texcoords.data();
    v.colors = s.colors.data(), v.radius = s.radius.data();
    scene.shapes.push_back(v);
  }
  for (auto& t : scene.texture_data) {
    ygl_texture v = {};
    v.width = t.width, v.height = t.height, v.linear = t.linear, v.nearest = t.nearest, v.clamp = t.clamp;
    v.pixelsf = t.pixelsf.empty() ? nullptr : t.pixelsf.data();
    v.pixelsb = t.pixelsb.empty() ? nullptr : t.pixelsb.data();
    scene.textures.push_back(v);
  }
  auto& d = scene.desc;
  d                  = {};
  d.num_cameras      = (int)scene.cameras.size();
  d.num_instances    = (int)scene.instances.size();
  d.num_environments = (int)scene.environments.size();
  d.num_shapes       = (int)scene.shapes.size();
  d.num_textures     = (int)scene.textures.size();
  d.num_materials    = (int)scene.materials.size();
  d.cameras = scene.cameras.data(), d.instances = scene.instances.data(), d.environments = scene.environments.data();
  d.shapes = scene.shapes.data(), d.textures = scene.textures.data(), d.materials = scene.materials.data();
}

}  // namespace

extern "C" void ygl_internal_set_error(const char* message);

extern "C" {

int ygl_scene_load(const char* filename, ygl_loaded_scene** out) {
  if (!filename || !out) return ygl_internal_set_error("null argument"), YGL_ERR_INVALID;
  try {
    auto        scene = std::make_unique<ygl_loaded_scene>();
    std::string error;
    // load_scene, yocto_sceneio.cpp:2761-2782: the extension picks the format
    const auto ext = path_extension(filename);
    bool       ok  = false;
    if (ext == ".json") ok = load_json_scene(filename, *scene, error);
    else if (ext == ".ply") ok = load_ply_scene(filename, *scene, error);
    else if (ext == ".gltf" || ext == ".glb") ok = load_gltf_scene(filename, *scene, error);
    else if (ext == ".obj") ok = load_obj_scene(filename, *scene, error);
    else error = "unsupported format " + std::string(filename), ok = false;
    if (!ok) return ygl_internal_set_error(error.c_str()), YGL_ERR_RUNTIME;
    make_desc(*scene);
    *out = scene.release();
    return YGL_OK;
  } catch (const std::exception& e) {  // a file that asks for more memory than there is, ...
    return ygl_internal_set_error((std::string("cannot load ") + filename + ": " + e.what()).c_str()), YGL_ERR_RUNTIME;
  }
}
const ygl_scene_desc* ygl_loaded_scene_desc(const ygl_loaded_scene* scene) { return scene ? &scene->desc : nullptr; }
const char* ygl_loaded_scene_name(const ygl_loaded_scene* scene, int kind, int index) {
  if (!scene || kind < 0 || kind >= 6 || index < 0 || index >= (int)scene->names[kind].size()) return nullptr;
  return scene->names[kind][index].c_str();
}
void ygl_loaded_scene_destroy(ygl_loaded_scene* scene) { delete scene; }

}  // extern "C"
