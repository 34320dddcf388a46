// Triangle meshes for the scene front end: OBJ reader, procedural stand-ins, bounding-box rescale
// and the generalized winding number used for inside tests.
//
// Reference behaviour: Simulation::loadMesh (src/mpm.cu:331-346: igl::readOBJ into a float vertex
// matrix, scale the longest bounding-box edge to `size`, move the lowest corner to `position`) and
// igl::winding_number (include/igl/winding_number.cpp:41-54, include/igl/solid_angle.cpp:12-55):
// the sum over faces of the signed solid angle / 2 pi, evaluated in float.  libigl sums through an
// AABB hierarchy; here the sum runs over the faces in file order, so values agree up to float
// summation order ("parity unpinned" for the front end).
//
// The scene meshes of the reference checkout are Git-LFS pointer stubs.  When an
// .obj file is such a stub (or missing), a declared procedural stand-in is used instead:
// sphere.obj -> UV sphere, cube.obj -> cube, rubber_duck.obj -> ellipsoid, stanford_bunny.obj ->
// torus (any other name -> UV sphere).  All stand-ins are single closed non-self-intersecting
// surfaces, so the winding number is 0 or 1.
#pragma once
#include <charconv>
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

namespace mpmh {

struct TriMesh {
  std::vector<float> V;  // n x 3
  std::vector<int> F;    // m x 3, 0-based
  size_t n_vertices() const { return V.size() / 3; }
  size_t n_faces() const { return F.size() / 3; }
};

inline bool is_lfs_stub(const std::string& path) {
  std::ifstream f(path);
  if (!f) return false;
  std::string first;
  std::getline(f, first);
  return first.rfind("version https://git-lfs", 0) == 0;
}
// the four meshes the reference's scenes name (all Git-LFS pointers in its checkout)
inline bool is_reference_mesh_name(const std::string& filename) {
  return filename == "sphere.obj" || filename == "cube.obj" || filename == "rubber_duck.obj" || filename == "stanford_bunny.obj";
}

// `v x y z [w]`, `f` corners as i, i/t, i/t/n or i//n (1-based; negative = relative to the end);
// vt/vn and everything else is skipped.  Faces with more than 3 corners are fan-triangulated.
inline bool read_obj(const std::string& path, TriMesh& m, std::string* err = nullptr) {
  std::ifstream f(path);
  if (!f) {
    if (err) *err = "cannot open '" + path + "'";
    return false;
  }
  m = TriMesh{};
  std::string line;
  int lineno = 0;
  while (std::getline(f, line)) {
    ++lineno;
    std::istringstream ss(line);
    std::string tag;
    if (!(ss >> tag)) continue;
    if (tag == "v") {
      float x, y, z;
      if (!(ss >> x >> y >> z)) {
        if (err) *err = path + ":" + std::to_string(lineno) + ": vertex needs 3 coordinates";
        return false;
      }
      m.V.insert(m.V.end(), {x, y, z});
    } else if (tag == "f") {
      std::vector<int> corners;
      std::string tok;
      while (ss >> tok) {
        const long idx = std::strtol(tok.c_str(), nullptr, 10);  // stops at '/'
        if (idx == 0) {
          if (err) *err = path + ":" + std::to_string(lineno) + ": bad face corner '" + tok + "'";
          return false;
        }
        corners.push_back(idx > 0 ? (int)idx - 1 : (int)m.n_vertices() + (int)idx);
      }
      if (corners.size() < 3) {
        if (err) *err = path + ":" + std::to_string(lineno) + ": face needs at least 3 corners";
        return false;
      }
      for (size_t c = 1; c + 1 < corners.size(); ++c) m.F.insert(m.F.end(), {corners[0], corners[c], corners[c + 1]});
    }
  }
  for (int v : m.F)
    if (v < 0 || (size_t)v >= m.n_vertices()) {
      if (err) *err = path + ": face index out of range";
      return false;
    }
  return true;
}

// OBJ text as igl::writeOBJ emits it (include/igl/writeOBJ.cpp:42-47, 78-94): %0.17g vertices, 1-based faces
// "v %0.17g %0.17g %0.17g" / "f %d %d %d" (1-based) lines; std::to_chars(general, 17) produces the characters of
// %.17g, chunks are formatted in parallel and written in order
inline bool write_obj(const std::string& path, const std::vector<double>& V, const std::vector<int>& F) {
  FILE* f = std::fopen(path.c_str(), "w");
  if (!f) return false;
  const size_t nv = V.size() / 3, nf = F.size() / 3, chunk = 1u << 13;
  const size_t cv = (nv + chunk - 1) / chunk, cf = (nf + chunk - 1) / chunk;
  std::vector<std::string> parts(cv + cf);
#pragma omp parallel for schedule(dynamic)
  for (long long c = 0; c < (long long)(cv + cf); ++c) {
    std::string& out = parts[(size_t)c];
    char buf[40];
    if ((size_t)c < cv) {
      const size_t b = (size_t)c * chunk, e = std::min(nv, b + chunk);
      out.reserve((e - b) * 72);
      for (size_t i = b; i < e; ++i) {
        out += 'v';
        for (int d = 0; d < 3; ++d) {
          out += ' ';
          const auto r = std::to_chars(buf, buf + sizeof(buf), V[3 * i + d], std::chars_format::general, 17);
          out.append(buf, r.ptr);
        }
        out += '\n';
      }
    } else {
      const size_t b = ((size_t)c - cv) * chunk, e = std::min(nf, b + chunk);
      out.reserve((e - b) * 28);
      for (size_t i = b; i < e; ++i) {
        out += 'f';
        for (int d = 0; d < 3; ++d) {
          out += ' ';
          const auto r = std::to_chars(buf, buf + sizeof(buf), F[3 * i + d] + 1);
          out.append(buf, r.ptr);
        }
        out += '\n';
      }
    }
  }
  bool ok = true;
  for (const std::string& part : parts) ok = ok && std::fwrite(part.data(), 1, part.size(), f) == part.size();
  return std::fclose(f) == 0 && ok;
}

// ---- procedural stand-ins (outward-facing triangles) ----
inline TriMesh make_ellipsoid(float ax, float ay, float az, int n_lat = 32, int n_lon = 64) {
  TriMesh m;
  const double pi = 3.14159265358979323846;
  m.V.insert(m.V.end(), {0.f, ay, 0.f});  // north pole
  for (int i = 1; i < n_lat; ++i) {
    const double th = pi * i / n_lat;
    for (int j = 0; j < n_lon; ++j) {
      const double ph = 2 * pi * j / n_lon;
      m.V.insert(m.V.end(), {(float)(ax * std::sin(th) * std::cos(ph)), (float)(ay * std::cos(th)), (float)(az * std::sin(th) * std::sin(ph))});
    }
  }
  m.V.insert(m.V.end(), {0.f, -ay, 0.f});  // south pole
  const int south = (int)m.n_vertices() - 1;
  auto ring = [&](int i, int j) { return 1 + (i - 1) * n_lon + (j % n_lon); };
  for (int j = 0; j < n_lon; ++j) {
    m.F.insert(m.F.end(), {0, ring(1, j + 1), ring(1, j)});
    m.F.insert(m.F.end(), {south, ring(n_lat - 1, j), ring(n_lat - 1, j + 1)});
  }
  for (int i = 1; i + 1 < n_lat; ++i)
    for (int j = 0; j < n_lon; ++j) {
      m.F.insert(m.F.end(), {ring(i, j), ring(i, j + 1), ring(i + 1, j + 1)});
      m.F.insert(m.F.end(), {ring(i, j), ring(i + 1, j + 1), ring(i + 1, j)});
    }
  return m;
}

inline TriMesh make_cube() {
  TriMesh m;
  for (int i = 0; i < 8; ++i) m.V.insert(m.V.end(), {(float)(i & 1), (float)((i >> 1) & 1), (float)((i >> 2) & 1)});
  const int q[6][4] = {{0, 2, 3, 1}, {4, 5, 7, 6}, {0, 1, 5, 4}, {2, 6, 7, 3}, {0, 4, 6, 2}, {1, 3, 7, 5}};
  for (auto& f : q) {
    m.F.insert(m.F.end(), {f[0], f[1], f[2]});
    m.F.insert(m.F.end(), {f[0], f[2], f[3]});
  }
  return m;
}

inline TriMesh make_torus(float R = 1.0f, float r = 0.45f, int n_major = 64, int n_minor = 32) {
  TriMesh m;
  const double pi = 3.14159265358979323846;
  for (int i = 0; i < n_major; ++i) {
    const double u = 2 * pi * i / n_major;
    for (int j = 0; j < n_minor; ++j) {
      const double v = 2 * pi * j / n_minor;
      m.V.insert(m.V.end(), {(float)((R + r * std::cos(v)) * std::cos(u)), (float)(r * std::sin(v)), (float)((R + r * std::cos(v)) * std::sin(u))});
    }
  }
  auto id = [&](int i, int j) { return (i % n_major) * n_minor + (j % n_minor); };
  for (int i = 0; i < n_major; ++i)
    for (int j = 0; j < n_minor; ++j) {
      m.F.insert(m.F.end(), {id(i, j), id(i, j + 1), id(i + 1, j + 1)});
      m.F.insert(m.F.end(), {id(i, j), id(i + 1, j + 1), id(i + 1, j)});
    }
  return m;
}

inline TriMesh stand_in_for(const std::string& filename) {
  if (filename.find("cube") != std::string::npos) return make_cube();
  if (filename.find("duck") != std::string::npos) return make_ellipsoid(1.0f, 0.8f, 0.6f);
  if (filename.find("bunny") != std::string::npos) return make_torus();
  return make_ellipsoid(1.0f, 1.0f, 1.0f);
}

// The file's mesh, or a procedural stand-in when the file is a Git-LFS pointer, or when it is absent
// AND carries the name of one of the reference's own meshes (whose geometry this repository cannot
// ship); any other missing or unreadable file is an error.  *substituted says which.
inline bool load_mesh_or_stand_in(const std::string& path, TriMesh& m, bool* substituted, std::string* err) {
  const size_t slash = path.find_last_of('/');
  const std::string name = slash == std::string::npos ? path : path.substr(slash + 1);
  const bool missing = !std::ifstream(path).good();
  if (is_lfs_stub(path) || (missing && is_reference_mesh_name(name))) {
    m = stand_in_for(name);
    if (substituted) *substituted = true;
    return true;
  }
  if (substituted) *substituted = false;
  return read_obj(path, m, err);
}

// Simulation::loadMesh (src/mpm.cu:331-346): float arithmetic on the vertex matrix, `size` double
inline void rescale_mesh(TriMesh& m, double size, const float position[3]) {
  float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
  for (size_t i = 0; i < m.n_vertices(); ++i)
    for (int d = 0; d < 3; ++d) {
      mn[d] = std::min(mn[d], m.V[3 * i + d]);
      mx[d] = std::max(mx[d], m.V[3 * i + d]);
    }
  const float length_max = std::max({mx[0] - mn[0], mx[1] - mn[1], mx[2] - mn[2]});
  const float scale = (float)(size / (double)length_max);
  float shift[3];
  for (int d = 0; d < 3; ++d) shift[d] = position[d] - scale * mn[d];
  for (size_t i = 0; i < m.n_vertices(); ++i)
    for (int d = 0; d < 3; ++d) {
      float v = m.V[3 * i + d];
      v *= scale;
      v += shift[d];
      m.V[3 * i + d] = v;
    }
}

// igl::solid_angle / (2 pi) for one triangle seen from p, float like the reference's instantiation
// (vertex matrix and query points are float); atan2 and the division run in double and the result
// is rounded to float on return.
inline float solid_angle_2pi(const float* A, const float* B, const float* C, const float* P) {
  float v[3][3];
  for (int d = 0; d < 3; ++d) {
    v[0][d] = A[d] - P[d];
    v[1][d] = B[d] - P[d];
    v[2][d] = C[d] - P[d];
  }
  float vl[3];
  for (int r = 0; r < 3; ++r) vl[r] = std::sqrt(v[r][0] * v[r][0] + v[r][1] * v[r][1] + v[r][2] * v[r][2]);
  const float detf = v[0][0] * v[1][1] * v[2][2] + v[1][0] * v[2][1] * v[0][2] + v[2][0] * v[0][1] * v[1][2] -
                     v[2][0] * v[1][1] * v[0][2] - v[1][0] * v[0][1] * v[2][2] - v[0][0] * v[2][1] * v[1][2];
  float dp[3];
  dp[0] = v[1][0] * v[2][0];
  dp[0] += v[1][1] * v[2][1];
  dp[0] += v[1][2] * v[2][2];
  dp[1] = v[2][0] * v[0][0];
  dp[1] += v[2][1] * v[0][1];
  dp[1] += v[2][2] * v[0][2];
  dp[2] = v[0][0] * v[1][0];
  dp[2] += v[0][1] * v[1][1];
  dp[2] += v[0][2] * v[1][2];
  const float den = vl[0] * vl[1] * vl[2] + dp[0] * vl[0] + dp[1] * vl[1] + dp[2] * vl[2];
  return (float)(std::atan2((double)detf, (double)den) / (2. * 3.1415926535897932384626433832795));
}

inline float winding_number(const TriMesh& m, const float* p) {
  float w = 0.0f;
  for (size_t f = 0; f < m.n_faces(); ++f)
    w += solid_angle_2pi(&m.V[3 * m.F[3 * f]], &m.V[3 * m.F[3 * f + 1]], &m.V[3 * m.F[3 * f + 2]], p);
  return w;
}

}  // namespace mpmh
