// Output of the scene front end: particle dumps and the surface mesh.
// Reference behaviour: ParticleWriter (include/particle_writer.h:14-44: attributes id, position,
// velocity, radius = 0.1 through Partio) and MeshBuilder::computeMesh (include/mesh_builder.h:165-211:
// min-distance splat capped at (r+1) voxel_dx, scalar field r voxel_dx - d, iso-surface at 0 on a
// mesh_grid^3 lattice, vertices scaled by voxel_dx, igl::writeOBJ).
//
// Partio and libigl are not available to this build, so:
//   * particles are written in Partio's own file formats, chosen by the file extension like
//     Partio::write does: .bgeo (the reference's choice, src/main.cu:109: Houdini's classic binary
//     geometry, version 5, big-endian: header, attribute dictionary, one (x, y, z, 1) + attributes record
//     per point, one particle-system primitive) and .pda (Partio's ASCII table).  Same four attributes
//     in the same order.  Written from the format description; with Partio absent the bytes are not
//     pinned against the reference's output.
//   * the iso-surface is extracted by marching tetrahedra (six tetrahedra per lattice cube around
//     the 0-6 diagonal) instead of libigl's GPL marching-cubes tables: same field, same iso-level,
//     same lattice, a different (finer) triangulation.  The reference has no golden meshes; the
//     tests compare enclosed volume and closedness.
//   * --laplacian_smooth and --mesh-face-count (mesh_builder.h:196-211) are mesh_ops.hpp.
#pragma once
#ifdef _OPENMP
#include <omp.h>
#endif
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <iostream>
#include <string>
#include <unordered_map>
#include <vector>

#include "mesh.hpp"
#include "mesh_ops.hpp"
#include "simulation.hpp"

namespace mpmh {

class ParticleWriter {
 public:
  // format by extension, like Partio::write: .bgeo (binary) or anything else = .pda (ASCII)
  bool writeParticles(const std::string& filepath, const std::vector<Particle>& particles) const {
    const bool bgeo = filepath.size() >= 5 && filepath.compare(filepath.size() - 5, 5, ".bgeo") == 0;
    FILE* f = std::fopen(filepath.c_str(), bgeo ? "wb" : "w");
    if (!f) {
      std::cout << "Warning: Particles could not be written to " << filepath << std::endl;
      return false;
    }
    if (bgeo) write_bgeo(f, particles); else write_pda(f, particles);
    std::fclose(f);
    return true;
  }

 private:
  static void write_pda(FILE* f, const std::vector<Particle>& particles) {
    std::fprintf(f, "ATTRIBUTES\n id position velocity radius\nTYPES\n I V V R\nNUMBER_OF_PARTICLES: %zu\nBEGIN DATA\n", particles.size());
    int i = 0;
    for (const Particle& p : particles) {
      std::fprintf(f, "%d %.9g %.9g %.9g %.9g %.9g %.9g %.9g\n", i, p.x[0], p.x[1], p.x[2], p.v[0], p.v[1], p.v[2], 0.1f);
      i++;
    }
  }
  // big-endian scalars
  static void be32(FILE* f, uint32_t v) {
    const unsigned char b[4] = {(unsigned char)(v >> 24), (unsigned char)(v >> 16), (unsigned char)(v >> 8), (unsigned char)v};
    std::fwrite(b, 1, 4, f);
  }
  static void be16(FILE* f, uint16_t v) {
    const unsigned char b[2] = {(unsigned char)(v >> 8), (unsigned char)v};
    std::fwrite(b, 1, 2, f);
  }
  static void bef(FILE* f, float v) {
    uint32_t u;
    std::memcpy(&u, &v, 4);
    be32(f, u);
  }
  static void hstr(FILE* f, const char* s) {
    const size_t n = std::strlen(s);
    be16(f, (uint16_t)n);
    std::fwrite(s, 1, n, f);
  }
  // attribute dictionary entry: name, size, Houdini type (0 float, 1 int, 4 index, 5 vector), defaults
  static void attr_def(FILE* f, const char* name, uint16_t size, uint32_t type) {
    hstr(f, name);
    be16(f, size);
    be32(f, type);
    for (uint16_t i = 0; i < size; ++i) be32(f, 0);
  }
  static void write_bgeo(FILE* f, const std::vector<Particle>& particles) {
    const uint32_t n = (uint32_t)particles.size();
    std::fwrite("Bgeo", 1, 4, f);
    std::fputc('V', f);
    be32(f, 5);  // version
    be32(f, n);  // points
    be32(f, 1);  // primitives: one particle system
    be32(f, 0);  // point groups
    be32(f, 0);  // primitive groups
    be32(f, 3);  // point attributes besides the position: id, velocity, radius
    be32(f, 0);  // vertex attributes
    be32(f, 1);  // primitive attributes: generator
    be32(f, 0);  // detail attributes
    attr_def(f, "id", 1, 1);
    attr_def(f, "velocity", 3, 5);
    attr_def(f, "radius", 1, 0);
    uint32_t i = 0;
    for (const Particle& p : particles) {
      bef(f, p.x[0]);
      bef(f, p.x[1]);
      bef(f, p.x[2]);
      bef(f, 1.0f);
      be32(f, i++);
      bef(f, p.v[0]);
      bef(f, p.v[1]);
      bef(f, p.v[2]);
      bef(f, 0.1f);
    }
    // primitive attribute dictionary: an indexed string "generator" with the single value "papi"
    hstr(f, "generator");
    be16(f, 1);
    be32(f, 4);
    be32(f, 1);
    hstr(f, "papi");
    // the particle system: key, vertex count, one vertex per point, then its generator index
    be32(f, 0x00008000u);
    be32(f, n);
    for (uint32_t q = 0; q < n; ++q) be32(f, q);
    be32(f, 0);
    std::fputc(0x00, f);
    std::fputc(0xff, f);
  }
};

// fillVoxelGrid_distance (mesh_builder.h:89-138): float arithmetic, C truncation of the ranges
inline void fill_voxel_grid_distance(const std::vector<Particle>& particles, real dx, real distance_cutoff, int G, std::vector<float>& grid) {
  grid.resize((size_t)G * G * G);
  const real r_gridpoints = distance_cutoff / dx;
  // every thread owns a range of i-planes and clips each particle's box to it: min() is order-independent, so
  // the result is the sequential one without atomics
#pragma omp parallel
  {
#ifdef _OPENMP
    const int nt = omp_get_num_threads(), me = omp_get_thread_num();
#else
    const int nt = 1, me = 0;
#endif
    const int i0 = (int)((long long)G * me / nt), i1 = (int)((long long)G * (me + 1) / nt);
    std::fill(grid.begin() + (size_t)i0 * G * G, grid.begin() + (size_t)i1 * G * G, distance_cutoff);
    for (const Particle& particle : particles) {
      int b[3], e[3];
      for (int d = 0; d < 3; ++d) {
        const real xg = particle.x[d] / dx;
        b[d] = std::max(0, (int)(xg - r_gridpoints));
        e[d] = std::min(G, (int)(xg + (r_gridpoints + 1.0f)));
      }
      for (int i = std::max(b[0], i0); i < std::min(e[0], i1); ++i) {
        const real dxn = particle.x[0] - i * dx;
        for (int j = b[1]; j < e[1]; ++j) {
          const real dyn = particle.x[1] - j * dx;
          for (int k = b[2]; k < e[2]; ++k) {
            const real dzn = particle.x[2] - k * dx;
            const real dist = std::sqrt(dxn * dxn + dyn * dyn + dzn * dzn);
            float& g = grid[((size_t)i * G + j) * G + k];
            g = std::min(dist, g);
          }
        }
      }
    }
  }
}

// iso-surface S = 0 of a scalar lattice (index i*G*G + j*G + k, i slowest like mesh_builder.h:185-189);
// vertices in lattice units; triangles wind so that normals point from S > 0 (inside) to S < 0
inline void marching_tetrahedra(const std::vector<double>& S, int G, std::vector<double>& V, std::vector<int>& F) {
  V.clear();
  F.clear();
  std::unordered_map<unsigned long long, int> edge_vertex;
  auto lattice = [&](int i, int j, int k) { return ((size_t)i * G + j) * G + k; };
  auto vertex_on = [&](size_t a, size_t b) {
    if (a > b) std::swap(a, b);
    const unsigned long long key = (unsigned long long)a * (unsigned long long)S.size() + b;
    auto it = edge_vertex.find(key);
    if (it != edge_vertex.end()) return it->second;
    const double sa = S[a], sb = S[b];
    const double tt = sa / (sa - sb);
    const size_t ia = a / ((size_t)G * G), ja = (a / G) % G, ka = a % G;
    const size_t ib = b / ((size_t)G * G), jb = (b / G) % G, kb = b % G;
    const int id = (int)(V.size() / 3);
    V.push_back(ia + tt * ((double)ib - (double)ia));
    V.push_back(ja + tt * ((double)jb - (double)ja));
    V.push_back(ka + tt * ((double)kb - (double)ka));
    edge_vertex.emplace(key, id);
    return id;
  };
  auto pos = [&](size_t a, double out[3]) {
    out[0] = (double)(a / ((size_t)G * G));
    out[1] = (double)((a / G) % G);
    out[2] = (double)(a % G);
  };
  auto emit = [&](int a, int b, int c, size_t inside_corner) {
    if (a == b || b == c || a == c) return;
    const double *pa = &V[3 * a], *pb = &V[3 * b], *pc = &V[3 * c];
    const double u[3] = {pb[0] - pa[0], pb[1] - pa[1], pb[2] - pa[2]}, w[3] = {pc[0] - pa[0], pc[1] - pa[1], pc[2] - pa[2]};
    const double n[3] = {u[1] * w[2] - u[2] * w[1], u[2] * w[0] - u[0] * w[2], u[0] * w[1] - u[1] * w[0]};
    double q[3];
    pos(inside_corner, q);
    const double side = n[0] * (q[0] - pa[0]) + n[1] * (q[1] - pa[1]) + n[2] * (q[2] - pa[2]);
    if (side > 0) F.insert(F.end(), {a, c, b}); else F.insert(F.end(), {a, b, c});
  };
  // cube corners: bit 0 = +i, bit 1 = +j, bit 2 = +k; six tetrahedra sharing the diagonal 0-7
  static const int tets[6][4] = {{0, 1, 3, 7}, {0, 3, 2, 7}, {0, 2, 6, 7}, {0, 6, 4, 7}, {0, 4, 5, 7}, {0, 5, 1, 7}};
  // the cells the surface passes through, found plane by plane in parallel and visited in lattice order
  std::vector<std::vector<int>> crossed(G > 1 ? G - 1 : 0);
#pragma omp parallel for schedule(dynamic, 4)
  for (int i = 0; i < G - 1; ++i)
    for (int j = 0; j + 1 < G; ++j)
      for (int k = 0; k + 1 < G; ++k) {
        int n_in = 0;
        for (int q = 0; q < 8; ++q) n_in += S[lattice(i + (q & 1), j + ((q >> 1) & 1), k + ((q >> 2) & 1))] > 0.0;
        if (n_in != 0 && n_in != 8) crossed[i].push_back(j * G + k);
      }
  for (int i = 0; i + 1 < G; ++i)
    for (int jk : crossed[i]) {
        const int j = jk / G, k = jk % G;
        size_t c[8];
        for (int q = 0; q < 8; ++q) c[q] = lattice(i + (q & 1), j + ((q >> 1) & 1), k + ((q >> 2) & 1));
        for (auto& tet : tets) {
          size_t in[4], out[4];
          int ni = 0, no = 0;
          for (int q = 0; q < 4; ++q) {
            if (S[c[tet[q]]] > 0.0) in[ni++] = c[tet[q]]; else out[no++] = c[tet[q]];
          }
          if (ni == 1) {
            emit(vertex_on(in[0], out[0]), vertex_on(in[0], out[1]), vertex_on(in[0], out[2]), in[0]);
          } else if (ni == 3) {
            emit(vertex_on(out[0], in[0]), vertex_on(out[0], in[1]), vertex_on(out[0], in[2]), in[0]);
          } else if (ni == 2) {
            const int a = vertex_on(in[0], out[0]), b = vertex_on(in[0], out[1]), cc = vertex_on(in[1], out[1]), d = vertex_on(in[1], out[0]);
            emit(a, b, cc, in[0]);
            emit(a, cc, d, in[0]);
          }
        }
      }
}

class MeshBuilder {
 public:
  MeshBuilder(const SimulationParameters& params, const CLIOptions& flags, u32 grid_size) : params_(params), flags_(flags), side_(grid_size) {}

  // returns the mesh as well as writing it (filename may be empty: no file)
  bool computeMesh(const std::string& filename, const std::vector<Particle>& particles, std::vector<double>* V_out = nullptr,
                   std::vector<int>* F_out = nullptr) const {
    const double voxel_dx = params_.N * params_.dx / double(side_);
    const int G = (int)side_;
    std::vector<float> sdf;
    fill_voxel_grid_distance(particles, (real)voxel_dx, (real)((flags_.mesh_particle_radius + 1) * voxel_dx), G, sdf);
    std::vector<double> S(sdf.size());
#pragma omp parallel for schedule(static)
    for (long long q = 0; q < (long long)sdf.size(); ++q) S[q] = flags_.mesh_particle_radius * voxel_dx - double(sdf[q]);
    std::vector<double> V;
    std::vector<int> F;
    marching_tetrahedra(S, G, V, F);
    for (double& v : V) v *= voxel_dx;
    if (flags_.laplacian_smooth != 0 && !smooth_mesh(V, F))  // mesh_builder.h:196-200
      std::cout << "Warning: the smoothing system is not positive definite (degenerate mesh); smoothing skipped" << std::endl;
    if (flags_.mesh_face_count != -1) decimate_mesh(V, F, (size_t)std::max(0, flags_.mesh_face_count));  // mesh_builder.h:202-208
    bool ok = true;
    if (!filename.empty()) ok = write_obj(filename, V, F);
    if (V_out) *V_out = std::move(V);
    if (F_out) *F_out = std::move(F);
    return ok;
  }

 private:
  const SimulationParameters& params_;
  const CLIOptions flags_;
  const u32 side_;
};

}  // namespace mpmh
