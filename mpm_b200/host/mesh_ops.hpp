// Post-processing of the output surface mesh: the reference's optional smoothing and decimation
// (include/mesh_builder.h:196-211, 213-245), without libigl / Eigen:
//   * smooth_mesh: ONE implicit step of cotangent-Laplacian (mean-curvature) flow,
//       (M - 0.001 L) U = M V,  M = barycentric (lumped) mass matrix, L = cotangent matrix of V,
//     solved per coordinate by Jacobi-preconditioned conjugate gradients (the matrix is SPD), then the
//     reference's normalisation U /= sqrt(total area) (mesh_builder.h:241 — it rescales the mesh to unit
//     area; kept because that is what the reference writes).
//   * decimate_mesh: shortest-edge collapse to the midpoint until at most `max_faces` faces remain —
//     the cost / placement pair igl::decimate(V, F, max_m, ...) uses by default
//     (shortest_edge_and_midpoint) — with the link condition as validity test so the surface stays
//     manifold.  Same criterion, own implementation: the reference has no golden meshes, the tests
//     check face count, closedness and volume.
#pragma once
#include <algorithm>
#include <array>
#include <cmath>
#include <map>
#include <queue>
#include <set>
#include <unordered_map>
#include <vector>

namespace mpmh {

inline double tri_area(const double* a, const double* b, const double* c) {
  const double u[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]}, w[3] = {c[0] - a[0], c[1] - a[1], c[2] - a[2]};
  const double n[3] = {u[1] * w[2] - u[2] * w[1], u[2] * w[0] - u[0] * w[2], u[0] * w[1] - u[1] * w[0]};
  return 0.5 * std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
}
inline double mesh_area(const std::vector<double>& V, const std::vector<int>& F) {
  double a = 0;
  for (size_t f = 0; f + 2 < F.size(); f += 3) a += tri_area(&V[3 * F[f]], &V[3 * F[f + 1]], &V[3 * F[f + 2]]);
  return a;
}

// sparse symmetric matrix as per-row (column, value) lists with the diagonal kept apart
struct SparseSym {
  std::vector<double> diag;
  std::vector<std::vector<std::pair<int, double>>> off;
  explicit SparseSym(size_t n) : diag(n, 0.0), off(n) {}
  void add(int i, int j, double v) {
    if (i == j) {
      diag[(size_t)i] += v;
      return;
    }
    for (auto& e : off[(size_t)i])
      if (e.first == j) {
        e.second += v;
        return;
      }
    off[(size_t)i].push_back({j, v});
  }
  void mul(const std::vector<double>& x, std::vector<double>& y) const {
    const long long n = (long long)diag.size();
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < n; ++i) {
      double s = diag[(size_t)i] * x[(size_t)i];
      for (auto& e : off[(size_t)i]) s += e.second * x[(size_t)e.first];
      y[(size_t)i] = s;
    }
  }
};

inline bool smooth_mesh(std::vector<double>& V, const std::vector<int>& F, double step = 0.001) {
  const size_t n = V.size() / 3;
  if (n == 0 || F.empty()) return true;
  SparseSym S(n);  // M - step * L
  std::vector<double> mass(n, 0.0);
  for (size_t f = 0; f + 2 < F.size(); f += 3) {
    const int v[3] = {F[f], F[f + 1], F[f + 2]};
    const double* p[3] = {&V[3 * v[0]], &V[3 * v[1]], &V[3 * v[2]]};
    const double area = tri_area(p[0], p[1], p[2]);
    if (!(area > 0)) continue;
    for (int c = 0; c < 3; ++c) mass[(size_t)v[c]] += area / 3.0;
    for (int c = 0; c < 3; ++c) {  // cotangent at corner c weighs the opposite edge (a, b)
      const int a = v[(c + 1) % 3], b = v[(c + 2) % 3];
      const double* pc = p[c];
      const double* pa = p[(c + 1) % 3];
      const double* pb = p[(c + 2) % 3];
      const double e1[3] = {pa[0] - pc[0], pa[1] - pc[1], pa[2] - pc[2]}, e2[3] = {pb[0] - pc[0], pb[1] - pc[1], pb[2] - pc[2]};
      const double dot = e1[0] * e2[0] + e1[1] * e2[1] + e1[2] * e2[2];
      const double w = 0.5 * dot / (2.0 * area);  // 0.5 * cot(angle at c); L_ab += w, L_aa -= w, L_bb -= w
      S.add(a, b, -step * w);
      S.add(b, a, -step * w);
      S.add(a, a, step * w);
      S.add(b, b, step * w);
    }
  }
  for (size_t i = 0; i < n; ++i) S.diag[i] += mass[i];
  std::vector<double> x(n), r(n), z(n), p(n), q(n);
  for (int d = 0; d < 3; ++d) {
    for (size_t i = 0; i < n; ++i) x[i] = V[3 * i + d];
    S.mul(x, q);
    double rz = 0, r0 = 0;
    for (size_t i = 0; i < n; ++i) {
      r[i] = mass[i] * V[3 * i + d] - q[i];
      z[i] = S.diag[i] > 0 ? r[i] / S.diag[i] : r[i];
      p[i] = z[i];
      rz += r[i] * z[i];
      r0 += r[i] * r[i];
    }
    for (int it = 0; it < 500 && r0 > 0; ++it) {
      S.mul(p, q);
      double pq = 0;
      for (size_t i = 0; i < n; ++i) pq += p[i] * q[i];
      if (!(pq > 0)) return false;  // not positive definite: degenerate mesh
      const double alpha = rz / pq;
      double rz_new = 0, rr = 0;
      for (size_t i = 0; i < n; ++i) {
        x[i] += alpha * p[i];
        r[i] -= alpha * q[i];
        z[i] = S.diag[i] > 0 ? r[i] / S.diag[i] : r[i];
        rz_new += r[i] * z[i];
        rr += r[i] * r[i];
      }
      if (rr <= 1e-24 * r0) break;
      const double beta = rz_new / rz;
      rz = rz_new;
      for (size_t i = 0; i < n; ++i) p[i] = z[i] + beta * p[i];
    }
    for (size_t i = 0; i < n; ++i) V[3 * i + d] = x[i];
  }
  const double area = mesh_area(V, F);
  if (area > 0)
    for (double& v : V) v /= std::sqrt(area);  // mesh_builder.h:241
  return true;
}

// collapses shortest edges to their midpoints until F holds at most max_faces triangles (or no valid
// collapse is left); compacts V and F; returns the number of collapses
inline size_t decimate_mesh(std::vector<double>& V, std::vector<int>& F, size_t max_faces) {
  const size_t nv = V.size() / 3;
  size_t n_faces = F.size() / 3;
  if (n_faces <= max_faces) return 0;
  std::vector<std::set<int>> vfaces(nv);  // faces around a vertex
  std::vector<char> alive(n_faces, 1);
  for (size_t f = 0; f < n_faces; ++f)
    for (int c = 0; c < 3; ++c) vfaces[(size_t)F[3 * f + c]].insert((int)f);
  std::vector<unsigned> stamp(nv, 0);  // bumped whenever a vertex moves: stale queue entries are skipped
  struct Item {
    double len2;
    int a, b;
    unsigned sa, sb;
    bool operator<(const Item& o) const { return len2 > o.len2; }
  };
  std::priority_queue<Item> pq;
  auto len2 = [&](int a, int b) {
    double s = 0;
    for (int d = 0; d < 3; ++d) s += (V[3 * a + d] - V[3 * b + d]) * (V[3 * a + d] - V[3 * b + d]);
    return s;
  };
  auto push = [&](int a, int b) {
    if (a > b) std::swap(a, b);
    pq.push({len2(a, b), a, b, stamp[(size_t)a], stamp[(size_t)b]});
  };
  for (size_t f = 0; f < n_faces; ++f)
    for (int c = 0; c < 3; ++c) {
      const int a = F[3 * f + c], b = F[3 * f + (c + 1) % 3];
      if (a < b) push(a, b);  // each interior edge once (closed surface: the other face sees it as (b, a))
    }
  auto neighbours = [&](int v) {
    std::set<int> nb;
    for (int f : vfaces[(size_t)v])
      for (int c = 0; c < 3; ++c)
        if (F[3 * f + c] != v) nb.insert(F[3 * f + c]);
    return nb;
  };
  size_t collapses = 0;
  while (n_faces > max_faces && !pq.empty()) {
    const Item it = pq.top();
    pq.pop();
    const int a = it.a, b = it.b;
    if (vfaces[(size_t)a].empty() || vfaces[(size_t)b].empty()) continue;
    if (stamp[(size_t)a] != it.sa || stamp[(size_t)b] != it.sb) continue;  // stale
    // faces sharing the edge, and the link condition: the common neighbours of a and b must be exactly
    // the vertices opposite the edge in those faces (else the collapse pinches the surface)
    std::vector<int> shared;
    for (int f : vfaces[(size_t)a])
      if (vfaces[(size_t)b].count(f)) shared.push_back(f);
    if (shared.empty()) continue;
    const std::set<int> na = neighbours(a), nb = neighbours(b);
    size_t common = 0;
    for (int v : na) common += nb.count(v);
    if (common != shared.size()) continue;
    for (int d = 0; d < 3; ++d) V[3 * a + d] = 0.5 * (V[3 * a + d] + V[3 * b + d]);  // a <- midpoint, b disappears
    for (int f : shared) {
      alive[(size_t)f] = 0;
      for (int c = 0; c < 3; ++c) vfaces[(size_t)F[3 * f + c]].erase(f);
      --n_faces;
    }
    for (int f : vfaces[(size_t)b]) {
      for (int c = 0; c < 3; ++c)
        if (F[3 * f + c] == b) F[3 * f + c] = a;
      vfaces[(size_t)a].insert(f);
    }
    vfaces[(size_t)b].clear();
    ++stamp[(size_t)a];
    ++stamp[(size_t)b];
    for (int v : neighbours(a)) push(a, v);
    ++collapses;
  }
  // compact
  std::vector<int> remap(nv, -1);
  std::vector<double> V2;
  std::vector<int> F2;
  for (size_t f = 0; f < alive.size(); ++f) {
    if (!alive[f]) continue;
    for (int c = 0; c < 3; ++c) {
      const int v = F[3 * f + c];
      if (remap[(size_t)v] < 0) {
        remap[(size_t)v] = (int)(V2.size() / 3);
        V2.insert(V2.end(), {V[3 * v], V[3 * v + 1], V[3 * v + 2]});
      }
      F2.push_back(remap[(size_t)v]);
    }
  }
  V.swap(V2);
  F.swap(F2);
  return collapses;
}

}  // namespace mpmh
