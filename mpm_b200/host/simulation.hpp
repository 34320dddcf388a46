// Host-side facade with the public surface of the reference's `class Simulation`
// (include/mpm.cuh:44-88, src/mpm.cu:180-394), implemented over the C ABI of include/mpm_b200.h:
//   Simulation(opts, kernel, material_models), addObject, initCuda, advance, syncDevice,
//   getFullParticleCount, getFullParticleList, getActiveParticleList; public par, N, t, objects,
//   active_particle_count, material_models.
// The plugin aliases of include/mpm.cuh:24-27 are kept as names: Particle = MpmParticle (same 104
// bytes), MaterialModel = MpmMaterial (same 7 floats, built by mpm_make_material with the
// constructor arithmetic of MaterialModel.cuh), InterpolationKernel = the quadratic B-spline tag.
//
// Differences, all deliberate:
//   * object lifetimes: the reference notices a changed active set only inside particlesToHost and
//     then re-uploads stale data; here advance() activates an object at the first substep with
//     t >= lifetime_begin by appending its particles on the device (mpm_append_particles_aos) and
//     retires one at t >= lifetime_end: its last state is read back into the host copy, then it is removed
//     by a compaction on the device (mpm_remove_particles); the survivors are not re-uploaded.
//   * syncDevice() downloads into the per-object vectors exactly like particlesToHost; the
//     cheaper positions-only read-back for viewers is syncPositions().
//   * errors are reported: every C-ABI failure throws std::runtime_error with mpm_last_error.
#pragma once
#include <cstdlib>
#include <limits>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/mpm_b200.h"
#include "mesh.hpp"
#include "options.hpp"

namespace mpmh {

using Particle = MpmParticle;
using MaterialModel = MpmMaterial;
struct QuadraticInterpolationKernel {  // tag: the kernels are instantiated for this one (mpm.cuh:26)
  static constexpr u32 size() { return 3; }
  static constexpr bool d_is_const() { return true; }
};
using InterpolationKernel = QuadraticInterpolationKernel;

struct Vec {
  real v[3] = {0, 0, 0};
  Vec() = default;
  Vec(real x, real y, real z) : v{x, y, z} {}
  real& operator()(int i) { return v[i]; }
  real operator()(int i) const { return v[i]; }
  real& operator[](int i) { return v[i]; }
  real operator[](int i) const { return v[i]; }
};

// MaterialModel(volume, density, E, Nu, hardening, lo, hi) as src/main.cu:36-42 calls it
inline MaterialModel make_material_model(double volume, double density = 700.0, double E = 1.4e5, double Nu = 0.2,
                                         double hardening = 10.0, double lo = 0.975, double hi = 1.0075) {
  MaterialModel m;
  mpm_make_material(volume, density, E, Nu, hardening, lo, hi, &m);
  return m;
}

// Particle(material_index, x, velocity): F = I, C = 0, Jp = 1 (types.h:31-33, TransferScheme.h:50-53)
inline Particle make_particle(u8 material, const Vec& x, const Vec& v) {
  Particle p{};
  p.material_type = material;
  for (int d = 0; d < 3; ++d) {
    p.x[d] = x[d];
    p.v[d] = v[d];
  }
  p.F[0] = p.F[4] = p.F[8] = 1.0f;
  p.Jp = 1.0f;
  return p;
}

struct SimulationParameters {  // include/TransferScheme.h:6-29
  real dt;
  u32 N;
  real N_real;
  real dx;
  real dx_inv;
  SimulationParameters(real dt_, u32 N_) : dt(dt_), N(N_), N_real(real(N_)), dx((real)(1.0 / N_)), dx_inv((real)(1.0 / (double)dx)) {}
};

struct SimObject {  // include/mpm.cuh:30-42
  explicit SimObject(MaterialModel model) : materialModel(model) {}
  std::vector<Particle> particles;
  MaterialModel materialModel;
  real lifetime_begin = 0.0;
  real lifetime_end = std::numeric_limits<real>::max();
  bool substituted_mesh = false;  // the .obj was an LFS stub: procedural stand-in used (mesh.hpp)
  bool isActive(real t) const { return (t >= lifetime_begin && t < lifetime_end); }
};

inline float get_random() { return float(rand()) / float(RAND_MAX); }  // src/mpm.cu:10-12

class Simulation {
 public:
  SimulationParameters par;
  u32& N = par.N;
  double t = 0.0;
  std::vector<SimObject> objects;
  int active_particle_count = 0;
  InterpolationKernel interpolationKernel;
  std::vector<MaterialModel> const& material_models;

  Simulation(const CLIOptions& opts, InterpolationKernel const& kernel, std::vector<MaterialModel> const& models)
      : par(opts.dt, opts.N), interpolationKernel(kernel), material_models(models), opts_(opts) {}
  ~Simulation() { mpm_destroy(sim_); }
  Simulation(const Simulation&) = delete;
  Simulation& operator=(const Simulation&) = delete;

  // src/mpm.cu:237-251.  The .obj may be a Git-LFS stub: see mesh.hpp.
  void addObject(std::string const& filepath, int material_model_index, real size, Vec position, Vec velocity,
                 real lifetime_begin = 0.0, real lifetime_end = std::numeric_limits<real>::max()) {
    TriMesh mesh;
    bool substituted = false;
    std::string err;
    if (!load_mesh_or_stand_in(filepath, mesh, &substituted, &err)) throw std::runtime_error(err);
    rescale_mesh(mesh, (double)size, position.v);
    const MaterialModel material = material_models.at((size_t)material_model_index);
    objects.push_back(SimObject(material));
    addParticles(mesh, u32(1.0 / material.particleVolume), velocity, objects.back().particles, (u8)material_model_index);
    objects.back().lifetime_begin = lifetime_begin;
    objects.back().lifetime_end = lifetime_end;
    objects.back().substituted_mesh = substituted;
  }

  // src/mpm.cu:197-207: creates the device handle and uploads the objects active at t
  void initCuda() {
    MpmParams p{};
    p.dt = par.dt;
    p.N = par.N;
    // mpm.cuh:25: MaterialModel = MMSnow; the other classes of MaterialModel.cuh by --model
    p.model = opts_.model == "fixed_corotated" ? MPM_MODEL_FIXED_COROTATED : opts_.model == "jelly" ? MPM_MODEL_JELLY : MPM_MODEL_SNOW;
    p.svd_mode = opts_.svd == "fast" ? MPM_SVD_FAST : MPM_SVD_EXACT;
    p.sort_every = opts_.sort_every;
    p.rebin_permille = opts_.rebin_permille;
    p.graph_mode = opts_.graphs == "off" ? MPM_GRAPH_OFF : opts_.graphs == "on" ? MPM_GRAPH_ON : MPM_GRAPH_AUTO;
    p.device = -1;
    p.capacity = std::max<size_t>(getFullParticleCount(), 1);  // every object fits: activation never reallocates
    if (material_models.empty()) throw std::runtime_error("no materials");
    if (mpm_create(&p, material_models.data(), (int)material_models.size(), &sim_)) throw std::runtime_error(mpm_last_error(nullptr));
    uploadActive();
  }

  // src/mpm.cu:323-329, plus the lifetime handling described above.  Like the reference's advance()
  // this returns before the GPU has done anything; here the substeps are also handed to the device
  // in batches (at most kBatch, flushed by syncDevice / syncPositions / a change of the active set):
  // within one mpm_advance call the kernels hand the next P2G's affine matrix over (MPM_PIPE_HANDOVER).
  static constexpr int kBatch = 20;  // the reference's sync cadence, src/main.cu:99
  void advance() {
    if (!sim_) throw std::runtime_error("advance() before initCuda()");
    if (activeSetChanged()) {
      flush();
      applyLifetimes();
    }
    ++pending_;
    t += par.dt;
    if (pending_ >= kBatch) flush();
  }
  void flush() {
    if (sim_ && pending_) {
      const int n = pending_;
      pending_ = 0;
      check(mpm_advance(sim_, n));
    }
  }

  // src/mpm.cu:209-211, 288-306: device -> per-object host vectors (blocking)
  void syncDevice() {
    if (!sim_) return;
    flush();
    size_t n = 0;
    host_.resize(uploaded_count());
    check(mpm_download_particles_aos(sim_, host_.data(), host_.size(), &n));
    size_t i = 0;
    for (size_t o : uploaded_)
      for (Particle& q : objects[o].particles) q = host_[i++];
  }

  // positions only (12 B/particle) in upload order (= getActiveParticleList order until an object is appended): what a viewer needs
  void syncPositions(std::vector<float>& xyz) {
    flush();
    xyz.resize(3 * uploaded_count());
    size_t n = 0;
    if (sim_ && !xyz.empty()) check(mpm_download_positions(sim_, xyz.data(), xyz.size() / 3, &n));
  }

  size_t getFullParticleCount() const {
    size_t n = 0;
    for (auto& object : objects) n += object.particles.size();
    return n;
  }
  std::vector<Particle>& getFullParticleList() {
    particles_all_.resize(0);
    for (auto& object : objects) particles_all_.insert(particles_all_.end(), object.particles.begin(), object.particles.end());
    return particles_all_;
  }
  std::vector<Particle>& getActiveParticleList() {
    particles_all_.resize(0);
    for (auto& object : objects) {
      if (!object.isActive((real)t)) continue;
      particles_all_.insert(particles_all_.end(), object.particles.begin(), object.particles.end());
    }
    return particles_all_;
  }

  MpmSim* handle() {
    flush();
    return sim_;
  }

  // src/mpm.cu:348-394: rejection sampling in the bounding box, glibc rand() in x, y, z order,
  // 2048 points per batch, a point is kept when its winding number truncates to 1
  static void addParticles(const TriMesh& mesh, u32 particle_density, Vec velocity, std::vector<Particle>& particles, u8 material_index) {
    const u32 BatchSize = 2048;
    std::vector<float> points(3 * BatchSize);
    double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300};
    for (size_t i = 0; i < mesh.n_vertices(); ++i)
      for (int d = 0; d < 3; ++d) {
        mn[d] = std::min(mn[d], (double)mesh.V[3 * i + d]);
        mx[d] = std::max(mx[d], (double)mesh.V[3 * i + d]);
      }
    const double range_x = mx[0] - mn[0], range_y = mx[1] - mn[1], range_z = mx[2] - mn[2];
    const double volume_bounding_box = range_x * range_y * range_z;
    const u32 particle_count_target = (u32)(particle_density * volume_bounding_box);
    u32 count_tot = 0;
    std::vector<int> W(BatchSize);
    while (count_tot < particle_count_target) {
      for (u32 i = 0; i < BatchSize; i++) {
        points[3 * i + 0] = (float)(mn[0] + get_random() * range_x);
        points[3 * i + 1] = (float)(mn[1] + get_random() * range_y);
        points[3 * i + 2] = (float)(mn[2] + get_random() * range_z);
      }
#pragma omp parallel for schedule(static)
      for (int i = 0; i < (int)BatchSize; i++) W[i] = (int)winding_number(mesh, &points[3 * i]);  // float -> int truncation (MatrixXi W)
      for (u32 i = 0; i < BatchSize; i++) {
        count_tot++;
        if (W[i] == 1) particles.push_back(make_particle(material_index, Vec(points[3 * i], points[3 * i + 1], points[3 * i + 2]), velocity));
        if (count_tot == particle_count_target) return;
      }
    }
  }

 private:
  CLIOptions opts_;
  MpmSim* sim_ = nullptr;
  int pending_ = 0;  // substeps advance() has accepted and not yet handed to the device
  std::vector<size_t> uploaded_;  // object indices on the device, in upload order
  std::vector<Particle> particles_all_, host_;

  void check(int rc) const {
    if (rc) throw std::runtime_error(mpm_last_error(sim_));
  }
  size_t uploaded_count() const {
    size_t n = 0;
    for (size_t o : uploaded_) n += objects[o].particles.size();
    return n;
  }
  std::vector<size_t> activeNow() const {
    std::vector<size_t> a;
    for (size_t o = 0; o < objects.size(); ++o)
      if (objects[o].isActive((real)t)) a.push_back(o);
    return a;
  }
  bool activeSetChanged() const {
    size_t k = 0;  // both lists are in object order only if nothing was appended out of order: compare as sets
    const std::vector<size_t> a = activeNow();
    if (a.size() != uploaded_.size()) return true;
    for (size_t o : a) {
      bool found = false;
      for (size_t u : uploaded_) found = found || u == o;
      if (!found) return true;
      ++k;
    }
    return false;
  }
  void uploadActive() {
    uploaded_ = activeNow();
    host_.clear();
    for (size_t o : uploaded_) host_.insert(host_.end(), objects[o].particles.begin(), objects[o].particles.end());
    check(mpm_upload_particles_aos(sim_, host_.data(), host_.size()));
    active_particle_count = (int)host_.size();
  }
  void applyLifetimes() {
    const std::vector<size_t> now = activeNow();
    // objects that ended: their last state goes to the host copies (as the reference's particlesToHost does when it
    // notices the change), then they are removed on the device by a stable compaction — no re-upload, no re-bin;
    // the upload order of the survivors closes up, exactly as uploaded_ does here
    bool any_ended = false;
    for (size_t u : uploaded_) {
      bool still = false;
      for (size_t o : now) still = still || o == u;
      any_ended = any_ended || !still;
    }
    if (any_ended) syncDevice();
    for (size_t k = 0; k < uploaded_.size();) {
      bool still = false;
      for (size_t o : now) still = still || o == uploaded_[k];
      if (still) {
        ++k;
        continue;
      }
      size_t first = 0;
      for (size_t q = 0; q < k; ++q) first += objects[uploaded_[q]].particles.size();
      check(mpm_remove_particles(sim_, first, objects[uploaded_[k]].particles.size()));
      uploaded_.erase(uploaded_.begin() + (long)k);
    }
    // objects entering: appended on the device
    for (size_t o : now) {
      bool have = false;
      for (size_t u : uploaded_) have = have || u == o;
      if (have) continue;
      check(mpm_append_particles_aos(sim_, objects[o].particles.data(), objects[o].particles.size()));
      uploaded_.push_back(o);
    }
    active_particle_count = (int)uploaded_count();
  }
};

}  // namespace mpmh
