// C entry points of the host front end (libmpm_b200_host.so), so that the parity tests can drive
// the C++ facade from Python: scene loading / sampling (no GPU needed), the Simulation facade
// (GPU), particle and mesh output.  Not part of the drop-in boundary; that is include/mpm_b200.h.
#include <cstring>
#include <memory>

#include "output.hpp"
#include "scene.hpp"

using namespace mpmh;

struct MpmhScene {
  CLIOptions flags;
  std::vector<MaterialModel> materials;
  std::unique_ptr<Simulation> sim;
  std::string err;
};

static thread_local std::string g_err;

extern "C" {

const char* mpmh_last_error(const MpmhScene* s) { return s ? s->err.c_str() : g_err.c_str(); }

// argv-style options exactly as the CLI takes them; seed >= 0 calls srand(seed) first (the
// reference never seeds: glibc default = srand(1))
MpmhScene* mpmh_scene_load(int argc, char** argv, int seed) {
  auto s = std::make_unique<MpmhScene>();
  std::string err;
  if (!s->flags.parse(argc, argv, err)) {
    g_err = err;
    return nullptr;
  }
  try {
    if (seed >= 0) srand((unsigned)seed);
    s->sim = std::make_unique<Simulation>(s->flags, InterpolationKernel(), s->materials);
    load_scene(s->flags, s->materials, *s->sim);
  } catch (const std::exception& e) {
    g_err = e.what();
    return nullptr;
  }
  return s.release();
}
void mpmh_scene_free(MpmhScene* s) { delete s; }

int mpmh_grid_size(const MpmhScene* s) { return (int)s->flags.N; }
int mpmh_n_materials(const MpmhScene* s) { return (int)s->materials.size(); }
void mpmh_get_materials(const MpmhScene* s, float* out7) { std::memcpy(out7, s->materials.data(), s->materials.size() * sizeof(MaterialModel)); }
int mpmh_n_objects(const MpmhScene* s) { return (int)s->sim->objects.size(); }
size_t mpmh_object_count(const MpmhScene* s, int o) { return s->sim->objects[(size_t)o].particles.size(); }
int mpmh_object_substituted(const MpmhScene* s, int o) { return s->sim->objects[(size_t)o].substituted_mesh ? 1 : 0; }
void mpmh_object_lifetime(const MpmhScene* s, int o, float* begin, float* end) {
  *begin = s->sim->objects[(size_t)o].lifetime_begin;
  *end = s->sim->objects[(size_t)o].lifetime_end;
}
size_t mpmh_full_count(const MpmhScene* s) { return s->sim->getFullParticleCount(); }
size_t mpmh_active_count(MpmhScene* s) { return s->sim->getActiveParticleList().size(); }
void mpmh_get_full(MpmhScene* s, MpmParticle* out) {
  auto& v = s->sim->getFullParticleList();
  std::memcpy(out, v.data(), v.size() * sizeof(MpmParticle));
}
void mpmh_get_active(MpmhScene* s, MpmParticle* out) {
  auto& v = s->sim->getActiveParticleList();
  std::memcpy(out, v.data(), v.size() * sizeof(MpmParticle));
}
double mpmh_time(const MpmhScene* s) { return s->sim->t; }

#define MPMH_TRY(stmt)        \
  try {                       \
    stmt;                     \
    return 0;                 \
  } catch (const std::exception& e) { \
    s->err = e.what();        \
    return 1;                 \
  }
int mpmh_init_cuda(MpmhScene* s) { MPMH_TRY(s->sim->initCuda()) }
int mpmh_advance(MpmhScene* s, int n) { MPMH_TRY(for (int i = 0; i < n; ++i) s->sim->advance()) }
int mpmh_sync_device(MpmhScene* s) { MPMH_TRY(s->sim->syncDevice()) }
int mpmh_write_particles(MpmhScene* s, const char* path) { MPMH_TRY(if (!ParticleWriter().writeParticles(path, s->sim->getActiveParticleList())) throw std::runtime_error("cannot write particles")) }
// mesh of the active particles; returns vertex / face counts, copies them out when the buffers are given
int mpmh_compute_mesh(MpmhScene* s, const char* path, size_t* n_vertices, size_t* n_faces) {
  MPMH_TRY({
    MeshBuilder mesher(s->sim->par, s->flags, s->flags.mesh_grid);
    std::vector<double> V;
    std::vector<int> F;
    if (!mesher.computeMesh(path ? path : "", s->sim->getActiveParticleList(), &V, &F)) throw std::runtime_error("cannot write mesh");
    *n_vertices = V.size() / 3;
    *n_faces = F.size() / 3;
  })
}

// building blocks, for unit tests
int mpmh_winding_numbers(const float* V, size_t nv, const int* F, size_t nf, const float* points, size_t np, float* w_out) {
  TriMesh m;
  m.V.assign(V, V + 3 * nv);
  m.F.assign(F, F + 3 * nf);
  for (size_t i = 0; i < np; ++i) w_out[i] = winding_number(m, points + 3 * i);
  return 0;
}
// the stand-in (or file) mesh after loadMesh's rescale; call twice: first with null buffers for the sizes
int mpmh_load_mesh(const char* path, double size, const float* position, float* V, int* F, size_t* nv, size_t* nf, int* substituted) {
  TriMesh m;
  bool sub = false;
  std::string err;
  if (!load_mesh_or_stand_in(path, m, &sub, &err)) {
    g_err = err;
    return 1;
  }
  rescale_mesh(m, size, position);
  *nv = m.n_vertices();
  *nf = m.n_faces();
  *substituted = sub;
  if (V) std::memcpy(V, m.V.data(), m.V.size() * sizeof(float));
  if (F) std::memcpy(F, m.F.data(), m.F.size() * sizeof(int));
  return 0;
}
int mpmh_marching_tetrahedra(const double* S, int G, double* V, size_t v_cap, int* F, size_t f_cap, size_t* nv, size_t* nf) {
  std::vector<double> s(S, S + (size_t)G * G * G), v;
  std::vector<int> f;
  marching_tetrahedra(s, G, v, f);
  *nv = v.size() / 3;
  *nf = f.size() / 3;
  if (V && v.size() <= 3 * v_cap) std::memcpy(V, v.data(), v.size() * sizeof(double));
  if (F && f.size() <= 3 * f_cap) std::memcpy(F, f.data(), f.size() * sizeof(int));
  return 0;
}

}  // extern "C"
