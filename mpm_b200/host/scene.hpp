// Scene loading: what main() does between constructing the Simulation and initCuda()
// (reference src/main.cu:25-66): read the TOML file, build one MaterialModel per [[material]]
// (volume = 1 / --particle-count, defaults density 700, E 1.4e5, Nu 0.2, hardening 10, clamps
// 0.975 / 1.0075), then addObject per [[object]] in file order (defaults mesh "sphere.obj",
// size 1.0, lifetime [0, max)), meshes looked up in <scene dir>/meshes/.
#pragma once
#include <map>
#include <string>

#include "simulation.hpp"
#include "toml_lite.hpp"

namespace mpmh {

inline std::string scene_meshes_dir(const std::string& scene_path) {
  const size_t slash = scene_path.find_last_of('/');
  return (slash == std::string::npos ? std::string("") : scene_path.substr(0, slash + 1)) + "meshes/";
}

inline Vec to_vector(const TomlTable& t, const std::string& key) {
  std::vector<double> a;
  Vec out;
  if (!t.numbers(key, a)) throw std::runtime_error("object is missing the array '" + key + "'");  // cpptoml: dereferencing an empty option
  int i = 0;
  for (const double value : a) {
    if (i < 3) out[i] = (real)value;
    i++;
  }
  return out;
}

// fills `material_models` (which `simulation` references) and adds every object
inline void load_scene(const CLIOptions& flags, std::vector<MaterialModel>& material_models, Simulation& simulation, bool verbose = false) {
  const std::string meshes_dir = scene_meshes_dir(flags.scene);
  const TomlDoc config = toml_parse_file(flags.scene);
  std::map<std::string, u8> material_index;
  int i = 0;
  for (const TomlTable& material : config.table_array("material")) {
    material_models.push_back(make_material_model(1.0 / flags.particle_count, material.number_or("density", 700.0), material.number_or("E", 1.4e5),
                                                  material.number_or("Nu", 0.2), material.number_or("hardening", 10.0),
                                                  material.number_or("plast_clamp_lower", 0.975), material.number_or("plast_clamp_higher", 1.0075)));
    material_index[material.string_or("name", "")] = (u8)i;
    i++;
  }
  i = 0;
  for (const TomlTable& object : config.table_array("object")) {
    if (verbose) std::cout << "Adding object " << i << "\r" << std::flush;
    const u32 material = material_index[object.string_or("material", "")];  // unknown name -> index 0, like std::map::operator[]
    simulation.addObject(meshes_dir + object.string_or("mesh", "sphere.obj"), (int)material, (real)object.number_or("size", 1.0),
                         to_vector(object, "position"), to_vector(object, "velocity"), (real)object.number_or("lifetime_begin", 0.0),
                         (real)object.number_or("lifetime_end", std::numeric_limits<double>::max()));
    i++;
  }
}

}  // namespace mpmh
