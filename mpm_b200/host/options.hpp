// Command-line options of the simulator: the same flags, defaults and derived fields as the
// reference's CLIOptions (include/options.h:6-53), parsed without cxxopts.
// Accepted forms: --name value, --name=value, and for the grid size the reference's own spelling
// `-N value` (cxxopts turns the one-letter option name "N" into a SHORT option, include/cxxopts.h:1457,
// 1634-1662, so the reference's README commands say `-N 16`); `-N16`, `-N=16` and `--N 16` are taken too.
// A parse error prints a message and exits with status 1, like the reference (options.h:48-51).  Extensions (not in the reference, all optional):
// --steps (stop after this many substeps; the reference loops until killed), --svd exact|fast,
// --model snow|fixed_corotated|jelly (the compile-time MaterialModel alias of include/mpm.cuh:25 as data),
// --sort-every, --rebin-permille, --sync-every, --frame-rate, --graphs auto|off|on, --particle-format bgeo|pda.
#pragma once
#include <cstdint>
#include <cstdlib>
#include <iostream>
#include <map>
#include <string>

namespace mpmh {

using real = float;
using u8 = uint8_t;
using u32 = uint32_t;
using i32 = int32_t;

struct CLIOptions {
  // Simulation parameters.
  float dt = 1e-4f;
  u32 N = 60;
  u32 particle_count = 500000;  // particles per unit cube (a density, src/main.cu:36)
  real dx = 1.0f / 60.0f;
  real N_real = 60.0f;
  std::string save_dir;
  std::string scene = "scenes/snowman.toml";
  // Mesh builder parameters.
  u32 mesh_grid = 250;
  u32 mesh_particle_radius = 5;
  i32 mesh_face_count = -1;
  int laplacian_smooth = 0;
  // Extensions.
  long long steps = -1;
  std::string svd = "exact";
  std::string model = "snow";  // which MaterialModel alias: snow = MMSnow (the reference's choice, mpm.cuh:25), fixed_corotated = MMFixedCorotated
  u32 sort_every = 8;
  u32 rebin_permille = 0;  // MpmParams.rebin_permille
  u32 sync_every = 20;  // src/main.cu:99
  u32 frame_rate = 240; // src/main.cu:8
  std::string graphs = "auto";  // MpmParams.graph_mode: auto | off | on
  std::string particle_format = "bgeo";  // bgeo (the reference, src/main.cu:109) or pda (Partio ASCII)

  CLIOptions() { derive(); }
  CLIOptions(int argc, char* argv[]) {
    std::string err;
    if (!parse(argc, argv, err)) {
      std::cout << err << std::endl;
      std::exit(1);
    }
  }

  void derive() {
    dx = (real)(1.0 / N);  // options.h:41
    N_real = real(N);
  }

  bool parse(int argc, char* argv[], std::string& err) {
    std::map<std::string, std::string> kv;
    for (int i = 1; i < argc; ++i) {
      std::string a = argv[i];
      if (a.rfind("-N", 0) == 0 && a.rfind("--", 0) != 0) {  // the short option of the reference
        std::string val = a.substr(2);
        if (!val.empty() && val[0] == '=') val = val.substr(1);
        if (val.empty()) {
          if (i + 1 >= argc) {
            err = "Option 'N' is missing an argument";
            return false;
          }
          val = argv[++i];
        }
        kv["N"] = val;
        continue;
      }
      if (a.rfind("--", 0) != 0) {
        err = "Unexpected argument '" + a + "'";
        return false;
      }
      a = a.substr(2);
      std::string val;
      const size_t eq = a.find('=');
      if (eq != std::string::npos) {
        val = a.substr(eq + 1);
        a = a.substr(0, eq);
      } else {
        if (i + 1 >= argc) {
          err = "Option '" + a + "' is missing an argument";
          return false;
        }
        val = argv[++i];
      }
      kv[a] = val;
    }
    try {
      for (auto& [k, v] : kv) {
        if (k == "dt") dt = std::stof(v);
        else if (k == "N") N = to_u32(v);
        else if (k == "save-dir") save_dir = v;
        else if (k == "scene") scene = v;
        else if (k == "particle-count") particle_count = to_u32(v);
        else if (k == "mesh-grid") mesh_grid = to_u32(v);
        else if (k == "mesh-particle-radius") mesh_particle_radius = to_u32(v);
        else if (k == "mesh-face-count") mesh_face_count = (i32)std::stol(v);
        else if (k == "laplacian_smooth") laplacian_smooth = std::stoi(v);
        else if (k == "steps") steps = std::stoll(v);
        else if (k == "svd") svd = v;
        else if (k == "model") model = v;
        else if (k == "sort-every") sort_every = to_u32(v);
        else if (k == "rebin-permille") rebin_permille = to_u32(v);
        else if (k == "sync-every") sync_every = to_u32(v);
        else if (k == "frame-rate") frame_rate = to_u32(v);
        else if (k == "particle-format") particle_format = v;
        else if (k == "graphs") graphs = v;
        else {
          err = "Option '" + k + "' does not exist";
          return false;
        }
      }
    } catch (const std::exception&) {
      err = "Argument could not be parsed";
      return false;
    }
    if (N == 0 || sync_every == 0 || (graphs != "auto" && graphs != "off" && graphs != "on") || (particle_format != "bgeo" && particle_format != "pda") || (svd != "exact" && svd != "fast") || (model != "snow" && model != "fixed_corotated" && model != "jelly")) {
      err = "Argument out of range";
      return false;
    }
    derive();
    return true;
  }

 private:
  static u32 to_u32(const std::string& s) {
    size_t pos = 0;
    const unsigned long long v = std::stoull(s, &pos);
    if (pos != s.size() || v > 0xffffffffull || (!s.empty() && s[0] == '-')) throw std::out_of_range("u32");
    return (u32)v;
  }
};

}  // namespace mpmh
