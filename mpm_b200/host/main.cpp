// mpm_b200_cli: the reference's main() (src/main.cu:22-115) over the B200-native substep.
// Same flags (options.hpp), same scene files, same loop: advance(); every 20 substeps syncDevice();
// with --save-dir, every 1 / 240 / dt substeps a surface mesh meshes/mesh_%05d.obj and a particle
// dump particles/particles_%d.bgeo (--particle-format pda: Partio's ASCII flavour).  Headless (the GLFW viewer is out of scope); --steps bounds the
// loop, which the reference runs until it is killed.
#include <sys/stat.h>

#include <future>
#include <iomanip>
#include <memory>
#include <sstream>

#include "output.hpp"
#include "scene.hpp"

using namespace mpmh;

const unsigned int FrameRateDefault = 240;  // for export of data
std::vector<MaterialModel> material_models;

int main(int argc, char* argv[]) {
  CLIOptions flags(argc, argv);
  try {
    Simulation simulation(flags, InterpolationKernel(), material_models);
    std::cout << "Loading scene file: " << flags.scene << std::endl;
    load_scene(flags, material_models, simulation, true);
    std::cout << "loaded materials" << std::endl;
    for (size_t o = 0; o < simulation.objects.size(); ++o)
      if (simulation.objects[o].substituted_mesh)
        std::cout << "object " << o << ": mesh file is a Git-LFS stub, procedural stand-in used" << std::endl;
    simulation.initCuda();

    ParticleWriter writer;
    MeshBuilder mesher(simulation.par, flags, flags.mesh_grid);
    const bool save = flags.save_dir != "";
    if (save) {
      mkdir(flags.save_dir.c_str(), 0777);
      mkdir((flags.save_dir + "/meshes").c_str(), 0777);
      mkdir((flags.save_dir + "/particles").c_str(), 0777);
    }
    const u32 frame_rate = flags.frame_rate ? flags.frame_rate : FrameRateDefault;
    const u32 save_every = std::max<u32>(1, u32(1. / float(frame_rate) / flags.dt));
    u32 frame_id = 0;
    std::future<void> output;  // the frame being written
    const unsigned long long n_steps = flags.steps >= 0 ? (unsigned long long)flags.steps : std::numeric_limits<u32>::max();
    for (unsigned long long i = 0; i < n_steps; i++) {
      if (i % 10 == 0) std::cout << "Step " << i << "\r" << std::flush;
      simulation.advance();
      const bool frame = save && (i % save_every) == 0;
      if (i % flags.sync_every == 0 || frame) simulation.syncDevice();
      if (frame) {
        // the frame's files are written from a snapshot by a second host thread, so that the substeps of the next
        // frame (queued asynchronously by advance()) and the meshing / file I/O of this one overlap; one frame in flight
        std::stringstream ss;
        ss << flags.save_dir << "/meshes/mesh_" << std::setfill('0') << std::setw(5) << frame_id << ".obj";
        const std::string mesh_path = ss.str();
        ss.str("");
        ss.clear();
        ss << flags.save_dir << "/particles/particles_" << frame_id << "." << flags.particle_format;
        const std::string particle_path = ss.str();
        if (output.valid()) output.get();
        auto snapshot = std::make_shared<std::vector<Particle>>(simulation.getActiveParticleList());
        output = std::async(std::launch::async, [&mesher, &writer, snapshot, mesh_path, particle_path] {
          mesher.computeMesh(mesh_path, *snapshot);
          writer.writeParticles(particle_path, *snapshot);
        });
        frame_id++;
      }
    }
    if (output.valid()) output.get();
    simulation.syncDevice();
    std::cout << "\ndone: " << n_steps << " substeps, t = " << simulation.t << ", " << simulation.getActiveParticleList().size() << " active particles"
              << std::endl;
  } catch (const std::exception& e) {
    std::cerr << "error: " << e.what() << std::endl;
    return 2;
  }
  return 0;
}
