/* mpm_b200 — C ABI of the B200-native MLS-MPM substep.
 *
 * Drop-in boundary for the hot path of kekeblom/mpm: everything Simulation::advance()
 * (reference src/mpm.cu:323-329) does on the device, plus the buffer management around it
 * (src/mpm.cu:197-215, 278-314).  Plain pointers and sizes only; no C++ or torch types cross
 * this boundary.  All int-returning functions return 0 on success and a non-zero CUDA/NCCL
 * derived code otherwise (text via mpm_last_error); no exceptions cross the ABI.  One handle =
 * one device = one host thread at a time; distinct handles are independent.
 *
 * Data crossing the boundary keeps the reference layouts:
 *   particle  = MLS_APIC_Particle, 104 bytes AoS: u8 material_type (+3 pad), x[3], v[3],
 *               F[9] column-major, C[9] column-major, Jp
 *               (reference include/types.h:24-35, include/TransferScheme.h:46-54)
 *   material  = MMSnow<Particle>, 7 floats (reference include/MaterialModel.cuh:23-24,47-48,70-72)
 *   grid node = float4 (px|vx, py|vy, pz|vz, m), index N*N*i + N*j + k (reference src/mpm.cu:49,55,66)
 */
#ifndef MPM_B200_H
#define MPM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MPM_B200_ABI_VERSION 4

/* which MaterialModel alias the kernels are instantiated for (reference include/mpm.cuh:25) */
enum { MPM_MODEL_SNOW = 0, MPM_MODEL_FIXED_COROTATED = 1 };
/* svd3 arithmetic: EXACT reproduces the reference svd3 bit for bit; FAST contracts to FMA and
 * uses the hardware rsqrt approximation (deviation reported by the tests) */
enum { MPM_SVD_EXACT = 0, MPM_SVD_FAST = 1 };
/* P2G kernel: RUNS pre-reduces same-cell particles in registers before the vector reductions
 * (default); DIRECT issues 27 vector reductions per particle (also used when N > 1019) */
enum { MPM_P2G_RUNS = 0, MPM_P2G_DIRECT = 1 };
/* G2P kernel: TILE stages the grid block and the particle streams of each CTA in shared memory
 * with bulk async copies (default); DIRECT gathers the 27 nodes per particle from global memory */
enum { MPM_G2P_TILE = 0, MPM_G2P_DIRECT = 1 };
/* substep pipeline: OFF (default) runs reset -> P2G -> grid update -> G2P as separate kernels, as the
 * reference does.  G2P2G runs G2P of substep s and P2G of substep s+1 as ONE warp-specialised kernel
 * over two alternating grids, so the particle state written by G2P is never read back from HBM
 * (needs P2G_RUNS + G2P_TILE, else the separate kernels run).  Same arithmetic per particle either
 * way.  Measured slower than the separate kernels so far, hence not the default.
 * In G2P2G mode the internal grid holds the NEXT substep's velocities after mpm_advance. */
enum { MPM_FUSE_OFF = 0, MPM_FUSE_G2P2G = 1 };
/* stage indices for mpm_get_stage_times */
enum { MPM_STAGE_SORT = 0, MPM_STAGE_RESET = 1, MPM_STAGE_P2G = 2, MPM_STAGE_GRID = 3, MPM_STAGE_G2P = 4,
       MPM_STAGE_EXCHANGE = 5, MPM_STAGE_G2P2G = 6, MPM_STAGE_COUNT = 7 };

/* replaces the reference's 104-byte particle record at the boundary */
typedef struct MpmParticle {
  uint8_t material_type;
  uint8_t pad_[3];
  float x[3];
  float v[3];
  float F[9]; /* column-major */
  float C[9]; /* column-major */
  float Jp;
} MpmParticle;

/* replaces MMSnow<Particle> (trivially copyable, memcpy'd to the device, src/mpm.cu:198-201) */
typedef struct MpmMaterial {
  float particleVolume;
  float particleMass;
  float mu0;
  float lambda0;
  float hardening;
  float plast_clamp_lower;
  float plast_clamp_higher;
} MpmMaterial;

/* replaces SimulationParameters (include/TransferScheme.h:6-29) + the compile-time aliases */
typedef struct MpmParams {
  float dt;             /* --dt */
  uint32_t N;           /* --N: global grid is N^3, domain is the unit cube */
  uint32_t model;       /* MPM_MODEL_* */
  uint32_t svd_mode;    /* MPM_SVD_* */
  uint32_t sort_every;  /* re-bin (sort + permute) every this many substeps; 0 = never after upload */
  uint32_t x_begin;     /* slab decomposition: first owned x-plane (0 for a single device) */
  uint32_t x_end;       /* one past the last owned x-plane (N for a single device; 0 means N) */
  int32_t device;       /* CUDA device ordinal, -1 = current */
  uint64_t capacity;    /* particle slots to allocate (0 = size of the first upload) */
  uint32_t p2g_mode;    /* MPM_P2G_* */
  uint32_t ghost;       /* slab handles: extra ghost x-planes either side, i.e. how many cells a particle
                           may drift out of its slab between re-bins (0 = default: 1 for slabs) */
  uint32_t g2p_mode;    /* MPM_G2P_* */
  uint32_t fuse_mode;   /* MPM_FUSE_* */
  uint32_t rebin_permille; /* 0 = fixed cadence only.  > 0: also re-bin as soon as the cell crossings counted by G2P
                           since the last re-bin exceed this many per mille of the particle count (needs
                           G2P_TILE, separate kernels, a single-device handle — slab handles keep the fixed
                           cadence, their ranks must re-bin in the same substep); sort_every = 0 then means
                           "only on demand" */
  uint32_t reserved_;   /* must be 0 */
} MpmParams;

typedef struct MpmSim MpmSim;

/* MaterialModel constructor arithmetic (MaterialModel.cuh:26-29, 50-54, 74-84 as called from
 * src/main.cu:36-42); host only */
void mpm_make_material(double volume, double density, double E, double Nu, double hardening,
                       double plast_clamp_lower, double plast_clamp_higher, MpmMaterial* out);

/* Simulation::Simulation + initCuda() minus the particle upload (src/mpm.cu:180-207) */
int mpm_create(const MpmParams* params, const MpmMaterial* materials, int n_materials, MpmSim** out);
/* Simulation::~Simulation (src/mpm.cu:190-195) */
void mpm_destroy(MpmSim* sim);
const char* mpm_last_error(const MpmSim* sim); /* sim may be NULL: last creation error */
int mpm_abi_version(void);

/* particlesToDevice (src/mpm.cu:278-286): host AoS -> device SoA; replaces the active set */
int mpm_upload_particles_aos(MpmSim* sim, const MpmParticle* particles, size_t count);
/* slab handles (multi-GPU): same, with caller-chosen global particle ids; such handles return
 * particles in their current (cell-sorted) order, ids via mpm_debug_download_sort */
int mpm_upload_particles_with_ids(MpmSim* sim, const MpmParticle* particles, const uint32_t* ids, size_t count);
/* adds particles to the active set on the device (an object whose lifetime begins, include/mpm.cuh:36-40;
 * the reference re-uploads everything from stale host copies, src/mpm.cu:298-314).  Their ids continue
 * the upload order.  Needs room: set MpmParams.capacity to the total over all objects. */
int mpm_append_particles_aos(MpmSim* sim, const MpmParticle* particles, size_t count);
/* particlesToHost (src/mpm.cu:288-306): blocking; particles come back in upload order */
int mpm_download_particles_aos(MpmSim* sim, MpmParticle* particles, size_t capacity, size_t* count);
/* positions only (12 B/particle), upload order — what the reference's viewer cadence needs */
int mpm_download_positions(MpmSim* sim, float* xyz, size_t capacity, size_t* count);
/* the same without blocking: the copy is queued behind the substeps issued so far and overlaps the
 * ones issued afterwards only if xyz is pinned host memory; read it after mpm_sync().  This is the
 * reference's viewer cadence (src/main.cu:99-102) without stalling the substep pipeline.  The staging
 * buffer is shared with the other transfers: one transfer in flight at a time. */
int mpm_download_positions_async(MpmSim* sim, float* xyz, size_t capacity, size_t* count);
/* synthetic dense block generated on the device: ids
 * [first_id, first_id+count), x = lo + (hi-lo)*u(hash(seed,id,axis)), v=0, F=I, C=0, Jp=1;
 * only particles whose base node lies in this handle's slab are kept */
int mpm_generate_dense_block(MpmSim* sim, uint64_t first_id, uint64_t count, uint32_t seed, float lo, float hi,
                             uint8_t material);
size_t mpm_particle_count(const MpmSim* sim);

/* Simulation::advance() x n_substeps (src/mpm.cu:323-329); asynchronous like the reference */
int mpm_advance(MpmSim* sim, int n_substeps);
/* block until the device is idle (what syncDevice's cudaMemcpy does implicitly, src/mpm.cu:289) */
int mpm_sync(MpmSim* sim);
double mpm_time(const MpmSim* sim);           /* Simulation::t */
uint64_t mpm_substeps_done(const MpmSim* sim);
uint64_t mpm_kernel_launches(const MpmSim* sim); /* kernels this handle has launched so far */
uint64_t mpm_rebins_done(const MpmSim* sim);     /* re-bins (sort + permute) so far, the one at upload included */

/* single stages, for parity tests and profiling (same kernels mpm_advance runs) */
int mpm_stage_sort(MpmSim* sim);        /* north-star stage (1): cell keys, radix sort, SoA permute */
int mpm_stage_reset_grid(MpmSim* sim);  /* resetGrid, src/mpm.cu:213-215 */
int mpm_stage_p2g(MpmSim* sim);         /* particleToGrid, src/mpm.cu:14-74 */
int mpm_stage_grid_update(MpmSim* sim); /* gridOpKernel, src/mpm.cu:76-107 */
int mpm_stage_g2p(MpmSim* sim);         /* gridToParticle, src/mpm.cu:109-178 */

/* debug access to internal state (parity tests) */
int mpm_debug_download_grid(MpmSim* sim, float* vec4, size_t n_nodes);       /* local slab incl. ghost planes */
int mpm_debug_upload_grid(MpmSim* sim, const float* vec4, size_t n_nodes);
/* replaces the particle data in place (particles in upload order, same count) WITHOUT re-binning:
 * tests use it to present the kernels with a stale cell order */
int mpm_debug_overwrite_particles_aos(MpmSim* sim, const MpmParticle* particles, size_t count);
int mpm_debug_download_sort(MpmSim* sim, uint32_t* keys, uint32_t* ids, size_t capacity); /* current order */
size_t mpm_grid_nodes(const MpmSim* sim);

/* accumulated CUDA-event time per stage since the last call (ms); enables timing on first call */
int mpm_get_stage_times(MpmSim* sim, float ms[MPM_STAGE_COUNT]);
/* stream the handle launches on (cudaStream_t), for callers that time with their own events */
void* mpm_stream(MpmSim* sim);

/* multi-GPU (one handle per rank, slabs along x): NCCL communicator from a unique id that the
 * host launcher distributes (128 bytes, mpm_comm_unique_id on rank 0) */
int mpm_comm_unique_id(void* id128);
int mpm_attach_comm(MpmSim* sim, const void* id128, int rank, int nranks);

/* linalg known-answer hooks (reference tests/test_linalg.cu:49-91): row-major 3x3 batches on the device */
int mpm_svd3_batch(const float* A, float* U, float* S, float* V, size_t n, int svd_mode);
int mpm_polar_batch(const float* A, float* R, size_t n, int svd_mode);
int mpm_determinant_batch(const float* A, float* det, size_t n);

#ifdef __cplusplus
}
#endif
#endif /* MPM_B200_H */
