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

#define MPM_B200_ABI_VERSION 6

/* which MaterialModel alias the kernels are instantiated for (reference include/mpm.cuh:25): the classes of
 * include/mpm_b200/MaterialModel.cuh.  Ids >= MPM_MODEL_USER are materials compiled in through
 * include/mpm_b200/plugin.cuh (MPM_B200_REGISTER_MATERIAL). */
enum { MPM_MODEL_SNOW = 0, MPM_MODEL_FIXED_COROTATED = 1, MPM_MODEL_JELLY = 2, MPM_MODEL_USER = 16 };
/* svd3 arithmetic: EXACT reproduces the reference svd3 bit for bit; FAST contracts to FMA and
 * uses the hardware rsqrt approximation (deviation reported by the tests) */
enum { MPM_SVD_EXACT = 0, MPM_SVD_FAST = 1 };
/* P2G kernel: RUNS pre-reduces same-cell particles in registers before the vector reductions
 * (default); DIRECT is the generic kernel over the plugin concepts, 27 vector reductions per particle
 * (also used when N > 1019 and for user-defined transfer schemes / interpolation kernels) */
enum { MPM_P2G_RUNS = 0, MPM_P2G_DIRECT = 1 };
/* G2P kernel: TILE stages the particle streams of each CTA in shared memory with bulk async copies
 * (default); DIRECT is the generic kernel over the plugin concepts */
enum { MPM_G2P_TILE = 0, MPM_G2P_DIRECT = 1 };
/* substep pipeline.  HANDOVER (default): between two substeps of ONE mpm_advance call, G2P computes the
 * affine matrix of the next P2G (stress + m C) while F, C, Jp are in its registers and stores it in place
 * of C; that P2G then reads 15 of the 25 particle streams and does not evaluate the material.  The last
 * G2P of every mpm_advance call stores C, so the particle state is always the reference's at the API
 * boundary.  Same arithmetic per particle either way.  CLASSIC: every P2G evaluates the material
 * itself, as the reference does (also what DIRECT kernels and single-substep calls get). */
enum { MPM_PIPE_HANDOVER = 0, MPM_PIPE_CLASSIC = 1 };
/* CUDA graphs.  A small scene is launch-bound (a substep is 4 launches of a few microseconds each): the
 * launches of one mpm_advance(n) call are captured once per (n, phase of the re-bin cadence, buffer parity)
 * and replayed afterwards.  AUTO: when the handle holds at most 4 M particles; never with rebin_permille,
 * on slab handles or while stage timing is on (those take decisions on the host between the launches). */
enum { MPM_GRAPH_AUTO = 0, MPM_GRAPH_OFF = 1, MPM_GRAPH_ON = 2 };
/* stage indices for mpm_get_stage_times */
enum { MPM_STAGE_SORT = 0, MPM_STAGE_RESET = 1, MPM_STAGE_P2G = 2, MPM_STAGE_GRID = 3, MPM_STAGE_G2P = 4,
       MPM_STAGE_EXCHANGE = 5, MPM_STAGE_COUNT = 6 };

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
  uint32_t pipeline;    /* MPM_PIPE_* */
  uint32_t rebin_permille; /* 0 = fixed cadence only.  > 0: also re-bin as soon as the cell crossings counted by G2P
                           since the last re-bin exceed this many per mille of the particle count (needs
                           G2P_TILE, a single-device handle — slab handles keep the fixed
                           cadence, their ranks must re-bin in the same substep); sort_every = 0 then means
                           "only on demand" */
  uint32_t graph_mode;  /* MPM_GRAPH_*: replay mpm_advance calls as CUDA graphs */
} MpmParams;

typedef struct MpmSim MpmSim;

/* MaterialModel constructor arithmetic (MaterialModel.cuh:26-29, 50-54, 74-84 as called from
 * src/main.cu:36-42); host only */
void mpm_make_material(double volume, double density, double E, double Nu, double hardening,
                       double plast_clamp_lower, double plast_clamp_higher, MpmMaterial* out);

/* Simulation::Simulation + initCuda() minus the particle upload (src/mpm.cu:180-207) */
int mpm_create(const MpmParams* params, const MpmMaterial* materials, int n_materials, MpmSim** out);
/* the same for materials of any registered model: `materials` = n_materials trivially copyable objects
 * of the model's own type, material_bytes each — what the reference memcpy's to the device
 * (src/mpm.cu:198-201).  mpm_create is this with the 7-float MpmMaterial records converted. */
int mpm_create_raw(const MpmParams* params, const void* materials, size_t material_bytes, int n_materials, MpmSim** out);
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
/* removes the particles at positions [first, first + count) of the upload order (an object whose lifetime ends,
 * include/mpm.cuh:36-40; the reference rebuilds its device arrays from host copies, src/mpm.cu:298-314).  On
 * the device: a stable compaction that keeps the cell order; the particles behind the range move up in the
 * upload order, so later downloads return the survivors contiguously.  Whole-domain handles only. */
int mpm_remove_particles(MpmSim* sim, size_t first, size_t count);
/* particlesToHost (src/mpm.cu:288-306): blocking; particles come back in upload order */
int mpm_download_particles_aos(MpmSim* sim, MpmParticle* particles, size_t capacity, size_t* count);
/* Overlapped transfers for loops that stream particle sets through one handle (pinned host memory).  The reference copies and computes strictly in turn (particlesToDevice / advance /
 * particlesToHost on the default stream, src/mpm.cu:278-306); these keep its data formats and overlap the
 * PCIe copies of one particle set with the substeps of another:
 *   mpm_prefetch_particles_aos        starts the host -> device copy of `particles` into a second staging buffer
 *                                     on a copy stream and returns.  A later mpm_upload_particles_aos of the SAME
 *                                     pointer and count uses that copy (and waits for it on the device) instead
 *                                     of copying again; any other upload ignores it.  The host buffer must stay
 *                                     unchanged until that upload.
 *   mpm_download_particles_aos_async  like mpm_download_particles_aos, but the device -> host copy runs on a copy
 *                                     stream and the call returns; substeps and uploads issued afterwards overlap
 *                                     it.  `particles` is valid after mpm_download_wait.  One read-back in
 *                                     flight at a time: a second one queues behind the first. */
int mpm_prefetch_particles_aos(MpmSim* sim, const MpmParticle* particles, size_t count);
int mpm_download_particles_aos_async(MpmSim* sim, MpmParticle* particles, size_t capacity, size_t* count);
int mpm_download_wait(MpmSim* sim);
/* positions only (12 B/particle), upload order — what the reference's viewer cadence needs */
int mpm_download_positions(MpmSim* sim, float* xyz, size_t capacity, size_t* count);
/* the same without blocking: the copy is queued behind the substeps issued so far and runs on a copy stream
 * of its own, so the substeps issued afterwards overlap it (xyz must be pinned host memory for that); read
 * the buffer after mpm_sync() or mpm_download_wait().  This is the reference's viewer cadence
 * (src/main.cu:99-102) without stalling the substep pipeline.  The staging buffer is shared with the other
 * read-backs: a second one queues behind the first. */
int mpm_download_positions_async(MpmSim* sim, float* xyz, size_t capacity, size_t* count);
/* synthetic dense block generated on the device: ids
 * [first_id, first_id+count), x = lo + (hi-lo)*u(hash(seed,id,axis)), v=0, F=I, C=0, Jp=1;
 * only particles whose base node lies in this handle's slab are kept */
int mpm_generate_dense_block(MpmSim* sim, uint64_t first_id, uint64_t count, uint32_t seed, float lo, float hi,
                             uint8_t material);
/* the same block under stress (bench.py --stress): v = shear * (y - 0.5, 0, 0.3 (x - 0.5)) and
 * F = I + f_noise * u, u uniform in [-1, 1) from the same counter hash */
int mpm_generate_dense_block_stressed(MpmSim* sim, uint64_t first_id, uint64_t count, uint32_t seed, float lo, float hi,
                                      uint8_t material, float shear, float f_noise);
size_t mpm_particle_count(const MpmSim* sim);

/* Simulation::advance() x n_substeps (src/mpm.cu:323-329); asynchronous like the reference */
int mpm_advance(MpmSim* sim, int n_substeps);
/* block until the device is idle (what syncDevice's cudaMemcpy does implicitly, src/mpm.cu:289) */
int mpm_sync(MpmSim* sim);
double mpm_time(const MpmSim* sim);           /* Simulation::t */
uint64_t mpm_substeps_done(const MpmSim* sim);
uint64_t mpm_kernel_launches(const MpmSim* sim); /* kernels this handle has launched so far */
uint64_t mpm_rebins_done(const MpmSim* sim);     /* re-bins (sort + permute) so far, the one at upload included */
uint64_t mpm_graph_replays(const MpmSim* sim);   /* mpm_advance calls served by replaying a captured CUDA graph */
uint64_t mpm_merge_rebins(const MpmSim* sim);    /* re-bins done by merging (few particles changed cell) instead of the radix sort */

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

/* accumulated CUDA-event time per stage since the last call (ms).  The first call switches the per-stage
 * timing on (events + a host synchronisation around every stage: for profiling, not for production);
 * mpm_set_stage_timing(sim, 0) switches it off again */
int mpm_get_stage_times(MpmSim* sim, float ms[MPM_STAGE_COUNT]);
int mpm_set_stage_timing(MpmSim* sim, int on);
/* counters kept by the kernels (NaN / out-of-domain detection; ADVICE r1: slab escapes) */
typedef struct MpmDiagnostics {
  uint32_t jp_not_one;    /* 1 if any uploaded particle had Jp != 1 (fixed-corotated handles then read the Jp stream) */
  uint32_t escaped;       /* slab handles: particle-substeps whose stencil left the planes held locally (mass was lost);
                             a re-bin that finds this non-zero fails */
  uint32_t nonfinite;     /* particles with a non-finite position at the last re-bin */
  uint32_t out_of_domain; /* particles whose stencil lay wholly outside the domain at the last re-bin (frozen, like the reference) */
} MpmDiagnostics;
int mpm_get_diagnostics(MpmSim* sim, MpmDiagnostics* out); /* blocking */
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
/* D^-1 of the quadratic kernel through the generic D_inv (sum_nodes w d d^T, inverted) at n positions;
 * must equal 4 dx^-2 I (reference include/InterpolationKernel.cuh:21-50 vs :71-73) */
int mpm_dinv_batch(const float* xyz, float* Dinv9, size_t n, uint32_t N);

#ifdef __cplusplus
}
#endif
#endif /* MPM_B200_H */
