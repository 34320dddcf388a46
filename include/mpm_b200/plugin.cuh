// mpm_b200 plugin surface — compiling a user-defined material (or a whole transfer tuple) in.
//
// The reference selects its plugins with four compile-time aliases (include/mpm.cuh:24-27) and is
// rebuilt to change them.  Here the substep kernels are templates over the same concepts
// (MaterialModel.cuh, InterpolationKernel.cuh, TransferScheme.cuh) and one macro instantiates them for
// a new type and registers the result under a model id:
//
//     // my_material.cu
//     #include <mpm_b200/plugin.cuh>
//     template <class Particle> class MyMaterial : public MaterialModelBase<Particle> { ... computePF ... endOfStepMutation ... };
//     MPM_B200_REGISTER_MATERIAL(16, MyMaterial<MLS_APIC_Particle>)
//
// and the file is compiled and linked with the library's own sources (mpm_b200/build.py:
// build(extra_sources=[...]); tests/plugin/user_material.cu is built and tested that way).  Handles are
// then created with MpmParams.model = 16 and the material objects passed as raw bytes
// (mpm_create_raw: what the reference memcpy's to the device, src/mpm.cu:198-201).
//
// A material registered this way runs through the staged production kernels (P2G with run
// pre-reduction, bulk-copy-fed G2P, hand-over).  A different InterpolationKernel / TransferScheme runs
// through the generic one-thread-per-particle kernels (mpm_b200/csrc/kernels.cuh), which call the
// concept methods node by node exactly like the reference's kernels.
#pragma once
#include "../../mpm_b200/csrc/substep.cuh"

#define MPM_B200_PP_CAT2(a, b) a##b
#define MPM_B200_PP_CAT(a, b) MPM_B200_PP_CAT2(a, b)

// ID >= MPM_MODEL_USER; the same type serves both MpmParams.svd_mode values
#define MPM_B200_REGISTER_MATERIAL(ID, ...)                                                              \
  namespace {                                                                                            \
  struct MPM_B200_PP_CAT(MpmB200Registrar, ID) {                                                         \
    MPM_B200_PP_CAT(MpmB200Registrar, ID)() {                                                            \
      static_assert((ID) >= MPM_MODEL_USER, "model ids below MPM_MODEL_USER belong to the shipped materials"); \
      const mpm::ModelOps* o = mpm::ModelImpl<__VA_ARGS__>::ops(#__VA_ARGS__);                           \
      mpm::register_model((ID), o, o);                                                                   \
    }                                                                                                    \
  } MPM_B200_PP_CAT(g_mpm_b200_registrar, ID);                                                           \
  }

// a material template with the policy parameter of the shipped ones: Tmpl<Particle, mpm::ExactOps> for
// MPM_SVD_EXACT handles, Tmpl<Particle, mpm::FastOps> for MPM_SVD_FAST
#define MPM_B200_REGISTER_MATERIAL_TEMPLATE(ID, Tmpl)                                                    \
  namespace {                                                                                            \
  struct MPM_B200_PP_CAT(MpmB200Registrar, ID) {                                                         \
    MPM_B200_PP_CAT(MpmB200Registrar, ID)() {                                                            \
      static_assert((ID) >= MPM_MODEL_USER, "model ids below MPM_MODEL_USER belong to the shipped materials"); \
      mpm::register_model((ID), mpm::ModelImpl<Tmpl<MLS_APIC_Particle, mpm::ExactOps>>::ops(#Tmpl "<ExactOps>"),       \
                          mpm::ModelImpl<Tmpl<MLS_APIC_Particle, mpm::FastOps>>::ops(#Tmpl "<FastOps>"));             \
    }                                                                                                    \
  } MPM_B200_PP_CAT(g_mpm_b200_registrar, ID);                                                           \
  }

// a whole tuple (material, interpolation kernel, transfer scheme): generic kernels only
#define MPM_B200_REGISTER_TUPLE(ID, Material, Kernel, Scheme)                                            \
  namespace {                                                                                            \
  struct MPM_B200_PP_CAT(MpmB200Registrar, ID) {                                                         \
    MPM_B200_PP_CAT(MpmB200Registrar, ID)() {                                                            \
      static_assert((ID) >= MPM_MODEL_USER, "model ids below MPM_MODEL_USER belong to the shipped materials"); \
      const mpm::ModelOps* o = mpm::ModelImpl<Material, Kernel, Scheme>::ops(#Material);                 \
      mpm::register_model((ID), o, o);                                                                   \
    }                                                                                                    \
  } MPM_B200_PP_CAT(g_mpm_b200_registrar, ID);                                                           \
  }
