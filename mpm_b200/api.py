"""ctypes binding of include/mpm_b200.h (one method per C entry point)."""
import ctypes
import os

import numpy as np

from . import build as _build

SNOW, FIXED_COROTATED, JELLY, MODEL_USER = 0, 1, 2, 16
SVD_EXACT, SVD_FAST = 0, 1
P2G_RUNS, P2G_DIRECT = 0, 1
G2P_TILE, G2P_DIRECT = 0, 1
PIPE_HANDOVER, PIPE_CLASSIC = 0, 1
GRAPH_AUTO, GRAPH_OFF, GRAPH_ON = 0, 1, 2
STAGES = ("sort", "reset", "p2g", "grid", "g2p", "exchange")

# MpmParticle == the reference's MLS_APIC_Particle (104 bytes, matrices column-major)
PARTICLE_DTYPE = np.dtype(
    [("material_type", "u1"), ("pad", "u1", 3), ("x", "f4", 3), ("v", "f4", 3), ("F", "f4", 9), ("C", "f4", 9),
     ("Jp", "f4")]
)
MATERIAL_DTYPE = np.dtype([(n, "f4") for n in ("particleVolume", "particleMass", "mu0", "lambda0", "hardening",
                                               "plast_clamp_lower", "plast_clamp_higher")])


class MpmParams(ctypes.Structure):
    _fields_ = [("dt", ctypes.c_float), ("N", ctypes.c_uint32), ("model", ctypes.c_uint32),
                ("svd_mode", ctypes.c_uint32), ("sort_every", ctypes.c_uint32), ("x_begin", ctypes.c_uint32),
                ("x_end", ctypes.c_uint32), ("device", ctypes.c_int32), ("capacity", ctypes.c_uint64),
                ("p2g_mode", ctypes.c_uint32), ("ghost", ctypes.c_uint32), ("g2p_mode", ctypes.c_uint32),
                ("pipeline", ctypes.c_uint32), ("rebin_permille", ctypes.c_uint32),
                ("graph_mode", ctypes.c_uint32)]


class MpmError(RuntimeError):
    pass


_lib = None
_vp = ctypes.c_void_p
_fp = ctypes.POINTER(ctypes.c_float)


def _preload_nccl():
    """The library depends on libnccl.so.2 by soname.  A PyTorch imported later in the same process needs the NCCL
    it ships with (newer than the system one), and the first libnccl.so.2 loaded wins — so load that one first
    when it is installed; otherwise the system library is found the usual way."""
    import importlib.util

    try:
        spec = importlib.util.find_spec("nvidia")
        for base in (spec.submodule_search_locations if spec else []):
            cand = os.path.join(base, "nccl", "lib", "libnccl.so.2")
            if os.path.exists(cand):
                ctypes.CDLL(cand, mode=ctypes.RTLD_GLOBAL)
                return
    except Exception:
        pass


def lib():
    """Loads (building if needed) libmpm_b200.so.  Raises if it cannot be had — no fallback."""
    global _lib
    if _lib is None:
        path = os.environ.get("MPM_B200_LIB") or _build.build()  # MPM_B200_LIB: experiment builds (tools/ab.py)
        _preload_nccl()
        L = ctypes.CDLL(path)
        L.mpm_last_error.restype = ctypes.c_char_p
        L.mpm_last_error.argtypes = [_vp]
        L.mpm_particle_count.restype = ctypes.c_size_t
        L.mpm_particle_count.argtypes = [_vp]
        L.mpm_grid_nodes.restype = ctypes.c_size_t
        L.mpm_grid_nodes.argtypes = [_vp]
        L.mpm_time.restype = ctypes.c_double
        L.mpm_time.argtypes = [_vp]
        L.mpm_substeps_done.restype = ctypes.c_uint64
        L.mpm_substeps_done.argtypes = [_vp]
        L.mpm_kernel_launches.restype = ctypes.c_uint64
        L.mpm_kernel_launches.argtypes = [_vp]
        L.mpm_rebins_done.restype = ctypes.c_uint64
        L.mpm_rebins_done.argtypes = [_vp]
        L.mpm_graph_replays.restype = ctypes.c_uint64
        L.mpm_graph_replays.argtypes = [_vp]
        L.mpm_merge_rebins.restype = ctypes.c_uint64
        L.mpm_merge_rebins.argtypes = [_vp]
        L.mpm_stream.restype = _vp
        L.mpm_stream.argtypes = [_vp]
        L.mpm_destroy.restype = None
        L.mpm_destroy.argtypes = [_vp]
        L.mpm_make_material.restype = None
        L.mpm_make_material.argtypes = [ctypes.c_double] * 7 + [_vp]
        L.mpm_create.argtypes = [ctypes.POINTER(MpmParams), _vp, ctypes.c_int, ctypes.POINTER(_vp)]
        L.mpm_create_raw.argtypes = [ctypes.POINTER(MpmParams), _vp, ctypes.c_size_t, ctypes.c_int, ctypes.POINTER(_vp)]
        L.mpm_generate_dense_block_stressed.argtypes = [_vp, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_float,
                                                        ctypes.c_float, ctypes.c_uint8, ctypes.c_float, ctypes.c_float]
        L.mpm_set_stage_timing.argtypes = [_vp, ctypes.c_int]
        L.mpm_get_diagnostics.argtypes = [_vp, _vp]
        L.mpm_dinv_batch.argtypes = [_vp, _vp, ctypes.c_size_t, ctypes.c_uint32]
        L.mpm_upload_particles_aos.argtypes = [_vp, _vp, ctypes.c_size_t]
        L.mpm_append_particles_aos.argtypes = [_vp, _vp, ctypes.c_size_t]
        L.mpm_upload_particles_with_ids.argtypes = [_vp, _vp, _vp, ctypes.c_size_t]
        L.mpm_download_particles_aos.argtypes = [_vp, _vp, ctypes.c_size_t, ctypes.POINTER(ctypes.c_size_t)]
        L.mpm_remove_particles.argtypes = [_vp, ctypes.c_size_t, ctypes.c_size_t]
        L.mpm_prefetch_particles_aos.argtypes = [_vp, _vp, ctypes.c_size_t]
        L.mpm_download_particles_aos_async.argtypes = [_vp, _vp, ctypes.c_size_t, ctypes.POINTER(ctypes.c_size_t)]
        L.mpm_download_wait.argtypes = [_vp]
        L.mpm_download_positions.argtypes = [_vp, _vp, ctypes.c_size_t, ctypes.POINTER(ctypes.c_size_t)]
        L.mpm_download_positions_async.argtypes = [_vp, _vp, ctypes.c_size_t, ctypes.POINTER(ctypes.c_size_t)]
        L.mpm_generate_dense_block.argtypes = [_vp, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_float,
                                               ctypes.c_float, ctypes.c_uint8]
        L.mpm_advance.argtypes = [_vp, ctypes.c_int]
        for n in ("mpm_sync", "mpm_stage_sort", "mpm_stage_reset_grid", "mpm_stage_p2g", "mpm_stage_grid_update",
                  "mpm_stage_g2p"):
            getattr(L, n).argtypes = [_vp]
        L.mpm_debug_download_grid.argtypes = [_vp, _vp, ctypes.c_size_t]
        L.mpm_debug_upload_grid.argtypes = [_vp, _vp, ctypes.c_size_t]
        L.mpm_debug_overwrite_particles_aos.argtypes = [_vp, _vp, ctypes.c_size_t]
        L.mpm_debug_download_sort.argtypes = [_vp, _vp, _vp, ctypes.c_size_t]
        L.mpm_get_stage_times.argtypes = [_vp, _vp]
        L.mpm_comm_unique_id.argtypes = [_vp]
        L.mpm_attach_comm.argtypes = [_vp, _vp, ctypes.c_int, ctypes.c_int]
        L.mpm_svd3_batch.argtypes = [_vp, _vp, _vp, _vp, ctypes.c_size_t, ctypes.c_int]
        L.mpm_polar_batch.argtypes = [_vp, _vp, ctypes.c_size_t, ctypes.c_int]
        L.mpm_determinant_batch.argtypes = [_vp, _vp, ctypes.c_size_t]
        _lib = L
    return _lib


def _ptr(a):
    return a.ctypes.data_as(_vp)


def make_material(volume, density=700.0, E=1.4e5, Nu=0.2, hardening=10.0, plast_clamp_lower=0.975,
                  plast_clamp_higher=1.0075):
    """MaterialModel(volume, density, E, Nu, hardening, lo, hi) as src/main.cu:36-42 builds it."""
    out = np.zeros(7, np.float32)
    lib().mpm_make_material(volume, density, E, Nu, hardening, plast_clamp_lower, plast_clamp_higher, _ptr(out))
    return out


class Sim:
    """One handle = one device.  Mirrors the device half of the reference's Simulation class."""

    def __init__(self, N, dt, materials, model=SNOW, svd_mode=SVD_EXACT, sort_every=0, x_begin=0, x_end=0,
                 device=-1, capacity=0, p2g_mode=P2G_RUNS, ghost=0, g2p_mode=G2P_TILE, pipeline=PIPE_HANDOVER, rebin_permille=0,
                 raw_materials=None, graph_mode=0):
        """materials: n x 7 floats (MpmMaterial = MMSnow's fields; the leading fields are used by the
        other shipped models).  raw_materials: bytes of n objects of a registered model's own
        material type instead (mpm_create_raw, user-defined materials)."""
        self._h = _vp()
        self.params = MpmParams(dt, N, model, svd_mode, sort_every, x_begin, x_end, device, capacity, p2g_mode, ghost, g2p_mode, pipeline, rebin_permille, graph_mode)
        if raw_materials is not None:
            raw, n = raw_materials
            buf = ctypes.create_string_buffer(bytes(raw), len(raw))
            rc = lib().mpm_create_raw(ctypes.byref(self.params), buf, len(raw) // n, n, ctypes.byref(self._h))
        else:
            mats = np.ascontiguousarray(materials, np.float32).reshape(-1, 7)
            rc = lib().mpm_create(ctypes.byref(self.params), _ptr(mats), mats.shape[0], ctypes.byref(self._h))
        if rc:
            raise MpmError(lib().mpm_last_error(None).decode())
        self.N = N
        self.dt = dt

    def _ck(self, rc):
        if rc:
            raise MpmError(lib().mpm_last_error(self._h).decode())

    def close(self):
        if self._h:
            lib().mpm_destroy(self._h)
            self._h = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # --- buffers (initCuda / particlesToDevice / particlesToHost) ---
    def upload(self, particles):
        assert particles.dtype == PARTICLE_DTYPE and particles.flags.c_contiguous
        self._ck(lib().mpm_upload_particles_aos(self._h, _ptr(particles), particles.shape[0]))

    def append(self, particles):
        """More particles into the active set (an object whose lifetime begins); ids continue."""
        assert particles.dtype == PARTICLE_DTYPE and particles.flags.c_contiguous
        self._ck(lib().mpm_append_particles_aos(self._h, _ptr(particles), particles.shape[0]))

    def remove(self, first, count):
        """Drops the particles at positions [first, first + count) of the upload order (an object whose lifetime ends)."""
        self._ck(lib().mpm_remove_particles(self._h, first, count))

    def overwrite(self, particles):
        """New particle data (upload order) into the existing slots, without re-binning."""
        assert particles.dtype == PARTICLE_DTYPE and particles.flags.c_contiguous
        self._ck(lib().mpm_debug_overwrite_particles_aos(self._h, _ptr(particles), particles.shape[0]))

    def upload_with_ids(self, particles, ids):
        assert particles.dtype == PARTICLE_DTYPE and particles.flags.c_contiguous
        ids = np.ascontiguousarray(ids, np.uint32)
        self._ck(lib().mpm_upload_particles_with_ids(self._h, _ptr(particles), _ptr(ids), particles.shape[0]))

    def upload_ptr(self, ptr, count):
        self._ck(lib().mpm_upload_particles_aos(self._h, _vp(ptr), count))

    def download(self, out=None):
        n = self.count
        if out is None:
            out = np.empty(n, PARTICLE_DTYPE)
        cnt = ctypes.c_size_t()
        self._ck(lib().mpm_download_particles_aos(self._h, _ptr(out), out.shape[0], ctypes.byref(cnt)))
        return out[: cnt.value]

    def download_ptr(self, ptr, capacity):
        cnt = ctypes.c_size_t()
        self._ck(lib().mpm_download_particles_aos(self._h, _vp(ptr), capacity, ctypes.byref(cnt)))
        return cnt.value

    def prefetch_ptr(self, ptr, count):
        """Starts the host -> device copy of a pinned AoS buffer; a later upload_ptr of the same buffer uses it."""
        self._ck(lib().mpm_prefetch_particles_aos(self._h, _vp(ptr), count))

    def download_ptr_async(self, ptr, capacity):
        """Queues the read-back into a pinned AoS buffer; valid after download_wait()."""
        cnt = ctypes.c_size_t()
        self._ck(lib().mpm_download_particles_aos_async(self._h, _vp(ptr), capacity, ctypes.byref(cnt)))
        return cnt.value

    def download_wait(self):
        self._ck(lib().mpm_download_wait(self._h))

    def download_positions(self):
        out = np.empty((self.count, 3), np.float32)
        cnt = ctypes.c_size_t()
        self._ck(lib().mpm_download_positions(self._h, _ptr(out), out.shape[0], ctypes.byref(cnt)))
        return out

    def download_positions_async(self, out):
        """Queues the positions read-back into `out` (float32 [count, 3], ideally pinned); valid after sync()."""
        cnt = ctypes.c_size_t()
        self._ck(lib().mpm_download_positions_async(self._h, _ptr(out), out.shape[0], ctypes.byref(cnt)))
        return cnt.value

    def generate_dense_block(self, count, seed=1234, lo=0.1, hi=0.9, material=0, first_id=0, shear=0.0, f_noise=0.0):
        self._ck(lib().mpm_generate_dense_block_stressed(self._h, first_id, count, seed, lo, hi, material, shear, f_noise))

    @property
    def count(self):
        return lib().mpm_particle_count(self._h)

    @property
    def grid_nodes(self):
        return lib().mpm_grid_nodes(self._h)

    @property
    def t(self):
        return lib().mpm_time(self._h)

    @property
    def launches(self):
        return lib().mpm_kernel_launches(self._h)

    @property
    def rebins(self):
        return lib().mpm_rebins_done(self._h)

    @property
    def graph_replays(self):
        return lib().mpm_graph_replays(self._h)

    @property
    def merge_rebins(self):
        return lib().mpm_merge_rebins(self._h)

    @property
    def stream(self):
        return lib().mpm_stream(self._h)

    # --- substep ---
    def advance(self, n=1):
        self._ck(lib().mpm_advance(self._h, n))

    def sync(self):
        self._ck(lib().mpm_sync(self._h))

    def stage(self, name):
        self._ck(getattr(lib(), "mpm_stage_" + name)(self._h))

    def grid(self):
        n = self.grid_nodes
        g = np.empty((n // (self.N * self.N), self.N, self.N, 4), np.float32)
        self._ck(lib().mpm_debug_download_grid(self._h, _ptr(g), n))
        return g

    def set_grid(self, g):
        g = np.ascontiguousarray(g, np.float32)
        self._ck(lib().mpm_debug_upload_grid(self._h, _ptr(g), g.size // 4))

    def sort_state(self):
        n = self.count
        keys, ids = np.empty(n, np.uint32), np.empty(n, np.uint32)
        self._ck(lib().mpm_debug_download_sort(self._h, _ptr(keys), _ptr(ids), n))
        return keys, ids

    def stage_times(self):
        ms = np.zeros(len(STAGES), np.float32)
        self._ck(lib().mpm_get_stage_times(self._h, _ptr(ms)))
        return dict(zip(STAGES, ms.tolist()))

    def set_stage_timing(self, on):
        self._ck(lib().mpm_set_stage_timing(self._h, 1 if on else 0))

    def diagnostics(self):
        d = np.zeros(4, np.uint32)
        self._ck(lib().mpm_get_diagnostics(self._h, _ptr(d)))
        return dict(zip(("jp_not_one", "escaped", "nonfinite", "out_of_domain"), (int(v) for v in d)))

    def attach_comm(self, unique_id, rank, nranks):
        self._ck(lib().mpm_attach_comm(self._h, ctypes.c_char_p(unique_id), rank, nranks))


def comm_unique_id():
    buf = ctypes.create_string_buffer(128)
    if lib().mpm_comm_unique_id(buf):
        raise MpmError("mpm_comm_unique_id failed")
    return buf.raw


def svd3_batch(A, mode=SVD_EXACT):
    A = np.ascontiguousarray(A, np.float32).reshape(-1, 9)
    n = A.shape[0]
    U, S, V = np.empty((n, 9), np.float32), np.empty((n, 3), np.float32), np.empty((n, 9), np.float32)
    if lib().mpm_svd3_batch(_ptr(A), _ptr(U), _ptr(S), _ptr(V), n, mode):
        raise MpmError("mpm_svd3_batch failed (no GPU?)")
    return U.reshape(n, 3, 3), S, V.reshape(n, 3, 3)


def polar_batch(A, mode=SVD_EXACT):
    A = np.ascontiguousarray(A, np.float32).reshape(-1, 9)
    R = np.empty_like(A)
    if lib().mpm_polar_batch(_ptr(A), _ptr(R), A.shape[0], mode):
        raise MpmError("mpm_polar_batch failed (no GPU?)")
    return R.reshape(-1, 3, 3)


def dinv_batch(x, N):
    """D^-1 through the generic D_inv of the interpolation kernel at positions x (n x 3): n x 3 x 3."""
    x = np.ascontiguousarray(x, np.float32).reshape(-1, 3)
    out = np.empty((x.shape[0], 9), np.float32)
    if lib().mpm_dinv_batch(_ptr(x), _ptr(out), x.shape[0], N):
        raise MpmError("mpm_dinv_batch failed (no GPU?)")
    return out.reshape(-1, 3, 3)


def determinant_batch(A):
    A = np.ascontiguousarray(A, np.float32).reshape(-1, 9)
    d = np.empty(A.shape[0], np.float32)
    if lib().mpm_determinant_batch(_ptr(A), _ptr(d), A.shape[0]):
        raise MpmError("mpm_determinant_batch failed (no GPU?)")
    return d
