"""ctypes binding of the host front end (mpm_b200/host/*.hpp behind libmpm_b200_host.so): scene
loading and sampling, the Simulation facade, particle / mesh output.  Mirrors src/main.cu."""
import ctypes
import os

import numpy as np

from . import build as _build
from .api import MATERIAL_DTYPE, PARTICLE_DTYPE, MpmError, lib as _core_lib

_lib = None
_vp = ctypes.c_void_p


def lib():
    global _lib
    if _lib is None:
        _core_lib()  # libmpm_b200.so first: the host library links against it
        L = ctypes.CDLL(_build.build_host())
        L.mpmh_last_error.restype = ctypes.c_char_p
        L.mpmh_last_error.argtypes = [_vp]
        L.mpmh_scene_load.restype = _vp
        L.mpmh_scene_load.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_char_p), ctypes.c_int]
        L.mpmh_scene_free.argtypes = [_vp]
        L.mpmh_scene_free.restype = None
        for n in ("mpmh_n_materials", "mpmh_n_objects", "mpmh_init_cuda", "mpmh_sync_device", "mpmh_grid_size"):
            getattr(L, n).argtypes = [_vp]
        L.mpmh_get_materials.argtypes = [_vp, _vp]
        L.mpmh_object_count.restype = ctypes.c_size_t
        L.mpmh_object_count.argtypes = [_vp, ctypes.c_int]
        L.mpmh_object_substituted.argtypes = [_vp, ctypes.c_int]
        L.mpmh_object_lifetime.argtypes = [_vp, ctypes.c_int, _vp, _vp]
        for n in ("mpmh_full_count", "mpmh_active_count"):
            getattr(L, n).restype = ctypes.c_size_t
            getattr(L, n).argtypes = [_vp]
        L.mpmh_get_full.argtypes = [_vp, _vp]
        L.mpmh_get_active.argtypes = [_vp, _vp]
        L.mpmh_time.restype = ctypes.c_double
        L.mpmh_time.argtypes = [_vp]
        L.mpmh_advance.argtypes = [_vp, ctypes.c_int]
        L.mpmh_write_particles.argtypes = [_vp, ctypes.c_char_p]
        L.mpmh_compute_mesh.argtypes = [_vp, ctypes.c_char_p, _vp, _vp]
        L.mpmh_winding_numbers.argtypes = [_vp, ctypes.c_size_t, _vp, ctypes.c_size_t, _vp, ctypes.c_size_t, _vp]
        L.mpmh_load_mesh.argtypes = [ctypes.c_char_p, ctypes.c_double, _vp, _vp, _vp, _vp, _vp, _vp]
        L.mpmh_marching_tetrahedra.argtypes = [_vp, ctypes.c_int, _vp, ctypes.c_size_t, _vp, ctypes.c_size_t, _vp, _vp]
        _lib = L
    return _lib


def _ptr(a):
    return a.ctypes.data_as(_vp)


class Scene:
    """A loaded scene = CLIOptions + material models + a Simulation with every object sampled."""

    def __init__(self, *cli_args, seed=1):
        """cli_args: the command line of the reference binary, e.g. "--scene", path, "--N", "32"."""
        args = [b"mpm_b200_cli"] + [str(a).encode() for a in cli_args]
        argv = (ctypes.c_char_p * len(args))(*args)
        self._h = lib().mpmh_scene_load(len(args), argv, seed)
        if not self._h:
            raise MpmError(lib().mpmh_last_error(None).decode())

    def close(self):
        if self._h:
            lib().mpmh_scene_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc:
            raise MpmError(lib().mpmh_last_error(self._h).decode())

    @property
    def materials(self):
        out = np.zeros(lib().mpmh_n_materials(self._h), MATERIAL_DTYPE)
        lib().mpmh_get_materials(self._h, _ptr(out))
        return out

    @property
    def n_objects(self):
        return lib().mpmh_n_objects(self._h)

    @property
    def N(self):
        return lib().mpmh_grid_size(self._h)

    @property
    def full_count(self):
        return lib().mpmh_full_count(self._h)

    def object_counts(self):
        return [lib().mpmh_object_count(self._h, o) for o in range(self.n_objects)]

    def object_substituted(self):
        return [bool(lib().mpmh_object_substituted(self._h, o)) for o in range(self.n_objects)]

    def object_lifetimes(self):
        out = []
        for o in range(self.n_objects):
            b, e = ctypes.c_float(), ctypes.c_float()
            lib().mpmh_object_lifetime(self._h, o, ctypes.byref(b), ctypes.byref(e))
            out.append((b.value, e.value))
        return out

    def full_particles(self):
        out = np.empty(lib().mpmh_full_count(self._h), PARTICLE_DTYPE)
        lib().mpmh_get_full(self._h, _ptr(out))
        return out

    def active_particles(self):
        out = np.empty(lib().mpmh_active_count(self._h), PARTICLE_DTYPE)
        lib().mpmh_get_active(self._h, _ptr(out))
        return out

    @property
    def t(self):
        return lib().mpmh_time(self._h)

    # --- Simulation facade (GPU) ---
    def init_cuda(self):
        self._ck(lib().mpmh_init_cuda(self._h))

    def advance(self, n=1):
        self._ck(lib().mpmh_advance(self._h, n))

    def sync_device(self):
        self._ck(lib().mpmh_sync_device(self._h))

    # --- output ---
    def write_particles(self, path):
        self._ck(lib().mpmh_write_particles(self._h, str(path).encode()))

    def compute_mesh(self, path=None):
        nv, nf = ctypes.c_size_t(), ctypes.c_size_t()
        self._ck(lib().mpmh_compute_mesh(self._h, str(path).encode() if path else None, ctypes.byref(nv), ctypes.byref(nf)))
        return nv.value, nf.value


def winding_numbers(V, F, points):
    V = np.ascontiguousarray(V, np.float32)
    F = np.ascontiguousarray(F, np.int32)
    points = np.ascontiguousarray(points, np.float32)
    w = np.empty(len(points), np.float32)
    lib().mpmh_winding_numbers(_ptr(V), len(V), _ptr(F), len(F), _ptr(points), len(points), _ptr(w))
    return w


def load_mesh(path, size, position):
    pos = np.ascontiguousarray(position, np.float32)
    nv, nf, sub = ctypes.c_size_t(), ctypes.c_size_t(), ctypes.c_int()
    args = (str(path).encode(), float(size), _ptr(pos))
    if lib().mpmh_load_mesh(*args, None, None, ctypes.byref(nv), ctypes.byref(nf), ctypes.byref(sub)):
        raise MpmError(lib().mpmh_last_error(None).decode())
    V, F = np.empty((nv.value, 3), np.float32), np.empty((nf.value, 3), np.int32)
    lib().mpmh_load_mesh(*args, _ptr(V), _ptr(F), ctypes.byref(nv), ctypes.byref(nf), ctypes.byref(sub))
    return V, F, bool(sub.value)


def marching_tetrahedra(S):
    S = np.ascontiguousarray(S, np.float64)
    G = S.shape[0]
    assert S.shape == (G, G, G)
    nv, nf = ctypes.c_size_t(), ctypes.c_size_t()
    lib().mpmh_marching_tetrahedra(_ptr(S), G, None, 0, None, 0, ctypes.byref(nv), ctypes.byref(nf))
    V, F = np.empty((nv.value, 3), np.float64), np.empty((nf.value, 3), np.int32)
    lib().mpmh_marching_tetrahedra(_ptr(S), G, _ptr(V), nv.value, _ptr(F), nf.value, ctypes.byref(nv), ctypes.byref(nf))
    return V, F
