"""ctypes bindings for the parity checker under oracle/ (test infrastructure only).

May be imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs; never by the product package mpm_b200/.
"""
import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")

# MLS_APIC_Particle, 104 bytes (reference include/types.h:24-35 + include/TransferScheme.h:46-54)
PARTICLE_DTYPE = np.dtype(
    [("material_type", "u1"), ("pad", "u1", 3), ("x", "f4", 3), ("v", "f4", 3), ("F", "f4", 9), ("C", "f4", 9),
     ("Jp", "f4")]
)
assert PARTICLE_DTYPE.itemsize == 104

SNOW, FIXED_COROTATED, JELLY = 0, 1, 2

_fp = ctypes.POINTER(ctypes.c_float)
_u32p = ctypes.POINTER(ctypes.c_uint32)
_ip = ctypes.POINTER(ctypes.c_int)


def _f(a):
    return a.ctypes.data_as(_fp)


def build(force=False):
    so = os.path.join(ORACLE_DIR, "liboracle.so")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(os.path.join(ORACLE_DIR, "mpm_oracle.cpp")):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "--no-print-directory"], stdout=subprocess.DEVNULL)
    return so


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.oracle_determinant.restype = ctypes.c_float
        _lib.oracle_max_threads.restype = ctypes.c_int
    return _lib


def ref_lib(name):
    """oracle/_ref/<name>.so (the reference's own code built for the host) or None."""
    p = os.path.join(ORACLE_DIR, "_ref", name + ".so")
    return ctypes.CDLL(p) if os.path.exists(p) else None


def new_particles(x, v=None, material=0):
    x = np.asarray(x, np.float32)
    p = np.zeros(x.shape[0], PARTICLE_DTYPE)
    p["x"] = x
    if v is not None:
        p["v"] = np.asarray(v, np.float32)
    p["F"][:, [0, 4, 8]] = 1.0
    p["Jp"] = 1.0
    p["material_type"] = material
    return p


def make_material(volume, density=700.0, E=1.4e5, Nu=0.2, hardening=10.0, lo=0.975, hi=1.0075):
    out = np.zeros(7, np.float32)
    lib().oracle_make_material(*[ctypes.c_double(a) for a in (volume, density, E, Nu, hardening, lo, hi)], _f(out))
    return out


def params(dt, N):
    dx, dxi = ctypes.c_float(), ctypes.c_float()
    lib().oracle_params(ctypes.c_float(dt), ctypes.c_uint32(N), ctypes.byref(dx), ctypes.byref(dxi))
    return dx.value, dxi.value


def set_threads(n):
    lib().oracle_set_threads(int(n))


def max_threads():
    return lib().oracle_max_threads()


def svd3(A, which="oracle"):
    A = np.ascontiguousarray(A, np.float32).reshape(-1, 9)
    n = A.shape[0]
    U, S, V = np.empty((n, 9), np.float32), np.empty((n, 3), np.float32), np.empty((n, 9), np.float32)
    if which == "oracle":
        lib().oracle_svd3_batch(_f(A), _f(U), _f(S), _f(V), ctypes.c_size_t(n))
    else:
        r = ref_lib("libref_svd3")
        if r is None:
            raise FileNotFoundError("oracle/_ref/libref_svd3.so")
        r.ref_svd3_batch(_f(A), _f(U), _f(S), _f(V), ctypes.c_size_t(n))
    return U.reshape(n, 3, 3), S, V.reshape(n, 3, 3)


def polar(A):
    A = np.ascontiguousarray(A, np.float32).reshape(-1, 9)
    n = A.shape[0]
    R, S = np.empty((n, 9), np.float32), np.empty((n, 9), np.float32)
    lib().oracle_polar_batch(_f(A), _f(R), _f(S), ctypes.c_size_t(n))
    return R.reshape(n, 3, 3), S.reshape(n, 3, 3)


def determinant(A):
    A = np.ascontiguousarray(A, np.float32).reshape(9)
    return lib().oracle_determinant(_f(A))


def weights(x, dx_inv):
    x = np.ascontiguousarray(x, np.float32)
    base = np.zeros(3, np.int32)
    w = np.zeros(9, np.float32)
    lib().oracle_weights(_f(x), ctypes.c_float(dx_inv), base.ctypes.data_as(_ip), _f(w))
    return base, w.reshape(3, 3)


def cell_keys(p, dt, N):
    keys = np.empty(p.shape[0], np.uint32)
    lib().oracle_cell_keys(p.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(p.shape[0]), ctypes.c_float(dt),
                           ctypes.c_uint32(N), keys.ctypes.data_as(_u32p))
    return keys


def sort_perm(keys):
    keys = np.ascontiguousarray(keys, np.uint32)
    perm = np.empty(keys.shape[0], np.uint32)
    lib().oracle_sort_perm(keys.ctypes.data_as(_u32p), ctypes.c_size_t(keys.shape[0]), perm.ctypes.data_as(_u32p))
    return perm


def new_grid(N):
    return np.zeros((N, N, N, 4), np.float32)


def p2g(p, mats, dt, N, kind, grid=None):
    if grid is None:
        grid = new_grid(N)
    mats = np.ascontiguousarray(mats, np.float32)
    lib().oracle_p2g(p.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(p.shape[0]), _f(mats), ctypes.c_float(dt),
                     ctypes.c_uint32(N), ctypes.c_int(kind), _f(grid))
    return grid


def p2g_magnitudes(p, mats, dt, N, kind):
    """Sum over particles of |contribution| per node and channel (the scale of a per-node P2G error)."""
    grid = new_grid(N)
    mats = np.ascontiguousarray(mats, np.float32)
    lib().oracle_p2g_magnitudes(p.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(p.shape[0]), _f(mats), ctypes.c_float(dt),
                                ctypes.c_uint32(N), ctypes.c_int(kind), _f(grid))
    return grid


def grid_update(grid, dt, N):
    lib().oracle_grid_update(_f(grid), ctypes.c_float(dt), ctypes.c_uint32(N))
    return grid


def g2p(grid, p, mats, dt, N, kind):
    mats = np.ascontiguousarray(mats, np.float32)
    lib().oracle_g2p(_f(grid), p.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(p.shape[0]), _f(mats),
                     ctypes.c_float(dt), ctypes.c_uint32(N), ctypes.c_int(kind))
    return p


def advance(p, mats, dt, N, kind, n_steps=1, grid=None):
    if grid is None:
        grid = new_grid(N)
    mats = np.ascontiguousarray(mats, np.float32)
    lib().oracle_advance(p.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(p.shape[0]), _f(mats), ctypes.c_float(dt),
                         ctypes.c_uint32(N), ctypes.c_int(kind), _f(grid), ctypes.c_int(n_steps))
    return p, grid


class Ref:
    """The reference's own kernels + plugin headers compiled for the host (oracle/_ref)."""

    def __init__(self, kind):
        self.kind = kind
        tag = {SNOW: "snow", FIXED_COROTATED: "fc", JELLY: "jelly"}[kind]
        self.pre = f"ref_{tag}_"
        self.lib = ref_lib("libref_mpm_" + tag)
        self.available = self.lib is not None

    def fn(self, name):
        return getattr(self.lib, self.pre + name)

    def make_material(self, volume, density=700.0, E=1.4e5, Nu=0.2, hardening=10.0, lo=0.975, hi=1.0075):
        out = np.zeros(7, np.float32)
        self.fn("make_material")(*[ctypes.c_double(a) for a in (volume, density, E, Nu, hardening, lo, hi)], _f(out))
        return out

    def params(self, dt, N):
        dx, dxi = ctypes.c_float(), ctypes.c_float()
        self.fn("params")(ctypes.c_float(dt), ctypes.c_uint32(N), ctypes.byref(dx), ctypes.byref(dxi))
        return dx.value, dxi.value

    def p2g(self, p, mats, dt, N, grid=None):
        if grid is None:
            grid = new_grid(N)
        mats = np.ascontiguousarray(mats, np.float32).reshape(-1, 7)
        self.fn("p2g")(p.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(p.shape[0]), _f(mats),
                       ctypes.c_int(mats.shape[0]), ctypes.c_float(dt), ctypes.c_uint32(N), _f(grid))
        return grid

    def grid_update(self, grid, dt, N):
        self.fn("grid_update")(_f(grid), ctypes.c_float(dt), ctypes.c_uint32(N))
        return grid

    def g2p(self, grid, p, mats, dt, N):
        mats = np.ascontiguousarray(mats, np.float32).reshape(-1, 7)
        rc = self.fn("g2p")(_f(grid), p.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(p.shape[0]), _f(mats),
                            ctypes.c_int(mats.shape[0]), ctypes.c_float(dt), ctypes.c_uint32(N))
        assert rc == 0
        return p

    def advance(self, p, mats, dt, N, n_steps=1, grid=None):
        if grid is None:
            grid = new_grid(N)
        mats = np.ascontiguousarray(mats, np.float32).reshape(-1, 7)
        rc = self.fn("advance")(p.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(p.shape[0]), _f(mats),
                                ctypes.c_int(mats.shape[0]), ctypes.c_float(dt), ctypes.c_uint32(N), _f(grid),
                                ctypes.c_int(n_steps))
        assert rc == 0
        return p, grid

    def polar(self, A):
        A = np.ascontiguousarray(A, np.float32).reshape(9)
        R, S = np.empty(9, np.float32), np.empty(9, np.float32)
        self.fn("polar")(_f(A), _f(R), _f(S))
        return R.reshape(3, 3), S.reshape(3, 3)


# ------------------------------------------------------------------------------------------------
# scene generators shared by oracle tests, GPU parity tests and bench.py
# ------------------------------------------------------------------------------------------------
def lowbias32(x):
    """Counter-based 32-bit hash (same constants as mpm_b200/csrc generate kernel)."""
    x = np.asarray(x, np.uint32).copy()
    x ^= x >> np.uint32(16)
    x *= np.uint32(0x7FEB352D)
    x ^= x >> np.uint32(15)
    x *= np.uint32(0x846CA68B)
    x ^= x >> np.uint32(16)
    return x


def dense_block_positions(count, seed=1234, lo=0.1, hi=0.9, first_id=0):
    """SURVEY.md §8(d) config 4/5 generator: x = lo + (hi-lo) * u(hash(seed, id, axis))."""
    ids = np.arange(first_id, first_id + count, dtype=np.uint64)
    out = np.empty((count, 3), np.float32)
    for axis in range(3):
        with np.errstate(over="ignore"):
            h = lowbias32(((ids * np.uint64(3) + np.uint64(axis)) & np.uint64(0xFFFFFFFF)).astype(np.uint32)
                          ^ lowbias32(np.uint32(seed)))
        u = (h >> np.uint32(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)
        out[:, axis] = np.float32(lo) + (np.float32(hi) - np.float32(lo)) * u
    return out


def dense_block_stress(p, seed=1234, shear=0.0, f_noise=0.0, first_id=0):
    """mpm_generate_dense_block_stressed (mpm_b200/csrc/handle_kernels.cuh) on the host, bit for bit:
    v = shear * (y - 0.5, 0, 0.3 (x - 0.5)), F = I + f_noise * u, u in [-1, 1) from the counter hash."""
    n = p.shape[0]
    f32 = np.float32
    if shear:
        p["v"][:, 0] = f32(shear) * (p["x"][:, 1] - f32(0.5))
        p["v"][:, 2] = (f32(0.3) * f32(shear)) * (p["x"][:, 0] - f32(0.5))
    if f_noise:
        ids = np.arange(first_id, first_id + n, dtype=np.uint64)
        with np.errstate(over="ignore"):
            salt = lowbias32(np.uint32(seed)) * np.uint32(0x9E3779B9) + np.uint32(77)
            for e in range(9):
                h = lowbias32(((ids * np.uint64(9) + np.uint64(e)) & np.uint64(0xFFFFFFFF)).astype(np.uint32) ^ salt)
                u = (h >> np.uint32(8)).astype(f32) * f32(2.0 / 16777216.0) - f32(1.0)
                r, c = divmod(e, 3)  # the device numbers the entries row-major; the AoS record is column-major
                p["F"][:, 3 * c + r] = p["F"][:, 3 * c + r] + f32(f_noise) * u
    return p


def sphere_positions(count_density, size, position, rng):
    """Stand-in for sphere.obj (LFS stub in the reference checkout): rejection-sample the ball of
    diameter `size` whose bounding box has its low corner at `position` (src/mpm.cu:331-394)."""
    n_target = int(count_density * size ** 3)
    pts = rng.random((n_target, 3), dtype=np.float32) * np.float32(size)
    c = np.float32(size / 2)
    keep = ((pts - c) ** 2).sum(1) < c * c
    return (pts[keep] + np.asarray(position, np.float32)).astype(np.float32)
