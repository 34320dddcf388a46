"""Builds mpm_b200/libmpm_b200.so (CUDA kernels + C ABI) in-tree for sm_100a with nvcc."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libmpm_b200.so")
SOURCES = ["mpm_sim.cu"]


def _newest_source():
    t = 0.0
    for root in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        for f in os.listdir(root):
            t = max(t, os.path.getmtime(os.path.join(root, f)))
    return t


def build(force=False, verbose=False, defines=(), out=None):
    """defines/out: experiment builds (tools/ab.py), e.g. defines=("MPM_G2P_VARIANT=0",)."""
    global LIB
    lib_default = LIB
    if out:
        LIB = out
    try:
        return _build(force, verbose, defines)
    finally:
        LIB = lib_default


def _build(force, verbose, defines):
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= _newest_source():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-shared",
           "-Xcompiler", "-fPIC", "-ccbin", "/usr/bin/g++", "-Xptxas", "-v" if verbose else "-warn-spills",
           "-o", LIB] + ["-D" + d for d in defines] + [os.path.join(CSRC, s) for s in SOURCES] + ["-lnccl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode:
        raise RuntimeError("nvcc failed building libmpm_b200.so")
    return LIB


HOST = os.path.join(HERE, "host")
HOST_LIB = os.path.join(HERE, "libmpm_b200_host.so")
HOST_CLI = os.path.join(HERE, "mpm_b200_cli")


def build_host(force=False):
    """g++ -> mpm_b200/libmpm_b200_host.so (scene front end + Simulation facade behind C entry points,
    used by the tests) and mpm_b200/mpm_b200_cli (the reference's main loop).  Both link the CUDA
    library through $ORIGIN; -ffp-contract=off keeps the float arithmetic of the sampler as written."""
    build()
    newest = max(os.path.getmtime(os.path.join(HOST, f)) for f in os.listdir(HOST))
    newest = max(newest, os.path.getmtime(LIB))
    if not force and all(os.path.exists(t) and os.path.getmtime(t) >= newest for t in (HOST_LIB, HOST_CLI)):
        return HOST_LIB
    common = ["/usr/bin/g++", "-O2", "-std=c++17", "-fopenmp", "-ffp-contract=off", "-Wall", "-Wno-unknown-pragmas"]
    link = ["-L" + HERE, "-lmpm_b200", "-Wl,-rpath,$ORIGIN"]
    for cmd in (common + ["-fPIC", "-shared", "-o", HOST_LIB, os.path.join(HOST, "host_capi.cpp")] + link,
                common + ["-o", HOST_CLI, os.path.join(HOST, "main.cpp")] + link):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("g++ failed building the host front end")
    return HOST_LIB


if __name__ == "__main__":
    print(build(force=True, verbose="-v" in sys.argv))
    print(build_host(force=True))
