"""Builds mpm_b200/libmpm_b200.so (CUDA kernels + C ABI) in-tree for sm_100a with nvcc."""
import concurrent.futures
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libmpm_b200.so")
OBJ = os.path.join(HERE, "build")
# one translation unit per shipped material model (they compile in parallel) + the handle / C ABI
SOURCES = ["mpm_sim.cu", "models_snow.cu", "models_fixed_corotated.cu", "models_jelly.cu"]


def _newest_source():
    t = 0.0
    for root in (CSRC, os.path.join(ROOT, "include"), os.path.join(ROOT, "include", "mpm_b200")):
        for f in os.listdir(root):
            t = max(t, os.path.getmtime(os.path.join(root, f)))
    return t


def nvcc_flags(verbose=False, defines=()):
    return ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-ccbin", "/usr/bin/g++",
            "-I" + os.path.join(ROOT, "include"), "-Xptxas", "-v" if verbose else "-warn-spills"] + ["-D" + d for d in defines]


def build(force=False, verbose=False, defines=(), out=None, extra_sources=()):
    """defines/out: experiment builds (tools/ab.py), e.g. defines=("MPM_P2G_MINBLK=3",).
    extra_sources: more .cu files linked into the library, e.g. a user-defined material
    (include/mpm_b200/plugin.cuh)."""
    lib = out or LIB
    if not force and os.path.exists(lib) and os.path.getmtime(lib) >= _newest_source():
        return lib
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    tag = hashlib.sha1(repr((sorted(defines), os.path.basename(lib))).encode()).hexdigest()[:10]
    objdir = os.path.join(OBJ, tag)
    os.makedirs(objdir, exist_ok=True)
    srcs = [os.path.join(CSRC, s) for s in SOURCES] + list(extra_sources)
    flags = nvcc_flags(verbose, defines)

    def compile_one(src):
        obj = os.path.join(objdir, os.path.basename(src) + ".o")
        r = subprocess.run([nvcc] + flags + ["-c", src, "-o", obj], capture_output=True, text=True)
        return src, obj, r

    objs = []
    with concurrent.futures.ThreadPoolExecutor(max_workers=len(srcs)) as ex:
        for src, obj, r in ex.map(compile_one, srcs):
            if verbose or r.returncode:
                sys.stderr.write(r.stdout + r.stderr)
            if r.returncode:
                raise RuntimeError("nvcc failed on " + src)
            objs.append(obj)
    r = subprocess.run([nvcc, "-shared", "-o", lib] + objs + ["-lnccl"], capture_output=True, text=True)
    if r.returncode:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed linking " + lib)
    return lib


PLUGIN_LIB = os.path.join(HERE, "libmpm_b200_plugin_example.so")
PLUGIN_SRC = os.path.join(ROOT, "tests", "plugin", "user_material.cu")


def build_plugin_example(force=False):
    """The library again with a user-defined material compiled in (tests/plugin/user_material.cu,
    registered through include/mpm_b200/plugin.cuh): what a user of the plugin surface builds."""
    if not force and os.path.exists(PLUGIN_LIB) and os.path.getmtime(PLUGIN_LIB) >= max(_newest_source(), os.path.getmtime(PLUGIN_SRC)):
        return PLUGIN_LIB
    return build(force=True, out=PLUGIN_LIB, extra_sources=(PLUGIN_SRC,))


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
