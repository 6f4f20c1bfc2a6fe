"""Builds libvoxelrt.so (hand-written sm_100a CUDA behind the C-ABI of include/voxelrt.h).

In-tree build with explicit nvcc commands; the resulting .so is git-ignored but travels to the
GPU box with the gpurun snapshot. There is no alternative backend: if nvcc is missing the build
fails loudly.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libvoxelrt.so")
SOURCES = ["vrt_api.cu", "vrt_render.cu", "vrt_restir.cu", "vrt_temporal.cu", "vrt_build.cu", "vrt_sky_precompute.cu"]
HEADERS = ["vrt_kshared.cuh", "vrt_common.cuh", "vrt_trace.cuh", "vrt_bsdf.cuh", "vrt_sky.cuh", "vrt_restir.cuh", "vrt_internal.h",
           os.path.join("..", "..", "include", "voxelrt.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
]


def _nvcc():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: libvoxelrt cannot be built (there is no CPU fallback)")
    return nvcc


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build_variant(name, defines):
    """Tuning variant of the library (e.g. a different __launch_bounds__), built next to the
    default one as variants/libvoxelrt_<name>.so and selected with VRT_LIB=<path>."""
    nvcc = _nvcc()
    ccbin = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else None
    vdir = os.path.join(HERE, "variants")
    os.makedirs(vdir, exist_ok=True)
    objs = []
    for src in SOURCES:
        obj = os.path.join(vdir, "%s_%s" % (name, src.replace(".cu", ".o")))
        cmd = [nvcc] + NVCC_FLAGS + ["-D" + d for d in defines] + (["-ccbin", ccbin] if ccbin else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("nvcc failed on %s" % src)
        objs.append(obj)
    lib = os.path.join(vdir, "libvoxelrt_%s.so" % name)
    r = subprocess.run([nvcc, "-shared", "-o", lib] + (["-ccbin", ccbin] if ccbin else []) + objs + ["-lcudart"], capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("link failed")
    for o in objs:
        os.remove(o)
    return lib


def build(force=False, verbose=False):
    """Compile every CUDA translation unit for sm_100a and link libvoxelrt.so."""
    if not force and not _stale():
        return LIB
    nvcc = _nvcc()
    env = dict(os.environ)
    # the image's CXX points at a gcc wrapper without OpenMP specs; nvcc wants the system one
    ccbin = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else None
    objs = []
    logs = []
    for src in SOURCES:
        obj = os.path.join(CSRC, src.replace(".cu", ".o"))
        cmd = [nvcc] + NVCC_FLAGS + (["-ccbin", ccbin] if ccbin else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True, env=env)
        logs.append(r.stderr)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("nvcc failed on %s" % src)
        objs.append(obj)
    cmd = [nvcc, "-shared", "-o", LIB] + (["-ccbin", ccbin] if ccbin else []) + objs + ["-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True, env=env)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("link failed")
    with open(os.path.join(CSRC, "ptxas.log"), "w") as f:
        f.write("\n".join(logs))
    if verbose:
        print("\n".join(logs))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
