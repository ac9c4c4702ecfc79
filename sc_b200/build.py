"""Build recipe: nvcc, sm_100a only, in-tree shared libraries (they travel to the GPU box with the snapshot).

  libscgpu.so       default: -fmad=true (FMA contraction, what the FP64 pipe is built for) and -DSCG_FAST_DIV (division
                    by Newton iteration without the IEEE special-operand tail, <= 1 ulp; see fdiv() in pair_energy.cuh)
  libscgpu_strict.so  -fmad=false: same operation order AND same roundings as the reference / oracle
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "scgpu.cu")
DEPS = [SRC, os.path.join(HERE, "csrc", "pair_energy.cuh"), os.path.join(HERE, "csrc", "sweep.cuh"), os.path.join(HERE, "csrc", "comm.cuh"), os.path.join(HERE, "csrc", "wall.cuh"), os.path.join(HERE, "csrc", "sweep_rounds.cuh"), os.path.join(HERE, "csrc", "sweep_phased.cuh"), os.path.join(HERE, "csrc", "wl_order.cuh"),
        os.path.join(os.path.dirname(HERE), "include", "scgpu.h")]
VARIANTS = {"fast": ("libscgpu.so", ["-fmad=true", "-DSCG_FAST_DIV"]), "strict": ("libscgpu_strict.so", ["-fmad=false"])}


def lib_path(variant="fast"):
    # SCGPU_LIB_FAST / SCGPU_LIB_STRICT: measure an experimental build (scripts/sweep_variants.py) through the same bindings
    return os.environ.get("SCGPU_LIB_" + variant.upper()) or os.path.join(HERE, VARIANTS[variant][0])


def _stale(out):
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in DEPS)


def build(variant=None, force=False, verbose=False):
    """Compile the CUDA library (both variants by default). nvcc cross-compiles without a GPU."""
    names = [variant] if variant else list(VARIANTS)
    procs = []
    for v in names:
        out, flags = lib_path(v), VARIANTS[v][1]
        if not force and not _stale(out):
            continue
        cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
               "-diag-suppress", "177", "-shared", "-Xcompiler", "-fPIC"] + flags + ["-o", out, SRC, "-ldl"]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd))
        procs.append((cmd, subprocess.Popen(cmd)))      # the variants compile side by side
    for cmd, p in procs:
        if p.wait() != 0:
            raise subprocess.CalledProcessError(p.returncode, cmd)
    return [lib_path(v) for v in names]


HOST_SRCS = [os.path.join(HERE, "csrc", "host", f) for f in ("topology.cpp", "calculator.cpp", "schost_capi.cpp", "mc_driver.cpp")]
HOST_DEPS = HOST_SRCS + [os.path.join(HERE, "csrc", "host", f) for f in ("topology.hpp", "calculator.hpp")]
HOST_LIB = os.path.join(HERE, "libschost.so")


def build_host(force=False):
    """C++ host mirror of the reference's formats and calculator interface (sc_b200/csrc/host), linked to libscgpu.so."""
    if not force and os.path.exists(HOST_LIB):
        t = os.path.getmtime(HOST_LIB)
        if not any(os.path.getmtime(d) > t for d in HOST_DEPS + [lib_path("fast")]):
            return HOST_LIB
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-ffp-contract=off", "-o", HOST_LIB] + HOST_SRCS + \
          ["-L" + HERE, "-lscgpu", "-Wl,-rpath,$ORIGIN"]
    subprocess.check_call(cmd)
    return HOST_LIB


if __name__ == "__main__":
    import sys
    build(force="--force" in sys.argv, verbose=True)
    build_host(force="--force" in sys.argv)
