"""Build recipe: nvcc, sm_100a only, in-tree shared libraries (they travel to the GPU box with the snapshot).

  libscgpu.so       default: -fmad=true (FMA contraction, what the FP64 pipe is built for)
  libscgpu_strict.so  -fmad=false: same operation order AND same roundings as the reference / oracle
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "scgpu.cu")
DEPS = [SRC, os.path.join(HERE, "csrc", "pair_energy.cuh"), os.path.join(HERE, "csrc", "sweep.cuh"),
        os.path.join(os.path.dirname(HERE), "include", "scgpu.h")]
VARIANTS = {"fast": ("libscgpu.so", ["-fmad=true"]), "strict": ("libscgpu_strict.so", ["-fmad=false"])}


def lib_path(variant="fast"):
    return os.path.join(HERE, VARIANTS[variant][0])


def _stale(out):
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in DEPS)


def build(variant=None, force=False, verbose=False):
    """Compile the CUDA library (both variants by default). nvcc cross-compiles without a GPU."""
    names = [variant] if variant else list(VARIANTS)
    for v in names:
        out, flags = lib_path(v), VARIANTS[v][1]
        if not force and not _stale(out):
            continue
        cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
               "-diag-suppress", "177", "-shared", "-Xcompiler", "-fPIC"] + flags + ["-o", out, SRC]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    return [lib_path(v) for v in names]


if __name__ == "__main__":
    import sys
    build(force="--force" in sys.argv, verbose=True)
