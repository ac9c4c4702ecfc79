"""sc_b200 -- B200-native energy engine for patchy-spherocylinder Monte Carlo (the scOOP hot path).

The product is the CUDA library behind the C ABI in include/scgpu.h (sc_b200/csrc/). This package holds the
build recipe, a thin ctypes binding used by tests/bench, the synthetic-system generator of the benchmark
configurations, and the replica-exchange plumbing over torch.distributed. There is no CPU fallback.
"""
from .build import build, lib_path  # noqa: F401
from .engine import Engine, ScgpuError  # noqa: F401
