"""locomouse_cpp_b200 — B200-native (sm_100a) LocoMouse per-frame detection path.

Only what the hot path needs lives here: `csrc/` (CUDA kernels + the C ABI of
include/locomouse_b200.h), `host/` (C++ mirror of the reference's LocoMouse classes), the ctypes
binding (`api`), synthetic inputs (`synth`) and frame-range sharding (`sharding`).
There is no CPU fallback: `api.Detector` raises if the CUDA library cannot be loaded.
"""
from .types import (BOTTOM, PAW, SIDE, SNOUT, TAIL, Config, Model, Results, diff_results)  # noqa: F401

__all__ = ["Config", "Model", "Results", "diff_results", "PAW", "SNOUT", "TAIL", "BOTTOM", "SIDE"]
