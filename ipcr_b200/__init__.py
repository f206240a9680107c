"""ipcr_b200 -- B200-native (sm_100a) replacement for ipcr's core/engine scan.

Only the hot path lives here: the CUDA library (csrc/, C ABI in
include/ipcr_gpu.h) and the host-side mirror of the reference engine interface
(engine.py).  There is no CPU implementation of the scan in this package.
"""
from .engine import (CompiledPanel, Config, Engine, GpuError, New, Pair, Product,
                     SimulationScratch, rev_comp)

__all__ = ["CompiledPanel", "Config", "Engine", "GpuError", "New", "Pair", "Product",
           "SimulationScratch", "rev_comp"]
