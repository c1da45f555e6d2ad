"""B200-native panorama ray march of fizyk20/atm-raytracer (the `gen` hot path).

`runtime` binds the CUDA C-ABI library (include/atmrt.h); `host` binds the C++ host helpers;
`config` mirrors the reference's YAML/CLI surface. Nothing here computes the hot path on the CPU.
"""
from . import abi  # noqa: F401

__all__ = ["abi", "config", "runtime", "host", "synth"]
