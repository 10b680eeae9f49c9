"""bath_b200 -- B200 (sm_100a) engine for the translated-search hot path of BATH's bathsearch.

The product is libbathgpu.so (hand-written CUDA behind the C ABI in include/bathgpu.h);
this package holds its sources (csrc/), the build recipe and a ctypes binding of the C ABI.
There is no CPU fallback: loading fails loudly if the library has not been built, and every
call fails if no CUDA device is present.
"""
from .build import build_library, library_path  # noqa: F401
