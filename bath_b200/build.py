"""Build libbathgpu.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_HERE, "csrc")
_SO = os.path.join(_HERE, "libbathgpu.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-ftz=true", "-shared", "-Xcompiler", "-fPIC"]


def library_path():
    """Path of the built library (BATHGPU_LIB overrides it, for A/B builds while tuning)."""
    return os.environ.get("BATHGPU_LIB", _SO)


def _sources():
    out = []
    for root, _, files in os.walk(_CSRC):
        out += [os.path.join(root, f) for f in files if f.endswith((".cu", ".cuh", ".h", ".cpp", ".hpp"))]
    out.append(os.path.join(_HERE, "..", "include", "bathgpu.h"))
    return out


_HOST = os.path.join(_HERE, "host")
_HOST_SO = os.path.join(_HERE, "libbathhost.so")
GXX_FLAGS = ["-O2", "-g", "-std=c++17", "-fPIC", "-shared", "-Wall", "-ffp-contract=off", "-fno-fast-math", "-pthread"]


def host_library_path():
    return _HOST_SO


def build_host_library(force=False):
    """libbathhost.so: the host side of the path (model set-up, length models, pipeline logic), g++ only."""
    srcs = [os.path.join(_HOST, f) for f in sorted(os.listdir(_HOST)) if f.endswith((".cpp", ".h"))]
    srcs.append(os.path.join(_HERE, "..", "include", "bathhost.h"))
    if not force and os.path.exists(_HOST_SO) and all(os.path.getmtime(s) <= os.path.getmtime(_HOST_SO) for s in srcs):
        return _HOST_SO
    cmd = [os.environ.get("CXX", "g++")] + GXX_FLAGS + ["-o", _HOST_SO] + [s for s in srcs if s.endswith(".cpp")]
    subprocess.check_call(cmd)
    return _HOST_SO


def build_library(force=False, verbose=False):
    srcs = _sources()
    if not force and os.path.exists(_SO) and all(os.path.getmtime(s) <= os.path.getmtime(_SO) for s in srcs):
        return _SO
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    units = sorted(s for s in srcs if s.endswith((".cu", ".cpp")))
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", _SO] + units
    subprocess.check_call(cmd)
    return _SO


if __name__ == "__main__":
    import sys
    print(build_host_library(force="--force" in sys.argv))
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
