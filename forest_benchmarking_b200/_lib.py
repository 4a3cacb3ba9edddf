"""ctypes binding of libqtomo.so (the sm_100a CUDA library under csrc/).

There is no CPU fallback: if the library cannot be built/loaded, or no CUDA device is present when a
compute entry point is called, the call raises.  PyTorch is used only to own device memory and streams.
"""
import ctypes
import glob
import os
import re
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_HERE, "csrc")
_SO = os.path.join(_HERE, "libqtomo.so")
_HEADER = os.path.join(os.path.dirname(_HERE), "include", "qtomo.h")
_lib = None

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC"]


class QtomoError(RuntimeError):
    pass


def _sources():
    return sorted(glob.glob(os.path.join(_CSRC, "*.cu")))


def _stale():
    if not os.path.exists(_SO):
        return True
    t = os.path.getmtime(_SO)
    deps = _sources() + glob.glob(os.path.join(_CSRC, "*.cuh")) + [_HEADER]
    return any(os.path.getmtime(p) > t for p in deps)


def build(force=False, verbose=False):
    """Compile csrc/*.cu into libqtomo.so for sm_100a (nvcc cross-compiles without a GPU)."""
    if not force and not _stale():
        return _SO
    nvcc = os.environ.get("NVCC", "nvcc")
    objs = []
    procs = []
    os.makedirs(os.path.join(_HERE, "build"), exist_ok=True)
    for src in _sources():
        obj = os.path.join(_HERE, "build", os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
               "-Xcompiler", "-fPIC", "-c", src, "-o", obj]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise QtomoError("nvcc failed: %s\n%s" % (" ".join(cmd), out.decode()))
        if verbose and out:
            print(out.decode())
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", _SO] + objs
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    if r.returncode != 0:
        raise QtomoError("nvcc link failed:\n" + r.stdout.decode())
    return _SO


def declared_symbols():
    """Every function name declared in include/qtomo.h."""
    with open(_HEADER) as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(?:int|int64_t)\s+(qt_\w+)\s*\(", text)))


def lib():
    """Load (building first if the .so is missing and nvcc is available) and return the CDLL."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_SO):
        build()
    try:
        _lib = ctypes.CDLL(_SO)
    except OSError as e:  # fail loudly: there is no fallback
        raise QtomoError(f"cannot load {_SO}: {e}") from e
    for name in declared_symbols():
        getattr(_lib, name).restype = ctypes.c_int64 if name.endswith("_bytes") else ctypes.c_int
    return _lib


def last_error():
    buf = ctypes.create_string_buffer(512)
    lib().qt_last_error(buf, 512)
    return buf.value.decode()


def check(rc, what=""):
    if rc != 0:
        raise QtomoError(f"{what} failed (code {rc}): {last_error()}")


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise QtomoError("forest_benchmarking_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return torch


def current_stream_ptr():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    """Device pointer of a torch tensor (must be contiguous) as c_void_p; None -> NULL."""
    if t is None:
        return ctypes.c_void_p(0)
    if not t.is_contiguous():
        raise QtomoError("non-contiguous tensor passed to libqtomo")
    return ctypes.c_void_p(t.data_ptr())
