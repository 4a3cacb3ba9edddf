"""ctypes binding of libqtomo.so (the sm_100a CUDA library under csrc/).

There is no CPU fallback: if the library cannot be built/loaded, or no CUDA device is present when a
compute entry point is called, the call raises.  PyTorch is used only to own device memory and streams.
"""
import contextlib
import ctypes
import glob
import os
import re
import shutil
import subprocess
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_HERE, "csrc")
_SO = os.path.join(_HERE, "libqtomo.so")
_STAMP = _SO + ".srchash"  # hash of the sources the .so was built from
_HEADER = os.path.join(os.path.dirname(_HERE), "include", "qtomo.h")
_lib = None
_lib_lock = threading.Lock()

NVCC_COMPILE = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
                "-Xcompiler", "-fPIC"]


class QtomoError(RuntimeError):
    pass


def _sources():
    return sorted(glob.glob(os.path.join(_CSRC, "*.cu")))


def _source_hash():
    """Content hash of everything the library is compiled from (mtimes do not survive every copy of the tree)."""
    import hashlib
    h = hashlib.sha256()
    for p in _sources() + sorted(glob.glob(os.path.join(_CSRC, "*.cuh"))) + [_HEADER]:
        h.update(os.path.basename(p).encode())
        with open(p, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def _stale():
    if not os.path.exists(_SO) or not os.path.exists(_STAMP):
        return True
    with open(_STAMP) as f:
        return f.read().strip() != _source_hash()


def _unit_hash(src):
    """Hash of one translation unit: its source plus every header it may include."""
    import hashlib
    h = hashlib.sha256(" ".join(NVCC_COMPILE).encode())
    for p in [src] + sorted(glob.glob(os.path.join(_CSRC, "*.cuh"))) + [_HEADER]:
        with open(p, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def build(force=False, verbose=False):
    """Compile csrc/*.cu into libqtomo.so for sm_100a (nvcc cross-compiles without a GPU).  Translation units
    whose source and headers are unchanged since their object was built are not recompiled."""
    if not force and not _stale():
        return _SO
    nvcc = os.environ.get("NVCC", "nvcc")
    objs = []
    procs = []
    os.makedirs(os.path.join(_HERE, "build"), exist_ok=True)
    for src in _sources():
        obj = os.path.join(_HERE, "build", os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        stamp, want = obj + ".srchash", _unit_hash(src)
        if not force and os.path.exists(obj) and os.path.exists(stamp):
            with open(stamp) as f:
                if f.read().strip() == want:
                    continue
        cmd = [nvcc] + NVCC_COMPILE + ["-c", src, "-o", obj]
        procs.append((cmd, stamp, want, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    failed = []
    for cmd, stamp, want, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed.append("nvcc failed: %s\n%s" % (" ".join(cmd), out.decode()))
            continue
        with open(stamp, "w") as f:
            f.write(want)
        if verbose and out:
            print(out.decode())
    if failed:
        raise QtomoError("\n".join(failed))
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", _SO] + objs
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    if r.returncode != 0:
        raise QtomoError("nvcc link failed:\n" + r.stdout.decode())
    with open(_STAMP, "w") as f:
        f.write(_source_hash())
    return _SO


def declared_symbols():
    """Every function name declared in include/qtomo.h."""
    with open(_HEADER) as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(?:int|int64_t)\s+(qt_\w+)\s*\(", text)))


def lib():
    """Load and return the CDLL.  The library is (re)built first when it is missing or older than any of
    csrc/*.cu, csrc/*.cuh or include/qtomo.h -- the .so is git-ignored, so after an edit or a pull a silent run
    on the old binary must not happen.  On a box without nvcc a stale library is an error, not a fallback."""
    global _lib
    if _lib is not None:
        return _lib
    with _lib_lock:  # plans may be created from several host threads
        if _lib is not None:
            return _lib
        if _stale():
            if shutil.which(os.environ.get("NVCC", "nvcc")) is None:
                if not os.path.exists(_SO):
                    raise QtomoError(f"{_SO} is missing and nvcc is not available to build it")
                raise QtomoError(f"{_SO} is older than its sources and nvcc is not available to rebuild it")
            build()
        try:
            handle = ctypes.CDLL(_SO)
        except OSError as e:  # fail loudly: there is no fallback
            raise QtomoError(f"cannot load {_SO}: {e}") from e
        for name in declared_symbols():
            try:
                fn = getattr(handle, name)
            except AttributeError as e:
                raise QtomoError(f"{_SO} does not export {name} declared in include/qtomo.h") from e
            fn.restype = ctypes.c_int64 if name.endswith("_bytes") else ctypes.c_int
        _lib = handle
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


def current_stream_ptr(device=None):
    """cudaStream_t of torch's current stream on `device` (default: the current device)."""
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def common_device(*tensors, plan=None):
    """The one CUDA device all `tensors` (None entries skipped) live on; raises if they disagree with each other
    or with the device `plan` was created on (plans cudaMalloc their tables on the device current at creation)."""
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise ValueError("expected CUDA tensors")
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise ValueError(f"tensors live on different devices ({dev} and {t.device})")
    pdev = getattr(plan, "device", None)
    if pdev is not None and dev is not None and pdev != dev:
        raise ValueError(f"plan was created on {pdev} but the tensors live on {dev}")
    return dev if dev is not None else pdev


@contextlib.contextmanager
def on_device(device):
    """Makes `device` current for the launch, so that the stream handed to the C ABI, the kernel launch and any
    output allocation all belong to the device that owns the operands."""
    import torch
    with torch.cuda.device(device):
        yield


def check_tensor(name, t, dtype, shape):
    """dtype / device / exact shape / contiguity of a tensor whose raw pointer goes to a kernel."""
    if t.dtype != dtype or not t.is_cuda or tuple(t.shape) != tuple(shape) or not t.is_contiguous():
        raise ValueError(f"{name} must be a contiguous CUDA {dtype} tensor of shape {list(shape)}, got "
                         f"{t.dtype} {list(t.shape)} on {t.device}{'' if t.is_contiguous() else ' (non-contiguous)'}")


def ptr(t):
    """Device pointer of a torch tensor (must be contiguous) as c_void_p; None -> NULL."""
    if t is None:
        return ctypes.c_void_p(0)
    if not t.is_contiguous():
        raise QtomoError("non-contiguous tensor passed to libqtomo")
    return ctypes.c_void_p(t.data_ptr())
