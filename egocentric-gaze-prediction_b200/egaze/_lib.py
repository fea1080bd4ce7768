"""ctypes binding of libegaze.so (C-ABI declared in include/egaze.h).

The prototypes are parsed from the header itself, so the header is the single source of truth and
`tests/test_abi.py` can check that every declared symbol is exported.  There is NO fallback: if the shared
library is missing or the device is not sm_100, calls raise.
"""
import ctypes
import os
import re

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(os.path.dirname(_HERE))
LIB_PATH = os.path.join(_HERE, "libegaze.so")
HEADER_PATH = os.path.join(_ROOT, "include", "egaze.h")

_SCALARS = {"int": ctypes.c_int, "long long": ctypes.c_longlong, "float": ctypes.c_float, "double": ctypes.c_double}


def parse_header(path=HEADER_PATH):
    """-> {name: [(ctype, argname), ...]} for every `int egaze_*(...)` declaration."""
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    protos = {}
    for m in re.finditer(r"\bint\s+(egaze_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        name, args = m.group(1), " ".join(m.group(2).split())
        sig = []
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                if "*" in a:
                    sig.append((ctypes.c_void_p, a.split("*")[-1].strip()))
                else:
                    typ, argname = a.rsplit(" ", 1)
                    sig.append((_SCALARS[typ.replace("const ", "").strip()], argname))
        protos[name] = sig
    return protos


_lib = None
_protos = None


def lib():
    global _lib, _protos
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "libegaze.so not found at %s -- build it with `python __graft_entry__.py build` "
                "(egaze has no CPU / PyTorch fallback)" % LIB_PATH)
        _protos = parse_header()
        handle = ctypes.CDLL(LIB_PATH)
        for name, sig in _protos.items():
            fn = getattr(handle, name)  # AttributeError here == header/library mismatch
            fn.restype = ctypes.c_int
            fn.argtypes = [t for t, _ in sig]
        _lib = handle
    return _lib


def prototypes():
    lib()
    return _protos


def last_error():
    buf = ctypes.create_string_buffer(512)
    lib().egaze_last_error(buf, 512)
    return buf.value.decode(errors="replace")


def _arg(a):
    if a is None:
        return None
    if isinstance(a, torch.Tensor):
        if not a.is_cuda:
            raise RuntimeError("egaze: expected a CUDA tensor (no CPU fallback), got device %s" % a.device)
        if not a.is_contiguous():
            raise RuntimeError("egaze: non-contiguous tensor passed to the C-ABI")
        return a.data_ptr()
    return a


# kernels launched per C-ABI call (bench.py's `gpu_launches` claim); entries not listed launch nothing
_LAUNCHES = {"egaze_floss_fwd": 2, "egaze_conv3x3_tiles": 0, "egaze_check_device": 0, "egaze_sm_count": 0,
             "egaze_bn_bwd_blocks": 0, "egaze_conv3x3_stats_shape": 0, "egaze_bn_bwd_reduce": 2,
             "egaze_conv3x3_set_prof": 0, "egaze_lf_scratch": 0, "egaze_lf_fwd": 7, "egaze_lf_bwd": 12}
_launch_count = 0


def launch_counter():
    return _launch_count


def call(name, *args):
    global _launch_count
    fn = getattr(lib(), name)
    rc = fn(*[_arg(a) for a in args])
    if name == "egaze_lstm_seq_fwd":
        _launch_count += int(args[9]) + 2  # T+1 wavefront launches (layer 0 at t | layer 1 at t-1) + one Linear over all steps
    elif name == "egaze_lstm_seq_bwd":
        _launch_count += 6 * int(args[6]) + 16  # per step: 2 x (gate grad + 2 small GEMMs); + Linear / weight-grad passes
    else:
        _launch_count += _LAUNCHES.get(name, 1)
    if rc != 0:
        raise RuntimeError("egaze: %s failed (rc=%d): %s" % (name, rc, last_error()))


def stream_ptr():
    """cudaStream_t of PyTorch's current stream on the current device, as an integer for the C-ABI.  (The public
    torch.cuda.current_stream() costs ~10 us per call -- several ms of host time per training step at ~300 launches.)"""
    try:
        return torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice())
    except AttributeError:  # private API moved: fall back to the public one
        return torch.cuda.current_stream().cuda_stream


_checked = set()


def check_device(device=None):
    """Fail loudly unless the current device is a B200-class (sm_100) GPU."""
    if not torch.cuda.is_available():
        raise RuntimeError("egaze: CUDA device required (hand-written sm_100a kernels, no CPU fallback)")
    idx = torch.cuda.current_device() if device is None else torch.device(device).index
    if idx in _checked:
        return
    call("egaze_check_device")
    _checked.add(idx)
