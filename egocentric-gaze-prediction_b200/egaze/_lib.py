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
# EGAZE_LIB: another build of the same ABI (A/B measurements of kernel variants inside one GPU session, tools/gpu_ab_lib.sh)
LIB_PATH = os.environ.get("EGAZE_LIB") or os.path.join(_HERE, "libegaze.so")
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


class _CurrentStream(object):
    """Placeholder for "the current stream of the device the call's tensors live on": resolved inside call()."""
    __slots__ = ()


STREAM = _CurrentStream()


def _raw_stream(dev_idx):
    """cudaStream_t of PyTorch's current stream on device dev_idx, as an integer.  (The public torch.cuda.current_stream()
    costs ~10 us per call -- several ms of host time per training step at ~300 launches.)"""
    try:
        return torch._C._cuda_getCurrentRawStream(dev_idx)
    except AttributeError:  # private API moved: fall back to the public one
        return torch.cuda.current_stream(dev_idx).cuda_stream


def _cur_device():
    try:
        return torch._C._cuda_getDevice()
    except AttributeError:
        return torch.cuda.current_device()


# kernels launched per C-ABI call (bench.py's `gpu_launches` claim); entries not listed launch nothing
_LAUNCHES = {"egaze_floss_fwd": 2, "egaze_conv3x3_tiles": 0, "egaze_check_device": 0, "egaze_sm_count": 0,
             "egaze_bn_bwd_blocks": 0, "egaze_conv3x3_stats_shape": 0, "egaze_bn_bwd_reduce": 2,
             "egaze_conv3x3_set_prof": 0, "egaze_lf_scratch": 0, "egaze_lf_fwd": 7, "egaze_lf_bwd": 12,
             "egaze_adam_job_bytes": 0, "egaze_jpeg_info": 0, "egaze_conv3x3_plan_create": 0, "egaze_plan_destroy": 0, "egaze_adam_multi": 2, "egaze_f16_weight_scale": 0}
_launch_count = 0


def launch_counter():
    return _launch_count


def call(name, *args):
    """Invoke one C-ABI entry point.  The library launches on the CUDA runtime's CURRENT device, so the call runs under a
    device guard for the device its tensors live on (the reference puts models on `cuda:N` without torch.cuda.set_device,
    gaze_full.py:37-38), and the stream argument is that device's current PyTorch stream.  Tensors on different devices, CPU
    tensors and non-contiguous tensors are errors (no fallback)."""
    global _launch_count
    fn = getattr(lib(), name)
    dev = None
    cargs = []
    for a in args:
        if isinstance(a, torch.Tensor):
            if not a.is_cuda:
                raise RuntimeError("egaze: expected a CUDA tensor (no CPU fallback), got device %s" % a.device)
            if not a.is_contiguous():
                raise RuntimeError("egaze: non-contiguous tensor passed to the C-ABI")
            if dev is None:
                dev = a.device.index
            elif a.device.index != dev:
                raise RuntimeError("egaze: %s got tensors on cuda:%d and cuda:%d" % (name, dev, a.device.index))
            cargs.append(a.data_ptr())
        else:
            cargs.append(a)
    if dev is None and not any(a is STREAM for a in cargs):
        rc = fn(*cargs)          # host-only entry point (sizes, versions, ...): no device involved
    else:
        cur = _cur_device()
        if dev is None:
            dev = cur
        for i, a in enumerate(cargs):
            if a is STREAM:
                cargs[i] = _raw_stream(dev)
        if dev != cur:
            with torch.cuda.device(dev):
                rc = fn(*cargs)
        else:
            rc = fn(*cargs)
    if name == "egaze_lstm_seq_fwd":
        _launch_count += int(args[9]) + 2  # T+1 wavefront launches (layer 0 at t | layer 1 at t-1) + one Linear over all steps
    elif name == "egaze_lstm_seq_bwd":
        _launch_count += 6 * int(args[6]) + 16  # per step: 2 x (gate grad + 2 small GEMMs); + Linear / weight-grad passes
    else:
        _launch_count += _LAUNCHES.get(name, 1)
    if rc != 0:
        raise RuntimeError("egaze: %s failed (rc=%d): %s" % (name, rc, last_error()))


def call_on(dev, name, *args):
    """call() for entry points whose arguments carry no tensor (plans, handles): runs on device index `dev`."""
    global _launch_count
    fn = getattr(lib(), name)
    cargs = [(_raw_stream(dev) if a is STREAM else a) for a in args]
    if dev != _cur_device():
        with torch.cuda.device(dev):
            rc = fn(*cargs)
    else:
        rc = fn(*cargs)
    _launch_count += _LAUNCHES.get(name, 1)
    if rc != 0:
        raise RuntimeError("egaze: %s failed (rc=%d): %s" % (name, rc, last_error()))


def stream_ptr():
    """The stream argument of a C-ABI call: resolved by call() to the current PyTorch stream of the tensors' device."""
    return STREAM


def stream_key(device=None):
    """Integer identity of the current stream on `device` (for per-stream scratch buffers)."""
    idx = _cur_device() if device is None else torch.device(device).index
    return _raw_stream(_cur_device() if idx is None else idx)


_checked = set()


def check_device(device=None):
    """Fail loudly unless `device` (default: the current device) is a B200-class (sm_100) GPU."""
    if not torch.cuda.is_available():
        raise RuntimeError("egaze: CUDA device required (hand-written sm_100a kernels, no CPU fallback)")
    if device is None:
        idx = _cur_device()
    else:
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("egaze: expected a CUDA tensor (no CPU fallback), got device %s" % device)
        idx = _cur_device() if device.index is None else device.index
    if idx in _checked:
        return
    with torch.cuda.device(idx):
        call("egaze_check_device")
    _checked.add(idx)
