"""Input pipeline on the device.

The reference's dataset (data/STdatas.py:50-73) decodes, per sample and on the CPU, one RGB JPEG and TWENTY optical-flow JPEGs
(flow_x / flow_y of the current and the nine previous frames, data/STdatas.py:18-20) -- while a video is walked in order every
flow frame is therefore decoded ten times -- then converts, normalises and stacks them in PyTorch.  Here:

  decode_jpeg()   nvJPEG (GPU) decode of JPEG bytes to a device uint8 tensor in cv2.imread's layout (BGR / gray)
  normalize_image()  the reference's image normalisation (bit-identical arithmetic)
  FlowWindow      a device-side ring of the last 10 decoded flow frames of V videos advancing in lockstep: each new frame is
                  pushed once, the 20-channel stack is assembled from the ring (bit-identical to the reference's tensor)

Everything returns the NCHW fp32 tensors the drop-in modules' forward() takes (models.model_SP.forward(x_s, x_t)).
"""
import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import call, stream_ptr


def jpeg_info(data):
    """(width, height, components) of JPEG bytes."""
    buf = np.frombuffer(data, dtype=np.uint8)
    w, h, c = ctypes.c_int(0), ctypes.c_int(0), ctypes.c_int(0)
    call("egaze_jpeg_info", buf.ctypes.data, buf.size, ctypes.addressof(w), ctypes.addressof(h), ctypes.addressof(c))
    return w.value, h.value, c.value


def decode_jpeg(data, gray=False, device=None, out=None):
    """JPEG bytes -> device uint8 tensor [H, W, 3] (BGR, like cv2.imread(path)) or [H, W] (gray, like cv2.imread(path, 0))."""
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    _lib.check_device(device)
    buf = np.frombuffer(data, dtype=np.uint8)
    w, h, _ = jpeg_info(data)
    shape = (h, w) if gray else (h, w, 3)
    if out is None:
        out = torch.empty(shape, dtype=torch.uint8, device=device)
    elif tuple(out.shape) != shape or out.dtype != torch.uint8:
        raise RuntimeError("egaze: decode_jpeg output must be a uint8 tensor of shape %s" % (shape,))
    call("egaze_jpeg_decode", buf.ctypes.data, buf.size, int(gray), out, h, w, stream_ptr())
    torch.cuda.current_stream(device).synchronize()      # the host buffer may go away once we return
    return out


def normalize_image(bgr_u8):
    """uint8 [N, H, W, 3] or [H, W, 3] (BGR) -> fp32 [N, 3, H, W]: data/STdatas.py:51-55."""
    x = bgr_u8 if bgr_u8.dim() == 4 else bgr_u8.unsqueeze(0)
    if x.dtype != torch.uint8 or x.shape[-1] != 3:
        raise RuntimeError("egaze: normalize_image expects uint8 [N,H,W,3]")
    x = x.contiguous()
    N, H, W, _ = x.shape
    out = torch.empty((N, 3, H, W), dtype=torch.float32, device=x.device)
    call("egaze_image_norm", x, N, H, W, out, stream_ptr())
    return out


class FlowWindow(object):
    """The 10-frame optical-flow window of V videos walked in temporal order (one push per frame instead of ten decodes)."""

    def __init__(self, videos, height, width, frames=10, device=None):
        device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        _lib.check_device(device)
        self.V, self.T, self.H, self.W = int(videos), int(frames), int(height), int(width)
        self.ring = torch.zeros((self.V, self.T, 2, self.H, self.W), dtype=torch.uint8, device=device)
        self.pushed = 0

    def reset(self):
        self.pushed = 0

    def push(self, flow_x, flow_y):
        """flow_x, flow_y: uint8 [V, H, W] (or [H, W] when V == 1): the newest frame of every video."""
        fx = flow_x.reshape(self.V, self.H, self.W).contiguous()
        fy = flow_y.reshape(self.V, self.H, self.W).contiguous()
        if fx.dtype != torch.uint8 or fy.dtype != torch.uint8:
            raise RuntimeError("egaze: flow frames must be uint8")
        call("egaze_flow_push", self.ring, fx, fy, self.V, self.T, self.H, self.W, self.pushed % self.T, stream_ptr())
        self.pushed += 1

    def stack(self):
        """-> fp32 [V, 2*frames, H, W]: channels [x_n, y_n, x_{n-1}, y_{n-1}, ...] normalised to [-1, 1] (STdatas.py:59-68)."""
        if self.pushed == 0:
            raise RuntimeError("egaze: FlowWindow.stack() before the first push")
        out = torch.empty((self.V, 2 * self.T, self.H, self.W), dtype=torch.float32, device=self.ring.device)
        call("egaze_flow_stack", self.ring, self.V, self.T, self.H, self.W, (self.pushed - 1) % self.T, min(self.pushed, self.T),
             out, stream_ptr())
        return out
