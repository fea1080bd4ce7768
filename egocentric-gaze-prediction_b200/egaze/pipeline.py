"""Streaming SP -> AT -> LF gaze pipeline on the device.

The reference runs its three stages through the file system: AT.extract_late (AT.py:199-253) walks every video frame by frame
(batch 1), writes the SP gaze map and the AT attention map as 8-bit images (AT.py:228-230, 249-252), and the LF stage reads them
back (data/lateDataset.py:21-34).  `GazePipeline.step()` is that loop body for B videos advancing together, entirely on the
device: SP forward (hook on features_s) -> gaze point -> 3x3 crop-mean (or the aligned 48x48 crop) -> fixation: reuse the crop
weights / saccade: lstmnet with per-video state -> channel-weighted map, min-max normalised -> x16 bilinear -> late fusion.

`quantize=True` reproduces the two uint8 round trips of the file-based hand-over bit for bit (np.uint8(255*x) / 255 on both
maps -- the reference's JPEG compression of those files is lossy and is not reproduced); `quantize=False` keeps fp32 between the
stages.  The gaze point is the arg-max of the target map when one is given (what AT.extract_late effectively uses: computeAAEAUC's
2-D branch returns the target's arg-max, utils.py:124-140, AT.py:233-238 -- SURVEY 0 "quirks to preserve"), else of the predicted map.
"""
import torch

from . import ops, _lib
from ._lib import call, stream_ptr


def quant_u8(x):
    """np.uint8(255 * x) / 255 on the device (egaze_quant_u8)."""
    x = x.contiguous().float()
    out = torch.empty_like(x)
    call("egaze_quant_u8", x, x.numel(), out, stream_ptr())
    return out


def argmax_point(maps):
    """[B,(1,)H,W] -> int32 [B,2] (row, col): centre of the arg-max plateau, rounded down like an index."""
    m = maps.reshape(maps.shape[0], maps.shape[-2], maps.shape[-1]).contiguous().float()
    return ops.floss_centroid(m).floor().to(torch.int32)


class GazePipeline(object):
    def __init__(self, model_sp, lstm, late_fusion, crop_size=3, align=False, quantize=False, lf_order="train"):
        """lf_order: 'train' = late_fusion(AT map, SP map) as LF.py:90,119 call it; 'demo' = late_fusion(SP map, AT map) as
        run_spatialstream.py:138 does (a reference quirk)."""
        self.model, self.lstm, self.lf = model_sp, lstm, late_fusion
        self.crop_size, self.align, self.quantize, self.lf_order = int(crop_size), bool(align), bool(quantize), lf_order
        self.hidden = None
        self._blobs = []
        self._hook = self.model._modules.get('features_s').register_forward_hook(lambda m, i, o: self._blobs.append(o))  # AT.py:105

    def close(self):
        self._hook.remove()

    def reset(self, batch=None):
        """Start of a new set of videos (AT.py:129-131 resets the LSTM state at video boundaries)."""
        self.hidden = None

    @torch.no_grad()
    def step(self, x_s, x_t, fixsac=None, target=None):
        """One frame of each of the B videos.  x_s [B,3,H,W], x_t [B,20,H,W]; fixsac [B] (1 = fixation, default: all saccades);
        target [B,1,H,W] optional ground-truth map.  -> dict(sp, at, fused, gaze)."""
        _lib.check_device(x_s.device)
        B = x_s.shape[0]
        self._blobs.clear()
        sp = self.model(x_s, x_t)                                                     # AT.py:225
        feat = self._blobs[0]                                                         # AT.py:226
        sp_q = quant_u8(sp) if self.quantize else sp                                  # AT.py:228-230
        gaze = argmax_point(target if target is not None else sp_q)                   # AT.py:233 (see module docstring)
        vec = (ops.crop_align_mean if self.align else ops.crop_mean)(feat, gaze, self.crop_size)   # AT.py:235-241
        if self.hidden is None:
            z = torch.zeros(2, B, 512, device=x_s.device)
            self.hidden = (z, z.clone())
        out, (h, c) = self.lstm(vec.unsqueeze(0), self.hidden)                        # AT.py:245-246
        if fixsac is None:
            w, self.hidden = out.squeeze(0), (h, c)
        else:
            sac = torch.as_tensor(fixsac, device=x_s.device).reshape(B) != 1
            w = torch.where(sac[:, None], out.squeeze(0), vec)                        # AT.py:242-248
            self.hidden = (torch.where(sac[None, :, None], h, self.hidden[0]), torch.where(sac[None, :, None], c, self.hidden[1]))
        at = ops.weighted_map(w, feat)                                                # AT.py:58-66
        at_q = quant_u8(at) if self.quantize else at                                  # AT.py:249-250
        up = ops.bilinear_up(at_q.unsqueeze(1), x_s.shape[-1] // at.shape[-1], False) # AT.py:251 / run_spatialstream.py:136
        fused = self.lf(up, sp_q) if self.lf_order == "train" else self.lf(sp_q, up)  # LF.py:119 / run_spatialstream.py:138
        return {"sp": sp, "at": at, "fused": fused, "gaze": gaze}
