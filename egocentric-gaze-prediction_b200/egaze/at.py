"""The attention-transition (AT) step of the reference on the device, batched over independent videos.

Reference: AT.extract_late's loop body, AT.py:224-252 -- per frame (batch 1, frames of one video in temporal order):
SP forward with a hook on `features_s`, gaze point from `computeAAEAUC` (the arg-max of the target map, utils.py:104),
3x3 crop of the 512x14x14 map around it -> channel weights; on a FIXATION frame (`fixsac == 1`) those weights are used
as they are, on a SACCADE frame they go through `lstmnet` (whose state is carried from saccade to saccade); the weighted
channel sum, min-max normalised, is the attention map.

Here B videos advance together: the LSTM runs for every sample and the fixation samples keep their old state and their
crop weights (a select), which is what running each video on its own through the reference loop gives.
"""
import torch

from . import ops


def at_step(feat, gaze, fixsac, lstm, hidden, crop_size=3, align=False):
    """feat [B,512,h,w] hooked conv5_3 maps (CUDA); gaze [B,2] int pixel coordinates in the 224x224 frame (row, col);
    fixsac [B] (1 = fixation); lstm: models.LSTMnet.lstmnet; hidden: (h, c) each [2,B,512].
    -> (attention map [B,h,w] in [0,1], channel weights [B,512], new hidden)."""
    B = feat.shape[0]
    vec = ops.crop_align_mean(feat, gaze, crop_size) if align else ops.crop_mean(feat, gaze, crop_size)   # AT.py:231-241
    with torch.no_grad():
        out, (h, c) = lstm(vec.unsqueeze(0), hidden)                                                      # AT.py:245-246
    sac = torch.as_tensor(fixsac, device=feat.device).reshape(B) != 1
    w = torch.where(sac[:, None], out.squeeze(0), vec)                                                   # AT.py:242-248
    h = torch.where(sac[None, :, None], h, hidden[0])
    c = torch.where(sac[None, :, None], c, hidden[1])
    return ops.weighted_map(w, feat), w, (h, c)


def at_sequence(feats, gazes, fixsacs, lstm, hidden=None, crop_size=3, align=False):
    """T steps of at_step: feats [T,B,512,h,w], gazes [T,B,2], fixsacs [T,B] -> maps [T,B,h,w], final hidden."""
    T, B = feats.shape[0], feats.shape[1]
    if hidden is None:
        hidden = (torch.zeros(2, B, 512, device=feats.device), torch.zeros(2, B, 512, device=feats.device))
    maps = []
    for t in range(T):
        m, _, hidden = at_step(feats[t], gazes[t], fixsacs[t], lstm, hidden, crop_size, align)
        maps.append(m)
    return torch.stack(maps), hidden
