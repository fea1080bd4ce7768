#!/usr/bin/env python
"""bench.py -- the driver's benchmark contract for the egaze-b200 hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload sp_train|sp_fwd|pipeline_fwd|full_train|at_seq]
                    [--impl reference]

Metric (BASELINE.json): SP+AT+LF gaze-map frames/s at 224x224, batch 32 per GPU.  One "step" = one pass of the hot
path over one synthetic batch.  `full_train` (the default: BASELINE configs[3] per rank, the configuration the
metric is quoted on) = SP two-stream train step (forward + floss + backward + Adam), AT step on the hooked conv5_3 map, LF
train step on (AT map, SP map) with floss + Adam; `sp_train` = the SP part alone (BASELINE configs[1]); `sp_fwd` = eval-mode
two-stream forward; `pipeline_fwd` = SP forward -> AT step -> LF forward (gaze-map inference); `at_seq` = BASELINE configs[2]: crop-mean -> 2-layer LSTM ->
channel-weighted map over 16 feature sequences of 30 steps (a "frame" is one sequence step).  N > 1 ranks (torchrun) shard frames: weak scaling, B=32 per rank, one NCCL
allreduce of the weight gradients per training step and no collective for inference.

Prints ONE JSON line (rank 0).  `value` = device-resident throughput; `e2e` = same metric through the public API
(egaze.graph.GraphedStep over the drop-in modules) with HOST (pinned) inputs and a host read of the result inside the timed
region; `e2e_dropin` = the reference's own loop bodies (SP.py:125-142, AT.py:224-248, LF.py:86-99) written out literally over
the drop-in modules: eager launches, blocking `.to(device)` copies, every `loss.item()` the reference calls.
"""
import argparse
import contextlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "egocentric-gaze-prediction_b200")
for _p in (PKG, ROOT):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

# one metric string per workload: the line must say what was timed
METRICS = {
    "full_train": "SP+AT+LF gaze-map frames/sec at 224x224 b32 (train step: SP fwd/bwd + AT step + LF fwd/bwd, BASELINE configs[3])",
    "sp_train": "SP two-stream gaze-map frames/sec at 224x224 b32 (train step: fwd + floss + bwd + Adam, BASELINE configs[1])",
    "sp_fwd": "SP two-stream gaze-map frames/sec at 224x224 b32 (eval forward)",
    "pipeline_fwd": "SP+AT+LF gaze-map frames/sec at 224x224 b32 (inference: SP fwd -> AT step -> LF fwd)",
    "at_seq": "AT sequence-steps/sec over 512x14x14 feature sequences, seq_len 30, batch 16 (BASELINE configs[2])",
}
UNIT = "frames/s"


def metric_name(workload, B=32, S=224):
    m = METRICS[workload]
    return m if (B, S) == (32, 224) else m.replace("224x224 b32", "%dx%d b%d" % (S, S, B))
# algorithmic FLOPs per frame (SURVEY 8d / BASELINE.md 4): 2 FLOP per MAC of the reference's direct convolutions
FLOP_SP_FWD = 114.167e9
FLOP_SP_TRAIN = 341.171e9
FLOP_LF_FWD = 1.2147e9


def load_traffic(workload):
    """DRAM bytes per launch of the tcgen05 conv kernels (dram__bytes_read.sum + dram__bytes_write.sum averaged over the
    conv/wgrad launches of one step), from the committed ncu capture of this same command (profiles/conv_traffic.json,
    raw per-launch list beside it)."""
    try:
        with open(os.path.join(ROOT, "profiles", "conv_traffic.json")) as fh:
            d = json.load(fh)
        return float(d[workload]["bytes_per_launch"])
    except Exception:
        return None


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            d = json.load(fh)
        return d, "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler(threading.Thread):
    """Samples SM clocks / throttle reasons while the timed region runs: NVML in-process every 20 ms when available
    (nvidia_ml_py), else one `nvidia-smi` query per 200 ms."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._halt = index, [], threading.Event()
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            # LOCAL_RANK indexes CUDA_VISIBLE_DEVICES; NVML indexes the physical devices
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        n = self.nvml
        sm = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
        mx = n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM)
        r = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
        bits = [n.nvmlClocksThrottleReasonHwSlowdown, n.nvmlClocksThrottleReasonHwThermalSlowdown,
                n.nvmlClocksThrottleReasonSwThermalSlowdown, n.nvmlClocksThrottleReasonSwPowerCap]
        return [str(sm), str(mx), "0"] + ["Active" if (r & b) else "Not Active" for b in bits]

    def run(self):
        while not self._halt.is_set():
            try:
                if self.nvml is not None:
                    self.rows.append(self._sample_nvml())
                else:
                    out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                    self.rows.append([c.strip() for c in out.strip().split(",")])
            except Exception:
                pass
            self._halt.wait(0.02 if self.nvml is not None else 0.2)

    def finish(self):
        self._halt.set()
        self.join(timeout=3)
        sm, mx, reasons = [], 0.0, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
                for n, v in zip(self.NAMES, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        # the sampler only runs between the two barriers of the timed region: every sample is "under load"
        return {"sm_mhz": float(np.median(sm)) if sm else 0.0, "sm_min_mhz": float(min(sm)) if sm else 0.0, "sm_max_mhz": mx,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvml" if self.nvml is not None else "nvidia-smi"}


# --------------------------------------------------------------------------------------------------------------------
# workloads
# --------------------------------------------------------------------------------------------------------------------
def synth_sp_inputs(B, S, seed=1234):
    """The bench's synthetic batch: an RGB frame at normalised-image scale (data/STdatas.py:54-55), a
    20-channel flow stack quantised to 256 levels in [-1, 1] (data/STdatas.py:66-68) and a min-max-normalised, 8-bit
    quantised Gaussian gaze blob (data/dataset_preprocessing.py:113-119, data/STdatas.py:64,69-71).  Same numbers as the
    generator the parity fixtures were made with (tests/test_bench_output.py checks that), kept here so that nothing under
    oracle/ is on the measured path."""
    rs = np.random.RandomState(seed)
    x_s = rs.randn(B, 3, S, S).astype(np.float32)
    x_t = ((rs.randint(0, 256, (B, 20, S, S)).astype(np.float32) / 255 - 0.5) / 0.5).astype(np.float32)
    rows, cols = np.meshgrid(np.arange(S, dtype=np.float64), np.arange(S, dtype=np.float64), indexing='ij')
    sig_r, sig_c = S * 16.3 / 224, S * 12.25 / 224
    gt = np.empty((B, 1, S, S), np.float32)
    for b in range(B):
        cr, cc = (rs.rand(2) * 0.7 + 0.15) * S
        blob = np.exp(-((rows - cr) ** 2 / (2 * sig_r ** 2) + (cols - cc) ** 2 / (2 * sig_c ** 2)))
        blob = (blob - blob.min()) / (blob.max() - blob.min())
        gt[b, 0] = np.round(blob * 255) / 255
    return x_s, x_t, gt


class Workload(object):
    def __init__(self, name, B, S, rank, world, device):
        from utils import make_layers, cfg
        from models.model_SP import model_SP
        from models.late_fusion import late_fusion
        from models.LSTMnet import lstmnet
        import floss as floss_mod
        self.name, self.B, self.S, self.device, self.world = name, B, S, device, world
        torch.manual_seed(0)  # identical replicas on every rank
        from egaze.ddp import shard_seed
        self.model = model_SP(make_layers(cfg['D'], 3), make_layers(cfg['D'], 20)).to(device) if name != "at_seq" else None
        self.crit = floss_mod.floss()
        x_s, x_t, gt = synth_sp_inputs(B, S, shard_seed(1234, rank))
        self.host = [torch.from_numpy(a).pin_memory() for a in (x_s, x_t, gt)]
        self.dev = [t.to(device) for t in self.host]
        self.h2d_bytes = sum(t.numel() * 4 for t in self.host[:2]) + (self.host[2].numel() * 4 if name == "sp_train" else 0)
        self.flat = None
        self.flat_lf = None
        self.reducer = None
        self.ddp_kind = None
        if name == "sp_train":
            self.model.train()
            self._make_optimizers(False)
            self.flop = FLOP_SP_TRAIN
            self.d2h_bytes = 4
            if world > 1:
                self._make_flat_grads()
        elif name == "sp_fwd":
            self.model.eval()
            self.flop = FLOP_SP_FWD
            self.d2h_bytes = B * S * S * 4
        elif name == "pipeline_fwd":
            self.model.eval()
            self.lstm = lstmnet().to(device).eval()
            self.lf = late_fusion().to(device).eval()
            self.feats = []
            self.model._modules.get('features_s').register_forward_hook(lambda m, i, o: self.feats.append(o))  # AT.py:105
            self.hidden = (torch.zeros(2, B, 512, device=device), torch.zeros(2, B, 512, device=device))
            self.gaze = torch.randint(0, S, (B, 2), generator=torch.Generator().manual_seed(5 + rank)).int().to(device)
            self.flop = FLOP_SP_FWD + FLOP_LF_FWD
            self.d2h_bytes = B * S * S * 4
        elif name == "full_train":
            # BASELINE configs[3]: SP train step + AT step on the hooked features_s map + LF train step (LF.py:79-105)
            self.model.train()
            self.lstm = lstmnet().to(device).eval()
            self.lf = late_fusion().to(device).train()
            self._make_optimizers(False)
            self.lf_stream = torch.cuda.Stream(device=device)
            self.feats = []
            self.model._modules.get('features_s').register_forward_hook(lambda m, i, o: self.feats.append(o))  # AT.py:105
            self.hidden = (torch.zeros(2, B, 512, device=device), torch.zeros(2, B, 512, device=device))
            self.gaze = torch.randint(0, S, (B, 2), generator=torch.Generator().manual_seed(5 + rank)).int().to(device)
            self.flop = FLOP_SP_TRAIN + 3 * FLOP_LF_FWD
            self.h2d_bytes = sum(t.numel() * 4 for t in self.host)
            self.d2h_bytes = 8
            if world > 1:
                self._make_flat_grads(list(self.model.parameters()) + list(self.lf.parameters()))
        elif name == "at_seq":
            # BASELINE configs[2]: 16 feature sequences x 30 steps of 512x14x14 post-ReLU maps
            self.T, self.NB = 30, 16
            g = torch.Generator().manual_seed(shard_seed(77, rank))
            self.model = None
            self.lstm = lstmnet().to(device).eval()
            feat = torch.relu(torch.randn(self.T * self.NB, 512, S // 16, S // 16, generator=g))
            gaze = torch.randint(0, S, (self.T * self.NB, 2), generator=g).int()
            self.host = [feat.pin_memory(), gaze.pin_memory(), torch.zeros(1).pin_memory()]
            self.dev = [t.to(device) for t in self.host]
            self.h2d_bytes = feat.numel() * 4 + gaze.numel() * 4
            self.d2h_bytes = self.T * self.NB * (S // 16) ** 2 * 4
            self.flop = 0.0
            self.B = self.T * self.NB   # "frames" per step = sequence steps
        else:
            raise SystemExit("unknown workload %r" % name)

    def _make_flat_grads(self, params=None):
        """Data-parallel gradient averaging.  model_SP: egaze.ddp.OverlappedGradReducer -- the backward node all-reduces each
        segment of its flat fp32 gradient buffer (decoder -> fusion/bn -> deep trunk layers -> the rest) on a communication
        stream as soon as that segment is final, and the optimiser reads the averaged values in place.  The 49 KB of late-fusion
        gradients go through one small flat bucket after the LF backward.  EGAZE_BENCH_DDP=flat: round-1 scheme (ONE all-reduce
        of one flat buffer after the whole backward)."""
        from egaze.ddp import FlatGradBucket, attach_reducer, broadcast_parameters
        broadcast_parameters(self.model, 0)
        if params is not None:
            broadcast_parameters(self.lf, 0)
        if os.environ.get("EGAZE_BENCH_DDP", "overlap") == "flat":
            self.flat = FlatGradBucket(self.model.parameters() if params is None else params, self.device)
            self.ddp_kind = "one flat all-reduce after the backward"
        else:
            self.reducer = attach_reducer(self.model)
            self.flat_lf = FlatGradBucket(self.lf.parameters(), self.device) if params is not None else None
            self.ddp_kind = "segmented all-reduce overlapped with the backward (4 segments) + 1 small LF bucket"

    def _make_optimizers(self, capturable, stock=False):
        """Adam as the reference builds it (gaze_full.py:11 default lr, SP.py:113, LF.py:77).  The bench's own loops use the fused
        multi-tensor egaze.optim.Adam (same arguments, state and arithmetic; EGAZE_BENCH_ADAM=torch: stock torch.optim.Adam);
        the drop-in loop (stock=True) always steps stock torch.optim.Adam like the reference."""
        if not stock and os.environ.get("EGAZE_BENCH_ADAM", "fused") == "fused":
            from egaze.optim import Adam
            kw = {}
            self.adam_kind = "egaze.optim.Adam (fused multi-tensor, maintains the packed weight copies)"
        else:
            Adam = torch.optim.Adam
            kw = {"capturable": True} if capturable else {}
            self.adam_kind = "torch.optim.Adam"
        self.opt = Adam(self.model.parameters(), lr=1e-7, **kw)
        if self.name == "full_train":
            self.opt_lf = Adam(self.lf.parameters(), lr=1e-7, **kw)

    def capture(self):
        """Capture one whole training step (forward, floss, backward on all streams, the AT step and the LF train step on
        their side stream, the NCCL gradient all-reduce under torchrun, the Adam steps) on the resident inputs into a CUDA
        graph (egaze.graph.GraphedStep); `replay()` then runs a step with no host work at all.  EGAZE_BENCH_GRAPH=0 turns it
        off (eager module calls everywhere); returns False -- and the caller stays on the eager path -- if the capture fails."""
        from egaze.graph import GraphedStep
        try:
            self._make_optimizers(True)
            opts = [self.opt] + ([self.opt_lf] if self.name == "full_train" else [])
            self.graph = GraphedStep(self.step, self.dev, optimizers=opts, own_inputs=True, restore=False)
            return True
        except Exception as exc:  # noqa: BLE001 -- any capture problem means: stay eager
            sys.stderr.write("bench: CUDA-graph capture failed (%s: %s); staying on the eager path\n" % (type(exc).__name__, exc))
            self.graph = None
            try:
                torch.cuda.synchronize(self.device)
            except Exception:
                pass
            self._make_optimizers(False)
            return False

    def replay(self):
        return self.graph.replay()

    def drop_graph(self):
        """Destroy the captured graph (and its memory pool) and go back to the eager optimisers."""
        import gc
        torch.cuda.synchronize(self.device)
        try:
            self.graph.release()
        except Exception as exc:  # noqa: BLE001 -- the measurements are taken; a failed clean-up must not lose them
            sys.stderr.write("bench: releasing the CUDA graph failed (%s: %s)\n" % (type(exc).__name__, exc))
        self.graph = None
        self._make_optimizers(False)
        gc.collect()
        torch.cuda.synchronize(self.device)

    def dropin_step(self, sample):
        """The reference's loop bodies written out literally over the drop-in modules: SP.trainSP (SP.py:126-142) and, for
        `full_train`, the AT step of AT.extract_late (AT.py:224-248, on the device) and LF.trainLate (LF.py:86-99, without its
        per-batch scipy metric).  `sample` holds HOST tensors as a DataLoader yields them.  Returns the number of host syncs."""
        from egaze import ops
        input_s = sample['image']
        target = sample['gt']
        input_t = sample['flow']
        input_s = input_s.float().to(self.device)
        input_t = input_t.float().to(self.device)
        target = target.float().to(self.device)
        if self.name == "full_train":
            self.feats.clear()
        output = self.model(input_s, input_t)
        target = target.view(output.size())
        loss = self.crit(output, target)
        self.loss_mini_batch = loss.item()
        loss.backward()
        if self.flat is not None:
            self.flat.allreduce()
        self.opt.step()
        self.opt.zero_grad()
        self.loss_avg = loss.item()
        if self.name != "full_train":
            return 2
        with torch.no_grad():
            feature_s = self.feats[0]
            chn_weight = ops.crop_mean(feature_s, self.gaze, 3)
            chn_weight, hidden = self.lstm(chn_weight.unsqueeze(0), self.hidden)
            for dst, src in zip(self.hidden, hidden):
                dst.copy_(src)
            feat = ops.bilinear_up(ops.weighted_map(chn_weight.squeeze(0), feature_s).unsqueeze(1), 16, False)
        out = self.lf(feat, output.detach())
        loss = self.crit(out, target)
        self.loss_lf_avg = loss.item()
        self.opt_lf.zero_grad()
        loss.backward()
        if self.flat is not None:
            self.flat.allreduce()
        if self.flat_lf is not None:
            self.flat_lf.allreduce()
        self.opt_lf.step()
        return 3

    def step(self, x_s, x_t, gt):
        """One pass of the hot path; returns the tensor a user would read back."""
        from egaze import ops
        if self.name == "sp_train":
            if self.flat is not None:
                self.flat.zero()
            else:
                self.opt.zero_grad(set_to_none=True)
            out = self.model(x_s, x_t)
            loss = self.crit(out, gt.view(out.size()))
            loss.backward()          # (with an attached OverlappedGradReducer the gradients come back already averaged)
            if self.flat is not None:
                self.flat.allreduce()
            self.opt.step()
            return loss
        if self.name == "full_train":
            if self.flat is not None:
                self.flat.zero()
            else:
                self.opt.zero_grad(set_to_none=True)
                if self.flat_lf is not None:
                    self.flat_lf.zero()
                else:
                    self.opt_lf.zero_grad(set_to_none=True)
            self.feats.clear()
            out = self.model(x_s, x_t)                                         # SP.py:132
            gtv = gt.view(out.size())
            loss = self.crit(out, gtv)
            # The AT step and the LF train step only need the SP forward's outputs: they are enqueued on a second stream and
            # run (small, HBM / latency-bound kernels) under the tensor-bound SP backward.  EGAZE_BENCH_LF_STREAM=0: serial.
            lf_stream = self.lf_stream if os.environ.get("EGAZE_BENCH_LF_STREAM", "1") != "0" else None
            main = torch.cuda.current_stream(self.device)
            if lf_stream is None:
                loss.backward()
            else:
                lf_stream.wait_stream(main)
            with torch.cuda.stream(lf_stream) if lf_stream is not None else contextlib.nullcontext():
                with torch.no_grad():                                          # AT.py:224-252 on the hooked map, batched
                    feat = self.feats[0]
                    vec = ops.crop_mean(feat, self.gaze, 3)
                    w, hidden = self.lstm(vec.unsqueeze(0), self.hidden)
                    for dst, src in zip(self.hidden, hidden):   # the state stays in the same buffers (CUDA-graph replay)
                        dst.copy_(src)
                    amap = ops.weighted_map(w.squeeze(0), feat)
                    up = ops.bilinear_up(amap.unsqueeze(1), 16, False)
                fused = self.lf(up, out.detach())                              # LF.py:90
                loss_lf = self.crit(fused, gtv)                                # LF.py:91
                loss_lf.backward()
            if lf_stream is not None:
                for t in (feat, out, gtv):
                    t.record_stream(lf_stream)
                loss.backward()                                                # SP backward on the main stream, concurrently
                main.wait_stream(lf_stream)
                for p_ in self.lf.parameters():
                    if p_.grad is not None:
                        p_.grad.record_stream(main)
                loss_lf.record_stream(main)
            if self.flat is not None:
                self.flat.allreduce()
            if self.flat_lf is not None:
                self.flat_lf.allreduce()
            self.opt.step()
            self.opt_lf.step()
            return torch.stack((loss.detach(), loss_lf.detach()))
        if self.name == "at_seq":
            feat, gaze = x_s, x_t
            with torch.no_grad():
                vec = ops.crop_mean(feat, gaze, 3)                             # AT.py:236-241, all T*B frames at once
                hid = (torch.zeros(2, self.NB, 512, device=self.device), torch.zeros(2, self.NB, 512, device=self.device))
                w, _ = self.lstm(vec.view(self.T, self.NB, 512), hid)          # AT.py:245-246: the recurrence over T
                return ops.weighted_map(w.reshape(self.T * self.NB, 512), feat)  # AT.py:248 per frame
        with torch.no_grad():
            if self.name == "sp_fwd":
                return self.model(x_s, x_t)
            self.feats.clear()
            out = self.model(x_s, x_t)
            feat = self.feats[0]
            vec = ops.crop_mean(feat, self.gaze, 3)                        # AT.py:236-241
            w, self.hidden = self.lstm(vec.unsqueeze(0), self.hidden)      # AT.py:245-246 (saccade branch)
            amap = ops.weighted_map(w.squeeze(0), feat)                    # AT.py:248
            up = ops.bilinear_up(amap.unsqueeze(1), 16, False)             # run_spatialstream.py:136
            return self.lf(up, out)                                        # LF.py:90 argument order (AT map, SP map)


def count_launches(fn):
    from egaze import _lib
    n0 = _lib.launch_counter()
    fn()
    return _lib.launch_counter() - n0


def cpu_reference_fps(workload, B, S, steps, warmup, threads):
    """The reference's CPU path (stock torch.nn modules + oneDNN, oracle/torch_cpu_ref.py) on this box's host cores."""
    from oracle import torch_cpu_ref as ref
    from oracle import egaze_oracle as orc
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    m = ref.ModelSP()
    x_s, x_t, gt = [torch.from_numpy(a) for a in synth_sp_inputs(B, S, 1234)]
    if workload == "at_seq":
        # BASELINE configs[2] the way AT.extract_late walks a video: frame by frame, batch = the 16 sequences
        T, NB = 30, 16
        lstm = ref.LSTMNet().eval()
        g = torch.Generator().manual_seed(77)
        feat = torch.relu(torch.randn(T, NB, 512, S // 16, S // 16, generator=g)).numpy()
        gaze = torch.randint(0, S, (T, NB, 2), generator=g).numpy()

        def step():
            hid = (torch.zeros(2, NB, 512), torch.zeros(2, NB, 512))
            with torch.no_grad():
                for t in range(T):
                    vec = torch.from_numpy(orc.crop_mean(feat[t], gaze[t], 3))
                    w, hid = lstm(vec.unsqueeze(0), hid)
                    orc.get_weighted(w.squeeze(0).numpy(), feat[t])
        B = T * NB
    elif workload in ("sp_train", "full_train"):
        m.train()
        opt = torch.optim.Adam(m.parameters(), lr=1e-7)
        full = workload == "full_train"
        if full:
            lf, lstm = ref.LateFusion().train(), ref.LSTMNet().eval()
            opt_lf = torch.optim.Adam(lf.parameters(), lr=1e-7)

        def step():
            opt.zero_grad()
            blobs = []
            h = m.features_s.register_forward_hook(lambda mod, i, o: blobs.append(o.detach()))
            out = m(x_s, x_t)
            h.remove()
            loss = ref.floss(out, gt)
            loss.backward()
            opt.step()
            if full:
                opt_lf.zero_grad()
                with torch.no_grad():
                    feat = blobs[0]
                    vec = torch.from_numpy(orc.crop_mean(feat.numpy(), [[S // 2, S // 2]] * B, 3))
                    w, _ = lstm(vec.unsqueeze(0), (torch.zeros(2, B, 512), torch.zeros(2, B, 512)))
                    amap = torch.from_numpy(orc.get_weighted(w.squeeze(0).numpy(), feat.numpy()))
                    up = torch.nn.functional.interpolate(amap.unsqueeze(1), scale_factor=16, mode="bilinear", align_corners=False)
                loss_lf = ref.floss(lf(up, out.detach()), gt)
                loss_lf.backward()
                opt_lf.step()
    else:
        m.eval()
        lf = ref.LateFusion().eval() if workload == "pipeline_fwd" else None
        lstm = ref.LSTMNet().eval() if workload == "pipeline_fwd" else None

        def step():
            with torch.no_grad():
                if lf is None:
                    return m(x_s, x_t)
                blobs = []
                h = m.features_s.register_forward_hook(lambda mod, i, o: blobs.append(o))
                out = m(x_s, x_t)
                h.remove()
                feat = blobs[0]
                gaze = [[S // 2, S // 2]] * B
                vec = torch.from_numpy(orc.crop_mean(feat.numpy(), gaze, 3))
                w, _ = lstm(vec.unsqueeze(0), (torch.zeros(2, B, 512), torch.zeros(2, B, 512)))
                amap = torch.from_numpy(orc.get_weighted(w.squeeze(0).numpy(), feat.numpy()))
                up = torch.nn.functional.interpolate(amap.unsqueeze(1), scale_factor=16, mode="bilinear", align_corners=False)
                return lf(up, out)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    sample = ("at_seq, 16 sequences x 30 steps of 512x%dx%d maps" % (S // 16, S // 16)) if workload == "at_seq" else \
        "%s, batch %d x %dx%d" % (workload, B, S, S)
    return B / dt, dt, sample


_json_out = [None]


def claim_stdout():
    """stdout carries the ONE JSON line and nothing else: keep a private handle on the real stdout and point fd 1 at
    stderr, so whatever a library prints (NCCL's version banner, oneDNN / OpenMP chatter) cannot land next to the line."""
    if _json_out[0] is None:
        sys.stdout.flush()
        real = os.dup(1)
        os.dup2(2, 1)
        _json_out[0] = os.fdopen(real, "w")
    return _json_out[0]


def emit(line):
    out = claim_stdout()
    out.write(json.dumps(line) + "\n")
    out.flush()


def finish(world):
    """Tear the process group down without ever hanging the launcher: the line is out by now, so a teardown that does not
    return within 30 s (an NCCL communicator that still has work or graphs attached) ends in a hard exit with status 0."""
    sys.stdout.flush()
    sys.stderr.flush()
    if world <= 1:
        return
    guard = threading.Timer(30.0, os._exit, (0,))
    guard.daemon = True
    guard.start()
    try:
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        if torch.distributed.is_initialized():
            torch.distributed.destroy_process_group()
    except Exception as exc:  # noqa: BLE001
        sys.stderr.write("bench: destroy_process_group: %s\n" % exc)
    sys.stderr.flush()
    os._exit(0)


def common_config(args, world):
    """The part of `config` both arms share (the reference arm times a bounded sample of this same workload)."""
    return {"workload": args.workload, "batch_per_gpu": args.batch, "size": args.size, "global_batch": args.batch * world,
            "parallelism": "dp%d" % world,
            "l2": "per-step inputs+activations (>1 GB) exceed the 126 MB L2; no explicit flush"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    claim_stdout()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    threads = os.cpu_count() or 1
    # The reference arm runs the workload's OWN batch (same config as the GPU arm) when the host has the memory for it (a batch-32
    # two-stream train step keeps ~20 GB of activations in PyTorch-CPU), else the bounded batch (--ref-batch).  A batch-32 step
    # takes ~7 s on the GPU box's 16 host threads: --steps / --warmup are honoured up to 5 / 1 there (10 / 2 for the small batch),
    # so the whole arm stays within about a minute.
    B = args.ref_batch
    try:
        import psutil
        if args.ref_batch_auto and psutil.virtual_memory().available > 96 * 2 ** 30:
            B = args.batch
    except Exception:  # noqa: BLE001
        pass
    full = B == args.batch
    steps = max(1, min(args.steps, 5 if full else 10))
    warmup = max(1, min(args.warmup, 1 if full else 2))
    fps, dt, sample = cpu_reference_fps(args.workload, B, args.size, steps, warmup, threads)
    line = {"impl": "reference", "metric": metric_name(args.workload, args.batch, args.size), "value": fps, "unit": UNIT,
            "n_gpus": args.gpus, "steps": steps,
            "warmup": warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": common_config(args, max(world, args.gpus)),
            "detail": {"note": ("reference CPU path (stock torch.nn modules, oneDNN) timed on ONE process on the same workload and "
                                "batch (%s)" if full else
                                "reference CPU path (stock torch.nn modules, oneDNN) timed on ONE process on a bounded sample "
                                "(%s) of the same workload") % sample,
                       "same_batch_as_gpu_arm": full},
            "cpu_baseline": {"value": fps, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": "%s, %d step(s), torch %s CPU (oneDNN), %d threads" % (
                                 sample, steps, torch.__version__, threads)},
            "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default=os.environ.get("EGAZE_BENCH_WORKLOAD", "full_train"), choices=sorted(METRICS))
    ap.add_argument("--impl", default="egaze")
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--size", type=int, default=224)
    ap.add_argument("--ref-batch", type=int, default=4)
    ap.add_argument("--ref-batch-auto", type=int, default=1,
                    help="--impl reference: 1 = time the workload's own batch when the host has > 96 GB of free memory")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-dropin", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    claim_stdout()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the egaze path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=device)
    from egaze import _lib, ops
    W = max(args.warmup, 3)
    K = args.steps
    wl = Workload(args.workload, args.batch, args.size, rank, world, device)
    training = args.workload in ("sp_train", "full_train")

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput ------------------------------------------------------------------------------------
    for _ in range(W):
        wl.step(*wl.dev)
    launches = count_launches(lambda: wl.step(*wl.dev))
    graphed = False
    if os.environ.get("EGAZE_BENCH_GRAPH", "1") == "1" and training:
        graphed = wl.capture()
    run_step = (lambda: wl.replay()) if graphed else (lambda: wl.step(*wl.dev))
    for _ in range(2):
        run_step()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    t_host = time.perf_counter()
    for _ in range(K):
        run_step()
    host_ms = (time.perf_counter() - t_host) * 1e3 / K   # time the host needs to ENQUEUE a step (no synchronisation inside)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / K
    clocks = sampler.finish() if sampler else None

    # ---- end to end: host (pinned) inputs in, result read back, every step -----------------------------------------------
    # What a data loader does: the H2D copy of batch i+1 runs on a copy stream while batch i computes (double-buffered
    # device inputs).  Every step's input bytes cross PCIe inside the timed region and every step's result is read back.
    copy_stream = torch.cuda.Stream(device=device)
    bufs = [[torch.empty_like(t, device=device) for t in wl.host] for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    freed = [torch.cuda.Event() for _ in range(2)]

    def prefetch(i):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(freed[i % 2])
            for b, h in zip(bufs[i % 2], wl.host):
                b.copy_(h, non_blocking=True)
            ready[i % 2].record(copy_stream)

    # Every step's result is read back to the host inside the timed region -- one step LATE (the copy of step i's result is
    # enqueued right behind it and the host waits for it after enqueueing step i+1; the last one before the timer stops),
    # the way a training loop logs its loss without stalling the launch queue.
    host_res = [None, None]
    res_ready = [torch.cuda.Event() for _ in range(2)]

    def e2e_run(n):
        for ev in freed:
            ev.record()
        prefetch(0)
        got = 0
        for i in range(n):
            torch.cuda.current_stream().wait_event(ready[i % 2])
            if i + 1 < n:
                prefetch(i + 1)
            # graphed: egaze.graph.GraphedStep.__call__ copies the batch into the graph's static inputs and replays the step
            res = (wl.graph(*bufs[i % 2]) if graphed else wl.step(*bufs[i % 2])).detach()
            freed[i % 2].record()
            if host_res[i % 2] is None or host_res[i % 2].shape != res.shape:
                host_res[i % 2] = torch.empty(res.shape, dtype=res.dtype, pin_memory=True)
            host_res[i % 2].copy_(res, non_blocking=True)
            res_ready[i % 2].record()
            if i > 0:
                res_ready[(i - 1) % 2].synchronize()      # the host now holds step i-1's result
                got += 1
        res_ready[(n - 1) % 2].synchronize()
        return got + 1

    e2e_run(2)
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    e2e_run(K)
    f1.record()
    barrier()
    ms_e2e = f0.elapsed_time(f1) / K
    if graphed:
        wl.drop_graph()   # before any teardown: the graph holds the captured all-reduce of the NCCL communicator
        barrier()

    # ---- the reference's own loop bodies over the drop-in modules (eager launches, blocking copies, loss.item() syncs) ------
    ms_dropin, dropin_syncs, dropin_host_ms = None, 0, None
    if training and not args.no_dropin:
        sample = {"image": wl.host[0], "flow": wl.host[1], "gt": wl.host[2]}
        wl._make_optimizers(False, stock=True)
        for _ in range(2):
            wl.dropin_step(sample)
        barrier()
        t0 = time.perf_counter()
        for _ in range(K):
            dropin_syncs = wl.dropin_step(sample)
        torch.cuda.synchronize()
        ms_dropin = (time.perf_counter() - t0) * 1e3 / K   # every step ends in a host sync: wall clock == device time
        barrier()
        wl._make_optimizers(False)
    adam_kind = wl.adam_kind if training else None

    # ---- roofline pass: duration of every tcgen05 conv / wgrad launch ------------------------------------------------------
    # Same workload, same process, with a CUDA-event pair around every launch on the stream it is launched on.  The
    # weight-gradient GEMMs normally run on a second stream and overlap other kernels (their event intervals then overlap too
    # and the sum double-counts time), and so do the two trunks; this pass keeps everything on one stream: each number is the
    # kernel's own duration inside a long step.
    R = max(1, min(K, 5))
    knobs = ("EGAZE_WGRAD_STREAM", "EGAZE_TRUNK_STREAM", "EGAZE_BENCH_LF_STREAM")
    prev = {k: os.environ.get(k) for k in knobs}
    for k in knobs:
        os.environ[k] = "0"
    wl.step(*wl.dev)
    ops.conv_timer_reset(True)
    for _ in range(R):
        wl.step(*wl.dev)
    conv_ms, conv_launches = ops.conv_timer_read()   # summed over the R steps
    ops.conv_timer_reset(False)
    for k in knobs:
        if prev[k] is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = prev[k]
    barrier()

    t = torch.tensor([ms, ms_e2e, ms_dropin or 0.0], device=device, dtype=torch.float64)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    ms, ms_e2e, ms_dropin_max = t.tolist()
    if rank != 0:
        return finish(world)

    peaks, peak_src = load_peaks()
    frames = wl.B * world
    value = frames / ms * 1e3
    conv_flop_step = (wl.flop - (FLOP_LF_FWD if args.workload == "pipeline_fwd" else 0.0)
                      - (3 * FLOP_LF_FWD if args.workload == "full_train" else 0.0)) * args.batch
    conv_tflops = conv_flop_step * R / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
    peak_tf = peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"])
    traffic = load_traffic(args.workload)
    if args.workload == "at_seq":
        # HBM-bound: the 480 x 512 x 14 x 14 fp32 feature stack is read once by the weighted-map kernel
        alg_bytes = wl.dev[0].numel() * 4 + 17.9e6
        roofline = {"bound": "hbm", "kernel": "weighted_map_kernel (reads the whole feature stack) + lstm weights",
                    "achieved": alg_bytes / (ms * 1e-3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                    "frac": alg_bytes / (ms * 1e-3) / 1e9 / peaks["hbm_gbs"], "peak_source": "%s hbm_gbs" % peak_src,
                    "note": "achieved = algorithmic bytes / whole-step time (the step is 3 C-ABI calls)", "traffic": None}
    else:
        roofline = {"bound": "tensor", "kernel": "tcgen05 conv kernels (conv3x3_tc fprop/dgrad + wgrad_tc launches of the SP step)",
                    "achieved": conv_tflops, "peak": peak_tf, "unit": "TFLOP/s", "frac": conv_tflops / peak_tf,
                    "peak_source": "%s bf16_tflops_sustained (kernel timed inside a long step)" % peak_src,
                    "launches_per_step": conv_launches / max(R, 1), "kernel_ms_per_step": conv_ms / max(R, 1),
                    "algorithmic_tflop_per_step": conv_flop_step / 1e12,
                    "measured": "CUDA events around every launch, %d steps after the timed regions, single-stream order" % R,
                    "step_tflops": wl.flop * args.batch / (ms * 1e-3) / 1e12,
                    "traffic": traffic}
    cfg = common_config(args, world)
    cfg["batch_per_gpu"] = wl.B
    cfg["global_batch"] = frames
    line = {
        "metric": metric_name(args.workload, args.batch, args.size), "value": value, "unit": UNIT, "n_gpus": world, "steps": K,
        "warmup": W, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": ops.dtype_string(), "data": "synthetic",
        "config": cfg,
        "detail": {"precision_mode": ops.precision(), "optimizer": adam_kind, "gradient_averaging": wl.ddp_kind,
                   "device_loop": "CUDA graph replay of the whole step" if graphed else "eager launches"},
        "clocks": clocks,
        "e2e": {"value": frames / ms_e2e * 1e3, "unit": UNIT, "h2d_bytes_per_step": wl.h2d_bytes,
                "d2h_bytes_per_step": wl.d2h_bytes,
                "call": "egaze.graph.GraphedStep (whole step replayed as one CUDA graph)" if graphed else "eager module calls",
                "note": "pinned host inputs prefetched one step ahead on a copy stream; every step's result copied to pinned "
                        "host memory and awaited by the host one step later (the last one inside the timed region)"},
        "gpu_launches": launches * K, "host_enqueue_ms_per_step": host_ms,
        "roofline": roofline,
    }
    if ms_dropin is not None:
        line["e2e_dropin"] = {"value": frames / ms_dropin_max * 1e3, "unit": UNIT, "ms_per_step": ms_dropin_max,
                              "h2d_bytes_per_step": wl.h2d_bytes, "d2h_bytes_per_step": 4 * dropin_syncs,
                              "host_syncs_per_step": dropin_syncs,
                              "call": "the reference's loop bodies verbatim over the drop-in modules (SP.py:126-142"
                                      + (", AT.py:224-248, LF.py:86-99" if args.workload == "full_train" else "")
                                      + "): eager launches, blocking .to(device) copies from pinned memory, every loss.item()"}
    if not args.no_cpu_baseline and world == 1:
        threads = os.cpu_count() or 1
        cpu_steps = 5   # ~5-10 s of CPU work on the GPU box's host (a batch-4 train step takes ~0.8 s on 16 threads)
        fps, dt, sample = cpu_reference_fps(args.workload, args.ref_batch, args.size, cpu_steps, 1, threads)
        line["cpu_baseline"] = {"value": fps, "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": "%s, 1 warm-up + %d timed steps, torch %s CPU (oneDNN), %d threads"
                                          % (sample, cpu_steps, torch.__version__, threads)}
    emit(line)
    finish(world)


if __name__ == "__main__":
    main()
