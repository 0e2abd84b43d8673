#!/usr/bin/env python
"""bench.py -- RoIRotate forward throughput on B200 (BASELINE.json metric: RoIRotate Mfeat-px/s).

  python bench.py [--gpus N] [--steps K] [--warmup W]                  # product arm (default)
  python bench.py --impl reference [--gpus N] [--steps K] [--warmup W]  # reference CPU arm

One STEP = one batch of `--launches-per-step` (default 2010) independent RoIRotate requests in BASELINE.json
configs[1] ("cfg1"): each request is a single 1280x720 image's shared feature map (180x320 at 1/4 scale), 64 random
rotated RoIs, 8x64 pooled output, fp32 -- ONE forward launch through the C ABI per request, so a step is a fixed CUDA
graph of 2010 launches (about 5 ms) and `--steps 20` times about 0.1 s.  `value` = output feature pixels
(N*C*PH*PW, zero tail included) per second, whole job (all ranks), inputs resident in HBM; `roofline` is per launch.
Consecutive requests use different (feature map, RoIs, output) buffer sets out of 67 whose total size is 12x the
126 MB L2, so every launch reads its features from HBM, and are issued round-robin on `--streams` streams the way a
serving pipeline keeps independent images in flight.  Timed with CUDA events on the launching stream, bracketed by
barrier + synchronize; max over ranks.

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

METRIC = "roirotate_fwd_mfeat_px_per_s"
UNIT = "Mfeat-px/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--channels", type=int, default=64, help="64 = the reference's focr map; 256 = FPN map")
    ap.add_argument("--layout", default="nhwc", choices=["nhwc", "nchw"])
    ap.add_argument("--images", type=int, default=1, help="images per step (cfg1 = 1)")
    ap.add_argument("--rois-per-image", type=int, default=64)
    ap.add_argument("--sets", type=int, default=0, help="rotating buffer sets (0 = enough to exceed 3x L2)")
    ap.add_argument("--launches-per-step", type=int, default=0,
                    help="requests (= forward launches) per step; 0 = 30 x the number of buffer sets (2010 for cfg1)")
    ap.add_argument("--pdl", type=int, default=1)
    ap.add_argument("--rois-ready", type=int, default=1,
                    help="RROI_B200_FLAG_ROIS_READY: the RoI rows are uploaded long before the launch (true for this bench)")
    ap.add_argument("--concurrency", type=int, default=-1,
                    help="rroi_b200_opts.concurrency hint; -1 = the number of streams")
    ap.add_argument("--streams", type=int, default=8,
                    help="independent steps (different images) are issued round-robin on this many streams")
    ap.add_argument("--variant", type=int, default=0,
                    help="rroi_b200_opts.variant: 0 = automatic (from grid size and the concurrency hint)")
    ap.add_argument("--no-extras", action="store_true", help="skip e2e / cpu_baseline / variants legs")
    ap.add_argument("--e2e-steps", type=int, default=200)
    ap.add_argument("--e2e-streams", type=int, default=4, help="host-buffer sets / streams of the e2e leg")
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--pipeline-images", type=int, default=32, help="images per GPU per end-to-end step (cfg4: 32)")
    ap.add_argument("--pipeline-steps", type=int, default=4)
    ap.add_argument("--pipeline-lanes", type=int, default=1, help="parallel graph branches the step's micro-batches are dealt to")
    ap.add_argument("--pipeline-micro", type=int, default=32, help="images per micro-batch of the end-to-end step")
    ap.add_argument("--no-pipeline", action="store_true")
    ap.add_argument("--no-train", action="store_true")
    ap.add_argument("--train-images", type=int, default=32, help="batch of the cfg3 training step")
    ap.add_argument("--train-steps", type=int, default=3)
    return ap.parse_args()


# ----------------------------------------------------------------------------------------- helpers

def host_threads():
    """Host cores this process may use; passed explicitly to the oracle (torchrun exports OMP_NUM_THREADS=1)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


class StdoutGuard:
    """Everything that writes to fd 1 while the benchmark runs (NCCL prints its version banner there) goes to
    stderr; the ONE JSON line is written to the real stdout at the end."""

    def __init__(self):
        sys.stdout.flush()
        self.real = os.dup(1)
        os.dup2(2, 1)

    def emit(self, text):
        sys.stdout.flush()
        os.write(self.real, (text + "\n").encode())


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index, period=0.02):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.stop_flag = [], set(), threading.Event()
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:  # pragma: no cover
            self.nv, self.err = None, repr(e)

    def sample(self):
        if self.nv is None:
            return
        nv = self.nv
        try:
            self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
            r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown",
                     0x4: "sw_power_cap", 0x80: "hw_power_brake", 0x2: "applications_clocks_setting"}
            for bit, name in names.items():
                if r & bit:
                    self.reasons.add(name)
        except Exception:  # pragma: no cover
            pass

    def run(self):
        while not self.stop_flag.is_set():
            self.sample()
            time.sleep(self.period)

    def finish(self):
        self.sample()
        self.stop_flag.set()
        self.join(timeout=1.0)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(key):
    """DRAM bytes per launch from the committed ncu --set full capture (profiles/roofline_traffic.json)."""
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            return json.load(f).get(key)
    except Exception:
        return None


# ----------------------------------------------------------------------------------------- workload

class Workload:
    """`sets` independent (features, rois, pooled) buffer sets of the cfg1 shape, resident on the device.
    Every launch goes through the C ABI with per-call options (rroi_b200_opts); nothing is process-global."""

    def __init__(self, args, device, torch, cabi=None):
        import workloads as WL
        self.WL = WL
        self.C, self.H, self.W, self.PH, self.PW, self.scale = args.channels, 180, 320, 8, 64, 0.25
        self.B = args.images
        self.N = args.images * args.rois_per_image
        self.layout = args.layout
        self.bf16 = getattr(args, "dtype", "fp32") == "bf16"         # rroi_b200_forward_bf16 (channels-last only)
        esz = 2 if self.bf16 else 4
        per_set = esz * (self.B * self.C * self.H * self.W + self.N * self.C * self.PH * self.PW)
        # rotate over ~12x the L2 capacity: with a non-LRU replacement policy a cyclic working set of k x L2 can
        # still hit ~1/k of the time, so 3x (the first choice) flattered the kernel by up to 30 %
        self.sets = args.sets if args.sets > 0 else max(4, int(np.ceil(12 * 126e6 / per_set)) + 1)
        self.working_set_mb = per_set * self.sets / 1e6
        fmt = torch.channels_last if self.layout == "nhwc" else torch.contiguous_format
        gen = torch.Generator(device=device).manual_seed(1234)
        self.feats, self.rois, self.rois_np, self.alg_bytes = [], [], [], []
        for s in range(self.sets):
            f = torch.randn(self.B, self.C, self.H, self.W, device=device, generator=gen)
            if self.bf16:
                f = f.to(torch.bfloat16)
            self.feats.append(f.contiguous(memory_format=fmt))
            r = WL.batch_rois(self.B, args.rois_per_image, seed0=s * self.B)
            self.rois_np.append(r)
            self.rois.append(torch.from_numpy(r).to(device))
            alg = WL.algorithmic_bytes_fwd(r, self.C, self.PH, self.PW)
            # bf16: the same element counts at 2 bytes (the 24-byte RoI rows stay fp32)
            self.alg_bytes.append((alg - 24 * self.N) / 2 + 24 * self.N if self.bf16 else alg)
        self.feat_px_per_step = self.N * self.C * self.PH * self.PW
        self.out = [torch.empty((self.N, self.C, self.PH, self.PW), device=device, memory_format=fmt,
                                dtype=torch.bfloat16 if self.bf16 else torch.float32) for _ in range(self.sets)]
        self.backward = False
        self.opts = None                  # ctypes rroi_b200_opts kept alive here; None = the library's defaults
        self.xform = None                 # optional per-set transform tables (rroi_b200_roi_xform)

    def set_opts(self, cabi, **kw):
        self.opts = cabi.opts(**kw) if kw else None
        return self

    def enable_xform(self, torch, lib, stream):
        self.xform = [torch.empty((self.N, 8), device=r.device) for r in self.rois]
        for x, r in zip(self.xform, self.rois):
            assert lib.rroi_b200_roi_xform(r.data_ptr(), x.data_ptr(), self.N, self.PH, self.scale, stream) == 0

    def enable_backward(self, torch):
        """Allocate top_diff / bottom_diff per buffer set and compute the backward's algorithmic bytes:
        4*C*sum V_n (top_diff of valid elements) + 4*B*C*H*W (the map is defined everywhere) + 2*4*C*U (RMW of the U
        distinct pixels that receive gradient) + 24*N.  U comes from this library's own forward centres."""
        from fots.pytorch_b200.rroi_align.functions.rroi_align import forward_raw
        self.gtop = [torch.randn_like(o) for o in self.out]
        self.gbot = [torch.empty_like(f) for f in self.feats]
        self.backward = True
        self.alg_bytes_bwd, self.touched = [], []
        for s in range(self.sets):
            _, ix, iy, _ = forward_raw(self.feats[s], self.rois[s], self.PH, self.PW, self.scale, want_idx=True)
            valid = torch.from_numpy(self.WL.valid_counts(self.rois_np[s], self.PH, self.PW) // self.PH).to(ix.device)
            inside = torch.arange(self.PW, device=ix.device)[None, None, :] < valid[:, None, None]
            b = self.rois[s][:, 0].long()[:, None, None].expand_as(ix)
            keys = []
            for fx, fy in ((torch.floor, torch.floor), (torch.ceil, torch.floor), (torch.ceil, torch.ceil), (torch.floor, torch.ceil)):
                x, y = fx(ix).long(), fy(iy).long()
                ok = inside & (x > 0) & (x < self.W - 1) & (y > 0) & (y < self.H - 1)      # kernel.cu:267-274
                keys.append(((b * self.H + y) * self.W + x)[ok])
            U = int(torch.unique(torch.cat(keys)).numel())
            V = int(self.WL.valid_counts(self.rois_np[s], self.PH, self.PW).sum())
            self.touched.append(U)
            self.alg_bytes_bwd.append(4 * self.C * V + 4 * self.B * self.C * self.H * self.W + 8 * self.C * U + 24 * self.N)

    def launch_bwd(self, s, lib, cabi, stream):
        st = lib.rroi_b200_backward_opt(self.gtop[s].data_ptr(), self.rois[s].data_ptr(), None, None,
                                        self.gbot[s].data_ptr(), self.N, self.B, self.C, self.H, self.W, self.PH,
                                        self.PW, self.scale,
                                        cabi.LAYOUT_NHWC if self.layout == "nhwc" else cabi.LAYOUT_NCHW, 1,
                                        cabi.opts_ref(self.opts), stream)
        if st != 0:
            raise RuntimeError("rroi_b200_backward_opt -> %d" % st)

    def launch(self, s, lib, cabi, stream):
        """One request on buffer set s: exactly one kernel launch through the C ABI (rroi_b200_forward_opt)."""
        if self.backward:
            return self.launch_bwd(s, lib, cabi, stream)
        xf = self.xform[s].data_ptr() if self.xform is not None else None
        if self.bf16:
            st = lib.rroi_b200_forward_bf16_opt(self.feats[s].data_ptr(), self.rois[s].data_ptr(), xf, self.out[s].data_ptr(),
                                                None, None, self.N, self.B, self.C, self.H, self.W, self.PH, self.PW,
                                                self.scale, cabi.opts_ref(self.opts), stream)
            if st != 0:
                raise RuntimeError("rroi_b200_forward_bf16_opt -> %d" % st)
            return
        st = lib.rroi_b200_forward_opt(self.feats[s].data_ptr(), self.rois[s].data_ptr(), xf, self.out[s].data_ptr(),
                                       None, None, self.N, self.B, self.C, self.H, self.W, self.PH, self.PW,
                                       self.scale, cabi.LAYOUT_NHWC if self.layout == "nhwc" else cabi.LAYOUT_NCHW,
                                       cabi.opts_ref(self.opts), stream)
        if st != 0:
            raise RuntimeError("rroi_b200_forward_opt -> %d" % st)


def launches_per_step_for(wl, want=0, target=2000):
    """A multiple of the number of buffer sets (every step then starts on set 0 and one graph serves all steps)."""
    if want > 0:
        return max(wl.sets, (want // wl.sets) * wl.sets)
    return wl.sets * max(1, int(round(target / wl.sets)))


def timed_steps(wl, steps, warmup, per_step, torch, lib, cabi, barrier, nstreams=1, launch=None):
    """One step = ONE CUDA graph of `per_step` launches (independent requests on rotating buffer sets), captured on
    `nstreams` parallel branches the way a serving pipeline keeps several images in flight.  W untimed warm-up steps
    (>= 3), then exactly K replays timed with CUDA events on the launching stream; returns elapsed ms."""
    launch = launch or (lambda s, st: wl.launch(s, lib, cabi, st))
    stream = torch.cuda.Stream()
    side = [torch.cuda.Stream() for _ in range(max(nstreams - 1, 0))]
    with torch.cuda.stream(stream):
        for i in range(3):   # lazy module load etc. outside any capture
            launch(i % wl.sets, stream.cuda_stream)
    stream.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=stream):
        cap = torch.cuda.current_stream()
        lanes = [cap] + side
        fork = torch.cuda.Event()
        fork.record(cap)
        for sd in side:
            sd.wait_event(fork)
        for i in range(per_step):
            launch(i % wl.sets, lanes[i % len(lanes)].cuda_stream)
        for sd in side:
            join = torch.cuda.Event()
            join.record(sd)
            cap.wait_event(join)
    # clock ramp (untimed, before the W warm-up steps): ~0.3 s of the same work
    t_end = time.time() + 0.3
    with torch.cuda.stream(stream):
        while time.time() < t_end:
            g.replay()
            stream.synchronize()
        for _ in range(max(warmup, 3)):
            g.replay()
    stream.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    torch.cuda.synchronize()
    with torch.cuda.stream(stream):
        e0.record(stream)
        for _ in range(steps):
            g.replay()
        e1.record(stream)
    torch.cuda.synchronize()
    barrier()
    return e0.elapsed_time(e1)


def verify_outputs(wl, torch, cabi):
    """After the timed region: every buffer set's pooled output (written by the graph replays, on whatever stream)
    must equal a fresh single launch with default options, bit for bit."""
    from fots.pytorch_b200.rroi_align.functions.rroi_align import forward_raw
    ok = True
    for s in range(0, wl.sets, max(1, wl.sets // 8)):
        want, _, _, _ = forward_raw(wl.feats[s], wl.rois[s], wl.PH, wl.PW, wl.scale, want_idx=False)
        ok = ok and bool(torch.equal(want, wl.out[s]))
    return ok


def e2e_leg(args, wl, torch, device, steps, barrier, nbuf=None):
    """Same metric end to end through the public module API with HOST buffers: per step a pinned-host ->
    device copy of the step's features + RoIs, _RRoiAlign.forward, and a device -> pinned-host read of the
    pooled result.  Two streams alternate so the H2D of step i+1 overlaps the D2H of step i.  Every rank runs it."""
    from fots.pytorch_b200 import _RRoiAlign
    fmt = torch.channels_last if wl.layout == "nhwc" else torch.contiguous_format
    nbuf = nbuf or getattr(args, "e2e_streams", 2)
    mod = _RRoiAlign(wl.PH, wl.PW, wl.scale)
    h_feat = [wl.feats[i % wl.sets].cpu().contiguous(memory_format=fmt).pin_memory() for i in range(nbuf)]
    h_rois = [wl.rois[i % wl.sets].cpu().pin_memory() for i in range(nbuf)]
    d_feat = [torch.empty_like(wl.feats[0]) for _ in range(nbuf)]
    d_rois = [torch.empty_like(wl.rois[0]) for _ in range(nbuf)]
    h_out = [torch.empty((wl.N, wl.C, wl.PH, wl.PW), memory_format=fmt).pin_memory() for _ in range(nbuf)]
    streams = [torch.cuda.Stream() for _ in range(nbuf)]

    def step(i):
        b = i % nbuf
        with torch.cuda.stream(streams[b]):
            d_feat[b].copy_(h_feat[b], non_blocking=True)
            d_rois[b].copy_(h_rois[b], non_blocking=True)
            y = mod(d_feat[b], d_rois[b])
            h_out[b].copy_(y, non_blocking=True)

    for i in range(6):
        step(i)
    torch.cuda.synchronize()
    barrier()
    t0 = time.perf_counter()
    for i in range(steps):
        step(i)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    barrier()
    h2d = h_feat[0].numel() * 4 + h_rois[0].numel() * 4
    d2h = h_out[0].numel() * 4
    return dt, h2d, d2h


def ref_gpu_kernel_leg(wl, torch, lib, cabi, steps=30):
    """Baseline leg (like cpu_baseline): the reference's own CUDA kernel (rroi_align_kernel.cu compiled unmodified for
    sm_100a into oracle/_ref) with the three zero-fills its Python wrapper issues
    (functions/rroi_align.py:17-20), graph-replayed over the same rotating NCHW buffer sets.  Checker-side code: it is
    only timed here, never part of the product path."""
    import ctypes
    path = os.path.join(ROOT, "oracle", "_ref", "libref_rroi_sm100a.so")
    if not os.path.exists(path):
        return {"unavailable": "oracle/_ref/libref_rroi_sm100a.so not built (needs /root/reference at build time)"}
    R = ctypes.CDLL(path)
    i, f, vp = ctypes.c_int, ctypes.c_float, ctypes.c_void_p
    R.RROIAlignForwardLaucher.restype = i
    R.RROIAlignForwardLaucher.argtypes = [vp, f, i, i, i, i, i, i, vp, vp, vp, vp, vp]
    nsets = min(wl.sets, 24)                      # 24 x (14.7 + 3 x 8.4) MB = 0.96 GB = 7.6 x L2
    feats = [wl.feats[s].contiguous() for s in range(nsets)]
    outs = [[torch.empty((wl.N, wl.C, wl.PH, wl.PW), device=feats[0].device) for _ in range(3)] for _ in range(nsets)]

    class Shim:
        sets = nsets

    def launch(s, st):
        with torch.cuda.stream(torch.cuda.ExternalStream(st)):
            for t in outs[s]:
                t.zero_()
        rc = R.RROIAlignForwardLaucher(feats[s].data_ptr(), wl.scale, wl.N, wl.H, wl.W, wl.C, wl.PH, wl.PW,
                                       wl.rois[s].data_ptr(), outs[s][0].data_ptr(), outs[s][1].data_ptr(),
                                       outs[s][2].data_ptr(), st)
        assert rc == 1

    per_step = nsets * 10
    ms = timed_steps(Shim, steps, 3, per_step, torch, lib, cabi, lambda: None, 1, launch=launch)
    us = ms / (steps * per_step) * 1e3
    ms8 = timed_steps(Shim, steps, 3, per_step, torch, lib, cabi, lambda: None, 8, launch=launch)
    us8 = ms8 / (steps * per_step) * 1e3
    # product, same layout (NCHW), same protocol
    same = {"sets": nsets}
    return {"us_per_call_1stream": us, "us_per_call_8streams": us8, "mfeat_px_per_s_1stream": wl.feat_px_per_step / us,
            "mfeat_px_per_s_8streams": wl.feat_px_per_step / us8, "layout": "nchw", "channels": wl.C, "rois": wl.N,
            "includes": "3 zero-fill kernels + RROIAlignForward (atomicAdd into top_data / con_idx_x / con_idx_y)",
            "protocol": "CUDA graph of %d calls over %d rotating buffer sets, %d replays" % (per_step, nsets, steps), **same}


def variants_leg(args, torch, device, lib, cabi, peak):
    """Same harness on neighbouring configurations (not the headline): one stream, the reference NCHW layout,
    the 256-channel FPN map, cfg4's per-GPU batch, and the backward.  Each entry: us per launch, Mfeat-px/s, roofline
    fraction (forward: SURVEY 8d forward bytes; backward: top_diff + whole map + RMW of the touched pixels)."""
    import types
    out = {}
    S = args.streams
    grid = [
        # the library's defaults (opts = NULL), one launch at a time on one stream: the latency-bound case
        ("serial_1stream", dict(images=1, streams=1)),
        ("serial_1stream_rois_ready", dict(images=1, streams=1, opts=dict(rois_ready=True, concurrency=1))),
        # 8 streams with the library's defaults (no concurrency hint, no flag)
        ("default_opts_%dstreams" % S, dict(images=1, streams=S)),
        ("nchw_reference_layout", dict(layout="nchw", images=1, streams=S, opts=dict(concurrency=S, rois_ready=True))),
        ("fpn_c256", dict(channels=256, images=1, streams=S, opts=dict(concurrency=S, rois_ready=True))),
        ("cfg4_per_gpu_32img_2048rois", dict(images=32, streams=1)),
        ("cfg4_per_gpu_nchw", dict(layout="nchw", images=32, streams=1)),
        # the inference pipeline's variant: bf16 map in, bf16 pooled out (half the bytes, same arithmetic)
        ("bf16_io", dict(images=1, streams=S, dtype="bf16", opts=dict(concurrency=S, rois_ready=True))),
        ("bf16_io_cfg4_per_gpu", dict(images=32, streams=1, dtype="bf16")),
        # backward (rroi_b200_backward_opt, zero_fill = 1: the gradient map is defined everywhere), cfg3/cfg4's per-GPU batch
        ("backward_nhwc_cfg4", dict(images=32, streams=1, backward=True)),
        ("backward_nchw_cfg4", dict(layout="nchw", images=32, streams=1, backward=True)),
        ("backward_nhwc_cfg1", dict(images=1, streams=1, backward=True)),
    ]
    for name, kw in grid:
        a = types.SimpleNamespace(channels=kw.get("channels", args.channels), layout=kw.get("layout", args.layout),
                                  images=kw["images"], rois_per_image=args.rois_per_image, sets=0, dtype=kw.get("dtype", "fp32"))
        w = Workload(a, device, torch)
        w.set_opts(cabi, **kw.get("opts", {}))
        if kw.get("xform"):
            w.enable_xform(torch, lib, torch.cuda.current_stream().cuda_stream)
        if kw.get("backward"):
            w.enable_backward(torch)
        torch.cuda.synchronize()
        per_step = launches_per_step_for(w, target=max(8, 2000 // (kw["images"] * (4 if kw.get("backward") else 1))))
        steps = 10
        ms = timed_steps(w, steps, 3, per_step, torch, lib, cabi, lambda: None, kw["streams"])
        us = ms / (steps * per_step) * 1e3
        alg = float(np.mean(w.alg_bytes_bwd if kw.get("backward") else w.alg_bytes))
        out[name] = {"us_per_launch": us, "mfeat_px_per_s": w.feat_px_per_step / us, "alg_mb": alg / 1e6,
                     "achieved_gbs": alg / us / 1e3, "frac": alg / us / 1e3 / peak, "streams": kw["streams"],
                     "layout": a.layout, "channels": a.channels, "rois": w.N, "dtype": a.dtype,
                     "opts": kw.get("opts", {}), "launches_timed": steps * per_step}
        if kw.get("backward"):
            out[name]["touched_pixels"] = float(np.mean(w.touched))
            no_rmw = alg - 8.0 * a.channels * float(np.mean(w.touched))         # without the 2*4*C*U read-modify-write term
            out[name]["frac_without_rmw_term"] = no_rmw / us / 1e3 / peak
        del w
        torch.cuda.empty_cache()
    return out


def train_leg(args, torch, device):
    """BASELINE.json configs[3]: one full training step at batch 32, 1280x720 synthetic images, random-init FOTSNet --
    bf16 autocast backbone / recogniser, fp32 RoIRotate forward AND backward kernels (through autograd), CTC loss
    (torch's ctc_loss: warp-ctc is absent, parity unpinned), detection losses, Adam update.  ms per step, images/s."""
    from fots.pytorch_b200.pipeline import FOTSNet
    from fots.pytorch_b200.pipeline.train import TrainStep, synthetic_targets
    torch.manual_seed(0)
    B = args.train_images
    net = FOTSNet(attention=True, nclass=89).to(device).to(memory_format=torch.channels_last)
    step = TrainStep(net, lr=1e-4)
    gen = torch.Generator(device=device).manual_seed(7)
    images = torch.randn(B, 3, 720, 1280, device=device, generator=gen)
    tgt = synthetic_targets(B, 64, 720, 1280, 89, device, seed=0)
    losses = [step(images, tgt)["total"]]                      # warm-up (cuDNN autotune, allocator)
    torch.cuda.synchronize()
    bad = [n for n, p in net.named_parameters() if p.grad is not None and not bool(torch.isfinite(p.grad).all())]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.train_steps):
        losses.append(step(images, tgt)["total"])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.train_steps
    out = {"ms_per_step": ms, "images_per_s": B / (ms * 1e-3), "batch": B, "rois_per_image": 64, "steps": args.train_steps,
           "losses": [float(x) for x in losses], "nonfinite_grads_after_first_step": bad[:8], "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9,
           "dtype": "bf16 autocast networks (cuDNN/cuBLAS under autograd), fp32 RoIRotate forward + backward kernels "
                    "(rroi_b200_forward_opt / rroi_b200_backward_opt), fp32 master weights, Adam",
           "loss": "dense EAST-style detection losses + F.ctc_loss(sum)/N (warp-ctc absent: parity unpinned)"}
    del net, step, images, tgt
    torch.cuda.empty_cache()
    return out


def pipeline_leg(args, torch, device, dist, world, rank):
    """End-to-end images/s (BASELINE.json configs[2]/[4]): random-init FOTSNet in bf16 channels-last, 1280x720
    synthetic images, 64 planted boxes per image, backbone + heads -> RoI rows -> RoIRotate (bf16 in/out) -> forward_ocr ->
    greedy CTC decode, image-sharded (32 images per GPU per step, one 32-image micro-batch, the rank-local part replayed
    from one CUDA graph) with ONE all_gather of the per-image records per step.  Images start on the device; timed with CUDA events, max over ranks."""
    from fots.pytorch_b200.pipeline import FOTSNet, FOTSPipeline
    from fots.pytorch_b200.pipeline.infer import planted_quads
    from fots.pytorch_b200.pipeline.shard import all_gather_records
    torch.manual_seed(0)
    net = FOTSNet(attention=True, nclass=89).to_b200(device, inference=True)
    pipe = FOTSPipeline(net, 8, 64, 0.25, amp_dtype=torch.bfloat16)
    per_gpu, micro = args.pipeline_images, 8                 # micro: the step with detection (host merge per 8-image micro-batch)
    micro_plain = args.pipeline_micro if per_gpu % args.pipeline_micro == 0 else micro
    batch = per_gpu * world
    gen = torch.Generator(device=device).manual_seed(100 + rank)
    # raw uint8 images [b, 720, 1280, 3] as cv2.imread returns them (viewed as channels-last [b, 3, 720, 1280]); the
    # reference's x / 128 - 1 (test.py:80-83) is applied inside the stem kernel, so the host uploads a quarter of the bytes
    images = torch.randint(0, 256, (per_gpu, 720, 1280, 3), device=device, generator=gen, dtype=torch.uint8).permute(0, 3, 1, 2)
    quads = torch.from_numpy(planted_quads(per_gpu, 64, seed0=rank * per_gpu)).to(device)

    # one CUDA graph for the rank-local part of the step; its micro-batches are independent and run as `lanes` parallel
    # branches of the graph (the latency-bound launches of one fill the bubbles of the other).  Measured on B200
    # (tools/step_time.py, 32 images, micro-batch x lanes): 32 x 1 13.74 ms, 16 x 2 13.98, 8 x 4 14.35, 8 x 1 15.7, 4 x 8 15.36
    local = pipe.capture(images, quads, micro_plain, lanes=args.pipeline_lanes)

    def step():
        return all_gather_records(local(), batch)

    def timed(fn, n):
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier(device_ids=[device.index])
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            out = fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device=device, dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) / n, out

    for _ in range(2):
        step()
    ms, out = timed(step, args.pipeline_steps)
    assert out.shape[0] == batch

    # the same step fed from pinned HOST images: H2D of the next step's images on a copy stream overlaps this step's
    # compute; the records come back to pinned host memory every step
    host_imgs = [torch.empty((per_gpu, 720, 1280, 3), dtype=torch.uint8).pin_memory().permute(0, 3, 1, 2) for _ in range(2)]
    for hbuf in host_imgs:
        hbuf.copy_(images)
    stage = [torch.empty_like(images) for _ in range(2)]
    host_rec = torch.empty(out.shape, dtype=out.dtype).pin_memory()
    copy_stream = torch.cuda.Stream()
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]
    state = {"i": 0}

    def prefetch(k):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[k])
            stage[k].copy_(host_imgs[k], non_blocking=True)
            ready[k].record(copy_stream)

    for k in range(2):
        consumed[k].record()
    prefetch(0)

    def step_host():
        k = state["i"] % 2
        prefetch(1 - k)
        cur = torch.cuda.current_stream()
        cur.wait_event(ready[k])
        images.copy_(stage[k])
        consumed[k].record(cur)
        rec = all_gather_records(local(), batch)
        host_rec.copy_(rec, non_blocking=True)
        state["i"] += 1
        return rec

    for _ in range(2):
        step_host()
    ms_host, _ = timed(step_host, args.pipeline_steps)
    # ---- the same step WITH the detector post-processing inside (SURVEY 8d protocol): the head outputs are overwritten
    # by seeded planted detection maps (64 boxes per image), GPU decode + host merge on this rank's thread pool,
    # merged boxes -> RoIs (padded / cut to 64 per image from the planted set)
    det_info = None
    try:
        import workloads as WL
        from fots.pytorch_b200.pipeline.detect import host_threads as det_threads
        q_np = planted_quads(micro, 64, seed0=rank * per_gpu)
        maps8 = WL.planted_maps_from_quads(q_np, 180, 320, seed0=rank * per_gpu)
        maps = [torch.from_numpy(np.concatenate([m] * (per_gpu // micro), 0)).to(device) for m in maps8]
        local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
        threads = max(1, det_threads() // max(local_world, 1))
        det = pipe.capture_with_detection(images, quads, override_maps=maps, micro=micro, threads=threads)
        found = {}

        def step_det():
            rec, f = det()
            found["n"] = f
            return all_gather_records(rec, batch)

        for _ in range(2):
            step_det()
        ms_det, _ = timed(step_det, args.pipeline_steps)
        det_info = {"images_per_s": batch / (ms_det * 1e-3), "ms_per_step": ms_det, "host_threads_per_rank": threads,
                    "boxes_found_per_image_mean": float(np.mean(found["n"])),
                    "boxes": "head outputs overwritten by planted detection maps (64 boxes per image, SURVEY 8d), then "
                             "fots_b200_decode_candidates (GPU) + fots_b200_merge_candidates_host_batch (one host worker per "
                             "8-image micro-batch; all head graphs enqueued up front, the recogniser graphs on a second stream as "
                             "their boxes arrive); images device-resident"}
        del det, maps
    except Exception as e:
        det_info = {"error": repr(e)}
    cpu = pipeline_cpu_baseline(torch) if (rank == 0 and world == 1) else None      # N=1 only, like cpu_baseline
    return {"images_per_s": batch / (ms_host * 1e-3), "ms_per_step": ms_host, "cpu_baseline": cpu,
            "images_per_s_device_resident": batch / (ms * 1e-3), "ms_per_step_device_resident": ms,
            "images_per_step": batch, "images_per_gpu": per_gpu, "rois_per_image": 64, "micro_batch": micro_plain,
            "graph_branches": args.pipeline_lanes,
            "h2d_bytes_per_step": images.numel() * images.element_size() * world, "d2h_bytes_per_step": out.numel() * 4,
            "dtype": "uint8 images in (normalised on load by the stem kernel); bf16 activations, fp32 accumulation: every convolution on "
                     "this repo's kernels (tcgen05 implicit GEMM, mma.sync stem / heads, depthwise), fused InstanceNorm / top-down "
                     "merge / max-pool kernels, bf16-in/bf16-out RoIRotate (fp32 arithmetic, == bf16(reference(float(features))) "
                     "bit for bit; NOT within 1e-4 of the fp32 pipeline by construction: one bf16 rounding per activation)",
            "collective": "one all_gather of int32 [images, 64, 74] records per step" if world > 1 else "none (1 rank)",
            "boxes": "planted (seeded) boxes are an INPUT of this number; `with_detection` has the decode + merge inside the step",
            "with_detection": det_info,
            "flop_per_image": 221.7e9}


def conv_leg(torch, device):
    """Tensor-pipe evidence for the consumer/feeder convolutions: the hand-written tcgen05 implicit-GEMM
    kernel (fots_b200_conv2d_nhwc_bf16, leaky-ReLU fused) on the recogniser's shapes at the 8-image micro-batch of the
    pipeline (512 RoIs), timed with CUDA events over 20 launches after 3 warm-ups, against the measured dense bf16 peak."""
    from fots.pytorch_b200.pipeline import conv as TC
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peak, src = float(json.load(f)["bf16_tflops"]), "measured (MEASURED_PEAKS.json bf16_tflops, burst)"
    except Exception:
        peak, src = 1590.0, "fallback (B200_PROFILING.md 1.59 PFLOP/s)"
    out = {"peak_tflops": peak, "peak_source": src, "bound": "tensor"}
    shapes = [("conv8_conv9_256to256_cta_pairs", 512, 4, 64, 256, 256), ("conv7_128to256_cta_pairs", 512, 4, 64, 128, 256),
              ("conv6_128to128", 512, 8, 64, 128, 128), ("layer0_1_64to64_360x640", 8, 360, 640, 64, 64)]
    with torch.no_grad():
        for name, N, H, W, cin, cout in shapes:
            x = torch.randn(N, cin, H, W, device=device).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
            w = (torch.randn(cout, cin, 3, 3, device=device) / (cin * 9) ** 0.5).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
            for _ in range(3):
                TC.conv2d(x, w, None, (1, 1), 0.01)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(20):
                TC.conv2d(x, w, None, (1, 1), 0.01)
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) / 20 * 1e3
            tf = 2.0 * N * H * W * cout * cin * 9 / us / 1e6
            out[name] = {"us_per_launch": us, "tflops": tf, "frac": tf / peak}
            del x, w
    torch.cuda.empty_cache()
    return out


def pipeline_cpu_baseline(torch):
    """The reference's CPU path for the end-to-end step, on this box's host cores: the same architecture in fp32 on
    torch CPU (fots.pytorch_b200.pipeline.nets is output-identical to tools/models.py with shared weights) + the
    oracle's CPU RoIRotate (the reference has none) + forward_ocr + arg-max, one 1280x720 image with 64 RoIs at a time."""
    from oracle import rroi_oracle as O
    from fots.pytorch_b200.pipeline import FOTSNet
    from fots.pytorch_b200.pipeline.infer import planted_quads
    threads = host_threads()
    old = torch.get_num_threads()
    torch.set_num_threads(threads)
    try:
        torch.manual_seed(0)
        net = FOTSNet(attention=True, nclass=89).eval()
        img = torch.randn(1, 3, 720, 1280)
        q = planted_quads(1, 64)[0]
        rois = np.zeros((64, 6), np.float32)
        for i in range(64):
            p4 = q[i, :8].reshape(4, 2)
            dw, dh = p4[2] - p4[1], p4[1] - p4[0]
            rois[i] = [0, int(p4[:, 0].mean()), int(p4[:, 1].mean()), np.hypot(*dh), np.hypot(*dw),
                       -np.degrees(np.arctan2(dw[1], dw[0]))]

        def one():
            with torch.no_grad():
                _, _, _, feats = net(img)
                pooled, _, _ = O.forward(feats[1].numpy(), rois, 8, 64, 0.25, threads=threads)
                return net.forward_ocr(torch.from_numpy(pooled)).argmax(1)

        one()
        t0, n = time.perf_counter(), 0
        while n < 3:
            one()
            n += 1
        dt = (time.perf_counter() - t0) / n
    finally:
        torch.set_num_threads(old)
    return {"images_per_s": 1.0 / dt, "cores": threads, "kind": "port",
            "sample": "%d images, fp32 torch-CPU backbone + heads + forward_ocr, oracle RoIRotate, %d threads" % (n, threads)}


def cpu_baseline_leg(wl, seconds):
    """The CPU oracle (a port: the reference has no CPU RoIRotate) on the same cfg1 step, all host threads."""
    from oracle import rroi_oracle as O
    feats = wl.feats[0].cpu().contiguous().numpy()
    rois = wl.rois_np[0]
    threads = host_threads()
    O.forward(feats, rois, wl.PH, wl.PW, wl.scale, threads=threads)
    t0, reps = time.perf_counter(), 0
    while True:
        O.forward(feats, rois, wl.PH, wl.PW, wl.scale, threads=threads)
        reps += 1
        dt = time.perf_counter() - t0
        if dt > seconds or reps >= 2000:
            break
    return {"value": wl.feat_px_per_step * reps / dt / 1e6, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": "%d full cfg1 forward passes (%d RoIs, C=%d, NCHW) in %.1f s, OpenMP over (RoI,channel) planes"
                      % (reps, wl.N, wl.C, dt)}


# ----------------------------------------------------------------------------------------- arms

def run_reference(args):
    """--impl reference: the reference's CPU path for this metric.  The reference ships no CPU RoIRotate
    (rroi_align/functions/rroi_align.py:22-25 is dead code), so this is the oracle port of
    rroi_align_kernel.cu:28-162 with all host threads.  Each step is a bounded sample of the cfg1 step."""
    rank, world, _ = dist_env()
    if rank != 0:
        return
    out = StdoutGuard()
    import workloads as WL
    from oracle import rroi_oracle as O
    C, PH, PW, scale = args.channels, 8, 64, 0.25
    steps = args.steps if args.steps is not None else 50
    warmup = args.warmup if args.warmup is not None else 3
    feats = WL.features(0, args.images, C, 180, 320)
    rois = WL.batch_rois(args.images, args.rois_per_image)
    threads = host_threads()
    O.forward(feats, rois, PH, PW, scale, threads=threads)           # page in, spin up the OpenMP team
    t0 = time.perf_counter()
    O.forward(feats, rois, PH, PW, scale, threads=threads)
    t_full = time.perf_counter() - t0
    # bounded sample: the first n RoIs x first c channels of the step, sized so K+W steps fit the budget
    budget = 60.0
    frac = min(1.0, budget / max((steps + warmup) * t_full, 1e-9))
    n = max(1, min(len(rois), int(round(len(rois) * frac))))
    c = C if n > 1 or frac * len(rois) >= 1 else max(1, int(C * frac * len(rois)))
    sample, fsample = rois[:n], np.ascontiguousarray(feats[:, :c])
    for _ in range(warmup):
        O.forward(fsample, sample, PH, PW, scale, threads=threads)
    t0 = time.perf_counter()
    for _ in range(steps):
        O.forward(fsample, sample, PH, PW, scale, threads=threads)
    dt = time.perf_counter() - t0
    px = n * c * PH * PW
    value = px * steps / dt / 1e6
    desc = "first %d of %d RoIs x first %d of %d channels of the cfg1 step per step (NCHW fp32), %d OpenMP threads" % (
        n, len(rois), c, C, threads)
    out.emit(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": dt / steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "cfg1: 1x%dx180x320 fp32 map, %d rotated RoIs, 8x64 pooled, forward" % (C, len(rois)),
                   "sample": desc, "note": "reference has no CPU RoIRotate; oracle port of rroi_align_kernel.cu:28-162"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": desc},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def run_b200(args):
    import torch
    rank, world, local = dist_env()
    out = StdoutGuard()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product has no CPU path)")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=device)
        barrier = lambda: dist.barrier(device_ids=[local])
    else:
        dist = None
        barrier = lambda: None
    from fots.pytorch_b200 import _cabi
    lib = _cabi.lib()

    steps = args.steps if args.steps is not None else 50
    warmup = args.warmup if args.warmup is not None else 5
    wl = Workload(args, device, torch)
    conc = args.streams if args.concurrency < 0 else args.concurrency
    head_opts = dict(pdl=bool(args.pdl), rois_ready=bool(args.rois_ready), concurrency=conc, variant=max(args.variant, 0))
    wl.set_opts(_cabi, **head_opts)
    per_step = launches_per_step_for(wl, args.launches_per_step)
    sampler = ClockSampler(local)
    sampler.start()
    ms = timed_steps(wl, steps, warmup, per_step, torch, lib, _cabi, barrier, args.streams)
    clocks = sampler.finish()
    verified = verify_outputs(wl, torch, _cabi)
    t = torch.tensor([ms], device=device, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    launches = steps * per_step
    value = wl.feat_px_per_step * launches * world / (ms_max * 1e-3) / 1e6

    peak, peak_src = measured_peak_gbs()
    alg = float(np.mean(wl.alg_bytes))
    launch_us = ms / launches * 1e3
    achieved = alg / (launch_us * 1e-6) / 1e9
    kname = "rroi_fwd_nhwc_packed_kernel<%d,256,2> (auto: 256-bin tiles under the concurrency hint)" % wl.C \
        if wl.layout == "nhwc" and wl.C in (32, 64, 128, 256) and args.variant <= 0 else "rroi_fwd_%s (variant %d)" % (wl.layout, args.variant)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": ms_max / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "cfg1 (BASELINE.json configs[1]) x %d independent requests per step: each request = one %dx%dx180x320 "
                               "fp32 feature map, %d random rotated RoIs, 8x64 pooled output, ONE RoIRotate forward launch" %
                               (per_step, wl.B, wl.C, wl.N),
                   "layout": "channels_last (NHWC in HBM)" if wl.layout == "nhwc" else "NCHW (reference layout)",
                   "launches_per_step": per_step, "images_per_step": wl.B * per_step, "rois_per_launch": wl.N, "channels": wl.C,
                   "l2": "inputs larger than L2: %d rotating buffer sets, %.0f MB working set vs 126 MB L2" % (wl.sets, wl.working_set_mb),
                   "launch": "one CUDA graph of %d launches per step, %d stream(s)" % (per_step, args.streams),
                   "opts": "rroi_b200_forward_opt with per-call rroi_b200_opts %r (no process-global tuning)" % head_opts,
                   "parallelism": "image-sharded, %d rank(s), no data-path collective in RoIRotate" % world},
        "gpu_launches": launches,
        "verified": verified,
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": ncu_traffic("%s_c%d" % (wl.layout, wl.C)),
                     "algorithmic_bytes_per_launch": alg, "avg_launch_us": launch_us, "peak_source": peak_src},
    }
    if not args.no_extras:
        # end to end through the public module with host buffers: EVERY rank runs it, value = whole-job aggregate
        dt, h2d, d2h = e2e_leg(args, wl, torch, device, args.e2e_steps, barrier)
        te = torch.tensor([dt], device=device, dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        line["e2e"] = {"value": wl.feat_px_per_step * args.e2e_steps * world / float(te.item()) / 1e6, "unit": UNIT,
                       "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world, "steps": args.e2e_steps,
                       "ranks": world,
                       "api": "fots.pytorch_b200._RRoiAlign(8,64,0.25)(features, rois) with pinned host buffers, 2 streams, on every rank"}
    if rank == 0 and not args.no_extras:
        line["variants"] = variants_leg(args, torch, device, lib, _cabi, peak)
        try:
            line["ref_gpu_kernel"] = ref_gpu_kernel_leg(wl, torch, lib, _cabi)
            nchw = line["variants"].get("nchw_reference_layout")
            if nchw and "us_per_call_8streams" in line["ref_gpu_kernel"]:
                line["ref_gpu_kernel"]["speedup_same_layout_8streams"] = line["ref_gpu_kernel"]["us_per_call_8streams"] / nchw["us_per_launch"]
                line["ref_gpu_kernel"]["speedup_headline_layout"] = line["ref_gpu_kernel"]["us_per_call_8streams"] / launch_us
        except Exception as e:
            line["ref_gpu_kernel"] = {"error": repr(e)}
        if world == 1:
            line["cpu_baseline"] = cpu_baseline_leg(wl, args.cpu_seconds)
        try:
            line["conv_tc"] = conv_leg(torch, device)
        except Exception as e:   # secondary evidence: never take the headline line down
            line["conv_tc"] = {"error": repr(e)}
    del wl
    torch.cuda.empty_cache()
    if not args.no_pipeline and not args.no_extras:
        try:
            line["pipeline"] = pipeline_leg(args, torch, device, dist, world, rank)
        except Exception as e:   # the headline line must survive a failure of the secondary leg
            line["pipeline"] = {"error": repr(e)}
    if rank == 0 and world == 1 and not args.no_train and not args.no_extras:
        try:
            line.setdefault("pipeline", {})["train_step"] = train_leg(args, torch, device)
        except Exception as e:
            line.setdefault("pipeline", {})["train_step"] = {"error": repr(e)}
    if dist is not None:
        dist.barrier(device_ids=[local])
        dist.destroy_process_group()
    if rank == 0:
        out.emit(json.dumps(line))


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
