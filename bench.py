#!/usr/bin/env python
"""bench.py — BASELINE.json's metric on BASELINE.json's config, on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c1|c3] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

Metric: voice-seconds rendered per second (= voices x rendered seconds / render time); realtime factor = rendered
seconds / render time.  Workload (N=1): BASELINE.json configs[1] = "C2": 64 voices x (stereo 10 s source -> GainNode with
a-rate automation -> ConvolverNode with a per-voice 2 s stereo IR) -> bus GainNode(1/8) -> destination, 12 s rendered at
48 kHz.  N > 1: weak scaling — 64 voices per GPU, voices sharded by rank, ONE ncclReduce(sum) of the [2, N] float32 bus
to rank 0 per render (the only exchange step of the path), bus gain applied on the root after the reduce.

A "step" is one complete render of the workload.
  value  = device time (CUDA events on the library's own stream, which is the stream the kernels launch on), inputs
           already resident in HBM (sources uploaded, IR spectra prepared, graph flattened).
  e2e    = the same render through the reference-facing API (PlayableAudioBuffer / ConvolverNode.Buffer / Connect /
           OfflineAudioContext.Render) starting from pinned HOST arrays: H2D of sources and IRs, IR preparation, render,
           D2H of the result are all inside the timed region (wall clock between device synchronisations).
  roofline: dominant kernel = the spectral MAC (K6), computed as a fast convolution along block time (a second FFT over the
           partition axis, csrc/fft2.cu).  achieved = the bytes this algorithm has to move per launch (XT, H2, YT once) / K6's
           CUDA-event duration, against the measured HBM peak, with the ncu-measured DRAM traffic of the same launch beside
           it.  `roofline_contract` restates it with SURVEY.md §8d's contract bytes (what the REFERENCE algorithm moves: per
           channel-convolver block 16*P*C + 8*C + 8*B, T = 1) — a fraction >> 1 that measures the algorithmic gain;
           `roofline_fp32` gives the flops K6 issues against the FP32 peak.
  cpu_baseline: the CPU oracle (a C++ restatement of the reference's algorithm; the reference is C#/.NET and cannot run
           here) on a bounded sample of the same workload, 1 thread (the reference renders a context on one thread).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from tests import synth  # noqa: E402

FS = 48000
WORKLOADS = {
    # name: (voices per GPU, source seconds, ir seconds, render seconds, bus gain, builder)
    "c2": dict(voices=64, src_s=10.0, ir_s=2.0, render_s=12.0, bus_gain=1.0 / 8, kind="c2",
               desc="C2: 64 voices x (stereo 10 s noise -> GainNode a-rate automation -> ConvolverNode 2 s stereo IR per voice) -> bus GainNode(1/8), 12 s @ 48 kHz"),
    "c1": dict(voices=1, src_s=10.0, ir_s=1.0, render_s=11.0, bus_gain=1.0, kind="c1",
               desc="C1: 1 voice, stereo 10 s noise -> ConvolverNode 1 s stereo IR, 11 s @ 48 kHz"),
    "c3": dict(voices=128, src_s=10.0, ir_s=2.0, render_s=12.0, bus_gain=1.0 / 32, kind="c3",
               desc="C3 (per-GPU shard): 128 voices x (BiQuad lowpass a-rate sweep -> GainNode -> ConvolverNode 2 s IR) -> bus GainNode(1/32), 12 s @ 48 kHz"),
}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)", float(d.get("sm_max_mhz", 1965.0))
    return 6650.0, "fallback (B200_PROFILING.md)", 1965.0


def ncu_traffic(variant_used, big_segments=0):
    """DRAM bytes per K6 launch (pair) on the C2 workload from the committed ncu --set full capture of the variant in use."""
    name = {3: "k6_fft2_mixed_traffic.json" if big_segments > 0 else "k6_fft2_traffic.json"}.get(variant_used, "mac_traffic.json")
    p = os.path.join(ROOT, "profiles", name)
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f)
    return None


class ClockSampler:
    def __init__(self, index):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index),
                 "--query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                 "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap",
                 "--format=csv,noheader,nounits", "-lms", "10"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t_begin=None, t_end=None):
        """Summary of the samples that arrived inside [t_begin, t_end] (the timed region)."""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ts, r in self.rows:
            if t_begin is not None and not (t_begin <= ts <= t_end + 0.1):
                continue
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_inputs(wl, rank, pinned):
    """Per-voice host arrays (pinned when torch/cuda is available): (src[2], ir[2], gains)."""
    nsrc, nir = int(wl["src_s"] * FS), int(wl["ir_s"] * FS)
    V = wl["voices"]
    alloc = None
    if pinned:
        import torch
        def alloc(rows, n):  # noqa: E306  one page-locked block per buffer, the channels are its rows
            return torch.empty((rows, n), dtype=torch.float32, pin_memory=True).numpy()
    voices = []
    for i in range(V):
        v = rank * V + i
        src, ir = synth.make_voice_inputs(v, nsrc, nir)
        if alloc:
            ps = alloc(len(src), src[0].shape[0])
            pi = alloc(len(ir), ir[0].shape[0])
            for c, a in enumerate(src):
                ps[c, :] = a
            for c, a in enumerate(ir):
                pi[c, :] = a
            src, ir = [ps[c] for c in range(len(src))], [pi[c] for c in range(len(ir))]
        voices.append((src, ir, synth.voice_gains(v)))
    return voices


def build_graph(api, wl, voices, **kw):
    if wl["kind"] == "c1":
        return synth.build_c1(api, FS, voices[0][0], voices[0][1], **kw)
    if wl["kind"] == "c3":
        return synth.build_c3(api, FS, voices, wl["bus_gain"], **kw)
    return synth.build_c2(api, FS, voices, wl["bus_gain"], **kw)


# ----------------------------------------------------------------------------------------------- reference arm / cpu baseline
def cpu_render_sample(wl, n_voices, render_s):
    """Times the CPU oracle on the first n_voices of the workload (same graph, same IR length), one thread."""
    from oracle import ga_oracle as O
    sub = dict(wl)
    sub["voices"] = n_voices
    voices = make_inputs(sub, 0, pinned=False)
    ctx = build_graph(O, sub, voices)
    n = int(render_s * FS)
    t0 = time.perf_counter()
    out = ctx.Render(n)
    dt = time.perf_counter() - t0
    assert np.isfinite(out).all()
    return n_voices * render_s / dt, dt


def run_reference(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    nv = min(2, wl["voices"])
    vals = []
    for i in range(args.warmup + args.steps):
        v, dt = cpu_render_sample(wl, nv, wl["render_s"])
        if i >= args.warmup:
            vals.append((v, dt))
    value = float(np.mean([v for v, _ in vals]))
    ms = float(np.mean([dt for _, dt in vals])) * 1e3
    sample = f"{nv} of {wl['voices']} voices of the workload, full {wl['render_s']} s render, per step"
    line = {
        "impl": "reference", "metric": "voice-seconds rendered/sec", "value": value, "unit": "voice-s/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": {"workload": wl["desc"], "note": "CPU oracle = C++ restatement of the reference algorithm (the reference is C#/.NET 9; no dotnet in this image); one context renders on one thread, as in the reference"},
        "cpu_baseline": {"value": value, "unit": "voice-s/s", "cores": 1, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "voice-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "realtime_factor": value / nv,
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------- our arm
def run_ours(args, wl):
    import torch
    import torch.distributed as dist
    import graphaudio_b200 as G
    from graphaudio_b200 import _native as N
    from graphaudio_b200.api import check

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: graphaudio_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    n = int(wl["render_s"] * FS)
    voices = make_inputs(wl, rank, pinned=True)
    h2d_bytes = sum(a.nbytes for v in voices for a in (v[0] + v[1]))
    d2h_bytes = 2 * n * 4 if rank == 0 else 0
    L = N.lib()

    def build():
        return build_graph(G, wl, voices, device_id=local, tile_blocks=args.tile_blocks, partition=args.partition, mac_variant=args.mac_variant,
                           uniform_segments=args.uniform_segments)

    def comm(ctx):
        if world > 1:
            idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
            if rank == 0:
                buf = (C.c_char * 128)()
                check(L.gac_comm_unique_id(buf))
                idt.copy_(torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8))
            dist.broadcast(idt, 0)
            raw = bytes(idt.cpu().numpy().tobytes())
            check(L.gac_comm_init(ctx._h, raw, rank, world))

    # ---- resident arm: inputs in HBM, graph flattened once
    ctx = build()
    if world > 1:
        ctx.MarkBus(ctx.bus)
    comm(ctx)
    graph = ctx._graph()
    out_pinned = torch.zeros((2, n), dtype=torch.float32, pin_memory=True)  # the result lands in page-locked host memory
    out_host = out_pinned.numpy()
    out_ptrs = (N.fp * 2)(*[out_host[c].ctypes.data_as(N.fp) for c in range(2)])
    d_out = torch.empty((2, n), dtype=torch.float32, device="cuda")
    st = N.gac_stats()

    def step_resident():
        if world > 1:
            check(L.gac_render_sharded(ctx._h, graph, n, 0, out_ptrs, 2))
        else:
            check(L.gac_render_device(ctx._h, graph, 0, n, C.c_void_p(d_out.data_ptr()), 2, 1))
        check(L.gac_get_stats(ctx._h, C.byref(st)))
        return st.as_dict()

    sampler = ClockSampler(local)
    sampler.start()  # nvidia-smi needs a few hundred ms to deliver its first sample: start it before the warm-up
    for _ in range(max(3, args.warmup)):
        step_resident()
    barrier()
    t0 = time.perf_counter()
    stats = [step_resident() for _ in range(args.steps)]
    barrier()
    t1 = time.perf_counter()
    wall = t1 - t0
    clocks = sampler.stop(t0, t1)
    dev_ms = float(sum(s["ms_total"] for s in stats))
    tmax = torch.tensor([dev_ms, wall * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    dev_ms_step = float(tmax[0]) / args.steps
    wall_ms_step = float(tmax[1]) / args.steps
    L.gac_graph_destroy(graph)

    # ---- e2e arm: host arrays -> public API -> host result, everything inside the timed region.
    # Per step: a fresh OfflineAudioContext, PlayableAudioBuffer uploads (H2D from pinned host memory),
    # ConvolverNode.Buffer (IR preparation), Connect, Render (D2H of the result on the root).
    # N > 1: the NCCL communicator of each fresh context is bootstrapped BEFORE its timed region (not part of a render).
    e2e_note = None
    if world > 1:
        e2e_note = "communicator bootstrap outside the timed region; H2D + IR prepare + sharded render + D2H inside"
    lat = []
    # the interpreter's cyclic collector is left ON, but everything allocated so far (torch, numpy, the input arrays) is moved to the
    # permanent generation: otherwise one step in ~50 pays a 40 ms full collection of objects that have nothing to do with the render
    import gc
    gc.collect()
    gc.freeze()
    for i in range(3 + args.steps):
        def fresh():
            return G.OfflineAudioContext(FS, device_id=local, tile_blocks=args.tile_blocks, partition=args.partition,
                                         mac_variant=args.mac_variant, async_upload=not args.sync_upload)
        if world > 1:
            c = fresh()
            comm(c)
        barrier()
        t0 = time.perf_counter()
        if world == 1:
            c = fresh()  # N = 1: creating the context is part of the step
        build_into(G, wl, voices, c)
        if world > 1:
            c.MarkBus(c.bus)
            g = c._graph()
            check(L.gac_render_sharded(c._h, g, n, 0, out_ptrs, 2))
            L.gac_graph_destroy(g)
        else:
            c.Render(out_host, n, 0)
        barrier()
        if i >= 3:
            lat.append(time.perf_counter() - t0)
        c.Dispose()
    tl = torch.tensor([sum(lat) * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tl, op=dist.ReduceOp.MAX)
    e2e_ms = float(tl[0]) / args.steps

    # ---- cpu baseline (rank 0, N = 1 only): bounded sample of the same workload on the host
    cpu = None
    if world == 1 and not args.no_cpu:
        nv = min(args.cpu_voices, wl["voices"])
        v, dt = cpu_render_sample(wl, nv, wl["render_s"])
        cpu = {"value": v, "unit": "voice-s/s", "cores": 1, "kind": "port",
               "sample": f"{nv} of {wl['voices']} voices, full {wl['render_s']} s render ({dt:.1f} s of CPU work); CPU oracle = C++ restatement of the reference algorithm, not the .NET binary"}

    if rank == 0:
        V = wl["voices"] * world
        peak, peak_src, sm_max = measured_peaks()
        s_last = stats[-1]
        mac_ms = float(np.mean([s["ms_mac"] for s in stats]))
        alg_bytes = s_last["algorithmic_bytes"]
        achieved = alg_bytes / (mac_ms * 1e-3) / 1e9
        used = int(s_last["mac_variant_used"])
        big = int(s_last.get("mac_big_segments", 0))
        traffic = ncu_traffic(used, big) if args.workload == "c2" and args.partition == 128 else None
        flops = s_last["mac_flops"]
        k6_name = {1: "k_mac_stream (K6, direct sum, reference op order)", 2: "k_mac_tiled (K6, register-tiled direct sum, FFMA)",
                   4: "k_mac_tiled (K6, register-tiled direct sum, FFMA2)",
                   3: "k_fft2_conv16 (K6 spectral MAC as a fast convolution along block time)"}.get(used, "K6")
        if used == 3 and big > 0:
            k6_name = (f"k_fft2_conv16<2M> + k_fft2_conv16<M> (K6 as a fast convolution along block time: {big} double-length overlap-save "
                       "segment(s) in front, two launches timed together)")
        moved = s_last["mac_bytes_moved"]
        conv_ms = float(np.mean([s["ms_fft_fwd"] + s["ms_mac"] + s["ms_fft_inv"] for s in stats]))
        # bytes the whole convolver (K5 + K6 + K7) has to move once: signal in (+ gain table), XT out/in, H2, YT out/in, signal out
        units = s_last["conv_units"]
        Bp = args.partition
        conv_bytes = units * (4.0 * Bp * 2 + 8.0 * (Bp + 1) * 4) + (moved - units * 16.0 * (Bp + 1) if used == 3 else 0.0)
        sm_mhz = clocks.get("sm_mhz") or sm_max
        fp32_peak_tf = 148 * 128 * 2 * sm_max * 1e6 / 1e12
        line = {
            "metric": "voice-seconds rendered/sec", "value": V * wl["render_s"] / (dev_ms_step * 1e-3), "unit": "voice-s/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": dev_ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["desc"], "voices_per_gpu": wl["voices"], "partition": args.partition, "sample_rate": FS,
                       "frames": n, "parallelism": f"voices sharded x{world}, one ncclReduce of the bus" if world > 1 else "single GPU",
                       "l2": "working set per step (sources 246 MB + spectrograms 1.2 GB) exceeds the 126 MB L2; no flush needed",
                       "timing": "CUDA events on the library's launch stream (gac_get_stats.ms_total), max over ranks"},
            "realtime_factor": wl["render_s"] / (dev_ms_step * 1e-3),
            "samples_per_s": V * 2 * n / (dev_ms_step * 1e-3),
            "wall_ms_per_step": wall_ms_step,
            "e2e": {"value": V * wl["render_s"] / (e2e_ms * 1e-3), "unit": "voice-s/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": h2d_bytes * world, "d2h_bytes_per_step": d2h_bytes, **({"note": e2e_note} if e2e_note else {})},
            "gpu_launches": int(sum(s["kernel_launches"] for s in stats)),
            "clocks": clocks,
            # the dominant kernel against the HBM roofline: bytes the implemented algorithm has to move per launch (XT, H2, YT once;
            # DESIGN.md §4) / its CUDA-event duration; `traffic` = DRAM bytes of the same launch measured by ncu (profiles/)
            "roofline": {"bound": "hbm", "achieved": moved / (mac_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": moved / (mac_ms * 1e-3) / 1e9 / peak,
                         "traffic": traffic["dram_bytes_per_launch"] if traffic else None,
                         "kernel": k6_name, "peak_source": peak_src, "ms_per_launch": mac_ms, "algorithmic_bytes_per_launch": moved,
                         "convolver_K5_K6_K7": {"ms": conv_ms, "bytes": conv_bytes, "achieved": conv_bytes / (conv_ms * 1e-3) / 1e9,
                                                "frac": conv_bytes / (conv_ms * 1e-3) / 1e9 / peak}},
            # SURVEY.md 8(d)'s contract: the bytes the REFERENCE algorithm moves for the same units (FDL + IR walked once per
            # quantum, T = 1, C = B + 1) over K6's time.  A fraction >> 1 is the algorithmic gain (each spectrogram moved once
            # instead of P times), not a measure of kernel quality: `roofline` above is.
            "roofline_contract": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                                  "algorithmic_bytes_per_launch": alg_bytes, "kernel": k6_name},
            "roofline_fp32": {"bound": "fp32", "achieved": flops / (mac_ms * 1e-3) / 1e12, "peak": fp32_peak_tf, "unit": "TFLOP/s",
                              "frac": flops / (mac_ms * 1e-3) / 1e12 / fp32_peak_tf,
                              "flops_per_launch": flops,
                              "peak_source": f"148 SMs x 128 FMA/clk x 2 x {sm_max:.0f} MHz (nominal max clock); median SM clock under load {sm_mhz} MHz"},
            "kernel_ms": {k: float(np.mean([s[k] for s in stats])) for k in
                          ["ms_source", "ms_automation", "ms_biquad", "ms_gain", "ms_fft_fwd", "ms_mac", "ms_fft_inv", "ms_mix", "ms_d2h"]},
        }
        if cpu:
            line["cpu_baseline"] = cpu
        line["e2e"]["async_upload"] = not args.sync_upload
        line["e2e"]["ms_per_step_quartiles"] = [float(x) * 1e3 for x in np.percentile(lat, [0, 25, 50, 75, 100])]  # this rank's steps
        line["e2e"]["slowest_step"] = int(np.argmax(lat))
        print(json.dumps(line))
    ctx.Dispose()
    if world > 1:
        dist.destroy_process_group()


def build_into(api, wl, voices, ctx):
    """Builds the workload's graph inside an existing context (used by the N>1 e2e arm)."""
    class _Shim:
        pass
    shim = _Shim()
    for name in dir(api):
        setattr(shim, name, getattr(api, name))
    shim.OfflineAudioContext = lambda fs, **kw: ctx
    return build_graph(shim, wl, voices)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--voices", type=int, default=0, help="voices per GPU (0: the workload's own count)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--partition", type=int, default=128)
    ap.add_argument("--tile-blocks", dest="tile_blocks", type=int, default=32)
    ap.add_argument("--mac-variant", dest="mac_variant", type=int, default=0,
                    help="K6 algorithm (gac_context_desc.mac_variant): 0 default (second-level FFT), 1 streaming direct sum, 4 register-tiled direct sum")
    ap.add_argument("--cpu-voices", dest="cpu_voices", type=int, default=32)
    ap.add_argument("--no-cpu", dest="no_cpu", action="store_true")
    ap.add_argument("--uniform-segments", dest="uniform_segments", action="store_true",
                    help="resident arm: K6 without the double-length overlap-save segments in front (GAC_FLAG_UNIFORM_SEGMENTS), for A/B runs")
    ap.add_argument("--no-extra", dest="no_extra", action="store_true", help="(accepted for compatibility; there is no extra measurement any more)")
    ap.add_argument("--sync-upload", dest="sync_upload", action="store_true",
                    help="e2e arm: copy every buffer during gac_buffer_create (reference semantics) instead of GAC_FLAG_ASYNC_UPLOAD")
    args = ap.parse_args()
    wl = dict(WORKLOADS[args.workload])
    if args.voices > 0:
        wl["voices"] = args.voices
    if args.impl == "reference":
        run_reference(args, wl)
    else:
        run_ours(args, wl)


if __name__ == "__main__":
    main()
