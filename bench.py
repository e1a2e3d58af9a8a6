#!/usr/bin/env python
"""bench.py — BASELINE.json's metric on BASELINE.json's north-star config, on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3|c2|c1] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

Metric: voice-seconds rendered per second (= voices x rendered seconds / render time); realtime factor = rendered
seconds / render time.

Workload (default `c3` = BASELINE.json configs[2], the north-star target): 128 voices PER GPU, each
AudioBufferSourceNode(stereo 10 s) -> BiQuadFilterNode(lowpass, a-rate cutoff sweep) -> GainNode(a-rate automation) ->
ConvolverNode(per-voice 2 s stereo IR) -> bus GainNode(1/32) -> destination, 12 s rendered at 48 kHz.  The voices are independent
until the bus, so the path shards by voice with no data-path collective but ONE ncclReduce(sum) of the [2, frames] float32 bus
to rank 0 per render, the bus gain on the root behind it: WEAK scaling, per-GPU work fixed — at N = 8 the job is BASELINE
configs[2] exactly as written (1024 voices on 8 GPUs), at N = 1 it is one GPU's share of it.  `--voices 1024` renders the whole
config on one GPU (it fits); `--workload c2` is BASELINE configs[1] (64 voices per GPU).

A "step" is one complete render of the workload.
  value  = device time of gac_render_sharded at EVERY N (N = 1 included: same entry point, same D2H of the 4.6 MB result into
           page-locked host memory inside the timed region), CUDA events on the library's own stream (the stream the kernels
           launch on), max over ranks; inputs already resident in HBM (sources uploaded, IR spectra prepared, graph flattened).
  e2e    = the same render through the reference-facing API (PlayableAudioBuffer / ConvolverNode.Buffer / Connect / Render)
           starting from pinned HOST arrays: H2D of sources and IRs, IR preparation, render, D2H of the result all inside
           the timed region (wall clock between device synchronisations), max over ranks.  Reported with asynchronous
           uploads (GAC_FLAG_ASYNC_UPLOAD: headline) and, beside it, with the reference's copy-during-the-call semantics.
  parity = inside the run: (1) `parity_max_err` — a sharded render of the first max(2, N) voices of the workload, sharded by
           the same function, at full length against the CPU oracle; (2) `reduce_check_max_err` — the timed, reduced bus against
           the float64 sum of every rank's own un-sharded render of its shard.
  roofline: dominant kernel = the spectral MAC (K6) as a fast convolution along block time (csrc/fft2.cu).  achieved =
           COMPULSORY bytes (XT read once, YT written once, one set of second-level IR spectra read once) / K6's CUDA-event
           duration, against the measured HBM peak; `traffic` = dram bytes of the K6 launches measured by ncu in a child
           process of THIS run (rank 0, N = 1).  `whole_render` gives the same for all kernels of the render.
  cpu_baseline: the CPU oracle (C++ restatement of the reference algorithm; the reference is C#/.NET 9 — bench.py probes
           `dotnet` and says so) on a bounded sample of the same workload, 1 thread (a context renders on one thread).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import shutil
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from tests import synth  # noqa: E402
from graphaudio_b200 import sharding  # noqa: E402

FS = 48000
WORKLOADS = {
    # total_voices: sharded over the ranks (strong scaling); voices_per_gpu: fixed per rank (weak scaling)
    "c3": dict(voices_per_gpu=128, src_s=10.0, ir_s=2.0, render_s=12.0, bus_gain=1.0 / 32, kind="c3", scaling="weak",
               desc="C3 (BASELINE configs[2], the north-star config: 1024 voices on 8 GPUs = 128 voices per GPU): per voice stereo 10 s noise -> "
                    "BiQuadFilterNode lowpass, a-rate cutoff 2 -> 12 kHz -> GainNode a-rate automation -> ConvolverNode 2 s stereo IR per voice; bus "
                    "GainNode(1/32), 12 s @ 48 kHz; voices sharded over the GPUs, one ncclReduce of the bus"),
    "c5": dict(voices_per_gpu=32, src_s=10.0, ir_s=10.0, render_s=20.0, bus_gain=1.0 / 16, kind="c5", scaling="weak", fs=96000, src_rate=44100,
               desc="C5 (BASELINE configs[4]: 256 voices on 8 GPUs = 32 voices per GPU): per voice stereo 10 s noise at 44.1 kHz -> CubicResampler to 96 kHz "
                    "-> GainNode a-rate automation -> convolver with a 10 s stereo IR, 512-frame partitions; bus GainNode(1/16), 20 s @ 96 kHz"),
    "c2": dict(voices_per_gpu=64, src_s=10.0, ir_s=2.0, render_s=12.0, bus_gain=1.0 / 8, kind="c2", scaling="weak",
               desc="C2 (BASELINE configs[1]): 64 voices per GPU x (stereo 10 s noise -> GainNode a-rate automation -> ConvolverNode 2 s stereo IR per voice) "
                    "-> bus GainNode(1/8), 12 s @ 48 kHz"),
    "c1": dict(voices_per_gpu=1, src_s=10.0, ir_s=1.0, render_s=11.0, bus_gain=1.0, kind="c1", scaling="weak",
               desc="C1 (BASELINE configs[0]): 1 voice, stereo 10 s noise -> ConvolverNode 1 s stereo IR, 11 s @ 48 kHz"),
}


def total_voices(wl, world):
    return wl["total_voices"] if "total_voices" in wl else wl["voices_per_gpu"] * world


def shard_of(wl, rank, world):
    """[lo, hi) global voice indices of `rank`."""
    return sharding.shard_range(total_voices(wl, world), rank, world)


def config_of(wl, world, args):
    """The `config` object of the JSON line — the same keys and values from both arms (ours / reference)."""
    V = total_voices(wl, world)
    return {"workload": wl["desc"], "voices": V, "voices_per_gpu": V // world if V % world == 0 else V / world, "partition": args.partition,
            "sample_rate": wl.get("fs", FS), "frames": int(wl["render_s"] * wl.get("fs", FS)),
            "parallelism": f"voices sharded x{world}, one ncclReduce of the bus" if world > 1 else "single GPU",
            "l2": "working set per step (sources, signals and spectrograms: GBs) exceeds the 126 MB L2; no flush needed",
            "timing": "CUDA events on the library's launch stream (gac_get_stats.ms_total), max over ranks"}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)", float(d.get("sm_max_mhz", 1965.0))
    return 6650.0, "fallback (B200_PROFILING.md)", 1965.0


def dotnet_probe():
    """the harness first probes `dotnet --version`; when present, tools/dotnet_crosscheck can time the real
    reference.  (Absent from this image and from the GPU box: the expected case.)"""
    exe = shutil.which("dotnet")
    if not exe:
        return {"present": False}
    try:
        v = subprocess.run([exe, "--version"], capture_output=True, text=True, timeout=20).stdout.strip()
    except Exception as e:  # noqa: BLE001
        v = f"error: {e}"
    return {"present": True, "version": v, "note": "run tools/dotnet_crosscheck/run.sh to time and cross-check the real reference"}


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


def bind_to_gpu_cpus(local):
    """Pins this rank to the CPUs next to its GPU (the PCIe device's local_cpulist) BEFORE any page-locked memory is allocated,
    so that the staging arrays live on the GPU's NUMA node.  Returns a description for the JSON line."""
    info = {"cpus_online": os.cpu_count()}
    try:
        bdf = subprocess.run(["nvidia-smi", "-i", str(local), "--query-gpu=pci.bus_id", "--format=csv,noheader"], capture_output=True, text=True,
                             timeout=20).stdout.strip().lower()
        if bdf.startswith("00000000:"):
            bdf = bdf[4:]
        base = f"/sys/bus/pci/devices/{bdf}"
        with open(base + "/local_cpulist") as f:
            cpulist = f.read().strip()
        node = None
        if os.path.exists(base + "/numa_node"):
            with open(base + "/numa_node") as f:
                node = int(f.read().strip())
        cpus = set()
        for part in cpulist.split(","):
            if not part:
                continue
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        info.update({"gpu_pci": bdf, "gpu_numa_node": node, "gpu_local_cpus": cpulist})
        allowed = os.sched_getaffinity(0)
        want = cpus & allowed
        if want and want != allowed:
            os.sched_setaffinity(0, want)
            info["bound"] = True
        else:
            info["bound"] = False  # one NUMA node (or no topology information): nothing to choose
    except Exception as e:  # noqa: BLE001
        info["bound"] = False
        info["error"] = str(e)[:120]
    return info


def make_inputs(wl, lo, hi, pinned, cheap=False):
    """Per-voice host arrays for global voices [lo, hi) (pinned when torch/cuda is available): (src[2], ir[2], gains).
    cheap: voice v re-uses the samples of voice lo + (v - lo) % 8 (content does not matter: the ncu traffic child)."""
    nsrc, nir = int(wl["src_s"] * wl.get("src_rate", wl.get("fs", FS))), int(wl["ir_s"] * wl.get("fs", FS))
    alloc = None
    if pinned:
        import torch
        def alloc(rows, n):  # noqa: E306  one page-locked block per buffer, the channels are its rows
            return torch.empty((rows, n), dtype=torch.float32, pin_memory=True).numpy()
    voices = []
    for v in range(lo, hi):
        if cheap and v - lo >= 8:
            src, ir, _ = voices[(v - lo) % 8]
            voices.append((src, ir, synth.voice_gains(v)))
            continue
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


def build_graph(api, wl, voices, bus_gain=None, **kw):
    bus_gain = wl["bus_gain"] if bus_gain is None else bus_gain
    if wl["kind"] == "c1":
        return synth.build_c1(api, FS, voices[0][0], voices[0][1], **kw)
    if wl["kind"] == "c3":
        return synth.build_c3(api, FS, voices, bus_gain, **kw)
    if wl["kind"] == "c5":
        return synth.build_c5(api, wl["fs"], wl["src_rate"], voices, bus_gain, **kw)
    return synth.build_c2(api, FS, voices, bus_gain, **kw)


def build_into(api, wl, voices, ctx, bus_gain=None):
    """Builds the workload's graph inside an existing context."""
    class _Shim:
        pass
    shim = _Shim()
    for name in dir(api):
        setattr(shim, name, getattr(api, name))
    shim.OfflineAudioContext = lambda fs, **kw: ctx
    return build_graph(shim, wl, voices, bus_gain)


# ----------------------------------------------------------------------------------------------- reference arm / cpu baseline
def host_threads():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_render_sample(wl, n_voices, render_s, threads=1):
    """Times the CPU oracle on the first n_voices of the workload (same graph, same IR length).  One OfflineAudioContext renders on one
    thread, as in the reference; with threads > 1 the voices are sharded over that many contexts, each rendered by its own host thread
    (the oracle is called through ctypes, which releases the GIL), and the buses are summed afterwards — what a user of the reference
    would do with a many-core host.  Returns (voice-seconds per second, seconds, threads used)."""
    from oracle import ga_oracle as O
    threads = max(1, min(threads, n_voices))
    voices = make_inputs(wl, 0, n_voices, pinned=False)
    n = int(render_s * wl.get("fs", FS))
    ctxs = [build_graph(O, wl, voices[sharding.shard_range(n_voices, t, threads)[0]:sharding.shard_range(n_voices, t, threads)[1]])
            for t in range(threads)]
    outs = [None] * threads

    def work(t):
        outs[t] = ctxs[t].Render(n)

    t0 = time.perf_counter()
    if threads == 1:
        work(0)
    else:
        ths = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
        for th in ths:
            th.start()
        for th in ths:
            th.join()
    out = outs[0]
    for o in outs[1:]:
        out = out + o
    dt = time.perf_counter() - t0
    assert np.isfinite(out).all()
    return n_voices * render_s / dt, dt, threads


CPU_NOTE = ("CPU oracle = C++ restatement of the reference algorithm (the reference is C#/.NET 9; no dotnet in this image); one "
            "OfflineAudioContext renders on one thread, as in the reference; the voices are sharded over one context per host thread")


def run_reference(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    V = total_voices(wl, world)
    T = host_threads()
    nv = min(max(args.ref_voices, 2 * T), V)  # at least two voices per host thread and step
    vals = []
    used = 1
    for i in range(args.warmup + args.steps):
        v, dt, used = cpu_render_sample(wl, nv, wl["render_s"], threads=T)
        if i >= args.warmup:
            vals.append((v, dt))
    value = float(np.mean([v for v, _ in vals]))
    ms = float(np.mean([dt for _, dt in vals])) * 1e3
    sample = f"{nv} of {V} voices of the workload on {used} host threads, full {wl['render_s']} s render, per step; {CPU_NOTE}"
    line = {
        "impl": "reference", "metric": "voice-seconds rendered/sec", "value": value, "unit": "voice-s/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": wl["scaling"], "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": config_of(wl, world, args),
        "cpu_baseline": {"value": value, "unit": "voice-s/s", "cores": used, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "voice-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "realtime_factor": value / nv, "dotnet": dotnet_probe(),
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------- ncu traffic (child process)
K6_REGEX = "regex:k_fft2_conv|k_fft2_sum16"


def ncu_traffic_live(args, wl, n_voices, timeout_s=420):
    """DRAM bytes of the K6 launches of ONE render of this workload (n_voices voices), measured by ncu on a child process that
    runs the resident arm with this very library.  Returns a dict or {"error": ...}."""
    ncu = shutil.which("ncu") or "/usr/local/cuda/bin/ncu"
    if not os.path.exists(ncu):
        return {"error": "ncu not found"}
    log = os.path.join("/tmp", f"gac_ncu_traffic_{os.getpid()}.csv")
    cmd = [ncu, "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum", "--kernel-name", K6_REGEX, "--clock-control", "none", "--csv",
           "--log-file", log, sys.executable, os.path.abspath(__file__), "--traffic-child", "--workload", args.workload, "--voices", str(n_voices),
           "--partition", str(args.partition)] + (["--uniform-segments"] if args.uniform_segments else [])
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout_s)
    except Exception as e:  # noqa: BLE001
        return {"error": f"ncu child failed: {e}"[:200]}
    if r.returncode != 0 or not os.path.exists(log):
        return {"error": ("ncu child rc=%d: " % r.returncode + (r.stderr or r.stdout)[-300:])}
    import csv
    unit_scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    with open(log) as f:
        rows = [ln for ln in f if not ln.startswith("==")]
    per_launch = {}
    for row in csv.DictReader(rows):
        name = row.get("Metric Name", "")
        if not name.startswith("dram__bytes"):
            continue
        val = float(row["Metric Value"].replace(",", "")) * unit_scale.get(row.get("Metric Unit", "byte"), 1.0)
        d = per_launch.setdefault(int(row.get("ID", "0")), {"rd": 0.0, "wr": 0.0})
        d["rd" if "read" in name else "wr"] += val
    os.remove(log)
    if not per_launch:
        return {"error": "no K6 launch in the ncu log"}
    # the child renders twice (the first render prepares the double-length IR spectra and runs the single-length plan):
    # the launches of the LAST render are the second half
    ids = sorted(per_launch)
    last = ids[len(ids) // 2:] if len(ids) % 2 == 0 and len(ids) >= 2 else ids
    rd = sum(per_launch[i]["rd"] for i in last)
    wr = sum(per_launch[i]["wr"] for i in last)
    return {"dram_bytes": rd + wr, "dram_bytes_read": rd, "dram_bytes_write": wr, "k6_launches": len(last), "voices": n_voices,
            "how": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum on a child process of this bench run (second of two renders, same library)"}


def traffic_child(args, wl):
    """Runs under ncu: two resident renders of the workload with cheap inputs (content does not change the traffic)."""
    import graphaudio_b200 as G
    n = int(wl["render_s"] * wl.get("fs", FS))
    voices = make_inputs(wl, 0, args.voices, pinned=False, cheap=True)
    ctx = build_graph(G, wl, voices, device_id=0, partition=args.partition, uniform_segments=args.uniform_segments)
    ctx.MarkBus(ctx.bus)
    out = np.zeros((2, n), np.float32)
    ctx.RenderSharded(out, n, 0)
    ctx.RenderSharded(out, n, 0)
    ctx.Dispose()


# ----------------------------------------------------------------------------------------------- our arm
def run_ours(args, wl):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    host = bind_to_gpu_cpus(local)  # before torch allocates page-locked memory
    import torch
    import torch.distributed as dist
    import graphaudio_b200 as G
    from graphaudio_b200 import _native as N
    from graphaudio_b200.api import check

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

    n = int(wl["render_s"] * wl.get("fs", FS))
    lo, hi = shard_of(wl, rank, world)
    V = total_voices(wl, world)
    t_gen = time.perf_counter()
    voices = make_inputs(wl, lo, hi, pinned=True)
    t_gen = time.perf_counter() - t_gen
    h2d_bytes = sum(a.nbytes for v in voices for a in (v[0] + v[1]))
    d2h_bytes = 2 * n * 4
    L = N.lib()
    if wl["kind"] == "c5" and args.partition == 128:
        args.partition = 512
    ctx_kw = dict(device_id=local, tile_blocks=args.tile_blocks, partition=args.partition, mac_variant=args.mac_variant)

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

    # ---- resident arm: inputs in HBM, graph flattened once; the SAME entry point at every N
    ctx = build_graph(G, wl, voices, uniform_segments=args.uniform_segments, **ctx_kw)
    ctx.MarkBus(ctx.bus)
    comm(ctx)
    graph = ctx._graph()
    out_pinned = torch.zeros((2, n), dtype=torch.float32, pin_memory=True)  # the result lands in page-locked host memory
    out_host = out_pinned.numpy()
    out_ptrs = (N.fp * 2)(*[out_host[c].ctypes.data_as(N.fp) for c in range(2)])
    st = N.gac_stats()

    def step_resident():
        check(L.gac_render_sharded(ctx._h, graph, n, 0, out_ptrs, 2))
        check(L.gac_get_stats(ctx._h, C.byref(st)))
        return st.as_dict()

    sampler = ClockSampler(local)
    sampler.start()  # nvidia-smi needs a few hundred ms to deliver its first sample: start it before the warm-up
    warm = max(3, args.warmup)
    for _ in range(warm):
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
    timed_result = out_host.copy()
    L.gac_graph_destroy(graph)

    # ---- parity (2): the reduced bus against the float64 sum of every rank's own, un-sharded render of its shard
    own = np.zeros((2, n), np.float32)
    ctx.Render(own, n, 0)  # bus gain applied by this rank itself
    acc = torch.from_numpy(own.astype(np.float64)).cuda()
    if world > 1:
        dist.all_reduce(acc, op=dist.ReduceOp.SUM)
    reduce_err = float(np.abs(acc.cpu().numpy() - timed_result.astype(np.float64)).max()) if rank == 0 else 0.0
    bus_peak = float(np.abs(timed_result).max()) if rank == 0 else 0.0
    del acc
    ctx.Dispose()
    del ctx

    # ---- parity (1): the first max(2, N) voices of the workload, sharded by the same function, full length, against the oracle
    parity = None
    if not args.no_parity:
        k = max(2, world)
        plo, phi = sharding.shard_range(k, rank, world)
        pg = 0.5 if k <= 4 else 0.25  # bus gain of the subset: peak in [0.25, 1]
        pv = [voices[i - lo] if lo <= i < hi else None for i in range(plo, phi)]
        if any(v is None for v in pv):
            pv = make_inputs(wl, plo, phi, pinned=False)
        pc = build_graph(G, wl, pv, bus_gain=pg, **ctx_kw)
        pc.MarkBus(pc.bus)
        comm(pc)
        pout = np.zeros((2, n), np.float32)
        pc.RenderSharded(pout, n, 0)
        pc.Dispose()
        if rank == 0:
            from oracle import ga_oracle as O
            t_or = time.perf_counter()
            ref = build_graph(O, wl, make_inputs(wl, 0, k, pinned=False), bus_gain=pg).Render(n)
            parity = {"max_err": float(np.abs(pout - ref).max()), "peak": float(np.abs(ref).max()), "voices": k,
                      "oracle_s": time.perf_counter() - t_or,
                      "what": f"voices 0..{k - 1} of the workload sharded over {world} rank(s) by the same function, full {wl['render_s']} s, "
                              f"bus gain {pg}, through gac_render_sharded, against the CPU oracle"}
        barrier()

    # ---- e2e arm: host arrays -> public API -> host result, everything inside the timed region.
    # Per step: a fresh OfflineAudioContext, PlayableAudioBuffer uploads (H2D from pinned host memory),
    # ConvolverNode.Buffer (IR preparation), Connect, Render (D2H of the result on the root).
    # N > 1: the NCCL communicator of each fresh context is bootstrapped BEFORE its timed region (not part of a render).
    import gc
    gc.collect()
    gc.freeze()  # objects allocated so far leave the cyclic collector's sight (one step in ~50 paid a 40 ms full collection)

    def e2e_arm(async_upload, steps, warm_steps):
        lat = []
        for i in range(warm_steps + steps):
            def fresh():
                return G.OfflineAudioContext(wl.get("fs", FS), async_upload=async_upload, **ctx_kw)
            if world > 1:
                c = fresh()
                comm(c)
            barrier()
            t0 = time.perf_counter()
            if world == 1:
                c = fresh()  # N = 1: creating the context is part of the step
            build_into(G, wl, voices, c)
            c.MarkBus(c.bus)
            g = c._graph()
            check(L.gac_render_sharded(c._h, g, n, 0, out_ptrs, 2))
            L.gac_graph_destroy(g)
            barrier()
            if i >= warm_steps:
                lat.append(time.perf_counter() - t0)
            c.Dispose()
        tl = torch.tensor([sum(lat) * 1e3], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tl, op=dist.ReduceOp.MAX)
        return float(tl[0]) / steps, lat

    e2e_steps = max(2, min(args.steps, args.e2e_steps))
    e2e_ms, lat = e2e_arm(not args.sync_upload, e2e_steps, 2)
    # the end-to-end arm renders the same graph as the resident arm (fresh context, uploads in flight, voice batches, impulse
    # responses prepared on the way): its last result against the resident arm's
    e2e_err = float(np.abs(out_host - timed_result).max()) if rank == 0 else 0.0
    e2e_other_ms, _ = e2e_arm(args.sync_upload, max(2, e2e_steps // 2), 1)  # the other upload mode, beside the headline

    # ---- cpu baseline (rank 0, N = 1 only): bounded sample of the same workload on the host
    cpu = None
    if world == 1 and not args.no_cpu:
        T = host_threads()
        nv = min(max(args.cpu_voices, 2 * T), V)
        v, dt, used = cpu_render_sample(wl, nv, wl["render_s"], threads=T)
        v1, dt1, _ = cpu_render_sample(wl, min(4, V), wl["render_s"], threads=1)
        cpu = {"value": v, "unit": "voice-s/s", "cores": used, "kind": "port", "single_thread_value": v1,
               "sample": f"{nv} of {V} voices on {used} host threads, full {wl['render_s']} s render ({dt:.1f} s of wall time; one thread alone: "
                         f"{v1:.1f} voice-s/s on {min(4, V)} voices); {CPU_NOTE}"}

    # ---- ncu traffic of the K6 launches (rank 0, N = 1 only; a child process under ncu, after every timed region)
    traffic = None
    if world == 1 and not args.no_traffic:
        torch.cuda.synchronize()
        traffic = ncu_traffic_live(args, wl, hi - lo)

    if rank == 0:
        peak, peak_src, sm_max = measured_peaks()
        s_last = stats[-1]
        mean = lambda k: float(np.mean([s[k] for s in stats]))  # noqa: E731
        mac_ms = mean("ms_mac")
        used = int(s_last["mac_variant_used"])
        big = int(s_last.get("mac_big_segments", 0))
        units = s_last["conv_units"]
        Bp, Cb = args.partition, args.partition + 1
        nvox = hi - lo
        Npad = -(-n // Bp) * Bp
        Qs = -(-(Npad // Bp) // 16) * 16
        moved = s_last["mac_bytes_moved"]            # XT + YT + every H2 table the plan reads (both lengths with mixed segments)
        # compulsory bytes of K6: XT in, YT out (B+1 rows of Qs blocks per channel-convolver), ONE set of second-level IR spectra
        spectro = units / (Npad // Bp) * Cb * Qs * 8.0
        h2_one = s_last.get("mac_h2_bytes_single", 0.0)
        fused = int(s_last.get("fanin_members", 0))   # convolvers summed as spectra (fan-in fusion): per-voice YT, K7 and fan-in inputs are gone
        # moved = XT once + every H2 table the plan reads (single length: M float2 per row, + the double-length rows, 2 M, with mixed
        # segments) + what K6 writes (per-voice YT, or one partial spectrogram per voice chunk with fan-in fusion)
        compulsory = moved - (2 * h2_one if big > 0 else 0.0) if used == 3 else moved
        yt_bytes = max(0.0, compulsory - spectro - h2_one) if used == 3 else spectro   # what K6 writes and K7 reads
        k6_name = {1: "k_mac_stream (K6, direct sum, reference op order)", 2: "k_mac_tiled (K6, register-tiled direct sum, FFMA)",
                   4: "k_mac_tiled (K6, register-tiled direct sum, FFMA2)",
                   3: "k_fft2_conv16 (K6 spectral MAC as a fast convolution along block time)"}.get(used, "K6")
        kn = "k_fft2_sum16" if fused else "k_fft2_conv16"
        if used == 3 and fused:
            k6_name = "k_fft2_sum16 (K6 as a fast convolution along block time, the voices of a fan-in summed as second-level spectra)"
        if used == 3 and big > 0:
            k6_name = (f"{kn}<2M> + {kn}<M> (K6 as a fast convolution along block time" + (", the voices of a fan-in summed as second-level spectra" if fused else "") +
                       f": {big} double-length overlap-save segment(s) in front, two launches timed together)")
        # bytes every kernel of the render has to move once (compulsory): source in (biquad or K5), biquad streams, signal rows,
        # automation tables out + in, K5 signal in + XT out, K6, K7 YT in + signal out, fan-in reads + bus write
        sig_bytes = nvox * 2 * Npad * 4.0
        n_tabs = nvox + (1 if wl["kind"] == "c3" else 0)  # per-voice gain tables (+ the shared cutoff sweep)
        whole = {
            "automation": n_tabs * Npad * 4.0,
            "biquad": (sig_bytes * 2 + Npad * 4.0) if wl["kind"] == "c3" else 0.0,      # source in, filtered signal out, cutoff table
            "K5": sig_bytes + nvox * Npad * 4.0 + spectro,                              # signal + gain table in, XT out
            "K6": compulsory,
            "K7": yt_bytes + (2 * Npad * 4.0 if fused else sig_bytes),                   # YT in, signal out (fused: the group's sum)
            "mix": (2 * Npad * 4.0 if fused else sig_bytes) + 2 * Npad * 4.0 * 2,        # fan-in reads, bus write + bus gain pass
        }
        whole_bytes = float(sum(whole.values()))
        conv_ms = mean("ms_fft_fwd") + mac_ms + mean("ms_fft_inv")
        conv_bytes = whole["K5"] + whole["K6"] + whole["K7"]
        sm_mhz = clocks.get("sm_mhz") or sm_max
        fp32_peak_tf = 148 * 128 * 2 * sm_max * 1e6 / 1e12
        flops = s_last["mac_flops"]
        line = {
            "metric": "voice-seconds rendered/sec", "value": V * wl["render_s"] / (dev_ms_step * 1e-3), "unit": "voice-s/s",
            "n_gpus": world, "steps": args.steps, "warmup": warm, "ms_per_step": dev_ms_step, "higher_is_better": True,
            "scaling": wl["scaling"], "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_of(wl, world, args),
            "realtime_factor": wl["render_s"] / (dev_ms_step * 1e-3),
            "samples_per_s": V * 2 * n / (dev_ms_step * 1e-3),
            "wall_ms_per_step": wall_ms_step,
            "parity_max_err": parity["max_err"] if parity else None,
            "parity": parity,
            "reduce_check_max_err": reduce_err,
            "bus_peak": bus_peak,
            "e2e": {"value": V * wl["render_s"] / (e2e_ms * 1e-3), "unit": "voice-s/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": int(V * h2d_bytes / max(1, nvox)), "d2h_bytes_per_step": d2h_bytes,
                    "async_upload": not args.sync_upload, "steps": e2e_steps, "vs_resident_max_err": e2e_err,
                    ("sync_upload_ms_per_step" if not args.sync_upload else "async_upload_ms_per_step"): e2e_other_ms,
                    "h2d_gb_per_s_per_rank": h2d_bytes / (e2e_ms * 1e-3) / 1e9,
                    "note": ("communicator bootstrap outside the timed region; " if world > 1 else "") +
                            "context creation (N = 1), H2D of this rank's sources and IRs, IR preparation, graph flattening, sharded render, "
                            "D2H of the bus inside; max over ranks"},
            "gpu_launches": int(sum(s["kernel_launches"] for s in stats)),
            "clocks": clocks,
            "host": {**host, "input_generation_s": t_gen},
            # the dominant kernel against the HBM roofline: COMPULSORY bytes per launch / its CUDA-event duration; `traffic` = DRAM
            # bytes of the same launches measured by ncu inside this run (N = 1)
            "roofline": {"bound": "hbm", "achieved": compulsory / (mac_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": compulsory / (mac_ms * 1e-3) / 1e9 / peak,
                         "traffic": traffic.get("dram_bytes") if traffic else None,
                         "traffic_detail": traffic,
                         "kernel": k6_name, "peak_source": peak_src, "ms_per_launch": mac_ms, "algorithmic_bytes_per_launch": compulsory,
                         "bytes_definition": "compulsory: XT read once + ONE set of second-level IR spectra (the second, double-length table the "
                                             "mixed-segment plan also reads is NOT counted) + what K6 writes (with fan-in fusion: one partial "
                                             "spectrogram per voice chunk and output channel instead of one YT per channel-convolver)",
                         "fanin_fusion": {"groups": int(s_last.get("fanin_groups", 0)), "convolvers": fused},
                         "moved_bytes_per_launch": moved, "frac_on_moved_bytes": moved / (mac_ms * 1e-3) / 1e9 / peak,
                         "convolver_K5_K6_K7": {"ms": conv_ms, "bytes": conv_bytes, "achieved": conv_bytes / (conv_ms * 1e-3) / 1e9,
                                                "frac": conv_bytes / (conv_ms * 1e-3) / 1e9 / peak},
                         "whole_render": {"ms": dev_ms_step, "bytes": whole_bytes, "bytes_by_stage": whole,
                                          "achieved": whole_bytes / (dev_ms_step * 1e-3) / 1e9,
                                          "frac": whole_bytes / (dev_ms_step * 1e-3) / 1e9 / peak,
                                          "note": "rank 0's shard; compulsory bytes of every stage over the whole step's device time"}},
            # what the REFERENCE algorithm would move for the same units (FDL + IR walked once per quantum, SURVEY.md §8d T = 1) over
            # K6's time: the algorithmic gain of doing the partition sum as a convolution along block time — not a roofline
            "algorithmic_gain": {"reference_bytes_per_launch": s_last["algorithmic_bytes"],
                                 "bytes_ratio": s_last["algorithmic_bytes"] / compulsory if compulsory else None},
            "fp32": {"achieved_tflops": flops / (mac_ms * 1e-3) / 1e12, "peak_tflops": fp32_peak_tf,
                     "frac": flops / (mac_ms * 1e-3) / 1e12 / fp32_peak_tf, "flops_per_launch": flops,
                     "peak_source": f"148 SMs x 128 FMA/clk x 2 x {sm_max:.0f} MHz; median SM clock under load {sm_mhz} MHz"},
            "kernel_ms": {k: mean(k) for k in ["ms_source", "ms_automation", "ms_biquad", "ms_gain", "ms_fft_fwd", "ms_mac", "ms_fft_inv",
                                               "ms_mix", "ms_d2h"]},
            "dotnet": dotnet_probe(),
        }
        line["kernel_ms"]["sum"] = float(sum(line["kernel_ms"].values()))
        if cpu:
            line["cpu_baseline"] = cpu
        line["e2e"]["ms_per_step_quartiles"] = [float(x) * 1e3 for x in np.percentile(lat, [0, 25, 50, 75, 100])]  # this rank's steps
        line["e2e"]["ms_per_step_all"] = [round(float(x) * 1e3, 3) for x in lat]
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--voices", type=int, default=0, help="total voices (c3) / voices per GPU (c2): 0 = the workload's own count")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--partition", type=int, default=128)
    ap.add_argument("--tile-blocks", dest="tile_blocks", type=int, default=32)
    ap.add_argument("--mac-variant", dest="mac_variant", type=int, default=0,
                    help="K6 algorithm (gac_context_desc.mac_variant): 0 default (second-level FFT), 1 streaming direct sum, 4 register-tiled direct sum")
    ap.add_argument("--cpu-voices", dest="cpu_voices", type=int, default=24, help="voices of the cpu_baseline sample")
    ap.add_argument("--ref-voices", dest="ref_voices", type=int, default=8, help="voices per step of the --impl reference arm")
    ap.add_argument("--e2e-steps", dest="e2e_steps", type=int, default=8)
    ap.add_argument("--no-cpu", dest="no_cpu", action="store_true")
    ap.add_argument("--no-parity", dest="no_parity", action="store_true")
    ap.add_argument("--no-traffic", dest="no_traffic", action="store_true", help="skip the ncu child that measures K6's DRAM traffic")
    ap.add_argument("--uniform-segments", dest="uniform_segments", action="store_true",
                    help="resident arm: K6 without the double-length overlap-save segments in front (GAC_FLAG_UNIFORM_SEGMENTS), for A/B runs")
    ap.add_argument("--sync-upload", dest="sync_upload", action="store_true",
                    help="e2e headline with the reference's copy-during-the-call semantics instead of GAC_FLAG_ASYNC_UPLOAD")
    ap.add_argument("--traffic-child", dest="traffic_child", action="store_true", help=argparse.SUPPRESS)
    args = ap.parse_args()
    wl = dict(WORKLOADS[args.workload])
    if args.voices > 0:
        wl["total_voices" if "total_voices" in wl else "voices_per_gpu"] = args.voices
    if args.traffic_child:
        traffic_child(args, wl)
    elif args.impl == "reference":
        run_reference(args, wl)
    else:
        run_ours(args, wl)


if __name__ == "__main__":
    main()
