"""Dumps the cross-check cases (inputs as little-endian .f32 + case.txt) for tools/dotnet_crosscheck/Program.cs and, beside
them, the CPU oracle's render of each case (oracle_<case>.f32) for eyeballing.

    python tools/dotnet_crosscheck/dump_cases.py [<cases dir>]     (default: tools/dotnet_crosscheck/cases, git-ignored)

The cases are C1 .. C5 of BASELINE.json (plus a looping, rate-swept variant of C5) at sizes the oracle renders in well under a second each, built by the same
tests/synth.py builders the parity tests use; tests/test_reference_crosscheck.py re-creates the inputs from the seeds (nothing
but the reference's outputs has to be committed: tests/golden/ref_<case>.f32)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests import synth  # noqa: E402

CASES = {
    # name: kind, fs, frames, voices, src frames, ir frames, extra
    "c1_small": dict(kind="c1", sample_rate=48000, frames=12000, voices=1, src=8000, ir=3000),
    "c2_small": dict(kind="c2", sample_rate=48000, frames=12000, voices=3, src=8000, ir=3000, bus_gain=0.5, t_scale=0.02),
    "c3_small": dict(kind="c3", sample_rate=48000, frames=12000, voices=3, src=8000, ir=3000, bus_gain=0.5, t_scale=0.02, f0=300.0, f1=9000.0,
                     q=2.0, sweep_end=10.0),
    "c4_small": dict(kind="c4", sample_rate=48000, frames=12000, voices=1, src=9000, ir=2000, t_scale=0.04, f0=2000.0, f1=12000.0, q=0.707,
                     sweep_end=4.0),
    "c5_small": dict(kind="c5", sample_rate=96000, source_rate=44100, frames=24000, voices=2, src=8000, ir=5000, bus_gain=0.5, t_scale=0.02),
    # looping sources on the CubicResampler path (the 512-float wrap buffer, Nodes/AudioBufferSourceNode.cs:236-358) under a k-rate
    # PlaybackRate sweep that crosses an effective rate of exactly 1 nowhere (44.1 kHz buffers in a 48 kHz context)
    "loop_small": dict(kind="loop", sample_rate=48000, source_rate=44100, frames=12000, voices=2, src=3000, ir=2000, bus_gain=0.5, t_scale=0.02,
                       loop_start=0.01, loop_end=0.05, rate0=0.8, rate1=1.6, rate_t1=0.2),
}


def case_inputs(name):
    c = CASES[name]
    voices = []
    for v in range(c["voices"]):
        src, ir = synth.make_voice_inputs(200 + 10 * list(CASES).index(name) + v, c["src"], c["ir"])
        voices.append((src, ir, synth.voice_gains(v)))
    return c, voices


def build_case(api, name):
    c, voices = case_inputs(name)
    fs, ts = c["sample_rate"], c.get("t_scale", 1.0)
    if c["kind"] == "c1":
        return synth.build_c1(api, fs, voices[0][0], voices[0][1]), c
    if c["kind"] == "c2":
        return synth.build_c2(api, fs, voices, c["bus_gain"], t_scale=ts), c
    if c["kind"] == "c3":
        return synth.build_c3(api, fs, voices, c["bus_gain"], f0=c["f0"], f1=c["f1"], t_scale=ts, q=c["q"]), c
    if c["kind"] == "c4":
        return synth.build_c4(api, fs, voices[0][0], voices[0][1], t_scale=ts), c
    if c["kind"] == "loop":
        return synth.build_c5(api, fs, c["source_rate"], voices, c["bus_gain"], t_scale=ts, loop=(c["loop_start"], c["loop_end"]),
                              rate_ramp=(c["rate0"], c["rate1"], c["rate_t1"])), c
    return synth.build_c5(api, fs, c["source_rate"], voices, c["bus_gain"], t_scale=ts), c


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tools", "dotnet_crosscheck", "cases")
    from oracle import ga_oracle as O
    for name in CASES:
        c, voices = case_inputs(name)
        d = os.path.join(out, name)
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, "case.txt"), "w") as f:
            for k, v in c.items():
                if k not in ("src", "ir"):
                    f.write(f"{k} = {v!r}\n".replace("'", ""))
            for v, (_, _, g) in enumerate(voices):
                for i in range(3):
                    f.write(f"v{v}_g{i} = {float(np.float32(g[i]))!r}\n")
        for v, (src, ir, _) in enumerate(voices):
            for ch, a in enumerate(src):
                a.astype("<f4").tofile(os.path.join(d, f"v{v}_src{ch}.f32"))
            for ch, a in enumerate(ir):
                a.astype("<f4").tofile(os.path.join(d, f"v{v}_ir{ch}.f32"))
        ctx, _ = build_case(O, name)
        y = ctx.Render(c["frames"])
        y.astype("<f4").tofile(os.path.join(out, f"oracle_{name}.f32"))
        print(name, y.shape, float(np.abs(y).max()))


if __name__ == "__main__":
    main()
