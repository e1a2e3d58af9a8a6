"""Multi-GPU parity check, run under torchrun on N B200s (not collected by pytest; the CPU twin is test_multirank_cpu.py):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py

Every rank renders its shard of a C2-shaped graph with gac_render_sharded (one ncclReduce of the bus inside the library);
rank 0 compares the result with (a) the same graph rendered on ONE GPU and (b) the CPU oracle.  Gate: 1e-5.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import graphaudio_b200 as G  # noqa: E402
from graphaudio_b200 import sharding  # noqa: E402
from tests import synth  # noqa: E402

FS, BUS_GAIN, NF = 48000, 0.25, 40000
NV = int(os.environ.get("GAC_CHECK_VOICES", "11"))  # 11 voices: uneven shards; single-voice shards at 8 ranks


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    voices = []
    for v in range(NV):
        src, ir = synth.make_voice_inputs(v, 30000, 9000)
        voices.append((src, ir, synth.voice_gains(v)))
    mine = sharding.shard_list(voices, rank, world)
    ctx = synth.build_c2(G, FS, mine, BUS_GAIN, t_scale=0.05, device_id=local)
    ctx.MarkBus(ctx.bus)  # a shard may hold a single voice (or none): the bus fan-in must stay explicit
    idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt.copy_(torch.frombuffer(bytearray(G.OfflineAudioContext.CommUniqueId()), dtype=torch.uint8))
    dist.broadcast(idt, 0)
    ctx.CommInit(bytes(idt.cpu().numpy().tobytes()), rank, world)
    out = np.zeros((2, NF), np.float32)
    ctx.RenderSharded(out, NF, root=0)
    ok = True
    if rank == 0:
        single = synth.build_c2(G, FS, voices, BUS_GAIN, t_scale=0.05, device_id=local).Render(NF)
        from oracle import ga_oracle as O
        ref = synth.build_c2(O, FS, voices, BUS_GAIN, t_scale=0.05).Render(NF)
        e1, e2 = float(np.abs(out - single).max()), float(np.abs(out - ref).max())
        print(f"multi_gpu_check world={world}: peak {np.abs(ref).max():.3f}  |sharded - single GPU| {e1:.3e}  |sharded - oracle| {e2:.3e}")
        ok = e1 <= 1e-5 and e2 <= 1e-5
    dist.barrier()
    ctx.Dispose()
    dist.destroy_process_group()
    if not ok:
        raise SystemExit(1)


if __name__ == "__main__":
    main()
