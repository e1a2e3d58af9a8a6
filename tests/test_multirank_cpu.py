"""CPU test of the N > 1 path (gloo, world_size 2): voices sharded by rank, per-rank partial buses summed with ONE
reduce to rank 0, bus GainNode applied on the root AFTER the reduce — the schedule gac_render_sharded implements with
ncclReduce.  The per-rank renders here come from the CPU oracle (the checker), so this test pins the host-side logic:
the shard partition, the reduce, the bus-op-after-reduce order, and the float32 association error it introduces."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FS = 48000
NV = 6
BUS_GAIN = 0.3
N_FRAMES = 12000


def _voices():
    from tests import synth
    out = []
    for v in range(NV):
        src, ir = synth.make_voice_inputs(v, 8000, 2000)
        out.append((src, ir, synth.voice_gains(v)))
    return out


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from graphaudio_b200 import sharding
    from oracle import ga_oracle as O
    from tests import synth
    mine = sharding.shard_list(_voices(), rank, world)
    # the shard renders up to the bus fan-in: bus gain 1.0 here, the real bus gain is applied on the root after the reduce
    partial = synth.build_c2(O, FS, mine, 1.0, t_scale=0.02).Render(N_FRAMES)
    t = torch.from_numpy(partial.copy())
    dist.reduce(t, dst=0, op=dist.ReduceOp.SUM)
    if rank == 0:
        bus = t.numpy() * np.float32(BUS_GAIN)  # GainNode: float32 multiply (Nodes/GainNode.cs:49-58)
        q.put(bus)
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.timeout(300)
def test_two_rank_sharded_bus_reduce_matches_single_render():
    from oracle import ga_oracle as O
    from tests import synth
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    sharded = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    whole = synth.build_c2(O, FS, _voices(), BUS_GAIN, t_scale=0.02).Render(N_FRAMES)
    err = np.abs(sharded - whole).max()
    assert np.abs(whole).max() > 1e-3
    # sequential fan-in vs (shard 0 sum) + (shard 1 sum): float32 association only
    assert err <= 1e-6, err
