"""The host's ceiling under the end-to-end step at N GPUs: every rank copies what one e2e step of the bench workload uploads (128 voices:
a [2, 480000] source and a [2, 96000] impulse response each, 590 MB, from page-locked memory) at the same time, nothing else running.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/h2d_concurrent.py
"""
import json, os, time
import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
hs = [torch.empty((2, n), dtype=torch.float32, pin_memory=True) for _ in range(128) for n in (480000, 96000)]
ds = [torch.empty_like(h, device="cuda") for h in hs]
nbytes = sum(h.numel() * 4 for h in hs)
res = {}
for mode in ("alone", "together"):
    ms = []
    for rep in range(6):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if mode == "together" or rank == 0:
            for h, d in zip(hs, ds):
                d.copy_(h, non_blocking=True)
        torch.cuda.synchronize()
        ms.append((time.perf_counter() - t0) * 1e3)
    t = torch.tensor([sorted(ms[1:])[len(ms[1:]) // 2]], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    res[mode] = float(t[0])
if rank == 0:
    print(json.dumps({"ranks": world, "mb_per_rank": nbytes / 1e6, "ms_rank0_alone": res["alone"], "gb_s_alone": nbytes / res["alone"] / 1e6,
                      "ms_all_ranks_together_max": res["together"], "gb_s_per_rank_together": nbytes / res["together"] / 1e6,
                      "gb_s_aggregate": world * nbytes / res["together"] / 1e6, "cpus": os.cpu_count()}))
if world > 1:
    dist.destroy_process_group()
