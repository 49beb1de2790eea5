"""Concurrent pinned H2D / D2H rates on N ranks of one box (the ceiling of the host entry point's e2e scaling):
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 profiles/pcie_probe_multi.py"""
import os, time, json
import torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
h_in = torch.empty(102_643_200 // 4).pin_memory(); d_in = torch.empty_like(h_in, device="cuda")
d_out = torch.empty(49_766_400 // 4, device="cuda"); h_out = torch.empty(d_out.shape).pin_memory()
s2 = torch.cuda.Stream()
def run(kind, n=10):
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        if kind in ("h2d", "both"): d_in.copy_(h_in, non_blocking=True)
        if kind in ("d2h", "both"):
            with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize()
    dt = torch.tensor([(time.perf_counter() - t0) / n], device="cuda", dtype=torch.float64)
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    return float(dt.item())
res = {"ranks": world}
for kind in ("h2d", "d2h", "both"):
    run(kind, 3)
    t = run(kind)
    nbytes = (h_in.nbytes if kind != "d2h" else 0) + (h_out.nbytes if kind != "h2d" else 0)
    res[kind] = {"ms_max_over_ranks": t * 1e3, "per_rank_GBps": nbytes / t / 1e9, "aggregate_GBps": world * nbytes / t / 1e9}
if rank == 0: print(json.dumps(res))
dist.destroy_process_group()
