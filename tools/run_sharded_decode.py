"""Multi-GPU check of the query-sharding launcher on real devices (NCCL), one process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/run_sharded_decode.py

Rank 0 owns the synthetic weights / latents (the encoder rank of SURVEY.md section 8e); they are broadcast once, every rank
decodes its share of the (pair, t) slabs -- or, when slabs < ranks, its row bands with a locally recomputed halo -- with
no collective inside the decode loop, the RGB is gathered to rank 0 and compared with rank 0's own single-GPU decode of
the whole job (same kernels, same per-query arithmetic: the comparison is exact)."""
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "stif-continuous-video-representation_b200"))
import stif_b200  # noqa: E402
from stif_b200 import synthetic as synth  # noqa: E402
from stif_b200.launcher import QueryShardLauncher, plan_units  # noqa: E402


def main() -> int:
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok = True
    # (pairs, H, W, out size, times, stress weights): slabs >= ranks, slabs < ranks (row bands + halo), odd sizes + big flows
    jobs = [(2, 64, 64, (256, 256), [i / 8 for i in range(8)], False),
            (1, 270, 480, (1080, 1920), [0.5], False),
            (1, 48, 40, (163, 141), [0.25], True)]
    if world >= 8:   # BASELINE.json config 4: 4K output, x8 temporal, one timestep per GPU
        jobs.append((1, 540, 960, (2160, 3840), [i / 8 for i in range(8)], False))
    for P, H, W, out_size, times, stress in jobs:
        launcher = QueryShardLauncher(mode="bf16")
        weights = synth.make_weights(3, stress) if rank == 0 else None
        launcher.broadcast_weights(weights)
        lat = fr = None
        if rank == 0:
            lat_np, fr_np = synth.make_inputs(11, P, H, W, 1.0 if stress else 0.05)
            lat, fr = torch.from_numpy(lat_np), torch.from_numpy(fr_np)
        launcher.broadcast_inputs(lat, fr, (P, H, W))
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        results = launcher.decode(times, out_size, halo=8)
        torch.cuda.synchronize()
        dt_local = time.perf_counter() - t0
        dist.barrier()
        t0 = time.perf_counter()                      # second pass: workspaces exist, geometry tables cached
        results = launcher.decode(times, out_size, halo=8)
        torch.cuda.synchronize()
        dt_local = time.perf_counter() - t0
        t = torch.tensor([dt_local], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        full = launcher.gather(results, times, out_size)
        if rank == 0:
            dec = stif_b200.STIFQueryDecoder(local, mode="bf16")
            dec.load_weights(weights)
            ref = dec.decode_stacked(launcher.latent, launcher.frames, times, out_size)   # [T,P,3,HH,WW]
            torch.cuda.synchronize()
            same = bool(torch.equal(full, ref))
            units = plan_units(P, len(times), out_size[0], world)
            kind = "slabs" if all(u.row_begin == 0 and u.row_end == out_size[0] for u in units) else "row bands"
            q = P * len(times) * out_size[0] * out_size[1]
            print(f"job {P}x{H}x{W} -> {out_size} T={len(times)} on {world} GPUs: {len(units)} units ({kind}), "
                  f"decode {float(t.item()) * 1e3:.2f} ms ({q / float(t.item()):.3e} q/s, second pass), "
                  f"gathered == single-GPU decode: {same}", flush=True)
            ok = ok and same
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    return 0 if int(flag.item()) == 1 else 1


if __name__ == "__main__":
    sys.exit(main())
