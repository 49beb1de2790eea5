"""Host logic of the query-sharding launcher under gloo, world_size 2, on CPU (no GPU needed).
The per-unit decoder is a stub (a deterministic function of pair, t and pixel) so that planning,
the one-off broadcasts, the collective-free decode loop and the stitching gather are all exercised."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import PKG, ROOT


def test_plan_whole_slabs_round_robin(stif):
    units = stif.plan_units(num_pairs=3, num_times=4, HH=64, world=8)          # 12 slabs >= 8 ranks
    assert len(units) == 12 and all(u.row_begin == 0 and u.row_end == 64 for u in units)
    assert [u.rank for u in units] == [i % 8 for i in range(12)]
    assert sorted({(u.pair, u.t_index) for u in units}) == [(p, c) for p in range(3) for c in range(4)]


def test_plan_row_bands_when_slabs_are_scarce(stif):
    units = stif.plan_units(num_pairs=1, num_times=2, HH=1080, world=8)        # config 2 on 8 GPUs: 2 slabs -> 4 bands each
    assert len(units) == 8 and sorted(u.rank for u in units) == list(range(8))
    for c in range(2):
        bands = sorted((u.row_begin, u.row_end) for u in units if u.t_index == c)
        assert bands[0][0] == 0 and bands[-1][1] == 1080
        assert all(a[1] == b[0] for a, b in zip(bands, bands[1:]))             # contiguous, no overlap
    one = stif.plan_units(1, 8, 2160, 8)                                       # config 4: one timestep per GPU
    assert [(u.t_index, u.rank) for u in one] == [(i, i) for i in range(8)]
    tiny = stif.plan_units(1, 1, 3, 8)                                         # more ranks than rows
    assert sum(u.row_end - u.row_begin for u in tiny) == 3
    with pytest.raises(ValueError):
        stif.plan_units(0, 1, 4, 2)


def _stub_decode(lat, fr, t, out_size, rows, halo):
    HH, WW = out_size
    yy = torch.arange(HH, dtype=torch.float32).view(1, HH, 1)
    xx = torch.arange(WW, dtype=torch.float32).view(1, 1, WW)
    ch = torch.arange(3, dtype=torch.float32).view(3, 1, 1)
    full = lat.sum() + 10.0 * t + yy + 0.01 * xx + 100.0 * ch + fr.mean()
    if rows is None:
        return full
    out = torch.zeros_like(full)
    out[:, rows[0]:rows[1]] = full[:, rows[0]:rows[1]]
    return out


def _worker(rank, world, port, P, times, out_size, q):
    for p in (ROOT, PKG):
        sys.path.insert(0, p)
    import stif_b200
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        H = W = 4
        lat = fr = None
        if rank == 0:
            g = torch.Generator().manual_seed(0)
            lat = torch.randn(P, 3, 64, H, W, generator=g)
            fr = torch.rand(P, 2, 3, H, W, generator=g)
        launcher = stif_b200.QueryShardLauncher(decode_fn=_stub_decode, device="cpu")
        from oracle import synth
        w = launcher.broadcast_weights(synth.make_weights(0) if rank == 0 else None)
        launcher.broadcast_inputs(lat, fr, (P, H, W))
        results = launcher.decode(times, out_size, halo=2)
        full = launcher.gather(results, times, out_size, dst=0)
        q.put((rank, [tuple(vars(u).values()) for u, _ in results], float(w["encode_imnet.net.4.weight"].sum()),
               None if full is None else full))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("P,times", [(3, [0.0, 0.5]), (1, [0.25])])
def test_two_rank_decode_and_gather(stif, P, times):
    world, out_size = 2, (6, 5)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + 7 * P) % 500
    procs = [ctx.Process(target=_worker, args=(r, world, port, P, times, out_size, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    got.sort(key=lambda g: g[0])
    assert got[0][2] == got[1][2]                                              # both ranks hold rank 0's weights
    all_units = got[0][1] + got[1][1]
    plan = stif.plan_units(P, len(times), out_size[0], world)
    assert sorted(all_units) == sorted(tuple(vars(u).values()) for u in plan)  # every unit decoded exactly once
    full = got[0][3]
    g = torch.Generator().manual_seed(0)
    lat = torch.randn(P, 3, 64, 4, 4, generator=g)
    fr = torch.rand(P, 2, 3, 4, 4, generator=g)
    for c, t in enumerate(times):
        for p_ in range(P):
            ref = _stub_decode(lat[p_:p_ + 1], fr[p_:p_ + 1], t, out_size, None, 0)
            assert torch.allclose(full[c, p_], ref, atol=1e-5)
