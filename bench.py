#!/usr/bin/env python
"""bench.py -- decoded space-time queries/s of the STIF query decoder on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--mode bf16|fp32] [--impl ours|reference]

A *step* is one pass of the hot path over one batch of synthetic input: BASELINE.json config 2
(x4 spatial / x2 temporal decode of a 480x270 latent to 1080p: B=1, H=270, W=480 -> 1080x1920,
times [0, 0.5]; 4 147 200 queries = K0 latent projection + 2 x (K1 stage A+B, K2 stage C+D+E)).
At N>1 every rank decodes its own frame pair of that shape (the path shards by (pair, t) slabs,
SURVEY.md section 8e): weak scaling, no collective inside the decode loop; the MLP weights are
broadcast once from rank 0 over NCCL before the timed region.

One JSON line is printed by rank 0 (see the keys at the bottom).  Nothing here reads /root/reference.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "stif-continuous-video-representation_b200")
for _p in (ROOT, PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402

FLOP_PER_QUERY = 416_896            # 2 x 208 448 MAC, unpadded
FLOP_K1 = 2 * (49_728 + 38_336)     # feat_imnet + flow_imnet  (stage A+B)
FLOP_K2 = 2 * 120_384               # encode_imnet              (stage E)
WORKLOADS = {
    # name: (H, W, HH, WW, times)
    "config2": (270, 480, 1080, 1920, [0.0, 0.5]),
    "config1": (64, 64, 256, 256, [i / 8.0 for i in range(8)]),
    "config4_slab": (540, 960, 2160, 3840, [0.375]),
}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return float(d["bf16_tflops"]), float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), "measured"
    return 1590.0, 1400.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def mark(self):
        return time.time()

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self, t0: float, t1: float) -> dict:
        sm, smax, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = [r for (ts, r) in self.rows if t0 - 0.05 <= ts <= t1 + 0.15] or [r for (_, r) in self.rows]
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            try:
                sm.append(float(f[1]))
                smax = max(smax, float(f[2]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax or None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------- CPU reference arm
CPU_SAMPLES = {
    # name: (H, W, (HH, WW), times, description) -- per-query work and scale factor are those of config 2
    "quarter": (135, 240, (540, 960), [0.5],
                "config2 quarter-area crop: 135x240 latent -> 540x960, t=[0.5], 518400 queries"),
    "full": (270, 480, (1080, 1920), [0.0, 0.5],
             "the whole config-2 step: 270x480 latent -> 1080x1920, t=[0.0, 0.5], 4147200 queries"),
}


def run_cpu_reference(steps: int, warmup: int, sample: str = "quarter"):
    """The reference's OWN implementation on the host cores: the unmodified `LunaTokis.decoding`
    (Sakuya_arch_test.py:364-459) imported from the staged tree oracle/_ref (oracle/stage_ref.py), fp32, all host threads.
    Falls back to the torch-CPU port (oracle/port_torch.py, kind "port") only if the staged tree is absent.
    The caller must have hidden the GPUs (CUDA_VISIBLE_DEVICES="") before torch was imported: the method hard-codes
    `.cuda()` (:372-375) and warplayer.py:5 picks its device at import."""
    import torch
    from oracle import ref_loader, synth
    assert not torch.cuda.is_available(), "the CPU reference arm must run with the GPUs hidden"
    torch.set_num_threads(os.cpu_count() or 1)
    H, W, scale, times, what = CPU_SAMPLES[sample]
    w = synth.make_weights(0, False)
    lat, fr = synth.make_inputs(0, 1, H, W, 0.05)
    nq = scale[0] * scale[1] * len(times)
    if ref_loader.reference_available():
        model = ref_loader.build_reference_model(w)
        kind = "reference"
        src = f"unmodified LunaTokis.decoding from {'oracle/_ref' if ref_loader.reference_kind() == 'staged' else ref_loader.REFERENCE_ROOT}"

        def step():
            ref_loader.reference_decode(lat, fr, w, times, scale, device="cpu", model=model)
    else:
        from oracle import port_torch
        kind, src = "port", "oracle/port_torch.py (staged reference absent)"

        def step():
            port_torch.decode(lat, fr, w, times, scale)
    for _ in range(warmup):
        step()
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        step()
        ts.append(time.perf_counter() - t0)
    sec = float(np.mean(ts))
    return {"value": nq / sec, "unit": "queries/s", "cores": torch.get_num_threads(), "kind": kind, "source": src,
            "sample": what, "sec_per_step": sec, "best_sec": float(np.min(ts))}


def workload_config(args, world: int) -> dict:
    """The `config` object of the JSON line -- identical for both arms (`--impl ours` / `--impl reference`)."""
    H, W, HH, WW, times = WORKLOADS[args.workload]
    return {"workload": f"{args.workload}: B=1 {H}x{W} latent -> {HH}x{WW}, times {times} "
                        f"({HH * WW * len(times)} queries per rank per step); one frame pair per rank",
            "weights": "SIREN-init seed 0" + (" stress variant" if args.stress_weights else ""),
            "l2": "no flush: per-step working set (latent 100 MB + tables + 50 MB output) exceeds the 126 MB L2"}


def main_reference(args):
    """`--impl reference`: the reference's own CPU implementation of the path on the box's host cores (oracle/_ref),
    each step a bounded sample of the workload (`cpu_baseline.sample`).  Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    os.environ["CUDA_VISIBLE_DEVICES"] = ""            # before torch is imported anywhere in this process
    res = run_cpu_reference(max(1, args.steps), max(0, min(args.warmup, 1)), args.cpu_sample)
    line = {"impl": "reference", "metric": "decoded space-time queries/sec", "value": res["value"], "unit": "queries/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["sec_per_step"] * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, 1),
            "cpu_baseline": {"value": res["value"], "unit": "queries/s", "cores": res["cores"], "kind": res["kind"],
                             "sample": res["sample"], "source": res["source"]},
            "e2e": {"value": res["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


def cpu_baseline_subprocess(sample: str, steps: int):
    """`cpu_baseline` leg of our arm: the reference arm in a child process with the GPUs hidden (this process holds a
    CUDA context and the reference picks its device at import)."""
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        env.pop(k, None)
    p = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", str(steps), "--warmup", "0",
                        "--cpu-sample", sample], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=900)
    for ln in reversed(p.stdout.strip().splitlines()):
        if ln.startswith("{"):
            return json.loads(ln)["cpu_baseline"]
    raise RuntimeError(f"CPU reference arm failed: {p.stderr[-2000:]}")


# --------------------------------------------------------------------------- sharded (strong-scaling) leg
SHARDED_JOBS = {
    # BASELINE.json config 4: 4K output, x8 temporal; 8 (t) slabs -> one per GPU at N=8, round-robin below
    "config4": (1, 540, 960, (2160, 3840), [i / 8.0 for i in range(8)]),
    # config 2 split across the GPUs: 2 slabs -> whole slabs at N<=2, row bands with a locally recomputed halo above
    "config2": (1, 270, 480, (1080, 1920), [0.0, 0.5]),
}


def sharded_leg(job: str, steps: int, world: int, rank: int, weights, mode: str, maxreduce, barrier) -> dict:
    """The north-star multi-GPU split: ONE job whose queries are partitioned over the N ranks by the query-sharding
    launcher (stif_b200/launcher.py) -- latent + frames + weights broadcast once from rank 0 over NCCL (timed apart,
    `broadcast_s`), then every rank decodes its (pair, t) slabs / row bands with no collective inside the loop.  Total work
    is fixed, so this is STRONG scaling.  The checksum is an exact integer sum over the fp32 bit patterns of every
    unit's rows: identical at every N iff the sharded result is bit-identical to the single-GPU decode."""
    import torch
    import torch.distributed as dist
    from stif_b200 import synthetic as synth
    from stif_b200.launcher import QueryShardLauncher, plan_units

    P, H, W, out_size, times = SHARDED_JOBS[job]
    launcher = QueryShardLauncher(mode=mode)
    lat = fr = None
    if rank == 0:
        lat_np, fr_np = synth.make_inputs(200, P, H, W, 0.05)
        lat, fr = torch.from_numpy(lat_np).cuda(), torch.from_numpy(fr_np).cuda()
    barrier()
    tb = time.perf_counter()
    launcher.broadcast_weights(weights if rank == 0 else None)
    launcher.broadcast_inputs(lat, fr, (P, H, W))
    torch.cuda.synchronize()
    t_bcast = maxreduce(time.perf_counter() - tb)
    for _ in range(2):
        results = launcher.decode(times, out_size, halo=16)
    del results
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(steps):
        results = launcher.decode(times, out_size, halo=16)
    ev1.record()
    barrier()
    ms = maxreduce(ev0.elapsed_time(ev1) / steps)
    local_sum = 0
    for u, t in results:
        local_sum += int(t[:, u.row_begin:u.row_end].contiguous().view(torch.int32).to(torch.int64).sum().item())
    tot = torch.tensor([local_sum], device="cuda", dtype=torch.int64)
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    units = plan_units(P, len(times), out_size[0], world)
    nq = P * len(times) * out_size[0] * out_size[1]
    kind = "slabs" if all(u.row_begin == 0 and u.row_end == out_size[0] for u in units) else "row bands + local halo"
    launcher._decoder.close()
    del launcher, results
    torch.cuda.empty_cache()
    return {"workload": f"{job}: {P} pair(s) {H}x{W} latent -> {out_size[0]}x{out_size[1]}, {len(times)} timesteps = {nq} queries in total",
            "scaling": "strong", "value": nq / (ms * 1e-3), "unit": "queries/s", "ms_per_step": ms, "units": len(units),
            "partition": kind, "broadcast_s": t_bcast, "collectives_in_decode_loop": 0,
            "checksum_i64": int(tot.item())}


# --------------------------------------------------------------------------- our arm
def main_ours(args):
    import torch
    import torch.distributed as dist

    import stif_b200
    from stif_b200 import synthetic as synth  # seeded numpy generators (product-side helper; nothing from oracle/ here)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    H, W, HH, WW, times = WORKLOADS[args.workload]
    T = len(times)
    nq_rank = HH * WW * T

    # weights: created on rank 0, broadcast once over NCCL (0.84 MB); latents: one frame pair per rank
    keys = stif_b200.weight_keys()
    shapes = synth.weight_shapes()
    flat = torch.zeros(sum(int(np.prod(shapes[k])) for k in keys), device="cuda")
    if rank == 0:
        w0 = synth.make_weights(0, args.stress_weights)
        flat.copy_(torch.from_numpy(np.concatenate([w0[k].ravel() for k in keys])))
    t_bcast = 0.0
    if world > 1:
        torch.cuda.synchronize()
        tb = time.perf_counter()
        dist.broadcast(flat, 0)
        torch.cuda.synchronize()
        t_bcast = time.perf_counter() - tb
    weights, off = {}, 0
    flat_h = flat.cpu().numpy()
    for k in keys:
        n = int(np.prod(shapes[k]))
        weights[k] = flat_h[off:off + n].reshape(shapes[k]).copy()
        off += n
    lat_h, fr_h = synth.make_inputs(100 + rank, 1, H, W, 0.05)
    dec = stif_b200.STIFQueryDecoder(local, mode=args.mode)
    dec.load_weights(weights)
    lat = torch.from_numpy(lat_h).cuda()
    fr = torch.from_numpy(fr_h).cuda()
    out = torch.empty((T, 1, 3, HH, WW), device="cuda")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def maxreduce(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        dec.decode_stacked(lat, fr, times, (HH, WW), out=out)
    barrier()
    # ---- timed region 1: inputs resident in HBM (value)
    dec.profile(True)
    dec.profile_read()
    l0 = dec.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = sampler.mark()
    ev0.record()
    for _ in range(args.steps):
        dec.decode_stacked(lat, fr, times, (HH, WW), out=out)
    ev1.record()
    barrier()
    w1 = sampler.mark()
    ms_step = maxreduce(ev0.elapsed_time(ev1) / args.steps)
    launches = dec.launch_count - l0
    prof = dec.profile_read()
    dec.profile(False)
    # ---- timed region 2: end to end through the C-ABI host entry point (pinned host buffers, H2D + D2H inside)
    lat_p, fr_p = torch.from_numpy(lat_h).pin_memory(), torch.from_numpy(fr_h).pin_memory()
    out_p = torch.empty((T, 1, 3, HH, WW)).pin_memory()
    for _ in range(2):
        dec.decode_host(lat_p, fr_p, times, (HH, WW), out=out_p)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        dec.decode_host(lat_p, fr_p, times, (HH, WW), out=out_p)   # synchronous: returns after the D2H copy
    e2e_sec = maxreduce((time.perf_counter() - t0) / args.steps)
    barrier()
    checksum = float(out_p.double().abs().mean())
    # ---- extra (not the headline): same call with STIF_FLAG_OUT_U8, i.e. the uint8 HWC frames the reference's caller
    # saves (custom_video_test.py:102) converted on the device, a quarter of the download
    out8_p = torch.empty((T, 1, HH, WW, 3), dtype=torch.uint8).pin_memory()
    for _ in range(2):
        dec.decode_host(lat_p, fr_p, times, (HH, WW), out=out8_p, uint8=True)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        dec.decode_host(lat_p, fr_p, times, (HH, WW), out=out8_p, uint8=True)
    e2e8_sec = maxreduce((time.perf_counter() - t0) / args.steps)
    barrier()
    # ---- extra: the byte-minimal host call -- bf16 latents in (stif_decode_host_bf16; the projection rounds to bf16 anyway, so the
    # result is bit-identical), uint8 frames out: 54 MB up + 12 MB down instead of 103 + 50
    lat16_p = torch.from_numpy(lat_h).to(torch.bfloat16).pin_memory()
    for _ in range(2):
        dec.decode_host(lat16_p, fr_p, times, (HH, WW), out=out8_p, uint8=True)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        dec.decode_host(lat16_p, fr_p, times, (HH, WW), out=out8_p, uint8=True)
    e2e16_sec = maxreduce((time.perf_counter() - t0) / args.steps)
    barrier()
    if rank == 0:
        sampler.stop()
    # ---- the same decoder behind the query-sharding launcher: ONE job split over the N GPUs (strong scaling)
    dec.close()
    del lat, fr, out
    torch.cuda.empty_cache()
    sharded = {}
    if not args.no_sharded:
        for job in ("config4", "config2"):
            sharded[job] = sharded_leg(job, max(3, args.steps // 2), world, rank, weights, args.mode, maxreduce, barrier)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak_burst, peak_sust, peak_src = measured_peaks()
    qps = world * nq_rank / (ms_step * 1e-3)
    # dominant kernel group and its roofline (tensor pipe): algorithmic FLOP per launch / mean launch duration
    groups = {"K0_project_latent": 0, "K1_stageAB_feat_flow": 1, "K2_stageCDE_warp_encode": 2}
    per = {}
    for name, k in groups.items():
        if prof["count"][k]:
            per[name] = prof["ms"][k] / prof["count"][k]
    dom = max((n for n in per if not n.startswith("K0")), key=lambda n: per[n])
    # queries one launch of the dominant kernel processes: the timesteps of a pair share launches (stif_workspace_bytes_grouped)
    q_launch = args.steps * T * HH * WW / prof["count"][groups[dom]]
    flop_launch = (FLOP_K1 if dom.startswith("K1") else FLOP_K2) * q_launch
    achieved = flop_launch / (per[dom] * 1e-3) / 1e12
    kernel_ms = sum(prof["ms"]) / args.steps
    traffic = None      # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu capture (config 2 only)
    tpath = os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")
    if args.workload == "config2" and os.path.isfile(tpath):
        t = json.load(open(tpath)).get(dom)
        if t:
            # the capture's launches cover t["queries_per_launch"] queries (one slab if absent); scale to this run's launches
            traffic = (t["dram_read_bytes"] + t["dram_write_bytes"]) * q_launch / t.get("queries_per_launch", HH * WW)
    roof = {"bound": "tensor", "kernel": dom, "achieved": achieved, "peak": peak_burst, "unit": "TFLOP/s",
            "frac": achieved / peak_burst, "peak_source": f"{peak_src} (burst; sustained {peak_sust})",
            "traffic": traffic, "ms_per_launch": per[dom], "queries_per_launch": q_launch,
            "all_kernels_ms_per_launch": per,
            "share_of_step": {n: per[n] * prof["count"][groups[n]] / args.steps / kernel_ms for n in per},
            "whole_step_frac": qps / world * FLOP_PER_QUERY / 1e12 / peak_burst}
    line = {"metric": "decoded space-time queries/sec", "value": qps, "unit": "queries/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16" if args.mode == "bf16" else "f32",
            "data": "synthetic",
            "config": workload_config(args, world),
            "setup": {"weight_broadcast_s": t_bcast, "mode": args.mode},
            "roofline": roof,
            "e2e": {"value": world * nq_rank / e2e_sec, "unit": "queries/s",
                    "h2d_bytes_per_step": int(lat_p.numel() * 4 + fr_p.numel() * 4),
                    "d2h_bytes_per_step": int(out_p.numel() * 4), "ms_per_step": e2e_sec * 1e3,
                    "api": "stif_decode_host (C ABI) on pinned host buffers"},
            "e2e_uint8": {"value": world * nq_rank / e2e8_sec, "unit": "queries/s", "d2h_bytes_per_step": int(out8_p.numel()),
                          "ms_per_step": e2e8_sec * 1e3, "api": "stif_decode_host with STIF_FLAG_OUT_U8 (custom_video_test.py:102 conversion on device)"},
            "e2e_bf16_latent_uint8": {"value": world * nq_rank / e2e16_sec, "unit": "queries/s", "ms_per_step": e2e16_sec * 1e3,
                                      "h2d_bytes_per_step": int(lat16_p.numel() * 2 + fr_p.numel() * 4), "d2h_bytes_per_step": int(out8_p.numel()),
                                      "api": "stif_decode_host_bf16 + STIF_FLAG_OUT_U8 (opt-in; bit-identical frames)"},
            "sharded": sharded,
            "gpu_launches": int(launches),
            "clocks": sampler.summary(w0, w1),
            "output_checksum": checksum}
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline_subprocess(args.cpu_sample if args.cpu_sample != "quarter" else "full", args.cpu_steps)
    emit(line)
    if world > 1:
        dist.destroy_process_group()


_RESULT_FD = None


def emit(line: dict) -> None:
    """The ONE line of this run's stdout (libraries that chat on fd 1 -- NCCL's version banner -- were sent to stderr)."""
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


def main():
    global _RESULT_FD
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)      # keep the real stdout for the result line ...
    os.dup2(2, 1)               # ... and point fd 1 at stderr for everything else in this process and its libraries
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--workload", default="config2", choices=sorted(WORKLOADS))
    ap.add_argument("--stress-weights", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sharded", action="store_true", help="skip the strong-scaling legs (config 4 / config 2 split over the ranks)")
    ap.add_argument("--cpu-steps", type=int, default=1)
    ap.add_argument("--cpu-sample", default="quarter", choices=sorted(CPU_SAMPLES),
                    help="bounded sample one CPU step covers (reference arm; our arm's cpu_baseline leg times 'full' once)")
    args = ap.parse_args()
    if args.impl == "reference":
        main_reference(args)
    else:
        main_ours(args)


if __name__ == "__main__":
    main()
