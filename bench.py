#!/usr/bin/env python
"""bench.py -- NeuS training throughput (rays/s, ms/step) at base.json settings on N B200s.

  python bench.py --gpus 1 --steps K --warmup W                     # this framework (CUDA, sm_100a)
  python bench.py --impl reference --gpus 1 --steps K --warmup W    # reference arm: CPU port of the reference path
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...   # data-parallel rays + NCCL all-reduce

One "step" = one Testbed::train() call (src/testbed.cu:2640): occupancy refresh when due, ray generation + marching,
first forward, compositing/loss, fused second forward + backward (incl. eikonal second order), Adam+EMA.
Workload (config.workload): BASELINE.json configs[1], the DTU-Scan24-shaped synthetic scene (49 views, 1600x1200
RGBA, analytic-SDF images), base.json network, batch 2^18 samples per GPU.  Prints ONE JSON line.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SCHEDULE_STEPS = 2000   # configs/nerf/base.json:123 first_frame_max_training_step


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--preroll", type=int, default=700, help="untimed training steps before warm-up so that rays_per_batch and the occupancy grid are in steady state")
    ap.add_argument("--views", type=int, default=49)
    ap.add_argument("--width", type=int, default=1600)
    ap.add_argument("--height", type=int, default=1200)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-prefetch", action="store_true", help="A/B: the reference's in-order step (config.prefetch_next_samples = 0) instead of the pipelined one")
    ap.add_argument("--global-movement", action="store_true", help="secondary measurement: time the step of frames >= 1 of a dynamic sequence (DeltaNetwork in the network kernels, delta gradients, global-move Adam)")
    ap.add_argument("--dp", default="p2p", choices=["p2p", "nccl"], help="N>1: optimizer fused with the gradient reduction over NVLink peer memory (p2p) or NCCL all-reduce + replicated Adam")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return p.get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    def __init__(self, device_index):
        self.idx = device_index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for n, v in zip(names, out[2:]):
                    if v.strip().lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def make_data(args):
    from neus2_b200 import synthetic
    # NEUS2_BENCH_DATA_CACHE=<file.npz>: back-to-back A/B runs on one box render the synthetic views once (untimed set-up either way)
    cache = os.environ.get("NEUS2_BENCH_DATA_CACHE")
    key = f"{args.views}x{args.width}x{args.height}"
    if cache and os.path.exists(cache):
        z = np.load(cache)
        if str(z["key"]) == key:
            data = {k: z[k] for k in ("pixels", "focal", "principal", "xforms")}
            data["spheres"] = [(c, float(r)) for c, r in zip(z["sphere_centers"], z["sphere_radii"])]
            return data
    data = synthetic.dtu_like_views(n_views=args.views, width=args.width, height=args.height, seed=0)
    if cache:
        np.savez(cache, key=key, sphere_centers=np.array([c for c, _ in data["spheres"]]), sphere_radii=np.array([r for _, r in data["spheres"]]),
                 **{k: data[k] for k in ("pixels", "focal", "principal", "xforms")})
    return data


# ------------------------------------------------------------------------------------------------------------------
# CPU port of the reference path (oracle), timed on a bounded sample of the same workload
# ------------------------------------------------------------------------------------------------------------------
def cpu_port_rays_per_s(data, params_f32, bitfield, rays_per_batch, n_u, n_compacted, batch, n_sample_rays=1024):
    """Times oracle sample generation + forward + loss + backward + Adam on `n_sample_rays` rays of the workload and
    scales to rays/s.  Returns (rays_per_s, description)."""
    from oracle import oracle as O
    cfg = O.make_cfg()
    ds = O.make_dataset(data["pixels"], data["focal"], data["principal"], data["xforms"])
    ph = params_f32.astype(np.float16).view(np.uint16)
    R = n_sample_rays
    t0 = time.perf_counter()
    gen = O.generate_training_samples(R, 16 * batch, 0, O.Pcg32(seed=1337), ds, bitfield)
    n = gen["n_samples"]
    k = gen["n_kept"]
    if n == 0:
        return None, "no samples"
    coords = np.ascontiguousarray(gen["coords"][:n])
    out = O.net_forward(cfg, coords, ph)
    lp = O.LossParams()
    lp.loss_scale, lp.random_bg_color, lp.snap_to_pixel_centers, lp.mask_loss_weight, lp.ek_loss_weight, lp.cos_anneal_ratio = 128.0, 1, 1, 0.0, 0.01, 1.0
    loss = O.compute_loss(R, k, 0, O.Pcg32(seed=1337), max(n, 128), lp, ds, out, gen["ray_indices"], gen["rays"], gen["numsteps"], gen["coords"])
    m = max(1, min(loss["n_compacted"], n))
    O.net_backward(cfg, np.ascontiguousarray(loss["coords_compacted"][:m]), ph, np.ascontiguousarray(loss["dloss"][:m]), batch)
    t_net = time.perf_counter() - t0
    # Adam + EMA over all parameters is per step, not per ray
    lay = O.param_layout(cfg)
    N = lay["total"]
    g = np.zeros(N, np.float32)
    g[:: 7] = 1e-3
    mm, m1, m2, st, ema = params_f32.copy(), np.zeros(N, np.float32), np.zeros(N, np.float32), np.zeros(N, np.uint32), np.zeros(N, np.uint16)
    php = ph.copy()
    t1 = time.perf_counter()
    O.adam_ema_step(lay["grid"], 128.0, 1e-3, 0.9, 0.99, 1e-15, 1e-6, 0.95, 1, g, mm, php, m1, m2, st, ema)
    t_adam = time.perf_counter() - t1
    # one full step = rays_per_batch rays  (sample work scales with rays) + one optimizer pass
    t_step = t_net * (rays_per_batch / R) + t_adam
    desc = (f"oracle (CPU port) on {R} rays of the workload: {n} samples fwd, {m} samples bwd in {t_net:.2f}s + Adam/EMA over {N} params "
            f"{t_adam:.2f}s; scaled to {rays_per_batch} rays/step")
    return rays_per_batch / t_step, desc


def run_reference_arm(args):
    """Reference arm: the reference has no CPU path (CUDA-only, and its ngp layer does not compile here: SURVEY §0.3), so
    this times the CPU restatement of its path (oracle/) on the host cores for the same config/metric."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as O
    data = make_data(args)
    cfg = O.make_cfg()
    lay = O.param_layout(cfg)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from tests import util as U
    # The state the CPU path is timed on.  The metric is defined in training steady state (rays_per_batch adapted, occupancy grid
    # converged, surfaces sharp enough that rays saturate), which the CPU port cannot reach in the time allowed (~2.5 s per step), so
    # when a GPU is present the state -- parameters, occupancy bitfield, rays per step -- is produced by --preroll untimed steps of the
    # measured arm and copied to the host; nothing of it runs inside the timed region.  Without a GPU: geometric-init-like parameters
    # and the analytic surface shell (fewer samples per ray than the steady state, i.e. a FASTER reference).
    state_note = "analytic shell occupancy + geometric-init parameters (no GPU: not the steady state, fewer samples per ray)"
    p, bitfield, rays_per_step = None, None, 4096 * 4
    try:
        import torch
        if torch.cuda.is_available():
            import neus2_b200 as N
            ctx = N.Context(N.default_config(), device=0)
            ctx.init_params()
            ctx.set_dataset(data["pixels"], data["focal"], data["principal"], data["xforms"])
            for _ in range(args.preroll):
                st = ctx.train_step()
            torch.cuda.synchronize()
            p = ctx.get_params()
            bitfield = ctx.debug_buffer("bitfield", "<u1", (128 ** 3,)).cpu().numpy()
            rays_per_step = int(st.rays_per_batch_next)
            ctx.close()
            state_note = f"steady state after {args.preroll} untimed steps (parameters + occupancy bitfield copied from the GPU, {rays_per_step} rays/step)"
    except Exception as ex:   # fall back to the analytic state
        state_note += f" [GPU state unavailable: {ex}]"
    if p is None:
        p = U.random_params(cfg, seed=0, grid_scale=1e-4)
        bitfield = shell_bitfield(data)
    cores = os.cpu_count()
    vals = []
    desc = ""
    for it in range(args.warmup + args.steps):
        v, desc = cpu_port_rays_per_s(data, p, bitfield, rays_per_step, 0, 0, 1 << 18, n_sample_rays=512)
        if it >= args.warmup:
            vals.append(v)
    value = float(np.mean(vals))
    line = {
        "impl": "reference", "metric": "NeuS training rays/s (base.json)", "value": value, "unit": "rays/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * rays_per_step / value, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f16", "data": "synthetic",
        "config": workload_config(args, args.gpus),
        "cpu_baseline": {"value": value, "unit": "rays/s", "cores": cores, "kind": "port", "sample": desc + "; state: " + state_note},
        "e2e": {"value": value, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "the reference is CUDA-only (no CPU implementation of the path) and its application (src/*.cu) does not compile offline (Eigen/Core missing); this is the CPU port of its path, pinned to the reference's own kernels (DESIGN.md, Oracle)",
    }
    print(json.dumps(line))


def shell_bitfield(data):
    """Occupancy bitfield of the analytic surface shell (|sdf| < 2 cells), all mips, via the oracle's bitfield builder."""
    from oracle import oracle as O
    from neus2_b200 import synthetic
    G = 128
    idx = np.arange(G ** 3, dtype=np.uint32)
    # vectorised morton decode
    def dec(x):
        x = x & 0x49249249
        x = (x | (x >> 2)) & 0xc30c30c3
        x = (x | (x >> 4)) & 0x0f00f00f
        x = (x | (x >> 8)) & 0xff0000ff
        x = (x | (x >> 16)) & 0x0000ffff
        return x
    pts = np.stack([dec(idx), dec(idx >> 1), dec(idx >> 2)], axis=1).astype(np.float64)
    pts = (pts + 0.5) / G
    sdf = synthetic.sphere_sdf(pts, data["spheres"])
    grid = np.where(np.abs(sdf) < 2.5 / G, 1.0, 0.0).astype(np.float32)
    bitfield, _ = O.update_bitfield(grid, 1)
    return bitfield


def workload_config(args, n_gpus):
    return {
        "workload": f"DTU-Scan24-shaped synthetic scene ({args.views} views {args.width}x{args.height} RGBA from an analytic SDF), base.json "
                    f"(hashgrid L=14 F=2 T=2^19, 64-wide fp16 SDF+colour MLPs, eikonal 0.01), batch 2^18 samples per GPU",
        "batch_samples_per_gpu": 1 << 18, "global_batch_samples": (1 << 18) * n_gpus, "parallelism": (f"dp{n_gpus} (rays sharded; " + ("Adam fused with the gradient reduction over NVLink peer memory, 1/N of the optimizer state per GPU)" if args.dp == "p2p" else "NCCL all-reduce of fp32 grads + replicated Adam)")) if n_gpus > 1 else "single GPU",
        "l2_policy": "per-step working set (>=470 MB of sample/activation buffers) exceeds the 126 MB L2; no explicit flush",
        "preroll_steps": args.preroll,
        **({"global_movement": "DeltaNetwork on (frames >= 1 of a dynamic sequence)"} if args.global_movement else {}),
        "step_order": "in-order (prefetch_next_samples=0)" if args.no_prefetch else "pipelined: the next step's sample generation overlaps this step's backward tail, weight gradients and optimizer; no host sync at the end of a step",
    }


def main():
    args = parse()
    if args.impl == "reference":
        # all host cores: torchrun exports OMP_NUM_THREADS=1 to its workers, which would time the OpenMP port on ONE core
        # (round 1's N>1 reference lines were 10x slower than the N=1 line for exactly that reason); set before libgomp is loaded
        os.environ["OMP_NUM_THREADS"] = str(os.cpu_count())
        os.environ.pop("OMP_PROC_BIND", None)
        run_reference_arm(args)
        return
    import torch
    import torch.distributed as dist
    import neus2_b200 as N

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    else:
        torch.cuda.set_device(0)
    dev = torch.device(f"cuda:{local_rank}")
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}"

    data = make_data(args)
    B = (1 << 18) * world                      # weak scaling: every GPU back-propagates 2^18 samples per step
    cfg = N.default_config(batch_size=B, world_size=world, max_rays_per_batch=(1 << 18) * world, prefetch_next_samples=0 if args.no_prefetch else 1)
    ctx = N.Context(cfg, device=local_rank)
    ctx.init_params()
    # pinned host copy of the dataset: the upload is part of the end-to-end number
    px_pinned = torch.from_numpy(data["pixels"]).pin_memory()
    torch.cuda.synchronize()
    t_up0 = time.perf_counter()
    ctx.set_dataset(px_pinned.numpy(), data["focal"], data["principal"], data["xforms"])
    torch.cuda.synchronize()
    t_upload = time.perf_counter() - t_up0
    dataset_bytes = int(data["pixels"].nbytes + data["focal"].nbytes + data["principal"].nbytes + data["xforms"].nbytes)

    grads = ctx.buffers()[3]
    stream = torch.cuda.current_stream()

    from neus2_b200 import parallel
    if world > 1 and args.dp == "p2p":
        parallel.attach_peers(ctx, rank, world)

    def step():
        return parallel.train_step(ctx, rank, world)   # world > 1: NCCL all-reduce of the 10.56 M fp32 gradients over NVLink

    for _ in range(args.preroll):
        st = step()
    if args.global_movement:
        ctx.begin_global_movement()                # m_use_delta = m_train_delta = true from here on
    for _ in range(args.warmup):
        st = step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()

    launches0 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    rays = 0
    n_u = n_c = n_fp = 0
    t_wall0 = time.perf_counter()
    with ClockSampler(local_rank) as clocks:
        e0.record(stream)
        for _ in range(args.steps):
            st = step()
            rays += st.rays_per_batch
            n_u += st.measured_batch_size_before_compaction
            n_c += st.measured_batch_size
            n_fp += st.first_pass_samples
        e1.record(stream)
        torch.cuda.synchronize()
        t_wall = time.perf_counter() - t_wall0
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    launches = ctx.launch_count() - launches0
    # per-stage CUDA-event timers: a second, untimed pass over as many steps (the ~16 event records per step stay out of the headline number)
    ctx.set_profiling(True)
    for _ in range(args.steps):
        step()
    torch.cuda.synchronize()
    stages = ctx.stage_times()
    ctx.set_profiling(False)
    stages_all = None
    if world > 1:
        stages_all = [None] * world
        dist.all_gather_object(stages_all, {k: round(v[0], 4) for k, v in stages.items()})
        dist.barrier()

    if rank == 0:
        K = args.steps
        value = rays / (ms * 1e-3)
        # a number taken on a mis-sized batch is invalid: rays_per_batch is adapted so that the compacted batch tracks 2^18 samples per GPU
        fill = (n_c / K) / B
        if not 0.96 < fill < 1.04:
            raise SystemExit(f"bench invalid: compacted batch is {fill:.2f} of its target over the timed steps")
        # end to end through the public API (Context.train_step): wall clock of the K steps incl. the per-step device->host
        # read of the counters/loss, plus the host->device upload of the training images from pinned memory, amortised over
        # the 2000-step schedule the metric is defined on.
        t_e2e = t_wall + t_upload * (K / SCHEDULE_STEPS)
        e2e = rays / t_e2e
        peak, peak_src = peaks()
        # dominant kernel = fused forward+backward over the compacted batch
        dom = "fwdbwd"
        dom_ms = stages[dom][0]
        rows = (1 << 18)
        alg_bytes = rows * (508 + 1376)        # 508 B/sample forward + ~1.4 KB/sample backward (algorithmic minimum)
        achieved = alg_bytes / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
        # the hash-grid tables (21 MB fp16) and their gradients (42 MB fp32) live in L2, so the honest ceiling for the gather /
        # scatter part is what L2 sustains for those access patterns (tools/l2_microbench.cu, measured on this pool's B200)
        l2 = {}
        try:
            with open(os.path.join(ROOT, "profiles", "r01_l2_microbench.json")) as f:
                mb = json.load(f)
            gathers_fwd = (n_fp / K) * 14 * 8      # samples the first pass actually evaluated (it stops where the loss stops), not all marched samples
            gathers_bwd, reds_bwd = rows * 14 * 8, rows * 14 * 8
            floor_bwd_ms = (gathers_bwd / mb["gather4B_hash_corners_Gops"] + reds_bwd / mb["red_v2f32_random_Gops"]) * 1e-6
            l2 = {"forward_gather_Gops": gathers_fwd / (stages["forward"][0] * 1e6), "gather_ceiling_Gops": mb["gather4B_hash_corners_Gops"],
                  "gather_ceiling_random_Gops": mb["gather4B_random_Gops"], "forward_frac_of_ceiling": gathers_fwd / (stages["forward"][0] * 1e6) / mb["gather4B_hash_corners_Gops"],
                  "fwdbwd_l2_floor_ms": floor_bwd_ms, "fwdbwd_frac_of_l2_floor": floor_bwd_ms / dom_ms if dom_ms > 0 else None,
                  "source": "profiles/r01_l2_microbench.json (random 4-byte gathers over 16 MiB / 8-byte vector reductions over 32 MiB, both L2-resident)"}
        except Exception:
            pass
        line = {
            "metric": "NeuS training rays/s (base.json)", "value": value, "unit": "rays/s", "n_gpus": world, "steps": K, "warmup": args.warmup,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16", "data": "synthetic",
            "config": workload_config(args, world),
            "clocks": clocks.summary(),
            "e2e": {"value": e2e, "unit": "rays/s", "h2d_bytes_per_step": dataset_bytes / SCHEDULE_STEPS, "d2h_bytes_per_step": 28,
                    "note": "wall clock through Context.train_step(); dataset upload from pinned host memory amortised over the 2000-step schedule"},
            "gpu_launches": int(launches),
            "roofline": {"kernel": "k_network_fwdbwd<0, 1> (fused second forward + backward + tcgen05 weight gradients)", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": 70.15e6, "traffic_source": "profiles/r02_captureC.md: dram__bytes_read 66.6 MB + dram__bytes_write 3.5 MB per launch (ncu --set full); far below the algorithmic bytes because the hash tables and their gradients are served by L2 (see l2)",
                         "peak_source": peak_src, "alg_bytes_per_launch": alg_bytes, "avg_launch_ms": dom_ms, "l2": l2},
            "stage_ms": {k: round(v[0], 4) for k, v in stages.items()},
            **({"stage_ms_per_rank": stages_all} if stages_all else {}),
            "per_step": {"rays": rays / K, "samples_before_compaction": n_u / K, "samples_compacted": n_c / K, "first_pass_samples_evaluated": n_fp / K, "batch_fill": fill, "training_step": int(st.training_step),
                         "first_pass_samples_evaluated_last_step": int(st.first_pass_samples)},
        }
        if not args.no_cpu_baseline and world == 1:      # reported on rank 0 at N = 1 only
            try:
                bitfield = ctx.debug_buffer("bitfield", "<u1", (128 ** 3,)).cpu().numpy()
                v, desc = cpu_port_rays_per_s(data, ctx.get_params(), bitfield, int(rays / K), n_u / K, n_c / K, 1 << 18, n_sample_rays=512)
                line["cpu_baseline"] = {"value": v, "unit": "rays/s", "cores": os.cpu_count(), "kind": "port", "sample": desc}
            except Exception as ex:   # the baseline is a report, it must not take the bench line down
                line["cpu_baseline"] = {"value": None, "unit": "rays/s", "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {ex}"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
