#!/usr/bin/env python
"""bench.py -- outer ICP iterations / second (correspondence + LM) on BASELINE.json's config.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 3]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

One "step" = one pass of the loop body main_multiview.cpp:150-169 minus rendering: correspondences of every frame
(Frame::computeClosestPointsToNeighbours) + one full LM solve (ceresOptimizer_sophusSE3).  The K timed steps are the
first K rounds of the real ICP trajectory from the seeded noisy poses (the reference hard-codes 20); the W warm-up
steps run the same rounds beforehand and the poses are then reset, so warm-up does not change the timed work.
Workload = BASELINE.json configs[2] (the config the metric is quoted on): 20 views x 200k pts, point-to-plane,
Sophus SE(3), robust, cutoff 0.05, knn 2, synthetic bunny-shaped scans (mv_lm_icp_b200/synth.py).
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

CONFIGS = {   # BASELINE.json configs[] (index = config id); knn = 2 everywhere (main_multiview.cpp:41)
    2: dict(views=10, points=100_000, param="aa", cost="p2plane"),
    3: dict(views=20, points=200_000, param="se3", cost="p2plane"),
    4: dict(views=40, points=500_000, param="quat", cost="mixed"),
    5: dict(views=64, points=1_000_000, param="se3", cost="p2plane"),
    0: dict(views=6, points=20_000, param="se3", cost="p2plane"),   # tiny, for dry runs
}
PARAM = {"aa": 0, "quat": 1, "se3": 2}
COST = {"p2p": 0, "p2plane": 1, "mixed": 2}
CUTOFF = 0.05


def load_scene(cfg_id, views, points):
    from mv_lm_icp_b200 import synth
    cache = f"/tmp/mvicp_scene_c{cfg_id}_{views}x{points}.npz"
    if os.path.exists(cache):
        try:
            z = np.load(cache)
            return {"pts": [z[f"p{i}"] for i in range(views)], "nor": [z[f"n{i}"] for i in range(views)],
                    "poses_gt": z["gt"], "poses_init": z["init"]}
        except Exception:
            pass
    sc = synth.make_scene(views, points, config_id=cfg_id)
    if int(os.environ.get("LOCAL_RANK", "0")) == 0:
        try:
            tmp = cache + f".{os.getpid()}.tmp.npz"
            np.savez(tmp, gt=sc["poses_gt"], init=sc["poses_init"], **{f"p{i}": p for i, p in enumerate(sc["pts"])},
                     **{f"n{i}": p for i, p in enumerate(sc["nor"])})
            os.replace(tmp, cache)
        except Exception:
            pass
    return sc


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (profiling guide's clocks line)."""

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.idx = gpu_index; self.samples = []; self.stop_flag = False; self.proc = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                if self.stop_flag:
                    break
                self.samples.append(line.strip())
        except Exception:
            pass

    def finish(self):
        self.stop_flag = True
        if self.proc:
            try:
                self.proc.terminate()
            except Exception:
                pass
        sm, mx, reasons = [], [], set()
        for s in self.samples:
            f = [x.strip() for x in s.split(",")]
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
                for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


def measured_traffic(kernel):
    """DRAM bytes per launch of a kernel from the committed `ncu --set full` capture (profiles/traffic.json), or None."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        return json.load(open(p)).get(kernel)
    except Exception:
        return None


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ======================================================================================================
# CPU arm: the reference path on the host cores (nanoflann verbatim from oracle/_ref when built, oracle LM port)
# ======================================================================================================
def host_threads():
    """OpenMP threads for the CPU arm: one per PHYSICAL core (hyper-thread siblings slow the static-schedule loops down:
    measured 27 s vs 4 s per LM round at 128 vs 64 threads on a 64-core box)."""
    from oracle import oracle as O
    n = O.max_threads()
    try:
        import psutil
        phys = psutil.cpu_count(logical=False)
        if phys:
            n = min(n, phys) if n > 0 else phys
    except Exception:
        pass
    return max(1, n)


def cpu_rounds(sc, views_sub, n_rounds, cfg, threads):
    """Runs n_rounds outer ICP rounds on the sub-problem made of the first `views_sub` views (ring edges among them),
    with every host thread.  Returns per-round seconds [(corr_s, lm_s)], number of edges, and which NN code ran."""
    from oracle import oracle as O
    from mv_lm_icp_b200 import synth
    full_edges = synth.ring_edges(cfg["views"], 2)
    edges = [(s, d) for s, d in full_edges if s < views_sub and d < views_sub and s != 0]
    kind = "ref" if O.ref_lib() is not None else "kd"
    pts = sc["pts"][:views_sub]; nor = sc["nor"][:views_sub]
    idx = {d: O.KdIndex(pts[d], kind) for d in set(d for _, d in edges)}   # one-time build: excluded, as the reference's lazily built index is amortised
    poses = sc["poses_init"][:views_sub].copy()
    times = []
    for _ in range(n_rounds):
        t0 = time.perf_counter()
        corr, w = [], []
        for s, d in edges:
            i, d2 = idx[d].closest_points(pts[s], poses[s], poses[d], threads=threads)
            f, sec, dist, ww, _ = O.filter_edge(i, d2, np.float32(CUTOFF))
            corr.append((f, sec)); w.append(ww)
        t1 = time.perf_counter()
        poses, summ, _ = O.optimize(pts, nor, poses, edges, corr, w, param=PARAM[cfg["param"]], cost=COST[cfg["cost"]], robust=True,
                                    se3_autodiff=True, threads=threads)
        t2 = time.perf_counter()
        times.append((t1 - t0, t2 - t1, summ["num_iterations"]))
    return times, len(edges), kind


def run_reference(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as O
    threads = host_threads()
    sc = load_scene(args.config, cfg["views"], cfg["points"])
    E_full = len([e for e in __import__("mv_lm_icp_b200").synth.ring_edges(cfg["views"], 2) if e[0] != 0])
    views_sub = 3   # frames 0,1,2 -> edges (1,0),(1,2),(2,1)
    if args.warmup > 0:
        cpu_rounds(sc, views_sub, min(args.warmup, 1), cfg, threads)      # page in code and data; CPU timings do not drift after that
    timed, E_sub, kind = cpu_rounds(sc, views_sub, args.steps, cfg, threads)   # the same rounds 0..K-1 the GPU arm times
    scale = E_full / E_sub
    per_round = float(np.mean([a + b for a, b, _ in timed])) * scale
    val = 1.0 / per_round
    sample = (f"rounds 0..{args.steps - 1} ({'after a warm-up round' if args.warmup > 0 else 'no warm-up'}) of the sub-problem views 0..{views_sub - 1} ({E_sub} of {E_full} directed "
              f"edges, {cfg['points']} queries each), extrapolated x{scale:.2f} by edge count; NN = "
              f"{'reference nanoflann.hpp (oracle/_ref)' if kind == 'ref' else 'oracle KD-tree port'}, LM = oracle port of the "
              f"Ceres path (Jet autodiff, dense Cholesky; Ceres not installable), {threads} OpenMP threads; index build excluded")
    out = {"impl": "reference", "metric": "ICP iterations/sec (corr+LM)", "value": val, "unit": "iter/s", "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_round * 1e3, "higher_is_better": True, "scaling": "strong",
           "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": f"multiview point-to-plane, {cfg['views']} views x {cfg['points']} pts, {cfg['param']} param, robust, knn 2, cutoff 0.05"},
           "cpu_baseline": {"value": val, "unit": "iter/s", "cores": threads, "kind": "reference" if kind == "ref" else "port", "sample": sample,
                            "corr_s_per_round": float(np.mean([a for a, _, _ in timed])) * scale,
                            "lm_s_per_round": float(np.mean([b for _, b, _ in timed])) * scale},
           "e2e": {"value": val, "unit": "iter/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(out), flush=True)


# ======================================================================================================
# GPU arm
# ======================================================================================================
def run_ours(args, cfg):
    import torch
    import mv_lm_icp_b200 as mv
    from mv_lm_icp_b200 import synth
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
    dev = local if world > 1 else 0
    torch.cuda.set_device(dev)

    sc = load_scene(args.config, cfg["views"], cfg["points"])
    M, N = cfg["views"], cfg["points"]
    edges = synth.ring_edges(M, 2)
    param, cost = PARAM[cfg["param"]], COST[cfg["cost"]]

    t_setup0 = time.perf_counter()
    eng = mv.Engine(device=dev, flags=args.flags)
    eng.set_frames(sc["pts"], sc["nor"])
    eng.set_graph(edges)
    if world > 1:
        import torch.distributed as dist
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(mv.nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        uid = bytes(idt.cpu().numpy().tobytes())
        eng.comm_init(uid, rank, world)
    eng.sync()
    setup_s = time.perf_counter() - t_setup0
    # Frame::recomputeNormals (default-on in the reference, before round 0; not part of the metric): timed once, on a
    # scratch engine so that the benchmark itself keeps the uploaded fp32-exact normals
    normals_ms = None
    if rank == 0 and world == 1:
        e2 = mv.Engine(device=dev); e2.set_frames(sc["pts"], sc["nor"])
        e2.recompute_normals(10); _, normals_ms = e2.recompute_normals(10)
        e2.close()

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize(); eng.sync()

    stream = torch.cuda.ExternalStream(eng.stream(), device=dev)

    def run_rounds(k, e2e):
        """k rounds from the initial poses; per-round device ms (events on the engine's stream) and stats."""
        eng.set_graph(edges)             # forget the previous trajectory's matches: round 0 is a cold, unseeded search
        eng.set_poses(sc["poses_init"])
        per = []
        poses = sc["poses_init"]
        for r in range(k):
            ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            ev0.record(stream)
            if e2e:
                eng.set_poses(poses)        # host buffer -> device, through the C ABI
            s = eng.icp_round(CUTOFF, param, cost, True)
            if e2e:
                poses = eng.get_poses()     # device -> host
            ev1.record(stream)
            eng.sync()
            wall = time.perf_counter() - t0
            st = eng.stats()
            per.append(dict(ms=ev0.elapsed_time(ev1), wall_ms=wall * 1e3, lm_iters=s["num_iterations"], evals=s["num_evaluations"],
                            knn_ms=st["knn_ms"], select_ms=st["select_ms"], lm_eval_ms=st["lm_eval_ms"], lm_other_ms=st["lm_other_ms"],
                            corr=st["correspondences"], queries=st["queries"]))
        return per

    # warm-up (untimed), then EXACTLY K timed steps between barriers
    run_rounds(max(3, args.warmup), False)
    sampler = ClockSampler(dev); sampler.start(); time.sleep(0.3)
    barrier()
    l0 = eng.stats()["kernel_launches"]
    t0 = time.perf_counter()
    per = run_rounds(args.steps, False)
    barrier()
    wall_total = time.perf_counter() - t0
    l1 = eng.stats()["kernel_launches"]
    # e2e: same K rounds, poses cross the C ABI as host buffers every step
    barrier()
    t0 = time.perf_counter()
    per_e2e = run_rounds(args.steps, True)
    barrier()
    wall_e2e = time.perf_counter() - t0
    clocks = sampler.finish()

    dev_ms = sum(p["ms"] for p in per)
    tot = torch.tensor([dev_ms, wall_total * 1e3, wall_e2e * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(tot, op=dist.ReduceOp.MAX)
    dev_ms, wall_ms, e2e_ms = [float(x) for x in tot.cpu()]
    mine = {k: round(sum(p[k] for p in per) / args.steps, 4) for k in ("knn_ms", "select_ms", "lm_eval_ms", "lm_other_ms")}
    mine["queries"] = int(per[0]["queries"])
    per_rank = [mine]
    if world > 1:   # lm_other of a rank includes its wait for the slowest rank's matrices: the imbalance shows here
        per_rank = [None] * world
        dist.all_gather_object(per_rank, mine)
    if rank == 0:
        peak, peak_src = hbm_peak()
        K = args.steps
        # the step is host-driven (one sync per LM iteration), so the honest whole-step time is the wall clock between the barriers
        ms_per_step = wall_ms / K
        knn_ms = float(np.mean([p["knn_ms"] for p in per])); q = per[0]["queries"] * world if world > 1 else per[0]["queries"]
        q_local = per[0]["queries"]
        knn_bytes = 44.0 * q_local
        evals = sum(p["evals"] for p in per); lm_eval_ms = sum(p["lm_eval_ms"] for p in per)
        corr_mean = float(np.mean([p["corr"] for p in per]))
        lm_bytes = 52.0 * corr_mean if cost != 0 else 36.0 * corr_mean
        knn_gbs = knn_bytes / (knn_ms * 1e-3) / 1e9
        lm_gbs = lm_bytes / ((lm_eval_ms / max(1, evals)) * 1e-3) / 1e9
        share = {"knn": sum(p["knn_ms"] for p in per) / dev_ms, "select": sum(p["select_ms"] for p in per) / dev_ms,
                 "lm_eval": lm_eval_ms / dev_ms, "lm_other": sum(p["lm_other_ms"] for p in per) / dev_ms}
        dominant = "knn" if share["knn"] >= share["lm_eval"] else "lm_eval"
        roof_knn = {"kernel": "knn_kernel", "bound": "hbm", "achieved": knn_gbs, "peak": peak, "unit": "GB/s", "frac": knn_gbs / peak,
                    "traffic": measured_traffic("knn_kernel") if world == 1 else None, "algorithmic_bytes_per_launch": knn_bytes,
                    "avg_launch_ms": knn_ms, "peak_source": peak_src}
        roof_lm = {"kernel": "lm_eval_kernel", "bound": "hbm", "achieved": lm_gbs, "peak": peak, "unit": "GB/s", "frac": lm_gbs / peak,
                   "traffic": measured_traffic("lm_eval_kernel") if world == 1 else None, "algorithmic_bytes_per_launch": lm_bytes,
                   "avg_launch_ms": lm_eval_ms / max(1, evals), "peak_source": peak_src}
        pose_bytes = M * 16 * 8
        out = {"metric": "ICP iterations/sec (corr+LM)", "value": 1e3 / ms_per_step, "unit": "iter/s", "n_gpus": world, "steps": K,
               "warmup": max(3, args.warmup), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
               "dtype": "f64", "data": "synthetic",
               "config": {"workload": f"multiview point-to-plane, {M} views x {N} pts, {cfg['param']} param, robust, knn 2, cutoff 0.05 (BASELINE configs[2])"
                          if args.config == 3 else f"config {args.config}: {M} views x {N} pts, {cfg['param']}, {cfg['cost']}",
                          "l2": "no flush: resident working set (clouds + trees + match arrays) = %.0f MB > 126 MB L2" %
                                ((M * N * 48 + len(edges) * N * 12) / 1e6),
                          "parallelism": f"edges (frame -> neighbour query sets) sharded over {world} GPU(s) by query count, one process per GPU",
                          "timing": "wall clock between barriers (host-driven LM loop); device-event sum = %.3f ms/step" % (dev_ms / K),
                          "setup_ms_excluded": setup_s * 1e3,
                          "normals_ms_excluded": normals_ms,
                          "lm_iterations_per_round": [p["lm_iters"] for p in per],
                          "per_round_ms": [round(p["ms"], 3) for p in per],
                          "per_rank_ms_per_step": per_rank,
                          "engine_flags": args.flags,
                          "storage": "fp32 records (lossless), fp64 arithmetic"},
               "e2e": {"value": 1e3 / (e2e_ms / K), "unit": "iter/s", "h2d_bytes_per_step": pose_bytes, "d2h_bytes_per_step": pose_bytes,
                       "note": "per step: poses host->device, correspond+optimize, poses device->host through the C ABI; clouds uploaded once "
                               "(setup_ms_excluded), as the reference keeps its clouds and KD-trees across rounds",
                       "incl_setup_value": 1e3 / ((e2e_ms + setup_s * 1e3) / K)},
               "gpu_launches": int(l1 - l0), "clocks": clocks,
               "roofline": roof_knn if dominant == "knn" else roof_lm, "roofline_knn": roof_knn, "roofline_lm": roof_lm,
               "time_share": share}
        if world == 1 and not args.no_cpu:
            from oracle import oracle as O
            th = host_threads()
            tms, E_sub, kind = cpu_rounds(sc, 3, 2, cfg, th)
            E_full = len([e for e in edges if e[0] != 0])
            per_round = float(np.mean([a + b for a, b, _ in tms])) * E_full / E_sub
            out["cpu_baseline"] = {"value": 1.0 / per_round, "unit": "iter/s", "cores": th, "kind": "reference" if kind == "ref" else "port",
                                   "sample": f"2 rounds of the 3-view sub-problem ({E_sub} of {E_full} edges x {N} queries), extrapolated by edge count; "
                                             f"NN = {'reference nanoflann (oracle/_ref)' if kind == 'ref' else 'oracle KD port'}, LM = oracle port "
                                             f"(Ceres not installable), {th} threads"}
            # the reference itself is single-threaded: the faithful figure, on one edge, beside the all-core one
            t1, E1, _ = cpu_rounds(sc, 2, 1, cfg, 1)
            per_round1 = float(np.mean([a + b for a, b, _ in t1])) * E_full / E1
            out["cpu_baseline"]["single_thread"] = {"value": 1.0 / per_round1, "unit": "iter/s", "cores": 1,
                                                    "sample": f"round 0 of the 2-view sub-problem ({E1} of {E_full} edges x {N} queries), extrapolated by edge count"}
        print(json.dumps(out), flush=True)
    eng.close()
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=3)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--flags", type=int, default=0, help="MVICP_FLAG_* bits for the engine (A/B measurements)")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    if args.impl == "reference":
        run_reference(args, cfg)
    else:
        run_ours(args, cfg)


if __name__ == "__main__":
    main()
