#!/usr/bin/env python
"""bench.py -- ICP pair-iterations/s of the batched registration engine (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1], "C2"): synthetic 1024 cluster pairs x 512 points per GPU, ICP stage only,
exactly 20 iterations -- the reference call  iterative_closest_point(src, dst, thres=0.1, max_iterations=20,
relative_rmse_thr=-1.0)  (SURVEY.md section 8d).  One *step* = one such call over one batch.  Weak scaling: every rank
owns its own 1024-pair shard and the resulting 4x4 transforms are all-gathered (NCCL) inside the step.

Printed line (rank 0): value = pair-iterations/s with inputs resident in HBM; e2e = the same through the host-buffer
path (pinned host inputs, H2D + kernels + D2H of the transforms every step); roofline = algorithmic bytes of the
dominant kernel / its CUDA-event duration against MEASURED_PEAKS.json; cpu_baseline = the CPU oracle (the reference
algorithm restated, oracle/icp_oracle.py) on this box's host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

PAIRS_PER_GPU = 1024
POINTS = 512
ICP_ITERS = 20
THRES = 0.1
METRIC = "icp_pair_iterations_per_sec"
UNIT = "pair-iters/s"
L2_BYTES = 126 * 1024 * 1024


def algorithmic_bytes_per_pair_iter(n_s: int, n_d: int) -> int:
    """both clouds as stored fp32 (x,y,z,flag) rows + one 4x4 fp32 transform."""
    return 16 * (n_s + n_d) + 64


def load_ncu_traffic():
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture of the same workload
    (profiles/ncu_r2_metrics.json, written by tools/ncu_metrics.py from `ncu --set full`), or None."""
    path = os.path.join(ROOT, "profiles", "ncu_r2_metrics.json")
    try:
        with open(path) as f:
            m = json.load(f)
        return float(m["dram_bytes_read"]) + float(m["dram_bytes_write"])
    except Exception:
        return None


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


# ---------------------------------------------------------------------------------------------- clock sampling
class ClockSampler:
    """Polls SM clock / throttle reasons through NVML while the timed regions run."""

    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x10: "sync_boost"}

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.005)

    def start(self):
        if self.nv is not None and self._thread is None:
            self._stop.clear()
            self._thread = threading.Thread(target=self._loop, daemon=True)
            self._thread.start()

    def stop(self):
        if self._thread is not None:
            self._stop.set()
            self._thread.join()
            self._thread = None

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ---------------------------------------------------------------------------------------------- CPU oracle timing
def time_cpu_oracle(target_seconds: float = 12.0, max_pairs: int = PAIRS_PER_GPU, seed: int = 1234):
    """The reference algorithm restated on CPU (oracle/icp_oracle.py + OpenMP knn leaf), all host threads."""
    import torch
    from icp_flow_b200 import synth
    from oracle import icp_oracle as O

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    src, dst, _ = synth.make_pairs(max_pairs, POINTS, seed=seed, residual_only=True)
    a, c = torch.from_numpy(src), torch.from_numpy(dst)
    # calibrate on a small slice, then size the sample for ~target_seconds
    t0 = time.perf_counter()
    O.icp_loop(a[:32], c[:32], thres=THRES, max_iterations=ICP_ITERS, relative_rmse_thr=-1.0)
    dt = time.perf_counter() - t0
    rate = 32 * ICP_ITERS / dt
    pairs = int(min(max_pairs, max(32, rate * target_seconds / ICP_ITERS / 2)))
    reps, dt = 0, 0.0
    while dt < target_seconds and reps < 64:        # repeat the sample until ~target_seconds of CPU work
        t0 = time.perf_counter()
        O.icp_loop(a[:pairs], c[:pairs], thres=THRES, max_iterations=ICP_ITERS, relative_rmse_thr=-1.0)
        dt += time.perf_counter() - t0
        reps += 1
    return {"value": reps * pairs * ICP_ITERS / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{reps} x ({pairs} of {max_pairs} pairs x {POINTS} pts x {ICP_ITERS} iterations), {dt:.2f} s, "
                      f"oracle/icp_oracle.py icp_loop (torch CPU fp32 + OpenMP C knn leaf)"}, pairs, dt


def run_reference_arm(args):
    """--impl reference: the reference's CPU implementation of the path (restated oracle) on the host cores.  One step =
    the engine arm's step: the full batch of PAIRS_PER_GPU pairs x ICP_ITERS iterations (same `config`)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import torch
    from icp_flow_b200 import synth
    from oracle import icp_oracle as O

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    src, dst, _ = synth.make_pairs(PAIRS_PER_GPU, POINTS, seed=1234, residual_only=True)
    a, c = torch.from_numpy(src), torch.from_numpy(dst)
    run = lambda: O.icp_loop(a, c, thres=THRES, max_iterations=ICP_ITERS, relative_rmse_thr=-1.0)
    budget_s = 200.0                 # the whole run has to end within a few minutes on any host
    t_w = time.perf_counter()
    warm = 0
    for _ in range(max(1, args.warmup)):
        run()
        warm += 1
        if time.perf_counter() - t_w > 0.2 * budget_s:
            break
    per_step = (time.perf_counter() - t_w) / warm
    steps = max(1, min(args.steps, int(budget_s / max(per_step, 1e-9))))
    t0 = time.perf_counter()
    for _ in range(steps):
        run()
    dt = time.perf_counter() - t0
    value = PAIRS_PER_GPU * ICP_ITERS * steps / dt
    sample = (f"each step = the full batch: {PAIRS_PER_GPU} pairs x {POINTS} pts x {ICP_ITERS} iterations; {steps} timed steps "
              f"of the {args.steps} requested (CPU time budget {budget_s:.0f} s), {warm} warm-up; "
              "oracle/icp_oracle.py icp_loop (torch CPU fp32 + OpenMP C knn leaf)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": dt / steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


def workload_config(n_gpus: int):
    return {"workload": f"C2: synthetic {PAIRS_PER_GPU} cluster pairs x {POINTS} pts per GPU, ICP only, "
                        f"{ICP_ITERS} forced iterations (max_iterations={ICP_ITERS}, relative_rmse_thr=-1), "
                        f"thres_dist={THRES}, residual-only motion",
            "pairs_per_gpu": PAIRS_PER_GPU, "points": POINTS, "icp_iterations": ICP_ITERS,
            "parallelism": f"pairs sharded over {n_gpus} GPU(s), all-gather of 4x4 transforms"}


# ---------------------------------------------------------------------------------------------- secondary configs
def _events_ms(fn, reps, sync):
    import torch
    fn()
    sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    sync()
    return e0.elapsed_time(e1) / reps


def time_c5_shard(dev, pairs=4096, steps=20, ext_fn=None, after_step=None):
    """One GPU's share of C5 (32768 pairs x 512 points over 8 GPUs = 4096 pairs per GPU, 20 forced iterations): four
    waves of CTAs instead of the single wave of the C2 batch.  Rotating pool of 5 batches (336 MB > 2x L2)."""
    import torch
    from icp_flow_b200 import _lib, ops, synth
    rank = int(os.environ.get("RANK", "0"))
    pool = []
    for i in range(5):
        s, d, _ = synth.make_pairs(pairs, POINTS, seed=4321 + rank + 1000 * i, residual_only=True)
        pool.append((torch.from_numpy(s).to(dev), torch.from_numpy(d).to(dev)))
    prm = ops.make_params(thres=THRES, max_iterations=ICP_ITERS, relative_rmse_thr=-1.0, early_exit=False, batch_stop=True)
    ws = torch.empty(_lib.lib().icpf_workspace_bytes(pairs, POINTS, 0, 0, 0), device=dev, dtype=torch.uint8)
    out = [None]

    def one(i):
        out[0] = ops.icp_batch(*pool[i % 5], prm, out=out[0], workspace=ws, ext=ext_fn(i) if ext_fn else None)
        if after_step:
            after_step(i, out[0])

    for i in range(3):
        one(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        one(3 + i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    assert int(out[0].iterations.sum().item()) == pairs * ICP_ITERS
    return ms, out[0]


def time_secondary(dev):
    """Not the headline metric: the other single-GPU BASELINE configs measured in the same run.  C1 (the demo-frame
    plumbing case), C3 at its full size (4096 pairs x 1024 points, full hist_icp, 135 x 135 x 3 histogram) with per-stage
    times, its own roofline accounting and a CPU hist_icp baseline on a stated subsample, C4 (Waymo-shape frame pair through
    match_pcds + flow) and one GPU's share of C5 (4096 pairs x 512 points)."""
    import types
    import numpy as np
    import torch
    import icp_flow_b200 as E
    from icp_flow_b200 import ops, scan, synth
    from oracle import icp_oracle as O
    out = {}
    peak, _ = load_peaks()
    sync = torch.cuda.synchronize
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)

    # ---- C1: demo.npz frame pair, the 32 cluster pairs <= 256 points of tests/golden/c1_demo.npz (F = 2.0, demo.sh)
    try:
        g = dict(np.load(os.path.join(ROOT, "tests", "golden", "c1_demo.npz")))
        s1, d1 = torch.from_numpy(g["src"]).to(dev), torch.from_numpy(g["dst"]).to(dev)
        a1 = types.SimpleNamespace(thres_dist=THRES, translation_frame=2.0, chunk_size=50)
        ms = _events_ms(lambda: E.hist_icp(a1, s1, d1), 20, sync)
        t0 = time.perf_counter()
        O.hist_icp(torch.from_numpy(g["src"]), torch.from_numpy(g["dst"]), O.PathParams(thres_dist=THRES, translation_frame=2.0))
        cpu_s = time.perf_counter() - t0
        out["c1_demo"] = {"pairs": int(s1.shape[0]), "points": int(s1.shape[1]), "hist_icp_ms": ms,
                          "pairs_per_s": s1.shape[0] / ms * 1e3,
                          "cpu_baseline": {"ms": cpu_s * 1e3, "cores": cores, "kind": "port",
                                           "sample": "the same 32 pairs in full, oracle/icp_oracle.py hist_icp"}}
    except Exception as exc:
        out["c1_demo"] = {"error": f"{type(exc).__name__}: {exc}"}

    # ---- C3 at size
    P3, N3, F3 = 4096, 1024, 6.666
    s_np, d_np, _ = synth.make_pairs(P3, N3, seed=99, ragged=False, residual_only=False)
    s, d = torch.from_numpy(s_np).to(dev), torch.from_numpy(d_np).to(dev)
    args = types.SimpleNamespace(thres_dist=THRES, translation_frame=F3, chunk_size=50)
    init = ops.estimate_init_pose(args, s, d, auto_swap=True)
    ms_init = _events_ms(lambda: ops.estimate_init_pose(args, s, d, auto_swap=True), 3, sync)
    ms_apply = _events_ms(lambda: ops.apply_icp(args, s, d, init, auto_swap=True), 3, sync)
    ms_all = _events_ms(lambda: E.hist_icp(args, s, d), 3, sync)
    _, dbg = ops.hist_icp(args, s, d, return_debug=True)
    batch_iters = int(dbg["batch"].tolist()[0])
    b_hist = P3 * (16 * (N3 + N3) + 16)                      # one read of both clouds + one translation per pair
    tests = float(P3) * N3 * N3                                # range tests (candidate votes) of the vote stage
    cpu_pairs = 128
    t0 = time.perf_counter()
    O.hist_icp(torch.from_numpy(s_np[:cpu_pairs]), torch.from_numpy(d_np[:cpu_pairs]), O.PathParams(thres_dist=THRES, translation_frame=F3))
    cpu_s = time.perf_counter() - t0
    out["c3_hist_icp"] = {
        "pairs": P3, "points": N3, "bins": [135, 135, 3], "ms": ms_all, "pairs_per_s": P3 / ms_all * 1e3,
        "stages_ms": {"estimate_init_pose": ms_init, "apply_icp": ms_apply}, "batch_iterations": batch_iters,
        "roofline": {"bound": "hbm", "stage": "estimate_init_pose (hist_fused + hist_score kernels)",
                     "achieved": b_hist / (ms_init * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                     "frac": b_hist / (ms_init * 1e-3) / 1e9 / peak, "algorithmic_bytes": b_hist,
                     "real_bound": "shared-memory atomics / instruction issue, not HBM: the vote stage is n_s*n_d range tests per pair",
                     "range_tests_per_s": tests / (ms_init * 1e-3),
                     "issue_ceiling_range_tests_per_s": 148 * 128 * 1.965e9 / 8.0,
                     "ceiling_note": "fp32 lane-op issue peak (148 SMs x 128 lanes x 1.965 GHz) / ~8 lane-ops per range test (SURVEY 8d)"},
        "cpu_baseline": {"pairs_per_s": cpu_pairs / cpu_s, "cores": cores, "kind": "port",
                         "sample": f"the first {cpu_pairs} of the {P3} pairs, full hist_icp, {cpu_s:.1f} s, oracle/icp_oracle.py"},
    }
    # the reference's own CUDA vote kernel (hist_cuda_core.cuh, built unmodified for sm_100a into oracle/_ref/) as the
    # kernel-to-beat of the vote stage, on the first 512 pairs (its dense [B,135,135,3] fp32 output is 219 KB per pair)
    try:
        from oracle import ref_hist
        if ref_hist.available():
            nb = 512
            hb = ops._hist_bins(THRES, F3, dev)
            sb, db = s[:nb].contiguous(), d[:nb].contiguous()
            buf = torch.empty(nb, *hb.lens, device=dev, dtype=torch.float32)
            ms_ref = _events_ms(lambda: ref_hist.hist(db, sb, list(hb.c.min), list(hb.c.max), hb.lens, out=buf), 3, sync)
            ref_counts = buf.clone()
            ms_seam = _events_ms(lambda: ops.hist(db, sb, *hb.c.min, *hb.c.max, *hb.lens), 3, sync)
            same = bool(torch.equal(ops.hist(db, sb, *hb.c.min, *hb.c.max, *hb.lens), ref_counts))
            ms_fused = _events_ms(lambda: ops.estimate_init_pose(args, sb, db, auto_swap=True), 3, sync)
            out["c3_hist_icp"]["vote_stage_vs_reference_kernel"] = {
                "pairs": nb, "reference_hist_cuda_kernel_ms": ms_ref, "icpf_hist_votes_f32_ms": ms_seam,
                "counts_bit_identical": same,
                "whole_estimate_init_pose_ms": ms_fused,
                "what": "reference: hist_cuda_kernel (one thread per (pair, i, j), global fp32 atomics, 64 launches of 8 pairs) "
                        "+ the zero-fill of its dense output; icpf_hist_votes_f32: the bit-compatible public seam (same dense "
                        "output); whole_estimate_init_pose: votes + NMS + top-5 + candidate scoring of the fused path, which "
                        "never materialises the dense histogram"}
            del buf, ref_counts
    except Exception as exc:
        out["c3_hist_icp"]["vote_stage_vs_reference_kernel"] = {"error": f"{type(exc).__name__}: {exc}"}
    del s, d, init

    # ---- C4: Waymo-shape frame pair
    sp, sl, dp, dl, _ = synth.make_scene()
    t = [torch.from_numpy(x).to(dev) for x in (sp, dp, sl, dl)]
    fargs = types.SimpleNamespace(thres_dist=THRES, translation_frame=3.34, chunk_size=50, min_cluster_size=30,
                                  thres_box=0.1, max_points=10000, thres_error=0.2, thres_iou=0.2, thres_rot=0.1)
    pose = torch.eye(4, device=dev)
    best, matched = 1e9, 0
    for _ in range(4):
        scan.clear_cache()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        rows, T = E.match_pcds(fargs, *t)
        E.flow_estimation_torch(fargs, t[0], t[1], t[2], t[3], rows, T, pose)
        torch.cuda.synchronize()
        best, matched = min(best, time.perf_counter() - t0), len(rows)
    out["c4_frame"] = {"points": [len(sp), len(dp)], "clusters": 200, "max_points": 10000, "matched_pairs": matched,
                       "ms": best * 1e3, "what": "cluster index + match_pcds (both stages) + flow, wall clock"}

    # ---- row f4: DBSCAN clustering of the C4 source scan (the reference: Open3D / sklearn on the CPU)
    try:
        from icp_flow_b200 import cluster
        from oracle import cluster_oracle as CO
        nonground = torch.from_numpy(sp[sl > -1e7]).to(dev)
        ms_db = _events_ms(lambda: cluster.dbscan_labels(nonground, 0.25, 30), 5, sync)
        t0 = time.perf_counter()
        want = CO.dbscan_labels(nonground.cpu().numpy(), 0.25, 30)
        cpu_s = time.perf_counter() - t0
        got = cluster.dbscan_labels(nonground, 0.25, 30).cpu().numpy()
        out["c4_dbscan"] = {"points": int(nonground.shape[0]), "eps": 0.25, "min_points": 30, "ms": ms_db,
                            "clusters": int(want.max() + 1), "labels_equal_sklearn": bool((got == want).all()),
                            "cpu_baseline": {"ms": cpu_s * 1e3, "kind": "port", "sample": "sklearn.cluster.DBSCAN (kd_tree) on the same points, one run"}}
    except Exception as exc:
        out["c4_dbscan"] = {"error": f"{type(exc).__name__}: {exc}"}
    # ---- row f4, the clusterer the reference's scripts select (--if_hdbscan): HDBSCAN of the same non-ground points
    try:
        pts_h = nonground[:, :3].contiguous()
        cluster.hdbscan_labels(pts_h[:2000], 30)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        got = cluster.hdbscan_labels(pts_h, 30)
        torch.cuda.synchronize()
        ms_h = (time.perf_counter() - t0) * 1e3
        t0 = time.perf_counter()
        want = CO.hdbscan_labels(pts_h.cpu().numpy(), 30)
        cpu_s = time.perf_counter() - t0
        from sklearn.metrics import adjusted_rand_score
        cluster.hdbscan_labels(pts_h[:2000], 30, exact_order=False)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fast = cluster.hdbscan_labels(pts_h, 30, exact_order=False)
        torch.cuda.synchronize()
        ms_fast = (time.perf_counter() - t0) * 1e3
        out["c4_hdbscan"] = {"points": int(pts_h.shape[0]), "min_cluster_size": 30, "ms": ms_h,
                             "any_order": {"ms": ms_fast, "adjusted_rand_vs_sklearn": float(adjusted_rand_score(want, fast)),
                                           "what": "exact_order=False: the spanning tree unique under (weight, min, max) by Boruvka "
                                                   "rounds; equal-weight edges merge in that order, not in sklearn's"},
                             "clusters": int(want.max() + 1), "partition_equals_sklearn": bool(CO.same_partition(got, want)),
                             "what": "core distances + Prim's spanning tree of the mutual-reachability graph on the GPU (one "
                                     "cooperative launch), condensed tree / excess of mass on the host; wall clock",
                             "cpu_baseline": {"ms": cpu_s * 1e3, "kind": "port",
                                              "sample": "sklearn.cluster.HDBSCAN (kd_tree, one core) on the same points, one run"}}
    except Exception as exc:
        out["c4_hdbscan"] = {"error": f"{type(exc).__name__}: {exc}"}

    # ---- C5, one GPU's share (the 8-GPU line carries the sharded run itself)
    ms5, _ = time_c5_shard(dev)
    out["c5_one_gpu_share"] = {"pairs": 4096, "points": POINTS, "icp_iterations": ICP_ITERS, "ms": ms5,
                               "pair_iters_per_s": 4096 * ICP_ITERS / ms5 * 1e3,
                               "roofline_frac": 4096 * ICP_ITERS * algorithmic_bytes_per_pair_iter(POINTS, POINTS) / (ms5 * 1e-3) / 1e9 / peak}
    return out


def bind_to_gpu_cpus(index: int):
    """Pin this rank to the CPUs NVML names as local to its GPU (same NUMA node / PCIe root), so that the pinned staging
    buffers are first-touched there and the H2D streams of the ranks do not all pull through one node."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        pynvml.nvmlDeviceSetCpuAffinity(h)
        cpus = sorted(os.sched_getaffinity(0))
        return f"{len(cpus)} cpus [{cpus[0]}..{cpus[-1]}]"
    except Exception as exc:
        return f"unbound ({type(exc).__name__})"


# ---------------------------------------------------------------------------------------------- engine arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the C3 / C4 side measurements (N=1 only)")
    ap.add_argument("--nn-mode", type=int, default=0)
    ap.add_argument("--gather", default="auto", choices=["auto", "p2p", "nccl"],
                    help="multi-GPU: fused peer-store all-gather from the kernel epilogue (p2p) or NCCL all_gather")
    ap.add_argument("--cell-factor", type=int, default=0, help="tuning: grid cell size in 1/1000 of the gate radius")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "engine" else args.warmup
    if args.impl == "reference":
        return run_reference_arm(args)

    import ctypes
    import torch
    import torch.distributed as dist
    from icp_flow_b200 import _lib, ops, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (the engine has no CPU path)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    affinity = bind_to_gpu_cpus(local_rank)      # before any pinned allocation: first touch lands on the GPU's NUMA node
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n_gpus = world

    P, N = PAIRS_PER_GPU, POINTS
    batch_bytes = 2 * P * N * 16
    pool_n = max(4, -(-int(2.2 * L2_BYTES) // batch_bytes))        # rotating pool > 2x L2 -> every step reads HBM
    src_pool, dst_pool = [], []
    host_packed = None
    for i in range(pool_n):
        s, d, _ = synth.make_pairs(P, N, seed=1234 + rank + 1000 * i, residual_only=True)
        if i == 0:
            # the host side of the e2e path ships the compact format: xyz of the valid rows + CSR offsets (12 B per row)
            host_packed = ops.pack_compact(torch.from_numpy(s), torch.from_numpy(d))      # one pinned buffer per batch
        src_pool.append(torch.from_numpy(s).to(dev))
        dst_pool.append(torch.from_numpy(d).to(dev))
    params = ops.make_params(thres=THRES, max_iterations=ICP_ITERS, relative_rmse_thr=-1.0, early_exit=False,
                             batch_stop=True, nn_mode=args.nn_mode)
    params.reserved[0] = args.cell_factor
    L = _lib.lib()
    ws = torch.empty(L.icpf_workspace_bytes(P, N, 0, 0, 0), device=dev, dtype=torch.uint8)
    out = None
    outs = [None, None]
    gathered = [torch.empty(n_gpus * P, 4, 4, device=dev, dtype=torch.float32) for _ in range(2)]
    launches_per_step = 2   # icp_pairs_kernel + icp_resolve_batch_kernel (batch stop + state at the stop, one launch)
    comm_stream = torch.cuda.Stream() if world > 1 else None
    peer = None
    gather_kind = "none"
    if world > 1:
        gather_kind = "nccl"
        if args.gather in ("auto", "p2p"):
            try:
                from icp_flow_b200.shard import PeerGather
                peer = PeerGather(P, dev, slots=2)
                gather_kind = "p2p"
            except Exception as exc:          # symmetric memory unavailable on this box
                if args.gather == "p2p":
                    raise
                if rank == 0:
                    print(f"[bench] peer gather unavailable ({type(exc).__name__}: {exc}); using NCCL", file=sys.stderr)
        ok = torch.tensor([1 if peer is not None else 0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0:
            peer, gather_kind = None, "nccl"
    computed = [torch.cuda.Event() for _ in range(2)]
    gathered_ev = [torch.cuda.Event() for _ in range(2)]

    def step(i, prof=None):
        # Batches are independent, so the all-gather of step i (comm stream, NVLink) overlaps the kernels of step i+1;
        # outputs are double-buffered and a slot is reused only after its gather has completed.
        nonlocal out
        s = i & 1
        if world > 1 and i >= 2:
            torch.cuda.current_stream().wait_event(gathered_ev[s])
        ev = (prof[0].cuda_event, prof[1].cuda_event) if prof is not None else (0, 0)
        if peer is not None:
            # fused: the kernel epilogue stores the transforms into every rank's gathered buffer (NVLink peer memory);
            # the cross-rank barrier that publishes them runs on the side stream, under the next batch's kernels
            outs[s] = out = ops.icp_batch(src_pool[i % pool_n], dst_pool[i % pool_n], params, out=outs[s], workspace=ws,
                                          ext=peer.ext(s, *ev))
            computed[s].record()
            with torch.cuda.stream(comm_stream):
                comm_stream.wait_event(computed[s])
                peer.finish(s)
                gathered_ev[s].record(comm_stream)
            return out
        outs[s] = out = ops.icp_batch(src_pool[i % pool_n], dst_pool[i % pool_n], params, out=outs[s], workspace=ws,
                                      ext=ops.icp_ext(start_event=ev[0], stop_event=ev[1]) if prof is not None else None)
        if world > 1:
            computed[s].record()
            with torch.cuda.stream(comm_stream):
                comm_stream.wait_event(computed[s])
                dist.all_gather_into_tensor(gathered[s].view(n_gpus * P, 16), out.pose.view(P, 16))
                gathered_ev[s].record(comm_stream)
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    stream = torch.cuda.current_stream()

    # ---- resident-input throughput: K steps, CUDA events on the launching stream, per-step events around the kernel
    for i in range(args.warmup + (args.warmup & 1)):
        step(i)
    prof_events = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                   for _ in range(args.steps)]
    for a, b in prof_events:        # force creation of the underlying CUDA events
        a.record(stream)
        b.record(stream)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.start()
    ev0.record(stream)
    for i in range(args.steps):
        step(i, prof_events[i])
    if world > 1:
        stream.wait_stream(comm_stream)      # the timed region ends when the last gather has landed
    ev1.record(stream)
    barrier()
    sampler.stop()
    ms_total = ev0.elapsed_time(ev1)
    kernel_ms = [a.elapsed_time(b) for a, b in prof_events]
    stats = ops.icp_stats(ws, P).sum(dim=0).tolist()
    iters_done = int(out.iterations.sum().item())
    assert iters_done == P * ICP_ITERS, iters_done     # every pair really executed 20 iterations
    t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    pair_iters_per_step = n_gpus * P * ICP_ITERS
    value = pair_iters_per_step * args.steps / (ms_total * 1e-3)

    # ---- end to end: pinned host inputs -> H2D -> kernels -> D2H of the transforms, every step
    h_pose = [torch.empty(P, 4, 4, dtype=torch.float32).pin_memory() for _ in range(2)]
    pipe = ops.IcpHostPipeline(P, N, params, device=dev)
    e2e_i = 0

    def e2e_step():
        # the public host-buffer call: H2D of this step's inputs, kernels, D2H of its transforms (double-buffered)
        nonlocal e2e_i
        s = e2e_i & 1
        o = pipe.submit_packed(host_packed, h_pose[s], ext=peer.ext(s) if peer is not None else None)
        if peer is not None:
            with torch.cuda.stream(pipe.compute_stream):
                peer.finish(s)
        elif world > 1:
            with torch.cuda.stream(pipe.compute_stream):
                dist.all_gather_into_tensor(gathered[s].view(n_gpus * P, 16), o.pose.view(P, 16))
        e2e_i += 1

    e2e_steps = max(10, min(args.steps, 100))
    for _ in range(4):
        e2e_step()
    pipe.synchronize()
    barrier()
    sampler.start()
    ev0.record(stream)
    pipe.copy_stream.wait_event(ev0)
    pipe.compute_stream.wait_event(ev0)
    for _ in range(e2e_steps):
        e2e_step()
    stream.wait_stream(pipe.copy_stream)
    stream.wait_stream(pipe.compute_stream)
    ev1.record(stream)
    pipe.synchronize()
    barrier()
    sampler.stop()
    t = torch.tensor([ev0.elapsed_time(ev1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = pair_iters_per_step * e2e_steps / (float(t.item()) * 1e-3)
    h2d_bytes = int(host_packed.buffer.numel())

    # ---- the gathered rows really are every rank's transforms: compare the peer-written buffer with one NCCL all_gather
    gather_verified = None
    c5 = None
    if world > 1:
        s_last = (e2e_i - 1) & 1
        ref = torch.empty(n_gpus * P, 16, device=dev, dtype=torch.float32)
        dist.all_gather_into_tensor(ref, pipe.outs[s_last].pose.view(P, 16).contiguous())
        got = peer.slots[s_last][0] if peer is not None else gathered[s_last].view(n_gpus * P, 16)
        torch.cuda.synchronize()
        okv = torch.tensor([1 if torch.equal(got.view(-1, 16), ref) else 0], device=dev)
        dist.all_reduce(okv, op=dist.ReduceOp.MIN)
        gather_verified = bool(int(okv.item()))
        # ---- C5 as named: 4096 pairs per GPU (32768 pairs over 8 GPUs), 20 forced iterations, gather of the transforms
        try:
            P5 = 4096
            peer5 = None
            if peer is not None:
                from icp_flow_b200.shard import PeerGather
                peer5 = PeerGather(P5, dev, slots=2)
            g5 = [torch.empty(n_gpus * P5, 16, device=dev, dtype=torch.float32) for _ in range(2)]

            def after(i, o):
                if peer5 is not None:
                    peer5.finish(i & 1)
                else:
                    dist.all_gather_into_tensor(g5[i & 1], o.pose.view(P5, 16))

            dist.barrier()
            ms5, o5 = time_c5_shard(dev, P5, steps=20, ext_fn=(lambda i: peer5.ext(i & 1)) if peer5 is not None else None,
                                    after_step=after)
            t5 = torch.tensor([ms5], device=dev, dtype=torch.float64)
            dist.all_reduce(t5, op=dist.ReduceOp.MAX)
            ref5 = torch.empty(n_gpus * P5, 16, device=dev, dtype=torch.float32)
            dist.all_gather_into_tensor(ref5, o5.pose.view(P5, 16).contiguous())
            got5 = peer5.slots[(3 + 20 - 1) & 1][0] if peer5 is not None else g5[(3 + 20 - 1) & 1]
            torch.cuda.synchronize()
            ok5 = torch.tensor([1 if torch.equal(got5.view(-1, 16), ref5) else 0], device=dev)
            dist.all_reduce(ok5, op=dist.ReduceOp.MIN)
            c5 = {"pairs_total": n_gpus * P5, "pairs_per_gpu": P5, "points": POINTS, "icp_iterations": ICP_ITERS,
                  "ms_per_step": float(t5.item()), "pair_iters_per_s": n_gpus * P5 * ICP_ITERS / (float(t5.item()) * 1e-3),
                  "gather": gather_kind, "gather_verified": bool(int(ok5.item())),
                  "what": "resident inputs, rotating pool of 5 batches per GPU, max over ranks of the CUDA-event time"}
        except Exception as exc:
            c5 = {"error": f"{type(exc).__name__}: {exc}"}

    if rank == 0:
        peak, peak_kind = load_peaks()
        k_ms = statistics.mean(kernel_ms)
        alg_bytes = P * ICP_ITERS * algorithmic_bytes_per_pair_iter(N, N)
        achieved = alg_bytes / (k_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(n_gpus),
            "run": {"l2": f"rotating pool of {pool_n} input batches ({pool_n * batch_bytes / 2**20:.0f} MiB > 2x 126 MiB L2): "
                          "every step reads its inputs from HBM", "nn_mode": args.nn_mode, "gather": gather_kind,
                    "gather_verified": gather_verified, "host_affinity": affinity},
            "clocks": sampler.summary(),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": P * 64,
                    "steps": e2e_steps, "host_format": "compact: xyz of the valid rows (12 B/row) + CSR offsets in ONE pinned buffer (one H2D copy per "
                    "step), expanded to the padded [P,N,4] batch by icpf_expand_rows_f32 on the device; the padded format would ship "
                    f"{batch_bytes} B per step", "h2d_gbs": h2d_bytes * e2e_steps / (float(t.item()) * 1e-3) / 1e9},
            # (+ the two expand_rows kernels of the compact host format in the e2e steps)
            "gpu_launches": launches_per_step * args.steps + (launches_per_step + 2) * e2e_steps,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": load_ncu_traffic(), "peak_kind": peak_kind, "kernel": "icp_pairs_kernel",
                         "kernel_ms": k_ms, "kernel_share_of_step": k_ms / (ms_total / args.steps),
                         "algorithmic_bytes_per_launch": alg_bytes},
            "nn_search": {"full_search_fraction": stats[0] / float(P * ICP_ITERS * N)},
        }
        if c5 is not None:
            line["secondary"] = {"c5_sharded": c5}
        if n_gpus == 1 and not args.no_secondary:
            try:
                line["secondary"] = time_secondary(dev)
            except Exception as exc:      # side measurements must never cost the headline line
                line["secondary"] = {"error": f"{type(exc).__name__}: {exc}"}
        if n_gpus == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = time_cpu_oracle()[0]
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
