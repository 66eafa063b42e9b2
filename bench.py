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
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture (profiles/), or None."""
    path = os.path.join(ROOT, "profiles", "ncu_r1_metrics.json")
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
    """--impl reference: the reference's CPU implementation of the path (restated oracle) on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import torch
    from icp_flow_b200 import synth
    from oracle import icp_oracle as O

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sample_pairs = 128
    src, dst, _ = synth.make_pairs(sample_pairs, POINTS, seed=1234, residual_only=True)
    a, c = torch.from_numpy(src), torch.from_numpy(dst)
    run = lambda: O.icp_loop(a, c, thres=THRES, max_iterations=ICP_ITERS, relative_rmse_thr=-1.0)
    for _ in range(args.warmup):
        run()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        run()
    dt = time.perf_counter() - t0
    value = sample_pairs * ICP_ITERS * args.steps / dt
    sample = (f"each step = {sample_pairs} of the {PAIRS_PER_GPU} pairs x {POINTS} pts x {ICP_ITERS} iterations; "
              "oracle/icp_oracle.py icp_loop (torch CPU fp32 + OpenMP C knn leaf)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
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
def time_secondary(dev):
    """Not the headline metric: the two other single-GPU BASELINE configs, timed with CUDA events / wall clock so that
    they are measured in the same run.  C3 at a quarter of its pair count (1024 of 4096 pairs x 1024 points, full
    hist_icp with the 135 x 135 x 3 histogram) and C4 (Waymo-shape frame pair through match_pcds + flow)."""
    import types
    import torch
    import icp_flow_b200 as E
    from icp_flow_b200 import scan, synth
    out = {}
    s, d, _ = synth.make_pairs(1024, 1024, seed=99, ragged=False, residual_only=False)
    s, d = torch.from_numpy(s).to(dev), torch.from_numpy(d).to(dev)
    args = types.SimpleNamespace(thres_dist=THRES, translation_frame=6.666, chunk_size=50)
    for _ in range(2):
        E.hist_icp(args, s, d)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(5):
        E.hist_icp(args, s, d)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    out["c3_hist_icp"] = {"pairs": 1024, "points": 1024, "bins": [135, 135, 3], "ms": ms, "pairs_per_s": 1024 / ms * 1e3}
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
    return out


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
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n_gpus = world

    P, N = PAIRS_PER_GPU, POINTS
    batch_bytes = 2 * P * N * 16
    pool_n = max(4, -(-int(2.2 * L2_BYTES) // batch_bytes))        # rotating pool > 2x L2 -> every step reads HBM
    src_pool, dst_pool = [], []
    host_src = host_dst = None
    for i in range(pool_n):
        s, d, _ = synth.make_pairs(P, N, seed=1234 + rank + 1000 * i, residual_only=True)
        if i == 0:
            host_src = torch.from_numpy(s).pin_memory()
            host_dst = torch.from_numpy(d).pin_memory()
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
    launches_per_step = 3   # icp_pairs_kernel + icp_resolve_batch_kernel + icp_select_batch_kernel (state at the batch stop)
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
        if prof is not None:
            L.icpf_profile_next_icp(ctypes.c_void_p(prof[0].cuda_event), ctypes.c_void_p(prof[1].cuda_event))
        if peer is not None:
            # fused: the kernel epilogue stores the transforms into every rank's gathered buffer (NVLink peer memory);
            # the cross-rank barrier that publishes them runs on the side stream, under the next batch's kernels
            peer.arm(s)
            outs[s] = out = ops.icp_batch(src_pool[i % pool_n], dst_pool[i % pool_n], params, out=outs[s], workspace=ws)
            computed[s].record()
            with torch.cuda.stream(comm_stream):
                comm_stream.wait_event(computed[s])
                peer.finish(s)
                gathered_ev[s].record(comm_stream)
            return out
        outs[s] = out = ops.icp_batch(src_pool[i % pool_n], dst_pool[i % pool_n], params, out=outs[s], workspace=ws)
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
        if peer is not None:
            peer.arm(s)
        o = pipe.submit(host_src, host_dst, h_pose[s])
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

    if rank == 0:
        peak, peak_kind = load_peaks()
        k_ms = statistics.mean(kernel_ms)
        alg_bytes = P * ICP_ITERS * algorithmic_bytes_per_pair_iter(N, N)
        achieved = alg_bytes / (k_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(workload_config(n_gpus), l2=f"rotating pool of {pool_n} input batches "
                           f"({pool_n * batch_bytes / 2**20:.0f} MiB > 2x 126 MiB L2): every step reads its inputs from HBM",
                           nn_mode=args.nn_mode, gather=gather_kind),
            "clocks": sampler.summary(),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": batch_bytes, "d2h_bytes_per_step": P * 64,
                    "steps": e2e_steps},
            "gpu_launches": launches_per_step * args.steps,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": load_ncu_traffic(), "peak_kind": peak_kind, "kernel": "icp_pairs_kernel",
                         "kernel_ms": k_ms, "kernel_share_of_step": k_ms / (ms_total / args.steps),
                         "algorithmic_bytes_per_launch": alg_bytes},
            "nn_search": {"full_search_fraction": stats[0] / float(P * ICP_ITERS * N),
                          "cache_refreshes_per_pair": stats[1] / float(P)},
        }
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
