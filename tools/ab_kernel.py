#!/usr/bin/env python
"""A/B harness for kernel variants (same ABI, different -D flags).

    python tools/ab_kernel.py build name=-DFLAG[,-DFLAG2] ...   # here (nvcc cross-compiles): icp_flow_b200/_variants/lib_<name>.so
    python tools/ab_kernel.py run                               # on the GPU box: times + output hashes of every variant
    python tools/ab_kernel.py one                               # (internal) one variant, selected by ICPF_LIB_PATH

`run` reports, per variant, the CUDA-event time of one C2 step (1024 pairs x 512 points x 20 forced iterations, rotating
pool of input batches > L2) and md5 hashes of the transforms of three fixed workloads, so that a variant that changes
any result bit is visible at once.
"""
import glob, hashlib, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VDIR = os.path.join(ROOT, "icp_flow_b200", "_variants")


def build(specs):
    from icp_flow_b200 import build as b
    os.makedirs(VDIR, exist_ok=True)
    for spec in specs:
        name, _, flags = spec.partition("=")
        out = os.path.join(VDIR, f"lib_{name}.so")
        b.build_library(force=True, extra_flags=[f for f in flags.split(",") if f], out=out)
        print("built", out)


def one():
    import numpy as np, torch
    from icp_flow_b200 import ops, synth
    dev = torch.device("cuda:0")
    res = {"lib": os.path.basename(os.environ.get("ICPF_LIB_PATH", "product"))}
    # timing: C2 step
    P, N = 1024, 512
    pool = []
    for k in range(18):
        s, d, _ = synth.make_pairs(P, N, seed=1234 + k, ragged=False, residual_only=True)
        pool.append((torch.from_numpy(s).to(dev), torch.from_numpy(d).to(dev)))
    prm = ops.make_params(thres=0.1, max_iterations=20, relative_rmse_thr=-1.0, early_exit=False, batch_stop=True)
    out = ws = None
    ws = torch.empty(ops._lib.lib().icpf_workspace_bytes(P, N, 0, 0, 0), device=dev, dtype=torch.uint8)
    for k in range(6):
        out = ops.icp_batch(*pool[k % 18], prm, out=out, workspace=ws)
    torch.cuda.synchronize()
    times = []
    for rep in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(36):
            out = ops.icp_batch(*pool[k % 18], prm, out=out, workspace=ws)
        e1.record(); torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1) / 36)
    res["c2_ms"] = min(times)
    res["c2_Mpi_s"] = P * 20 / min(times) / 1e3
    st = ops.icp_stats(ws, P).cpu().numpy()
    res["search_frac"] = float(st[:, 0].sum()) / (P * N * 20)
    h = lambda *ts: hashlib.md5(b"".join(t.detach().cpu().numpy().tobytes() for t in ts)).hexdigest()[:12]
    r = ops.icp_batch(*pool[0], prm)
    res["h_c2"] = h(r.R, r.T, r.rmse)
    # reference stopping rule, ragged clusters, full motion (zero-inlier pairs, early exits, re-run pass)
    s, d, _ = synth.make_pairs(512, 512, seed=77, ragged=True, residual_only=False)
    s, d = torch.from_numpy(s).to(dev), torch.from_numpy(d).to(dev)
    r = ops.icp_batch(s, d, ops.make_params())
    res["h_stop"] = h(r.R, r.T, r.rmse, r.iterations, r.batch)
    s, d, _ = synth.make_pairs(64, 5000, seed=5, ragged=True, residual_only=True)
    s, d = torch.from_numpy(s).to(dev), torch.from_numpy(d).to(dev)
    r = ops.icp_batch(s, d, ops.make_params())
    res["h_big"] = h(r.R, r.T, r.rmse, r.iterations, r.batch)
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record(); r = ops.icp_batch(s, d, ops.make_params()); t1.record(); torch.cuda.synchronize()
    res["big_ms"] = t0.elapsed_time(t1)
    # histogram initialisation on a C3-shaped batch (1024 pairs x 1024 points, 135 x 135 x 3 bins)
    import types
    s, d, _ = synth.make_pairs(1024, 1024, seed=99, ragged=False, residual_only=False)
    s, d = torch.from_numpy(s).to(dev), torch.from_numpy(d).to(dev)
    args = types.SimpleNamespace(thres_dist=0.1, translation_frame=6.666, chunk_size=50)
    init = ops.estimate_init_pose(args, s, d, auto_swap=True)
    torch.cuda.synchronize()
    t0.record(); init = ops.estimate_init_pose(args, s, d, auto_swap=True); t1.record(); torch.cuda.synchronize()
    res["init_ms"] = t0.elapsed_time(t1)
    res["h_init"] = h(init)
    t0.record(); T = ops.hist_icp(args, s, d); t1.record(); torch.cuda.synchronize()
    res["hist_icp_ms"] = t0.elapsed_time(t1)
    res["h_hist_icp"] = h(T)
    print("AB " + json.dumps(res))


def run():
    libs = [None] + sorted(glob.glob(os.path.join(VDIR, "lib_*.so")))
    for lib in libs:
        env = dict(os.environ)
        if lib:
            env["ICPF_LIB_PATH"] = lib
        else:
            env.pop("ICPF_LIB_PATH", None)
        p = subprocess.run([sys.executable, os.path.abspath(__file__), "one"], env=env, capture_output=True, text=True)
        lines = [l for l in p.stdout.splitlines() if l.startswith("AB ")]
        print(lines[-1] if lines else f"AB FAILED {lib}: {p.stderr[-800:]}")


if __name__ == "__main__":
    {"build": lambda: build(sys.argv[2:]), "run": run, "one": one}[sys.argv[1]]()
