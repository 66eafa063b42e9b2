#!/usr/bin/env python
"""C2 step time and search fraction for a list of grid cell factors (and any variant library in ICPF_LIB_PATH).

    python tools/sweep_c2.py 2.2 2.6 3.0      # on the GPU box
"""
import os, sys, json, glob, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def one(cfs):
    import torch
    from icp_flow_b200 import ops, synth
    dev = torch.device("cuda:0")
    P, N = 1024, 512
    pool = []
    for k in range(18):
        s, d, _ = synth.make_pairs(P, N, seed=1234 + k, ragged=False, residual_only=True)
        pool.append((torch.from_numpy(s).to(dev), torch.from_numpy(d).to(dev)))
    ws = torch.empty(ops._lib.lib().icpf_workspace_bytes(P, N, 0, 0, 0), device=dev, dtype=torch.uint8)
    for cf in cfs:
        prm = ops.make_params(thres=0.1, max_iterations=20, relative_rmse_thr=-1.0, early_exit=False, batch_stop=True)
        prm.reserved[0] = int(round(cf * 1000))
        out = None
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
        st = ops.icp_stats(ws, P).cpu().numpy()
        print("SW " + json.dumps({"lib": os.path.basename(os.environ.get("ICPF_LIB_PATH", "product")), "cf": cf,
                                  "c2_ms": round(min(times), 4), "search_frac": round(float(st[:, 0].sum()) / (P * N * 20), 4)}), flush=True)


if __name__ == "__main__":
    if sys.argv[1] == "one":
        one([float(x) for x in sys.argv[2:]])
    else:
        vdir = os.path.join(ROOT, "icp_flow_b200", "_variants")
        for lib in [None] + sorted(glob.glob(os.path.join(vdir, "lib_*.so"))):
            env = dict(os.environ)
            if lib:
                env["ICPF_LIB_PATH"] = lib
            else:
                env.pop("ICPF_LIB_PATH", None)
            p = subprocess.run([sys.executable, os.path.abspath(__file__), "one"] + sys.argv[1:], env=env, capture_output=True, text=True)
            out = [l for l in p.stdout.splitlines() if l.startswith("SW ")]
            print("\n".join(out) if out else f"SW FAILED {lib}: {p.stderr[-600:]}")
