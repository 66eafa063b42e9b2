import sys, os, types, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from icp_flow_b200 import ops, synth, _lib
P, N, F = 256, 1024, 6.666
src, dst, meta = synth.make_pairs(P, N, seed=99, residual_only=False)
dev = torch.device("cuda:0")
s, d = torch.from_numpy(src).to(dev), torch.from_numpy(dst).to(dev)
hb = ops._hist_bins(0.1, F, dev)
L = _lib.lib()
ws = torch.zeros(L.icpf_workspace_bytes(P, N, *hb.lens), device=dev, dtype=torch.uint8)
pose = torch.empty(P, 4, 4, device=dev)
rc = L.icpf_hist_init_f32(s.data_ptr(), d.data_ptr(), P, N, ctypes.byref(hb.c), 1, pose.data_ptr(), None, None, None, None, ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
up = lambda v: (v + 255) // 256 * 256
off = up(P*4) + up(P*16) + 256 + up(P*8) + up(P*36) + up(P*12) + up(P*64) + 2*up(P*20)
need = ws[off:off + P*4].view(torch.int32)
print("rc", rc, "need_global:", int(need.sum()), "of", P, need[:16].tolist())
