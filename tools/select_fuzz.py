#!/usr/bin/env python
"""icpf_match_select_f32 against the reference's matrices (utils_match.py:70-75, 94-135) restated with torch on random
inputs -- duplicated errors (ties), NaN errors, rejected pairs, labels missing from the lists:   (GPU)
    python tools/select_fuzz.py [n_cases] [first_seed]"""
import os, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from icp_flow_b200 import ops


def matrices(pairs, su, du, ev, accept, T, thres):
    """the reference's scatter loop + match_segments_descend, vectorised (pairs are unique, so the scatter order is moot)"""
    ns, nd = len(su), len(du)
    m_err = torch.full((ns, nd, 2), 1e8)
    m = [torch.zeros((ns, nd, 2)) for _ in range(3)]
    m_T = torch.zeros((ns, nd, 4, 4))
    for k in range(len(pairs)):
        if not accept[k]:
            continue
        si, di = torch.nonzero(su == pairs[k, 0]), torch.nonzero(du == pairs[k, 1])
        if len(si) == 0 or len(di) == 0:
            continue
        si, di = int(si), int(di)
        m_err[si, di] = ev[0][k]
        for j in range(3):
            m[j][si, di] = ev[1 + j][k]
        m_T[si, di] = T[k]
    e_min = m_err.min(-1)[0]
    rows_i = torch.arange(ns)
    cols_i = torch.argmin(e_min, dim=1)
    ok = e_min[rows_i, cols_i] < thres
    rows_i, cols_i = rows_i[ok], cols_i[ok]
    rows = torch.cat([su[rows_i][:, None].float(), du[cols_i][:, None].float(), m_err[rows_i, cols_i]] + [x[rows_i, cols_i] for x in m], dim=1)
    return rows, m_T[rows_i, cols_i]


n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 200
seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 0
bad = 0
for i in range(n_cases):
    g = torch.Generator().manual_seed(seed0 + i)
    ns, nd = int(torch.randint(1, 40, (1,), generator=g)), int(torch.randint(1, 40, (1,), generator=g))
    su = torch.sort(torch.randperm(200, generator=g)[:ns])[0]
    du = torch.sort(torch.randperm(200, generator=g)[:nd])[0]
    allp = torch.stack([su.repeat_interleave(nd), du.repeat(ns)], 1)
    P = int(torch.randint(1, len(allp) + 1, (1,), generator=g))
    pairs = allp[torch.randperm(len(allp), generator=g)[:P]]
    if i % 4 == 0:
        pairs = torch.cat([pairs, torch.tensor([[777, int(du[0])], [int(su[0]), 888]])])       # labels outside the lists
        P += 2
    err = torch.rand(P, 2, generator=g) * 0.4
    err = torch.round(err * 20) / 20 if i % 2 else err                                           # many exact ties
    if i % 5 == 0:
        err[torch.randint(0, P, (max(1, P // 10),), generator=g), 0] = float("nan")
    other = [torch.rand(P, 2, generator=g) for _ in range(3)]
    accept = (torch.rand(P, generator=g) < 0.7).int()
    T = torch.rand(P, 4, 4, generator=g)
    args = types.SimpleNamespace(thres_error=0.2)
    want_rows, want_T = matrices(pairs, su, du, [err] + other, accept, T, 0.2)
    rows, Tm, s_left, d_left = ops.match_select(args, pairs.cuda(), su.cuda(), du.cuda(), [x.cuda() for x in [err] + other],
                                                accept.cuda(), T.cuda(), return_left=True)
    rows, Tm = rows.cpu(), Tm.cpu()
    same = rows.shape == want_rows.shape and torch.equal(rows.nan_to_num(-7.0), want_rows.nan_to_num(-7.0)) and torch.equal(Tm, want_T)
    same = same and torch.equal(s_left.cpu(), su[~torch.isin(su, want_rows[:, 0].long())]) and torch.equal(d_left.cpu(), du[~torch.isin(du, want_rows[:, 1].long())])
    if not same:
        bad += 1
        print("selection differs: case", seed0 + i, ns, nd, P)
print(f"select fuzz (seeds {seed0}..{seed0 + n_cases - 1}): {n_cases} cases; differing from the reference's matrices: {bad}")
