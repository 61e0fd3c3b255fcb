"""Profiling target: the CTC kernel alone on a batch larger than L2 (for ncu; not a benchmark).
usage: python tools/prof_ctc.py [B] [T] [K] [L] [reps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "stanford-ctc_b200")]
import numpy as np
import torch
import ctc_fast
from _ctcb import lib

B, T, K, L, reps = [int(x) for x in (sys.argv[1:6] + ["8192", "200", "62", "30", "3"][len(sys.argv) - 1:])]
g = torch.Generator(device="cuda").manual_seed(3)
acts = torch.randn(B, T, K, device="cuda", generator=g)
rng = np.random.RandomState(4)
seqs = torch.from_numpy((1 + rng.randint(0, K - 1, size=(B * L))).astype(np.int32)).cuda()
offs = torch.arange(0, (B + 1) * L, L, dtype=torch.int32, device="cuda")
lens = torch.full((B,), T, dtype=torch.int32, device="cuda")
grad = torch.empty_like(acts)
ws = torch.empty(lib.ctcb_ctc_workspace_bytes(B, T, L), dtype=torch.uint8, device="cuda")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(reps):
    e0.record()
    nll, _, skip = ctc_fast.ctc_loss_batch(acts, lens, seqs, offs, L, grad=grad, workspace=ws)
    e1.record()
    torch.cuda.synchronize()
    alg = B * (8.0 * K * T + 4.0 * L + 4.0)
    print("rep %d: %.3f ms  %.1f GB/s algorithmic  %.2f M utt/s  skips=%d" % (
        i, e0.elapsed_time(e1), alg / e0.elapsed_time(e1) / 1e6, B / e0.elapsed_time(e1) / 1e3, int(skip.sum())))
