"""Pointwise accuracy of the recurrent sweep against a float64 NumPy recurrence over a LONG utterance (error growth
with t), with the reference's own weight scale (brnnet.py:66-70).  usage: sweep_accuracy.py T B H"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "stanford-ctc_b200")]
import numpy as np, torch
import _ctcb
from _ctcb import lib, check, ptr
T, B, H = [int(x) for x in (sys.argv[1:4] + ["800", "4", "1024"][len(sys.argv) - 1:])]
rng = np.random.RandomState(3)
s = np.sqrt(6) / np.sqrt(2 * H)
Wf = rng.uniform(-s, s, (H, H)).astype(np.float32); Wb = rng.uniform(-s, s, (H, H)).astype(np.float32)
pre = (rng.randn(T, B, H) * 1.0).astype(np.float32)
lens = np.full(B, T, dtype=np.int32)
F = np.zeros((T, B, H)); W64 = Wf.astype(np.float64); p64 = pre.astype(np.float64)
for t in range(T):
    F[t] = np.clip(p64[t] + (F[t - 1] @ W64.T if t > 0 else 0.0), 0.0, 20.0)
dev = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()
oF = torch.empty(T, B, H, device="cuda"); oB = torch.empty_like(oF)
nscr = int(lib.ctcb_brnn_sweep_workspace_bytes(H, B))
scr = torch.zeros(nscr // 4, dtype=torch.int32, device="cuda")
d_lens, d_pre, d_Wf, d_Wb = dev(lens), dev(pre), dev(Wf), dev(Wb)       # keep the device copies alive
check(lib.ctcb_brnn_sweep_f32(0, T, B, H, ptr(d_lens), ptr(d_pre), ptr(d_Wf), ptr(d_Wb), ptr(oF), ptr(oB), None, None,
                              20.0, ptr(scr), nscr, _ctcb.current_stream()))
torch.cuda.synchronize()
g = oF.cpu().numpy().astype(np.float64)
print("tc=%s sweep=%s T=%d B=%d H=%d flag=%d  mean|F|=%.3f  frac(F>0)=%.3f" % (os.environ.get("CTCB_SWEEP_TC", "auto"), os.environ.get("CTCB_SWEEP", "auto"), T, B, H, int(scr[0]), np.abs(F).mean(), (F > 0).mean()))
for t in sorted(set([0, 1, 2, 10, 50, 100, T // 2, T - 1])):
    if t < T:
        e = np.abs(g[t] - F[t])
        print("  t=%4d  max abs err %.3e   rms err %.3e   rms F %.3e   sign flips %d" % (t, e.max(), np.sqrt((e ** 2).mean()), np.sqrt((F[t] ** 2).mean()), int(((g[t] > 0) != (F[t] > 0)).sum())))
