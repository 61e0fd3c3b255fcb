"""Time the two sweep launches of the C2 workload in isolation (forward + BPTT), events on the stream."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "stanford-ctc_b200")]
import numpy as np, torch
import _ctcb
from _ctcb import lib, check, ptr
T, B, H = [int(x) for x in (sys.argv[1:4] + ["200", "32", "512"][len(sys.argv) - 1:])]
g = torch.Generator(device="cuda").manual_seed(1)
s = 0.9 / np.sqrt(H / 3.0)
Wf = (torch.rand(H, H, device="cuda", generator=g) * 2 - 1) * s
Wb = (torch.rand(H, H, device="cuda", generator=g) * 2 - 1) * s
pre = torch.randn(T, B, H, device="cuda", generator=g)
lens = torch.full((B,), T, dtype=torch.int32, device="cuda")
oF = torch.empty(T, B, H, device="cuda"); oB = torch.empty_like(oF); dF = torch.empty_like(oF); dB = torch.empty_like(oF)
nscr = int(lib.ctcb_brnn_sweep_workspace_bytes(H, B))
scr = torch.zeros(nscr // 4, dtype=torch.int32, device="cuda")
st = _ctcb.current_stream()
def fwd(): check(lib.ctcb_brnn_sweep_f32(0, T, B, H, ptr(lens), ptr(pre), ptr(Wf), ptr(Wb), ptr(oF), ptr(oB), None, None, 20.0, ptr(scr), nscr, st))
def bwd(): check(lib.ctcb_brnn_sweep_f32(1, T, B, H, ptr(lens), ptr(pre), ptr(Wf), ptr(Wb), ptr(dF), ptr(dB), ptr(oF), ptr(oB), 20.0, ptr(scr), nscr, st))
for _ in range(3): fwd(); bwd()
torch.cuda.synchronize()
for name, fn in (("fwd", fwd), ("bptt", bwd)):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): fn()
    e1.record(); torch.cuda.synchronize()
    print("%s sweep=%s tc=%s nb=%s T=%d B=%d H=%d: %.3f ms/launch (%.2f us/step) flag=%d" % (
        name, os.environ.get("CTCB_SWEEP", "auto"), os.environ.get("CTCB_SWEEP_TC", "auto"), os.environ.get("CTCB_SWEEP_NB", "auto"), T, B, H,
        e0.elapsed_time(e1) / 20, 1e3 * e0.elapsed_time(e1) / 20 / T, int(scr[0])))

if os.environ.get("CTCB_SWEEP_TRACE"):
    al = lambda x: (x + 1023) // 1024 * 1024
    off = (4096 + al(4 * H * H * 4) + al(8 * B * H * 4)) // 8
    tr = scr.view(torch.int64)[off:off + 64 * 16].cpu().numpy().reshape(64, 16)
    names = ["step top", "counter seen", "TMA issued", "first ready (MMA)", "done commit issued", "done seen (epi)",
             "partial in smem", "after cluster sync", "after finalize stores", "after syncthreads", "after arrive",
             "-", "-", "k-block 0 MMAs issued", "k-block 1 data seen", "k-block 4 data seen"]
    import numpy as np
    rows = tr[8:56]
    base = rows[:, 0:1]
    rel = (rows[:, :16] - base).astype(np.float64)
    med = np.median(rel, axis=0)
    period = np.median(np.diff(tr[8:56, 0]))
    print("trace of CTA (0,0,0), last launch (BPTT), SM cycles relative to the step top, median over steps 8..55; step period %.0f cycles" % period)
    for i in np.argsort(med):
        if names[i] == "-":
            continue
        print("  %-24s %8.0f" % (names[i], med[i]))
