"""GPU parity tests of the options of the hot path: the blank-forced CTC
(ctc_fast_blankforce.pyx), the uni-directional net (nnets/rnnet.py) and the feed-forward net (nnets/nnet.py).

Blank-forced CTC: golden vectors produced by the UNMODIFIED reference module, tolerance 1e-4 relative on loss
and gradient (float64 state on the device; float32 in/out).  Nets: golden vectors of the float64 restatement
(oracle/brnn_oracle.py), tolerances of tests/test_brnn_gpu.py."""
import numpy as np
import pytest

import recipes
from oracle import brnn_oracle, ctc_oracle

pytestmark = pytest.mark.gpu
TOL = 1e-4
COST_TOL, GRAD_TOL = 1e-4, 1e-3


def _rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)


# ------------------------------------------------------------------------------------ blank-forced CTC
@pytest.mark.parametrize("name", recipes.ALL_BF)
def test_blankforce_golden_through_dropin_module(name, golden_bf, cuda):
    import ctc_fast_blankforce as bf
    probs, seq = recipes.bf_case(name)
    nll, grad, skip = bf.ctc_loss(np.asfortranarray(probs.astype(np.float64)), seq)
    assert isinstance(nll, float) and grad.dtype == np.float64 and grad.flags.f_contiguous
    assert skip == bool(golden_bf[name + "/skip"]) and not skip
    g_nll = float(golden_bf[name + "/nll"])
    assert abs(nll - g_nll) / abs(g_nll) <= TOL
    st = recipes.golden_stride(*probs.shape)
    assert _rel(grad[:, ::st], golden_bf[name + "/grad"].astype(np.float64)) <= TOL
    assert abs(np.linalg.norm(grad) - float(golden_bf[name + "/gradnorm"])) <= TOL * float(golden_bf[name + "/gradnorm"])


def test_blankforce_ragged_batch_from_logits(cuda):
    """Time-major [T][B][K] logits, ragged lengths and state counts, fused float32 softmax; rows beyond an
    utterance's length carry zero gradient; a zero-probability frame gives skip with zero gradient."""
    import ctc_fast_blankforce as bf
    torch = cuda
    rng = np.random.RandomState(12)
    B, Tmax, K = 6, 90, 29
    lens = np.array([90, 61, 33, 90, 12, 75], dtype=np.int32)
    nlab = [20, 9, 4, 44, 2, 0]
    logits = rng.randn(Tmax, B, K).astype(np.float32) * 2.0
    seqs = []
    for n in nlab:
        s = np.zeros(2 * n + 1, dtype=np.int32)
        s[1::2] = 1 + rng.randint(K - 1, size=n)
        seqs.append(s)
    off = np.concatenate([[0], np.cumsum([len(s) for s in seqs])]).astype(np.int32)
    acts = torch.from_numpy(logits).cuda()
    nll, grad, skip = bf.ctc_loss_batch(acts, torch.from_numpy(lens).cuda(), torch.from_numpy(np.concatenate(seqs)).cuda(),
                                        torch.from_numpy(off).cuda(), max(len(s) for s in seqs), is_prob=False,
                                        utt_stride=K, frame_stride=B * K)
    nll, grad, skip = nll.cpu().numpy(), grad.cpu().numpy(), skip.cpu().numpy()
    for u in range(B):
        T = lens[u]
        z = logits[:T, u].astype(np.float32)
        e = np.exp(z - z.max(axis=1, keepdims=True))
        p = (e / e.sum(axis=1, keepdims=True)).astype(np.float32)          # float32 hand-over, brnnet.py:170
        o_nll, o_grad, o_skip = ctc_oracle.ctc_loss_blankforce(np.asfortranarray(p.T.astype(np.float64)), seqs[u])
        assert bool(skip[u]) == o_skip and not o_skip
        assert abs(nll[u] - o_nll) / abs(o_nll) <= TOL, u
        assert _rel(grad[:T, u].T.astype(np.float64), o_grad) <= TOL, u
        assert not grad[T:, u].any()
    # skip path: probabilities with an all-zero frame
    probs, seq = recipes.bf_case("bf_small")
    p64 = np.asfortranarray(probs.astype(np.float64))
    p64[:, 4] = 0.0
    n2, g2, s2 = bf.ctc_loss(p64, seq)
    assert s2 and not g2.any()


def test_blankforce_best_path_and_errors(cuda):
    import ctc_fast_blankforce as bf
    probs = np.zeros((10, 9))
    path = [0, 3, 3, 0, 3, 1, 5, 5, 8]
    probs[path, np.arange(9)] = 1.0
    assert bf.decode_best_path(np.asfortranarray(probs)) == [3, 3, 1, 5, 8]      # no label filter in this module
    assert bf.decode_best_path(np.asfortranarray(probs)) == ctc_oracle.decode_best_path_blankforce(probs)
    p, s = recipes.bf_case("bf_small")
    with pytest.raises(ValueError):
        bf.ctc_loss(np.ascontiguousarray(p.astype(np.float64)), s)                # C order: the reference raises too
    with pytest.raises(ValueError):
        bf.ctc_loss(np.asfortranarray(p.astype(np.float64)), np.zeros(1100, dtype=np.int32))   # > 1024 states


# ------------------------------------------------------------------------------------ rnnet / nnet
def _load(nn, golden, tag):
    import torch
    for i, (w, b) in enumerate(nn.stack):
        w.copy_(torch.from_numpy(golden["%s/w%d" % (tag, i)]))
        b.copy_(torch.from_numpy(golden["%s/b%d" % (tag, i)].reshape(tuple(b.shape))))


@pytest.mark.parametrize("tag", ["uni", "dnn"])
def test_rnnet_and_nnet_golden_minibatch(tag, golden_rnn, cuda):
    import nnets.nnet
    import nnets.rnnet
    datas, labelss = recipes.rnn_variant_batch()
    np.random.seed(9)
    if tag == "uni":
        nn = nnets.rnnet.NNet(13, 11, 64, 3, 41, temporalLayer=2, maxUtts=4, maxLabels=15)
    else:
        nn = nnets.nnet.NNet(13, 11, 64, 3, 41, maxUtts=4, maxLabels=15)
    nn.initParams()
    assert len(nn.stack) == (5 if tag == "uni" else 4)
    # same np.random draw order as rnnet.py:38-61 / nnet.py:20-23: the untouched recurrent matrix scales by 1.2
    for i, (w, b) in enumerate(nn.stack):
        assert tuple(w.shape) == golden_rnn["%s/w%d" % (tag, i)].shape
    if tag == "dnn":
        assert np.array_equal(nn.stack[0][0].cpu().numpy(), golden_rnn["dnn/w0"])
    _load(nn, golden_rnn, tag)
    costs, grad, skips = nn.costAndGradBatch(datas, labelss)
    assert not skips.any()
    np.testing.assert_allclose(costs, golden_rnn[tag + "/costs"], rtol=COST_TOL)
    for i, (dw, db) in enumerate(grad):
        assert _rel(dw.cpu().numpy().astype(np.float64), golden_rnn["%s/dw%d" % (tag, i)]) <= GRAD_TOL, i
        if i < 4:
            assert _rel(db.cpu().numpy().astype(np.float64), golden_rnn["%s/db%d" % (tag, i)]) <= GRAD_TOL, i
    # the reference signature, one utterance at a time, sums to the minibatch gradient
    tot = [np.zeros(tuple(dw.shape)) for dw, _ in grad]
    g_batch = [dw.cpu().numpy().astype(np.float64) for dw, _ in grad]
    for d, l in zip(datas, labelss):
        c, g, s = nn.costAndGrad(d, l)
        assert not s
        for t, (dw, _) in zip(tot, g):
            t += dw.cpu().numpy()
    for t, gb in zip(tot, g_batch):
        assert _rel(gb, t) < 1e-4


@pytest.mark.parametrize("H,B", [(128, 5), (256, 7), (512, 6), (96, 3)])
def test_rnnet_register_resident_sizes_vs_oracle(H, B, cuda):
    """One-direction launches of the cluster kernels (H = 128/256/512) and of the generic kernel (H = 96)."""
    import nnets.rnnet
    D, K, N, tl = 41, 62, 3, 2
    rng = np.random.RandomState(H + B)
    lens = [int(x) for x in rng.randint(30, 60, size=B)]
    nlabs = [int(x) for x in rng.randint(3, 15, size=B)]
    datas, labelss = recipes.synth_batch(D, K, lens, nlabs, seed=H)
    np.random.seed(23)
    on = brnn_oracle.NNet(D, K, H, N, 60, temporalLayer=tl, dtype=np.float64, unidirectional=True)
    on.initParams()
    np.random.seed(23)
    nn = nnets.rnnet.NNet(D, K, H, N, 60, temporalLayer=tl, maxUtts=B, maxLabels=16)
    nn.initParams()
    for (w, _), (ow, _) in zip(nn.stack, on.stack):
        assert np.array_equal(w.cpu().numpy(), ow.astype(np.float32))
    o_costs, o_grad, o_skips = on.costAndGradBatch(datas, labelss)
    costs, grad, skips = nn.costAndGradBatch(datas, labelss)
    assert np.array_equal(skips, o_skips)
    np.testing.assert_allclose(costs, o_costs, rtol=COST_TOL)
    for i, ((dw, db), (odw, odb)) in enumerate(zip(grad, o_grad)):
        assert _rel(dw.cpu().numpy().astype(np.float64), odw) <= GRAD_TOL, i


def test_rnnet_forward_only_returns_best_path(cuda):
    """rnnet.py:138-139: with train=False costAndGrad returns ctc.decode_best_path(probs)."""
    import ctc_fast
    import nnets.rnnet
    datas, _ = recipes.synth_batch(13, 11, [25], [4], seed=3)
    np.random.seed(2)
    nn = nnets.rnnet.NNet(13, 11, 64, 3, 41, train=False, temporalLayer=2)
    nn.initParams()
    np.random.seed(2)
    on = brnn_oracle.NNet(13, 11, 64, 3, 41, train=False, temporalLayer=2, unidirectional=True)
    on.initParams()
    probs = on.forward(datas[0])[3]
    want = ctc_oracle.decode_best_path(np.asfortranarray(probs.astype(np.float64)))
    got = nn.costAndGrad(datas[0])
    assert isinstance(got, tuple) and (list(got[0]), list(got[1])) == (want[0], want[1])


def test_unidirectional_sweep_c_abi(cuda):
    """ctcb_brnn_sweep_f32 with Wb == NULL: the forward-in-time recurrence and its BPTT alone."""
    import _ctcb
    from _ctcb import lib, check, ptr
    torch = cuda
    rng = np.random.RandomState(8)
    T, B, H = 30, 6, 512
    s = 0.9 / np.sqrt(H / 3.0)
    W = rng.uniform(-s, s, (H, H))
    pre = rng.randn(T, B, H) * 2.0 + 0.5
    pre[T // 2] += 25.0
    lens = rng.randint(T // 2, T + 1, size=B); lens[0] = T
    for b in range(B):
        pre[lens[b]:, b] = 0.0
    F = np.zeros_like(pre)
    for b in range(B):
        for t in range(lens[b]):
            F[t, b] = np.clip(pre[t, b] + (W @ F[t - 1, b] if t > 0 else 0.0), 0.0, 20.0)
    dev = lambda x: torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).cuda()
    d_len = torch.from_numpy(lens.astype(np.int32)).cuda()
    oF = torch.empty(T, B, H, device="cuda")
    scratch = torch.zeros(1024, dtype=torch.int32, device="cuda")
    d_W = dev(W)
    check(lib.ctcb_brnn_sweep_f32(0, T, B, H, ptr(d_len), ptr(dev(pre)), ptr(d_W), None, ptr(oF), None, None, None,
                                  20.0, ptr(scratch), 4096, _ctcb.current_stream()))
    torch.cuda.synchronize()
    assert scratch[:3].tolist() == [0, 0, 0]
    gF = oF.cpu().numpy().astype(np.float64)
    assert (F >= 20.0).any() and np.abs(gF - F).max() < 2e-4 * np.abs(F).max()
    d = rng.randn(T, B, H)
    for b in range(B):
        d[lens[b]:, b] = 0.0
    m = (gF > 0) & (gF < 20.0)
    dF = np.zeros_like(d)
    for b in range(B):
        for t in range(lens[b] - 1, -1, -1):
            dF[t, b] = m[t, b] * (d[t, b] + (W.T @ dF[t + 1, b] if t + 1 < lens[b] else 0.0))
    odF = torch.empty(T, B, H, device="cuda")
    check(lib.ctcb_brnn_sweep_f32(1, T, B, H, ptr(d_len), ptr(dev(d)), ptr(d_W), None, ptr(odF), None, ptr(oF), None,
                                  20.0, ptr(scratch), 4096, _ctcb.current_stream()))
    torch.cuda.synchronize()
    assert scratch[:3].tolist() == [0, 0, 0]
    assert np.abs(odF.cpu().numpy() - dF).max() < 2e-4 * max(1.0, np.abs(dF).max())
