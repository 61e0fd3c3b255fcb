"""GPU parity tests of the BRNN step through the reference call surface (nnets.brnnet.NNet, sgd.SGD) and
the C ABI, against the committed golden vectors and the float64 oracle restatement of brnnet.py.

Tolerances: the dense arithmetic is exact fp32 (FFMA), the reference's own precision (cudamat sgemm);
against the float64 oracle we require  |cost - ref|/|ref| <= 1e-4  and, per weight tensor,
||dW - dW_ref||_F / ||dW_ref||_F <= 1e-4 (north_star's bound; accumulated fp32 rounding over T recurrent steps is
typically ~1e-5)."""
import io

import numpy as np
import pytest

import recipes
from oracle import brnn_oracle

pytestmark = pytest.mark.gpu
COST_TOL, GRAD_TOL = 1e-4, 1e-4


def _rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)


def _load_stack(nn, golden, prefix):
    import torch
    for i, (w, b) in enumerate(nn.stack):
        w.copy_(torch.from_numpy(golden["%s/w%d" % (prefix, i)]))
        b.copy_(torch.from_numpy(golden["%s/b%d" % (prefix, i)].reshape(tuple(b.shape))))


def test_gemm_all_layouts_vs_numpy(cuda):
    import _ctcb
    from _ctcb import lib, check, ptr
    torch = cuda
    rng = np.random.RandomState(0)
    ws = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")
    for (M, N, K) in [(300, 62, 41), (257, 130, 19), (128, 128, 4096), (64, 512, 20000), (1, 7, 3), (513, 512, 512),
                      (6400, 512, 62), (640, 512, 41), (200, 40, 32), (129, 33, 33)]:
        for ta in (0, 1):
            for tb in (0, 1):
                A = rng.randn(*((K, M) if ta else (M, K))).astype(np.float32)
                B = rng.randn(*((N, K) if tb else (K, N))).astype(np.float32)
                bias = rng.randn(N).astype(np.float32)
                msk = rng.randn(M, N).astype(np.float32)
                ref = (A.T if ta else A).astype(np.float64) @ (B.T if tb else B).astype(np.float64) + bias
                ref = np.maximum(ref, 0) * (msk > 0)
                dA, dB, dbias, dmsk = (torch.from_numpy(x).cuda() for x in (A, B, bias, msk))
                C = torch.empty(M, N, device="cuda")
                check(lib.ctcb_gemm_f32(ta, tb, M, N, K, 1.0, ptr(dA), A.shape[1], ptr(dB), B.shape[1], 0.0, ptr(C), N,
                                        ptr(dbias), 1, ptr(dmsk), ptr(ws), ws.numel(), _ctcb.current_stream()))
                assert _rel(C.cpu().numpy().astype(np.float64), ref) < 1e-5, (M, N, K, ta, tb)


def test_reference_cpu_brnn_recipe(golden_brnn, cuda):
    """rnnetcpu.py:181-193 recipe (D=20,T=10,K=6,H=30,N=3,temporalLayer=2, labels [0,1,2])."""
    import nnets.brnnet as rnnet
    cfg, data, labels = recipes.rnnetcpu()
    np.random.seed(33); np.random.randn(20, 10)
    nn = rnnet.NNet(cfg["inputDim"], cfg["outputDim"], cfg["layerSize"], cfg["numLayers"], cfg["maxBatch"],
                    temporalLayer=cfg["temporalLayer"])
    nn.initParams()
    for i, (w, b) in enumerate(nn.stack):       # same np.random draw order as brnnet.py:38-41,66-70
        assert np.array_equal(w.cpu().numpy(), golden_brnn["rnnetcpu/w%d" % i])
    assert nn.paramCount() == sum(golden_brnn["rnnetcpu/w%d" % i].size + golden_brnn["rnnetcpu/b%d" % i].size
                                  for i in range(6))
    cost, grad, skip = nn.costAndGrad(data.astype(np.float32), labels)
    assert not skip
    assert abs(cost - float(golden_brnn["rnnetcpu/cost"])) / float(golden_brnn["rnnetcpu/cost"]) <= COST_TOL
    for i, (dw, db) in enumerate(grad):
        assert _rel(dw.cpu().numpy().astype(np.float64), golden_brnn["rnnetcpu/dw%d" % i]) <= GRAD_TOL, i
        if i < 4:
            assert _rel(db.cpu().numpy().astype(np.float64), golden_brnn["rnnetcpu/db%d" % i]) <= GRAD_TOL, i


def test_ragged_minibatch_with_clip_and_l2(golden_brnn, cuda):
    """5 utterances of different lengths, recurrence driven into the 20.0 clip, reg > 0."""
    import nnets.brnnet as rnnet
    lens, nlabs = [37, 50, 21, 50, 8], [9, 14, 5, 20, 3]
    datas, labelss = recipes.synth_batch(13, 11, lens, nlabs, seed=7)
    assert int(golden_brnn["ragged/clip_hits"]) > 0
    nn = rnnet.NNet(13, 11, 64, 3, 50, temporalLayer=2, reg=1e-3, maxUtts=5, maxLabels=20)
    np.random.seed(0)
    nn.initParams()
    _load_stack(nn, golden_brnn, "ragged")
    costs, grad, skips = nn.costAndGradBatch(datas, labelss)
    assert not skips.any()
    np.testing.assert_allclose(costs, golden_brnn["ragged/costs"], rtol=COST_TOL)
    assert abs(nn.regcost - float(golden_brnn["ragged/regcost"])) / float(golden_brnn["ragged/regcost"]) < 1e-5
    for i, (dw, db) in enumerate(grad):
        assert _rel(dw.cpu().numpy().astype(np.float64), golden_brnn["ragged/dw%d" % i]) <= GRAD_TOL, i
        if i < 4:
            assert _rel(db.cpu().numpy().astype(np.float64), golden_brnn["ragged/db%d" % i]) <= GRAD_TOL, i
    # batch == sum of per-utterance calls (B = 1 path of the reference signature)
    tot = [np.zeros(tuple(dw.shape)) for dw, _ in grad]
    nn1 = rnnet.NNet(13, 11, 64, 3, 50, temporalLayer=2, reg=0.0, maxUtts=1, maxLabels=20)
    nn1.initParams()
    _load_stack(nn1, golden_brnn, "ragged")
    for d, l in zip(datas, labelss):
        c, g, s = nn1.costAndGrad(d, l)
        for t, (dw, _) in zip(tot, g):
            t += dw.cpu().numpy()
    nn0 = rnnet.NNet(13, 11, 64, 3, 50, temporalLayer=2, reg=0.0, maxUtts=5, maxLabels=20)
    nn0.initParams()
    _load_stack(nn0, golden_brnn, "ragged")
    _, g0, _ = nn0.costAndGradBatch(datas, labelss)
    for t, (dw, _) in zip(tot, g0):
        assert _rel(dw.cpu().numpy().astype(np.float64), t) < 1e-4


@pytest.mark.parametrize("H,N,tl,B", [(128, 2, 1, 8), (256, 3, 2, 9), (512, 2, 1, 4), (96, 1, 1, 3), (64, 2, -1, 4)])
def test_register_resident_sweep_sizes_vs_oracle(H, N, tl, B, cuda):
    """Layer sizes that take the register-resident recurrence (H = 128/256/512), the generic one
    (H = 96, 64), the literal 1-layer extension (temporalLayer == numLayers) and the no-temporal DNN."""
    import nnets.brnnet as rnnet
    D, K = 41, 62
    rng = np.random.RandomState(H + B)
    lens = [int(x) for x in rng.randint(30, 60, size=B)]
    nlabs = [int(x) for x in rng.randint(3, 15, size=B)]
    datas, labelss = recipes.synth_batch(D, K, lens, nlabs, seed=H)
    top = (tl == N)
    np.random.seed(21)
    on = brnn_oracle.NNet(D, K, H, N, 60, temporalLayer=tl, dtype=np.float64, allow_top_temporal=top)
    on.initParams()
    np.random.seed(21)
    nn = rnnet.NNet(D, K, H, N, 60, temporalLayer=tl, maxUtts=B, maxLabels=16, allowTopTemporal=top)
    nn.initParams()
    assert (nn.temporalLayer > 0) == (tl > 0)
    o_costs, o_grad, o_skips = on.costAndGradBatch(datas, labelss)
    costs, grad, skips = nn.costAndGradBatch(datas, labelss)
    assert np.array_equal(skips, o_skips)
    np.testing.assert_allclose(costs, o_costs, rtol=COST_TOL)
    for i, ((dw, db), (odw, odb)) in enumerate(zip(grad, o_grad)):
        assert _rel(dw.cpu().numpy().astype(np.float64), odw) <= GRAD_TOL, i
    # forward-only mode returns K x T float32 probabilities (brnnet.py:171-173)
    nt = rnnet.NNet(D, K, H, N, 60, train=False, temporalLayer=tl, allowTopTemporal=top)
    np.random.seed(21)
    nt.initParams()
    probs = nt.costAndGrad(datas[0])
    assert probs.dtype == np.float32 and probs.shape == (K, lens[0])
    o_probs = on.forward(datas[0])[3]
    assert np.abs(probs - o_probs).max() < 1e-5


def test_nesterov_step_matches_reference_update(cuda):
    """sgd.py:91-161 for three consecutive steps (momentum warm-up 0.5), clip active."""
    import nnets.brnnet as rnnet
    import sgd
    D, K, H, N, B = 13, 11, 32, 2, 4
    datas, labelss = recipes.synth_batch(D, K, [25, 30, 18, 30], [6, 8, 4, 9], seed=9)
    np.random.seed(2)
    on = brnn_oracle.NNet(D, K, H, N, 30, temporalLayer=1, dtype=np.float64)
    on.initParams()
    vel = [[np.zeros_like(w), np.zeros_like(b)] for w, b in on.stack]
    np.random.seed(2)
    nn = rnnet.NNet(D, K, H, N, 30, temporalLayer=1, maxUtts=B, maxLabels=10)
    nn.initParams()
    opt = sgd.SGD(nn, 30, alpha=1e-3, momentum=0.9, maxGradNorm=5.0, batchSize=B, verbose=False)
    for it in range(1, 4):
        mom = 0.5 if it <= 10 else 0.9
        on.updateParams(mom, vel)
        _, g, _ = on.costAndGradBatch(datas, labelss)
        on.updateParams(-mom, vel)
        gn, _ = brnn_oracle.sgd_step(on, vel, g, it, 1e-3, 0.9, maxGNorm=5.0)
        assert gn > 5.0                                      # the clip is exercised
        opt.it = it
        nn._batch.pack(datas, labelss).upload()
        opt.step_device(nn._batch, mom)
        for (w, b), (ow, ob) in zip(nn.stack, on.stack):
            assert _rel(w.cpu().numpy().astype(np.float64), ow) < 2e-5
        assert abs(float(opt._gnorm2.sqrt().item()) - gn) / gn < 1e-3


def test_sgd_run_reference_schedule_and_checkpoint(cuda, tmp_path):
    """SGD.run with batchSize=1 follows the reference's per-utterance schedule; toFile/fromFile use the
    reference's two-pickle params.pk layout (sgd.py:36-42, brnnet.py:258-267)."""
    import pickle
    import random
    import nnets.brnnet as rnnet
    import sgd
    datas, labelss = recipes.synth_batch(13, 11, [25, 30, 18, 30, 5], [6, 8, 4, 9, 7], seed=9)
    keys = ["k%d" % i for i in range(5)]
    data_dict = dict(zip(keys, datas)); alis = dict(zip(keys, [list(map(str, l)) for l in labelss]))
    np.random.seed(2); random.seed(33)
    nn = rnnet.NNet(13, 11, 32, 2, 30, temporalLayer=1)
    nn.initParams()
    opt = sgd.SGD(nn, 30, alpha=1e-4, momentum=0.95, verbose=False)
    opt.run(data_dict, alis, list(keys), None)
    assert opt.it == 5 and len(opt.costt) == 4              # the T < |l| utterance was skipped (sgd.py:84-88)
    assert all(np.isfinite(opt.costt)) and len(opt.expcost) == 4
    f = str(tmp_path / "params.pk")
    with open(f, "wb") as fid:
        opt.toFile(fid); nn.toFile(fid)
    with open(f, "rb") as fid:
        it, costt, expcost, vstack = pickle.load(fid)
        stack = pickle.load(fid)
    assert it == 5 and len(vstack) == len(stack) == 5 and stack[0][0].shape == (32, 13) and stack[3][1].shape == (1, 1)
    nn2 = rnnet.NNet(13, 11, 32, 2, 30, temporalLayer=1); nn2.initParams()
    opt2 = sgd.SGD(nn2, 30, alpha=1e-4, momentum=0.95, verbose=False)
    with open(f, "rb") as fid:
        opt2.fromFile(fid); nn2.fromFile(fid)
    assert opt2.it == 5 and bool((nn2.params == nn.params).all()) and bool((opt2._vflat == opt._vflat).all())


def test_check_grad_finite_differences(cuda):
    """NNet.check_grad (brnnet.py:279-297) on a tiny net: forward differences in fp32 are noisy, so only
    a loose agreement is asserted (the tight check is the float64 oracle comparison above)."""
    import nnets.brnnet as rnnet
    datas, labelss = recipes.synth_batch(6, 5, [12], [3], seed=4)
    np.random.seed(8)
    nn = rnnet.NNet(6, 5, 8, 2, 12, temporalLayer=1)
    nn.initParams()
    pairs = nn.check_grad(datas[0], labelss[0], epsilon=1e-2, maxChecks=2, verbose=False)
    ana = np.array([p[0] for p in pairs]); num = np.array([p[1] for p in pairs])
    assert np.abs(ana - num).max() < 0.05 * max(1.0, np.abs(ana).max())


def test_graph_replayed_steps_equal_eager_steps(cuda):
    """sgd.SGD.step_device captures the step into a CUDA graph on its third occurrence and replays it: six steps with
    graphs must leave bit-identical parameters to six eager steps, and the library's launch counter keeps counting."""
    import nnets.brnnet as rnnet
    import sgd
    from _ctcb import lib
    torch = cuda
    D, K, H, N, B = 13, 11, 128, 2, 4
    datas, labelss = recipes.synth_batch(D, K, [25, 30, 18, 30], [6, 8, 4, 9], seed=9)
    outs = []
    for use in (False, True):
        np.random.seed(2)
        nn = rnnet.NNet(D, K, H, N, 30, temporalLayer=1, maxUtts=B, maxLabels=10)
        nn.initParams()
        opt = sgd.SGD(nn, 30, alpha=1e-3, momentum=0.9, maxGradNorm=5.0, batchSize=B, verbose=False)
        opt.useGraphs = use
        st = torch.cuda.Stream()
        st.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(st):
            nn._batch.pack(datas, labelss).upload()
            n0 = lib.ctcb_launch_count()
            for it in range(6):
                opt.it = 20 + it
                opt.step_device(nn._batch, 0.9)
            per_step = (lib.ctcb_launch_count() - n0) / 6.0
        st.synchronize()
        assert (len(opt._graphs) == 1) == use
        outs.append((nn.params.clone(), per_step))
    assert bool((outs[0][0] == outs[1][0]).all())
    assert outs[0][1] == outs[1][1] and outs[0][1] > 10
