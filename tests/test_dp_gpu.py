"""GPU data-parallel test (needs >= 2 GPUs; skipped otherwise; the same arithmetic without NCCL is covered on one GPU by
tests/test_configs_gpu.py::test_two_shards_summed_equal_the_full_batch_with_l2): two ranks over NCCL (reg > 0, so the L2
term must be added once, after the sum), each with its shard of the step's utterances, must end with bit-identical parameters that match the single-GPU step on the full
minibatch (the all-reduce sums the flat gradient + statistics tail; sgd.py:91-161 semantics)."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
ROOT = %r
sys.path[:0] = [ROOT, os.path.join(ROOT, "stanford-ctc_b200"), os.path.join(ROOT, "tests")]
import numpy as np, torch, torch.distributed as dist
rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
if world > 1:
    dist.init_process_group("nccl")
import random, recipes, nnets.brnnet as rnnet, sgd, parallel
datas, labelss = recipes.synth_batch(13, 11, [25, 30, 18, 30, 22, 27], [6, 8, 4, 9, 5, 7], seed=9)
keys = ["k%%d" %% i for i in range(6)]
dd = dict(zip(keys, datas)); alis = dict(zip(keys, [list(map(str, l)) for l in labelss]))
np.random.seed(2); random.seed(33)
nn = rnnet.NNet(13, 11, 64, 2, 30, temporalLayer=1, reg=1e-3, maxUtts=parallel.per_rank_capacity(6, world), maxLabels=10)
nn.initParams()
opt = sgd.SGD(nn, 30, alpha=1e-3, momentum=0.9, maxGradNorm=5.0, batchSize=6, verbose=False)
for _ in range(3):
    opt.run(dd, alis, list(keys), None)
torch.cuda.synchronize()
np.save(os.path.join(sys.argv[1], "params_w%%d_r%%d.npy" %% (world, rank)), nn.params.cpu().numpy())
np.save(os.path.join(sys.argv[1], "cost_w%%d_r%%d.npy" %% (world, rank)), np.array(opt.costt))
if world > 1:
    dist.barrier(); dist.destroy_process_group()
'''


def test_two_rank_step_matches_single_gpu(cuda, tmp_path):
    if cuda.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    # the single-GPU reference run and the 2-rank run start together (most of their wall time is start-up)
    p1_ = subprocess.Popen([sys.executable, str(script), str(tmp_path)], env=env, stdout=subprocess.PIPE,
                           stderr=subprocess.PIPE, text=True)
    p2_ = subprocess.Popen([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                            "--master-addr", "127.0.0.1", "--master-port", "29621", str(script), str(tmp_path)],
                           env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    try:
        _, e1 = p1_.communicate(timeout=600)
        _, e2 = p2_.communicate(timeout=600)
    finally:
        for p_ in (p1_, p2_):
            if p_.poll() is None:
                p_.kill()
    assert p1_.returncode == 0, e1[-2000:]
    assert p2_.returncode == 0, e2[-2000:]
    p1 = np.load(tmp_path / "params_w1_r0.npy")
    p20, p21 = np.load(tmp_path / "params_w2_r0.npy"), np.load(tmp_path / "params_w2_r1.npy")
    assert np.array_equal(p20, p21)                                   # replicas stay bit-identical
    assert np.linalg.norm(p20 - p1) / np.linalg.norm(p1) < 1e-5       # and follow the single-GPU trajectory
    c1, c2 = np.load(tmp_path / "cost_w1_r0.npy"), np.load(tmp_path / "cost_w2_r0.npy")
    np.testing.assert_allclose(c1, c2, rtol=1e-4)
