"""CPU tests of the host-side mirror of the reference interface: data files, run utilities, likelihood
writer format, step sharding -- including a world_size-2 gloo run of the data-parallel reduction."""
import os
import struct
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_dataloader_roundtrip(tmp_path):
    import dataLoader as dl
    d = str(tmp_path) + "/"
    dl.write_synthetic_file(d, 1, num_utts=5, rawsize=12, outputDim=7, T_range=(5, 9), L_range=(2, 4))
    loader = dl.DataLoader(d, 12, 8)                      # centre crop 12 -> 8 (dataLoader.py:62-67)
    data_dict, alis, keys, sizes = loader.loadDataFileDict(1)
    assert len(keys) == 5 and sizes.sum() == sum(v.shape[1] for v in data_dict.values())
    raw = np.fromfile(d + "feats1.bin", np.float32).reshape(-1, 12)
    k0 = keys[0]
    assert data_dict[k0].shape == (8, sizes[0]) and data_dict[k0].dtype == np.float32
    assert np.array_equal(data_dict[k0], raw[:sizes[0], 2:10].T)
    assert all(1 <= int(x) < 7 for x in alis[k0])
    loader.loadDataFileAsynch(1)
    dd2, _, keys2, _ = loader.getDataAsynch()
    assert keys2 == keys and np.array_equal(dd2[k0], data_dict[k0])


def test_run_utils(tmp_path):
    import run_utils
    f = str(tmp_path / "cfg.json")
    run_utils.dump_config({"b": 1, "a": [1, 2]}, f)
    assert run_utils.load_config(f) == {"a": [1, 2], "b": 1}
    assert run_utils.TimeString.match(str(run_utils.TimeString()))
    s = run_utils.CfgStruct(x=3)
    assert s.x == 3
    run_utils.touch_file(str(tmp_path / "sentinel"))
    assert os.path.exists(str(tmp_path / "sentinel"))


def test_kaldi_header_format(tmp_path):
    import writeLikelihoods as wl
    f = str(tmp_path / "x.ark")
    with open(f, "wb") as fid:
        wl.writeUttHeader(fid, "utt1", 7, 35)
    b = open(f, "rb").read()
    assert b[:5] == b"utt1 " and b[5:6] == b"\x00" and b[6:10] == b"BFM "
    assert struct.unpack("b", b[10:11])[0] == 4 and struct.unpack("i", b[11:15])[0] == 7
    assert struct.unpack("b", b[15:16])[0] == 4 and struct.unpack("i", b[16:20])[0] == 35


def test_shard_partitions_every_step():
    import parallel
    items = list(range(13))
    for world in (1, 2, 3, 4, 8):
        parts = [parallel.shard(items, r, world) for r in range(world)]
        assert sorted(sum(parts, [])) == items
        assert max(map(len, parts)) - min(map(len, parts)) <= 1
        assert max(map(len, parts)) <= parallel.per_rank_capacity(len(items), world)


def _dp_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path[:0] = [ROOT, os.path.join(ROOT, "stanford-ctc_b200"), os.path.join(ROOT, "tests")]
    import torch
    import torch.distributed as dist
    import parallel
    import recipes
    from oracle import brnn_oracle
    dist.init_process_group("gloo", rank=rank, world_size=world)
    d, r, w = parallel.world_info()
    assert (r, w) == (rank, world)
    datas, labelss = recipes.synth_batch(9, 7, [12, 15, 9, 14, 11], [3, 4, 2, 5, 3], seed=2)
    np.random.seed(11)
    nn = brnn_oracle.NNet(9, 7, 16, 2, 15, temporalLayer=1, dtype=np.float64)
    nn.initParams()
    # this rank's shard of the step, summed gradient flattened with the 4-float statistics tail
    md, ml = parallel.shard(datas, rank, world), parallel.shard(labelss, rank, world)
    costs, grad, skips = nn.costAndGradBatch(md, ml)
    flat = np.concatenate([np.concatenate([g[0].ravel(), g[1].ravel()]) for g in grad] +
                          [np.array([np.sum(~skips), costs[~skips].sum(), np.sum(skips), 0.0])])
    t = torch.from_numpy(flat.copy())
    parallel.allreduce_sum(d, t)
    q.put((rank, t.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_data_parallel_reduction_world2_gloo():
    """Two ranks (gloo, CPU): sharding + one all-reduce of the flat gradient reproduces the
    single-process minibatch gradient and statistics."""
    import torch.multiprocessing as mp
    import recipes
    from oracle import brnn_oracle
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    datas, labelss = recipes.synth_batch(9, 7, [12, 15, 9, 14, 11], [3, 4, 2, 5, 3], seed=2)
    np.random.seed(11)
    nn = brnn_oracle.NNet(9, 7, 16, 2, 15, temporalLayer=1, dtype=np.float64)
    nn.initParams()
    costs, grad, skips = nn.costAndGradBatch(datas, labelss)
    full = np.concatenate([np.concatenate([g[0].ravel(), g[1].ravel()]) for g in grad] +
                          [np.array([5.0, costs.sum(), 0.0, 0.0])])
    assert np.array_equal(res[0], res[1])                 # identical on every rank
    np.testing.assert_allclose(res[0], full, rtol=1e-12, atol=1e-12)


def test_bench_reference_arm_json_contract():
    """`bench.py --impl reference` (the CPU arm the driver times beside ours) prints one JSON line with the
    contract's keys, runs the oracle port on host cores only and needs no GPU."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    env["CTCB_REF_WORKERS"] = "2"
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--gpus", "1",
                        "--steps", "1", "--warmup", "0", "--config", "tiny"], env=env, capture_output=True, text=True,
                       timeout=300)
    assert r.returncode == 0, r.stderr[-1500:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert key in d, key
    assert d["impl"] == "reference" and d["value"] > 0 and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] == 2
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    # a non-zero rank of a multi-GPU launch exits silently
    env2 = dict(env, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r2 = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--gpus", "2",
                         "--steps", "1", "--warmup", "0", "--config", "tiny"], env=env2, capture_output=True, text=True,
                        timeout=300)
    assert r2.returncode == 0 and not [l for l in r2.stdout.splitlines() if l.startswith("{")]


def test_step_utterance_selection_is_consistent_across_ranks():
    """sgd.py:76-88 admission rules, then round-robin sharding: every admitted utterance is owned by exactly one
    rank, skipped ones by none, and all ranks log the same skips."""
    import parallel
    rng = np.random.RandomState(0)
    keys = ["k%d" % i for i in range(23)]
    data = {k: np.zeros((5, int(rng.randint(3, 60))), dtype=np.float32) for k in keys}
    alis = {k: list(range(int(rng.randint(1, 20)))) for k in keys}
    data["k3"] = np.zeros((5, 80), dtype=np.float32)          # longer than the buffers
    alis["k7"] = list(range(data["k7"].shape[1] + 1))         # more labels than frames
    ok = [k for k in keys if data[k].shape[1] <= 64 and data[k].shape[1] >= len(alis[k])]
    assert "k3" not in ok and "k7" not in ok
    for world in (1, 2, 3, 8):
        logs = [[] for _ in range(world)]
        shards = [parallel.select_step_utterances(data, alis, keys, 64, r, world, log=logs[r].append) for r in range(world)]
        assert sorted(sum(shards, [])) == sorted(ok)
        assert max(len(s) for s in shards) - min(len(s) for s in shards) <= 1
        assert shards[0] == ok[0::world]
        assert all(l == logs[0] for l in logs) and len(logs[0]) == len(keys) - len(ok)
        assert any("exceeds batch length" in m for m in logs[0]) and any("less than label length" in m for m in logs[0])


def test_length_bucketed_chunks_cover_every_key_once_and_cut_padding():
    import random
    import parallel
    rng = random.Random(5)
    lens = {("k%d" % i): rng.randint(160, 240) for i in range(1000)}      # the +-20 % ragged-T variant of SURVEY 8d
    keys = list(lens)
    rng.shuffle(keys)
    chunks, padded, real = parallel.bucketed_chunks(keys, lens.get, 32, 16, rng=random.Random(1))
    flat = [k for c in chunks for k in c]
    assert sorted(flat) == sorted(keys) and all(len(c) <= 32 for c in chunks)
    plain = [keys[i:i + 32] for i in range(0, len(keys), 32)]
    plain_padded = sum(max(lens[k] for k in c) * len(c) for c in plain)
    assert real == sum(lens.values())
    assert padded < plain_padded and (padded - real) < 0.25 * (plain_padded - real)     # most of the padding is gone
    # batch size 1 is the reference's schedule: the shuffled order, untouched
    one, p1, r1 = parallel.bucketed_chunks(keys, lens.get, 1)
    assert [c[0] for c in one] == keys and p1 == r1


def test_label_capacity_admission():
    import numpy as np
    import parallel
    dd = {"a": np.zeros((3, 50)), "b": np.zeros((3, 50)), "c": np.zeros((3, 10))}
    alis = {"a": list(range(40)), "b": list(range(12)), "c": list(range(12))}
    logs = []
    used = parallel.select_step_utterances(dd, alis, ["a", "b", "c"], 60, 0, 1, log=logs.append, max_labels=20)
    assert used == ["b"] and len(logs) == 2          # "a": too many labels for the buffers; "c": fewer frames than labels
