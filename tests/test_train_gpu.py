"""GPU end-to-end test of the driver surface: runNNet.run on synthetic reference-format files writes the
run directory the reference writes (cfg.json, params.pk, epoch, num_files, last_cost, sentinel,
train.log), the cost goes down, --cfg_file resumes, and --test writes Kaldi arks."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_run_resume_and_test_mode(cuda, tmp_path):
    import dataLoader as dl
    import runNNet
    data = str(tmp_path / "data") + "/"
    for fn in (1, 2):
        dl.write_synthetic_file(data, fn, num_utts=12, rawsize=20, outputDim=9, T_range=(20, 39), L_range=(3, 8))
    args = ["--layerSize", "64", "--numLayers", "3", "--temporalLayer", "2", "--epochs", "2", "--step", "1e-3",
            "--momentum", "0.9", "--dataDir", data, "--numFiles", "2", "--inputDim", "16", "--rawDim", "20",
            "--outputDim", "9", "--maxUttLen", "40", "--maxLabels", "8", "--save_every", "1", "--batchSize", "4",
            "--runDir", str(tmp_path / "runs"), "--quiet"]
    SGD, nn = runNNet.run(args)
    run_dir = os.path.join(str(tmp_path / "runs"), os.listdir(str(tmp_path / "runs"))[0])
    for f in ("cfg.json", "params.pk", "epoch", "num_files", "last_cost", "sentinel", "train.log",
              "params.pk.epoch00", "params.pk.epoch01"):
        assert os.path.exists(os.path.join(run_dir, f)), f
    assert open(os.path.join(run_dir, "epoch")).read() == "1"
    assert len(SGD.costt) == 12 and np.isfinite(SGD.costt).all()
    assert np.mean(SGD.costt[-3:]) < np.mean(SGD.costt[:3])          # it learns
    # resume: nothing left to do (epochs reached) but state must load
    SGD2, nn2 = runNNet.run(["--cfg_file", os.path.join(run_dir, "cfg.json")])
    assert SGD2.it == SGD.it and bool((nn2.params == nn.params).all())
    # forward-only likelihood dump (runNNet.py:208-237)
    out = str(tmp_path / "ll")
    runNNet.run(["--cfg_file", os.path.join(run_dir, "cfg.json"), "--test", "--dataDir", data, "--numFiles", "1",
                 "--outDir", out])
    ark = open(os.path.join(out, "loglikelihoods1.ark"), "rb").read()
    assert ark.startswith(b"utt1_0000 \x00BFM ") and os.path.exists(os.path.join(out, "loglikelihoods_1.pk"))
