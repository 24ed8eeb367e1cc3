"""The reference's OWN evaluation entry points, unmodified, driven through this package's module.

north_star: "test_PMG.py --mode benchmark and Functions/PMG_evaluation.py load the … checkpoint and run unchanged".
``python -m pmg_b200.compat /root/reference/test_PMG.py …`` is what a user runs; here it is executed for real in a
subprocess with: test-only stubs for the GUI / plotting packages the reference imports at module level and this image
lacks (tests/stubs: matplotlib, tkinter, pytorch_grad_cam, screeninfo, pykalman), a 4-image folder, a Tasks.json, a
hyper-parameter JSON, a mapping JSON and a seeded synthetic ``best_model_PMG.pth`` (the real one is absent from the mount).

The scripts must get through argument parsing → ``build_model_from_hparams`` (Models.models_PMG resolves to this
package) → ``load_best_model(strict=True)`` → the image folder walk / DataLoader → the first ``model(x)``:
  * in this container (no GPU) that call must fail with THIS package's "no CPU fallback" error, raised from the reference's
    own call site (utils/benchmark_utils.py:70, Functions/PMG_evaluation.py:129) — proof that the reference reached the
    boundary with the B200 module in place and that no eager fallback exists;
  * on a box with a GPU and a reference checkout the same command runs to completion (exit code 0).
/root/reference does not exist on the GPU box, so these tests are skipped there."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from helpers import REFERENCE, ROOT, load_contract, torch_state_dict
from oracle import synth

pytestmark = pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="reference checkout only exists in the build container")
STUBS = os.path.join(ROOT, "tests", "stubs")


@pytest.fixture(scope="module")
def fixture_dir(tmp_path_factory):
    from PIL import Image
    d = tmp_path_factory.mktemp("pmg_callers")
    c, hp, tasks = load_contract()
    (d / "Tasks.json").write_text(json.dumps(c["tasks"]))
    (d / "hyperparameters_PMG.json").write_text(json.dumps({"hparams": hp}))
    torch.save(torch_state_dict(synth.pmg_state_dict(0)), d / "best_model_PMG.pth")
    rng = np.random.default_rng(0)
    data = {"set": {}}
    bench = d / "bench"
    classes = {"clear": "Clear", "rain": "Rain"}
    for i in range(4):
        cls = "clear" if i < 2 else "rain"
        (bench / cls).mkdir(parents=True, exist_ok=True)
        p = bench / cls / f"img{i}.png"
        Image.fromarray(rng.integers(0, 256, (300, 340, 3), dtype=np.uint8)).save(p)     # Resize(256)+CenterCrop(224) do real work
        rec = {"image_path": str(p)}
        for t, names in c["tasks"].items():
            rec[t] = classes[cls] if t == "Weather Type" else names[i % len(names)]
        data["set"][f"img{i}"] = rec
    (d / "data.json").write_text(json.dumps(data))
    (d / "mapping.json").write_text(json.dumps({"Weather Type": {"clear": ["Clear", "Sun and Clear"], "rain": ["Rain", "Fog and Rain"]}}))
    return d


def _run(fixture_dir, *extra):
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([STUBS, ROOT, env.get("PYTHONPATH", "")])
    cmd = [sys.executable, "-m", "pmg_b200.compat", os.path.join(REFERENCE, "test_PMG.py"),
           "--build_classifier", str(fixture_dir / "Tasks.json"), "--model_path", str(fixture_dir / "best_model_PMG.pth"),
           "--config_path", str(fixture_dir / "hyperparameters_PMG.json"), "--save_dir", str(fixture_dir / "out"),
           "--count_params", *extra]
    return subprocess.run(cmd, cwd=str(fixture_dir), env=env, capture_output=True, text=True, timeout=600)


def _check(r, call_site):
    assert "Model loaded (strict=True)" in r.stdout, r.stdout + r.stderr           # PMG_model_utils.py:187-188
    assert "2,431,284" in r.stdout                                                 # print_model_parameters on our module
    if torch.cuda.is_available():
        assert r.returncode == 0, r.stderr[-3000:]
        return
    assert r.returncode != 0
    assert "no CPU fallback" in r.stderr, r.stderr[-3000:]
    tb = r.stderr
    assert call_site in tb and "models_PMG.py" in tb, tb[-3000:]
    assert tb.index(call_site) < tb.rindex("models_PMG.py")                        # raised from inside the reference's call


def test_reference_benchmark_mode_reaches_the_b200_forward(fixture_dir):
    """test_PMG.py --mode benchmark (test_PMG.py:206-222 → utils/benchmark_utils.py:62-75: B = 1 per image)."""
    r = _run(fixture_dir, "--mode", "benchmark", "--benchmark_folder", str(fixture_dir / "bench"),
             "--benchmark_mapping", str(fixture_dir / "mapping.json"), "--roc_output", str(fixture_dir / "roc"))
    _check(r, os.path.join("utils", "benchmark_utils.py"))


def test_reference_classifier_mode_reaches_the_b200_forward(fixture_dir):
    """test_PMG.py --mode classifier (test_PMG.py:135-145 → Functions/PMG_evaluation.py:98,124-150: DataLoader batches,
    ragged last batch)."""
    r = _run(fixture_dir, "--mode", "classifier", "--data", str(fixture_dir / "data.json"), "--batch_size", "3")
    _check(r, os.path.join("Functions", "PMG_evaluation.py"))
