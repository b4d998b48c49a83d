"""The reference's unmodified training entry point against the drop-in `losses` package.

Runs only where the reference checkout exists (the build container; it cannot travel to the GPU box, and no reference source
may be copied into this repository), on the CPU: tests/ref_entry_driver.py imports Point_Cloud_Resistration/train_W_COS.py
as is and calls its train() -- train_one_epoch, test_one_epoch, the torch.save snapshot code -- and load_checkpoint(),
once with the reference's own losses and once with dropin/losses on sys.path.  The two runs must report the same losses,
end with the same weights, and load each other's snapshots."""
import json
import os
import subprocess
import sys

import pytest

REF = "/root/reference/Point_Cloud_Resistration"
HERE = os.path.dirname(os.path.abspath(__file__))

pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="needs the reference checkout (build container only)")


def _run(mode, workdir, other=""):
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE="1", OMP_NUM_THREADS="4")
    r = subprocess.run([sys.executable, os.path.join(HERE, "ref_entry_driver.py"), mode, str(workdir), other], capture_output=True,
                       text=True, env=env, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("RESULT ")][-1]
    return json.loads(line[len("RESULT "):])


def test_reference_train_entry_runs_unmodified_on_the_dropin_losses(tmp_path):
    (tmp_path / "ref").mkdir()
    (tmp_path / "ours").mkdir()
    ref = _run("reference", tmp_path / "ref")
    ours = _run("dropin", tmp_path / "ours", ref["snapshot"])
    ref2 = _run("reference", tmp_path / "ref", ours["snapshot"])
    # same reported losses, epoch by epoch (tensorboard scalars of train(): train_W_COS.py:240-241)
    assert len(ref["scalars"]) == len(ours["scalars"]) == 4
    for (tag_r, v_r, e_r), (tag_o, v_o, e_o) in zip(ref["scalars"], ours["scalars"]):
        assert (tag_r, e_r) == (tag_o, e_o)
        assert v_o == pytest.approx(v_r, rel=2e-4), (tag_r, e_r, v_r, v_o)
    # same trained weights (PCRNet 4.2 M parameters through two epochs of Adam; phi after its ascent steps)
    assert ours["model_digest"] == pytest.approx(ref["model_digest"], rel=1e-5)
    assert ours["phi_digest"] == pytest.approx(ref["phi_digest"], rel=1e-5)
    assert ours["phi_keys"] == ref["phi_keys"] and ours["n_phi_params"] == ref["n_phi_params"] == 1284
    # resume: own snapshot, and the other implementation's snapshot, through the reference's load_checkpoint
    for run, other in ((ours, ref), (ref2, ours)):
        assert run["resume_own"]["epoch"] >= 1 and run["resume_own"]["phi_op_states"] > 0
        assert run["resume_other"]["phi_digest"] == pytest.approx(other["resume_own"]["phi_digest"], rel=1e-6)
        assert run["resume_other"]["phi_out"] == pytest.approx(other["resume_own"]["phi_out"], rel=1e-5)
