"""The reference's training-loop calling sequence through the drop-in modules (tools/dropin_loop.py), in a fresh process so
that `model` / `src.sk_utils` resolve to dropin/ exactly as they would for main.py with dropin/ on PYTHONPATH."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_reference_training_loop_through_dropin(cuda_device):
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "dropin_loop.py")], cwd=ROOT, stdout=subprocess.PIPE,
                         stderr=subprocess.STDOUT, text=True, timeout=400)
    print(res.stdout[-3000:])
    assert res.returncode == 0 and "DROPIN_LOOP OK" in res.stdout


def test_dropin_modules_resolve_to_the_b200_mirrors():
    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r); import model, src.sk_utils as s, datasets.audio_utils as a; "
            "import selavi_b200.model as m; assert model.load_model is m.load_model and model.AVModel is m.AVModel; "
            "import selavi_b200.sk_utils as k; assert s.cluster is k.cluster and s.optimize_L_sk_gpu is k.optimize_L_sk_gpu; "
            "print('ok')") % (ROOT, os.path.join(ROOT, "dropin"))
    res = subprocess.run([sys.executable, "-c", code], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert res.returncode == 0 and res.stdout.strip().endswith("ok"), res.stdout[-2000:]
