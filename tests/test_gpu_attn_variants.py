"""The attention kernel is built in several forms that one process cannot switch between (the knobs are read once):
the fixed-offset softmax under a proven score bound (the DiT's path, every 4th exponential pair on the FMA pipe), the
same with all exponentials on the MUFU (K5_ATTN_BOUNDED is honoured per call, the polynomial share is compile time), and
the general running-max kernel with lazy rescaling.  Each parity run (dense with ragged KV tails, cross-attention shape,
large logits, block-sparse with a partial last query item; tolerances as in test_gpu_ops.py, stated in
tests/gpu_attn_variants.py) happens in a subprocess."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("name,env", [("general", {}), ("bounded", {"K5_VARIANT_BOUND": "1"}),
                                      ("bounded_declined", {"K5_VARIANT_BOUND": "1", "K5_ATTN_BOUNDED": "0"}),
                                      ("general_poly1", {"K5_ATTN_POLY": "1"})])
def test_attention_forms_match_the_fp32_restatement(name, env):
    spec = name + "=" + ",".join(f"{k}:{v}" for k, v in env.items())
    r = subprocess.run([sys.executable, os.path.join(HERE, "gpu_attn_variants.py"), spec],
                       env=dict(os.environ, K5_BENCH_S="4608"), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert f"{name}: parity" in r.stdout and "-> OK" in r.stdout, r.stdout + r.stderr
    assert "exit code" not in r.stdout, r.stdout
