"""The attention kernel has a second implementation behind K5_ATTN_IMPL=4 (csrc/attention4.cu: software-pipelined
softmax over 64-row KV tiles, double-buffered S and P in TMEM).  The knob is read once per process, so the parity run
(dense with ragged KV tails, cross-attention shape, large logits -> lazy rescale, block-sparse with a partial last query
item; tolerances as in test_gpu_ops.py, stated in tests/gpu_attn_variants.py) happens in a subprocess."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("impl", ["2", "4"])
def test_attention_implementations_match_the_fp32_restatement(impl):
    env = dict(os.environ, K5_ATTN_IMPL=impl, K5_BENCH_S="4608")
    r = subprocess.run([sys.executable, os.path.join(HERE, "gpu_attn_variants.py"), f"impl{impl}=K5_ATTN_IMPL:{impl}"],
                       env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert f"impl{impl}: parity" in r.stdout and "-> OK" in r.stdout, r.stdout + r.stderr
    assert "exit code" not in r.stdout, r.stdout
