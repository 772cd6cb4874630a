"""The GEMM and conv3d kernels have tuning modes selected by environment variables that are read once per process
(K5_GEMM_CLUSTER = 0 single CTA / 1 CTA pair + TMA multicast / 3 CTA pair under one 2-SM MMA = default;
K5_CONV_CLUSTER = 0 / 1).  The default modes are what every other GPU test exercises; here the operator-level parity
tests are re-run in a subprocess for each non-default mode so that the fallbacks keep working."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _rerun(env_overrides, target, keyword):
    env = dict(os.environ, **env_overrides)
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(HERE, target), "-x", "-q", "-k", keyword, "-p",
                        "no:cacheprovider"], env=env, capture_output=True, text=True, timeout=900, cwd=os.path.dirname(HERE))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert " passed" in r.stdout and "failed" not in r.stdout, r.stdout[-3000:]


@pytest.mark.parametrize("mode", ["0", "1"])
def test_gemm_parity_in_the_non_default_cluster_modes(mode):
    _rerun({"K5_GEMM_CLUSTER": mode}, "test_gpu_ops.py", "gemm")


def test_conv3d_parity_with_the_single_cta_kernel():
    _rerun({"K5_CONV_CLUSTER": "0"}, "test_gpu_vae.py", "conv3d or decoder")
