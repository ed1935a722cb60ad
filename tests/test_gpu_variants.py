"""GPU: the alternative kernel variants kept in-tree behind environment switches stay correct.
The switches are read once per process, so each variant runs the kernel parity tests in a fresh interpreter."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra, select):
    env = dict(os.environ, **env_extra)
    r = subprocess.run([sys.executable, "-m", "pytest", "tests/test_gpu_kernels.py", "-q", "-x", "-k", select,
                        "-p", "no:cacheprovider"], cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]


def test_single_cta_gemm_tiles():
    """MOLLY_GEMM_PAIR=0: 128x256 / 128x128 single-CTA tcgen05 tiles instead of cta_group::2 pairs."""
    _run({"MOLLY_GEMM_PAIR": "0"}, "gemm")


def test_persistent_attention_p_in_tmem():
    """MOLLY_ATTN_PERSISTENT=1: streamed work items, P kept in TMEM (A operand from tensor memory), K/V rings."""
    _run({"MOLLY_ATTN_PERSISTENT": "1"}, "attention")


def test_polynomial_exp2_attention():
    """MOLLY_ATTN_POLY=1|2: a quarter / half of the softmax exponentials evaluated on the FMA pipe (cubic, 1e-4 rel)."""
    _run({"MOLLY_ATTN_POLY": "2"}, "attention")
