"""GPU: the alternative kernel variants kept in-tree behind environment switches stay correct.
The switches are read once per process, so each variant runs the kernel parity tests in a fresh interpreter."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra, select, files=("tests/test_gpu_kernels.py",)):
    env = dict(os.environ, **env_extra)
    r = subprocess.run([sys.executable, "-m", "pytest", *files, "-q", "-x", "-k", select,
                        "-p", "no:cacheprovider"], cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]


def test_single_cta_gemm_tiles():
    """MOLLY_GEMM_PAIR=0: 128x256 / 128x128 single-CTA tcgen05 tiles instead of cta_group::2 pairs."""
    _run({"MOLLY_GEMM_PAIR": "0"}, "gemm")


def test_one_cta_per_item_attention():
    """MOLLY_ATTN_STREAM=0: grid = number of work items (no item streaming inside a CTA)."""
    _run({"MOLLY_ATTN_STREAM": "0"}, "attention")


def test_one_cta_per_sm_two_tile_attention():
    """MOLLY_ATTN_V2=1: attention2.cu -- one CTA per SM, two 128-row tiles, P in tensor memory, an epilogue warp-group and
    sequenced exp2 phases (FlashAttention-4 style; head_dim <= 64).  Forward, log-sum-exp and the backward that consumes it."""
    _run({"MOLLY_ATTN_V2": "1"}, "attention")


def test_register_pipelined_attention():
    """MOLLY_ATTN_PIPE=1: 64-key blocks, S(g+1) loaded from TMEM under the exp2 phase of block g, S issued a block ahead."""
    _run({"MOLLY_ATTN_PIPE": "1"}, "attention")


def test_polynomial_exp2_attention():
    """MOLLY_ATTN_POLY=1|2: a quarter / half of the softmax exponentials evaluated on the FMA pipe (cubic, 1e-4 rel)."""
    _run({"MOLLY_ATTN_POLY": "2"}, "attention")


def test_64_key_blocks_attention():
    """MOLLY_ATTN_KVB=64: 64-key KV blocks (128 TMEM columns, 64 KB smem: three CTAs per SM at head_dim <= 64)."""
    _run({"MOLLY_ATTN_KVB": "64"}, "attention")


def test_128_key_blocks_attention():
    """MOLLY_ATTN_KVB=128: the 128-key kernel, whichever is the default."""
    _run({"MOLLY_ATTN_KVB": "128"}, "attention")


def test_concurrent_modalities():
    """MOLLY_CONCURRENT_MODALITIES=1: the DNA/RNA and the protein encoder run on two streams (two branches of the CUDA
    graph in the captured form); results must not change."""
    _run({"MOLLY_CONCURRENT_MODALITIES": "1"}, "golden or graph or embed_and_process or ids_on_device",
         files=("tests/test_gpu_path.py", "tests/test_gpu_graph.py", "tests/test_gpu_inputs.py"))


def test_encoder_backward_with_per_layer_recompute():
    """MOLLY_TRAIN_RECOMPUTE=1: the training tape keeps only each layer's fp32 input and the backward recomputes the layer
    (the default keeps the activations when they fit the memory budget)."""
    _run({"MOLLY_TRAIN_RECOMPUTE": "1"}, "backward", files=("tests/test_gpu_train.py",))


def test_wgrad_through_transposes():
    """MOLLY_WGRAD_TRANSPOSE=1: the first wgrad version (explicit bf16 transposes + the K-major GEMM)."""
    _run({"MOLLY_WGRAD_TRANSPOSE": "1"}, "linear_wgrad")


def test_backward_gemms_legacy_kernels():
    """MOLLY_WGRAD_LEGACY=1 + MOLLY_DGRAD_TRANSPOSE=1: the round-1 backward GEMMs (single-CTA 128x128 MN-major wgrad kernel,
    weight transposes + K-major dgrad) instead of the main tcgen05 kernel with MN-major operands and split-K."""
    _run({"MOLLY_WGRAD_LEGACY": "1", "MOLLY_DGRAD_TRANSPOSE": "1"}, "linear_wgrad or tiny", files=("tests/test_gpu_kernels.py", "tests/test_gpu_train.py"))


def test_training_step_without_graphs():
    """MOLLY_TRAIN_GRAPH=0: the encoder training forward / backward always launched eagerly (a live nn.Module stepped several
    times: the path-level training tests)."""
    _run({"MOLLY_TRAIN_GRAPH": "0"}, "(train or backward or grad) and not cuda_graphs",
         files=("tests/test_gpu_path.py", "tests/test_gpu_real_class.py"))


def test_residual_gemms_through_reduce_add():
    """MOLLY_RESID_REDUCE=1: the in-place residual GEMMs add (acc + bias) into the fp32 stream with TMA reduce-adds
    (EPI_BIAS_ACCUM) instead of loading the residual tile; the golden fixtures must still match."""
    _run({"MOLLY_RESID_REDUCE": "1"}, "golden", files=("tests/test_gpu_path.py",))
