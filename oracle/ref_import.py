"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py) -- run the reference's OWN code in this container.

Only usable where ``/root/reference`` exists (the build container; NOT the GPU box).  It is used by
``make_golden.py`` to generate the committed fixtures and by ``tests/test_oracle_vs_reference.py`` to pin
``esm_oracle.py`` against the reference.  Nothing here is imported at bench / smoke / gpu-test time.

The reference module ``/root/reference/src/model/omics_one.py`` is executed unmodified.  Two non-arithmetic
imports it makes are absent from this image and are stubbed:
  * ``utils.tools.time_count``  (pulls in deepspeed; only used around ``nn.Linear`` construction, :22,27)
  * ``trainer.CausalLMOutputWithPast``  (only a return-type annotation, :154)
"""
from __future__ import annotations

import contextlib
import importlib.util
import os
import sys
import types
from typing import Dict

import torch

from .esm_oracle import EncoderSpec

REFERENCE_ROOT = "/root/reference"
# A copy of the ONE reference file this path executes, made by ``__graft_entry__.build()`` where /root/reference exists.
# oracle/_ref/ is git-ignored (reference sources never enter the history) but travels to the GPU box with gpurun, so the
# reference arm of bench.py and the real-class install test can run the reference's own code there.
LOCAL_COPY = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "omics_one.py")
_REF_MOD = None


def reference_file():
    for p in (os.path.join(REFERENCE_ROOT, "src/model/omics_one.py"), LOCAL_COPY):
        if os.path.isfile(p):
            return p
    return None


def reference_available() -> bool:
    return reference_file() is not None


def stage_reference_copy() -> bool:
    """build(): copy /root/reference/src/model/omics_one.py -> oracle/_ref/omics_one.py (unmodified)."""
    src = os.path.join(REFERENCE_ROOT, "src/model/omics_one.py")
    if not os.path.isfile(src):
        return os.path.isfile(LOCAL_COPY)
    import shutil
    os.makedirs(os.path.dirname(LOCAL_COPY), exist_ok=True)
    shutil.copyfile(src, LOCAL_COPY)
    return True


def load_reference_module():
    global _REF_MOD
    if _REF_MOD is not None:
        return _REF_MOD
    from transformers.modeling_outputs import CausalLMOutputWithPast
    u, t, tr = types.ModuleType("utils"), types.ModuleType("utils.tools"), types.ModuleType("trainer")
    t.time_count = contextlib.contextmanager(lambda name="block": (yield))
    u.tools = t
    tr.CausalLMOutputWithPast = CausalLMOutputWithPast
    saved = {k: sys.modules.get(k) for k in ("utils", "utils.tools", "trainer")}
    sys.modules.update({"utils": u, "utils.tools": t, "trainer": tr})
    try:
        spec = importlib.util.spec_from_file_location("ref_omics_one", reference_file())
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    _REF_MOD = mod
    return mod


def build_hf_encoder(spec: EncoderSpec, weights: Dict[str, torch.Tensor]):
    """Stock ``EsmForMaskedLM`` (eager, fp32, eval) carrying ``weights``.  For ``ffn_type == "glu"`` the two FFN
    modules are swapped for the NT-v2 gated form (restated from the public model card: PARITY UNPINNED)."""
    from transformers import EsmConfig, EsmForMaskedLM
    cfg = EsmConfig(
        vocab_size=spec.vocab_size, mask_token_id=spec.mask_token_id, pad_token_id=spec.pad_token_id,
        hidden_size=spec.hidden_size, num_hidden_layers=spec.num_hidden_layers,
        num_attention_heads=spec.num_attention_heads, intermediate_size=spec.intermediate_size,
        max_position_embeddings=spec.max_position_embeddings, layer_norm_eps=spec.layer_norm_eps,
        position_embedding_type=spec.position_embedding_type, token_dropout=spec.token_dropout,
        emb_layer_norm_before=spec.emb_layer_norm_before, hidden_dropout_prob=0.0,
        attention_probs_dropout_prob=0.0, attn_implementation="eager")
    model = EsmForMaskedLM(cfg)
    if spec.ffn_type == "glu":
        import torch.nn as nn

        class _GluIntermediate(nn.Module):
            def __init__(self, h, f):
                super().__init__()
                self.dense = nn.Linear(h, 2 * f, bias=False)

            def forward(self, x):
                x1, x2 = self.dense(x).split(self.dense.out_features // 2, dim=-1)
                return torch.nn.functional.silu(x1) * x2

        class _GluOutput(nn.Module):
            def __init__(self, h, f):
                super().__init__()
                self.dense = nn.Linear(f, h, bias=False)

            def forward(self, x, inp):
                return self.dense(x) + inp

        for layer in model.esm.encoder.layer:
            layer.intermediate = _GluIntermediate(spec.hidden_size, spec.intermediate_size)
            layer.output = _GluOutput(spec.hidden_size, spec.intermediate_size)
    if weights:
        missing, unexpected = model.load_state_dict(weights, strict=False)
        assert not unexpected, unexpected
        bad = [k for k in missing if not (k.startswith("lm_head") or "inv_freq" in k or "contact_head" in k
                                          or "position_ids" in k)]
        assert not bad, f"weights missing for {bad}"
    return model.float().eval()


def build_reference_omics_random(nt_spec: EncoderSpec, pr_spec: EncoderSpec, d_llm: int, k_tokens: int, seed: int = 0):
    """The reference's ``OmicsOne`` with stock HF encoders in their own random init (timing only: bench.py --impl reference)."""
    ref = load_reference_module()
    cfg = types.SimpleNamespace(
        text_config=types.SimpleNamespace(hidden_size=d_llm, use_return_dict=True),
        dna_rna_config=types.SimpleNamespace(hidden_size=nt_spec.hidden_size),
        protein_config=types.SimpleNamespace(hidden_size=pr_spec.hidden_size),
        dna_rna_project_token_num=k_tokens, protein_project_token_num=k_tokens)
    torch.manual_seed(seed)
    om = ref.OmicsOne(cfg)
    om.dna_rna_model = build_hf_encoder(nt_spec, {})
    om.protein_model = build_hf_encoder(pr_spec, {})
    return om.eval()


def build_reference_omics(nt, pr, d_llm: int):
    """``nt`` / ``pr`` are ``oracle.esm_oracle.OracleModality``.  Returns the reference's ``OmicsOne`` with HF encoders."""
    ref = load_reference_module()
    cfg = types.SimpleNamespace(
        text_config=types.SimpleNamespace(hidden_size=d_llm, use_return_dict=True),
        dna_rna_config=types.SimpleNamespace(hidden_size=nt.spec.hidden_size),
        protein_config=types.SimpleNamespace(hidden_size=pr.spec.hidden_size),
        dna_rna_project_token_num=nt.project_token_num,
        protein_project_token_num=pr.project_token_num)
    om = ref.OmicsOne(cfg)
    om.dna_rna_model = build_hf_encoder(nt.spec, nt.weights)
    om.protein_model = build_hf_encoder(pr.spec, pr.weights)
    om.dna_rna_projector.load_state_dict(nt.projector)
    om.protein_projector.load_state_dict(pr.projector)
    return om.eval()
