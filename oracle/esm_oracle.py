"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py) -- fp32 CPU restatement of the
reference's omics-embedding hot path.

Every function cites the reference lines it follows.  ``REF`` = /root/reference,
``HF`` = transformers 5.5.0 ``models/esm/modeling_esm.py`` (the reference pins 4.53.0,
``REF/requirements.txt:21``; 5.5.0 is what is installed and what the goldens were
generated with).

Weights are plain ``dict[str, Tensor]`` keyed exactly like ``EsmForMaskedLM.state_dict()``
so that the same dict feeds (a) this oracle, (b) a real HF model in ``ref_import.py`` and
(c) the product's weight packer.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, asdict
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# --------------------------------------------------------------------------------------
# Encoder shapes (ESM-2 from public checkpoint configs, NT from model cards)
# --------------------------------------------------------------------------------------
@dataclass(frozen=True)
class EncoderSpec:
    name: str
    hidden_size: int
    num_hidden_layers: int
    num_attention_heads: int
    intermediate_size: int
    vocab_size: int
    pad_token_id: int = 1
    mask_token_id: int = 32
    position_embedding_type: str = "rotary"     # "rotary" | "absolute"
    max_position_embeddings: int = 1026
    ffn_type: str = "gelu"                      # "gelu" (bias, erf-GELU) | "glu" (gated SiLU; biases iff the weights have them)
    glu_gate_first: bool = True                 # NT-v2 variant switch: silu(x1) * x2 (True) or silu(x2) * x1
    token_dropout: bool = True
    emb_layer_norm_before: bool = False
    layer_norm_eps: float = 1e-5

    @property
    def head_dim(self) -> int:
        return self.hidden_size // self.num_attention_heads

    def as_dict(self) -> dict:
        return asdict(self)


def _esm2(name, h, L, H, F_):
    return EncoderSpec(name, h, L, H, F_, vocab_size=33, mask_token_id=32, position_embedding_type="rotary",
                       max_position_embeddings=1026, ffn_type="gelu", token_dropout=True, layer_norm_eps=1e-5)


def _ntv2(name, h, L, H, F_):
    return EncoderSpec(name, h, L, H, F_, vocab_size=4107, mask_token_id=2, position_embedding_type="rotary",
                       max_position_embeddings=2050, ffn_type="glu", token_dropout=False, layer_norm_eps=1e-12)


def _ntv1(name, h, L, H, F_):
    return EncoderSpec(name, h, L, H, F_, vocab_size=4105, mask_token_id=2, position_embedding_type="absolute",
                       max_position_embeddings=1002, ffn_type="gelu", token_dropout=False, layer_norm_eps=1e-12)


SPECS: Dict[str, EncoderSpec] = {
    s.name: s for s in [
        _esm2("esm2_t6_8m", 320, 6, 20, 1280),
        _esm2("esm2_t33_650m", 1280, 33, 20, 5120),
        _ntv2("nt_v2_50m", 512, 12, 16, 2048),
        _ntv2("nt_v2_500m", 1024, 29, 16, 4096),
        _ntv1("nt_v1_2p5b", 2560, 32, 20, 10240),
        # tiny shapes for fixtures / fast unit tests (same code paths, head_dim 16/32/64)
        _esm2("tiny_esm2", 64, 2, 4, 256),
        _ntv2("tiny_ntv2", 128, 2, 4, 256),
        _ntv1("tiny_ntv1", 128, 2, 2, 256),
    ]
}


# --------------------------------------------------------------------------------------
# Weight construction (random init of the named architecture; HF state_dict key names)
# --------------------------------------------------------------------------------------
def _bf16_round(t: Tensor) -> Tensor:
    return t.to(torch.bfloat16).to(torch.float32)


def init_encoder_weights(spec: EncoderSpec, seed: int = 0, std: float = 0.02,
                         bf16_exact: bool = True) -> Dict[str, Tensor]:
    """Random weights with ``EsmForMaskedLM.state_dict()`` key names (HF:161-186, 285-316, 365-431, 485-492).

    Linear / embedding weights ~ N(0, std) like HF ``_init_weights``; unlike HF, biases and LayerNorm
    affine parameters are *also* randomised so that parity tests exercise them.  With ``bf16_exact`` the
    matrices are rounded to bf16-representable values so the bf16 candidate and the fp32 oracle read
    identical weights and only activation rounding separates them.
    """
    g = torch.Generator().manual_seed(seed)
    h, Fi = spec.hidden_size, spec.intermediate_size
    rnd = lambda *shape, s=std: torch.randn(*shape, generator=g) * s
    q = _bf16_round if bf16_exact else (lambda t: t)
    W: Dict[str, Tensor] = {}
    emb = rnd(spec.vocab_size, h)
    emb[spec.pad_token_id].zero_()                        # nn.Embedding(padding_idx=pad) (HF:168)
    W["esm.embeddings.word_embeddings.weight"] = q(emb)
    if spec.position_embedding_type == "absolute":
        pe = rnd(spec.max_position_embeddings, h)
        pe[spec.pad_token_id].zero_()                     # HF:182-184
        W["esm.embeddings.position_embeddings.weight"] = q(pe)
    if spec.emb_layer_norm_before:
        W["esm.embeddings.layer_norm.weight"] = 1.0 + rnd(h, s=0.1)
        W["esm.embeddings.layer_norm.bias"] = rnd(h, s=0.05)
    for i in range(spec.num_hidden_layers):
        p = f"esm.encoder.layer.{i}."
        W[p + "attention.LayerNorm.weight"] = 1.0 + rnd(h, s=0.1)
        W[p + "attention.LayerNorm.bias"] = rnd(h, s=0.05)
        for nm in ("query", "key", "value"):
            W[p + f"attention.self.{nm}.weight"] = q(rnd(h, h))
            W[p + f"attention.self.{nm}.bias"] = rnd(h)
        W[p + "attention.output.dense.weight"] = q(rnd(h, h))
        W[p + "attention.output.dense.bias"] = rnd(h)
        W[p + "LayerNorm.weight"] = 1.0 + rnd(h, s=0.1)
        W[p + "LayerNorm.bias"] = rnd(h, s=0.05)
        if spec.ffn_type == "glu":
            W[p + "intermediate.dense.weight"] = q(rnd(2 * Fi, h))
            W[p + "output.dense.weight"] = q(rnd(h, Fi))
        else:
            W[p + "intermediate.dense.weight"] = q(rnd(Fi, h))
            W[p + "intermediate.dense.bias"] = rnd(Fi)
            W[p + "output.dense.weight"] = q(rnd(h, Fi))
            W[p + "output.dense.bias"] = rnd(h)
    W["esm.encoder.emb_layer_norm_after.weight"] = 1.0 + rnd(h, s=0.1)
    W["esm.encoder.emb_layer_norm_after.bias"] = rnd(h, s=0.05)
    return W


def init_projector(h_enc: int, d_llm: int, seed: int = 0, bf16_exact: bool = True) -> Dict[str, Tensor]:
    """``nn.Linear(h_enc, d_llm)`` state dict (REF/src/model/omics_one.py:22-30): weight [D,h], bias [D]."""
    g = torch.Generator().manual_seed(seed)
    w = torch.randn(d_llm, h_enc, generator=g) / math.sqrt(h_enc)
    b = torch.randn(d_llm, generator=g) * 0.02
    if bf16_exact:
        w = _bf16_round(w)
    return {"weight": w, "bias": b}


# --------------------------------------------------------------------------------------
# Encoder forward  (third-party arithmetic the reference executes: HF EsmForMaskedLM)
# --------------------------------------------------------------------------------------
def gelu_erf(x: Tensor) -> Tensor:
    """HF:57-61 -- exact erf GELU, ``x * 0.5 * (1 + erf(x / sqrt(2)))``."""
    return x * 0.5 * (1.0 + torch.erf(x / math.sqrt(2.0)))


def rotate_half(x: Tensor) -> Tensor:
    """HF:43-45."""
    x1, x2 = x.chunk(2, dim=-1)
    return torch.cat((-x2, x1), dim=-1)


def rotary_tables(seq_len: int, dim: int, dtype=torch.float32) -> Tuple[Tensor, Tensor]:
    """HF:81-115 -- inv_freq = 10000^(-2i/d); angle = row index * inv_freq; cat(freqs, freqs)."""
    inv_freq = 1.0 / (10000 ** (torch.arange(0, dim, 2, dtype=torch.int64).float() / dim))
    t = torch.arange(seq_len).to(inv_freq.dtype)
    freqs = torch.outer(t, inv_freq)
    emb = torch.cat((freqs, freqs), dim=-1)
    return emb.cos().to(dtype), emb.sin().to(dtype)


def position_ids_from_input_ids(ids: Tensor, pad_idx: int) -> Tensor:
    """HF:971-984 -- cumsum(mask) * mask + pad_idx."""
    mask = ids.ne(pad_idx).int()
    return (torch.cumsum(mask, dim=1).type_as(mask) * mask).long() + pad_idx


def esm_embeddings(spec: EncoderSpec, W: Dict[str, Tensor], ids: Tensor, mask: Tensor) -> Tensor:
    """HF:189-236 (EsmEmbeddings.forward)."""
    x = W["esm.embeddings.word_embeddings.weight"][ids]
    if spec.token_dropout:
        is_mask_tok = ids == spec.mask_token_id
        x = x.masked_fill(is_mask_tok.unsqueeze(-1), 0.0)
        mask_ratio_train = 0.15 * 0.8
        src_lengths = mask.sum(-1)
        mask_ratio_observed = is_mask_tok.sum(-1).float() / src_lengths
        x = (x * (1 - mask_ratio_train) / (1 - mask_ratio_observed)[:, None, None]).to(x.dtype)
    if spec.position_embedding_type == "absolute":
        pos = position_ids_from_input_ids(ids, spec.pad_token_id)
        x = x + W["esm.embeddings.position_embeddings.weight"][pos]
    if spec.emb_layer_norm_before:
        x = F.layer_norm(x, (spec.hidden_size,), W["esm.embeddings.layer_norm.weight"],
                         W["esm.embeddings.layer_norm.bias"], spec.layer_norm_eps)
    x = (x * mask.unsqueeze(-1)).to(x.dtype)
    return x


def esm_self_attention(spec: EncoderSpec, W: Dict[str, Tensor], p: str, x_ln: Tensor, key_mask: Tensor) -> Tensor:
    """HF:318-362 (EsmSelfAttention.forward) + HF:257-282 (eager attention), eval mode."""
    N, K, h = x_ln.shape
    H, d = spec.num_attention_heads, spec.head_dim
    lin = lambda nm: F.linear(x_ln, W[p + f"attention.self.{nm}.weight"], W[p + f"attention.self.{nm}.bias"])
    q = lin("query").view(N, K, H, d).transpose(1, 2)
    k = lin("key").view(N, K, H, d).transpose(1, 2)
    v = lin("value").view(N, K, H, d).transpose(1, 2)
    q = q * d ** -0.5                                           # HF:341 scale BEFORE rotary
    if spec.position_embedding_type == "rotary":
        cos, sin = (t.to(q.device) for t in rotary_tables(K, d, q.dtype))   # (tests also run this checker on a GPU)
        q = q * cos + rotate_half(q) * sin                      # HF:48-54
        k = k * cos + rotate_half(k) * sin
    scores = torch.matmul(q, k.transpose(2, 3))                 # scaling = 1.0 (HF:315)
    # HF:679-709 create_bidirectional_mask -> additive finfo.min at padded KEYS only (queries all computed)
    add = torch.zeros(N, 1, 1, K, dtype=scores.dtype, device=scores.device)
    add.masked_fill_(~key_mask.bool()[:, None, None, :], torch.finfo(scores.dtype).min)
    probs = torch.softmax(scores + add, dim=-1)
    out = torch.matmul(probs, v).transpose(1, 2).reshape(N, K, h)
    return out


def esm_layer(spec: EncoderSpec, W: Dict[str, Tensor], i: int, x: Tensor, key_mask: Tensor) -> Tensor:
    """HF:446-482 (EsmLayer), HF:386-403 (EsmAttention pre-LN), HF:365-375, 406-427."""
    p = f"esm.encoder.layer.{i}."
    h = spec.hidden_size
    x_ln = F.layer_norm(x, (h,), W[p + "attention.LayerNorm.weight"], W[p + "attention.LayerNorm.bias"],
                        spec.layer_norm_eps)
    a = esm_self_attention(spec, W, p, x_ln, key_mask)
    a = F.linear(a, W[p + "attention.output.dense.weight"], W[p + "attention.output.dense.bias"]) + x
    a_ln = F.layer_norm(a, (h,), W[p + "LayerNorm.weight"], W[p + "LayerNorm.bias"], spec.layer_norm_eps)
    if spec.ffn_type == "glu":
        # NT-v2 remote code (PARITY UNPINNED): dense(h -> 2F, no bias); x1, x2 = split halves; silu(x1) * x2
        u = F.linear(a_ln, W[p + "intermediate.dense.weight"], W.get(p + "intermediate.dense.bias"))
        x1, x2 = u.split(u.size(-1) // 2, dim=-1)
        mid = F.silu(x1) * x2 if spec.glu_gate_first else F.silu(x2) * x1
        y = F.linear(mid, W[p + "output.dense.weight"], W.get(p + "output.dense.bias")) + a
    else:
        mid = gelu_erf(F.linear(a_ln, W[p + "intermediate.dense.weight"], W[p + "intermediate.dense.bias"]))
        y = F.linear(mid, W[p + "output.dense.weight"], W[p + "output.dense.bias"]) + a
    return y


def esm_encoder_forward(spec: EncoderSpec, W: Dict[str, Tensor], ids: Tensor,
                        return_all: bool = False):
    """``EsmForMaskedLM(ids, attention_mask=(ids != 1), output_hidden_states=True).hidden_states[-1]``

    i.e. the post-``emb_layer_norm_after`` hidden state (HF:494-514; REF omics_one.py:70-91).  The LM head
    (HF:797-815) is not computed: the reference discards its output.
    """
    mask = (ids != 1).long()                                     # REF omics_one.py:70 hard-codes pad id 1
    x = esm_embeddings(spec, W, ids, mask)
    hs = [x]
    for i in range(spec.num_hidden_layers):
        x = esm_layer(spec, W, i, x, mask)
        hs.append(x)
    x = F.layer_norm(x, (spec.hidden_size,), W["esm.encoder.emb_layer_norm_after.weight"],
                     W["esm.encoder.emb_layer_norm_after.bias"], spec.layer_norm_eps)
    return (x, hs) if return_all else x


# --------------------------------------------------------------------------------------
# The boundary: OmicsOne.process_omic_sequences (REF/src/model/omics_one.py:49-136)
# --------------------------------------------------------------------------------------
@dataclass
class OracleModality:
    spec: EncoderSpec
    weights: Dict[str, Tensor]
    projector: Dict[str, Tensor]        # {"weight": [D,h], "bias": [D]}
    project_token_num: int              # config.*_project_token_num  (K cap)


def process_omic_sequences(hidden_states: Tensor, omic_ids_list, omic_info_list,
                           dna_rna: Optional[OracleModality], protein: Optional[OracleModality]) -> Tensor:
    """Restates REF omics_one.py:49-136 line by line (in place; returns the same tensor object)."""

    def _inject(ids: List[Tensor], mappings: List[Tuple[int, int, int]], mod: Optional[OracleModality]):
        if not ids:                                              # :67-68
            return
        padded = torch.stack(ids, dim=0)                         # :69
        assert (padded < mod.spec.vocab_size).all(), \
            f"out-of-range token: {padded[padded >= mod.spec.vocab_size]}"      # :71-72
        try:
            last = esm_encoder_forward(mod.spec, mod.weights, padded)            # :73-88
        except Exception as e:                                   # :89-90
            raise RuntimeError(f"Error processing omic sequences: {e}")
        emb = F.linear(last, mod.projector["weight"], mod.projector["bias"])    # :91
        for idx, (b, start_pos, _) in enumerate(mappings):       # :93-97
            if start_pos == -1:
                continue
            k = min(mod.project_token_num, emb.size(1))
            hidden_states[b, start_pos + 1:start_pos + 1 + k] = emb[idx, :k].to(hidden_states.dtype)

    batch_size = hidden_states.shape[0]
    nt_ids, nt_map, pr_ids, pr_map = [], [], [], []
    for b in range(batch_size):                                  # :104-118
        for omic_id, info in zip(omic_ids_list[b], omic_info_list[b]):
            t, start = info["type"], info["start"]
            if t in ("dna", "rna"):
                nt_ids.append(omic_id)
                nt_map.append((b, start, len(omic_id)))
            elif t == "protein":
                pr_ids.append(omic_id)
                pr_map.append((b, start, len(omic_id)))
            elif t == "pad":
                continue
            else:
                raise ValueError(f"Unsupported omic type: {t}")
    _inject(nt_ids, nt_map, dna_rna)                             # :120-126  (NT first)
    _inject(pr_ids, pr_map, protein)                             # :128-134
    return hidden_states
