"""Encoder training path (``--train-bio``: src/utils/tools.py:313-331 unfreezes the encoders).

Forward = the same sm_100a kernels as the inference path, launched layer by layer so that each layer's fp32 input can be
kept; backward = per-layer recompute from that input, then the autograd of every HF module the kernels replace:

    nn.Linear      dgrad: tcgen05 GEMM on the transposed weight        wgrad / bias: ``molly_linear_wgrad``
    attention      ``molly_attention_bwd`` (dQ / dK,dV kernels) from the forward's row log-sum-exp
    LayerNorm      ``molly_layernorm_bwd``          GELU / gated SiLU   ``molly_act_fwd_bwd``
    rotary         the forward kernel with -sin (inverse rotation), then the q scale
    embeddings     ``molly_scatter_add_rows`` (token-dropout scale, pad / <mask> rows skipped like the forward)

Gradients come back keyed by the HF ``state_dict`` names of the encoder (q / k / v and the NT-v2 gate / up halves are
un-packed), fp32.  Orchestration is Python (one ctypes call per kernel): this is the training path, not the hot path.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import torch

from . import _lib, ops
from .packing import PackedEncoder, glu_deinterleave


def _emb_meta(enc: PackedEncoder, ids: torch.Tensor):
    """Per-row scatter indices / scales of the embedding backward, from the ids alone (tiny integer work, HF:189-236)."""
    cfg = enc.cfg
    mask = ids != 1                       # the attention / padding mask is the reference's literal 1 (omics_one.py:70), as in
    word_scale = mask.float()
    if cfg.token_dropout:
        is_mask_tok = ids == cfg.mask_token_id
        ratio = is_mask_tok.sum(-1).float() / mask.sum(-1).clamp(min=1).float()
        word_scale = word_scale * (~is_mask_tok).float() * ((1 - 0.15 * 0.8) / (1 - ratio))[:, None]
    pos_index = None
    if cfg.position_embedding_type == "absolute":
        m = mask.int()
        # the forward kernels (rowwise.cu); cfg.pad_token_id is only HF's padding_idx offset of the position ids (HF:971-984)
        pos_index = ((torch.cumsum(m, dim=1) * m).long() + cfg.pad_token_id).to(torch.int32).reshape(-1).contiguous()
    return (ids.to(torch.int32).reshape(-1).contiguous(), word_scale.reshape(-1).contiguous(), pos_index,
            mask.float().reshape(-1).contiguous())


def reduce_schedule(enc: PackedEncoder) -> List[List[Tuple[str, Tuple[int, ...]]]]:
    """The gradient groups ``_InjectTrainFn.backward`` hands to a ``LayerwiseGradReducer``, in order, as (HF name, shape)
    lists: the projector, the encoder layers L-1 .. 0, then final LayerNorm + embeddings.  ``encoder_backward`` reduces
    exactly these lists; a rank whose micro-batch lacks the modality all-reduces zeros laid out the same way
    (``omics_path._AbsentModalityFn``) and so receives the other ranks' averaged gradients."""
    cfg = enc.cfg
    h, F = cfg.hidden_size, cfg.intermediate_size
    glu = cfg.ffn_type == "glu"
    groups = [[("projector.weight", tuple(enc.proj_w.shape)), ("projector.bias", tuple(enc.proj_b.shape))]]
    for i in range(cfg.num_hidden_layers - 1, -1, -1):
        p = f"esm.encoder.layer.{i}."
        ffn_bias = enc.layer_tensors[i]["b_ffn1"] is not None
        g = [(p + "LayerNorm.weight", (h,)), (p + "LayerNorm.bias", (h,)), (p + "output.dense.weight", (h, F)),
             (p + "intermediate.dense.weight", ((2 * F if glu else F), h))]
        if ffn_bias:
            g += [(p + "intermediate.dense.bias", ((2 * F if glu else F),)), (p + "output.dense.bias", (h,))]
        g += [(p + "attention.LayerNorm.weight", (h,)), (p + "attention.LayerNorm.bias", (h,)),
              (p + "attention.output.dense.weight", (h, h)), (p + "attention.output.dense.bias", (h,))]
        for nm in ("query", "key", "value"):
            g += [(p + f"attention.self.{nm}.weight", (h, h)), (p + f"attention.self.{nm}.bias", (h,))]
        groups.append(g)
    last = [("esm.encoder.emb_layer_norm_after.weight", (h,)), ("esm.encoder.emb_layer_norm_after.bias", (h,)),
            ("esm.embeddings.word_embeddings.weight", tuple(enc.word_emb.shape))]
    if enc.pos_emb is not None:
        last.append(("esm.embeddings.position_embeddings.weight", tuple(enc.pos_emb.shape)))
    groups.append(last)
    return groups


class EncoderTape:
    """What the training forward keeps for the backward.  ``saved[i]`` holds layer i's activations (ln1, qkv, attn, lse2,
    x_mid, ln2, FFN pre-activation) when they fit the memory budget; otherwise only the fp32 layer input is kept and the
    backward recomputes the layer (activation checkpointing per layer)."""

    def __init__(self):
        self.layer_inputs: List[torch.Tensor] = []
        self.saved: List[tuple] = []
        self.x_final = None
        self.kv_info = self.key_mask = self.ids = None


def _save_activations(enc: PackedEncoder, n_tokens: int) -> bool:
    """Keep every layer's activations (no recompute) when they take less than 35 % of the free device memory;
    MOLLY_TRAIN_RECOMPUTE=1 / 0 forces per-layer recompute / saving."""
    import os
    env = os.environ.get("MOLLY_TRAIN_RECOMPUTE")
    if env is not None:
        return env == "0"
    cfg = enc.cfg
    f_pre = cfg.intermediate_size * (2 if cfg.ffn_type == "glu" else 1)
    per_layer = n_tokens * (cfg.hidden_size * (4 + 2 + 6 + 2 + 4 + 2) + f_pre * 2)
    return per_layer * cfg.num_hidden_layers < 0.35 * torch.cuda.mem_get_info(enc.device)[0]


def encoder_forward_train(enc: PackedEncoder, ids: torch.Tensor) -> Tuple[torch.Tensor, EncoderTape]:
    """``hidden_states[-1]`` (bf16 [n*K, h]) with the tape; numerically the inference forward (same kernels, same order)."""
    cfg, L = enc.cfg, _lib
    if cfg.emb_layer_norm_before:
        raise NotImplementedError("emb_layer_norm_before encoders are not covered by the training path")
    n_seq, K = ids.shape
    tape = EncoderTape()
    tape.ids = ids
    x, tape.kv_info, tape.key_mask = ops.embed(ids, enc.c_config, enc.word_emb, enc.pos_emb)
    save = _save_activations(enc, n_seq * K)
    for lt in enc.layer_tensors:
        tape.layer_inputs.append(x.clone())
        if save:
            kept = _layer_forward(enc, lt, x, n_seq, K, tape.kv_info, tape.key_mask, keep=True)
            tape.saved.append(kept[1:])
            mid, _ = ops.act_fwd_bwd(cfg.ffn_type == "glu", kept[7], None)       # finish the layer: FFN activation + FFN2
            ops.gemm_bf16(mid, lt["w_ffn2"], L.EPI_BIAS_RESIDUAL, bias=lt["b_ffn2"], residual=x, out=x)
        else:
            x = _layer_forward(enc, lt, x, n_seq, K, tape.kv_info, tape.key_mask)[0]
    tape.x_final = x
    out = ops.layernorm(x, enc.final_ln_w, enc.final_ln_b, cfg.layer_norm_eps, torch.bfloat16)
    return out, tape


def _layer_forward(enc, lt, x, n_seq, K, kv_info, key_mask, keep: bool = False):
    """One pre-LN layer on the fp32 stream ``x`` (updated in place).  ``keep``: also return what the backward needs."""
    cfg, L = enc.cfg, _lib
    h, H = cfg.hidden_size, cfg.num_attention_heads
    d = h // H
    rope = cfg.position_embedding_type == "rotary"
    ln1 = ops.layernorm(x, lt["ln1_w"], lt["ln1_b"], cfg.layer_norm_eps, torch.bfloat16)
    if rope and d <= 64:
        qkv = ops.gemm_bf16(ln1, lt["w_qkv"], L.EPI_BIAS_ROPE, bias=lt["b_qkv"], seq_k=K, scale_cols=h, scale=d ** -0.5,
                            rope_cos_t=enc.rope_cos_t, rope_sin_t=enc.rope_sin_t, rope_cols=2 * h, rope_head_dim=d)
    else:
        qkv = ops.gemm_bf16(ln1, lt["w_qkv"], L.EPI_BIAS, bias=lt["b_qkv"], scale_cols=h, scale=d ** -0.5)
        if rope:
            ops.rotary_(qkv, K, H, enc.rope_cos, enc.rope_sin)
    attn, lse2 = ops.attention_lse(qkv, n_seq, K, H, kv_info, key_mask)
    ops.gemm_bf16(attn, lt["w_attn_out"], L.EPI_BIAS_RESIDUAL, bias=lt["b_attn_out"], residual=x, out=x)
    x_mid = x.clone() if keep else None
    ln2 = ops.layernorm(x, lt["ln2_w"], lt["ln2_b"], cfg.layer_norm_eps, torch.bfloat16)
    glu = cfg.ffn_type == "glu"
    if keep:                                   # the backward needs the pre-activation: plain bias epilogue, activation apart
        pre = ops.gemm_bf16(ln2, lt["w_ffn1"], L.EPI_BIAS, bias=lt["b_ffn1"])
        return x, ln1, qkv, attn, lse2, x_mid, ln2, pre
    mid = ops.gemm_bf16(ln2, lt["w_ffn1"], L.EPI_GLU if glu else L.EPI_BIAS_GELU, bias=lt["b_ffn1"])
    ops.gemm_bf16(mid, lt["w_ffn2"], L.EPI_BIAS_RESIDUAL, bias=lt["b_ffn2"], residual=x, out=x)
    return (x,)


def encoder_backward(enc: PackedEncoder, tape: EncoderTape, d_out: torch.Tensor, reducer=None) -> Dict[str, torch.Tensor]:
    """Gradients of every encoder parameter (HF ``state_dict`` names, fp32) from ``d_out`` = d(loss)/d(hidden_states[-1]),
    bf16 [n*K, h].  ``reducer`` (``dist.LayerwiseGradReducer``): each layer's gradients are handed to it as soon as they
    exist, so their all-reduce overlaps the layers below; they come back averaged, in the reducer's dtype."""
    cfg, L = enc.cfg, _lib
    n_seq, K = tape.ids.shape
    h, H, F = cfg.hidden_size, cfg.num_attention_heads, cfg.intermediate_size
    d = h // H
    dev = d_out.device
    glu = cfg.ffn_type == "glu"
    rope = cfg.position_embedding_type == "rotary"
    grads: Dict[str, torch.Tensor] = {}
    # LayerNorm weight / bias gradients accumulate into pre-zeroed fp32 rows: one buffer (one fill) for the whole backward
    zero_rows = torch.zeros(4 * cfg.num_hidden_layers + 2, h, dtype=torch.float32, device=dev)
    zero_next = [0]

    def z(n):
        assert n == h
        row = zero_rows[zero_next[0]]
        zero_next[0] += 1
        return row
    schedule = reduce_schedule(enc) if reducer is not None else None     # [0] projector, [1 + (L-1-i)] layer i, [-1] the rest

    # emb_layer_norm_after
    d_x = torch.empty_like(tape.x_final)
    g_w, g_b = z(h), z(h)
    ops.layernorm_bwd(tape.x_final, d_out, enc.final_ln_w, cfg.layer_norm_eps, d_x, False, g_w, g_b)
    grads["esm.encoder.emb_layer_norm_after.weight"], grads["esm.encoder.emb_layer_norm_after.bias"] = g_w, g_b

    for i in range(cfg.num_hidden_layers - 1, -1, -1):
        lt = enc.layer_tensors[i]
        p = f"esm.encoder.layer.{i}."
        x_in = tape.layer_inputs[i]
        if tape.saved:
            ln1, qkv, attn, lse2, x_mid, ln2, pre = tape.saved[i]
            tape.saved[i] = None                                              # release as we go
        else:
            _, ln1, qkv, attn, lse2, x_mid, ln2, pre = _layer_forward(enc, lt, x_in.clone(), n_seq, K, tape.kv_info,
                                                                       tape.key_mask, keep=True)
        # ---- feed-forward block: x_out = x_mid + W2 act(W1 LN2(x_mid) + b1) + b2
        dy = ops.cast_bf16(d_x)
        d_act = ops.gemm_bf16(dy, ops.transpose_bf16(lt["w_ffn2"]), L.EPI_BIAS)
        act, d_pre = ops.act_fwd_bwd(glu, pre, d_act)
        ffn_bias = lt["b_ffn1"] is not None                                     # NT-v2's gated FFN has none (add_bias_fnn = False)
        dW2, db2 = ops.linear_wgrad(dy, act, with_bias=ffn_bias)
        dW1, db1 = ops.linear_wgrad(d_pre, ln2, with_bias=ffn_bias)
        d_ln2 = ops.gemm_bf16(d_pre, ops.transpose_bf16(lt["w_ffn1"]), L.EPI_BIAS)
        g_w, g_b = z(h), z(h)
        ops.layernorm_bwd(x_mid, d_ln2, lt["ln2_w"], cfg.layer_norm_eps, d_x, True, g_w, g_b)      # d_x is now d(x_mid)
        grads[p + "LayerNorm.weight"], grads[p + "LayerNorm.bias"] = g_w, g_b
        grads[p + "output.dense.weight"] = dW2
        if glu:                                # packed rows are (gate_0, up_0, gate_1, up_1, ...): back to the HF halves
            grads[p + "intermediate.dense.weight"] = glu_deinterleave(dW1, cfg.glu_gate_first)
            if ffn_bias:
                grads[p + "intermediate.dense.bias"] = glu_deinterleave(db1, cfg.glu_gate_first)
        else:
            grads[p + "intermediate.dense.weight"] = dW1
            grads[p + "intermediate.dense.bias"] = db1
        if ffn_bias:
            grads[p + "output.dense.bias"] = db2
        # ---- attention block: x_mid = x_in + Wo Attn(LN1(x_in)) + bo
        dy = ops.cast_bf16(d_x)
        d_attn = ops.gemm_bf16(dy, ops.transpose_bf16(lt["w_attn_out"]), L.EPI_BIAS)
        dWo, dbo = ops.linear_wgrad(dy, attn)
        d_qkv = ops.attention_bwd(qkv, attn, d_attn, lse2, n_seq, K, H, tape.kv_info, tape.key_mask)
        if rope:
            ops.rotary_(d_qkv, K, H, enc.rope_cos, enc.rope_neg_sin)          # inverse rotation of d(q'), d(k')
        ops.scale_cols_(d_qkv, h, d ** -0.5)                                  # q = (W_q x + b_q) * d^-1/2   (HF:341)
        dWqkv, dbqkv = ops.linear_wgrad(d_qkv, ln1)
        d_ln1 = ops.gemm_bf16(d_qkv, ops.transpose_bf16(lt["w_qkv"]), L.EPI_BIAS)
        g_w, g_b = z(h), z(h)
        ops.layernorm_bwd(x_in, d_ln1, lt["ln1_w"], cfg.layer_norm_eps, d_x, True, g_w, g_b)        # d_x is now d(x_in)
        grads[p + "attention.LayerNorm.weight"], grads[p + "attention.LayerNorm.bias"] = g_w, g_b
        grads[p + "attention.output.dense.weight"], grads[p + "attention.output.dense.bias"] = dWo, dbo
        for j, nm in enumerate(("query", "key", "value")):
            grads[p + f"attention.self.{nm}.weight"] = dWqkv[j * h:(j + 1) * h]
            grads[p + f"attention.self.{nm}.bias"] = dbqkv[j * h:(j + 1) * h]
        if reducer is not None:
            names = [n for n, _ in schedule[cfg.num_hidden_layers - i]]
            assert sorted(names) == sorted(n for n in grads if n.startswith(p)), "reduce_schedule is out of date"
            reducer.reduce_(grads, names)

    # ---- embeddings
    word_index, word_scale, pos_index, pos_scale = _emb_meta(enc, tape.ids)
    d_word = torch.zeros(enc.word_emb.shape, dtype=torch.float32, device=dev)
    ops.scatter_add_rows_(d_word, d_x, word_index, word_scale)
    grads["esm.embeddings.word_embeddings.weight"] = d_word
    if pos_index is not None:
        d_pos = torch.zeros(enc.pos_emb.shape, dtype=torch.float32, device=dev)
        ops.scatter_add_rows_(d_pos, d_x, pos_index, pos_scale)
        grads["esm.embeddings.position_embeddings.weight"] = d_pos
    if reducer is not None:
        names = [n for n, _ in schedule[-1]]
        assert sorted(names) == sorted(n for n in grads if not n.startswith("esm.encoder.layer.")), "reduce_schedule is out of date"
        reducer.reduce_(grads, names)
    return grads
