"""Weight ingestion: ``EsmForMaskedLM.state_dict()`` + projector ``nn.Linear.state_dict()`` -> packed device tensors and
a ``molly_encoder_t`` handle.

Keys consumed (the same three modules the reference calls, omics_one.py:18-30):
  esm.embeddings.word_embeddings.weight, [esm.embeddings.position_embeddings.weight], [esm.embeddings.layer_norm.*],
  esm.encoder.layer.{i}.attention.{LayerNorm, self.query|key|value, output.dense}.*,
  esm.encoder.layer.{i}.{LayerNorm, intermediate.dense, output.dense}.*, esm.encoder.emb_layer_norm_after.*;
  projector: weight [D, h], bias [D].
Packing done once: q/k/v stacked into one [3h, h] matrix (one GEMM instead of three), NT-v2 gate/up rows interleaved so
one accumulator tile holds both halves of each GLU pair, matrices cast to bf16, vectors to fp32, rotary cos/sin tables
in fp32 (the bf16 reference computes them in bf16; fp32 is closer to the fp32 oracle).
The LM head (``lm_head.*``) is ignored: the reference discards its output (omics_one.py:91).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Mapping, Optional

import torch

from . import _lib
from .config import EncoderConfig


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def glu_interleave(t: torch.Tensor, gate_first: bool = True) -> torch.Tensor:
    """``intermediate.dense`` weight [2F, h] (or bias [2F]) of the gated FFN -> rows (gate_0, up_0, gate_1, up_1, ...): one
    accumulator tile of the FFN1 GEMM then holds both halves of every pair and its epilogue computes silu(even) * odd.
    ``gate_first``: the SiLU half is rows [0, F) (x1 of ``x1, x2 = split``), else rows [F, 2F)."""
    f = t.shape[0] // 2
    gate, up = (t[:f], t[f:]) if gate_first else (t[f:], t[:f])
    return torch.stack([gate, up], dim=1).reshape(t.shape)


def glu_deinterleave(t: torch.Tensor, gate_first: bool = True) -> torch.Tensor:
    """Inverse of ``glu_interleave`` (gradients go back in the HF layout)."""
    f = t.shape[0] // 2
    pair = t.reshape(f, 2, *t.shape[1:])
    gate, up = pair[:, 0], pair[:, 1]
    return torch.cat([gate, up] if gate_first else [up, gate], dim=0).contiguous()


class PackedEncoder:
    """Owns the packed weights of one modality (encoder + projector) and the native handle built on them."""

    def __init__(self, cfg: EncoderConfig, state_dict: Mapping[str, torch.Tensor], projector: Mapping[str, torch.Tensor],
                 project_token_num: int, device: torch.device, rope_len: int = 4096):
        if device.type != "cuda":
            raise RuntimeError("molly_b200 runs on CUDA devices only (no CPU fallback)")
        self.cfg = cfg
        self.device = device
        self.project_token_num = int(project_token_num)
        self.llm_hidden_size = int(projector["weight"].shape[0])
        self._keep: List[torch.Tensor] = []
        self._handle = C.c_void_p()
        self._lib = _lib.load()
        h, L, Fi = cfg.hidden_size, cfg.num_hidden_layers, cfg.intermediate_size
        if projector["weight"].shape[1] != h:
            raise ValueError(f"projector in_features {projector['weight'].shape[1]} != encoder hidden {h}")

        self._recipes = []          # (kept tensor, builder(state_dict) -> source tensor): replayed by reload()

        def mat(build, parts=None):   # bf16 matrix on device
            o = build(state_dict).detach().to(device=device, dtype=torch.float32).to(torch.bfloat16).contiguous()
            self._keep.append(o)
            self._recipes.append((o, build, parts))
            return o

        def vec(build, parts=None):   # fp32 vector on device
            o = build(state_dict).detach().to(device=device, dtype=torch.float32).contiguous()
            self._keep.append(o)
            self._recipes.append((o, build, parts))
            return o

        def qkv_parts(p, kind):       # reload() copies q, k, v straight into their row blocks of the packed tensor (no cat)
            return lambda sd_: [sd_[p + f"attention.self.{n}.{kind}"] for n in ("query", "key", "value")]

        def key(name):
            return lambda sd_: sd_[name]

        def const(t):     # tables that do not depend on the checkpoint
            o = t.to(device=device, dtype=torch.float32).contiguous()
            self._keep.append(o)
            return o

        sd = state_dict
        w = _lib.EncoderWeights()
        self.word_emb = mat(key("esm.embeddings.word_embeddings.weight"))
        w.word_emb_dev = _ptr(self.word_emb)
        w.pos_emb_dev = None
        self.pos_emb = None
        if cfg.position_embedding_type == "absolute":
            self.pos_emb = mat(key("esm.embeddings.position_embeddings.weight"))
            w.pos_emb_dev = _ptr(self.pos_emb)
        w.emb_ln_w_dev = w.emb_ln_b_dev = None
        if cfg.emb_layer_norm_before:
            w.emb_ln_w_dev = _ptr(vec(key("esm.embeddings.layer_norm.weight")))
            w.emb_ln_b_dev = _ptr(vec(key("esm.embeddings.layer_norm.bias")))
        # rotary tables exactly as HF:81-115 builds them, in fp32
        d = cfg.head_dim
        self.rope_len = max(int(rope_len), int(cfg.max_position_embeddings))
        inv_freq = 1.0 / (10000 ** (torch.arange(0, d, 2, dtype=torch.int64).float() / d))
        freqs = torch.outer(torch.arange(self.rope_len).float(), inv_freq)
        self.rope_cos, self.rope_sin = const(freqs.cos()), const(freqs.sin())
        self.rope_neg_sin = const(-freqs.sin())              # inverse rotation = rotary backward
        self.rope_cos_t, self.rope_sin_t = const(freqs.cos().t()), const(freqs.sin().t())
        w.rope_cos_dev, w.rope_sin_dev = _ptr(self.rope_cos), _ptr(self.rope_sin)
        w.rope_len = self.rope_len
        w.rope_cos_t_dev = _ptr(self.rope_cos_t)             # frequency-major copies for the fused QKV epilogue
        w.rope_sin_t_dev = _ptr(self.rope_sin_t)

        names = ["ln1_w", "ln1_b", "w_qkv", "b_qkv", "w_attn_out", "b_attn_out", "ln2_w", "ln2_b", "w_ffn1", "b_ffn1",
                 "w_ffn2", "b_ffn2"]
        arrays: Dict[str, List[Optional[int]]] = {n: [] for n in names}
        self.layer_tensors: List[Dict[str, Optional[torch.Tensor]]] = []     # named views of the packed weights (train.py)
        for i in range(L):
            p = f"esm.encoder.layer.{i}."
            arrays["ln1_w"].append(_ptr(vec(key(p + "attention.LayerNorm.weight"))))
            arrays["ln1_b"].append(_ptr(vec(key(p + "attention.LayerNorm.bias"))))
            arrays["w_qkv"].append(_ptr(mat(lambda sd_, p=p: torch.cat(
                [sd_[p + f"attention.self.{n}.weight"] for n in ("query", "key", "value")], dim=0), qkv_parts(p, "weight"))))
            arrays["b_qkv"].append(_ptr(vec(lambda sd_, p=p: torch.cat(
                [sd_[p + f"attention.self.{n}.bias"] for n in ("query", "key", "value")], dim=0), qkv_parts(p, "bias"))))
            arrays["w_attn_out"].append(_ptr(mat(key(p + "attention.output.dense.weight"))))
            arrays["b_attn_out"].append(_ptr(vec(key(p + "attention.output.dense.bias"))))
            arrays["ln2_w"].append(_ptr(vec(key(p + "LayerNorm.weight"))))
            arrays["ln2_b"].append(_ptr(vec(key(p + "LayerNorm.bias"))))
            has_bias = (p + "intermediate.dense.bias") in sd
            if cfg.ffn_bias is not None and bool(cfg.ffn_bias) != has_bias:
                raise ValueError(f"EncoderConfig.ffn_bias={cfg.ffn_bias} but the state dict "
                                 f"{'has' if has_bias else 'lacks'} {p}intermediate.dense.bias")
            if cfg.ffn_type != "glu" and not has_bias:
                raise ValueError(f"the GELU FFN needs {p}intermediate.dense.bias (HF:406-427)")
            if cfg.ffn_type == "glu":
                if sd[p + "intermediate.dense.weight"].shape[0] != 2 * Fi:
                    raise ValueError(f"GLU intermediate.dense.weight must be [2F, h], got "
                                     f"{tuple(sd[p + 'intermediate.dense.weight'].shape)}")
                gf = cfg.glu_gate_first
                arrays["w_ffn1"].append(_ptr(mat(lambda sd_, p=p: glu_interleave(sd_[p + "intermediate.dense.weight"], gf))))
                arrays["b_ffn1"].append(_ptr(vec(lambda sd_, p=p: glu_interleave(sd_[p + "intermediate.dense.bias"], gf)))
                                        if has_bias else None)
            else:
                arrays["w_ffn1"].append(_ptr(mat(key(p + "intermediate.dense.weight"))))
                arrays["b_ffn1"].append(_ptr(vec(key(p + "intermediate.dense.bias"))))
            arrays["w_ffn2"].append(_ptr(mat(key(p + "output.dense.weight"))))
            arrays["b_ffn2"].append(_ptr(vec(key(p + "output.dense.bias"))) if has_bias else None)
        by_ptr = {t.data_ptr(): t for t in self._keep}
        for i in range(L):
            self.layer_tensors.append({n: (by_ptr[arrays[n][i]] if arrays[n][i] is not None else None) for n in names})
        self._ptr_arrays = {}
        for n in names:
            arr = (C.c_void_p * L)(*arrays[n])
            self._ptr_arrays[n] = arr
            setattr(w, n + "_dev", C.cast(arr, _lib.c_void_pp))
        self.final_ln_w = vec(key("esm.encoder.emb_layer_norm_after.weight"))
        self.final_ln_b = vec(key("esm.encoder.emb_layer_norm_after.bias"))
        w.final_ln_w_dev, w.final_ln_b_dev = _ptr(self.final_ln_w), _ptr(self.final_ln_b)
        # projector buffers are refreshed in place when the nn.Linear trains (--train-mlp)
        self.proj_w = torch.empty(self.llm_hidden_size, h, dtype=torch.bfloat16, device=device)
        self.proj_b = torch.empty(self.llm_hidden_size, dtype=torch.float32, device=device)
        self.load_projector(projector["weight"], projector["bias"])
        w.w_proj_dev = _ptr(self.proj_w)
        w.b_proj_dev = _ptr(self.proj_b)

        c = _lib.EncoderConfig(
            hidden_size=h, num_layers=L, num_heads=cfg.num_attention_heads, intermediate_size=Fi,
            vocab_size=cfg.vocab_size, pad_token_id=cfg.pad_token_id, mask_token_id=cfg.mask_token_id,
            position_type=_lib.POS_ABSOLUTE if cfg.position_embedding_type == "absolute" else _lib.POS_ROTARY,
            max_positions=cfg.max_position_embeddings,
            ffn_type=_lib.FFN_GLU if cfg.ffn_type == "glu" else _lib.FFN_GELU,
            token_dropout=int(cfg.token_dropout), emb_layer_norm_before=int(cfg.emb_layer_norm_before),
            layer_norm_eps=cfg.layer_norm_eps, llm_hidden_size=self.llm_hidden_size,
            project_token_num=self.project_token_num)
        self.c_config = c
        self._weights_struct = w
        with torch.cuda.device(device):
            _lib.check(self._lib.molly_encoder_create(C.byref(c), C.byref(w), C.byref(self._handle)),
                       "molly_encoder_create")

    # ------------------------------------------------------------------
    @torch.no_grad()
    def reload(self, state_dict: Mapping[str, torch.Tensor]) -> None:
        """Refresh the packed ENCODER weights in place from a new state dict (same shapes): device addresses, the native
        handle and its cached TMA descriptors stay valid.  SURVEY.md 8b: packed copies are caches that must follow the
        ``nn.Module`` weights when the encoders train (``--train-bio``) or a checkpoint is loaded after construction."""
        same_d, same_s, conv_d, conv_s = [], [], [], []
        def put(dst, src):
            src = src.detach()
            if tuple(src.shape) != tuple(dst.shape):
                raise ValueError(f"reload: shape {tuple(src.shape)} does not match the packed {tuple(dst.shape)}")
            if src.device != dst.device:
                src = src.to(dst.device, non_blocking=True)
            (same_d if src.dtype == dst.dtype else conv_d).append(dst)
            (same_s if src.dtype == dst.dtype else conv_s).append(src)

        for dst, build, parts in self._recipes:
            if parts is None:
                put(dst, build(state_dict))
                continue
            row = 0
            for src in parts(state_dict):              # row blocks of dst, in order
                put(dst[row:row + src.shape[0]], src)
                row += src.shape[0]
            if row != dst.shape[0]:
                raise ValueError(f"reload: parts cover {row} of {dst.shape[0]} rows")
        # trainable encoders are re-packed on EVERY call (omics_path.refresh_encoders): a few multi-tensor launches instead
        # of ~800 single copies (bf16 parameters under DeepSpeed bf16 go straight into the bf16 matrices)
        if same_d:
            torch._foreach_copy_(same_d, same_s)
        if conv_d:
            torch._foreach_copy_(conv_d, conv_s)

    def load_projector(self, weight: torch.Tensor, bias: torch.Tensor) -> None:
        """(Re)pack ``nn.Linear`` projector parameters into the buffers the kernels read."""
        self.proj_w.copy_(weight.detach())
        self.proj_b.copy_(bias.detach())

    @property
    def handle(self) -> C.c_void_p:
        if not self._handle:
            raise RuntimeError("encoder handle already destroyed")
        return self._handle

    def workspace_bytes(self, n_seq: int, k_tokens: int) -> int:
        return int(self._lib.molly_encoder_workspace_bytes(self.handle, n_seq, k_tokens))

    def close(self) -> None:
        for attr in ("_train_graphs", "_train_shapes_seen", "_grad_plan"):     # graphs of the training step hold gigabytes
            self.__dict__.pop(attr, None)
        if getattr(self, "_handle", None):
            self._lib.molly_encoder_destroy(self._handle)
            self._handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
