"""Encoder shape/config record of the path -- the fields of ``transformers.EsmConfig`` that the reference's encoder
forward actually reads (HF ``modeling_esm.py``:161-186, 285-316, 406-427) plus the FFN flavour of NT-v2."""
from __future__ import annotations

from dataclasses import dataclass, fields
from typing import Any, Mapping, Optional


@dataclass(frozen=True)
class EncoderConfig:
    hidden_size: int
    num_hidden_layers: int
    num_attention_heads: int
    intermediate_size: int
    vocab_size: int
    pad_token_id: int = 1
    mask_token_id: int = 32
    position_embedding_type: str = "rotary"     # "rotary" | "absolute"
    max_position_embeddings: int = 1026
    ffn_type: str = "gelu"                      # "gelu" = Linear+bias, erf-GELU ; "glu" = gated SiLU (NT-v2)
    # NT-v2's architecture is hub remote code that cannot be read offline (parity unpinned).  The two places where
    # a plausible variant of it would silently change results are explicit switches instead of assumptions:
    glu_gate_first: bool = True                 # silu(x1) * x2 with x1 = rows [0, F) of intermediate.dense (False: silu(x2) * x1)
    ffn_bias: Optional[bool] = None             # FFN Linear biases; None = whatever the state dict holds (`add_bias_fnn`)
    token_dropout: bool = True
    emb_layer_norm_before: bool = False
    layer_norm_eps: float = 1e-5
    name: str = ""

    @property
    def head_dim(self) -> int:
        return self.hidden_size // self.num_attention_heads

    @classmethod
    def from_mapping(cls, m: Mapping[str, Any]) -> "EncoderConfig":
        names = {f.name for f in fields(cls)}
        return cls(**{k: v for k, v in m.items() if k in names})

    @classmethod
    def from_hf_config(cls, hf_cfg: Any, state_dict: Optional[Mapping[str, Any]] = None) -> "EncoderConfig":
        """Read an ``EsmConfig`` (stock, or NT-v2's remote-code variant).  The gated FFN is recognised from the config
        (``add_bias_fnn == False`` in the NT-v2 remote code) or from the ``intermediate.dense`` weight being ``[2F, h]``."""
        ffn = "gelu"
        ffn_bias = None
        for attr in ("add_bias_fnn", "add_bias_ffn"):          # the remote code spells it "fnn"; accept the sane spelling too
            if hasattr(hf_cfg, attr):
                ffn_bias = bool(getattr(hf_cfg, attr))
                if ffn_bias is False:
                    ffn = "glu"
        if state_dict is not None:
            w = state_dict.get("esm.encoder.layer.0.intermediate.dense.weight")
            if w is not None and w.shape[0] == 2 * hf_cfg.intermediate_size:
                ffn = "glu"
            has_b = "esm.encoder.layer.0.intermediate.dense.bias" in state_dict
            if ffn_bias is not None and ffn_bias != has_b:
                raise ValueError(f"config says FFN bias = {ffn_bias} but the state dict "
                                 f"{'has' if has_b else 'lacks'} esm.encoder.layer.0.intermediate.dense.bias")
            ffn_bias = has_b
        pos = getattr(hf_cfg, "position_embedding_type", "absolute")
        if pos not in ("rotary", "absolute"):
            raise ValueError(f"Unsupported position_embedding_type: {pos}")
        return cls(
            hidden_size=hf_cfg.hidden_size, num_hidden_layers=hf_cfg.num_hidden_layers,
            num_attention_heads=hf_cfg.num_attention_heads, intermediate_size=hf_cfg.intermediate_size,
            vocab_size=hf_cfg.vocab_size, pad_token_id=hf_cfg.pad_token_id,
            mask_token_id=hf_cfg.mask_token_id if hf_cfg.mask_token_id is not None else -1,
            position_embedding_type=pos, max_position_embeddings=hf_cfg.max_position_embeddings, ffn_type=ffn,
            glu_gate_first=bool(getattr(hf_cfg, "glu_gate_first", True)), ffn_bias=ffn_bias,
            token_dropout=bool(getattr(hf_cfg, "token_dropout", False)),
            emb_layer_norm_before=bool(getattr(hf_cfg, "emb_layer_norm_before", False) or False),
            layer_norm_eps=float(hf_cfg.layer_norm_eps), name=getattr(hf_cfg, "name_or_path", "") or "")
