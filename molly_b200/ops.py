"""PyTorch custom ops (``torch.ops.molly_b200.*``) over the C ABI of ``libmolly_b200.so``.

torch is plumbing here: it owns device memory and the current stream; every op body is one ctypes call with raw
pointers.  There is no CPU implementation registered for any op -- calling one with CPU tensors raises.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Tuple

import torch
from torch import Tensor

from . import _lib
from .packing import PackedEncoder

_ENCODERS: Dict[int, PackedEncoder] = {}
_NEXT_ID = [1]
_WORKSPACES: Dict[Tuple[int, int], Tensor] = {}
_ERR_FLAGS: Dict[int, Tensor] = {}


# ------------------------------------------------------------------------------------------------
# registries / helpers
# ------------------------------------------------------------------------------------------------
def register_encoder(enc: PackedEncoder) -> int:
    eid = _NEXT_ID[0]
    _NEXT_ID[0] += 1
    _ENCODERS[eid] = enc
    return eid


def unregister_encoder(eid: int) -> None:
    enc = _ENCODERS.pop(eid, None)
    if enc is not None:
        enc.close()


def get_encoder(eid: int) -> PackedEncoder:
    return _ENCODERS[eid]


def _require_cuda(*tensors: Tensor) -> torch.device:
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("molly_b200 ops take CUDA tensors only (there is no CPU fallback)")
        if dev is not None and t.device != dev:
            raise RuntimeError(f"tensors on different devices: {t.device} vs {dev}")
        dev = t.device
    return dev


def _stream(dev: torch.device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _dtype_code(t: Tensor) -> int:
    if t.dtype == torch.bfloat16:
        return _lib.DTYPE_BF16
    if t.dtype == torch.float32:
        return _lib.DTYPE_F32
    raise ValueError(f"unsupported dtype {t.dtype}: the path writes bf16 or fp32 hidden_states")


_WS_PINNED: Dict[int, Tensor] = {}      # device index -> caller-owned workspace (CUDA-graph capture)


class use_workspace:
    """``with use_workspace(ws):`` every op on ``ws.device`` scratches in ``ws`` (a uint8 tensor that outlives the block's
    kernels).  CUDA-graph capture needs it: the workspace must be allocated outside the capture and keep its address."""

    def __init__(self, ws: Tensor):
        self.ws, self.idx = ws, ws.device.index if ws.device.index is not None else torch.cuda.current_device()

    def __enter__(self):
        self.prev = _WS_PINNED.get(self.idx)
        _WS_PINNED[self.idx] = self.ws
        return self.ws

    def __exit__(self, *exc):
        if self.prev is None:
            _WS_PINNED.pop(self.idx, None)
        else:
            _WS_PINNED[self.idx] = self.prev
        return False


def workspace_is_pinned(dev: torch.device) -> bool:
    return (dev.index if dev.index is not None else torch.cuda.current_device()) in _WS_PINNED


def workspace(dev: torch.device, nbytes: int) -> Tensor:
    """Grow-only scratch buffer per (device, stream); stable pointers keep the library's TMA-descriptor plan cached."""
    pinned = _WS_PINNED.get(dev.index if dev.index is not None else torch.cuda.current_device())
    if pinned is not None:
        if pinned.numel() < nbytes + 1024:
            raise RuntimeError(f"pinned workspace too small: {pinned.numel()} < {nbytes + 1024} bytes")
        return pinned
    key = (dev.index if dev.index is not None else torch.cuda.current_device(),
           torch.cuda.current_stream(dev).cuda_stream)
    ws = _WORKSPACES.get(key)
    if ws is None or ws.numel() < nbytes:
        if ws is not None:
            del _WORKSPACES[key]
            del ws
        ws = torch.empty(int(nbytes * 1.05) + 4096, dtype=torch.uint8, device=dev)
        _WORKSPACES[key] = ws
    return ws


def _aligned(ws: Tensor, align: int = 1024) -> Tuple[int, int]:
    p = ws.data_ptr()
    off = (-p) % align
    return p + off, ws.numel() - off


def error_flag(dev: torch.device) -> Tensor:
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    f = _ERR_FLAGS.get(idx)
    if f is None:
        f = torch.zeros(1, dtype=torch.int32, device=dev)
        _ERR_FLAGS[idx] = f
    return f


def check_device_errors(dev: torch.device, what: str = "process_omic_sequences") -> None:
    """Synchronising read of the device-side error flag; raises what the reference raises for the same input."""
    f = error_flag(dev)
    bits = int(f.item())
    if bits == 0:
        return
    f.zero_()
    if bits & _lib.ERRBIT_OOV:
        raise AssertionError(f"{what}: out-of-range token id (>= vocab_size) in omic_ids")          # omics_one.py:71-72
    if bits & _lib.ERRBIT_OVERFLOW:
        raise RuntimeError(f"{what}: omics placeholder run exceeds the text sequence length")       # slice shape mismatch
    if bits & _lib.ERRBIT_POSITION:
        raise RuntimeError(f"Error processing omic sequences: position id exceeds max_position_embeddings")  # :89-90
    if bits & _lib.ERRBIT_TOKEN:
        raise IndexError(f"{what}: index out of range in self (input_ids entry outside the LLM embedding table)")
    if bits & _lib.ERRBIT_LAYOUT:
        raise RuntimeError(f"{what}: placeholder runs in input_ids do not pair with the omic_ids slots")
    raise RuntimeError(f"{what}: device error flag {bits}")


def kernel_launch_count() -> int:
    return int(_lib.load().molly_kernel_launch_count())


_profiling = False            # per-launch events cannot be recorded inside a replayed graph: graphed paths check this


def profiling_active() -> bool:
    return _profiling


def add_kernel_launches(n: int) -> None:
    """Report the replay of a CUDA graph that holds ``n`` of this library's launches (keeps ``kernel_launch_count`` honest)."""
    _lib.check(_lib.load().molly_add_kernel_launches(int(n)), "molly_add_kernel_launches")


def profile_start() -> None:
    """Bracket every kernel the library launches from now on with CUDA events on the launching stream."""
    global _profiling
    _lib.check(_lib.load().molly_profile_start(), "molly_profile_start")
    _profiling = True


def profile_stop() -> Dict[str, dict]:
    """Per kernel family: launches, total ms, algorithmic work (FLOP or HBM bytes) since profile_start()."""
    global _profiling
    _profiling = False
    arr = (_lib.ProfileEntry * _lib.PROFILE_FAMILIES)()
    _lib.check(_lib.load().molly_profile_stop(arr, _lib.PROFILE_FAMILIES), "molly_profile_stop")
    out = {}
    for e in arr:
        if e.launches:
            out[e.name.decode()] = {"launches": int(e.launches), "ms": float(e.total_ms), "work": float(e.work),
                                    "unit": "flop" if e.work_is_flops else "byte"}
    return out


# ------------------------------------------------------------------------------------------------
# hot-path ops
# ------------------------------------------------------------------------------------------------
@torch.library.custom_op("molly_b200::encode_project_merge", mutates_args=("hidden_states",), device_types="cuda")
def encode_project_merge(hidden_states: Tensor, ids: Tensor, seq_table: Tensor, enc_id: int,
                         save_enc_out: bool) -> Tensor:
    """One modality of ``_inject_omic`` (omics_one.py:57-97): encoder forward -> projector -> in-place scatter."""
    enc = _ENCODERS[enc_id]
    dev = _require_cuda(hidden_states, ids, seq_table)
    if hidden_states.dim() != 3 or not hidden_states.is_contiguous():
        raise ValueError("hidden_states must be a contiguous [B, T, D] tensor")
    if ids.dtype != torch.int64 or ids.dim() != 2 or not ids.is_contiguous():
        raise ValueError("ids must be a contiguous int64 [n_seq, K] tensor")
    if seq_table.dtype != torch.int32 or tuple(seq_table.shape) != (ids.shape[0], 2) or not seq_table.is_contiguous():
        raise ValueError("seq_table must be a contiguous int32 [n_seq, 2] tensor")
    B, T, D = hidden_states.shape
    n_seq, K = ids.shape
    lib = _lib.load()
    ws = workspace(dev, enc.workspace_bytes(n_seq, K))
    ws_ptr, ws_bytes = _aligned(ws)
    enc_out = (torch.empty(n_seq * K, enc.cfg.hidden_size, dtype=torch.bfloat16, device=dev) if save_enc_out
               else torch.empty(0, dtype=torch.bfloat16, device=dev))
    with torch.cuda.device(dev):
        _lib.check(lib.molly_encode_project_merge_fwd(
            enc.handle, ids.data_ptr(), seq_table.data_ptr(), n_seq, K, hidden_states.data_ptr(),
            _dtype_code(hidden_states), B, T, D, ws_ptr, ws_bytes, error_flag(dev).data_ptr(),
            enc_out.data_ptr() if save_enc_out else None, _stream(dev)), "molly_encode_project_merge_fwd")
    return enc_out


@torch.library.custom_op("molly_b200::encode", mutates_args=(), device_types="cuda")
def encode(ids: Tensor, enc_id: int) -> Tensor:
    """``EsmForMaskedLM(ids, attention_mask=ids != 1, output_hidden_states=True).hidden_states[-1]`` -> bf16 [n, K, h]."""
    enc = _ENCODERS[enc_id]
    dev = _require_cuda(ids)
    if ids.dtype != torch.int64 or ids.dim() != 2 or not ids.is_contiguous():
        raise ValueError("ids must be a contiguous int64 [n_seq, K] tensor")
    n_seq, K = ids.shape
    lib = _lib.load()
    ws = workspace(dev, enc.workspace_bytes(n_seq, K))
    ws_ptr, ws_bytes = _aligned(ws)
    out = torch.empty(n_seq, K, enc.cfg.hidden_size, dtype=torch.bfloat16, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.molly_encode_fwd(enc.handle, ids.data_ptr(), n_seq, K, out.data_ptr(), ws_ptr, ws_bytes,
                                        error_flag(dev).data_ptr(), _stream(dev)), "molly_encode_fwd")
    return out


@torch.library.custom_op("molly_b200::pool", mutates_args=(), device_types="cuda")
def pool(enc_out: Tensor, ids: Tensor, mode: int) -> Tensor:
    """mode 0: masked mean over non-pad tokens (embed_text.py:112-129); mode 1: CLS token (baselines/model.py:104-120)."""
    dev = _require_cuda(enc_out, ids)
    n_seq, K, h = enc_out.shape
    if enc_out.dtype != torch.bfloat16 or not enc_out.is_contiguous() or not ids.is_contiguous():
        raise ValueError("pool: enc_out must be contiguous bf16 [n, K, h]")
    out = torch.empty(n_seq, h, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().molly_pool_fwd(enc_out.data_ptr(), ids.data_ptr(), n_seq, K, h, mode, out.data_ptr(),
                                              _stream(dev)), "molly_pool_fwd")
    return out


@torch.library.custom_op("molly_b200::project_bwd", mutates_args=("d_hidden",), device_types="cuda")
def project_bwd(d_hidden: Tensor, seq_table: Tensor, enc_out: Tensor, enc_id: int, k_tokens: int,
                zero_rows: bool) -> Tuple[Tensor, Tensor]:
    """Projector grads through the slice-assign: returns (dW [D,h] fp32, db [D] fp32); optionally zeroes the written rows
    of ``d_hidden`` in place (they carry no gradient to the LLM embedding table)."""
    enc = _ENCODERS[enc_id]
    dev = _require_cuda(d_hidden, seq_table, enc_out)
    if not d_hidden.is_contiguous() or d_hidden.dim() != 3:
        raise ValueError("d_hidden must be contiguous [B, T, D]")
    B, T, D = d_hidden.shape
    n_seq = seq_table.shape[0]
    h = enc.cfg.hidden_size
    M = n_seq * k_tokens
    Mp = (M + 7) // 8 * 8
    need = M * D * 2 + 1024 + (D + h) * Mp * 2 + 4096
    ws = workspace(dev, need)
    ws_ptr, ws_bytes = _aligned(ws)
    dW = torch.empty(D, h, dtype=torch.float32, device=dev)
    db = torch.empty(D, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().molly_project_bwd(
            enc.handle, d_hidden.data_ptr(), _dtype_code(d_hidden), seq_table.data_ptr(), n_seq, k_tokens, B, T, D,
            enc_out.data_ptr(), dW.data_ptr(), db.data_ptr(), int(zero_rows), ws_ptr, ws_bytes, _stream(dev)),
            "molly_project_bwd")
    return dW, db


@torch.library.custom_op("molly_b200::placeholder_scan", mutates_args=(), device_types="cuda")
def placeholder_scan(input_ids: Tensor, pad0: int, pad1: int, pad2: int) -> Tuple[Tensor, Tensor, Tensor]:
    """Positions / kinds / per-sample counts of the three omics pad-token ids in ``input_ids`` [B, T]."""
    dev = _require_cuda(input_ids)
    if input_ids.dtype != torch.int64 or input_ids.dim() != 2 or not input_ids.is_contiguous():
        raise ValueError("input_ids must be contiguous int64 [B, T]")
    B, T = input_ids.shape
    pos = torch.empty(B, T, dtype=torch.int32, device=dev)
    kind = torch.empty(B, T, dtype=torch.int32, device=dev)
    cnt = torch.empty(B, dtype=torch.int32, device=dev)
    pads = (C.c_int64 * 3)(pad0, pad1, pad2)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().molly_placeholder_scan(input_ids.data_ptr(), B, T, pads, pos.data_ptr(), kind.data_ptr(),
                                                      cnt.data_ptr(), _stream(dev)), "molly_placeholder_scan")
    return pos, kind, cnt


def placeholder_runs(input_ids: Tensor, pad_ids: Tuple[int, int, int], n_slots: Optional[Tensor], max_runs: int):
    """Runs of *_pad tokens per sample in text order -> (run_start, run_kind, run_len) [B, max_runs], n_runs [B] and
    pos_j [B, T] (index inside the run, -1 for text / unpaired runs).  SURVEY.md 8f row N1."""
    dev = _require_cuda(input_ids)
    if input_ids.dtype != torch.int64 or input_ids.dim() != 2 or not input_ids.is_contiguous():
        raise ValueError("input_ids must be contiguous int64 [B, T]")
    B, T = input_ids.shape
    run_start = torch.full((B, max_runs), -1, dtype=torch.int32, device=dev)
    run_kind = torch.full((B, max_runs), -1, dtype=torch.int32, device=dev)
    run_len = torch.empty(B, max_runs, dtype=torch.int32, device=dev)
    n_runs = torch.empty(B, dtype=torch.int32, device=dev)
    pos_j = torch.empty(B, T, dtype=torch.int32, device=dev)
    pads = (C.c_int64 * 3)(*[int(p) for p in pad_ids])
    with torch.cuda.device(dev):
        _lib.check(_lib.load().molly_placeholder_runs(
            input_ids.data_ptr(), B, T, pads, n_slots.data_ptr() if n_slots is not None else None, max_runs,
            run_start.data_ptr(), run_kind.data_ptr(), run_len.data_ptr(), n_runs.data_ptr(), pos_j.data_ptr(),
            _stream(dev)), "molly_placeholder_runs")
    return run_start, run_kind, run_len, n_runs, pos_j


def placeholder_reject(runs, slot_expect: Tensor, cap_dna_rna: int, cap_protein: int) -> None:
    """In place on ``runs[4]`` (pos_j): runs that ``build_seq_table`` will reject go back to -1 (embedded like text)."""
    run_start, run_kind, run_len, n_runs, pos_j = runs
    dev = _require_cuda(pos_j, slot_expect)
    B, T = pos_j.shape
    if slot_expect.dtype != torch.int32 or tuple(slot_expect.shape) != tuple(run_start.shape) or not slot_expect.is_contiguous():
        raise ValueError("slot_expect must be contiguous int32 [B, max_runs]")
    with torch.cuda.device(dev):
        _lib.check(_lib.load().molly_placeholder_reject(
            pos_j.data_ptr(), run_start.data_ptr(), run_kind.data_ptr(), run_len.data_ptr(), n_runs.data_ptr(),
            slot_expect.data_ptr(), B, T, run_start.shape[1], int(cap_dna_rna), int(cap_protein), _stream(dev)),
            "molly_placeholder_reject")


def build_seq_table(b_idx: Tensor, run_idx: Tensor, runs, expect_protein: bool, k_need: int = 1) -> Tensor:
    """seq_table [n, 2] = (b, x_start position) from the device-side runs: the reference's info["start"]."""
    dev = _require_cuda(b_idx, run_idx)
    run_start, run_kind, run_len, n_runs, _ = runs
    n = b_idx.numel()
    table = torch.empty(n, 2, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().molly_build_seq_table(
            b_idx.data_ptr(), run_idx.data_ptr(), n, run_start.data_ptr(), run_kind.data_ptr(), run_len.data_ptr(),
            n_runs.data_ptr(), run_start.shape[1], int(expect_protein), k_need, table.data_ptr(),
            error_flag(dev).data_ptr(), _stream(dev)), "molly_build_seq_table")
    return table


def embed_tokens_skip(input_ids: Tensor, pos_j: Tensor, pad_ids: Tuple[int, int, int], cap_dna_rna: int, cap_protein: int,
                      table: Tensor, out: Optional[Tensor] = None) -> Tensor:
    """``embed_tokens(input_ids)`` (omics_one.py:164, :209) except the rows the omics path is about to overwrite."""
    dev = _require_cuda(input_ids, pos_j, table)
    B, T = input_ids.shape
    vocab, D = table.shape
    if not table.is_contiguous():
        raise ValueError("embedding table must be contiguous")
    if out is None:
        out = torch.empty(B, T, D, dtype=table.dtype, device=dev)
    pads = (C.c_int64 * 3)(*[int(p) for p in pad_ids])
    with torch.cuda.device(dev):
        _lib.check(_lib.load().molly_embed_tokens_skip(
            input_ids.data_ptr(), pos_j.data_ptr(), pads, cap_dna_rna, cap_protein, table.data_ptr(), _dtype_code(table),
            vocab, D, out.data_ptr(), B, T, error_flag(dev).data_ptr(), _stream(dev)), "molly_embed_tokens_skip")
    return out


# ------------------------------------------------------------------------------------------------
# single-kernel ops (unit-parity surface)
# ------------------------------------------------------------------------------------------------
def gemm_bf16(a: Tensor, w: Tensor, epilogue: int, bias: Optional[Tensor] = None, residual: Optional[Tensor] = None,
              out: Optional[Tensor] = None, out_dtype: torch.dtype = torch.bfloat16,
              seq_table: Optional[Tensor] = None, seq_k: int = 0, k_cap: int = 0, scale_cols: int = 0,
              scale: float = 1.0, rope_cos_t: Optional[Tensor] = None, rope_sin_t: Optional[Tensor] = None,
              rope_cols: int = 0, rope_head_dim: int = 0) -> Tensor:
    dev = _require_cuda(a, w, bias, residual, out, seq_table)
    M, K = a.shape
    N = w.shape[0]
    n_out = N // 2 if epilogue == _lib.EPI_GLU else N
    B = T = 0
    if epilogue == _lib.EPI_SCATTER:
        B, T, _ = out.shape
        ldo = out.shape[-1]
    else:
        if out is None:
            out = torch.empty(M, n_out, dtype=out_dtype, device=dev)
        ldo = out.stride(0)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().molly_gemm_bf16(
            a.data_ptr(), a.stride(0), w.data_ptr(), w.stride(0), M, N, K, epilogue,
            None if bias is None else bias.data_ptr(), None if residual is None else residual.data_ptr(),
            out.data_ptr(), _dtype_code(out), ldo, None if seq_table is None else seq_table.data_ptr(), seq_k, B, T,
            k_cap, error_flag(dev).data_ptr(), scale_cols, scale,
            None if rope_cos_t is None else rope_cos_t.data_ptr(), None if rope_sin_t is None else rope_sin_t.data_ptr(),
            0 if rope_cos_t is None else rope_cos_t.shape[1], rope_cols, rope_head_dim, _stream(dev)),
            "molly_gemm_bf16")
    return out


def layernorm(x: Tensor, w: Tensor, b: Tensor, eps: float, out_dtype: torch.dtype = torch.bfloat16) -> Tensor:
    dev = _require_cuda(x, w, b)
    rows, h = x.shape
    out = torch.empty(rows, h, dtype=out_dtype, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().molly_layernorm(x.data_ptr(), w.data_ptr(), b.data_ptr(), rows, h, eps, out.data_ptr(),
                                               _dtype_code(out), _stream(dev)), "molly_layernorm")
    return out


def embed(ids: Tensor, c_config: "_lib.EncoderConfig", word_emb: Tensor, pos_emb: Optional[Tensor]):
    dev = _require_cuda(ids, word_emb, pos_emb)
    n_seq, K = ids.shape
    h = c_config.hidden_size
    x = torch.empty(n_seq * K, h, dtype=torch.float32, device=dev)
    kv_info = torch.empty(n_seq, 2, dtype=torch.int32, device=dev)
    key_mask = torch.empty(n_seq * K, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().molly_embed(ids.data_ptr(), n_seq, K, C.byref(c_config), word_emb.data_ptr(),
                                           None if pos_emb is None else pos_emb.data_ptr(), x.data_ptr(),
                                           kv_info.data_ptr(), key_mask.data_ptr(), error_flag(dev).data_ptr(),
                                           _stream(dev)), "molly_embed")
    return x, kv_info, key_mask


def rotary_(qkv: Tensor, k_tokens: int, heads: int, cos: Tensor, sin: Tensor) -> Tensor:
    dev = _require_cuda(qkv, cos, sin)
    rows, h3 = qkv.shape
    with torch.cuda.device(dev):
        _lib.check(_lib.load().molly_rotary(qkv.data_ptr(), rows, k_tokens, h3 // 3, heads, cos.data_ptr(),
                                            sin.data_ptr(), _stream(dev)), "molly_rotary")
    return qkv


def attention(qkv: Tensor, n_seq: int, k_tokens: int, heads: int, kv_info: Tensor, key_mask: Tensor) -> Tensor:
    dev = _require_cuda(qkv, kv_info, key_mask)
    h = qkv.shape[1] // 3
    out = torch.empty(n_seq * k_tokens, h, dtype=torch.bfloat16, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().molly_attention(qkv.data_ptr(), n_seq, k_tokens, h, heads, kv_info.data_ptr(),
                                               key_mask.data_ptr(), out.data_ptr(), _stream(dev)), "molly_attention")
    return out


def attention_lse(qkv: Tensor, n_seq: int, k_tokens: int, heads: int, kv_info: Tensor, key_mask: Tensor):
    """``attention`` that also returns the row log-sum-exp in the log2 domain, fp32 [n_seq, heads, k_tokens]."""
    dev = _require_cuda(qkv, kv_info, key_mask)
    h = qkv.shape[1] // 3
    out = torch.empty(qkv.shape[0], h, dtype=torch.bfloat16, device=dev)
    lse2 = torch.empty(n_seq, heads, k_tokens, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().molly_attention_lse(qkv.data_ptr(), n_seq, k_tokens, h, heads, kv_info.data_ptr(),
                                                   key_mask.data_ptr(), out.data_ptr(), lse2.data_ptr(), _stream(dev)),
                   "molly_attention_lse")
    return out, lse2


def attention_bwd(qkv: Tensor, out: Tensor, d_out: Tensor, lse2: Tensor, n_seq: int, k_tokens: int, heads: int,
                  kv_info: Tensor, key_mask: Tensor) -> Tensor:
    """d(q', k', v) packed like ``qkv`` (bf16 [n_seq*k, 3h]) from ``d_out`` (bf16 [n_seq*k, h])."""
    dev = _require_cuda(qkv, out, d_out, lse2, kv_info, key_mask)
    h = qkv.shape[1] // 3
    for t, shape in ((out, (qkv.shape[0], h)), (d_out, (qkv.shape[0], h))):
        if t.dtype != torch.bfloat16 or tuple(t.shape) != shape or not t.is_contiguous():
            raise ValueError("out / d_out must be contiguous bf16 [n_seq*k, h]")
    d_qkv = torch.empty_like(qkv)
    delta = torch.empty(n_seq * heads * k_tokens, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().molly_attention_bwd(qkv.data_ptr(), out.data_ptr(), d_out.data_ptr(), lse2.data_ptr(), n_seq,
                                                   k_tokens, h, heads, kv_info.data_ptr(), key_mask.data_ptr(),
                                                   d_qkv.data_ptr(), delta.data_ptr(), _stream(dev)), "molly_attention_bwd")
    return d_qkv


# ---- encoder-backward building blocks ------------------------------------------------------------
def linear_wgrad(dy: Tensor, x: Tensor, with_bias: bool = True):
    """Autograd of ``nn.Linear`` w.r.t. its parameters: (dW [N, K] = dy^T x, db [N] = colsum(dy) or None), fp32."""
    dev = _require_cuda(dy, x)
    M, N = dy.shape
    K = x.shape[1]
    if dy.dtype != torch.bfloat16 or x.dtype != torch.bfloat16 or x.shape[0] != M or not (dy.is_contiguous() and x.is_contiguous()):
        raise ValueError("linear_wgrad: dy [M, N] and x [M, K] must be contiguous bf16")
    dW = torch.empty(N, K, dtype=torch.float32, device=dev)
    db = torch.empty(N, dtype=torch.float32, device=dev) if with_bias else None
    ws = workspace(dev, (N + K) * ((M + 7) // 8 * 8) * 2 + 4096)
    ws_ptr, ws_bytes = _aligned(ws)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().molly_linear_wgrad(dy.data_ptr(), x.data_ptr(), M, N, K, dW.data_ptr(),
                                                  None if db is None else db.data_ptr(), ws_ptr, ws_bytes, _stream(dev)),
                   "molly_linear_wgrad")
    return dW, db


def gather_rows(d_hidden: Tensor, seq_table: Tensor, k_tokens: int, k_cap: int, zero_rows: bool) -> Tensor:
    """bf16 [n_seq*k, D]: the rows of ``d_hidden`` [B, T, D] the forward's slice-assign wrote (0 beyond k_cap)."""
    dev = _require_cuda(d_hidden, seq_table)
    B, T, D = d_hidden.shape
    n_seq = seq_table.shape[0]
    dy = torch.empty(n_seq * k_tokens, D, dtype=torch.bfloat16, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().molly_gather_rows(d_hidden.data_ptr(), _dtype_code(d_hidden), seq_table.data_ptr(), n_seq,
                                                 k_tokens, k_cap, B, T, D, dy.data_ptr(), int(zero_rows), _stream(dev)),
                   "molly_gather_rows")
    return dy


def transpose_bf16(x: Tensor) -> Tensor:
    dev = _require_cuda(x)
    rows, cols = x.shape
    out = torch.empty(cols, rows, dtype=torch.bfloat16, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().molly_transpose_bf16(x.data_ptr(), rows, cols, out.data_ptr(), _stream(dev)),
                   "molly_transpose_bf16")
    return out


def layernorm_bwd(x: Tensor, dy: Tensor, gamma: Tensor, eps: float, d_x: Tensor, accumulate: bool,
                  d_gamma: Optional[Tensor] = None, d_beta: Optional[Tensor] = None) -> Tensor:
    """``d_x`` (fp32 [rows, h]) = or += LayerNorm-backward(dy); ``d_gamma`` / ``d_beta`` (fp32, pre-zeroed) accumulate."""
    dev = _require_cuda(x, dy, gamma, d_x, d_gamma, d_beta)
    rows, h = x.shape
    stats = torch.empty(rows, 2, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().molly_layernorm_bwd(x.data_ptr(), dy.data_ptr(), gamma.data_ptr(), rows, h, eps, d_x.data_ptr(),
                                                   int(accumulate), stats.data_ptr(),
                                                   None if d_gamma is None else d_gamma.data_ptr(),
                                                   None if d_beta is None else d_beta.data_ptr(), _stream(dev)),
                   "molly_layernorm_bwd")
    return d_x


def act_fwd_bwd(glu: bool, pre: Tensor, d_act: Optional[Tensor]):
    """(act, d_pre) of the FFN activation: erf-GELU on ``pre`` [rows, F] or gated SiLU on interleaved ``pre`` [rows, 2F];
    ``d_act`` None: forward only, returns (act, None)."""
    dev = _require_cuda(pre, d_act)
    rows = pre.shape[0]
    f_out = pre.shape[1] // 2 if glu else pre.shape[1]
    act = torch.empty(rows, f_out, dtype=torch.bfloat16, device=dev)
    d_pre = torch.empty_like(pre) if d_act is not None else None
    with torch.cuda.device(dev):
        _lib.check(_lib.load().molly_act_fwd_bwd(int(glu), pre.data_ptr(), None if d_act is None else d_act.data_ptr(), rows,
                                                 f_out, act.data_ptr(), None if d_pre is None else d_pre.data_ptr(),
                                                 _stream(dev)), "molly_act_fwd_bwd")
    return act, d_pre


def cast_bf16(x: Tensor) -> Tensor:
    dev = _require_cuda(x)
    out = torch.empty(x.shape, dtype=torch.bfloat16, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().molly_cast_f32_bf16(x.data_ptr(), x.numel(), out.data_ptr(), _stream(dev)),
                   "molly_cast_f32_bf16")
    return out


def scale_cols_(x: Tensor, cols: int, scale: float) -> Tensor:
    dev = _require_cuda(x)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().molly_scale_cols(x.data_ptr(), x.shape[0], x.stride(0), cols, scale, _stream(dev)),
                   "molly_scale_cols")
    return x


def scatter_add_rows_(table: Tensor, src: Tensor, index: Tensor, scale: Tensor) -> Tensor:
    dev = _require_cuda(table, src, index, scale)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().molly_scatter_add_rows(src.data_ptr(), index.data_ptr(), scale.data_ptr(), src.shape[0],
                                                      src.shape[1], table.data_ptr(), _stream(dev)),
                   "molly_scatter_add_rows")
    return table


def merge_rows_(hidden_states: Tensor, src: Tensor, seq_table: Tensor, k_tokens: int, k_cap: int) -> Tensor:
    dev = _require_cuda(hidden_states, src, seq_table)
    B, T, D = hidden_states.shape
    with torch.cuda.device(dev):
        _lib.check(_lib.load().molly_merge_rows(src.data_ptr(), seq_table.data_ptr(), seq_table.shape[0], k_tokens,
                                                k_cap, hidden_states.data_ptr(), _dtype_code(hidden_states), B, T, D,
                                                error_flag(dev).data_ptr(), _stream(dev)), "molly_merge_rows")
    return hidden_states
