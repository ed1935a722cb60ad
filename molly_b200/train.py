"""Encoder training path (``--train-bio``: src/utils/tools.py:313-331 unfreezes the encoders).

Forward = the same sm_100a kernels as the inference path, run layer by layer by ``molly_encode_train_fwd`` with what the
backward needs kept in a tape (every layer's activations, or -- when they do not fit -- only each layer's fp32 input, the
backward then recomputing one layer at a time); backward = ``molly_encode_train_bwd``, the autograd of every HF module the
kernels replace:

    nn.Linear      dgrad: tcgen05 GEMM on the transposed weight        wgrad / bias: ``molly_linear_wgrad``
    attention      ``molly_attention_bwd`` (dQ / dK,dV kernels) from the forward's row log-sum-exp
    LayerNorm      ``molly_layernorm_bwd``          GELU / gated SiLU   ``molly_act_fwd_bwd``
    rotary         the forward kernel with -sin (inverse rotation), then the q scale
    embeddings     ``molly_scatter_add_rows`` (token-dropout scale, pad / <mask> rows skipped like the forward)

Orchestration is native (csrc/train.cu: one C call for the forward, one per range of layers for the backward); the first
version issued one ctypes call per kernel and the step was bound by the host (133 ms to enqueue 115 ms of kernels).  The
library writes fp32 gradients into ONE flat buffer in the packed layout of the weights (``molly_encoder_grad_layout``);
``GradPlan`` maps it to the HF ``state_dict`` names (q / k / v are slices, the NT-v2 gate / up halves are un-interleaved).
Once a shape is steady state the two calls are replayed from CUDA graphs (``_GraphedStep``).
"""
from __future__ import annotations

import ctypes as C
import os
import weakref
from typing import Dict, List, Optional, Tuple

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


class GradPlan:
    """The library's flat fp32 gradient layout of one encoder (``molly_encoder_grad_layout``) and its HF names.

    ``groups`` is the reduction schedule of ``--train-bio`` at N > 1, in the order the backward produces it: the projector
    (not part of the flat buffer), the encoder layers L-1 .. 0, then final LayerNorm + embeddings; each group is a list of
    (HF name, shape).  A rank whose micro-batch lacks the modality all-reduces zeros of ``group_numel(g)`` elements per group
    (``omics_path._AbsentModalityFn``) and maps what comes back with the same ``layer_views`` / ``tail_views``."""

    def __init__(self, enc: PackedEncoder):
        cfg = enc.cfg
        self.cfg = cfg
        self.L, self.h, self.F = cfg.num_hidden_layers, cfg.hidden_size, cfg.intermediate_size
        self.glu = cfg.ffn_type == "glu"
        self.F1 = 2 * self.F if self.glu else self.F
        off = (C.c_int64 * _lib.GRAD_SLOTS)()
        tail = (C.c_int64 * _lib.GRAD_TAIL_SLOTS)()
        group, total = C.c_int64(), C.c_int64()
        _lib.check(_lib.load().molly_encoder_grad_layout(enc.handle, off, C.byref(group), tail, C.byref(total)),
                   "molly_encoder_grad_layout")
        self.off, self.tail, self.group, self.total = list(off), list(tail), int(group.value), int(total.value)
        self.word_shape = tuple(enc.word_emb.shape)
        self.pos_shape = tuple(enc.pos_emb.shape) if enc.pos_emb is not None else None
        self.proj_shapes = (tuple(enc.proj_w.shape), tuple(enc.proj_b.shape))
        self.tail_begin = self.group * self.L

    # -- (HF name, slot, offset inside the slot, shape) of one layer group, in flat order
    def _layer_entries(self, i: int):
        h, F, F1 = self.h, self.F, self.F1
        p = f"esm.encoder.layer.{i}."
        G = _lib
        ent = [(p + "LayerNorm.weight", G.GRAD_LN2_W, 0, (h,)), (p + "LayerNorm.bias", G.GRAD_LN2_B, 0, (h,)),
               (p + "output.dense.bias", G.GRAD_B_FFN2, 0, (h,)), (p + "intermediate.dense.bias", G.GRAD_B_FFN1, 0, (F1,)),
               (p + "attention.LayerNorm.weight", G.GRAD_LN1_W, 0, (h,)), (p + "attention.LayerNorm.bias", G.GRAD_LN1_B, 0, (h,)),
               (p + "attention.output.dense.bias", G.GRAD_B_O, 0, (h,))]
        ent += [(p + f"attention.self.{nm}.bias", G.GRAD_B_QKV, j * h, (h,)) for j, nm in enumerate(("query", "key", "value"))]
        ent += [(p + "output.dense.weight", G.GRAD_W_FFN2, 0, (h, F)), (p + "intermediate.dense.weight", G.GRAD_W_FFN1, 0, (F1, h)),
                (p + "attention.output.dense.weight", G.GRAD_W_O, 0, (h, h))]
        ent += [(p + f"attention.self.{nm}.weight", G.GRAD_W_QKV, j * h * h, (h, h))
                for j, nm in enumerate(("query", "key", "value"))]
        return [e for e in ent if self.off[e[1]] >= 0]

    def _tail_entries(self):
        G = _lib
        ent = [("esm.encoder.emb_layer_norm_after.weight", G.GRAD_TAIL_FINAL_LN_W, (self.h,)),
               ("esm.encoder.emb_layer_norm_after.bias", G.GRAD_TAIL_FINAL_LN_B, (self.h,)),
               ("esm.embeddings.word_embeddings.weight", G.GRAD_TAIL_WORD_EMB, self.word_shape)]
        if self.pos_shape is not None:
            ent.append(("esm.embeddings.position_embeddings.weight", G.GRAD_TAIL_POS_EMB, self.pos_shape))
        return ent

    @property
    def groups(self) -> List[List[Tuple[str, Tuple[int, ...]]]]:
        out = [[("projector.weight", self.proj_shapes[0]), ("projector.bias", self.proj_shapes[1])]]
        for i in range(self.L - 1, -1, -1):
            out.append([(n, shape) for n, _, _, shape in self._layer_entries(i)])
        out.append([(n, shape) for n, _, shape in self._tail_entries()])
        return out

    def group_numel(self, g: int) -> int:
        """Elements all-reduced for schedule group ``g`` (0 projector, 1 .. L layers L-1 .. 0, L+1 the tail)."""
        if g == 0:
            return self.proj_shapes[0][0] * self.proj_shapes[0][1] + self.proj_shapes[1][0]
        return self.group if g <= self.L else self.total - self.tail_begin

    def layer_views(self, i: int, flat: torch.Tensor) -> Dict[str, torch.Tensor]:
        """HF-named gradients of layer ``i`` from its flat group: VIEWS only (no kernel reads ``flat`` here, so this is safe
        while an all-reduce of ``flat`` is still in flight).  The NT-v2 ``intermediate.dense`` entries are still in the packed
        row order (gate_0, up_0, gate_1, up_1, ...): ``finalize_`` un-interleaves them once the values are final."""
        out = {}
        for name, slot, sub, shape in self._layer_entries(i):
            n = 1
            for v in shape:
                n *= v
            out[name] = flat[self.off[slot] + sub:self.off[slot] + sub + n].view(shape)
        return out

    def finalize_(self, grads: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        """Packed -> HF layout for the tensors whose layouts differ (copies): call after the last collective has finished."""
        if self.glu:
            for i in range(self.L):
                for suffix in ("intermediate.dense.weight", "intermediate.dense.bias"):
                    name = f"esm.encoder.layer.{i}.{suffix}"
                    if name in grads:
                        grads[name] = glu_deinterleave(grads[name], self.cfg.glu_gate_first)
        return grads

    def tail_views(self, flat: torch.Tensor) -> Dict[str, torch.Tensor]:
        out = {}
        for name, slot, shape in self._tail_entries():
            n = 1
            for v in shape:
                n *= v
            o = self.tail[slot] - self.tail_begin
            out[name] = flat[o:o + n].view(shape)
        return out


def grad_plan(enc: PackedEncoder) -> GradPlan:
    plan = getattr(enc, "_grad_plan", None)
    if plan is None:
        plan = enc._grad_plan = GradPlan(enc)
    return plan


def reduce_schedule(enc: PackedEncoder) -> List[List[Tuple[str, Tuple[int, ...]]]]:
    """The gradient groups handed to a ``LayerwiseGradReducer``, in order, as (HF name, shape) lists (``GradPlan.groups``)."""
    return grad_plan(enc).groups


class EncoderTape:
    """What the training forward keeps for the backward: one device buffer laid out by the library (fp32 layer inputs,
    key mask, and either every layer's activations -- LN outputs, q/k/v, attention output, log-sum-exp, the FFN pre-activation
    and its activation -- or one activation slot the backward refills: ``recompute``)."""

    def __init__(self):
        self.buf: Optional[torch.Tensor] = None
        self.recompute = False
        self.ids = None
        self.graph = None                      # the _GraphedStep whose buffers this tape borrows (None: eager)
        self.gen = 0                           # ... and the generation of that loan


def _save_activations(enc: PackedEncoder, n_tokens: int) -> bool:
    """Keep every layer's activations (no recompute) when they take less than 35 % of the free device memory;
    MOLLY_TRAIN_RECOMPUTE=1 / 0 forces per-layer recompute / saving."""
    env = os.environ.get("MOLLY_TRAIN_RECOMPUTE")
    if env is not None:
        return env == "0"
    cache = enc.__dict__.setdefault("_save_activations_cache", {})
    if n_tokens not in cache:          # decided once per batch size: cudaMemGetInfo costs ~0.5 ms of device idle time per call
        cfg = enc.cfg
        f_pre = cfg.intermediate_size * (2 if cfg.ffn_type == "glu" else 1)
        per_layer = n_tokens * (cfg.hidden_size * (4 + 2 + 6 + 2 + 4 + 2) + f_pre * 2 + cfg.intermediate_size * 2)
        cache[n_tokens] = per_layer * cfg.num_hidden_layers < 0.35 * torch.cuda.mem_get_info(enc.device)[0]
    return cache[n_tokens]


def _sizes(enc: PackedEncoder, n_seq: int, K: int, recompute: bool) -> Tuple[int, int, int]:
    tape_b, ws_b, gf = C.c_size_t(), C.c_size_t(), C.c_int64()
    _lib.check(_lib.load().molly_encoder_train_sizes(enc.handle, n_seq, K, int(recompute), C.byref(tape_b), C.byref(ws_b),
                                                      C.byref(gf)), "molly_encoder_train_sizes")
    return int(tape_b.value), int(ws_b.value), int(gf.value)


def _device_buffer(nbytes: int, dev) -> torch.Tensor:
    """uint8 buffer whose base is 1024-B aligned (the caching allocator hands out 512-B aligned blocks)."""
    raw = torch.empty(nbytes + 1024, dtype=torch.uint8, device=dev)
    shift = (-raw.data_ptr()) % 1024
    return raw[shift:shift + nbytes]


class _GraphedStep:
    """CUDA graphs of one encoder's training forward and backward for one (n_seq, K): the ~35 launches per layer are
    replayed as two graphs instead of being enqueued one by one (the ``--train-bio`` step had ~7 ms of inter-kernel gaps over
    ~1 960 launches).  The buffers the graphs touch are owned here: ids, the forward output, the tape, the backward
    workspace, d_out and the flat gradient buffer.  ``busy`` is set between a forward and its backward; a second forward in
    between (re-entrant use) takes the eager path."""

    def __init__(self, enc: PackedEncoder, n_seq: int, K: int, recompute: bool, dev):
        h = enc.cfg.hidden_size
        self.n_seq, self.K, self.recompute = n_seq, K, recompute
        self.tape_bytes, self.ws_bytes, gf = _sizes(enc, n_seq, K, recompute)
        self.ids = torch.ones(n_seq, K, dtype=torch.int64, device=dev)
        self.out = torch.empty(n_seq * K, h, dtype=torch.bfloat16, device=dev)
        self.d_out = torch.empty(n_seq * K, h, dtype=torch.bfloat16, device=dev)
        self.tape = _device_buffer(self.tape_bytes, dev)
        self.ws = _device_buffer(self.ws_bytes, dev)
        self.flat = torch.empty(gf, dtype=torch.float32, device=dev)
        self.fwd = self.bwd = None
        self.fwd_launches = self.bwd_launches = 0
        self._owner = None                     # weak reference to the tape of the forward whose backward is still to come
        self.gen = 0                           # bumped by every forward that takes the buffers: a tape knows if it was overtaken
        self._passed_over = 0

    def try_acquire(self, tape) -> bool:
        """Take the buffers for ``tape``'s forward.  Refused while an earlier forward still waits for its backward (its tape is
        alive) -- the caller then runs eagerly -- except that a forward which has been passed over ``_GRAPH_ABANDON`` times is
        taken for abandoned (a forward in training mode whose loss was never back-propagated) and loses the buffers: its
        backward, should it still come, raises."""
        if self._owner is not None and self._owner() is not None:
            self._passed_over += 1
            if self._passed_over <= _GRAPH_ABANDON:
                return False
        self._owner = weakref.ref(tape)
        self._passed_over = 0
        self.gen += 1
        tape.gen = self.gen
        return True

    @property
    def busy(self) -> bool:
        return self._owner is not None and self._owner() is not None

    def release(self, tape) -> None:
        if tape.gen == self.gen:
            self._owner = None


def _graphs_enabled() -> bool:
    """MOLLY_TRAIN_GRAPH=0 turns the graphed training step off; it is also off while per-launch profiling is on (events cannot
    be recorded inside a replay) and inside somebody else's capture."""
    return (os.environ.get("MOLLY_TRAIN_GRAPH", "1") != "0" and not ops.profiling_active()
            and not torch.cuda.is_current_stream_capturing())


_GRAPH_MIN_SIGHTINGS = 3       # a shape is captured on its third use ...
_GRAPH_SHAPES = 2              # ... at most two shapes are held per encoder (the tape is gigabytes) ...
_GRAPH_MAX_CAPTURES = 8        # ... and an encoder whose shapes keep changing stops capturing (a capture costs ~0.1 s)
_GRAPH_ABANDON = 3             # forwards that may pass over a pending one before its buffers are reclaimed


def _graphed_step(enc: PackedEncoder, n_seq: int, K: int, recompute: bool, dev) -> Optional[_GraphedStep]:
    """The graphed step for this shape once it is steady state: the first two uses run eagerly (they configure the kernels and
    tell one-off shapes -- a training set whose number of omics sequences per micro-batch varies -- from the steady state);
    a new shape replaces a held one only when it has been seen twice as often."""
    if not _graphs_enabled():
        return None
    key = (n_seq, K, recompute)
    seen = enc.__dict__.setdefault("_train_shapes_seen", {})
    seen[key] = seen.get(key, 0) + 1
    cache = enc.__dict__.setdefault("_train_graphs", {})
    gs = cache.get(key)
    if gs is None:
        captures = enc.__dict__.get("_train_captures", 0)
        if seen[key] < _GRAPH_MIN_SIGHTINGS or captures >= _GRAPH_MAX_CAPTURES:
            return None
        if len(cache) >= _GRAPH_SHAPES:
            coldest = min(cache, key=lambda k: seen.get(k, 0))
            if cache[coldest].busy or seen[key] < 2 * seen.get(coldest, 0):
                return None
            del cache[coldest]
        tape_b, ws_b, gf = _sizes(enc, n_seq, K, recompute)
        if tape_b + ws_b + 4 * gf > 0.4 * torch.cuda.mem_get_info(dev)[0]:    # the graph's buffers stay allocated: only when
            seen[key] = -(1 << 30)                                             # they fit comfortably (decided once per shape)
            return None
        enc._train_captures = captures + 1
        gs = cache[key] = _GraphedStep(enc, n_seq, K, recompute, dev)
    return gs


def _capture(fn):
    """Capture ``fn`` (kernel launches on the current stream only) into a CUDA graph; returns (graph, launches captured)."""
    g = torch.cuda.CUDAGraph()
    l0 = ops.kernel_launch_count()
    with torch.cuda.graph(g, capture_error_mode="thread_local"):       # (the backward is captured on autograd's thread)
        fn()
    return g, ops.kernel_launch_count() - l0


def _call_fwd(enc: PackedEncoder, ids, n_seq, K, out, tape_buf, tape_bytes, recompute, dev) -> None:
    with torch.cuda.device(dev):
        _lib.check(_lib.load().molly_encode_train_fwd(enc.handle, ids.data_ptr(), n_seq, K, out.data_ptr(), tape_buf.data_ptr(),
                                                      tape_bytes, int(recompute), ops.error_flag(dev).data_ptr(),
                                                      ops._stream(dev)), "molly_encode_train_fwd")


def encoder_forward_train(enc: PackedEncoder, ids: torch.Tensor) -> Tuple[torch.Tensor, EncoderTape]:
    """``hidden_states[-1]`` (bf16 [n*K, h]) with the tape; numerically the inference forward (same kernels, same order).
    From the third call with a shape on, the launches are replayed from a CUDA graph (``_GraphedStep``): the returned tensor
    and the tape are then buffers of that graph, valid until the next training forward of this encoder with the same shape."""
    cfg = enc.cfg
    if cfg.emb_layer_norm_before:
        raise NotImplementedError("emb_layer_norm_before encoders are not covered by the training path")
    dev = ops._require_cuda(ids)
    ids = ids.contiguous()
    n_seq, K = ids.shape
    tape = EncoderTape()
    tape.recompute = not _save_activations(enc, n_seq * K)
    gs = _graphed_step(enc, n_seq, K, tape.recompute, dev)
    if gs is not None and gs.try_acquire(tape):
        gs.ids.copy_(ids)
        if gs.fwd is None:                     # (the launches were counted while capturing: this first replay is not added)
            gs.fwd, gs.fwd_launches = _capture(lambda: _call_fwd(enc, gs.ids, n_seq, K, gs.out, gs.tape, gs.tape_bytes,
                                                                 gs.recompute, dev))
        else:
            ops.add_kernel_launches(gs.fwd_launches)
        gs.fwd.replay()
        tape.ids, tape.buf, tape.graph = gs.ids, gs.tape, gs
        return gs.out, tape
    tape.ids = ids
    tape_bytes, _, _ = _sizes(enc, n_seq, K, tape.recompute)
    tape.buf = _device_buffer(tape_bytes, dev)
    out = torch.empty(n_seq * K, cfg.hidden_size, dtype=torch.bfloat16, device=dev)
    _call_fwd(enc, ids, n_seq, K, out, tape.buf, tape_bytes, tape.recompute, dev)
    return out, tape


def _embedding_backward(enc: PackedEncoder, plan: "GradPlan", ids: torch.Tensor, d_x: torch.Tensor, tail: torch.Tensor) -> None:
    """Embedding tables: scatter-add of the residual-stream gradient (token-dropout scale, pad / <mask> rows skipped)."""
    views = plan.tail_views(tail)
    word_index, word_scale, pos_index, pos_scale = _emb_meta(enc, ids)
    d_word = views["esm.embeddings.word_embeddings.weight"]
    d_word.zero_()
    ops.scatter_add_rows_(d_word, d_x, word_index, word_scale)
    if pos_index is not None:
        d_pos = views["esm.embeddings.position_embeddings.weight"]
        d_pos.zero_()
        ops.scatter_add_rows_(d_pos, d_x, pos_index, pos_scale)


def encoder_backward(enc: PackedEncoder, tape: EncoderTape, d_out: torch.Tensor, reducer=None) -> Dict[str, torch.Tensor]:
    """Gradients of every encoder parameter (HF ``state_dict`` names, fp32 views of one flat buffer) from ``d_out`` =
    d(loss)/d(hidden_states[-1]), bf16 [n*K, h].  ``reducer`` (``dist.LayerwiseGradReducer``): each layer's flat group is
    handed to it as soon as the layer's backward is enqueued, so its all-reduce overlaps the layers below; those gradients
    come back averaged, in the reducer's dtype (the reducer is finished -- current stream ordered after every collective --
    before this returns).  After a graphed forward (and without a reducer) the backward is one graph replay as well and the
    returned tensors are views of that graph's gradient buffer: consume them before the next backward of this encoder."""
    cfg = enc.cfg
    n_seq, K = tape.ids.shape
    h, L = cfg.hidden_size, cfg.num_hidden_layers
    dev = ops._require_cuda(d_out)
    if d_out.dtype != torch.bfloat16 or tuple(d_out.shape) != (n_seq * K, h) or not d_out.is_contiguous():
        raise ValueError("encoder_backward: d_out must be contiguous bf16 [n_seq*K, h]")
    plan = grad_plan(enc)
    gs: Optional[_GraphedStep] = tape.graph
    if gs is not None and tape.gen != gs.gen:
        raise RuntimeError("encoder_backward: the CUDA-graph buffers of this forward were taken over by a later training forward "
                           f"of the same encoder and shape (it was passed over more than {_GRAPH_ABANDON} times)")
    tape_bytes, ws_bytes, grad_floats = _sizes(enc, n_seq, K, tape.recompute)
    assert grad_floats == plan.total
    graphed = gs is not None and reducer is None and _graphs_enabled()
    if graphed:
        flat, ws, d_src = gs.flat, gs.ws, gs.d_out
        d_src.copy_(d_out)
    else:
        flat = torch.empty(plan.total, dtype=torch.float32, device=dev)
        ws = _device_buffer(ws_bytes, dev)
        d_src = d_out
    lib = _lib.load()

    def run(layer_begin: int, layer_end: int) -> None:
        with torch.cuda.device(dev):
            _lib.check(lib.molly_encode_train_bwd(enc.handle, n_seq, K, tape.buf.data_ptr(), tape_bytes, int(tape.recompute),
                                                  d_src.data_ptr(), flat.data_ptr(), ws.data_ptr(), ws_bytes, layer_begin,
                                                  layer_end, ops._stream(dev)), "molly_encode_train_bwd")

    # the running gradient of the residual stream is the first M*h floats of the workspace: d(embedding output) after layer 0
    d_x = ws[:n_seq * K * h * 4].view(torch.float32).view(n_seq * K, h)
    tail = flat[plan.tail_begin:]
    grads: Dict[str, torch.Tensor] = {}
    try:
        if graphed:
            if gs.bwd is None:
                def whole():
                    run(L, 0)
                    _embedding_backward(enc, plan, tape.ids, d_x, tail)
                gs.bwd, gs.bwd_launches = _capture(whole)
            else:
                ops.add_kernel_launches(gs.bwd_launches)
            gs.bwd.replay()
        elif reducer is None:
            run(L, 0)
        else:
            for i in range(L - 1, -1, -1):
                run(L if i == L - 1 else i, i)                    # (the first call also does emb_layer_norm_after)
                grads.update(plan.layer_views(i, reducer.reduce_flat_(flat[i * plan.group:(i + 1) * plan.group])))
        if reducer is None:
            for i in range(L):
                grads.update(plan.layer_views(i, flat[i * plan.group:(i + 1) * plan.group]))
        if not graphed:
            _embedding_backward(enc, plan, tape.ids, d_x, tail)
        if reducer is not None:
            tail = reducer.reduce_flat_(tail)
            reducer.finish()                       # every group is averaged before anything reads the flat buffers
        grads.update(plan.tail_views(tail))
    finally:
        if gs is not None:
            gs.release(tape)
    return plan.finalize_(grads)
