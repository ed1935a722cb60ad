"""Multi-GPU plumbing of the path.  The forward shards by SAMPLE with no collective at all; the only
exchange is the projector(/LoRA) gradient all-reduce of the training configuration, which the reference gets from
DeepSpeed ZeRO-0/2 (src/configs/ds_z0_config.json:18-27, 5e8-element buckets, overlap_comm false).  Here: ONE flat
bucket per step (projector grads are 4.7 M elements at Molly-1.7B -- latency-, not bandwidth-bound over NVSwitch),
all-reduced (mean) with NCCL on a side stream so it overlaps whatever backward work is still queued.
Works with the gloo backend too (CPU tests)."""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch
import torch.distributed as dist

from . import planner


def shard_batch(rank: int, world_size: int, hidden_states: torch.Tensor, omic_ids, omic_info_list):
    """Rank r owns samples [r*B/W, (r+1)*B/W) of the global batch (views, no copies)."""
    r = planner.shard_samples(hidden_states.shape[0], world_size, rank)
    sl = slice(r.start, r.stop)
    return hidden_states[sl], omic_ids[sl], omic_info_list[sl]


class FlatGradBucket:
    """One persistent flat buffer holding the grads of the trainable path parameters (projector weight/bias of both
    modalities, optionally LoRA adapters), all-reduced in a single collective per step."""

    def __init__(self, params: Sequence[torch.nn.Parameter], group: Optional[dist.ProcessGroup] = None,
                 dtype: Optional[torch.dtype] = None):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("FlatGradBucket needs at least one trainable parameter")
        self.group = group
        dev = self.params[0].device
        self.dtype = dtype or self.params[0].dtype
        self.numel = sum(p.numel() for p in self.params)
        # (a one-tensor bucket reduces its gradient in place and only needs `flat` if that gradient is missing / strided)
        self.flat = torch.zeros(self.numel if len(self.params) > 1 else 1, dtype=self.dtype, device=dev)
        self.stream = torch.cuda.Stream(device=dev) if dev.type == "cuda" else None
        self._work = None
        self._inplace = False

    def _views(self) -> List[torch.Tensor]:
        out, off = [], 0
        for p in self.params:
            out.append(self.flat[off:off + p.numel()].view_as(p))
            off += p.numel()
        return out

    def _mean_all_reduce(self, t: torch.Tensor):
        """Mean over ranks: NCCL averages inside the collective (ReduceOp.AVG); gloo has no AVG -> pre-divide + SUM."""
        if dist.get_backend(self.group) == "nccl":
            return dist.all_reduce(t, op=dist.ReduceOp.AVG, group=self.group, async_op=True)
        t.div_(dist.get_world_size(self.group))
        return dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group, async_op=True)

    def launch(self) -> None:
        """Start the (async) mean all-reduce.  Call right after the backward produced the gradients.  A bucket of ONE
        contiguous gradient (the LoRA-sized tensor of cfg-5) is reduced in place: no pack / unpack passes over it."""
        if self.stream is not None:
            self.stream.wait_stream(torch.cuda.current_stream(self.flat.device))
        ctx = torch.cuda.stream(self.stream) if self.stream is not None else _Null()
        with ctx:
            g = self.params[0].grad if len(self.params) == 1 else None
            if g is not None and g.is_contiguous() and g.dtype == self.dtype:
                self._inplace = True
                if self.stream is not None:
                    g.record_stream(self.stream)
                self._work = self._mean_all_reduce(g.view(-1))
                return
            self._inplace = False
            if self.flat.numel() != self.numel:
                self.flat = torch.zeros(self.numel, dtype=self.dtype, device=self.flat.device)
            for v, p in zip(self._views(), self.params):
                if p.grad is None:
                    v.zero_()
                else:
                    v.copy_(p.grad)
            self._work = self._mean_all_reduce(self.flat)

    def finish(self) -> None:
        """Wait for the collective and scatter the averaged grads back into ``p.grad``."""
        if self._work is None:
            return
        ctx = torch.cuda.stream(self.stream) if self.stream is not None else _Null()
        with ctx:
            # Work.wait() orders the CURRENT stream after NCCL's internal stream: it has to run inside the side-stream
            # context, or the copy-back below could read `flat` while it is still being reduced.
            self._work.wait()
            if not self._inplace:
                for v, p in zip(self._views(), self.params):
                    if p.grad is None:
                        p.grad = v.clone()
                    else:
                        p.grad.copy_(v)
        if self.stream is not None:
            torch.cuda.current_stream(self.flat.device).wait_stream(self.stream)
        self._work = None


class LayerwiseGradReducer:
    """Mean all-reduce of the encoder gradients WHILE the backward is still running (``--train-bio`` at N > 1).

    ``train.encoder_backward`` hands over each layer's gradients as soon as they exist: they are flattened into one
    buffer in the parameters' dtype (the dict entries become views of it), and its all-reduce is launched on a side stream
    so that NCCL traffic over NVLink overlaps the backward kernels of the layers below.  ``finish()`` makes the current
    stream wait for every collective; the gradients returned to autograd are already averaged, so the caller must not
    reduce them again.  Works with gloo on CPU tensors too (no streams)."""

    def __init__(self, group: Optional[dist.ProcessGroup] = None, dtype: torch.dtype = torch.bfloat16):
        self.group, self.dtype = group, dtype
        self._stream = None
        self._works = []
        self.bytes_reduced = 0

    def reduce_(self, grads: dict, names: Sequence[str]) -> None:
        names = [n for n in names if n in grads and grads[n] is not None]
        if not names or dist.get_world_size(self.group) == 1:
            return
        world = dist.get_world_size(self.group)
        flat = torch.cat([grads[n].reshape(-1).to(self.dtype) for n in names])
        flat.div_(world)
        off = 0
        for n in names:
            k = grads[n].numel()
            grads[n] = flat[off:off + k].view(grads[n].shape)
            off += k
        self.bytes_reduced += flat.numel() * flat.element_size()
        if flat.is_cuda:
            cur = torch.cuda.current_stream(flat.device)
            if self._stream is None:
                self._stream = torch.cuda.Stream(flat.device)
            self._stream.wait_stream(cur)
            flat.record_stream(self._stream)
            with torch.cuda.stream(self._stream):
                self._works.append(dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        else:
            self._works.append(dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def reduce_flat_(self, flat: torch.Tensor) -> torch.Tensor:
        """One group that already is one contiguous buffer (the library's flat gradient layout, ``train.GradPlan``): cast to
        the reducer's dtype (one pass), mean all-reduce on the side stream; returns the buffer the averaged values land in."""
        g = {"_flat": flat}
        self.reduce_(g, ["_flat"])
        return g["_flat"]

    def reduce_zeros_(self, numel: int, device) -> torch.Tensor:
        """The collective a rank issues in place of ``reduce_`` for gradients it does not have (a modality absent from its
        micro-batch): same size, same dtype, same position in the schedule, all zeros.  Returns what every rank gets back:
        the averaged gradients of the ranks that had the modality."""
        z = torch.zeros(max(0, int(numel)), dtype=self.dtype, device=device)
        if numel <= 0 or dist.get_world_size(self.group) == 1:
            return z
        return self.reduce_flat_(z)

    def finish(self) -> None:
        for w in self._works:
            w.wait()
        self._works = []
        if self._stream is not None:
            torch.cuda.current_stream(self._stream.device).wait_stream(self._stream)


class _Null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False
