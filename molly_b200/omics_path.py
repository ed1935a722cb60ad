"""Drop-in for ``OmicsOne.process_omic_sequences`` (reference ``src/model/omics_one.py:49-136``).

``FastOmicsPath.process_omic_sequences(hidden_states, omic_ids_list, omic_info_list, device)`` keeps the reference's
signature, tensor layouts, in-place + same-object return, modality order (DNA/RNA first, then protein, :120-134),
pairing-by-index quirk and exception types, while the work underneath is the sm_100a library:

    host: route pairs -> one id matrix + one (b, start) table per modality, one H2D each
    device: embed -> L x [LN, fused QKV GEMM, rotary, fused MHA, out-proj GEMM(+bias+residual), LN,
                          FFN1 GEMM(+bias+erf-GELU | gated SiLU), FFN2 GEMM(+bias+residual)] -> final LN
            -> projector GEMM whose epilogue stores rows straight into hidden_states[b, start+1+j, :]

``install(omics_one)`` swaps the method on a live reference ``OmicsOne`` instance; weights are read from its
``dna_rna_model`` / ``protein_model`` / ``*_projector`` modules' state dicts.
"""
from __future__ import annotations

import os
import types
from typing import Any, List, Mapping, Optional

import torch

from . import _lib, ops, planner
from .config import EncoderConfig
from .packing import PackedEncoder


class _InjectFn(torch.autograd.Function):
    """Autograd of one modality: grads reach the projector ``weight`` / ``bias``; the overwritten rows of
    ``hidden_states`` get zero grad (what autograd derives for the reference's slice-assign); encoders are frozen
    (reference: ``set_up_trainable_param`` / ``pre_train_lora``, src/utils/tools.py:313-338, 345-396)."""

    @staticmethod
    def forward(ctx, hidden_states, proj_weight, proj_bias, ids, seq_table, enc_id: int):
        need_grad = proj_weight.requires_grad or proj_bias.requires_grad
        enc_out = ops.encode_project_merge(hidden_states, ids, seq_table, enc_id, bool(need_grad))
        ctx.mark_dirty(hidden_states)
        ctx.enc_id = enc_id
        ctx.k_tokens = ids.shape[1]
        ctx.save_for_backward(seq_table, enc_out)
        ctx.w_meta, ctx.b_meta = (proj_weight.dtype, proj_weight.device), (proj_bias.dtype, proj_bias.device)
        return hidden_states

    @staticmethod
    def backward(ctx, grad_out):
        seq_table, enc_out = ctx.saved_tensors
        need_h = ctx.needs_input_grad[0]
        g = grad_out.contiguous()
        if need_h:
            g = g.clone() if g.data_ptr() == grad_out.data_ptr() else g
        dW = db = None
        if enc_out.numel() > 0:
            dW, db = ops.project_bwd(g, seq_table, enc_out, ctx.enc_id, ctx.k_tokens, bool(need_h))
            dW = dW.to(device=ctx.w_meta[1], dtype=ctx.w_meta[0])
            db = db.to(device=ctx.b_meta[1], dtype=ctx.b_meta[0])
        elif need_h:
            # frozen projector (--train-llm without --train-mlp): nothing was saved for dW, but the rows the slice-assign
            # overwrote must still get ZERO gradient (omics_one.py:97), or the *_pad embedding rows pick up the upstream one
            k_cap = min(ops.get_encoder(ctx.enc_id).project_token_num, ctx.k_tokens)
            ops.gather_rows(g, seq_table, ctx.k_tokens, k_cap, True)
        return (g if need_h else None), dW, db, None, None, None


class _InjectTrainFn(torch.autograd.Function):
    """Autograd of one modality with a TRAINABLE encoder (``--train-bio``, src/utils/tools.py:313-331): forward through
    ``train.encoder_forward_train`` (same kernels, layer inputs kept), backward through ``train.encoder_backward``; grads
    reach every encoder parameter (by HF name), the projector, and -- zeroed on the overwritten rows -- ``hidden_states``."""

    @staticmethod
    def forward(ctx, hidden_states, proj_weight, proj_bias, ids, seq_table, enc_id: int, names, reducer, *enc_params):
        from . import train
        enc = ops.get_encoder(enc_id)
        ctx.reducer = reducer
        k_tokens = ids.shape[1]
        enc_out, tape = train.encoder_forward_train(enc, ids)
        ops.gemm_bf16(enc_out, enc.proj_w, _lib.EPI_SCATTER, bias=enc.proj_b, out=hidden_states, seq_table=seq_table,
                      seq_k=k_tokens, k_cap=min(enc.project_token_num, k_tokens))
        ctx.mark_dirty(hidden_states)
        ctx.enc_id, ctx.k_tokens, ctx.names, ctx.tape = enc_id, k_tokens, names, tape
        ctx.save_for_backward(seq_table, enc_out)
        ctx.metas = ((proj_weight.dtype, proj_weight.device), (proj_bias.dtype, proj_bias.device),
                     [(p.dtype, p.device) for p in enc_params])          # grads go back in each parameter's dtype / device
        return hidden_states

    @staticmethod
    def backward(ctx, grad_out):
        from . import train
        seq_table, enc_out = ctx.saved_tensors
        enc = ops.get_encoder(ctx.enc_id)
        need_h = ctx.needs_input_grad[0]
        g = grad_out.contiguous()
        if need_h:
            g = g.clone() if g.data_ptr() == grad_out.data_ptr() else g
        dy = ops.gather_rows(g, seq_table, ctx.k_tokens, min(enc.project_token_num, ctx.k_tokens), bool(need_h))
        dW, db = ops.linear_wgrad(dy, enc_out)
        d_enc = ops.gemm_bf16(dy, ops.transpose_bf16(enc.proj_w), _lib.EPI_BIAS)
        reducer = ctx.reducer
        if reducer is not None:
            pg = {"projector.weight": dW, "projector.bias": db}
            reducer.reduce_(pg, list(pg))
            dW, db = pg["projector.weight"], pg["projector.bias"]
        static = ctx.tape.graph is not None          # gradients are views of a graph's buffer: never hand those to autograd
        grads = train.encoder_backward(enc, ctx.tape, d_enc, reducer)
        if reducer is not None:
            reducer.finish()
        ctx.tape = None
        w_m, b_m, p_ms = ctx.metas
        # every gradient goes back in its parameter's dtype / device: ONE multi-tensor cast instead of a kernel per parameter
        enc_grads, srcs, dsts = [], [], []
        for n, (dt, dv), need in zip(ctx.names, p_ms, ctx.needs_input_grad[8:]):
            if n not in grads or not need:
                enc_grads.append(None)
                continue
            src = grads[n]
            if src.dtype == dt and src.device == dv and not static:
                enc_grads.append(src)
                continue
            dst = torch.empty(src.shape, dtype=dt, device=dv)
            if dv == src.device:
                srcs.append(src)
                dsts.append(dst)
            else:
                dst.copy_(src)
            enc_grads.append(dst)
        if dsts:
            torch._foreach_copy_(dsts, srcs)
        return ((g if need_h else None), dW.to(device=w_m[1], dtype=w_m[0]), db.to(device=b_m[1], dtype=b_m[0]), None, None,
                None, None, None, *enc_grads)


class _AbsentModalityFn(torch.autograd.Function):
    """A modality with a TRAINABLE encoder that has no sequence in this rank's micro-batch while a gradient reducer is
    attached: the other ranks all-reduce that encoder's gradients layer by layer inside their backward, so this rank must
    issue the SAME sequence of collectives at the same point of its backward, or NCCL pairs the wrong buffers / hangs.  It
    contributes zeros sized like the groups of ``train.GradPlan`` and gets back what every rank gets: the averaged gradients.
    Forward is the identity on ``hidden_states``."""

    @staticmethod
    def forward(ctx, hidden_states, reducer, plan, names, *params):
        ctx.reducer, ctx.plan, ctx.names = reducer, plan, names
        ctx.metas = [(p.dtype, p.device) for p in params]
        ctx.mark_dirty(hidden_states)
        return hidden_states

    @staticmethod
    def backward(ctx, grad_out):
        plan, red, dev = ctx.plan, ctx.reducer, grad_out.device
        got = {}
        flat = red.reduce_zeros_(plan.group_numel(0), dev)                   # the projector group
        n_w = plan.proj_shapes[0][0] * plan.proj_shapes[0][1]
        got["projector.weight"], got["projector.bias"] = flat[:n_w].view(plan.proj_shapes[0]), flat[n_w:]
        for g in range(1, plan.L + 1):                                       # layers L-1 .. 0, as the backward produces them
            got.update(plan.layer_views(plan.L - g, red.reduce_zeros_(plan.group_numel(g), dev)))
        got.update(plan.tail_views(red.reduce_zeros_(plan.group_numel(plan.L + 1), dev)))
        red.finish()
        plan.finalize_(got)                                                  # (copies: only after the collectives are done)
        grads = [got[n].to(device=dv, dtype=dt) if (n in got and need) else None
                 for n, (dt, dv), need in zip(ctx.names, ctx.metas, ctx.needs_input_grad[4:])]
        return (grad_out if ctx.needs_input_grad[0] else None), None, None, None, *grads


class FastOmicsPath:
    """B200-native encode -> project -> merge behind the reference's call boundary."""

    def __init__(self, dna_rna: Optional[PackedEncoder], protein: Optional[PackedEncoder], strict: bool = False):
        self.dna_rna, self.protein = dna_rna, protein
        self._ids = {}
        for name, enc in (("dna_rna", dna_rna), ("protein", protein)):
            if enc is not None:
                self._ids[name] = ops.register_encoder(enc)
        self.strict = strict                # True: synchronise and raise device-side errors inside the call
        # Run the two modalities' encoders on two streams when their written rows are disjoint (always, for dataset-built
        # batches; checked per call) and the batch is small enough to leave SMs idle: -33 % latency at B=1, -9 % at B=4,
        # nothing at the power-capped headline batch (profiles/r01n_graph.md), hence the row threshold.
        # MOLLY_CONCURRENT_MODALITIES=0 restores strictly sequential launches, =2 lifts the threshold.
        mode = os.environ.get("MOLLY_CONCURRENT_MODALITIES", "1")
        self.concurrent = mode != "0"
        self.concurrent_max_rows = (1 << 62) if mode == "2" else 32768
        self._side_streams = {}
        self._proj_modules = {}             # name -> nn.Linear (live parameters, for --train-mlp)
        self._proj_versions = {}
        self._enc_modules = {}              # name -> EsmForMaskedLM (live parameters: --train-bio, late checkpoint loads)
        self._enc_versions = {}
        # optional dist.LayerwiseGradReducer: the path then returns ALREADY AVERAGED encoder / projector gradients
        self.grad_reducer = None

    # ------------------------------------------------------------------ construction
    @classmethod
    def from_state_dicts(cls, *, dna_rna_cfg: Optional[EncoderConfig], dna_rna_state: Optional[Mapping[str, Any]],
                         dna_rna_projector: Optional[Mapping[str, Any]], dna_rna_project_token_num: int,
                         protein_cfg: Optional[EncoderConfig], protein_state: Optional[Mapping[str, Any]],
                         protein_projector: Optional[Mapping[str, Any]], protein_project_token_num: int,
                         device, strict: bool = False) -> "FastOmicsPath":
        device = torch.device(device)
        nt = pr = None
        if dna_rna_cfg is not None:
            nt = PackedEncoder(dna_rna_cfg, dna_rna_state, dna_rna_projector, dna_rna_project_token_num, device)
        if protein_cfg is not None:
            pr = PackedEncoder(protein_cfg, protein_state, protein_projector, protein_project_token_num, device)
        return cls(nt, pr, strict=strict)

    @classmethod
    def from_omics_one(cls, om, device, strict: bool = False) -> "FastOmicsPath":
        """Build from a reference ``OmicsOne`` (omics_one.py:10-30): same three modules per modality."""
        def one(model, projector, k):
            if model is None:
                return None
            sd = model.state_dict()
            return PackedEncoder(EncoderConfig.from_hf_config(model.config, sd), sd, projector.state_dict(), k,
                                 torch.device(device))
        self = cls(one(om.dna_rna_model, om.dna_rna_projector, om.dna_rna_project_token_num),
                   one(om.protein_model, om.protein_projector, om.protein_project_token_num), strict=strict)
        self._proj_modules = {"dna_rna": om.dna_rna_projector, "protein": om.protein_projector}
        self._enc_modules = {k: m for k, m in (("dna_rna", om.dna_rna_model), ("protein", om.protein_model))
                             if m is not None}
        for name, m in self._enc_modules.items():
            self._enc_versions[name] = self._module_version(m)
        return self

    def install(self, om) -> None:
        """Replace ``om.process_omic_sequences`` (the frozen signature of omics_one.py:49-55) with this path."""
        path = self

        def process_omic_sequences(self_om, hidden_states, omic_ids_list, omic_info_list, device):
            return path.process_omic_sequences(hidden_states, omic_ids_list, omic_info_list, device)

        om.process_omic_sequences = types.MethodType(process_omic_sequences, om)

    def close(self) -> None:
        for eid in self._ids.values():
            ops.unregister_encoder(eid)
        self._ids = {}

    # ------------------------------------------------------------------ the boundary
    def process_omic_sequences(self, hidden_states: torch.Tensor, omic_ids_list, omic_info_list: List[List[dict]],
                               device) -> torch.Tensor:
        if not hidden_states.is_cuda:
            raise RuntimeError("molly_b200.process_omic_sequences needs hidden_states on a CUDA device "
                               "(the B200 path has no CPU fallback)")
        dev = hidden_states.device
        batch_size = hidden_states.shape[0]
        if self._enc_modules:
            self.refresh_encoders()
        nt_plan, pr_plan = planner.route(batch_size, omic_ids_list, omic_info_list)          # may raise ValueError
        # reference order: all DNA/RNA sequences, then all protein sequences (omics_one.py:120-134)
        work = [(name, plan) for name, plan in (("dna_rna", nt_plan), ("protein", pr_plan)) if len(plan)]  # :67-68
        if (self.concurrent and len(work) == 2 and hidden_states.is_contiguous()
                and max(len(nt_plan), len(pr_plan)) * self._k_ids(omic_ids_list, work[0][1]) <= self.concurrent_max_rows
                and not (torch.is_grad_enabled() and self._proj_modules)
                and planner.ranges_disjoint(nt_plan, self._k_rows("dna_rna", omic_ids_list, nt_plan),
                                            pr_plan, self._k_rows("protein", omic_ids_list, pr_plan))):
            cur = torch.cuda.current_stream(dev)
            side = self._side_streams.setdefault(dev.index, torch.cuda.Stream(dev))
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                self._inject(work[1][0], work[1][1], hidden_states, omic_ids_list, dev)
            self._inject(work[0][0], work[0][1], hidden_states, omic_ids_list, dev)
            cur.wait_stream(side)
        else:
            for name, plan in (("dna_rna", nt_plan), ("protein", pr_plan)):
                if len(plan):
                    self._inject(name, plan, hidden_states, omic_ids_list, dev)
                elif self.grad_reducer is not None and torch.is_grad_enabled():
                    self._absent_modality(name, hidden_states)       # same position in every rank's autograd graph
        if self.strict:
            ops.check_device_errors(dev)
        return hidden_states

    @staticmethod
    def _k_ids(omic_ids_list, plan: planner.ModalityPlan) -> int:
        return int(omic_ids_list.shape[-1] if isinstance(omic_ids_list, torch.Tensor) else
                   omic_ids_list[plan.b_idx[0]][plan.slot_idx[0]].shape[-1])

    def _k_rows(self, name: str, omic_ids_list, plan: planner.ModalityPlan) -> int:
        """Rows written per sequence: min(project_token_num, K) (omics_one.py:96)."""
        k_ids = self._k_ids(omic_ids_list, plan)
        enc_id = self._ids.get(name)
        cap = ops.get_encoder(enc_id).project_token_num if enc_id is not None else int(k_ids)
        return min(cap, int(k_ids))

    def _absent_modality(self, name: str, hidden_states: torch.Tensor) -> None:
        """Keep the reducer's collective schedule independent of the batch (see ``_AbsentModalityFn``)."""
        proj, enc_module, enc_id = self._proj_modules.get(name), self._enc_modules.get(name), self._ids.get(name)
        if proj is None or enc_module is None or enc_id is None:
            return
        named = [(n, prm) for n, prm in enc_module.named_parameters()]
        if not any(prm.requires_grad for _, prm in named):        # frozen encoder: its backward launches no collective
            return
        from . import train
        named += [("projector.weight", proj.weight), ("projector.bias", proj.bias)]
        _AbsentModalityFn.apply(hidden_states, self.grad_reducer, train.grad_plan(ops.get_encoder(enc_id)),
                                tuple(n for n, _ in named), *[prm for _, prm in named])

    def _inject(self, name: str, plan: planner.ModalityPlan, hidden_states: torch.Tensor, omic_ids_list, dev) -> None:
        enc_id = self._ids.get(name)
        if enc_id is None:
            raise RuntimeError(f"Error processing omic sequences: no {name} encoder is loaded")
        enc = ops.get_encoder(enc_id)
        ids = planner.gather_ids(omic_ids_list, plan)                                        # may raise RuntimeError
        if not ids.is_cuda:
            planner.check_vocab(ids, enc.cfg.vocab_size)                                     # AssertionError, :71-72
            ids = ids.pin_memory().to(dev, non_blocking=True) if torch.cuda.is_available() else ids.to(dev)
        _, T, D = hidden_states.shape
        k = min(enc.project_token_num, ids.shape[1])
        planner.check_placement(plan, k, hidden_states.shape[0], T)                          # RuntimeError, :97
        seq_table = plan.seq_table().pin_memory().to(dev, non_blocking=True)
        target = hidden_states
        if not hidden_states.is_contiguous():
            target = hidden_states.contiguous()
        proj = self._proj_modules.get(name)
        if proj is not None:
            self._refresh_projector(name, enc, proj)
        enc_module = self._enc_modules.get(name)
        train_enc = (proj is not None and enc_module is not None and torch.is_grad_enabled()
                     and any(prm.requires_grad for prm in enc_module.parameters()))
        if train_enc:                                        # --train-bio: the encoder's own parameters get gradients
            named = list(enc_module.named_parameters())
            _InjectTrainFn.apply(target, proj.weight, proj.bias, ids, seq_table, enc_id, tuple(n for n, _ in named),
                                 self.grad_reducer, *[prm for _, prm in named])
        elif proj is not None and torch.is_grad_enabled() and (proj.weight.requires_grad or proj.bias.requires_grad
                                                                or hidden_states.requires_grad):
            _InjectFn.apply(target, proj.weight, proj.bias, ids, seq_table, enc_id)
        else:
            ops.encode_project_merge(target, ids, seq_table, enc_id, False)
        if target is not hidden_states:
            hidden_states.copy_(target)

    # ------------------------------------------------------------------ the input producer, fused
    @torch.no_grad()
    def embed_and_process(self, input_ids: torch.Tensor, embed_weight: torch.Tensor, omic_ids_list,
                          omic_info_list: List[List[dict]], pad_token_ids) -> torch.Tensor:
        """``process_omic_sequences(embed_tokens(input_ids), omic_ids, omic_info_list)`` (omics_one.py:164-170 and
        :209-215) in one pass, forward only (eval / ``generate``).  The placeholder runs found in ``input_ids`` ON THE DEVICE
        are the index source: only ``info["type"]`` is read, ``info["start"]`` is what the run scan reproduces
        (omics_dataset.py:270-288).  The LLM embedding lookup never touches the rows the projector is about to overwrite.
        ``pad_token_ids`` = (dna_pad, rna_pad, protein_pad) token ids.  A layout that does not pair with the ids raises
        ``RuntimeError`` in strict mode (MOLLY_ERRBIT_LAYOUT)."""
        if input_ids.dim() != 2 or embed_weight.dim() != 2:
            raise ValueError("input_ids must be [B, T] and embed_weight [vocab, D]")
        dev = embed_weight.device
        input_ids = input_ids.to(dev, torch.int64, non_blocking=True).contiguous()
        meta = self._fused_meta(input_ids.shape[0], omic_ids_list, omic_info_list, dev)
        ids = {}
        for name, plan in meta["plans"].items():
            t = planner.gather_ids(omic_ids_list, plan)
            if not t.is_cuda:
                planner.check_vocab(t, ops.get_encoder(self._ids[name]).cfg.vocab_size)
                t = t.pin_memory().to(dev, non_blocking=True)
            ids[name] = t
        hidden = self._fused_run(input_ids, embed_weight, ids, meta, tuple(pad_token_ids))
        if self.strict:
            ops.check_device_errors(dev, "embed_and_process")
        return hidden

    def _fused_meta(self, batch_size: int, omic_ids_list, omic_info_list, dev) -> dict:
        """Host-side routing of one batch layout -> device index tables (constant for a CUDA-graph bucket)."""
        for i in range(len(omic_ids_list)):                                                  # omics_one.py:166-170
            assert len(omic_ids_list[i]) == len(omic_info_list[i]), f"Mismatch in omic count vs info count at index {i}"
        nt_plan, pr_plan = planner.route(batch_size, omic_ids_list, omic_info_list)         # may raise ValueError
        n_slots = [0] * batch_size
        plans, caps, idx = {}, {"dna_rna": 0, "protein": 0}, {}
        for name, plan in (("dna_rna", nt_plan), ("protein", pr_plan)):
            if len(plan) == 0:
                continue
            if name not in self._ids:
                raise RuntimeError(f"Error processing omic sequences: no {name} encoder is loaded")
            for b, r in zip(plan.b_idx, plan.run_idx):
                n_slots[b] = max(n_slots[b], r + 1)
            k_ids = omic_ids_list.shape[-1] if isinstance(omic_ids_list, torch.Tensor) else \
                omic_ids_list[plan.b_idx[0]][plan.slot_idx[0]].shape[-1]
            caps[name] = min(ops.get_encoder(self._ids[name]).project_token_num, int(k_ids))
            plans[name] = plan
            idx[name] = torch.tensor([plan.b_idx, plan.run_idx, plan.slot_idx], dtype=torch.int32).pin_memory().to(
                dev, non_blocking=True)
        slots_dev = torch.tensor(n_slots, dtype=torch.int32).pin_memory().to(dev, non_blocking=True)
        max_runs = max(1, max(n_slots))
        expect = torch.full((batch_size, max_runs), -1, dtype=torch.int32)       # modality each run pairs with (0 / 1), -1: none
        for name, plan in plans.items():
            for b, r in zip(plan.b_idx, plan.run_idx):
                expect[b, r] = 1 if name == "protein" else 0
        return {"plans": plans, "caps": caps, "idx": idx, "slots": slots_dev, "max_runs": max_runs,
                "slot_expect": expect.pin_memory().to(dev, non_blocking=True)}

    def _fused_run(self, input_ids, embed_weight, ids: dict, meta: dict, pad_token_ids) -> torch.Tensor:
        """Device-only part (capturable): run scan -> skipping embedding lookup -> per modality seq_table + encode/merge."""
        runs = ops.placeholder_runs(input_ids, pad_token_ids, meta["slots"], meta["max_runs"])
        # runs the seq_table will reject are embedded like text: no row of the (uninitialised) output stays unwritten
        ops.placeholder_reject(runs, meta["slot_expect"], max(1, meta["caps"]["dna_rna"]), max(1, meta["caps"]["protein"]))
        hidden = ops.embed_tokens_skip(input_ids, runs[4], pad_token_ids, meta["caps"]["dna_rna"], meta["caps"]["protein"],
                                       embed_weight)

        def one(name):
            enc_id = self._ids[name]
            idx = meta["idx"][name]
            seq_table = ops.build_seq_table(idx[0], idx[1], runs, expect_protein=(name == "protein"),
                                            k_need=max(1, meta["caps"][name]))    # shorter run = text rows overwritten
            proj = self._proj_modules.get(name)
            if proj is not None:
                self._refresh_projector(name, ops.get_encoder(enc_id), proj)
            ops.encode_project_merge(hidden, ids[name], seq_table, enc_id, False)

        names = [n for n in ("dna_rna", "protein") if n in meta["plans"]]                   # reference order, :120-134
        side_ws = meta.get("side_workspace")
        small = max(ids[n].shape[0] * ids[n].shape[1] for n in names) <= self.concurrent_max_rows if names else False
        if (self.concurrent and small and len(names) == 2
                and (side_ws is not None or not ops.workspace_is_pinned(hidden.device))):
            dev = hidden.device                        # two branches (also inside a CUDA-graph capture): disjoint rows
            cur = torch.cuda.current_stream(dev)
            side = self._side_streams.setdefault(dev.index, torch.cuda.Stream(dev))
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                if side_ws is not None:
                    with ops.use_workspace(side_ws):
                        one(names[1])
                else:
                    one(names[1])
            one(names[0])
            cur.wait_stream(side)
        else:
            for name in names:
                one(name)
        return hidden

    # ------------------------------------------------------------------ one CUDA graph per bucket
    def graphed(self, embed_weight: torch.Tensor, batch_size: int, seq_len: int, omic_types: List[List[str]], k_tokens: int,
                pad_token_ids) -> "GraphedOmicsCall":
        """Capture ``embed_and_process`` for one fixed bucket -- (B, T), the per-sample modality layout ``omic_types`` and K --
        into a CUDA graph (``generate`` runs the path once per batch right before a long decode, omics_one.py:187-233: at
        small B the ~200-450 launches, not the math, set the latency).  The returned callable copies the new ``input_ids``
        and ``omic_ids`` into its static buffers, replays, and returns its static ``inputs_embeds`` buffer."""
        return GraphedOmicsCall(self, embed_weight, batch_size, seq_len, omic_types, k_tokens, tuple(pad_token_ids))

    @staticmethod
    def _module_version(module) -> tuple:
        """Changes when a parameter is updated through autograd-visible in-place ops (``load_state_dict``, ``p.add_``) or
        re-bound.  It does NOT see ``p.data.copy_`` or writes through a flat buffer the parameters are views of -- which is
        how DeepSpeed's ZeRO optimizers update bf16 parameters -- so it is trusted for FROZEN modules only."""
        v = p_sum = 0
        for prm in module.parameters():
            v += prm._version
            p_sum ^= prm.data_ptr()
        return (v, p_sum)

    @staticmethod
    def _trainable(module) -> bool:
        return any(prm.requires_grad for prm in module.parameters())

    def mark_weights_dirty(self) -> None:
        """Force a re-pack of every live module on the next call (e.g. after writing frozen weights through ``.data``)."""
        self._enc_versions = {}
        self._proj_versions = {}

    def refresh_encoders(self) -> List[str]:
        """Re-pack (in place) the encoders whose live ``nn.Module`` weights may have changed since they were packed.
        A module with TRAINABLE parameters (``--train-bio``, src/utils/tools.py:313-331) is re-packed on every call: the
        reference trains under DeepSpeed ZeRO-2 in bf16 (scripts/train/*.sh, src/configs/ds_z2_config.json), whose
        optimizer step writes the parameters through ``.data`` / flat-buffer views and leaves no trace in the version
        counters.  A frozen module is re-packed when its version counters or storage moved (checkpoint loaded late), or
        after ``mark_weights_dirty()``.  Called at the top of every ``process_omic_sequences`` with live modules."""
        done = []
        for name, module in self._enc_modules.items():
            if name not in self._ids:
                continue
            ver = self._module_version(module)
            if self._trainable(module) or self._enc_versions.get(name) != ver:
                ops.get_encoder(self._ids[name]).reload(module.state_dict())
                self._enc_versions[name] = ver
                done.append(name)
        return done

    def _refresh_projector(self, name: str, enc: PackedEncoder, proj) -> None:
        """Same rule for the projector (``--train-mlp``): trainable -> re-packed every call (4.7 M elements at Molly-1.7B)."""
        ver = (proj.weight._version, proj.bias._version, proj.weight.data_ptr())
        if proj.weight.requires_grad or proj.bias.requires_grad or self._proj_versions.get(name) != ver:
            enc.load_projector(proj.weight, proj.bias)
            self._proj_versions[name] = ver

    # ------------------------------------------------------------------ projector checkpoints
    PROJECTOR_FILES = {"dna_rna": "dna_rna_projector.bin", "protein": "protein_projector.bin"}

    def save_projectors(self, output_dir: str) -> List[str]:
        """Write ``dna_rna_projector.bin`` / ``protein_projector.bin`` exactly as ``OmicsTrainer.save_model`` does
        (src/trainer/omics_trainer.py:92-103): ``torch.save`` of the ``nn.Linear`` state dict {"weight" [D, h], "bias" [D]}.
        Live ``nn.Linear`` modules (``from_omics_one``) are the source when present, else the packed device copies."""
        import os
        os.makedirs(output_dir, exist_ok=True)
        written = []
        for name, fname in self.PROJECTOR_FILES.items():
            if name not in self._ids:
                continue
            proj = self._proj_modules.get(name)
            if proj is not None:
                sd = {k: v.detach().cpu() for k, v in proj.state_dict().items()}
            else:
                enc = ops.get_encoder(self._ids[name])
                sd = {"weight": enc.proj_w.detach().cpu().clone(), "bias": enc.proj_b.detach().cpu().clone()}
            torch.save(sd, os.path.join(output_dir, fname))
            written.append(os.path.join(output_dir, fname))
        return written

    def load_projectors(self, trained_model_path: str) -> List[str]:
        """Load the two ``*.bin`` projector state dicts if present (src/inference_lora.py:218-234: missing files are
        skipped silently) into the live modules and the packed kernel buffers."""
        import os
        loaded = []
        for name, fname in self.PROJECTOR_FILES.items():
            fpath = os.path.join(trained_model_path, fname)
            if name not in self._ids or not os.path.exists(fpath):
                continue
            sd = torch.load(fpath, map_location="cpu")
            enc = ops.get_encoder(self._ids[name])
            if tuple(sd["weight"].shape) != tuple(enc.proj_w.shape) or tuple(sd["bias"].shape) != tuple(enc.proj_b.shape):
                raise RuntimeError(f"size mismatch for {fname}: weight {tuple(sd['weight'].shape)} vs "
                                   f"{tuple(enc.proj_w.shape)}")                 # what load_state_dict raises
            proj = self._proj_modules.get(name)
            if proj is not None:
                proj.load_state_dict(sd)
                self._refresh_projector(name, enc, proj)
            else:
                enc.load_projector(sd["weight"].to(enc.proj_w.device), sd["bias"].to(enc.proj_b.device))
            loaded.append(fpath)
        return loaded

    # ------------------------------------------------------------------ other consumers of the encoder
    def encode(self, name: str, ids: torch.Tensor) -> torch.Tensor:
        """``hidden_states[-1]`` of the named encoder for ``ids`` [n, K] -> bf16 [n, K, h]."""
        return ops.encode(ids.to(torch.int64).contiguous(), self._ids[name])

    def pooled(self, name: str, ids: torch.Tensor, mode: str = "mean") -> torch.Tensor:
        """Masked mean-pool (embed_text.py:112-129) or CLS read-out (baselines/model.py:104-120), fp32 [n, h]."""
        ids = ids.to(torch.int64).contiguous()
        return ops.pool(self.encode(name, ids), ids, 0 if mode == "mean" else 1)


class GraphedOmicsCall:
    """One captured bucket of ``FastOmicsPath.embed_and_process`` (see ``FastOmicsPath.graphed``)."""

    def __init__(self, path: FastOmicsPath, embed_weight: torch.Tensor, batch_size: int, seq_len: int,
                 omic_types: List[List[str]], k_tokens: int, pad_token_ids):
        dev = embed_weight.device
        self.path, self.embed_weight, self.pad_token_ids = path, embed_weight, pad_token_ids
        infos = [[{"type": t, "start": -1} for t in row] for row in omic_types]
        n_max = max(1, max(len(r) for r in omic_types))
        for row in infos:                                                       # collate-style padding (omics_dataset.py:480-492)
            row.extend({"type": "pad", "start": -1} for _ in range(n_max - len(row)))
        self.input_ids = torch.zeros(batch_size, seq_len, dtype=torch.int64, device=dev)
        self.omic_ids = torch.ones(batch_size, n_max, k_tokens, dtype=torch.int64, device=dev)
        self.meta = path._fused_meta(batch_size, self.omic_ids, infos, dev)
        self.gather = {name: (idx[0].long(), idx[2].long()) for name, idx in self.meta["idx"].items()}
        need = 0
        for name, plan in self.meta["plans"].items():
            need = max(need, ops.get_encoder(path._ids[name]).workspace_bytes(len(plan), k_tokens))
        self.workspace = torch.empty(need + 8192, dtype=torch.uint8, device=dev)
        if (path.concurrent and len(self.meta["plans"]) == 2
                and max(len(p) for p in self.meta["plans"].values()) * k_tokens <= path.concurrent_max_rows):   # 2nd branch
            self.meta["side_workspace"] = torch.empty(need + 8192, dtype=torch.uint8, device=dev)
        with ops.use_workspace(self.workspace):
            side = torch.cuda.Stream(dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):                                       # warm-up: lazy attribute setting, plan caches
                self._run()
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize(dev)
            ops.error_flag(dev).zero_()              # the all-zero warm-up input has no placeholder runs: ERRBIT_LAYOUT
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.out = self._run()

    @torch.no_grad()
    def _run(self) -> torch.Tensor:
        ids = {name: self.omic_ids[bi, si].contiguous() for name, (bi, si) in self.gather.items()}
        return self.path._fused_run(self.input_ids, self.embed_weight, ids, self.meta, self.pad_token_ids)

    def __call__(self, input_ids: torch.Tensor, omic_ids: torch.Tensor) -> torch.Tensor:
        """``input_ids`` [B, T] and collated ``omic_ids`` [B, Nmax, K] of this bucket (host or device) -> inputs_embeds
        (the call's static output buffer: consume or clone it before the next call)."""
        self.input_ids.copy_(input_ids, non_blocking=True)
        self.omic_ids.copy_(omic_ids, non_blocking=True)
        self.graph.replay()
        if self.path.strict:
            ops.check_device_errors(self.out.device, "GraphedOmicsCall")
        return self.out
