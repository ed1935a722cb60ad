"""GPU parity of the whole path through the reference-facing boundary ``process_omic_sequences``:
CUDA path (bf16 tensor-core arithmetic, fp32 residual stream) vs the fp32 oracle / the reference's committed outputs.

Parity bar (BASELINE.json north_star):
  * written-row index set and merged layout: bit-exact; rows outside the placeholder runs: bit-identical to the input
  * values: max|cand - ref_fp32| / max|ref_fp32| <= 2e-2 over the WHOLE merged [B,T,D] tensor (pad rows included)
"""
import os

import numpy as np
import pytest
import torch

from oracle import cases, synth
from oracle.esm_oracle import esm_encoder_forward, process_omic_sequences as oracle_process
from tests.util import assert_close, assert_parity, rel_max_err

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL = 2e-2
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def build_path(case, strict=True):
    from molly_b200.config import EncoderConfig
    from molly_b200.omics_path import FastOmicsPath
    return FastOmicsPath.from_state_dicts(
        dna_rna_cfg=EncoderConfig.from_mapping(case.nt.spec.as_dict()), dna_rna_state=case.nt.weights,
        dna_rna_projector=case.nt.projector, dna_rna_project_token_num=case.nt.project_token_num,
        protein_cfg=EncoderConfig.from_mapping(case.pr.spec.as_dict()), protein_state=case.pr.weights,
        protein_projector=case.pr.projector, protein_project_token_num=case.pr.project_token_num,
        device=DEV, strict=strict)


def oracle_out(case):
    hs = case.batch.hidden_states.clone()
    with torch.no_grad():
        return oracle_process(hs, case.batch.omic_ids, case.batch.omic_info_list, case.nt, case.pr)


def check_merged(name, case, got, ref, in_dtype):
    got = got.float().cpu()
    inp = case.batch.hidden_states.to(in_dtype).float()
    exp = synth.expected_rows(case.batch.omic_info_list, case.K, case.nt.project_token_num, case.pr.project_token_num)
    written = torch.zeros(got.shape[:2], dtype=torch.bool)
    for (b, t) in exp:
        written[b, t] = True
    assert torch.equal(got[~written], inp[~written]), f"{name}: rows outside the placeholder runs changed"
    changed = (got != inp).any(-1)
    assert torch.equal(changed, written), f"{name}: written-row index set differs from the reference's"
    assert_close(name, got, ref if in_dtype == torch.float32 else ref, TOL)
    assert_parity(name + " (written rows only)", got[written], ref[written], TOL)


@pytest.mark.parametrize("spec_name", ["tiny_esm2", "tiny_ntv2", "tiny_ntv1", "esm2_t6_8m", "nt_v2_50m"])
def test_encoder_forward_vs_oracle(spec_name):
    from molly_b200.config import EncoderConfig
    from molly_b200.packing import PackedEncoder
    from molly_b200 import ops
    from oracle.esm_oracle import SPECS, init_encoder_weights, init_projector
    spec = SPECS[spec_name]
    W = init_encoder_weights(spec, 50)
    g = torch.Generator().manual_seed(51)
    k = 150
    rows = []
    for valid in (150, 129, 40, 7):
        rows.append(synth.protein_ids(g, k, valid) if spec.vocab_size == 33 else synth.nucleotide_ids(g, k, valid, spec.vocab_size))
    ids = torch.stack(rows)
    with torch.no_grad():
        ref = esm_encoder_forward(spec, W, ids)
    enc = PackedEncoder(EncoderConfig.from_mapping(spec.as_dict()), W, init_projector(spec.hidden_size, 64, 52), k,
                        torch.device(DEV))
    eid = ops.register_encoder(enc)
    try:
        got = ops.encode(ids.to(DEV), eid)
        ops.check_device_errors(torch.device(DEV))
        assert_close(f"encoder {spec_name}", got, ref, TOL)
    finally:
        ops.unregister_encoder(eid)


@pytest.mark.parametrize("gate_first,ffn_bias", [(True, True), (False, False), (False, True)])
def test_nt_v2_variant_switches_vs_oracle(gate_first, ffn_bias):
    """NT-v2's gated FFN is hub remote code that cannot be read offline (parity unpinned).  The two places where a
    variant of it would change results -- which half of ``intermediate.dense`` goes through SiLU, and whether the FFN
    Linears carry biases (``add_bias_fnn``) -- are switches; each setting must match the oracle with the same setting,
    through the forward AND through the --train-bio backward."""
    import dataclasses
    from molly_b200.config import EncoderConfig
    from molly_b200.packing import PackedEncoder
    from molly_b200 import ops, train
    from oracle.esm_oracle import SPECS, init_encoder_weights, init_projector
    spec = dataclasses.replace(SPECS["tiny_ntv2"], glu_gate_first=gate_first)
    W = init_encoder_weights(spec, 53)
    if ffn_bias:
        g = torch.Generator().manual_seed(54)
        for i in range(spec.num_hidden_layers):
            p = f"esm.encoder.layer.{i}."
            W[p + "intermediate.dense.bias"] = torch.randn(2 * spec.intermediate_size, generator=g) * 0.05
            W[p + "output.dense.bias"] = torch.randn(spec.hidden_size, generator=g) * 0.05
    g = torch.Generator().manual_seed(55)
    k = 150
    ids = torch.stack([synth.nucleotide_ids(g, k, valid, spec.vocab_size) for valid in (150, 129, 40)])
    Wg = {n: v.clone().requires_grad_(v.dtype.is_floating_point) for n, v in W.items()}
    ref = esm_encoder_forward(spec, Wg, ids)
    d_out = torch.randn(ref.shape, generator=g).to(torch.bfloat16).float()
    (ref * d_out).sum().backward()
    cfg = EncoderConfig.from_mapping(spec.as_dict())
    assert cfg.glu_gate_first == gate_first
    enc = PackedEncoder(cfg, W, init_projector(spec.hidden_size, 64, 56), k, torch.device(DEV))
    eid = ops.register_encoder(enc)
    try:
        got = ops.encode(ids.to(DEV), eid)
        ops.check_device_errors(torch.device(DEV))
        assert_close(f"NT-v2 variant gate_first={gate_first} ffn_bias={ffn_bias}", got, ref.detach(), TOL)
        out, tape = train.encoder_forward_train(enc, ids.to(DEV))
        grads = train.encoder_backward(enc, tape, d_out.reshape(-1, spec.hidden_size).to(DEV).to(torch.bfloat16))
        for key in ("esm.encoder.layer.0.intermediate.dense.weight", "esm.encoder.layer.1.output.dense.weight") + (
                ("esm.encoder.layer.0.intermediate.dense.bias", "esm.encoder.layer.1.output.dense.bias") if ffn_bias else ()):
            assert_close(f"d {key} gate_first={gate_first}", grads[key].float().cpu(), Wg[key].grad, 3e-2)
        assert ("esm.encoder.layer.0.intermediate.dense.bias" in grads) == ffn_bias
    finally:
        ops.unregister_encoder(eid)


@pytest.mark.parametrize("name", list(cases.golden_cases().keys()))
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_path_vs_reference_golden(name, dtype):
    """CUDA path against the REFERENCE's own committed output (tests/golden) -- and against the live oracle."""
    case = cases.golden_cases()[name]
    z = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    assert cases.case_digest(case) == bytes(z["digest"]).decode()
    ref = torch.from_numpy(z["merged"])
    path = build_path(case)
    try:
        hs = case.batch.hidden_states.to(dtype).to(DEV)
        out = path.process_omic_sequences(hs, case.batch.omic_ids, case.batch.omic_info_list, hs.device)
        assert out is hs                                           # in place + identity (omics_one.py:97,136)
        if dtype == torch.bfloat16:                                # reference run on the bf16-rounded input rows
            ref = ref.clone()
            exp = synth.expected_rows(case.batch.omic_info_list, case.K, case.nt.project_token_num,
                                      case.pr.project_token_num)
            w = torch.zeros(ref.shape[:2], dtype=torch.bool)
            for (b, t) in exp:
                w[b, t] = True
            ref[~w] = case.batch.hidden_states.to(dtype).float()[~w]
        check_merged(f"path {name} {dtype}", case, out, ref, dtype)
        assert rel_max_err(oracle_out(case), torch.from_numpy(z["merged"])) < 1e-4
    finally:
        path.close()


@pytest.mark.parametrize("varlen", [False, True])
def test_path_molly_mini_full_size(varlen):
    """BASELINE.json configs[0]: Molly-mini, B=4, 512 omics tokens per sequence, both modalities, full tensor."""
    case = cases.molly_mini(varlen=varlen)
    ref = oracle_out(case)
    z = np.load(os.path.join(GOLDEN, f"{case.name}.npz"))
    path = build_path(case)
    try:
        hs = case.batch.hidden_states.to(DEV)
        out = path.process_omic_sequences(hs, case.batch.omic_ids.to(DEV), case.batch.omic_info_list, hs.device)
        check_merged(f"path {case.name}", case, out, ref, torch.float32)
        sel = torch.from_numpy(z["sel"]).long()
        assert_close(f"path {case.name} vs reference fixture rows", out.cpu()[sel[:, 0], sel[:, 1]],
                     torch.from_numpy(z["vals"]), TOL)
    finally:
        path.close()


def test_ids_on_device_equal_ids_on_host_and_list_form():
    case = cases.golden_cases()["tiny_rotary_glu"]
    path = build_path(case)
    try:
        outs = []
        for form in ("cpu_tensor", "cuda_tensor", "list_of_lists"):
            hs = case.batch.hidden_states.to(DEV)
            ids = case.batch.omic_ids
            if form == "cuda_tensor":
                ids = ids.to(DEV)
            elif form == "list_of_lists":
                ids = [[row for row in sample] for sample in ids]
            outs.append(path.process_omic_sequences(hs, ids, case.batch.omic_info_list, hs.device).cpu())
        assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])        # deterministic kernels
    finally:
        path.close()


def test_error_conventions_match_reference():
    case = cases.golden_cases()["tiny_rotary_glu"]
    path = build_path(case, strict=True)
    try:
        hs = case.batch.hidden_states.to(DEV)
        infos = [[dict(i) for i in row] for row in case.batch.omic_info_list]
        infos[0][0]["type"] = "lipid"
        with pytest.raises(ValueError, match="Unsupported omic type"):
            path.process_omic_sequences(hs, case.batch.omic_ids, infos, hs.device)
        bad = case.batch.omic_ids.clone()
        bad[0, 1, 3] = 999
        with pytest.raises(AssertionError):                      # host ids: checked before launch
            path.process_omic_sequences(hs, bad, case.batch.omic_info_list, hs.device)
        with pytest.raises(AssertionError):                      # device ids: error flag, strict mode syncs
            path.process_omic_sequences(hs, bad.to(DEV), case.batch.omic_info_list, hs.device)
        infos = [[dict(i) for i in row] for row in case.batch.omic_info_list]
        infos[1][0]["start"] = case.T - 5
        with pytest.raises(RuntimeError):
            path.process_omic_sequences(hs, case.batch.omic_ids, infos, hs.device)
        # empty modality lists are a no-op (omics_one.py:67-68); 'pad' / start == -1 silently skipped
        hs2 = case.batch.hidden_states.to(DEV)
        none = [[{"type": "pad", "start": -1} for _ in row] for row in case.batch.omic_info_list]
        out = path.process_omic_sequences(hs2, case.batch.omic_ids, none, hs2.device)
        assert out is hs2 and torch.equal(out.cpu(), case.batch.hidden_states)
        with pytest.raises(RuntimeError):
            path.process_omic_sequences(case.batch.hidden_states.clone(), case.batch.omic_ids,
                                        case.batch.omic_info_list, "cpu")          # no CPU fallback
    finally:
        path.close()


def test_sample_sharding_equals_single_rank():
    """each rank encodes/projects/merges its own samples; concatenated shards == the single-rank result."""
    from molly_b200 import planner
    case = cases.golden_cases()["tiny_rotary_glu"]
    path = build_path(case)
    try:
        hs = case.batch.hidden_states.to(DEV)
        full = path.process_omic_sequences(hs, case.batch.omic_ids, case.batch.omic_info_list, hs.device).cpu()
        for world in (2, 4):
            parts = []
            for rank in range(world):
                r = planner.shard_samples(case.batch.hidden_states.shape[0], world, rank)
                if len(r) == 0:
                    continue
                sl = slice(r.start, r.stop)
                h = case.batch.hidden_states[sl].to(DEV)
                parts.append(path.process_omic_sequences(h, case.batch.omic_ids[sl], case.batch.omic_info_list[sl],
                                                         h.device).cpu())
            assert torch.equal(torch.cat(parts), full)
    finally:
        path.close()


def test_install_on_omics_one_like_object_and_backward():
    """install() swaps the method on an OmicsOne-shaped object; projector grads match autograd of the oracle."""
    import types
    from molly_b200.omics_path import FastOmicsPath
    from molly_b200.config import EncoderConfig
    from molly_b200.packing import PackedEncoder
    case = cases.golden_cases()["tiny_rotary_glu"]
    D = case.D
    om = types.SimpleNamespace()
    om.dna_rna_projector = torch.nn.Linear(case.nt.spec.hidden_size, D).to(DEV)
    om.protein_projector = torch.nn.Linear(case.pr.spec.hidden_size, D).to(DEV)
    om.dna_rna_projector.load_state_dict(case.nt.projector)
    om.protein_projector.load_state_dict(case.pr.projector)
    dev = torch.device(DEV)
    path = FastOmicsPath(
        PackedEncoder(EncoderConfig.from_mapping(case.nt.spec.as_dict()), case.nt.weights, case.nt.projector, case.K, dev),
        PackedEncoder(EncoderConfig.from_mapping(case.pr.spec.as_dict()), case.pr.weights, case.pr.projector, case.K, dev))
    path._proj_modules = {"dna_rna": om.dna_rna_projector, "protein": om.protein_projector}
    path.install(om)
    try:
        emb = case.batch.hidden_states.to(DEV).requires_grad_(True)
        hs = emb * 1.0                                             # non-leaf, like embed_tokens(input_ids)
        out = om.process_omic_sequences(hs, case.batch.omic_ids, case.batch.omic_info_list, hs.device)
        assert out is hs
        gw = torch.randn(out.shape, generator=torch.Generator().manual_seed(60)).to(DEV)
        (out * gw).sum().backward()
        # oracle autograd in fp32
        import copy
        nt, pr = copy.copy(case.nt), copy.copy(case.pr)
        nt.projector = {k: v.clone().requires_grad_(True) for k, v in case.nt.projector.items()}
        pr.projector = {k: v.clone().requires_grad_(True) for k, v in case.pr.projector.items()}
        emb_ref = case.batch.hidden_states.clone().requires_grad_(True)
        o = oracle_process(emb_ref * 1.0, case.batch.omic_ids, case.batch.omic_info_list, nt, pr)
        (o * gw.cpu()).sum().backward()
        assert_close("d protein_projector.weight", om.protein_projector.weight.grad, pr.projector["weight"].grad, TOL)
        assert_close("d protein_projector.bias", om.protein_projector.bias.grad, pr.projector["bias"].grad, TOL)
        assert_close("d dna_rna_projector.weight", om.dna_rna_projector.weight.grad, nt.projector["weight"].grad, TOL)
        assert_close("d dna_rna_projector.bias", om.dna_rna_projector.bias.grad, nt.projector["bias"].grad, TOL)
        # grad wrt the incoming embeddings: exactly zero on overwritten rows, pass-through elsewhere
        assert torch.equal(emb.grad.cpu(), emb_ref.grad)
    finally:
        path.close()


def test_pooled_heads():
    """Other in-tree consumers of the encoder forward: masked mean (embed_text.py:112-129), CLS (baselines/model.py:104-120)."""
    case = cases.golden_cases()["tiny_rotary_glu"]
    path = build_path(case)
    try:
        g = torch.Generator().manual_seed(61)
        ids = torch.stack([synth.protein_ids(g, 60, v) for v in (60, 31, 5)])
        with torch.no_grad():
            last = esm_encoder_forward(case.pr.spec, case.pr.weights, ids)
        m = (ids != 1).float().unsqueeze(-1)
        mean_ref = (last * m).sum(1) / m.sum(1).clamp(min=1e-9)
        assert_close("masked mean pool", path.pooled("protein", ids.to(DEV), "mean"), mean_ref, TOL)
        assert_close("cls", path.pooled("protein", ids.to(DEV), "cls"), last[:, 0], TOL)
    finally:
        path.close()


def test_projector_checkpoint_round_trip(tmp_path):
    """``dna_rna_projector.bin`` / ``protein_projector.bin`` in the reference's format (omics_trainer.py:92-103 writes them,
    inference_lora.py:218-234 reads them): an ``nn.Linear`` state dict that loads into a fresh path and gives the same rows."""
    case = cases.golden_cases()["tiny_rotary_glu"]
    a = build_path(case)
    b = build_path(case)
    try:
        files = a.save_projectors(str(tmp_path))
        assert sorted(os.path.basename(f) for f in files) == ["dna_rna_projector.bin", "protein_projector.bin"]
        sd = torch.load(files[0], map_location="cpu")
        lin = torch.nn.Linear(sd["weight"].shape[1], sd["weight"].shape[0])
        lin.load_state_dict(sd)                                          # the reference's load path accepts it
        for enc in (b.dna_rna, b.protein):                               # scramble b's projectors, then restore from disk
            enc.load_projector(torch.zeros_like(enc.proj_w), torch.ones_like(enc.proj_b))
        hs = case.batch.hidden_states.to(DEV)
        want = a.process_omic_sequences(hs.clone(), case.batch.omic_ids, case.batch.omic_info_list, hs.device)
        off = b.process_omic_sequences(hs.clone(), case.batch.omic_ids, case.batch.omic_info_list, hs.device)
        assert not torch.equal(off, want)
        assert len(b.load_projectors(str(tmp_path))) == 2
        got = b.process_omic_sequences(hs.clone(), case.batch.omic_ids, case.batch.omic_info_list, hs.device)
        assert torch.equal(got, want)
        assert b.load_projectors(str(tmp_path / "nowhere")) == []        # missing files are skipped like the reference
    finally:
        a.close()
        b.close()


def test_overlapping_ranges_keep_reference_order():
    """Not produced by the dataset, but legal at the boundary: a protein range that overlaps a DNA range in the same sample.
    The reference writes DNA/RNA first and protein second (omics_one.py:120-134), so the protein rows win; the two-stream
    schedule must notice the overlap and fall back to that order."""
    case = cases.golden_cases()["tiny_rotary_glu"]
    infos = [[dict(i) for i in row] for row in case.batch.omic_info_list]
    dna = next(i for i in infos[0] if i["type"] == "dna")
    pro = next(i for i in infos[0] if i["type"] == "protein")
    pro["start"] = dna["start"] + 7                                  # rows [s+8, s+48) overlap the DNA rows [s+1, s+41)
    ref = case.batch.hidden_states.clone()
    with torch.no_grad():
        oracle_process(ref, case.batch.omic_ids, infos, case.nt, case.pr)
    path = build_path(case)
    try:
        assert path.concurrent and case.K * 4 <= path.concurrent_max_rows
        hs = case.batch.hidden_states.to(DEV)
        for _ in range(3):                                           # a race would not lose every time
            out = path.process_omic_sequences(hs.clone(), case.batch.omic_ids, infos, hs.device)
            assert_close("overlapping ranges", out.float().cpu(), ref, TOL)
            lo = pro["start"] + 1
            assert_close("protein rows win", out[0, lo:lo + case.K].float().cpu(), ref[0, lo:lo + case.K], TOL)
    finally:
        path.close()


def test_from_omics_one_follows_live_encoder_weights():
    """SURVEY 8b ownership: the packed weights are caches of the three ``nn.Module``s.  Built from an ``OmicsOne``-shaped
    object holding stock HF ``EsmForMaskedLM`` encoders, the path must (1) reproduce the oracle and (2) follow in-place weight
    updates (``--train-bio`` optimizer steps, late ``load_state_dict``) without being rebuilt."""
    import types
    from molly_b200.omics_path import FastOmicsPath
    from oracle.ref_import import build_hf_encoder
    case = cases.golden_cases()["tiny_absolute_leftpad"]             # tiny_ntv1 (stock ESM classes) + tiny_esm2
    om = types.SimpleNamespace()
    om.dna_rna_model = build_hf_encoder(case.nt.spec, case.nt.weights)
    om.protein_model = build_hf_encoder(case.pr.spec, case.pr.weights)
    om.dna_rna_projector = torch.nn.Linear(case.nt.spec.hidden_size, case.D)
    om.protein_projector = torch.nn.Linear(case.pr.spec.hidden_size, case.D)
    om.dna_rna_projector.load_state_dict(case.nt.projector)
    om.protein_projector.load_state_dict(case.pr.projector)
    om.dna_rna_project_token_num = case.nt.project_token_num
    om.protein_project_token_num = case.pr.project_token_num
    path = FastOmicsPath.from_omics_one(om, DEV, strict=True)
    try:
        hs = case.batch.hidden_states.to(DEV)
        with torch.no_grad():
            out = path.process_omic_sequences(hs.clone(), case.batch.omic_ids, case.batch.omic_info_list, hs.device)
        check_merged("from_omics_one", case, out, oracle_out(case), torch.float32)
        assert sorted(path.refresh_encoders()) == ["dna_rna", "protein"]    # trainable modules are re-packed on every call
        for m in (om.dna_rna_model, om.protein_model):
            for prm in m.parameters():
                prm.requires_grad_(False)
        assert path.refresh_encoders() == []                                # frozen: only when the version counters moved
        # an "optimizer step" on the protein encoder + a checkpoint load into the DNA/RNA encoder
        import copy
        new_pr = {k: (v * 1.05 if v.dtype.is_floating_point and v.dim() == 2 else v) for k, v in case.pr.weights.items()}
        new_nt = {k: (v * 0.97 if v.dtype.is_floating_point and v.dim() == 2 else v) for k, v in case.nt.weights.items()}
        with torch.no_grad():
            for name, prm in om.protein_model.named_parameters():
                if name in new_pr:
                    prm.copy_(new_pr[name])
        om.dna_rna_model.load_state_dict(new_nt, strict=False)
        pr2, nt2 = copy.copy(case.pr), copy.copy(case.nt)
        pr2.weights, nt2.weights = new_pr, new_nt
        ref2 = case.batch.hidden_states.clone()
        with torch.no_grad():
            oracle_process(ref2, case.batch.omic_ids, case.batch.omic_info_list, nt2, pr2)
            out2 = path.process_omic_sequences(hs.clone(), case.batch.omic_ids, case.batch.omic_info_list, hs.device)
        assert not torch.equal(out2, out)
        assert_close("after in-place weight updates", out2.float().cpu(), ref2, TOL)
        # --train-bio: trainable encoder parameters receive gradients through the path (tests/test_gpu_train.py checks every one)
        for prm in om.protein_model.parameters():
            prm.requires_grad_(True)
        hs2 = hs.clone().requires_grad_(True)
        out3 = path.process_omic_sequences(hs2 * 1.0, case.batch.omic_ids, case.batch.omic_info_list, hs.device)
        gw = torch.randn(out3.shape, generator=torch.Generator().manual_seed(77)).to(DEV)
        (out3 * gw).sum().backward()
        pr3 = copy.copy(pr2)
        pr3.weights = {k: v.clone().requires_grad_(v.dtype.is_floating_point) for k, v in new_pr.items()}
        pr3.projector = {k: v.clone().requires_grad_(True) for k, v in case.pr.projector.items()}
        emb_ref = case.batch.hidden_states.clone().requires_grad_(True)
        o3 = oracle_process(emb_ref * 1.0, case.batch.omic_ids, case.batch.omic_info_list, nt2, pr3)
        (o3 * gw.cpu()).sum().backward()
        assert all(prm.grad is None for prm in om.dna_rna_model.parameters())          # frozen encoder: untouched
        got_g = dict(om.protein_model.named_parameters())
        for key in ("esm.encoder.layer.0.attention.self.value.weight", "esm.encoder.layer.1.output.dense.weight",
                    "esm.encoder.emb_layer_norm_after.weight", "esm.embeddings.word_embeddings.weight"):
            assert_close(f"train-bio d {key}", got_g[key].grad.float().cpu(), pr3.weights[key].grad, 3e-2)
        assert_close("train-bio d protein_projector.weight", om.protein_projector.weight.grad.float().cpu(),
                     pr3.projector["weight"].grad, TOL)
        assert torch.equal(hs2.grad.cpu(), emb_ref.grad)                                 # zero on overwritten rows
    finally:
        path.close()


def _omics_one_like(case):
    import types
    from oracle.ref_import import build_hf_encoder
    om = types.SimpleNamespace()
    om.dna_rna_model = build_hf_encoder(case.nt.spec, case.nt.weights)
    om.protein_model = build_hf_encoder(case.pr.spec, case.pr.weights)
    om.dna_rna_projector = torch.nn.Linear(case.nt.spec.hidden_size, case.D)
    om.protein_projector = torch.nn.Linear(case.pr.spec.hidden_size, case.D)
    om.dna_rna_projector.load_state_dict(case.nt.projector)
    om.protein_projector.load_state_dict(case.pr.projector)
    om.dna_rna_project_token_num = case.nt.project_token_num
    om.protein_project_token_num = case.pr.project_token_num
    return om


def test_trainable_weights_follow_data_copy_and_flat_buffer_views():
    """DeepSpeed's ZeRO optimizers (the reference trains with ds_z2_config.json, bf16) update parameters with
    ``p.data.copy_`` or by writing a flat buffer the parameters are views of.  Neither bumps ``p._version`` nor moves
    ``data_ptr``; the packed copies of TRAINABLE modules must follow anyway."""
    import copy
    from molly_b200.omics_path import FastOmicsPath
    case = cases.golden_cases()["tiny_absolute_leftpad"]
    om = _omics_one_like(case)
    for prm in om.dna_rna_model.parameters():
        prm.requires_grad_(False)                                    # --train-bio protein only + --train-mlp
    path = FastOmicsPath.from_omics_one(om, DEV, strict=True)
    try:
        hs = case.batch.hidden_states.to(DEV)
        with torch.no_grad():
            path.process_omic_sequences(hs.clone(), case.batch.omic_ids, case.batch.omic_info_list, hs.device)
        # (1) p.data.copy_ on encoder matrices and on the projector
        new_pr = {k: (v * 1.04 if v.dtype.is_floating_point and v.dim() == 2 else v) for k, v in case.pr.weights.items()}
        versions = {n: p._version for n, p in om.protein_model.named_parameters()}
        for name, prm in om.protein_model.named_parameters():
            if name in new_pr:
                prm.data.copy_(new_pr[name])
        new_proj = {"weight": case.pr.projector["weight"] * 0.9, "bias": case.pr.projector["bias"] + 0.01}
        om.protein_projector.weight.data.copy_(new_proj["weight"])
        om.protein_projector.bias.data.copy_(new_proj["bias"])
        assert all(p._version == versions[n] for n, p in om.protein_model.named_parameters())   # invisible to the counters
        pr2 = copy.copy(case.pr)
        pr2.weights, pr2.projector = new_pr, new_proj
        ref = case.batch.hidden_states.clone()
        with torch.no_grad():
            oracle_process(ref, case.batch.omic_ids, case.batch.omic_info_list, case.nt, pr2)
            out = path.process_omic_sequences(hs.clone(), case.batch.omic_ids, case.batch.omic_info_list, hs.device)
        assert_close("after p.data.copy_", out.float().cpu(), ref, TOL)
        # (2) parameters re-bound as views of one flat buffer, then the buffer is written
        prms = list(om.protein_projector.parameters())
        flat = torch.cat([p.data.reshape(-1) for p in prms]).clone()
        off = 0
        for p in prms:
            p.data = flat[off:off + p.numel()].view_as(p)
            off += p.numel()
        with torch.no_grad():
            path.process_omic_sequences(hs.clone(), case.batch.omic_ids, case.batch.omic_info_list, hs.device)
        flat.mul_(1.1)                                               # the "optimizer step"
        pr3 = copy.copy(pr2)
        pr3.projector = {"weight": new_proj["weight"] * 1.1, "bias": new_proj["bias"] * 1.1}
        ref3 = case.batch.hidden_states.clone()
        with torch.no_grad():
            oracle_process(ref3, case.batch.omic_ids, case.batch.omic_info_list, case.nt, pr3)
            out3 = path.process_omic_sequences(hs.clone(), case.batch.omic_ids, case.batch.omic_info_list, hs.device)
        assert_close("after a flat-buffer write", out3.float().cpu(), ref3, TOL)
        # (3) frozen weights written through .data need the explicit hook
        om.dna_rna_projector.weight.requires_grad_(False)
        om.dna_rna_projector.bias.requires_grad_(False)
        with torch.no_grad():
            path.process_omic_sequences(hs.clone(), case.batch.omic_ids, case.batch.omic_info_list, hs.device)
        om.dna_rna_projector.weight.data.mul_(0.5)
        path.mark_weights_dirty()
        nt4 = copy.copy(case.nt)
        nt4.projector = {"weight": case.nt.projector["weight"] * 0.5, "bias": case.nt.projector["bias"]}
        ref4 = case.batch.hidden_states.clone()
        with torch.no_grad():
            oracle_process(ref4, case.batch.omic_ids, case.batch.omic_info_list, nt4, pr3)
            out4 = path.process_omic_sequences(hs.clone(), case.batch.omic_ids, case.batch.omic_info_list, hs.device)
        assert_close("after mark_weights_dirty", out4.float().cpu(), ref4, TOL)
    finally:
        path.close()


def test_frozen_projector_still_zeroes_the_overwritten_rows_of_the_embedding_gradient():
    """--train-llm without --train-mlp (src/utils/tools.py:332-335): the projector is frozen, ``hidden_states`` requires
    grad.  The reference's slice-assign gives the overwritten rows zero gradient (omics_one.py:97)."""
    from molly_b200.omics_path import FastOmicsPath
    case = cases.golden_cases()["tiny_rotary_glu"]
    om = _omics_one_like(case)
    for m in (om.dna_rna_model, om.protein_model, om.dna_rna_projector, om.protein_projector):
        for prm in m.parameters():
            prm.requires_grad_(False)
    path = FastOmicsPath.from_omics_one(om, DEV, strict=True)
    try:
        emb = case.batch.hidden_states.to(DEV).requires_grad_(True)
        out = path.process_omic_sequences(emb * 1.0, case.batch.omic_ids, case.batch.omic_info_list, emb.device)
        gw = torch.randn(out.shape, generator=torch.Generator().manual_seed(61)).to(DEV)
        (out * gw).sum().backward()
        emb_ref = case.batch.hidden_states.clone().requires_grad_(True)
        o = oracle_process(emb_ref * 1.0, case.batch.omic_ids, case.batch.omic_info_list, case.nt, case.pr)
        (o * gw.cpu()).sum().backward()
        assert torch.equal(emb.grad.cpu(), emb_ref.grad)
        written = (emb_ref.grad == 0).all(dim=-1)
        assert int(written.sum()) > 0
    finally:
        path.close()


def test_train_bio_steps_through_cuda_graphs_match_the_eager_step_and_accumulate():
    """``--train-bio`` stepped repeatedly with one batch shape: from the third step on the encoder forward / backward are CUDA
    graph replays (``train._GraphedStep``).  The parameter gradients of a replayed step equal those of the first (eager) step,
    the step follows an in-place weight update in between, and gradients ACCUMULATE across steps (what autograd receives is
    never a view of a graph's buffer)."""
    from molly_b200 import train
    from molly_b200.omics_path import FastOmicsPath
    case = cases.golden_cases()["tiny_rotary_glu"]
    om = _omics_one_like(case)
    for mod in (om.dna_rna_model, om.protein_model, om.dna_rna_projector, om.protein_projector):
        mod.to(DEV)
    path = FastOmicsPath.from_omics_one(om, DEV, strict=True)
    params = [p for m in (om.dna_rna_model, om.protein_model, om.dna_rna_projector, om.protein_projector)
              for p in m.parameters() if p.requires_grad]
    gw = torch.randn(case.batch.hidden_states.shape, generator=torch.Generator().manual_seed(61)).to(DEV)

    def step(zero=True):
        if zero:
            for p in params:
                p.grad = None
        hs = case.batch.hidden_states.to(DEV) * 1.0
        out = path.process_omic_sequences(hs, case.batch.omic_ids, case.batch.omic_info_list, hs.device)
        (out * gw).sum().backward()
        return [None if p.grad is None else p.grad.float().clone() for p in params]

    try:
        first = step()
        assert sum(g is not None and float(g.abs().sum()) > 0 for g in first) > 20
        step()
        third = step()                                             # captured and replayed
        graphs = [g for name in ("dna_rna", "protein") for g in getattr(path, name).__dict__.get("_train_graphs", {}).values()]
        assert graphs and all(g.fwd is not None and g.bwd is not None for g in graphs), "the steady-state step was not graphed"
        for a, b in zip(third, first):
            if b is not None:
                assert_close("graphed step vs eager step", a.cpu(), b.cpu(), 1e-4)
        doubled = step(zero=False)                                 # accumulate on top of the third step's gradients
        for a, b in zip(doubled, first):
            if b is not None:
                assert_close("accumulated over two steps", a.cpu(), 2 * b.cpu(), 1e-3)
        with torch.no_grad():                                      # an optimizer step written through .data
            for p in om.protein_model.parameters():
                if p.dim() == 2:
                    p.data.mul_(1.05)
        moved = step()
        assert any(b is not None and float((a - b).abs().max()) > 1e-3 * float(b.abs().max()) for a, b in zip(moved, first)), \
            "the graphed step did not follow the weight update"
        os.environ["MOLLY_TRAIN_GRAPH"] = "0"
        try:
            eager_moved = step()
        finally:
            del os.environ["MOLLY_TRAIN_GRAPH"]
        for a, b in zip(moved, eager_moved):
            if b is not None:
                assert_close("graphed vs eager after the update", a.cpu(), b.cpu(), 1e-4)
    finally:
        path.close()
