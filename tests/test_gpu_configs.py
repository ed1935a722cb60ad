"""GPU parity at the FULL model shapes of BASELINE.json configs[2] and configs[3] (few sequences, so the fp32 CPU oracle
finishes in about a minute), plus size-independent properties at the full headline size configs[1].

configs[2] "Molly-4B": ESM-2 650M + NT-v2 500M -> D=2560, mixed dna/rna/protein, K=2048, lengths in [64, 2048]
configs[3] "Molly-8B": NT-v1 2.5B multispecies (h=2560, F=10240, head_dim 128, learned absolute positions) -> D=4096,
                        1000-bp DNA windows: K=256, 171 valid tokens
"""
import pytest
import torch

from oracle import cases, synth
from oracle.esm_oracle import process_omic_sequences as oracle_process
from tests.test_gpu_path import TOL, build_path, check_merged
from tests.util import assert_close

pytestmark = pytest.mark.gpu
DEV = "cuda"


def test_cfg3_molly4b_varlen_full_encoders():
    import psutil
    if psutil.virtual_memory().available < 24 * 2 ** 30:
        pytest.skip("fp32 CPU oracle of ESM-2 650M + NT-v2 500M needs ~12 GB of host memory")
    case = cases.build_case("cfg3_molly4b", "nt_v2_500m", "esm2_t33_650m", D=2560, K=2048, T=4400, seed=3000,
                            samples=[[("dna", 2048), ("protein", 701)], [("rna", 97)]], mask_tokens=False)
    ref = case.batch.hidden_states.clone()
    with torch.no_grad():
        oracle_process(ref, case.batch.omic_ids, case.batch.omic_info_list, case.nt, case.pr)
    path = build_path(case)
    try:
        hs = case.batch.hidden_states.to(DEV)
        out = path.process_omic_sequences(hs, case.batch.omic_ids, case.batch.omic_info_list, hs.device)
        check_merged("cfg3 Molly-4B varlen K=2048", case, out, ref, torch.float32)
    finally:
        path.close()


def test_cfg4_molly8b_nt_v1_2p5b():
    import psutil
    if psutil.virtual_memory().available < 48 * 2 ** 30:
        pytest.skip("fp32 CPU oracle of NT-2.5B needs ~25 GB of host memory")
    case = cases.build_case("cfg4_molly8b", "nt_v1_2p5b", "tiny_esm2", D=4096, K=256, T=600, seed=4000,
                            samples=[[("dna", 171), ("dna", 171)]], mask_tokens=False)
    ref = case.batch.hidden_states.clone()
    with torch.no_grad():
        oracle_process(ref, case.batch.omic_ids, case.batch.omic_info_list, case.nt, case.pr)
    path = build_path(case)
    try:
        hs = case.batch.hidden_states.to(DEV)
        out = path.process_omic_sequences(hs, case.batch.omic_ids, case.batch.omic_info_list, hs.device)
        check_merged("cfg4 Molly-8B NT-v1 2.5B K=256", case, out, ref, torch.float32)
    finally:
        path.close()


def test_cfg2_full_size_properties():
    """BASELINE configs[1] at FULL size (B=64 x (1 DNA + 1 protein) x 1024 tokens, ESM-2 650M + NT-v2 500M, D=2048, bf16):
    size-independent properties instead of an (hours-long) CPU oracle run:
      * written-row index set == the reference's; every other row bit-identical to the input
      * sample sharding: ranks' shards concatenated == the unsharded result, bit for bit
      * determinism: two runs are bit-identical
      * linearity of the projector+merge tail is implied by the per-kernel tests; here: no NaN/Inf, sane magnitude."""
    import bench
    wl = bench.WORKLOADS["molly_1p7b"]
    dev = torch.device(DEV)
    path = bench.build_path(wl, dev, strict=True)
    try:
        omic_ids, infos = bench.make_inputs(wl, seed=99)
        g = torch.Generator(device=DEV).manual_seed(5)
        base = (torch.randn(wl["B"], wl["T"], wl["D"], device=DEV, generator=g) * 0.02).to(torch.bfloat16)
        hs = base.clone()
        out = path.process_omic_sequences(hs, omic_ids, infos, dev)
        assert out is hs
        exp = synth.expected_rows(infos, wl["K"], wl["K"], wl["K"])
        written = torch.zeros(wl["B"], wl["T"], dtype=torch.bool)
        idx = torch.tensor(list(exp.keys()))
        written[idx[:, 0], idx[:, 1]] = True
        changed = (out != base).any(-1).cpu()
        assert torch.equal(changed, written)
        assert torch.isfinite(out.float()).all()
        rms = float(out[written.to(DEV)].float().pow(2).mean().sqrt())
        assert 0.05 < rms < 50.0, rms
        # determinism
        hs2 = base.clone()
        path.process_omic_sequences(hs2, omic_ids.to(DEV), infos, dev)
        assert torch.equal(hs2, out)
        # sample sharding: two "ranks" of 32 samples each
        parts = []
        for sl in (slice(0, 32), slice(32, 64)):
            h = base[sl].clone()
            parts.append(path.process_omic_sequences(h, omic_ids[sl], infos[sl], dev))
        assert torch.equal(torch.cat(parts), out)
    finally:
        path.close()


def _trained_like(weights, seed):
    """Make random-init weights look like a trained checkpoint where it matters for numerics: sharp attention (q/k x6),
    a few dominant residual channels (LayerNorm gains x25, as in ESM-2's outlier features), non-zero biases."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for k, v in weights.items():
        v = v.clone()
        if k.endswith(("query.weight", "key.weight")):
            v *= 6.0
        elif k.endswith("LayerNorm.weight") or k.endswith("emb_layer_norm_after.weight"):
            idx = torch.randperm(v.numel(), generator=g)[:3]
            v[idx] *= 25.0
        elif k.endswith(".bias") and v.dim() == 1:
            v += torch.randn(v.shape, generator=g) * 0.05
        out[k] = v.to(torch.bfloat16).float()                      # keep the case bf16-exact like the other cases
    return out


def _encoder_errors(path, case, name, mod):
    """Normalised max error of the last hidden state against the fp32 oracle: (the reference's own stack -- HF modules in
    bf16 on the GPU, eager attention --, this repo)."""
    from oracle.esm_oracle import esm_encoder_forward
    from oracle.ref_import import build_hf_encoder
    rows = [case.batch.omic_ids[b, i] for b, row in enumerate(case.batch.omic_info_list) for i, info in enumerate(row)
            if info["type"] != "pad" and (info["type"] == "protein") == (name == "protein")]
    ids = torch.stack(rows)
    with torch.no_grad():
        ref = esm_encoder_forward(mod.spec, mod.weights, ids)
        hf = build_hf_encoder(mod.spec, mod.weights).to(DEV).to(torch.bfloat16)
        hf_out = hf(input_ids=ids.to(DEV), attention_mask=(ids != 1).to(DEV),
                    output_hidden_states=True).hidden_states[-1].float().cpu()
        ours = path.encode(name, ids.to(DEV)).float().cpu().view_as(ref)
    scale = float(ref.abs().max())
    return float((hf_out - ref).abs().max()) / scale, float((ours - ref).abs().max()) / scale


@pytest.mark.parametrize("trained_like", [False, True])
def test_closer_to_fp32_than_the_references_own_bf16_run(trained_like):
    """The reference executes its encoders in bf16 (src/inference_lora.py:249, DeepSpeed bf16).  Against the fp32 oracle this
    path must be at least as accurate as THAT execution (stock HF modules, bf16, same GPU) -- also for weights with trained-
    checkpoint statistics (peaked softmax rows, outlier channels), where ANY bf16 evaluation is chaotic (near-ties in a hard
    softmax flip) and the absolute 2e-2 bar of the random-init cases cannot be the criterion."""
    case = cases.build_case("trained_like", "tiny_ntv2", "tiny_esm2", D=128, K=200, T=700, seed=5000,
                            samples=[[("dna", 200), ("protein", 150)], [("rna", 77), ("protein", 200)]], mask_tokens=True)
    if trained_like:
        case.nt.weights = _trained_like(case.nt.weights, 1)
        case.pr.weights = _trained_like(case.pr.weights, 2)
    path = build_path(case)
    try:
        for name, mod in (("protein", case.pr), ("dna_rna", case.nt)):
            hf_err, our_err = _encoder_errors(path, case, name, mod)
            print(f"[{name} trained_like={trained_like}] vs fp32 oracle: HF bf16 {hf_err:.4f}  molly_b200 {our_err:.4f}")
            assert our_err <= max(TOL, hf_err), (name, our_err, hf_err)
            if not trained_like:
                assert our_err <= TOL
    finally:
        path.close()


@pytest.mark.parametrize("workload", ["molly_1p7b", "molly_4b", "molly_8b"])
def test_full_size_vs_fp32_oracle_run_on_the_gpu(workload):
    """BASELINE configs[1..3] at FULL size -- [1]: ESM-2 650M + NT-v2 500M, D=2048, 64 x (1 DNA + 1 protein) x 1024 tokens;
    [2]: D=2560, 64 sequences of mixed kinds with log-uniform lengths in K=2048 slots; [3]: NT-v1 2.5B, D=4096, 256 x 1000-bp
    DNA windows -- against the fp32 oracle itself.  The oracle is plain torch code, so as the CHECKER it can run on the GPU in
    fp32 (TF32 off), eight samples at a time; the candidate is the bf16 product path.  Same (bf16-exact) weights on both sides."""
    import bench
    from oracle.esm_oracle import SPECS, OracleModality
    from molly_b200.config import EncoderConfig
    from molly_b200.omics_path import FastOmicsPath
    from molly_b200.packing import PackedEncoder
    torch.backends.cuda.matmul.allow_tf32 = False
    wl = dict(bench.WORKLOADS[workload])
    if torch.cuda.mem_get_info()[0] < 80 * 2 ** 30:
        wl["B"] = 16
    dev = torch.device(DEV, 0)
    mods, packed = [], []
    for i, key in enumerate(("nt", "pr")):
        e = bench.ENC[wl[key]]
        sd = {k: v.to(torch.bfloat16).float() for k, v in bench.gpu_state_dict(e, dev, 70 + i).items()}
        g = torch.Generator(device=DEV).manual_seed(80 + i)
        proj = {"weight": (torch.randn(wl["D"], e["hidden_size"], device=DEV, generator=g) / e["hidden_size"] ** 0.5
                           ).to(torch.bfloat16).float(),
                "bias": (torch.randn(wl["D"], device=DEV, generator=g) * 0.02).to(torch.bfloat16).float()}
        mods.append(OracleModality(SPECS[wl[key]], sd, proj, wl["K"]))
        packed.append(PackedEncoder(EncoderConfig.from_mapping(dict(e, name=wl[key])), sd, proj, wl["K"], dev))
    path = FastOmicsPath(packed[0], packed[1], strict=True)
    try:
        omic_ids, infos = bench.make_inputs(wl, seed=321)
        g = torch.Generator(device=DEV).manual_seed(9)
        base = (torch.randn(wl["B"], wl["T"], wl["D"], device=DEV, generator=g) * 0.02).to(torch.bfloat16)
        got = path.process_omic_sequences(base.clone(), omic_ids, infos, dev).float()
        ref = base.float()
        ids_dev = omic_ids.to(DEV)
        with torch.no_grad():
            for s0 in range(0, wl["B"], 8):
                oracle_process(ref[s0:s0 + 8], ids_dev[s0:s0 + 8], infos[s0:s0 + 8], mods[0], mods[1])
        scale = float(ref.abs().max())
        err = float((got - ref).abs().max()) / scale
        print(f"[{workload} full size, B={wl['B']}] normalised max err vs fp32 oracle (run on the GPU): {err:.4f}")
        assert err <= TOL, err
        written = (ref != base.float()).any(-1)
        assert torch.equal((got != base.float()).any(-1), written)            # same written-row index set
        # the stricter readings SURVEY 8d asks to report: element-wise pass-rate and per-row cosine over the written rows
        w_got, w_ref = got[written], ref[written]
        rms = w_ref.pow(2).mean().sqrt()
        pass_rate = float(((w_got - w_ref).abs() <= TOL * w_ref.abs() + TOL * rms).float().mean())
        min_cos = float(torch.nn.functional.cosine_similarity(w_got, w_ref, dim=1).min())
        print(f"[{workload} full size] element-wise pass-rate {pass_rate:.6f} (|d| <= {TOL}|ref| + {TOL} rms), "
              f"min row cosine {min_cos:.6f}")
        assert pass_rate >= 0.999, pass_rate
        del w_got, w_ref
        assert int(written.sum()) == wl["K"] * sum(len(r) for r in infos)
    finally:
        path.close()


def test_cfg5_projector_grads_full_size_vs_fp32_oracle_run_on_the_gpu():
    """BASELINE configs[4] (train step, --train-mlp): projector weight / bias gradients of both modalities at full model size
    (ESM-2 650M + NT-v2 500M, D=2048, B=8 x (1 DNA + 1 protein) x 1024) against autograd of the fp32 oracle run on the GPU."""
    import bench
    from oracle.esm_oracle import SPECS, OracleModality
    from molly_b200.config import EncoderConfig
    from molly_b200.omics_path import FastOmicsPath
    from molly_b200.packing import PackedEncoder
    torch.backends.cuda.matmul.allow_tf32 = False
    wl = dict(bench.WORKLOADS["train_1p7b"])
    dev = torch.device(DEV, 0)
    mods, packed, lins = [], [], {}
    for i, (key, name) in enumerate((("nt", "dna_rna"), ("pr", "protein"))):
        e = bench.ENC[wl[key]]
        sd = {k: v.to(torch.bfloat16).float() for k, v in bench.gpu_state_dict(e, dev, 90 + i).items()}
        g = torch.Generator(device=DEV).manual_seed(95 + i)
        proj = {"weight": (torch.randn(wl["D"], e["hidden_size"], device=DEV, generator=g) / e["hidden_size"] ** 0.5
                           ).to(torch.bfloat16).float(),
                "bias": (torch.randn(wl["D"], device=DEV, generator=g) * 0.02).to(torch.bfloat16).float()}
        mods.append(OracleModality(SPECS[wl[key]], sd, {k: v.clone().requires_grad_(True) for k, v in proj.items()},
                                   wl["K"]))
        packed.append(PackedEncoder(EncoderConfig.from_mapping(dict(e, name=wl[key])), sd, proj, wl["K"], dev))
        lin = torch.nn.Linear(e["hidden_size"], wl["D"], device=dev, dtype=torch.bfloat16)
        lin.load_state_dict(proj)
        lins[name] = lin
    path = FastOmicsPath(packed[0], packed[1], strict=True)
    path._proj_modules = lins
    try:
        omic_ids, infos = bench.make_inputs(wl, seed=654)
        g = torch.Generator(device=DEV).manual_seed(19)
        base = (torch.randn(wl["B"], wl["T"], wl["D"], device=DEV, generator=g) * 0.02).to(torch.bfloat16)
        d_out = (torch.randn(wl["B"], wl["T"], wl["D"], device=DEV, generator=g) * 1e-2).to(torch.bfloat16)
        out = path.process_omic_sequences(base.clone(), omic_ids, infos, dev)
        out.backward(d_out)
        ref = oracle_process(base.float(), omic_ids.to(DEV), infos, mods[0], mods[1])
        ref.backward(d_out.float())
        for name, mod in (("dna_rna", mods[0]), ("protein", mods[1])):
            assert_close(f"cfg5 d {name}_projector.weight", lins[name].weight.grad.float(), mod.projector["weight"].grad, TOL)
            assert_close(f"cfg5 d {name}_projector.bias", lins[name].bias.grad.float(), mod.projector["bias"].grad, TOL)
    finally:
        path.close()
