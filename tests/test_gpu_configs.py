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
from tests.test_gpu_path import build_path, check_merged
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
