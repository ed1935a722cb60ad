"""CPU, build container only: the oracle restatement against the reference's own code run live
(/root/reference/src/model/omics_one.py + HF EsmForMaskedLM).  Skipped where /root/reference is absent (GPU box)."""
import pytest
import torch

from oracle import cases, ref_import
from oracle.esm_oracle import SPECS, esm_encoder_forward, init_encoder_weights, process_omic_sequences

pytestmark = pytest.mark.skipif(not ref_import.reference_available(), reason="/root/reference not mounted")


@pytest.mark.parametrize("spec_name", ["tiny_esm2", "tiny_ntv1", "tiny_ntv2"])
def test_encoder_matches_hf(spec_name):
    spec = SPECS[spec_name]
    W = init_encoder_weights(spec, 7)
    model = ref_import.build_hf_encoder(spec, W)
    g = torch.Generator().manual_seed(8)
    ids = torch.randint(4, min(spec.vocab_size, 30), (3, 50), generator=g)
    ids[0, 40:] = 1
    ids[1, 10:] = 1
    ids[2, 7] = spec.mask_token_id
    with torch.no_grad():
        out = model(ids, attention_mask=(ids != 1).long(), output_hidden_states=True, return_dict=True)
        mine, hs = esm_encoder_forward(spec, W, ids, return_all=True)
    assert len(out["hidden_states"]) == spec.num_hidden_layers + 1
    assert torch.allclose(out["hidden_states"][-1], mine, atol=1e-5, rtol=1e-5)      # [-1] is post-final-LN
    assert torch.allclose(out["hidden_states"][0], hs[0], atol=1e-6)


def test_boundary_matches_reference_and_error_conventions():
    case = cases.golden_cases()["tiny_rotary_glu"]
    om = ref_import.build_reference_omics(case.nt, case.pr, case.D)
    h1, h2 = case.batch.hidden_states.clone(), case.batch.hidden_states.clone()
    with torch.no_grad():
        r = om.process_omic_sequences(h1, case.batch.omic_ids, case.batch.omic_info_list, h1.device)
        o = process_omic_sequences(h2, case.batch.omic_ids, case.batch.omic_info_list, case.nt, case.pr)
    assert r is h1 and o is h2
    assert float((r - o).abs().max()) <= 1e-5
    # unknown type -> ValueError (omics_one.py:118)
    infos = [[dict(i) for i in row] for row in case.batch.omic_info_list]
    infos[0][0]["type"] = "lipid"
    for fn in (lambda: om.process_omic_sequences(h1, case.batch.omic_ids, infos, h1.device),
               lambda: process_omic_sequences(h2, case.batch.omic_ids, infos, case.nt, case.pr)):
        with pytest.raises(ValueError, match="Unsupported omic type"):
            fn()
    # out-of-vocab id -> AssertionError (omics_one.py:71-72)
    bad = case.batch.omic_ids.clone()
    bad[0, 1, 3] = 999
    for fn in (lambda: om.process_omic_sequences(h1, bad, case.batch.omic_info_list, h1.device),
               lambda: process_omic_sequences(h2, bad, case.batch.omic_info_list, case.nt, case.pr)):
        with pytest.raises(AssertionError):
            fn()
    # placement past T -> RuntimeError (slice shape mismatch)
    infos = [[dict(i) for i in row] for row in case.batch.omic_info_list]
    infos[1][0]["start"] = case.T - 5
    for fn in (lambda: om.process_omic_sequences(h1, case.batch.omic_ids, infos, h1.device),
               lambda: process_omic_sequences(h2, case.batch.omic_ids, infos, case.nt, case.pr)):
        with pytest.raises(RuntimeError):
            fn()
