"""GPU parity of SURVEY.md 8f row N2: one CUDA graph per (B, T, modality layout, K) bucket of the fused input path.
A replay must be bit-identical to the eager ``embed_and_process`` on the same inputs, for every new batch of the bucket."""
import pytest
import torch

from oracle import cases, synth
from tests.test_gpu_path import build_path

pytestmark = pytest.mark.gpu
DEV = "cuda"
VOCAB = synth.PLACEHOLDER_BASE + 16


@pytest.mark.parametrize("name", ["tiny_rotary_glu", "tiny_absolute_leftpad"])
def test_graph_replay_equals_eager(name):
    case = cases.golden_cases()[name]
    g = torch.Generator().manual_seed(3)
    table = (torch.randn(VOCAB, case.D, generator=g) * 0.02).to(torch.bfloat16).to(DEV)
    types = [[i["type"] for i in row if i["type"] != "pad"] for row in case.batch.omic_info_list]
    B, T = case.batch.input_ids.shape
    path = build_path(case)
    try:
        call = path.graphed(table, B, T, types, case.K, synth.PAD_TOKEN_IDS)
        for trial in range(3):
            ids = case.batch.input_ids.clone()
            text = ~torch.isin(ids, torch.arange(synth.PLACEHOLDER_BASE, synth.PLACEHOLDER_BASE + 9))
            ids[text] = (ids[text] + 17 * trial) % 1000                     # new text tokens, same placeholder layout
            omic = case.batch.omic_ids.clone()
            if trial:                                                       # new sequences: rotate each one's valid tokens
                for row in omic.view(-1, omic.shape[-1]):
                    n = int((row != 1).sum())
                    if n > 3:
                        row[1:n - 1] = row[1:n - 1].roll(trial)
            got = call(ids, omic).clone()
            want = path.embed_and_process(ids.to(DEV), table, omic, case.batch.omic_info_list, synth.PAD_TOKEN_IDS)
            assert torch.equal(got, want), f"trial {trial}: graph replay differs from the eager call"
            if trial:
                assert not torch.equal(got, first)
            else:
                first = got
    finally:
        path.close()


def test_graph_replay_reports_device_errors():
    case = cases.golden_cases()["tiny_rotary_glu"]
    table = torch.zeros(VOCAB, case.D, dtype=torch.bfloat16, device=DEV)
    types = [[i["type"] for i in row if i["type"] != "pad"] for row in case.batch.omic_info_list]
    B, T = case.batch.input_ids.shape
    path = build_path(case, strict=True)
    try:
        call = path.graphed(table, B, T, types, case.K, synth.PAD_TOKEN_IDS)
        call(case.batch.input_ids, case.batch.omic_ids)                    # fine
        wrong = case.batch.input_ids.clone()
        wrong[wrong == synth.PAD_TOKEN_IDS[0]] = 5                          # the DNA runs vanish from the text
        with pytest.raises(RuntimeError, match="do not pair"):
            call(wrong, case.batch.omic_ids)
    finally:
        path.close()
