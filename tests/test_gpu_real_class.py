"""GPU: the drop-in installed on the REFERENCE'S OWN class.  ``oracle/_ref/omics_one.py`` is an unmodified copy of
``/root/reference/src/model/omics_one.py`` staged by ``__graft_entry__.build()`` (git-ignored, shipped to the GPU box by gpurun).
A real ``OmicsOne`` is built with stock HF ``EsmForMaskedLM`` encoders and a stub LLM that only exposes what the prologue of
``forward`` (:163-173) and ``generate`` (:209-218) touches -- ``get_input_embeddings()``, ``__call__(inputs_embeds=...)``,
``generate(inputs_embeds=...)``, ``config`` -- then ``FastOmicsPath.from_omics_one(model).install(model)`` swaps the method and
the SAME forward / generate code runs on the GPU.  Compared with the un-installed instance running the reference's own
``process_omic_sequences`` in fp32 on the CPU."""
import copy
import types

import pytest
import torch

from oracle import cases, ref_import, synth
from tests.util import assert_close, assert_parity

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL = 2e-2


class _StubLLM(torch.nn.Module):
    """What OmicsOne.forward / generate need from ``self.model``: the embedding table and a consumer of inputs_embeds."""

    def __init__(self, vocab: int, d: int):
        super().__init__()
        self.embed_tokens = torch.nn.Embedding(vocab, d)
        self.config = types.SimpleNamespace(pad_token_id=0, eos_token_id=1)
        self.seen = {}

    def get_input_embeddings(self):
        return self.embed_tokens

    def forward(self, inputs_embeds=None, **kw):
        self.seen = dict(kw)
        return inputs_embeds

    def generate(self, inputs_embeds=None, **kw):
        self.seen = dict(kw)
        return inputs_embeds


def _build(case):
    om = ref_import.build_reference_omics(case.nt, case.pr, case.D)        # the reference's OmicsOne + HF encoders, fp32
    torch.manual_seed(5)
    om.model = _StubLLM(300, case.D)
    return om.eval()


@pytest.mark.skipif(not ref_import.reference_available(), reason="oracle/_ref/omics_one.py not staged (run build())")
@pytest.mark.parametrize("name", ["tiny_rotary_glu", "tiny_absolute_leftpad"])
def test_install_on_the_real_omics_one_forward_and_generate(name):
    from molly_b200.omics_path import FastOmicsPath
    case = cases.golden_cases()[name]
    ref_om = _build(case)
    B, T = case.batch.hidden_states.shape[:2]
    input_ids = torch.randint(0, 300, (B, T), generator=torch.Generator().manual_seed(6))
    with torch.no_grad():
        want = ref_om(input_ids=input_ids, omic_ids=case.batch.omic_ids, omic_info_list=case.batch.omic_info_list)
    text = ref_om.model.get_input_embeddings()(input_ids)

    fast_om = copy.deepcopy(ref_om).to(DEV)
    path = FastOmicsPath.from_omics_one(fast_om, DEV, strict=True)
    path.install(fast_om)
    try:
        ids_dev = input_ids.to(DEV)
        with torch.no_grad():
            got = fast_om(input_ids=ids_dev, omic_ids=case.batch.omic_ids, omic_info_list=case.batch.omic_info_list,
                          attention_mask=torch.ones(B, T, dtype=torch.long, device=DEV))
        assert got.is_cuda and "attention_mask" in fast_om.model.seen       # the rest of forward ran as written
        exp = synth.expected_rows(case.batch.omic_info_list, case.K, case.nt.project_token_num, case.pr.project_token_num)
        written = torch.zeros(B, T, dtype=torch.bool)
        for (b, t) in exp:
            written[b, t] = True
        g = got.float().cpu()
        assert torch.equal((g != text).any(-1), written), "written-row index set differs from the reference's"
        assert torch.equal(g[~written], text[~written])
        assert_close(f"real OmicsOne.forward {name}", g, want, TOL)
        assert_parity(f"real OmicsOne.forward {name} (written rows)", g[written], want[written], TOL)
        gen = fast_om.generate(input_ids=ids_dev, omic_ids=case.batch.omic_ids, omic_info_list=case.batch.omic_info_list)
        assert fast_om.model.seen.get("max_new_tokens") == 3072               # generate() body ran as written (:219-231)
        assert_close(f"real OmicsOne.generate {name}", gen.float().cpu(), want, TOL)
        # the sanity check of the prologue is the reference's own (:166-170)
        with pytest.raises(AssertionError):
            fast_om(input_ids=ids_dev, omic_ids=case.batch.omic_ids, omic_info_list=[r[:-1] for r in case.batch.omic_info_list])
        # autograd through the installed method: projector grads, zero grad on the overwritten embedding rows
        out = fast_om(input_ids=ids_dev, omic_ids=case.batch.omic_ids, omic_info_list=case.batch.omic_info_list)
        gw = torch.randn(out.shape, generator=torch.Generator().manual_seed(7))
        (out * gw.to(DEV)).sum().backward()
        out_ref = ref_om(input_ids=input_ids, omic_ids=case.batch.omic_ids, omic_info_list=case.batch.omic_info_list)
        (out_ref * gw).sum().backward()
        assert_close("d protein_projector.weight", fast_om.protein_projector.weight.grad.float().cpu(),
                     ref_om.protein_projector.weight.grad, TOL)
        assert_close("d dna_rna_projector.bias", fast_om.dna_rna_projector.bias.grad.float().cpu(),
                     ref_om.dna_rna_projector.bias.grad, TOL)
        assert_close("d embed_tokens.weight", fast_om.model.embed_tokens.weight.grad.float().cpu(),
                     ref_om.model.embed_tokens.weight.grad, TOL)
    finally:
        path.close()
