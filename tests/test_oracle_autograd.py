"""CPU: pins the ORACLE's gradients (autograd of the restated forward, the checker of tests/test_gpu_train.py) to autograd of
the modules the reference actually trains with -- stock HF ``EsmForMaskedLM`` (plus the NT-v2 gated-FFN subclass statement,
parity unpinned) in fp32, eager attention.  Every encoder parameter, three architectures."""
import pytest
import torch

from oracle import synth
from oracle.esm_oracle import SPECS, esm_encoder_forward, init_encoder_weights
from oracle.ref_import import build_hf_encoder


@pytest.mark.parametrize("spec_name", ["tiny_esm2", "tiny_ntv1", "tiny_ntv2"])
def test_oracle_autograd_equals_hf_autograd(spec_name):
    spec = SPECS[spec_name]
    W = init_encoder_weights(spec, 17)
    g = torch.Generator().manual_seed(18)
    K = 40
    ids = torch.stack([synth.protein_ids(g, K, v) if spec.vocab_size == 33 else synth.nucleotide_ids(g, K, v, spec.vocab_size)
                       for v in (40, 13, 29)])
    if spec.token_dropout:
        ids[1, 3] = spec.mask_token_id
    d_out = torch.randn(3, K, spec.hidden_size, generator=g)
    Wg = {k: v.clone().requires_grad_(v.dtype.is_floating_point) for k, v in W.items()}
    (esm_encoder_forward(spec, Wg, ids) * d_out).sum().backward()
    hf = build_hf_encoder(spec, W)
    for prm in hf.parameters():
        prm.requires_grad_(True)
    out = hf(input_ids=ids, attention_mask=(ids != 1).long(), output_hidden_states=True).hidden_states[-1]
    (out * d_out).sum().backward()
    checked = 0
    for name, prm in hf.named_parameters():
        if prm.grad is None or name not in Wg:
            continue                                     # lm_head / contact head: not on the path (omics_one.py:91)
        ref, mine = prm.grad, Wg[name].grad
        assert mine is not None, name
        scale = float(ref.abs().max())
        assert float((mine - ref).abs().max()) <= 1e-4 * max(scale, 1e-3), (name, float((mine - ref).abs().max()), scale)
        checked += 1
    assert checked >= 12 * spec.num_hidden_layers + 3
