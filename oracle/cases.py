"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py) -- named, seeded parity cases shared by the golden generator,
the CPU tests, the GPU tests, smoke() and the bench's CPU leg.  Everything is regenerated from seeds; the committed
fixtures under tests/golden/ hold the REFERENCE's outputs for the small cases plus checksums of the regenerated
weights / inputs so that a drifting RNG stream is detected instead of silently comparing different problems."""
from __future__ import annotations

import hashlib
from dataclasses import dataclass
from typing import Dict, List, Sequence, Tuple

import torch

from . import synth
from .esm_oracle import SPECS, EncoderSpec, OracleModality, init_encoder_weights, init_projector


@dataclass
class Case:
    name: str
    nt: OracleModality
    pr: OracleModality
    batch: synth.Batch
    D: int
    K: int
    T: int


def _tensor_digest(tensors: Sequence[torch.Tensor]) -> str:
    h = hashlib.sha256()
    for t in tensors:
        h.update(t.detach().contiguous().cpu().numpy().tobytes())
    return h.hexdigest()[:16]


def case_digest(c: Case) -> str:
    ts = [c.nt.weights[k] for k in sorted(c.nt.weights)] + [c.pr.weights[k] for k in sorted(c.pr.weights)]
    ts += [c.nt.projector["weight"], c.nt.projector["bias"], c.pr.projector["weight"], c.pr.projector["bias"]]
    ts += [c.batch.input_ids, c.batch.omic_ids, c.batch.hidden_states]
    return _tensor_digest(ts)


def build_case(name: str, nt_spec: str, pr_spec: str, D: int, K: int, T: int,
               samples: Sequence[Sequence[Tuple[str, int]]], seed: int, left_pad: bool = False,
               mask_tokens: bool = True, k_cap_nt: int = None, k_cap_pr: int = None) -> Case:
    nts, prs = SPECS[nt_spec], SPECS[pr_spec]
    nt = OracleModality(nts, init_encoder_weights(nts, seed + 1), init_projector(nts.hidden_size, D, seed + 2),
                        K if k_cap_nt is None else k_cap_nt)
    pr = OracleModality(prs, init_encoder_weights(prs, seed + 3), init_projector(prs.hidden_size, D, seed + 4),
                        K if k_cap_pr is None else k_cap_pr)
    bt = synth.make_batch(seed, samples, T=T, D=D, K=K, nt_spec=nts, pr_spec=prs, left_pad=left_pad)
    if mask_tokens:            # exercise the token-dropout rescale (HF:213-222) on the first protein sequence
        for b, infos in enumerate(bt.omic_info_list):
            for i, info in enumerate(infos):
                if info["type"] == "protein" and K > 8:
                    bt.omic_ids[b, i, 5] = prs.mask_token_id
                    return Case(name, nt, pr, bt, D, K, T)
    return Case(name, nt, pr, bt, D, K, T)


def lengths_loguniform(g: torch.Generator, n: int, lo: int, hi: int) -> List[int]:
    import math
    u = torch.rand(n, generator=g)
    return [int(round(math.exp(math.log(lo) + float(x) * (math.log(hi) - math.log(lo))))) for x in u]


# ---- small cases whose REFERENCE outputs are committed as fixtures (tests/golden/<name>.npz) -------------------------
def golden_cases() -> Dict[str, Case]:
    cs = {}
    cs["tiny_rotary_glu"] = build_case(
        "tiny_rotary_glu", "tiny_ntv2", "tiny_esm2", D=96, K=40, T=200, seed=100,
        samples=[[("dna", 40), ("protein", 17)], [("protein", 40)], [("rna", 9), ("dna", 25), ("protein", 33)], []])
    cs["tiny_absolute_leftpad"] = build_case(
        "tiny_absolute_leftpad", "tiny_ntv1", "tiny_esm2", D=64, K=48, T=260, seed=200, left_pad=True,
        samples=[[("rna", 48), ("protein", 3)], [("dna", 20), ("dna", 31), ("protein", 48)], [("protein", 12)]])
    cs["tiny_kcap"] = build_case(                       # project_token_num < K: only the first k rows are written
        "tiny_kcap", "tiny_ntv2", "tiny_esm2", D=96, K=40, T=160, seed=300, k_cap_nt=25, k_cap_pr=32,
        samples=[[("dna", 40), ("protein", 40)], [("rna", 11)]])
    cs["tiny_long"] = build_case(                       # K > 128: several KV blocks, partial last block
        "tiny_long", "tiny_ntv2", "tiny_esm2", D=64, K=300, T=700, seed=400,
        samples=[[("dna", 300), ("protein", 170)], [("protein", 300)], [("rna", 129)]])
    return cs


# ---- BASELINE.json configs[0] "Molly-mini": parity at full size against the oracle (and a row subsample in golden) -----
def molly_mini(varlen: bool = False, seed: int = 1234) -> Case:
    g = torch.Generator().manual_seed(seed)
    if varlen:
        lens = [int(v) for v in torch.randint(32, 513, (8,), generator=g)]
    else:
        lens = [512] * 8
    samples = [[("dna", lens[2 * b]), ("protein", lens[2 * b + 1])] for b in range(4)]
    return build_case("molly_mini_varlen" if varlen else "molly_mini", "nt_v2_50m", "esm2_t6_8m", D=1024, K=512,
                      T=2048, samples=samples, seed=seed, mask_tokens=False)
