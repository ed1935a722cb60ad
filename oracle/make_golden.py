"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz by running the REFERENCE's own code
(/root/reference/src/model/omics_one.py unmodified + HuggingFace EsmForMaskedLM, eager, fp32, CPU).

Run in the build container only (the GPU box has no /root/reference):
    python -m oracle.make_golden
Each fixture stores: the reference's merged hidden_states (fp32), the written-row index set, and a digest of the
seed-regenerated weights/inputs.  `molly_mini` stores a strided row subsample (full tensor would be 32 MB).
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

from . import cases, ref_import, synth

OUT_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def run_reference(case: cases.Case) -> torch.Tensor:
    om = ref_import.build_reference_omics(case.nt, case.pr, case.D)
    hs = case.batch.hidden_states.clone()
    with torch.no_grad():
        out = om.process_omic_sequences(hs, case.batch.omic_ids, case.batch.omic_info_list, hs.device)
    assert out is hs
    return out


def written_rows(case: cases.Case, out: torch.Tensor) -> np.ndarray:
    changed = (out != case.batch.hidden_states).any(-1)
    return changed.nonzero().to(torch.int32).numpy()


def main() -> None:
    if not ref_import.reference_available():
        sys.exit("make_golden needs /root/reference (build container only)")
    os.makedirs(OUT_DIR, exist_ok=True)
    import transformers
    meta = {"transformers": transformers.__version__, "torch": torch.__version__, "attn_implementation": "eager",
            "dtype": "float32", "reference": "src/model/omics_one.py:49-136 (unmodified, two import stubs)"}
    for name, case in cases.golden_cases().items():
        out = run_reference(case)
        rows = written_rows(case, out)
        exp = synth.expected_rows(case.batch.omic_info_list, case.K, case.nt.project_token_num,
                                  case.pr.project_token_num)
        assert {tuple(r) for r in rows.tolist()} == set(exp.keys()), name
        np.savez_compressed(os.path.join(OUT_DIR, f"{name}.npz"), merged=out.numpy(), rows=rows,
                            digest=np.frombuffer(cases.case_digest(case).encode(), dtype=np.uint8),
                            meta=np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8))
        print(f"{name}: merged {tuple(out.shape)} rows {len(rows)} digest {cases.case_digest(case)}")
    for varlen in (False, True):
        case = cases.molly_mini(varlen=varlen)
        out = run_reference(case)
        rows = written_rows(case, out)
        sel = rows[::29]                                         # strided subsample of the written rows
        vals = out[torch.from_numpy(sel[:, 0]).long(), torch.from_numpy(sel[:, 1]).long()].numpy()
        np.savez_compressed(os.path.join(OUT_DIR, f"{case.name}.npz"), rows=rows, sel=sel, vals=vals,
                            digest=np.frombuffer(cases.case_digest(case).encode(), dtype=np.uint8),
                            meta=np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8))
        print(f"{case.name}: rows {len(rows)} sampled {len(sel)} digest {cases.case_digest(case)}")


if __name__ == "__main__":
    main()
