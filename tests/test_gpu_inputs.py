"""GPU parity of SURVEY.md 8f row N1: the input producer on the device.

``FastOmicsPath.embed_and_process(input_ids, embed_weight, omic_ids, omic_info_list, pad_token_ids)`` must equal the
reference's ``process_omic_sequences(embed_tokens(input_ids), ...)`` (omics_one.py:164-170 / :209-215) while reading only
``info["type"]``: the ``start`` positions come from the run scan of ``input_ids`` on the device.
  * runs / seq_table: bit-exact against the Python restatement of the dataset's bookkeeping (oracle.synth.placeholder_runs)
  * text rows (and pad rows beyond the K cap): bit-identical to ``embed_weight[input_ids]``
  * written rows: <= 2e-2 normalised max error against the fp32 oracle
"""
import pytest
import torch

from oracle import cases, synth
from tests.test_gpu_path import build_path, TOL
from tests.util import assert_close

pytestmark = pytest.mark.gpu
DEV = "cuda"
VOCAB = synth.PLACEHOLDER_BASE + 16


def random_layout(seed, B, T, max_runs, pads=synth.PAD_TOKEN_IDS):
    """input_ids with random text and random runs (start at 0, end at T-1, across 256-token chunk borders ...)."""
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(0, 1000, (B, T), generator=g)
    for b in range(B):
        n = int(torch.randint(0, max_runs + 1, (1,), generator=g))
        t = 0 if b % 3 == 0 else int(torch.randint(0, 8, (1,), generator=g))
        for _ in range(n):
            if t >= T:
                break
            ln = int(torch.randint(1, 400, (1,), generator=g))
            kind = int(torch.randint(0, 3, (1,), generator=g))
            ids[b, t:t + ln] = pads[kind]
            t += ln + int(torch.randint(1, 300, (1,), generator=g))
    ids[B - 1, T - 5:] = pads[2]                       # a run that ends exactly at T
    return ids


@pytest.mark.parametrize("B,T,max_runs", [(1, 7, 1), (5, 255, 2), (4, 256, 3), (7, 1000, 4), (3, 4097, 12)])
def test_placeholder_runs_vs_python(B, T, max_runs):
    from molly_b200 import ops
    ids = random_layout(B * 1000 + T, B, T, max_runs)
    want = synth.placeholder_runs(ids, synth.PAD_TOKEN_IDS)
    cap = max(1, max(len(r) for r in want))
    rs, rk, rl, nr, pj = [t.cpu() for t in ops.placeholder_runs(ids.to(DEV), synth.PAD_TOKEN_IDS, None, cap)]
    want_j = torch.full((B, T), -1, dtype=torch.int32)
    for b, runs in enumerate(want):
        assert int(nr[b]) == len(runs)
        for r, (s, kind, ln) in enumerate(runs):
            assert (int(rs[b, r]), int(rk[b, r]), int(rl[b, r])) == (s, kind, ln), (b, r)
            want_j[b, s:s + ln] = torch.arange(ln, dtype=torch.int32)
    assert torch.equal(pj, want_j)
    # a too-small run table must truncate, not overflow
    if cap > 1:
        rs2, _, _, nr2, pj2 = [t.cpu() for t in ops.placeholder_runs(ids.to(DEV), synth.PAD_TOKEN_IDS, None, 1)]
        assert torch.equal(nr2, nr) and torch.equal(pj2, want_j) and torch.equal(rs2[:, 0], rs[:, 0])


def test_runs_beyond_slots_are_text():
    """Runs past the sample's omic_ids slots are never overwritten (the reference's zip stops): pos_j = -1 there."""
    from molly_b200 import ops
    ids = random_layout(77, 4, 1500, 4)
    want = synth.placeholder_runs(ids, synth.PAD_TOKEN_IDS)
    n_slots = torch.tensor([max(0, len(r) - 1) for r in want], dtype=torch.int32)
    pj = ops.placeholder_runs(ids.to(DEV), synth.PAD_TOKEN_IDS, n_slots.to(DEV), 8)[4].cpu()
    for b, runs in enumerate(want):
        for r, (s, _, ln) in enumerate(runs):
            exp = torch.arange(ln, dtype=torch.int32) if r < int(n_slots[b]) else torch.full((ln,), -1, dtype=torch.int32)
            assert torch.equal(pj[b, s:s + ln], exp), (b, r)


def _case(name):
    return cases.golden_cases()[name]


@pytest.mark.parametrize("name", ["tiny_rotary_glu", "tiny_absolute_leftpad", "tiny_kcap", "tiny_long"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_embed_and_process_vs_oracle(name, dtype):
    case = _case(name)
    g = torch.Generator().manual_seed(11)
    table = (torch.randn(VOCAB, case.D, generator=g) * 0.02).to(dtype)
    with torch.no_grad():
        ref = synth.embed_then_process(case.batch.input_ids, table.float(), case.batch.omic_ids,
                                       case.batch.omic_info_list, case.nt, case.pr)
    # the fused path may read ONLY the types: poison the starts
    infos = [[{"type": i["type"], "start": -12345} for i in row] for row in case.batch.omic_info_list]
    path = build_path(case)
    try:
        got = path.embed_and_process(case.batch.input_ids.to(DEV), table.to(DEV), case.batch.omic_ids, infos,
                                     synth.PAD_TOKEN_IDS)
        assert got.dtype == dtype and got.shape == ref.shape
        got = got.float().cpu()
        exp = synth.expected_rows(case.batch.omic_info_list, case.K, case.nt.project_token_num, case.pr.project_token_num)
        written = torch.zeros(got.shape[:2], dtype=torch.bool)
        for (b, t) in exp:
            written[b, t] = True
        plain = table.float()[case.batch.input_ids]
        assert torch.equal(got[~written], plain[~written]), "text rows must be the plain embedding lookup, bit for bit"
        assert_close(f"N1 {name} {dtype}", got, ref, TOL)
        assert_close(f"N1 {name} {dtype} (written rows)", got[written], ref[written], TOL)
        # and it equals the two-step product path bit for bit (same kernels, same order)
        hs = table.to(DEV)[case.batch.input_ids.to(DEV)]
        two = path.process_omic_sequences(hs, case.batch.omic_ids, case.batch.omic_info_list, hs.device)
        assert torch.equal(two.float().cpu(), got)
    finally:
        path.close()


def test_seq_table_from_runs_equals_reference_starts():
    from molly_b200 import ops, planner
    case = _case("tiny_absolute_leftpad")
    B = case.batch.input_ids.shape[0]
    nt, pr = planner.route(B, case.batch.omic_ids, case.batch.omic_info_list)
    n_slots = [0] * B
    for plan in (nt, pr):
        for b, r in zip(plan.b_idx, plan.run_idx):
            n_slots[b] = max(n_slots[b], r + 1)
    runs = ops.placeholder_runs(case.batch.input_ids.to(DEV), synth.PAD_TOKEN_IDS,
                                torch.tensor(n_slots, dtype=torch.int32, device=DEV), max(n_slots))
    for plan, is_pr in ((nt, False), (pr, True)):
        if len(plan) == 0:
            continue
        idx = torch.tensor([plan.b_idx, plan.run_idx], dtype=torch.int32, device=DEV)
        table = ops.build_seq_table(idx[0], idx[1], runs, expect_protein=is_pr).cpu()
        assert torch.equal(table, plan.seq_table())
    ops.check_device_errors(torch.device(DEV, 0))


def test_layout_mismatch_and_bad_token_raise():
    case = _case("tiny_rotary_glu")
    table = torch.zeros(VOCAB, case.D, device=DEV)
    wrong = case.batch.input_ids.clone()                # the text says "protein run" where the ids hold a DNA sequence
    wrong[wrong == synth.PAD_TOKEN_IDS[0]] = synth.PAD_TOKEN_IDS[2]
    path = build_path(case)
    try:
        with pytest.raises(RuntimeError, match="do not pair"):
            path.embed_and_process(wrong.to(DEV), table, case.batch.omic_ids, case.batch.omic_info_list,
                                   synth.PAD_TOKEN_IDS)
        bad = case.batch.input_ids.clone()
        bad[0, 0] = VOCAB + 3
        with pytest.raises(IndexError):
            path.embed_and_process(bad.to(DEV), table, case.batch.omic_ids, case.batch.omic_info_list, synth.PAD_TOKEN_IDS)
        with pytest.raises(AssertionError, match="Mismatch in omic count"):
            path.embed_and_process(case.batch.input_ids.to(DEV), table, case.batch.omic_ids,
                                   [row[:-1] for row in case.batch.omic_info_list], synth.PAD_TOKEN_IDS)
    finally:
        path.close()


def test_rejected_runs_are_embedded_like_text_in_lazy_mode():
    """strict=False (the default): a layout whose runs do not pair with the ids (wrong kind, or a run shorter than K after
    text truncation, omics_dataset.py:370-373) only sets MOLLY_ERRBIT_LAYOUT.  The rows of those runs must then hold the
    plain embedding lookup -- the output buffer is torch.empty, nothing may be left unwritten."""
    from molly_b200 import ops
    case = _case("tiny_rotary_glu")
    g = torch.Generator().manual_seed(12)
    table = (torch.randn(VOCAB, case.D, generator=g) * 0.02)
    ids = case.batch.input_ids.clone()
    b0 = 0
    dna_run = (ids[b0] == synth.PAD_TOKEN_IDS[0]).nonzero().flatten()
    ids[b0, dna_run[-3:]] = 17                              # sample 0: the DNA run is 3 tokens short of K
    pr_rows = (ids == synth.PAD_TOKEN_IDS[2]).any(1).nonzero().flatten()
    b1 = int(pr_rows[-1])
    pr_pos = ids[b1] == synth.PAD_TOKEN_IDS[2]
    ids[b1, pr_pos] = synth.PAD_TOKEN_IDS[1]                # last sample: "rna" text where the ids hold a protein
    path = build_path(case, strict=False)
    try:
        # poison the caching allocator's next block so that an unwritten row cannot look right by accident
        junk = torch.full((ids.shape[0], ids.shape[1], case.D), float("nan"), device=DEV)
        del junk
        got = path.embed_and_process(ids.to(DEV), table.to(DEV), case.batch.omic_ids, case.batch.omic_info_list,
                                     synth.PAD_TOKEN_IDS)
        assert int(ops.error_flag(torch.device(DEV, 0)).item()) & 8          # MOLLY_ERRBIT_LAYOUT, not raised in lazy mode
        ops.error_flag(torch.device(DEV, 0)).zero_()
        got = got.float().cpu()
        assert not torch.isnan(got).any()
        plain = table[ids]
        rejected = torch.zeros(ids.shape, dtype=torch.bool)
        rejected[b0, dna_run[:-3]] = True
        rejected[b1, pr_pos] = True
        assert torch.equal(got[rejected], plain[rejected]), "rows of rejected runs must be the plain lookup"
        text = ~torch.isin(ids, torch.tensor(synth.PAD_TOKEN_IDS))
        assert torch.equal(got[text], plain[text])
    finally:
        path.close()
