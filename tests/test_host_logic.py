"""CPU: host-side planner (routing, seq table, error conventions), config ingestion, ABI export checks."""
import ctypes
import os
import subprocess
import sys

import pytest
import torch

from molly_b200 import planner
from molly_b200.config import EncoderConfig
from oracle import cases, synth


def test_route_matches_reference_order_and_pairs_by_index():
    case = cases.golden_cases()["tiny_rotary_glu"]
    nt, pr = planner.route(4, case.batch.omic_ids, case.batch.omic_info_list)
    # reference appends batch-major, slot order; dna/rna share one list (omics_one.py:104-118)
    exp_nt, exp_pr = [], []
    for b, infos in enumerate(case.batch.omic_info_list):
        for i, info in enumerate(infos):
            if info["type"] in ("dna", "rna"):
                exp_nt.append((b, i, info["start"]))
            elif info["type"] == "protein":
                exp_pr.append((b, i, info["start"]))
    assert list(zip(nt.b_idx, nt.slot_idx, nt.starts)) == exp_nt
    assert list(zip(pr.b_idx, pr.slot_idx, pr.starts)) == exp_pr
    tbl = nt.seq_table()
    assert tbl.dtype == torch.int32 and tuple(tbl.shape) == (len(exp_nt), 2)
    assert tbl.tolist() == [[b, s] for b, _, s in exp_nt]
    ids = planner.gather_ids(case.batch.omic_ids, nt)
    assert torch.equal(ids, torch.stack([case.batch.omic_ids[b, i] for b, i, _ in exp_nt]))
    # list-of-lists input form (omics_one.py signature) gives the same matrix
    as_lists = [[row for row in sample] for sample in case.batch.omic_ids]
    assert torch.equal(planner.gather_ids(as_lists, nt), ids)


def test_route_errors_and_skips():
    ids = torch.ones(1, 3, 8, dtype=torch.int64)
    infos = [[{"type": "pad", "start": -1}, {"type": "dna", "start": -1}, {"type": "lipid", "start": 3}]]
    with pytest.raises(ValueError, match="Unsupported omic type: lipid"):
        planner.route(1, ids, infos)
    nt, pr = planner.route(1, ids, [infos[0][:2]])          # zip stops at the shorter list; 'pad' skipped
    assert len(nt) == 1 and len(pr) == 0 and nt.starts == [-1]
    planner.check_placement(nt, 8, 1, 10)                   # start == -1 is never checked (omics_one.py:94-95)


def test_check_placement_and_vocab():
    p = planner.ModalityPlan([0, 1], [0, 0], [5, 91])
    planner.check_placement(p, 8, 2, 100)                   # 91+1+8 == 100 fits exactly
    with pytest.raises(RuntimeError):
        planner.check_placement(p, 9, 2, 100)
    with pytest.raises(AssertionError, match="out-of-range token"):
        planner.check_vocab(torch.tensor([[1, 2, 33]]), 33)
    planner.check_vocab(torch.tensor([[1, 2, 32]]), 33)


def test_ragged_ids_raise_like_torch_stack():
    ids = [[torch.ones(8, dtype=torch.int64), torch.ones(9, dtype=torch.int64)]]
    infos = [[{"type": "dna", "start": 0}, {"type": "rna", "start": 20}]]
    nt, _ = planner.route(1, ids, infos)
    with pytest.raises(RuntimeError):
        planner.gather_ids(ids, nt)


def test_sharding_helpers():
    assert [list(planner.shard_samples(10, 4, r)) for r in range(4)] == [[0, 1, 2], [3, 4, 5], [6, 7, 8], [9]]
    parts = planner.balance_by_cost([9, 1, 1, 1, 8, 2, 2, 2], 2)
    assert sorted(sum(parts, [])) == list(range(8))
    loads = [sum([9, 1, 1, 1, 8, 2, 2, 2][i] for i in p) for p in parts]
    assert abs(loads[0] - loads[1]) <= 1


def test_config_from_hf():
    from transformers import EsmConfig
    hf = EsmConfig(vocab_size=33, mask_token_id=32, pad_token_id=1, hidden_size=320, num_hidden_layers=6,
                   num_attention_heads=20, intermediate_size=1280, max_position_embeddings=1026, layer_norm_eps=1e-5,
                   position_embedding_type="rotary", token_dropout=True, emb_layer_norm_before=False)
    c = EncoderConfig.from_hf_config(hf)
    assert (c.hidden_size, c.head_dim, c.ffn_type, c.token_dropout, c.position_embedding_type) == (320, 16, "gelu", True, "rotary")
    sd = {"esm.encoder.layer.0.intermediate.dense.weight": torch.zeros(2 * 1280, 320)}
    assert EncoderConfig.from_hf_config(hf, sd).ffn_type == "glu"


def test_abi_library_exports_every_declared_symbol():
    from molly_b200 import _lib
    declared = _lib.declared_symbols()
    assert set(declared) == set(_lib.SIGNATURES), "ctypes table and include/molly_b200.h disagree"
    lib = _lib.load()                                        # raises if the .so was not built
    for name in declared:
        assert hasattr(lib, name), f"{name} not exported by libmolly_b200.so"
    assert lib.molly_abi_version() == 1
    assert lib.molly_kernel_launch_count() >= 0
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    assert set(declared) <= exported


def test_abi_rejects_bad_arguments_without_a_gpu():
    """Argument validation happens before any CUDA call, so it is testable on the CPU box."""
    from molly_b200 import _lib
    lib = _lib.load()
    h = ctypes.c_void_p()
    assert lib.molly_encoder_create(None, None, ctypes.byref(h)) == _lib.ERR_INVALID
    assert b"NULL" in lib.molly_last_error()
    cfg = _lib.EncoderConfig(hidden_size=100, num_layers=1, num_heads=3, intermediate_size=64, vocab_size=10)
    w = _lib.EncoderWeights()
    assert lib.molly_encoder_create(ctypes.byref(cfg), ctypes.byref(w), ctypes.byref(h)) == _lib.ERR_INVALID
    assert b"not divisible" in lib.molly_last_error()
    cfg = _lib.EncoderConfig(hidden_size=96, num_layers=1, num_heads=4, intermediate_size=64, vocab_size=10,
                             llm_hidden_size=64)
    assert lib.molly_encoder_create(ctypes.byref(cfg), ctypes.byref(w), ctypes.byref(h)) == _lib.ERR_UNSUPPORTED
    assert b"head_dim 24" in lib.molly_last_error()
    assert lib.molly_encoder_workspace_bytes(None, 4, 4) == 0


def test_product_does_not_import_oracle():
    """The product path must never route through the oracle (or any CPU fallback)."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for dirpath, _, files in os.walk(os.path.join(root, "molly_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f


def test_cpu_tensors_are_refused():
    from molly_b200 import ops
    with pytest.raises((RuntimeError, NotImplementedError)):
        ops.layernorm(torch.zeros(4, 64), torch.ones(64), torch.zeros(64), 1e-5)
    with pytest.raises((RuntimeError, NotImplementedError)):
        torch.ops.molly_b200.placeholder_scan(torch.zeros(2, 8, dtype=torch.int64), 1, 2, 3)


def test_route_run_index_skips_pad_slots():
    """Run index = rank among the sample's non-'pad' slots, whatever the slot's position."""
    infos = [[{"type": "dna", "start": 3}, {"type": "pad", "start": -1}, {"type": "protein", "start": 50}],
             [{"type": "protein", "start": 1}, {"type": "rna", "start": 9}, {"type": "pad", "start": -1}]]
    ids = [[torch.zeros(4, dtype=torch.long)] * 3] * 2
    nt, pr = planner.route(2, ids, infos)
    assert (nt.b_idx, nt.slot_idx, nt.run_idx) == ([0, 1], [0, 1], [0, 1])
    assert (pr.b_idx, pr.slot_idx, pr.run_idx) == ([0, 1], [2, 0], [1, 0])


def test_oracle_placeholder_runs_reproduce_dataset_starts():
    from oracle import cases, synth
    for case in cases.golden_cases().values():
        runs = synth.placeholder_runs(case.batch.input_ids, synth.PAD_TOKEN_IDS)
        for rr, infos in zip(runs, case.batch.omic_info_list):
            real = [i for i in infos if i["type"] != "pad"]
            assert [r[0] - 1 for r in rr] == [i["start"] for i in real]
            assert [("dna", "rna", "protein")[r[1]] for r in rr] == [i["type"] for i in real]


def test_bench_workloads_build_valid_batches():
    """Every bench workload's synthetic batch obeys the dataset's layout (start, K pads, end) and the collate shapes."""
    import bench
    for name, wl in bench.WORKLOADS.items():
        small = dict(wl, B=min(wl["B"], 6))
        omic_ids, infos = bench.make_inputs(small, seed=5)
        assert omic_ids.shape[0] == small["B"] and omic_ids.shape[2] == small["K"] and omic_ids.dtype == torch.int64
        rows, valid, flops = bench.batch_stats(small, omic_ids, infos)
        assert rows == small["K"] * sum(len(r) for r in infos) and 0 < valid <= rows and flops > 0
        ids = bench.build_input_ids(small, infos)
        runs = synth.placeholder_runs(ids, bench.PAD_TOKEN_IDS)
        for rr, row in zip(runs, infos):
            assert [r[0] - 1 for r in rr] == [i["start"] for i in row], name
            assert all(r[2] == small["K"] for r in rr)
        nt, pr = planner.route(small["B"], omic_ids, infos)
        for plan, vocab in ((nt, bench.ENC[wl["nt"]]["vocab_size"]), (pr, bench.ENC[wl["pr"]]["vocab_size"])):
            if len(plan):
                planner.check_vocab(planner.gather_ids(omic_ids, plan), vocab)
                planner.check_placement(plan, small["K"], small["B"], small["T"])


def test_ranges_disjoint_decides_concurrency():
    nt = planner.ModalityPlan([0, 1], [0, 0], [10, 5])
    pr = planner.ModalityPlan([0, 1], [1, 1], [60, -1])
    assert planner.ranges_disjoint(nt, 40, pr, 40)              # [11,51) vs [61,101); start -1 is skipped
    assert not planner.ranges_disjoint(nt, 40, planner.ModalityPlan([0], [1], [49]), 40)   # [11,51) vs [50,90)
    assert planner.ranges_disjoint(nt, 40, planner.ModalityPlan([1], [1], [10]), 40) is False   # sample 1: [6,46) vs [11,51)
    assert planner.ranges_disjoint(nt, 40, planner.ModalityPlan([], [], []), 40)


def test_balance_equal_count():
    costs = [64, 2048, 100, 1500, 90, 700, 1024, 80]
    parts = planner.balance_equal_count(costs, 2)
    assert sorted(parts[0] + parts[1]) == list(range(8)) and len(parts[0]) == len(parts[1]) == 4
    loads = [sum(costs[i] for i in p) for p in parts]
    assert abs(loads[0] - loads[1]) <= 0.15 * max(loads), loads
    naive = [sum(costs[:4]), sum(costs[4:])]
    assert abs(loads[0] - loads[1]) < abs(naive[0] - naive[1])
    with pytest.raises(ValueError):
        planner.balance_equal_count(costs[:7], 2)


def test_embedding_backward_metadata_matches_autograd():
    """train._emb_meta (scatter indices / scales of the embedding backward) against autograd of the oracle's embedding block:
    pad rows and <mask> rows receive nothing, token-dropout rescales per sequence, learned positions follow cumsum(mask)."""
    import types
    from molly_b200 import train
    from oracle.esm_oracle import SPECS, esm_embeddings, init_encoder_weights
    for spec_name in ("tiny_esm2", "tiny_ntv1"):
        spec = SPECS[spec_name]
        W = {k: v.clone().requires_grad_(v.dtype.is_floating_point) for k, v in init_encoder_weights(spec, 3).items()}
        g = torch.Generator().manual_seed(4)
        K = 24
        ids = torch.stack([synth.protein_ids(g, K, v) if spec.vocab_size == 33 else synth.nucleotide_ids(g, K, v, spec.vocab_size)
                           for v in (24, 9, 17)])
        if spec.token_dropout:
            ids[0, 3] = spec.mask_token_id
            ids[2, 5] = spec.mask_token_id
        d_x = torch.randn(ids.numel(), spec.hidden_size, generator=g)
        x = esm_embeddings(spec, W, ids, (ids != 1).long())
        (x.reshape(-1, spec.hidden_size) * d_x).sum().backward()
        enc = types.SimpleNamespace(cfg=EncoderConfig.from_mapping(spec.as_dict()))
        word_index, word_scale, pos_index, pos_scale = train._emb_meta(enc, ids)
        d_word = torch.zeros_like(W["esm.embeddings.word_embeddings.weight"])
        d_word.index_add_(0, word_index.long(), d_x * word_scale[:, None])
        assert torch.allclose(d_word, W["esm.embeddings.word_embeddings.weight"].grad, atol=1e-5)
        if pos_index is not None:
            d_pos = torch.zeros_like(W["esm.embeddings.position_embeddings.weight"])
            d_pos.index_add_(0, pos_index.long(), d_x * pos_scale[:, None])
            assert torch.allclose(d_pos, W["esm.embeddings.position_embeddings.weight"].grad, atol=1e-5)


@pytest.mark.timeout(300)
def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (runs on the CPU): one JSON line with the contract's keys, same metric / unit / config keys
    as the product arm, `impl: reference`, a cpu_baseline describing the run and a zero-copy e2e object."""
    import json
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--workload", "molly_mini", "--steps", "1",
                        "--warmup", "0"], cwd=root, capture_output=True, text=True, timeout=280)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    line = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["higher_is_better"] is True and line["vs_baseline"] is None
    assert line["metric"] == "omics tokens/sec (encode+project+merge)" and line["unit"] == "omics tokens/s"
    from oracle import ref_import                      # "reference": oracle/_ref/omics_one.py staged by build(); else the port
    want_kind = "reference" if ref_import.reference_available() else "port"
    assert line["value"] > 0 and line["cpu_baseline"]["kind"] == want_kind and line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    for key in ("workload", "B_per_gpu", "K", "T", "D", "parallelism", "l2"):
        assert key in line["config"], key


def test_planner_properties_randomised():
    """Property tests (hypothesis): ranges_disjoint == brute force over written rows; balance_equal_count is a partition with
    equal counts whose load spread never exceeds the largest single cost."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=200, deadline=None)
    @given(st.lists(st.tuples(st.integers(0, 3), st.integers(-1, 60)), max_size=6),
           st.lists(st.tuples(st.integers(0, 3), st.integers(-1, 60)), max_size=6), st.integers(1, 12), st.integers(1, 12))
    def disjoint(nt_items, pr_items, k_nt, k_pr):
        nt = planner.ModalityPlan([b for b, _ in nt_items], list(range(len(nt_items))), [s for _, s in nt_items])
        pr = planner.ModalityPlan([b for b, _ in pr_items], list(range(len(pr_items))), [s for _, s in pr_items])
        rows_nt = {(b, s + 1 + j) for b, s in nt_items if s != -1 for j in range(k_nt)}
        rows_pr = {(b, s + 1 + j) for b, s in pr_items if s != -1 for j in range(k_pr)}
        assert planner.ranges_disjoint(nt, k_nt, pr, k_pr) == (not (rows_nt & rows_pr))

    @settings(max_examples=200, deadline=None)
    @given(st.integers(1, 4), st.integers(1, 6), st.data())
    def balanced(world, per_rank, data):
        costs = data.draw(st.lists(st.floats(1.0, 4096.0), min_size=world * per_rank, max_size=world * per_rank))
        parts = planner.balance_equal_count(costs, world)
        assert sorted(i for p in parts for i in p) == list(range(len(costs)))
        assert all(len(p) == per_rank for p in parts)
        loads = [sum(costs[i] for i in p) for p in parts]
        assert max(loads) - min(loads) <= max(costs) + 1e-6

    disjoint()
    balanced()


def test_nt_v2_loader_variants_pack_what_the_module_computes():
    """The gated-FFN packing for both half orders and with / without biases, checked on the CPU against the formula of the
    module it replaces: ``x1, x2 = dense(x).split(F); silu(x1) * x2`` (NT-v2 remote code, `add_bias_fnn`), or the swapped
    order.  The FFN1 epilogue computes silu(even column) * odd column of ``x @ W_packed^T + b_packed``."""
    import types
    import torch
    from molly_b200.config import EncoderConfig
    from molly_b200.packing import glu_deinterleave, glu_interleave
    g = torch.Generator().manual_seed(0)
    h, F = 8, 6
    W, b, x = torch.randn(2 * F, h, generator=g), torch.randn(2 * F, generator=g), torch.randn(5, h, generator=g)
    u = x @ W.t() + b
    x1, x2 = u.split(F, dim=-1)
    for gate_first, want in ((True, torch.nn.functional.silu(x1) * x2), (False, torch.nn.functional.silu(x2) * x1)):
        acc = x @ glu_interleave(W, gate_first).t() + glu_interleave(b, gate_first)
        got = torch.nn.functional.silu(acc[:, 0::2]) * acc[:, 1::2]
        assert torch.allclose(got, want, atol=1e-6)
        assert torch.equal(glu_deinterleave(glu_interleave(W, gate_first), gate_first), W)       # gradients go back unchanged
        assert torch.equal(glu_deinterleave(glu_interleave(b, gate_first), gate_first), b)

    # key names + config flags of the remote code as documented on the model card: same module paths as stock ESM,
    # intermediate.dense is [2F, h], biases present iff add_bias_fnn
    def hf_cfg(**kw):
        base = dict(hidden_size=h, num_hidden_layers=1, num_attention_heads=2, intermediate_size=F, vocab_size=4107,
                    pad_token_id=1, mask_token_id=2, position_embedding_type="rotary", max_position_embeddings=2050,
                    token_dropout=False, emb_layer_norm_before=False, layer_norm_eps=1e-12)
        base.update(kw)
        return types.SimpleNamespace(**base)
    sd_nobias = {"esm.encoder.layer.0.intermediate.dense.weight": W}
    sd_bias = dict(sd_nobias, **{"esm.encoder.layer.0.intermediate.dense.bias": b})
    c = EncoderConfig.from_hf_config(hf_cfg(add_bias_fnn=False), sd_nobias)
    assert (c.ffn_type, c.ffn_bias, c.glu_gate_first) == ("glu", False, True)
    c = EncoderConfig.from_hf_config(hf_cfg(add_bias_fnn=True), sd_bias)                          # biased gated FFN
    assert (c.ffn_type, c.ffn_bias) == ("glu", True)
    c = EncoderConfig.from_hf_config(hf_cfg(glu_gate_first=False), sd_nobias)                      # explicit half order
    assert (c.ffn_type, c.glu_gate_first) == ("glu", False)
    import pytest
    with pytest.raises(ValueError):                                                                # config and checkpoint disagree
        EncoderConfig.from_hf_config(hf_cfg(add_bias_fnn=False), sd_bias)
    stock = EncoderConfig.from_hf_config(hf_cfg(), {"esm.encoder.layer.0.intermediate.dense.weight": W[:F],
                                                    "esm.encoder.layer.0.intermediate.dense.bias": b[:F]})
    assert (stock.ffn_type, stock.ffn_bias) == ("gelu", True)


def test_graphed_step_ownership_protocol():
    """``train._GraphedStep`` (buffers of the CUDA-graphed training step): one forward owns them until its backward releases
    them; a second forward in between is refused (it runs eagerly); a tape that dies frees them; a forward that is passed over
    more than ``_GRAPH_ABANDON`` times loses them and its late backward is detectable by its generation.  Pure host logic."""
    from molly_b200 import train

    class Tape:
        gen = 0

    gs = object.__new__(train._GraphedStep)                # no device buffers: only the ownership state
    gs._owner, gs.gen, gs._passed_over = None, 0, 0
    a, b = Tape(), Tape()
    assert gs.try_acquire(a) and a.gen == 1 and gs.busy
    assert not gs.try_acquire(b) and b.gen == 0            # re-entrant forward: refused, buffers untouched
    gs.release(a)
    assert not gs.busy
    assert gs.try_acquire(b) and b.gen == 2
    del b                                                  # its backward never comes and the tape dies
    assert not gs.busy
    c = Tape()
    assert gs.try_acquire(c) and c.gen == 3
    others = [Tape() for _ in range(train._GRAPH_ABANDON + 1)]
    assert [gs.try_acquire(t) for t in others] == [False] * train._GRAPH_ABANDON + [True]
    assert others[-1].gen == 4 and c.gen != gs.gen         # c was taken for abandoned: its backward must refuse (gen mismatch)
    gs.release(c)                                          # a stale release does not free the new owner
    assert gs.busy
    gs.release(others[-1])
    assert not gs.busy
