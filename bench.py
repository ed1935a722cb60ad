#!/usr/bin/env python
"""bench.py -- omics tokens/s of the encode -> project -> merge path (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

A "step" is ONE call of the reference-facing boundary ``process_omic_sequences`` over one synthetic batch of the named
workload (default: BASELINE.json configs[1], Molly-1.7B: ESM-2 650M + NT-v2 500M -> D=2048, B=64 samples x (1 DNA + 1
protein) x 1024 omics tokens, T=3072, bf16, random-init weights).  An "omics token" is one row of the [N, K] encoder
input (pad rows are computed and written by the reference, so they count).  Multi-GPU = sample sharding, no forward
collective: every rank runs the same per-GPU batch ("weak" scaling); value = tokens of all ranks / max-over-ranks time.

  value        : inputs (ids, hidden_states) already resident in HBM when the timed region starts
  e2e          : same call with omic_ids in pinned HOST memory (the inference caller, src/inference_lora.py:291), i.e. the
                 H2D of the ids + sequence table inside the timed region, strict error semantics (device flag read back)
                 and one merged row read back to the host every step.  hidden_states stays on the device on both sides of
                 the boundary exactly as in the reference (omics_one.py:164 -> :175).
  roofline     : all tcgen05 GEMM launches (the dominant kernel) of one extra profiled step, CUDA events per launch on the
                 launching stream, algorithmic FLOP / time vs the measured sustained bf16 peak
  cpu_baseline : the fp32 CPU oracle (port of the reference path) on a bounded sample, all host threads (rank 0, N=1)
  --impl reference : times that CPU port as the reference arm.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "omics tokens/sec (encode+project+merge)"
UNIT = "omics tokens/s"

# encoder shapes -- kept here so that the measured arm never imports oracle/
ENC = {
    "esm2_t6_8m": dict(hidden_size=320, num_hidden_layers=6, num_attention_heads=20, intermediate_size=1280,
                       vocab_size=33, mask_token_id=32, position_embedding_type="rotary", max_position_embeddings=1026,
                       ffn_type="gelu", token_dropout=True, layer_norm_eps=1e-5),
    "esm2_t33_650m": dict(hidden_size=1280, num_hidden_layers=33, num_attention_heads=20, intermediate_size=5120,
                          vocab_size=33, mask_token_id=32, position_embedding_type="rotary",
                          max_position_embeddings=1026, ffn_type="gelu", token_dropout=True, layer_norm_eps=1e-5),
    "nt_v2_50m": dict(hidden_size=512, num_hidden_layers=12, num_attention_heads=16, intermediate_size=2048,
                      vocab_size=4107, mask_token_id=2, position_embedding_type="rotary", max_position_embeddings=2050,
                      ffn_type="glu", token_dropout=False, layer_norm_eps=1e-12),
    "nt_v2_500m": dict(hidden_size=1024, num_hidden_layers=29, num_attention_heads=16, intermediate_size=4096,
                       vocab_size=4107, mask_token_id=2, position_embedding_type="rotary", max_position_embeddings=2050,
                       ffn_type="glu", token_dropout=False, layer_norm_eps=1e-12),
    "nt_v1_2p5b": dict(hidden_size=2560, num_hidden_layers=32, num_attention_heads=20, intermediate_size=10240,
                       vocab_size=4105, mask_token_id=2, position_embedding_type="absolute",
                       max_position_embeddings=1002, ffn_type="gelu", token_dropout=False, layer_norm_eps=1e-12),
}
WORKLOADS = {
    # BASELINE.json configs[1] -- the configuration the metric is quoted on (fits one GPU)
    "molly_1p7b": dict(desc="Molly-1.7B: ESM-2 650M + NT-v2 500M -> Qwen3-1.7B merge (D=2048), bf16",
                       nt="nt_v2_500m", pr="esm2_t33_650m", D=2048, B=64, K=1024, T=3072, valid=1024),
    # BASELINE.json configs[0] -- the reference's own CPU-runnable case (parity config; selectable for quick runs)
    "molly_mini": dict(desc="Molly-mini: ESM-2 t6-8M + NT-v2-50M -> Qwen3-0.6B merge (D=1024)",
                       nt="nt_v2_50m", pr="esm2_t6_8m", D=1024, B=4, K=512, T=2048, valid=512),
    # BASELINE.json configs[4] -- the path's share of one training step (--train-mlp): forward with the encoder output kept,
    # projector backward (dW, db of both modalities) from a given d(inputs_embeds), one flat grad all-reduce (mean)
    # BASELINE.json configs[2] -- variable lengths: one sequence per sample, kind uniform in {dna, rna, protein}, length
    # log-uniform in [64, 2048], K = 2048 (pad rows are computed and written like the reference; keys beyond the length are not)
    "molly_4b": dict(desc="Molly-4B: ESM-2 650M + NT-v2 500M -> Qwen3-4B merge (D=2560), mixed kinds, log-uniform lengths",
                     nt="nt_v2_500m", pr="esm2_t33_650m", D=2560, B=64, K=2048, T=4096, valid=2048, layout="mixed_varlen"),
    # BASELINE.json configs[3] -- the biggest GEMMs: NT-v1 2.5B (h=2560, F=10240, head_dim 128, learned positions), 1000-bp
    # DNA windows = 171 tokens in K=256 slots; the protein encoder is loaded but idle
    "molly_8b": dict(desc="Molly-8B: NT-v1 2.5B multispecies -> Qwen3-8B merge (D=4096), 256 x 1000-bp DNA windows",
                     nt="nt_v1_2p5b", pr="esm2_t6_8m", D=4096, B=256, K=256, T=512, valid=171, layout="dna_only"),
    "train_1p7b": dict(desc="Molly-1.7B train step, path only: fwd + projector bwd + grad all-reduce (B=8/GPU)",
                       nt="nt_v2_500m", pr="esm2_t33_650m", D=2048, B=8, K=1024, T=3072, valid=1024, train=True),
    # the same step with the encoders unfrozen (--train-bio): training forward with a tape + the full encoder backward
    "train_bio_1p7b": dict(desc="Molly-1.7B train step with trainable encoders (--train-bio), path only (B=8/GPU)",
                           nt="nt_v2_500m", pr="esm2_t33_650m", D=2048, B=8, K=1024, T=3072, valid=1024, train=True,
                           train_encoders=True),
}


def flops_per_token(e: dict, kv_len: int, D: int) -> float:
    """2*L*(4h^2 + g*h*F) + 4*L*kv_len*h + 2*h*D, LM head excluded."""
    h, L, F = e["hidden_size"], e["num_hidden_layers"], e["intermediate_size"]
    g = 3 if e["ffn_type"] == "glu" else 2
    return 2.0 * L * (4 * h * h + g * h * F) + 4.0 * L * kv_len * h + 2.0 * h * D


# ------------------------------------------------------------------------------------------------------------------
def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        with open(p) as f:
            d = json.load(f)
        return {"tflops_sustained": d.get("bf16_tflops_sustained"), "tflops_burst": d.get("bf16_tflops"),
                "hbm_gbs": d.get("hbm_gbs"), "source": "measured (MEASURED_PEAKS.json)"}
    return {"tflops_sustained": 1400.0, "tflops_burst": 1590.0, "hbm_gbs": 6650.0,
            "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, uuid: str):
        self.uuid, self.proc, self.lines = uuid, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", self.uuid], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); smax.append(float(parts[1])); pw.append(float(parts[2]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[3:7]):
                if v == "Active":
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------------------------
def make_inputs(wl: dict, seed: int = 1234):
    """Host-side synthetic batch shaped like the reference's dataset + collate: collated
    ``omic_ids [B, Nmax, K]`` (pad id 1) and ``omic_info_list`` with the dataset's ``start`` positions."""
    import math
    import torch
    g = torch.Generator().manual_seed(seed)
    B, K, T, valid = wl["B"], wl["K"], wl["T"], wl["valid"]
    nt_vocab = ENC[wl["nt"]]["vocab_size"]
    layout = wl.get("layout", "pair")

    def nt_seq(n):
        row = torch.ones(K, dtype=torch.int64)
        row[:n] = torch.randint(6, nt_vocab - 5, (n,), generator=g)             # k-mers, <cls>=3
        row[0] = 3
        return row

    def pr_seq(n):
        row = torch.ones(K, dtype=torch.int64)
        row[:n] = torch.randint(4, 24, (n,), generator=g)                       # residues, <cls>=0 <eos>=2
        row[0] = 0
        row[n - 1] = 2
        return row

    rows, infos = [], []
    for b in range(B):
        s0 = 20 + (b % 7)                                                       # text prefix before the first run
        if layout == "pair":                                                    # DNA run, 30 text tokens, protein run
            s1 = s0 + K + 2 + 30
            assert s1 + K + 2 <= T
            rows.append(torch.stack([nt_seq(valid), pr_seq(valid)]))
            infos.append([{"type": "dna", "start": s0}, {"type": "protein", "start": s1}])
        elif layout == "mixed_varlen":
            kind = ("dna", "rna", "protein")[int(torch.randint(0, 3, (1,), generator=g))]
            n = int(round(math.exp(float(torch.rand(1, generator=g)) * math.log(valid / 64.0)) * 64))
            rows.append(torch.stack([pr_seq(n) if kind == "protein" else nt_seq(n)]))
            infos.append([{"type": kind, "start": s0}])
        elif layout == "dna_only":
            rows.append(torch.stack([nt_seq(valid)]))
            infos.append([{"type": "dna", "start": s0}])
        else:
            raise ValueError(layout)
        assert infos[-1][-1]["start"] + K + 2 <= T
    return torch.stack(rows), infos


def batch_stats(wl: dict, omic_ids, infos):
    """(omics rows per step, valid tokens per step, model FLOP per step) of one rank's batch."""
    rows = valid = 0
    flops = 0.0
    for b, row in enumerate(infos):
        for i, info in enumerate(row):
            e = ENC[wl["pr"] if info["type"] == "protein" else wl["nt"]]
            n = int((omic_ids[b, i] != 1).sum())
            rows += wl["K"]
            valid += n
            flops += wl["K"] * flops_per_token(e, n, wl["D"])
    return rows, valid, flops


def sample_cost(wl: dict, ids_b, infos_b) -> float:
    """Model FLOP of one sample (balance by work, not by count) -- the weight of planner.balance_equal_count."""
    c = 0.0
    for i, info in enumerate(infos_b):
        if info["type"] == "pad":
            continue
        e = ENC[wl["pr"] if info["type"] == "protein" else wl["nt"]]
        c += wl["K"] * flops_per_token(e, int((ids_b[i] != 1).sum()), wl["D"])
    return c


LLM_VOCAB = 151936                                         # Qwen3 embedding rows (config.json vocab_size)
PLACEHOLDER_BASE = 151669                                  # first id after Qwen3's own specials: the 9 added omics tags
PAD_TOKEN_IDS = (PLACEHOLDER_BASE + 1, PLACEHOLDER_BASE + 4, PLACEHOLDER_BASE + 7)


def build_input_ids(wl: dict, infos, seed: int = 7):
    """Text tokens + ``x_start, x_pad * K, x_end`` at every info["start"] (the dataset's layout, omics_dataset.py:270-288)."""
    import torch
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(0, 150000, (wl["B"], wl["T"]), generator=g)
    base = {"dna": 0, "rna": 3, "protein": 6}
    for b, row in enumerate(infos):
        for info in row:
            s, o = info["start"], PLACEHOLDER_BASE + base[info["type"]]
            ids[b, s] = o
            ids[b, s + 1:s + 1 + wl["K"]] = o + 1
            ids[b, s + 1 + wl["K"]] = o + 2
    return ids


def gpu_state_dict(e: dict, device, seed: int):
    """Random-init weights of the named architecture with EsmForMaskedLM.state_dict() key names, built on the GPU."""
    import torch
    g = torch.Generator(device=device).manual_seed(seed)
    h, L, F = e["hidden_size"], e["num_hidden_layers"], e["intermediate_size"]
    rn = lambda *s, std=0.02: torch.randn(*s, generator=g, device=device, dtype=torch.float32) * std
    sd = {"esm.embeddings.word_embeddings.weight": rn(e["vocab_size"], h)}
    sd["esm.embeddings.word_embeddings.weight"][1].zero_()
    if e["position_embedding_type"] == "absolute":
        sd["esm.embeddings.position_embeddings.weight"] = rn(e["max_position_embeddings"], h)
    for i in range(L):
        p = f"esm.encoder.layer.{i}."
        sd[p + "attention.LayerNorm.weight"] = 1 + rn(h, std=0.1)
        sd[p + "attention.LayerNorm.bias"] = rn(h, std=0.05)
        for nm in ("query", "key", "value"):
            sd[p + f"attention.self.{nm}.weight"] = rn(h, h)
            sd[p + f"attention.self.{nm}.bias"] = rn(h)
        sd[p + "attention.output.dense.weight"] = rn(h, h)
        sd[p + "attention.output.dense.bias"] = rn(h)
        sd[p + "LayerNorm.weight"] = 1 + rn(h, std=0.1)
        sd[p + "LayerNorm.bias"] = rn(h, std=0.05)
        if e["ffn_type"] == "glu":
            sd[p + "intermediate.dense.weight"] = rn(2 * F, h)
            sd[p + "output.dense.weight"] = rn(h, F)
        else:
            sd[p + "intermediate.dense.weight"] = rn(F, h)
            sd[p + "intermediate.dense.bias"] = rn(F)
            sd[p + "output.dense.weight"] = rn(h, F)
            sd[p + "output.dense.bias"] = rn(h)
    sd["esm.encoder.emb_layer_norm_after.weight"] = 1 + rn(h, std=0.1)
    sd["esm.encoder.emb_layer_norm_after.bias"] = rn(h, std=0.05)
    return sd


def build_path(wl: dict, device, strict: bool):
    import torch
    from molly_b200.config import EncoderConfig
    from molly_b200.omics_path import FastOmicsPath
    from molly_b200.packing import PackedEncoder
    encs = []
    for i, key in enumerate(("nt", "pr")):
        e = ENC[wl[key]]
        sd = gpu_state_dict(e, device, 10 + i)
        proj = {"weight": torch.randn(wl["D"], e["hidden_size"], device=device) / e["hidden_size"] ** 0.5,
                "bias": torch.randn(wl["D"], device=device) * 0.02}
        encs.append(PackedEncoder(EncoderConfig.from_mapping(dict(e, name=wl[key])), sd, proj, wl["K"], device,
                                  rope_len=max(4096, wl["K"])))
        del sd
    torch.cuda.empty_cache()
    return FastOmicsPath(encs[0], encs[1], strict=strict)


# ------------------------------------------------------------------------------------------------------------------
def cpu_reference_run(wl: dict, steps: int, warmup: int, budget_s: float):
    """Times the reference's CPU path on a bounded sample of the workload: 1 sample (1 DNA + 1 protein sequence, or what the
    layout has), fp32, all host threads.  kind "reference": the reference's OWN ``OmicsOne.process_omic_sequences``
    (oracle/_ref/omics_one.py, an unmodified copy staged by build()) driving stock HF ``EsmForMaskedLM`` modules (eager
    attention; NT-v2's gated FFN through the oracle's HF subclass) -- when that copy is absent, kind "port": the fp32 oracle
    restatement.  Returns (tokens/s, ms/step, sample text, cores, kind)."""
    import torch
    from oracle import ref_import
    from oracle.esm_oracle import SPECS, OracleModality, init_encoder_weights, init_projector, process_omic_sequences
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    nts, prs = SPECS[wl["nt"]], SPECS[wl["pr"]]
    D = wl["D"]
    kind = "reference" if ref_import.reference_available() else "port"
    if kind == "reference":
        om = ref_import.build_reference_omics_random(nts, prs, D, wl["K"])
    else:
        nt = OracleModality(nts, init_encoder_weights(nts, 11, bf16_exact=False), init_projector(nts.hidden_size, D, 12), wl["K"])
        pr = OracleModality(prs, init_encoder_weights(prs, 13, bf16_exact=False), init_projector(prs.hidden_size, D, 14), wl["K"])

    def one(k_tokens: int) -> float:
        sub = dict(wl, B=1, K=k_tokens, valid=min(wl["valid"], k_tokens), T=2 * k_tokens + 128)
        ids, infos = make_inputs(sub)
        one.n_seq = sum(len(r) for r in infos)
        hs = torch.zeros(1, sub["T"], D)
        t0 = time.perf_counter()
        with torch.no_grad():
            if kind == "reference":
                om.dna_rna_project_token_num = om.protein_project_token_num = k_tokens
                om.process_omic_sequences(hs, ids, infos, hs.device)
            else:
                nt.project_token_num = pr.project_token_num = k_tokens
                process_omic_sequences(hs, ids, infos, nt, pr)
        return time.perf_counter() - t0

    one(64)                                                    # thread-pool / allocator warm-up
    probe_k = min(128, wl["K"])
    t_probe = one(probe_k)
    k_tokens = wl["K"]
    while k_tokens > probe_k and (steps + warmup) * t_probe * (k_tokens / probe_k) * 1.3 > budget_s:
        k_tokens //= 2
    for _ in range(warmup):
        one(k_tokens)
    times = [one(k_tokens) for _ in range(steps)]
    total = sum(times)
    tokens = one.n_seq * k_tokens * steps
    what = ("the reference's own OmicsOne.process_omic_sequences + stock HF EsmForMaskedLM (eager), fp32 CPU" if kind == "reference"
            else "fp32 CPU oracle port")
    sample = (f"{steps} step(s) x 1 sample ({one.n_seq} sequence(s) x {k_tokens} tokens) of {wl['desc']}; {what}, "
              f"{cores} torch threads")
    return tokens / total, 1e3 * total / steps, sample, cores, kind


def l2_note(wl: dict, n_seqs: int) -> str:
    """Timing rule: inputs larger than L2, or a flush between iterations -- say which."""
    weights = sum(2.0 * ENC[k]["num_hidden_layers"] * (4 * ENC[k]["hidden_size"] ** 2 + (3 if ENC[k]["ffn_type"] == "glu" else 2)
                                                        * ENC[k]["hidden_size"] * ENC[k]["intermediate_size"]) for k in (wl["nt"], wl["pr"]))
    acts = n_seqs * wl["K"] * max(ENC[wl["nt"]]["hidden_size"], ENC[wl["pr"]]["hidden_size"]) * 12.0   # fp32 x, bf16 ln, qkv, attn
    hidden = wl["B"] * wl["T"] * wl["D"] * 2.0
    total = weights + acts + hidden
    if total > 4 * 126e6:
        return (f"per-step working set {total / 1e9:.1f} GB (weights {weights / 1e9:.1f} + activations {acts / 1e9:.1f} + "
                f"hidden_states {hidden / 1e9:.1f}) is far larger than the 126 MB L2: no flush needed")
    return (f"per-step working set {total / 1e6:.0f} MB is comparable to the 126 MB L2 and steps run back to back (warm L2): "
            "not a headline configuration")


def make_config(wl: dict, world: int, n_seqs: int, valid_per_step: int) -> dict:
    """The `config` object of the bench line (both arms print the same one)."""
    return {"workload": wl["desc"], "B_per_gpu": wl["B"], "layout": wl.get("layout", "pair: 1 dna + 1 protein per sample"),
            "seqs_per_gpu": n_seqs, "K": wl["K"], "valid_tokens_per_gpu": valid_per_step, "T": wl["T"], "D": wl["D"],
            "parallelism": f"sample-sharded x{world}",
            "l2": l2_note(wl, n_seqs),
            "residual_stream": "fp32", "pad_rows": "computed and written (reference-exact)"}


def run_reference_arm(args, wl: dict, rank: int, world: int) -> None:
    if rank != 0:
        return
    omic_ids, infos = make_inputs(wl, seed=1234)
    _, valid_per_step, _ = batch_stats(wl, omic_ids, infos)
    config = make_config(wl, world, sum(len(r) for r in infos), valid_per_step)
    config["note"] = ("reference arm = the reference's own CPU path on a bounded sample of this workload (its process_omic_sequences, "
                      "unmodified, on stock HF modules when oracle/_ref/ is staged; else the fp32 oracle port); the reference "
                      "ships no GPU kernel of its own")
    tps, ms, sample, cores, kind = cpu_reference_run(wl, args.steps, args.warmup, budget_s=200.0)
    line = {"impl": "reference", "metric": METRIC, "value": tps, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": config,
            "cpu_baseline": {"value": tps, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": tps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------------
class ParamBag:
    """Duck-typed stand-in for an ``EsmForMaskedLM``: live bf16 parameters under their HF ``state_dict`` names."""

    def __init__(self, state_dict):
        import torch
        self._p = {k: torch.nn.Parameter(v.to(torch.bfloat16)) for k, v in state_dict.items()}

    def named_parameters(self):
        return list(self._p.items())

    def parameters(self):
        return list(self._p.values())

    def state_dict(self):
        return {k: v.detach() for k, v in self._p.items()}


LORA_ELEMS_QWEN3_1P7B = 28 * 64 * (2 * (2048 + 2048) + 2 * (2048 + 1024) + 3 * (2048 + 6144))   # r=64 on every linear: 69 730 304


def train_section(args, wl: dict, path, dev, rank: int, world: int, steps: int, warmup: int) -> dict:
    """cfg-5: what the path contributes to one training step.  The LLM's own forward/backward is out of
    scope; its products -- d(loss)/d(inputs_embeds) and the LoRA adapter gradients (pre_train_lora, r=64 on every Qwen3
    linear: src/utils/tools.py:345-396) -- are fixed synthetic bf16 tensors.  Communication = what DeepSpeed ZeRO-0/2 does
    for the reference (src/configs/ds_z0_config.json:18-27): a mean all-reduce of every trainable gradient.  Here the LoRA
    bucket (69.7 M elements, complete when the LLM backward ends) is all-reduced on a side stream UNDER the path's backward,
    the projector bucket (4.7 M elements, the last gradients of the step) right after it."""
    import torch
    import torch.distributed as dist
    from molly_b200 import ops
    from molly_b200.dist import FlatGradBucket
    saved = (path._proj_modules, dict(path._enc_modules), dict(path._enc_versions), path.grad_reducer, path.concurrent)
    projs = {}
    for name, enc in (("dna_rna", path.dna_rna), ("protein", path.protein)):     # live nn.Linear modules, as in OmicsOne
        lin = torch.nn.Linear(enc.proj_w.shape[1], enc.proj_w.shape[0], device=dev, dtype=torch.bfloat16)
        with torch.no_grad():
            lin.weight.copy_(enc.proj_w)
            lin.bias.copy_(enc.proj_b)
        projs[name] = lin
    path._proj_modules = projs
    params = [p for lin in projs.values() for p in lin.parameters()]
    if wl.get("train_encoders"):
        for i, (name, key) in enumerate((("dna_rna", "nt"), ("protein", "pr"))):
            bag = ParamBag(gpu_state_dict(ENC[wl[key]], dev, 10 + i))       # same seeds as build_path: same weights
            path._enc_modules[name] = bag
            path._enc_versions[name] = path._module_version(bag)
            params += bag.parameters()
    overlap = bool(wl.get("train_encoders")) and world > 1 and os.environ.get("MOLLY_BENCH_FLAT_BUCKET", "0") != "1"
    bucket = lora_bucket = None
    lora = None
    if not wl.get("train_encoders"):          # cfg-5 proper: --train-mlp + LoRA
        lora = torch.nn.Parameter(torch.zeros(LORA_ELEMS_QWEN3_1P7B, device=dev, dtype=torch.bfloat16))
        lora.grad = (torch.randn(LORA_ELEMS_QWEN3_1P7B, device=dev) * 1e-3).to(torch.bfloat16)
    if overlap:                                # gradients are averaged layer by layer under the backward: no flat bucket
        from molly_b200.dist import LayerwiseGradReducer
        path.grad_reducer = LayerwiseGradReducer()
    elif world > 1:
        bucket = FlatGradBucket(params)
        if lora is not None:
            lora_bucket = FlatGradBucket([lora])
    omic_ids, infos = make_inputs(wl, seed=1234 + rank)
    omic_ids_dev = omic_ids.to(dev)
    base = (torch.randn(wl["B"], wl["T"], wl["D"], device=dev) * 0.02).to(torch.bfloat16)
    d_out = (torch.randn(wl["B"], wl["T"], wl["D"], device=dev) * 1e-3).to(torch.bfloat16)
    tokens_per_step = wl["B"] * 2 * wl["K"]

    def step(comm: bool = True):
        for p in params:                       # optimizer.zero_grad() (set_to_none=True, the PyTorch / HF Trainer default)
            p.grad = None
        hs = base.clone()
        out = path.process_omic_sequences(hs, omic_ids_dev, infos, dev)
        # (the LLM forward + backward run here in the real step: they leave d_out and the LoRA gradients)
        if comm and lora_bucket is not None:
            lora_bucket.launch()               # under the path's backward
        out.backward(d_out)
        if comm and bucket is not None:
            bucket.launch()
            bucket.finish()
        if comm and lora_bucket is not None:
            lora_bucket.finish()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record()
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
        host_ms[0] = (time.perf_counter() - t0) * 1e3 / n          # how long the host needs to enqueue one step
        ev1.record()
        barrier()
        ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()) / n

    host_ms = [0.0]
    for _ in range(max(warmup, 6)):            # the caching allocator needs a few steps to settle on the tape's block sizes
        step()
    assert all(p.grad is not None and torch.isfinite(p.grad.float()).all() for p in params)
    props = torch.cuda.get_device_properties(dev)
    sampler = ClockSampler("GPU-" + str(props.uuid))
    l0 = ops.kernel_launch_count()
    sampler.start()
    ms_step = timed(step, steps)
    host_enqueue_ms = host_ms[0]
    clocks = sampler.stop()
    launches = ops.kernel_launch_count() - l0
    comm = None
    if bucket is not None or overlap:
        ms_nocomm = timed(lambda: step(False), steps) if not overlap else None
        comm = {"buckets": [], "nccl_version": ".".join(str(v) for v in torch.cuda.nccl.version())}
        if bucket is not None:
            def allreduce_only():
                if lora_bucket is not None:
                    lora_bucket.launch()
                bucket.launch()
                bucket.finish()
                if lora_bucket is not None:
                    lora_bucket.finish()
            allreduce_only()
            comm["allreduce_ms"] = round(timed(allreduce_only, max(steps, 5)), 4)
            comm["buckets"].append({"what": "projector weight + bias, both modalities", "elements": bucket.numel,
                                    "bytes": bucket.numel * 2})
            if lora_bucket is not None:
                comm["buckets"].append({"what": "LoRA adapters r=64 on every Qwen3-1.7B linear (synthetic gradients)",
                                        "elements": lora_bucket.numel, "bytes": lora_bucket.numel * 2})
            comm["bytes"] = sum(b["bytes"] for b in comm["buckets"])
            comm["allreduce_bus_gbs"] = round(comm["bytes"] * 2 * (world - 1) / world / (comm["allreduce_ms"] * 1e-3) / 1e9, 1)
            comm["ms_per_step_without_allreduce"] = round(ms_nocomm, 4)
            comm["allreduce_exposed_ms"] = round(ms_step - ms_nocomm, 4)
            comm["allreduce_exposed_frac"] = round((ms_step - ms_nocomm) / ms_step, 4)
            comm["schedule"] = ("LoRA bucket launched when the (skipped) LLM backward ends, on a side stream under the path's "
                                "backward; projector bucket after the projector backward; both mean all-reduces, NCCL")
        else:
            comm["bytes"] = path.grad_reducer.bytes_reduced // max(1, max(warmup, 6) + steps)
            comm["schedule"] = "layer-wise all-reduce on a side stream under the encoder backward"
    path.concurrent = False                                       # per-launch events need one stream
    ops.profile_start()
    step(False)
    torch.cuda.synchronize()
    prof = ops.profile_stop()
    kernels = {k: {"launches": v["launches"], "ms": round(v["ms"], 3)} for k, v in prof.items()}
    grad_elems = sum(p.numel() for p in params) + (lora.numel() if lora is not None else 0)
    out = {"workload": wl["desc"], "B_per_gpu": wl["B"], "K": wl["K"], "T": wl["T"], "D": wl["D"],
           "tokens_per_s": world * tokens_per_step / (ms_step * 1e-3), "ms_per_step": ms_step,
           "host_enqueue_ms_per_step": round(host_enqueue_ms, 3),
           "parallelism": (f"sample-sharded x{world}, " + ("layer-wise grad all-reduce overlapped with the backward"
                                                            if overlap else "flat grad buckets, mean all-reduce")
                           + f" ({grad_elems} elements)"),
           "trainable": ("both projectors and every encoder parameter (--train-bio)" if wl.get("train_encoders")
                         else "both projectors (weight + bias) + LoRA-sized adapter gradients; encoders frozen"),
           "comm": comm, "clocks": clocks, "gpu_launches": launches, "kernels": kernels}
    path._proj_modules, path._enc_modules, path._enc_versions, path.grad_reducer, path.concurrent = saved
    return out


def run_train_step(args, wl: dict, dev, rank: int, world: int, warmup: int) -> None:
    """``--workload train_1p7b | train_bio_1p7b``: the training section as its own bench line."""
    path = build_path(wl, dev, strict=False)
    sec = train_section(args, wl, path, dev, rank, world, args.steps, warmup)
    line = {"metric": METRIC, "value": sec["tokens_per_s"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warmup,
            "ms_per_step": sec["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {k: sec[k] for k in ("workload", "B_per_gpu", "K", "T", "D", "parallelism", "trainable")},
            "host_enqueue_ms_per_step": sec["host_enqueue_ms_per_step"], "clocks": sec["clocks"], "gpu_launches": sec["gpu_launches"], "kernels": sec["kernels"], "comm": sec["comm"]}
    if rank == 0:
        print(json.dumps(line), flush=True)
    path.close()


def varlen_section(args, dev, rank: int, world: int, steps: int, warmup: int) -> dict:
    """cfg-3 (BASELINE.json configs[2]): one global batch of mixed dna / rna / protein sequences with log-uniform lengths,
    dealt to the ranks by ``planner.balance_equal_count`` (equal sample counts, balanced attention cost; host-side only --
    SURVEY.md 8e).  Reports the per-rank step times: their spread is the imbalance the planner left."""
    import torch
    import torch.distributed as dist
    from molly_b200 import planner
    wl = WORKLOADS["molly_4b"]
    path = build_path(wl, dev, strict=False)
    g_ids, g_infos = make_inputs(dict(wl, B=wl["B"] * world), seed=1234)
    # cost of a sample = model FLOP of its sequences: K rows through the GEMMs of ITS encoder (ESM-2 650M is 1.34x NT-v2 500M
    # per row; pad rows are computed like the reference) + attention over its valid keys
    costs = [sample_cost(wl, g_ids[b], g_infos[b]) for b in range(g_ids.shape[0])]
    mine = planner.balance_equal_count(costs, world)[rank]
    omic_ids, infos = g_ids[mine].contiguous(), [g_infos[b] for b in mine]
    omic_ids_dev = omic_ids.to(dev)
    hs = (torch.randn(wl["B"], wl["T"], wl["D"], device=dev) * 0.02).to(torch.bfloat16)
    tokens, valid, flops = batch_stats(wl, omic_ids, infos)
    fn = lambda: path.process_omic_sequences(hs, omic_ids_dev, infos, dev)
    for _ in range(warmup):
        fn()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev0.record()
    for _ in range(steps):
        fn()
    ev1.record()
    torch.cuda.synchronize()
    mine_ms = ev0.elapsed_time(ev1) / steps
    stats = torch.tensor([mine_ms, float(tokens), float(valid), flops, sum(costs[b] for b in mine)], device=dev,
                         dtype=torch.float64)
    allst = [torch.zeros_like(stats) for _ in range(world)]
    if world > 1:
        dist.all_gather(allst, stats)
    else:
        allst = [stats]
    ms = [float(t[0]) for t in allst]
    path.close()
    del path
    torch.cuda.empty_cache()
    return {"workload": wl["desc"], "B_per_gpu": wl["B"], "K": wl["K"], "T": wl["T"], "D": wl["D"],
            "split": "planner.balance_equal_count over the global batch (equal sample counts, balanced model FLOP)",
            "ms_per_step_max": round(max(ms), 3), "ms_per_step_min": round(min(ms), 3),
            "rank_imbalance": round(max(ms) / min(ms) - 1.0, 4),
            "tokens_per_s": round(sum(float(t[1]) for t in allst) / (max(ms) * 1e-3), 1),
            "valid_tokens_per_s": round(sum(float(t[2]) for t in allst) / (max(ms) * 1e-3), 1),
            "model_tflops_per_gpu": round(sum(float(t[3]) for t in allst) / world / (max(ms) * 1e-3) / 1e12, 1),
            "model_tflop_per_rank": [round(float(t[4]) / 1e12, 2) for t in allst]}


def gpu_library_baseline(dev, path_cfg2, steps: int = 5, warmup: int = 2) -> dict:
    """"also time the reference on GPU in bf16 -- that is the real bar".  What the reference executes per modality
    is stock HF ``EsmForMaskedLM`` (LM head included: it is computed and discarded, omics_one.py:83-91) + ``nn.Linear``
    (:92).  Timed here with random-init weights in bf16 on the same GPU, SDPA and eager attention, CUDA-event median,
    next to this repo's path on the same ids: the protein half of cfg-2 (ESM-2 650M, 64 x 1024) and the DNA half of cfg-4
    (NT-v1 2.5B, 256 x 256, 171 valid).  NT-v2 is hub remote code (not on disk), so it has no library arm."""
    import torch
    from transformers import EsmConfig, EsmForMaskedLM

    def med(fn):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(steps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        return statistics.median(ts)

    def one(wl_name: str, key: str, kind: str, path) -> dict:
        wl = WORKLOADS[wl_name]
        e = ENC[wl[key]]
        n_seq, K, D = wl["B"], wl["K"], wl["D"]
        omic_ids, infos = make_inputs(wl)
        col = [i for i, info in enumerate(infos[0]) if (info["type"] == "protein") == (kind == "protein")][0]
        ids = omic_ids[:, col].contiguous().to(dev)
        res = {"encoder": wl[key], "n_seq": n_seq, "K": K, "D": D, "tokens": n_seq * K}
        proj = torch.nn.Linear(e["hidden_size"], D, device=dev, dtype=torch.bfloat16)
        for impl in ("sdpa", "eager"):
            try:
                cfg = EsmConfig(vocab_size=e["vocab_size"], hidden_size=e["hidden_size"], num_hidden_layers=e["num_hidden_layers"],
                                num_attention_heads=e["num_attention_heads"], intermediate_size=e["intermediate_size"],
                                position_embedding_type=e["position_embedding_type"], token_dropout=e["token_dropout"],
                                mask_token_id=e["mask_token_id"], pad_token_id=1, layer_norm_eps=e["layer_norm_eps"],
                                max_position_embeddings=e["max_position_embeddings"], emb_layer_norm_before=False,
                                attn_implementation=impl)
                with torch.device(dev):
                    model = EsmForMaskedLM(cfg).to(torch.bfloat16).eval()
                mask = ids != 1

                @torch.no_grad()
                def ref_step():                                     # omics_one.py:83-91 + :92
                    o = model(input_ids=ids, attention_mask=mask, output_hidden_states=True)
                    return proj(o.hidden_states[-1])

                ms = med(ref_step)
                res[impl] = {"ms": round(ms, 3), "tokens_per_s": round(n_seq * K / ms * 1e3, 1)}
                del model, ref_step
            except Exception as ex:                                 # reported, never required
                res[impl] = {"ms": None, "tokens_per_s": None, "error": f"{type(ex).__name__}: {str(ex)[:120]}"}
            torch.cuda.empty_cache()
        hs = torch.zeros(n_seq, K + 8, D, dtype=torch.bfloat16, device=dev)
        one_infos = [[{"type": kind, "start": 2}] for _ in range(n_seq)]
        ids3 = ids.view(n_seq, 1, K)
        ms = med(lambda: path.process_omic_sequences(hs, ids3, one_infos, dev))
        res["ours"] = {"ms": round(ms, 3), "tokens_per_s": round(n_seq * K / ms * 1e3, 1)}
        for impl in ("sdpa", "eager"):
            if res[impl]["ms"]:
                res[f"ratio_vs_{impl}"] = round(res[impl]["ms"] / ms, 3)
        return res

    out = {"what": ("stock HF EsmForMaskedLM (LM head computed and discarded, like the reference) + nn.Linear, bf16, random "
                    "init, same GPU, CUDA-event median; 'ours' = this repo's process_omic_sequences on the same ids; "
                    "top-level keys = the protein half of cfg-2")}
    out.update(one("molly_1p7b", "pr", "protein", path_cfg2))
    try:
        path8 = build_path(WORKLOADS["molly_8b"], dev, strict=False)
        out["cfg4_nt_v1_2p5b"] = one("molly_8b", "nt", "dna", path8)
        path8.close()
        del path8
    except Exception as ex:
        out["cfg4_nt_v1_2p5b"] = {"error": f"{type(ex).__name__}: {str(ex)[:160]}"}
    torch.cuda.empty_cache()
    return out


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="molly_1p7b", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--headline-only", action="store_true",
                    help="skip the train_step / varlen / gpu_library_baseline sections of the default workload (profiling runs)")
    ap.add_argument("--cpu-budget-s", type=float, default=25.0)
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        run_reference_arm(args, wl, rank, world)
        return

    import torch
    import torch.distributed as dist
    from molly_b200 import _lib, ops
    _lib.load()                                                   # fail loudly: no CPU fallback
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    warmup = max(args.warmup, 3)

    if wl.get("train"):
        run_train_step(args, wl, dev, rank, world, warmup)
        if world > 1:
            dist.destroy_process_group()
        return
    path = build_path(wl, dev, strict=False)
    if wl.get("layout") == "mixed_varlen" and world > 1:
        # one global batch, samples dealt to ranks with equal counts and balanced attention cost
        from molly_b200 import planner
        g_ids, g_infos = make_inputs(dict(wl, B=wl["B"] * world), seed=1234)
        costs = [sample_cost(wl, g_ids[b], g_infos[b]) for b in range(g_ids.shape[0])]
        mine = planner.balance_equal_count(costs, world)[rank]
        omic_ids, infos = g_ids[mine].contiguous(), [g_infos[b] for b in mine]
    else:
        omic_ids, infos = make_inputs(wl, seed=1234 + rank)
    omic_ids_dev = omic_ids.to(dev)
    omic_ids_pinned = omic_ids.pin_memory()
    hs = (torch.randn(wl["B"], wl["T"], wl["D"], device=dev) * 0.02).to(torch.bfloat16)
    tokens_per_step, valid_per_step, model_flops = batch_stats(wl, omic_ids, infos)
    n_seqs = sum(len(r) for r in infos)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record()
        for _ in range(steps):
            fn()
        ev1.record()
        barrier()
        ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # ---- device-resident arm
    step_dev = lambda: path.process_omic_sequences(hs, omic_ids_dev, infos, dev)
    for _ in range(warmup):
        step_dev()
    ops.check_device_errors(dev)
    props = torch.cuda.get_device_properties(dev)
    sampler = ClockSampler("GPU-" + str(props.uuid))
    launches0 = ops.kernel_launch_count()
    sampler.start()
    total_ms = timed(step_dev, args.steps)
    clocks = sampler.stop()
    launches = ops.kernel_launch_count() - launches0
    value = world * tokens_per_step * args.steps / (total_ms / 1e3)

    # ---- end-to-end arm (the contract's `e2e`): ids from pinned HOST memory, strict error semantics (device flag read back),
    #      and the WHOLE merged [B, T, D] tensor copied back to pinned host memory every step
    path.strict = True
    b0, t0 = 0, infos[0][0]["start"] + 1
    h2d = omic_ids.numel() * 8 + n_seqs * 2 * 4                      # ids + the (b, start) tables
    host_out = torch.empty(hs.shape, dtype=hs.dtype, pin_memory=True)

    def step_e2e():
        path.process_omic_sequences(hs, omic_ids_pinned, infos, dev)
        host_out.copy_(hs, non_blocking=True)

    for _ in range(2):
        step_e2e()
    e2e_ms = timed(step_e2e, args.steps)
    e2e_value = world * tokens_per_step * args.steps / (e2e_ms / 1e3)
    d2h = hs.numel() * hs.element_size() + 2 * 4
    del host_out
    # the same call under the reference's own contract: inputs_embeds stays on the device for the LLM (omics_one.py:164 ->
    # :175), only one merged row + the error flag come back
    def step_resident():
        path.process_omic_sequences(hs, omic_ids_pinned, infos, dev)
        return hs[b0, t0].cpu()

    step_resident()
    res_ms = timed(step_resident, args.steps)
    e2e_resident = {"value": world * tokens_per_step * args.steps / (res_ms / 1e3), "unit": UNIT,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": wl["D"] * 2 + 2 * 4, "ms_per_step": res_ms / args.steps,
                    "note": "inputs_embeds left on the device as the reference does; one merged row + error flag read back"}
    path.strict = False

    # ---- embed_tokens(input_ids) fused with the path (run scan on the device is the index source)
    input_ids = build_input_ids(wl, infos).to(dev)
    table = (torch.randn(LLM_VOCAB, wl["D"], device=dev) * 0.02).to(torch.bfloat16)
    two_step = lambda: path.process_omic_sequences(torch.nn.functional.embedding(input_ids, table), omic_ids_dev, infos, dev)
    fused = lambda: path.embed_and_process(input_ids, table, omic_ids_dev, infos, PAD_TOKEN_IDS)
    for _ in range(2):
        a, b = two_step(), fused()
    n1_equal = bool(torch.equal(a, b))
    del a, b
    two_ms = timed(two_step, args.steps) / args.steps
    fused_ms = timed(fused, args.steps) / args.steps
    text_rows = wl["B"] * wl["T"] - tokens_per_step
    # the lookup alone (the step-level difference is below run-to-run noise): torch's gather vs scan + skipping gather
    slots = torch.tensor([len(r) for r in infos], dtype=torch.int32, device=dev)
    max_runs = max(len(r) for r in infos)

    def ours_lookup():
        runs = ops.placeholder_runs(input_ids, PAD_TOKEN_IDS, slots, max_runs)
        return ops.embed_tokens_skip(input_ids, runs[4], PAD_TOKEN_IDS, wl["K"], wl["K"], table)

    lookup_torch_ms = timed(lambda: torch.nn.functional.embedding(input_ids, table), 20) / 20
    lookup_ours_ms = timed(ours_lookup, 20) / 20
    input_fusion = {"two_step_ms": round(two_ms, 3), "fused_ms": round(fused_ms, 3), "bit_identical": n1_equal,
                    "tokens_per_s_fused": round(world * tokens_per_step / (fused_ms * 1e-3), 1),
                    "lookup_ms_torch": round(lookup_torch_ms, 4), "lookup_ms_fused": round(lookup_ours_ms, 4),
                    "lookup_bytes_torch": 2 * wl["B"] * wl["T"] * wl["D"] * 2,
                    "lookup_bytes_fused": 2 * text_rows * wl["D"] * 2,
                    "lookup_gbs_fused": round(2 * text_rows * wl["D"] * 2 / (lookup_ours_ms * 1e-3) / 1e9, 1)}
    del table

    # ---- one profiled step: per-launch CUDA events on the launching stream -> roofline of the dominant kernel
    peaks = measured_peaks()
    path.concurrent = False                                       # per-launch events need one stream
    ops.profile_start()
    step_dev()
    torch.cuda.synchronize()
    prof = ops.profile_stop()
    tot_ms = sum(v["ms"] for v in prof.values()) or 1.0
    kernels = {}
    for name, v in prof.items():
        rate = v["work"] / (v["ms"] * 1e-3) if v["ms"] > 0 else 0.0
        kernels[name] = {"launches": v["launches"], "ms": round(v["ms"], 3), "share": round(v["ms"] / tot_ms, 4),
                         ("tflops" if v["unit"] == "flop" else "gbs"): round(rate / (1e12 if v["unit"] == "flop" else 1e9), 1)}
    g_ms = sum(v["ms"] for k, v in prof.items() if k.startswith("gemm"))
    g_fl = sum(v["work"] for k, v in prof.items() if k.startswith("gemm"))
    g_n = sum(v["launches"] for k, v in prof.items() if k.startswith("gemm"))
    achieved = g_fl / (g_ms * 1e-3) / 1e12 if g_ms > 0 else 0.0
    # DRAM bytes per GEMM launch from the committed ncu --set full capture: quoted only for the workload it was taken on
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.isfile(tpath) and args.workload == "molly_1p7b":
        try:
            tj = json.load(open(tpath))
            vals = [l["dram_bytes"] for l in tj.get("launches", []) if "gemm" in l["kernel"]]
            if vals:
                traffic, traffic_src = sum(vals) / len(vals), tj.get("source")
        except Exception:
            pass
    roofline = {"kernel": "gemm_tcgen05_kernel (all encoder + projector GEMMs)", "bound": "tensor",
                "achieved": round(achieved, 1), "peak": peaks["tflops_sustained"], "unit": "TFLOP/s",
                "frac": round(achieved / peaks["tflops_sustained"], 4), "traffic": traffic,
                "traffic_source": traffic_src,
                "peak_source": peaks["source"] + ", sustained bf16 (kernel timed inside a long step)",
                "launches": g_n, "avg_launch_ms": round(g_ms / max(g_n, 1), 4), "share_of_step": round(g_ms / tot_ms, 4),
                "frac_of_burst": round(achieved / peaks["tflops_burst"], 4),
                "frac_of_nominal": round(achieved / 2250.0, 4)}
    step_ms = total_ms / args.steps

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warmup,
        "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic",
        "config": make_config(wl, world, n_seqs, valid_per_step),
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_ms / args.steps,
                "note": "ids from pinned host memory, strict device-error check, whole merged tensor read back to the host"},
        "e2e_device_resident": e2e_resident,
        "gpu_launches": launches,
        "roofline": roofline,
        "model_tflops_per_gpu": round(model_flops / (step_ms * 1e-3) / 1e12, 1),
        "valid_tokens_per_s": round(world * valid_per_step / (step_ms * 1e-3), 1),
        "input_fusion": input_fusion,
        "kernels": kernels,
    }
    if args.workload == "molly_1p7b" and not args.headline_only:
        # ---- the library bar, the collective-bearing train step (cfg-5) and the varlen split (cfg-3), every run
        if rank == 0 and world == 1:
            try:
                line["gpu_library_baseline"] = gpu_library_baseline(dev, path)
            except Exception as ex:
                line["gpu_library_baseline"] = {"error": f"{type(ex).__name__}: {str(ex)[:200]}"}
        path.concurrent = os.environ.get("MOLLY_CONCURRENT_MODALITIES", "1") != "0"
        line["train_step"] = train_section(args, WORKLOADS["train_1p7b"], path, dev, rank, world, max(args.steps, 5), warmup)
        path.close()
        torch.cuda.empty_cache()
        line["varlen"] = varlen_section(args, dev, rank, world, max(3, min(args.steps, 10)), warmup)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            tps, ms, sample, cores, kind = cpu_reference_run(wl, 1, 0, budget_s=args.cpu_budget_s)
            line["cpu_baseline"] = {"value": tps, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample}
        except Exception as ex:                                   # the baseline is reported, never required
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                                    "sample": f"failed: {ex}"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    path.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
