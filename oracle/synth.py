"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py) -- seeded synthetic batches shaped the way the
reference's dataset + collate produce them (REF = /root/reference):

  * placeholder run ``x_start, x_pad * K, x_end`` with ``info["start"] = len(input_ids)`` before the run
    (REF/src/dataset/omics_dataset.py:270-288), so rows ``start+1 .. start+K`` are the pad tokens;
  * omics ids padded / truncated to exactly K with tokenizer pad id 1 (``_encode_sequence`` :420-447);
  * collate pads ``omic_ids`` rows with 1 and infos with ``{"type": "pad", "start": -1}`` (:480-492) and
    ``input_ids`` with 0; Test mode left-pads and shifts ``start`` (:387-391);
  * ids are listed in kind order dna -> rna -> protein (:252-263, 304-307), infos in text order (:270); the
    generator lays the runs out in kind order so both orders agree (the reference mis-pairs otherwise).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Sequence, Tuple

import torch

from .esm_oracle import EncoderSpec

# ids the 9 added special tokens get on the Qwen3 tokenizer (REF/src/train.py:73-85): appended after 151668
PLACEHOLDER_BASE = 151669
SPECIAL = {
    "dna": {"start": PLACEHOLDER_BASE + 0, "pad": PLACEHOLDER_BASE + 1, "end": PLACEHOLDER_BASE + 2},
    "rna": {"start": PLACEHOLDER_BASE + 3, "pad": PLACEHOLDER_BASE + 4, "end": PLACEHOLDER_BASE + 5},
    "protein": {"start": PLACEHOLDER_BASE + 6, "pad": PLACEHOLDER_BASE + 7, "end": PLACEHOLDER_BASE + 8},
}
PAD_TOKEN_IDS = (SPECIAL["dna"]["pad"], SPECIAL["rna"]["pad"], SPECIAL["protein"]["pad"])
KIND_ORDER = {"dna": 0, "rna": 1, "protein": 2}


def protein_ids(g: torch.Generator, K: int, valid: int) -> torch.Tensor:
    """ESM-2 vocabulary: <cls>=0 <pad>=1 <eos>=2 <unk>=3, residues 4.., <mask>=32."""
    ids = torch.full((K,), 1, dtype=torch.int64)
    valid = max(2, min(valid, K))
    ids[:valid] = torch.randint(4, 24, (valid,), generator=g)
    ids[0] = 0
    ids[valid - 1] = 2
    return ids


def nucleotide_ids(g: torch.Generator, K: int, valid: int, vocab: int) -> torch.Tensor:
    """NT vocabulary: <cls>=3 at position 0, k-mers in [6, vocab-5), pad=1."""
    ids = torch.full((K,), 1, dtype=torch.int64)
    valid = max(1, min(valid, K))
    ids[:valid] = torch.randint(6, max(7, vocab - 5), (valid,), generator=g)
    ids[0] = 3
    return ids


@dataclass
class Batch:
    input_ids: torch.Tensor            # [B, T] int64
    omic_ids: torch.Tensor             # [B, Nmax, K] int64 (pad rows all 1)
    omic_info_list: List[List[dict]]   # [B][Nmax] {"type", "start"}
    hidden_states: torch.Tensor        # [B, T, D]
    n_omic_tokens: int                 # rows the path encodes (= sum over non-pad sequences of K)
    n_valid_tokens: int


def make_batch(seed: int, samples: Sequence[Sequence[Tuple[str, int]]], T: int, D: int, K: int,
               nt_spec: EncoderSpec, pr_spec: EncoderSpec, *, text_vocab: int = 151000,
               left_pad: bool = False, dtype=torch.float32, gap: Tuple[int, int] = (3, 12)) -> Batch:
    """``samples[b]`` = list of ``(kind, valid_len)``; both modalities share one K (collate needs a rectangular
    ``omic_ids``; REF scripts always set ``--dna-rna-k-tokens == --protein-k-tokens``)."""
    g = torch.Generator().manual_seed(seed)
    B = len(samples)
    n_max = max(1, max(len(s) for s in samples))
    input_rows, infos_all, ids_all = [], [], []
    n_tok = n_valid = 0
    for b, seqs in enumerate(samples):
        seqs = sorted(seqs, key=lambda kv: KIND_ORDER[kv[0]])
        toks: List[int] = torch.randint(0, text_vocab, (int(torch.randint(gap[0], gap[1] + 1, (1,), generator=g)),),
                                        generator=g).tolist()
        infos, ids = [], []
        for kind, valid in seqs:
            infos.append({"type": kind, "start": len(toks)})
            toks.append(SPECIAL[kind]["start"])
            toks.extend([SPECIAL[kind]["pad"]] * K)
            toks.append(SPECIAL[kind]["end"])
            toks.extend(torch.randint(0, text_vocab, (int(torch.randint(gap[0], gap[1] + 1, (1,), generator=g)),),
                                      generator=g).tolist())
            if kind == "protein":
                ids.append(protein_ids(g, K, valid))
            else:
                ids.append(nucleotide_ids(g, K, valid, nt_spec.vocab_size))
            n_tok += K
            n_valid += max(1, min(valid, K))
        if len(toks) > T:
            raise ValueError(f"sample {b} needs {len(toks)} tokens > T={T}")
        pad_len = T - len(toks)
        if left_pad:                                          # Test mode: shift starts (:387-391)
            toks = [0] * pad_len + toks
            for i in infos:
                i["start"] += pad_len
        else:
            toks = toks + [0] * pad_len
        while len(ids) < n_max:                               # collate padding (:480-492)
            ids.append(torch.full((K,), 1, dtype=torch.int64))
            infos.append({"type": "pad", "start": -1})
        input_rows.append(torch.tensor(toks, dtype=torch.int64))
        infos_all.append(infos)
        ids_all.append(torch.stack(ids))
    input_ids = torch.stack(input_rows)
    hidden = (torch.randn(B, T, D, generator=g) * 0.02).to(dtype)
    return Batch(input_ids, torch.stack(ids_all), infos_all, hidden, n_tok, n_valid)


def expected_rows(omic_info_list: List[List[dict]], K: int, k_cap_nt: int, k_cap_pr: int) -> Dict[Tuple[int, int], Tuple[int, int, int]]:
    """The reference's index set: ``(b, start+1+j) -> (b, i, j)`` for every non-pad sequence (omics_one.py:93-97)."""
    rows = {}
    for b, infos in enumerate(omic_info_list):
        for i, info in enumerate(infos):
            if info["type"] == "pad" or info["start"] == -1:
                continue
            k = min(k_cap_pr if info["type"] == "protein" else k_cap_nt, K)
            for j in range(k):
                rows[(b, info["start"] + 1 + j)] = (b, i, j)
    return rows


def placeholder_runs(input_ids: torch.Tensor, pad_ids: Sequence[int]) -> List[List[Tuple[int, int, int]]]:
    """Per sample, the maximal runs of *_pad tokens in text order as ``(first_pad_position, kind, length)``.
    ``first_pad_position - 1`` is the x_start token, i.e. ``info["start"]`` as the dataset records it
    (REF/src/dataset/omics_dataset.py:270-288; Test-mode left padding shifts it, :387-391).  Plain Python loop."""
    pad_ids = [int(p) for p in pad_ids]
    out = []
    for row in input_ids.tolist():
        runs, t, T = [], 0, len(row)
        while t < T:
            if row[t] in pad_ids:
                s = t
                while t < T and row[t] in pad_ids:
                    t += 1
                runs.append((s, pad_ids.index(row[s]), t - s))
            else:
                t += 1
        out.append(runs)
    return out


def embed_then_process(input_ids: torch.Tensor, embed_weight: torch.Tensor, omic_ids_list, omic_info_list, nt, pr):
    """``inputs_embeds = embed_tokens(input_ids)`` followed by ``process_omic_sequences`` (REF/src/model/omics_one.py:164-170
    in ``forward`` and :209-215 in ``generate``)."""
    from .esm_oracle import process_omic_sequences
    hidden = embed_weight[input_ids]
    return process_omic_sequences(hidden, omic_ids_list, omic_info_list, nt, pr)
