"""Host-side planning of one ``process_omic_sequences`` call: route every (omic_ids, info) pair to its modality exactly
like the reference's Python loop (omics_one.py:104-118), but emit ONE packed id matrix and ONE int32 sequence table per
modality instead of one ``.to(device)`` per sequence.  Pure host logic (numpy / CPU torch), testable without a GPU.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Sequence, Tuple

import torch

NT_TYPES = ("dna", "rna")


@dataclass
class ModalityPlan:
    """Sequences of one modality, in the reference's append order (batch-major, then slot index)."""
    b_idx: List[int]
    slot_idx: List[int]
    starts: List[int]
    run_idx: List[int] = field(default_factory=list)    # index among the sample's non-'pad' slots == placeholder run index

    def __len__(self) -> int:
        return len(self.b_idx)

    def seq_table(self) -> torch.Tensor:
        """int32 [n, 2] = (b, start_pos); start -1 is kept (the kernels skip it, omics_one.py:94-95)."""
        return torch.tensor(list(zip(self.b_idx, self.starts)), dtype=torch.int32).reshape(len(self.b_idx), 2)


def route(batch_size: int, omic_ids_list, omic_info_list) -> Tuple[ModalityPlan, ModalityPlan]:
    """omics_one.py:104-118 -- pairs are zipped BY INDEX (shorter of the two wins, like ``zip``); 'pad' is skipped;
    an unknown type raises ``ValueError("Unsupported omic type: ...")``."""
    nt, pr = ModalityPlan([], [], []), ModalityPlan([], [], [])
    for b in range(batch_size):
        ids_b, infos_b = omic_ids_list[b], omic_info_list[b]
        run = 0
        for i in range(min(len(ids_b), len(infos_b))):
            info = infos_b[i]
            omic_type = info["type"]
            start_pos = info["start"]
            if omic_type in NT_TYPES:
                tgt = nt
            elif omic_type == "protein":
                tgt = pr
            elif omic_type == "pad":
                continue
            else:
                raise ValueError(f"Unsupported omic type: {omic_type}")
            tgt.b_idx.append(b)
            tgt.slot_idx.append(i)
            tgt.starts.append(int(start_pos))
            tgt.run_idx.append(run)
            run += 1
    return nt, pr


def gather_ids(omic_ids_list, plan: ModalityPlan) -> torch.Tensor:
    """``torch.stack(omic_ids)`` (omics_one.py:69) for one modality: int64 [n, K] on the device the ids live on.
    Ragged lengths raise ``RuntimeError`` exactly as ``torch.stack`` would."""
    if isinstance(omic_ids_list, torch.Tensor):               # collated [B, Nmax, K]
        bi = torch.as_tensor(plan.b_idx, dtype=torch.long, device=omic_ids_list.device)
        si = torch.as_tensor(plan.slot_idx, dtype=torch.long, device=omic_ids_list.device)
        ids = omic_ids_list[bi, si]
    else:
        ids = torch.stack([omic_ids_list[b][i] for b, i in zip(plan.b_idx, plan.slot_idx)], dim=0)
    if ids.dim() != 2:
        raise RuntimeError(f"omic ids must be 1-D per sequence, got stacked shape {tuple(ids.shape)}")
    return ids.to(torch.int64).contiguous()


def check_placement(plan: ModalityPlan, k: int, batch_size: int, seq_len: int) -> None:
    """The reference's slice-assign (omics_one.py:97) raises ``RuntimeError`` when ``start+1+k`` runs past T."""
    for b, start in zip(plan.b_idx, plan.starts):
        if start == -1:
            continue
        if start < -1 or start + 1 + k > seq_len:
            raise RuntimeError(
                f"The expanded size of the tensor ({max(0, seq_len - start - 1)}) must match the existing size ({k}) "
                f"at non-singleton dimension 0 (sample {b}: start={start}, k={k}, T={seq_len})")


def ranges_disjoint(nt: ModalityPlan, k_nt: int, pr: ModalityPlan, k_pr: int) -> bool:
    """True when no DNA/RNA row range ``[start+1, start+1+k)`` intersects a protein one in the same sample -- always the
    case for dataset-built batches (omics_dataset.py:270-288).  If they did intersect, the reference's order (DNA/RNA first,
    protein second, omics_one.py:120-134) decides who wins, so the two modalities must then run one after the other."""
    by_sample = {}
    for b, s0 in zip(nt.b_idx, nt.starts):
        if s0 != -1:
            by_sample.setdefault(b, []).append((s0 + 1, s0 + 1 + k_nt))
    for b, s0 in zip(pr.b_idx, pr.starts):
        if s0 == -1:
            continue
        lo, hi = s0 + 1, s0 + 1 + k_pr
        for a_lo, a_hi in by_sample.get(b, ()):
            if lo < a_hi and a_lo < hi:
                return False
    return True


def check_vocab(ids: torch.Tensor, vocab_size: int) -> None:
    """omics_one.py:71-72 for ids that are already on the host."""
    bad = ids >= vocab_size
    if bool(bad.any()):
        raise AssertionError(f"out-of-range token: {ids[bad]}")


def shard_samples(batch_size: int, world_size: int, rank: int) -> range:
    """Sample-sharding of SURVEY.md 8e: rank r owns samples [r*B/W, (r+1)*B/W)."""
    per = (batch_size + world_size - 1) // world_size
    return range(min(batch_size, rank * per), min(batch_size, (rank + 1) * per))


def balance_by_cost(costs: Sequence[float], world_size: int) -> List[List[int]]:
    """Greedy longest-processing-time assignment of samples to ranks (cfg-3: cost = sum kv_len^2 + sum K)."""
    order = sorted(range(len(costs)), key=lambda i: -costs[i])
    loads = [0.0] * world_size
    out: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda j: loads[j])
        out[r].append(i)
        loads[r] += costs[i]
    for lst in out:
        lst.sort()
    return out


def balance_equal_count(costs: Sequence[float], world_size: int) -> List[List[int]]:
    """Same number of samples on every rank (the per-sample cost is dominated by the K rows every sequence computes,
    whatever its length), variable part balanced: samples sorted by cost and dealt in snake order (0..W-1, W-1..0, ...).
    ``len(costs)`` must be a multiple of ``world_size``."""
    if len(costs) % world_size:
        raise ValueError(f"{len(costs)} samples do not split evenly over {world_size} ranks")
    order = sorted(range(len(costs)), key=lambda i: -costs[i])
    out: List[List[int]] = [[] for _ in range(world_size)]
    for pos, i in enumerate(order):
        rnd, off = divmod(pos, world_size)
        out[off if rnd % 2 == 0 else world_size - 1 - off].append(i)
    for lst in out:
        lst.sort()
    return out
