"""CPU, world_size 2, gloo: the N>1 host logic -- sample sharding covers the batch exactly once, per-rank routing plans
concatenate to the single-rank plan, the flat projector-gradient bucket averages across ranks, and the bench's
max-over-ranks timing reduction works."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from molly_b200 import planner
from oracle import cases


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from molly_b200.dist import FlatGradBucket, shard_batch
        case = cases.golden_cases()["tiny_rotary_glu"]
        hs, ids, infos = shard_batch(rank, world, case.batch.hidden_states, case.batch.omic_ids,
                                     case.batch.omic_info_list)
        nt, pr = planner.route(hs.shape[0], ids, infos)
        r = planner.shard_samples(case.batch.hidden_states.shape[0], world, rank)
        plan = {"rank": rank, "samples": list(r),
                "nt": [(b + r.start, i, s) for b, i, s in zip(nt.b_idx, nt.slot_idx, nt.starts)],
                "pr": [(b + r.start, i, s) for b, i, s in zip(pr.b_idx, pr.slot_idx, pr.starts)]}
        # flat-bucket mean all-reduce of projector grads
        lin = torch.nn.Linear(8, 4)
        lin.weight.grad = torch.full_like(lin.weight, float(rank + 1))
        lin.bias.grad = torch.full_like(lin.bias, float(10 * (rank + 1)))
        bucket = FlatGradBucket(list(lin.parameters()))
        bucket.launch()
        bucket.finish()
        # a one-tensor bucket (the LoRA-sized gradient of cfg-5) is averaged in place: same storage before and after
        lora = torch.nn.Parameter(torch.zeros(1000))
        lora.grad = torch.full((1000,), float(3 * (rank + 1)))
        ptr = lora.grad.data_ptr()
        lb = FlatGradBucket([lora])
        lb.launch()
        lb.finish()
        plan["lora"] = (float(lora.grad.mean()), lora.grad.data_ptr() == ptr)
        plan["w_grad"] = float(lin.weight.grad.mean())
        plan["b_grad"] = float(lin.bias.grad.mean())
        # layer-wise reducer of the encoder backward: entries become views of one averaged buffer per call
        from molly_b200.dist import LayerwiseGradReducer
        red = LayerwiseGradReducer(dtype=torch.float32)
        grads = {"layer.0.w": torch.full((3, 4), float(rank + 1)), "layer.0.b": torch.full((4,), float(2 * rank)),
                 "layer.1.w": torch.full((2, 2), 7.0 * (rank + 1))}
        red.reduce_(grads, ["layer.0.w", "layer.0.b"])
        red.reduce_(grads, ["layer.1.w", "missing"])
        red.finish()
        plan["lw"] = [float(grads["layer.0.w"].mean()), float(grads["layer.0.b"].mean()), float(grads["layer.1.w"].mean())]
        plan["lw_shapes"] = [tuple(grads[k].shape) for k in ("layer.0.w", "layer.0.b", "layer.1.w")]
        # asymmetric micro-batches: rank 1 has no sequence of the modality and issues zero all-reduces of the same sizes
        red2 = LayerwiseGradReducer(dtype=torch.float32)
        if rank == 0:
            g2 = {"a": torch.full((5,), 4.0), "b": torch.full((2, 3), 8.0)}
            red2.reduce_(g2, ["a"])
            red2.reduce_(g2, ["b"])
            red2.finish()
            plan["asym"] = [float(g2["a"].mean()), float(g2["b"].mean())]
        else:
            red2.reduce_zeros_(5, torch.device("cpu"))
            red2.reduce_zeros_(6, torch.device("cpu"))
            red2.finish()
            plan["asym"] = [2.0, 4.0]
        # bench.py's timing reduction: max over ranks
        t = torch.tensor([float(rank + 1)])
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        plan["t_max"] = float(t)
        out_q.put(plan)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_two_rank_gloo_sharding_and_grad_bucket():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    plans = sorted([q.get(timeout=150) for _ in range(world)], key=lambda d: d["rank"])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    case = cases.golden_cases()["tiny_rotary_glu"]
    B = case.batch.hidden_states.shape[0]
    assert sorted(sum((p["samples"] for p in plans), [])) == list(range(B))
    nt, pr = planner.route(B, case.batch.omic_ids, case.batch.omic_info_list)
    assert sum((p["nt"] for p in plans), []) == list(zip(nt.b_idx, nt.slot_idx, nt.starts))
    assert sum((p["pr"] for p in plans), []) == list(zip(pr.b_idx, pr.slot_idx, pr.starts))
    for p in plans:
        assert p["w_grad"] == pytest.approx(1.5) and p["b_grad"] == pytest.approx(15.0)
        assert p["t_max"] == 2.0
        assert p["lora"][0] == pytest.approx(4.5) and p["lora"][1]
        assert p["lw"] == pytest.approx([1.5, 1.0, 10.5])            # means over ranks of (1,2), (0,2), (7,14)
        assert p["lw_shapes"] == [(3, 4), (4,), (2, 2)]
        assert p["asym"] == pytest.approx([2.0, 4.0])                # (4 + 0) / 2, (8 + 0) / 2: no hang, no mis-pairing
