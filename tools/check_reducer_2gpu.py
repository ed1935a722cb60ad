#!/usr/bin/env python
"""Run under torchrun with 2+ GPUs: every rank trains on the SAME batch, so the layer-wise averaged gradients
(dist.LayerwiseGradReducer, all-reduces overlapped with the backward) must equal the single-rank gradients."""
import os, sys
import torch
import torch.distributed as dist
sys.path.insert(0, ".")
import bench
from molly_b200.dist import LayerwiseGradReducer

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(0)                 # bench.build_path draws the projector weights from the default generator: same on every rank
torch.cuda.manual_seed_all(0)
wl = dict(desc="reducer check", nt="nt_v2_50m", pr="esm2_t6_8m", D=1024, B=2, K=512, T=2048, valid=512)
path = bench.build_path(wl, dev, strict=True)
projs, params = {}, []
for name, enc in (("dna_rna", path.dna_rna), ("protein", path.protein)):
    lin = torch.nn.Linear(enc.proj_w.shape[1], enc.proj_w.shape[0], device=dev, dtype=torch.bfloat16)
    with torch.no_grad():
        lin.weight.copy_(enc.proj_w); lin.bias.copy_(enc.proj_b)
    projs[name] = lin
    params += list(lin.parameters())
path._proj_modules = projs
for i, (name, key) in enumerate((("dna_rna", "nt"), ("protein", "pr"))):
    bag = bench.ParamBag(bench.gpu_state_dict(bench.ENC[wl[key]], dev, 10 + i))
    path._enc_modules[name] = bag
    path._enc_versions[name] = path._module_version(bag)
    params += bag.parameters()
omic_ids, infos = bench.make_inputs(wl, seed=5)                     # the same batch on every rank
g = torch.Generator(device=dev).manual_seed(6)
base = (torch.randn(wl["B"], wl["T"], wl["D"], device=dev, generator=g) * 0.02).to(torch.bfloat16)
d_out = (torch.randn(wl["B"], wl["T"], wl["D"], device=dev, generator=g) * 1e-2).to(torch.bfloat16)

def grads_of_one_step():
    for p in params:
        p.grad = None
    out = path.process_omic_sequences(base.clone(), omic_ids.to(dev), infos, dev)
    out.backward(d_out)
    torch.cuda.synchronize()
    return [p.grad.float().clone() for p in params]

ref = grads_of_one_step()
path.grad_reducer = LayerwiseGradReducer()
red = grads_of_one_step()
# aggregate criterion: a few gradients (key biases) are near-total cancellations whose last bits depend on the order of
# the fp32 atomics in the column sums, so they differ run to run on their own tiny scale
num = sum(float((a - b).pow(2).sum()) for a, b in zip(red, ref)) ** 0.5
den = sum(float(b.pow(2).sum()) for b in ref) ** 0.5
mb = path.grad_reducer.bytes_reduced / 2 ** 20
print(f"rank {rank}/{world}: {len(params)} parameter gradients, averaged vs single-rank: relative error of the whole "
      f"gradient {num / den:.2e}, {mb:.0f} MiB all-reduced in layer-wise buckets", flush=True)
assert num / den < 2e-3, num / den
dist.destroy_process_group()
