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
if num / den >= 2e-3:            # diagnostics: which tensors, and is it the local reference or the reduced run that moved?
    pnames = ([f"proj.{m}.{n}" for m in ("dna_rna", "protein") for n in ("weight", "bias")]
              + [f"{m}.{n}" for m in ("dna_rna", "protein") for n, _ in path._enc_modules[m].named_parameters()])
    errs = sorted(((float((a - b).norm() / (b.norm() + 1e-30)), n, float(a.norm()), float(b.norm()))
                   for n, a, b in zip(pnames, red, ref)), reverse=True)
    for e, n, na, nb in errs[:12]:
        print(f"rank {rank}: {n}: rel err {e:.3e}  |reduced| {na:.4e}  |reference| {nb:.4e}", flush=True)
for tag, gl in (("reference", ref), ("reduced", red)):
    chk = torch.tensor([float(sum(t.double().abs().sum() for t in gl))], device=dev, dtype=torch.float64)
    both = [torch.zeros_like(chk) for _ in range(world)]
    dist.all_gather(both, chk)
    print(f"rank {rank}: checksum of the {tag} gradients per rank: {[float(b) for b in both]}", flush=True)
assert num / den < 2e-3, num / den

# ---- asymmetric micro-batches: rank 1's batch has NO protein sequence.  Its backward must issue zero
#      all-reduces of the same sizes at the same point (omics_path._AbsentModalityFn) -- without them NCCL hangs or pairs
#      the wrong buffers.  Expected averaged gradients: dna_rna = the single-rank ones (every rank has the same DNA batch),
#      protein = single-rank x (ranks that had the protein) / world.
infos_asym = [[i if (i["type"] != "protein" or rank != 1) else {"type": "pad", "start": -1} for i in row] for row in infos]
def asym_step():
    for p in params:
        p.grad = None
    out = path.process_omic_sequences(base.clone(), omic_ids.to(dev), infos_asym, dev)
    out.backward(d_out)
    torch.cuda.synchronize()
    return [None if p.grad is None else p.grad.float().clone() for p in params]
asym = asym_step()
n_pr_params = len(list(projs["protein"].parameters())) + len(path._enc_modules["protein"].parameters())
names = ([("dna_rna", p) for p in projs["dna_rna"].parameters()] + [("protein", p) for p in projs["protein"].parameters()]
         + [("dna_rna", p) for p in path._enc_modules["dna_rna"].parameters()]
         + [("protein", p) for p in path._enc_modules["protein"].parameters()])
scale = {"dna_rna": 1.0, "protein": (world - 1) / world if world > 1 else 1.0}
num = den = 0.0
for (mod, _), a, b in zip(names, asym, ref):
    a = torch.zeros_like(b) if a is None else a
    num += float((a - b * scale[mod]).pow(2).sum()); den += float((b * scale[mod]).pow(2).sum())
print(f"rank {rank}/{world}: asymmetric modalities (rank 1 without protein): relative error of the averaged gradient "
      f"{(num / den) ** 0.5:.2e}", flush=True)
assert (num / den) ** 0.5 < 2e-3

# ---- FlatGradBucket over NCCL (its stream path was only covered by the gloo CPU test): multi-tensor bucket with
#      pack / unpack on the side stream, and the one-tensor bucket reduced in place
from molly_b200.dist import FlatGradBucket
lin = torch.nn.Linear(256, 128, device=dev, dtype=torch.bfloat16)
big = torch.nn.Parameter(torch.zeros(8 << 20, device=dev, dtype=torch.bfloat16))
for trial in range(3):
    lin.weight.grad = torch.full_like(lin.weight, float(rank + 1 + trial))
    lin.bias.grad = torch.full_like(lin.bias, float(10 * (rank + 1)))
    big.grad = torch.full_like(big, float(2 * rank + trial))
    b1, b2 = FlatGradBucket(list(lin.parameters())), FlatGradBucket([big])
    b2.launch(); b1.launch(); b1.finish(); b2.finish()
    torch.cuda.synchronize()
    want_w = sum(r + 1 + trial for r in range(world)) / world
    want_b = sum(10 * (r + 1) for r in range(world)) / world
    want_big = sum(2 * r + trial for r in range(world)) / world
    assert float((lin.weight.grad.float() - want_w).abs().max()) < 1e-2 and float((lin.bias.grad.float() - want_b).abs().max()) < 1e-1
    assert float((big.grad.float() - want_big).abs().max()) < 1e-2, (float(big.grad.float().mean()), want_big)
print(f"rank {rank}/{world}: FlatGradBucket (packed + in-place) over NCCL ok", flush=True)
dist.destroy_process_group()
