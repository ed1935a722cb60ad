"""Per-kernel device times of one --train-bio step (torch profiler / CUPTI; every kernel incl. torch's own copies)."""
import collections
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    wl = bench.WORKLOADS["train_bio_1p7b"]
    path = bench.build_path(wl, dev, strict=False)
    projs, params = {}, []
    for name, enc in (("dna_rna", path.dna_rna), ("protein", path.protein)):
        lin = torch.nn.Linear(enc.proj_w.shape[1], enc.proj_w.shape[0], device=dev, dtype=torch.bfloat16)
        with torch.no_grad():
            lin.weight.copy_(enc.proj_w)
            lin.bias.copy_(enc.proj_b)
        projs[name] = lin
    path._proj_modules = projs
    params = [p for lin in projs.values() for p in lin.parameters()]
    for i, (name, key) in enumerate((("dna_rna", "nt"), ("protein", "pr"))):
        bag = bench.ParamBag(bench.gpu_state_dict(bench.ENC[wl[key]], dev, 10 + i))
        path._enc_modules[name] = bag
        path._enc_versions[name] = path._module_version(bag)
        params += bag.parameters()
    omic_ids, infos = bench.make_inputs(wl, seed=1234)
    ids = omic_ids.to(dev)
    base = (torch.randn(wl["B"], wl["T"], wl["D"], device=dev) * 0.02).to(torch.bfloat16)
    d_out = (torch.randn(wl["B"], wl["T"], wl["D"], device=dev) * 1e-3).to(torch.bfloat16)

    def step():
        for p in params:
            p.grad = None
        out = path.process_omic_sequences(base.clone(), ids, infos, dev)
        out.backward(d_out)

    for _ in range(6):
        step()
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        step()
        torch.cuda.synchronize()
    agg = collections.OrderedDict()
    kernels = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    t0 = min(e.time_range.start for e in kernels)
    t1 = max(e.time_range.end for e in kernels)
    busy = 0.0
    for e in kernels:
        a = agg.setdefault(e.name[:90], [0, 0.0])
        a[0] += 1
        a[1] += e.time_range.elapsed_us()
        busy += e.time_range.elapsed_us()
    print(f"span {(t1 - t0) / 1e3:.2f} ms, kernel time {busy / 1e3:.2f} ms, {len(kernels)} device activities")
    for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:60]:
        print(f"{us / 1e3:9.3f} ms {n:6d} x {us / n:9.1f} us  {name}")
    # gaps: idle time between consecutive device activities (single stream view)
    ev = sorted(kernels, key=lambda e: e.time_range.start)
    gaps, end = [], ev[0].time_range.end
    for e in ev[1:]:
        if e.time_range.start > end:
            gaps.append((e.time_range.start - end, e.name[:60]))
        end = max(end, e.time_range.end)
    print(f"idle {sum(g for g, _ in gaps) / 1e3:.2f} ms in {len(gaps)} gaps; largest:")
    for g, n in sorted(gaps, reverse=True)[:15]:
        print(f"   {g:9.1f} us before {n}")
    path.close()


if __name__ == "__main__":
    main()
