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
    n_steps = 3
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(n_steps):
            step()
        torch.cuda.synchronize()
    # launch -> start delay per kernel (chrome trace: runtime launch events and kernels share a correlation id): a queue that
    # stays full shows delays of many milliseconds; delays of a few microseconds mean the GPU is waiting for the host
    import json
    import tempfile
    with tempfile.NamedTemporaryFile(suffix=".json") as f:
        prof.export_chrome_trace(f.name)
        tr = json.load(open(f.name))
    launch, kern = {}, {}
    for ev in tr["traceEvents"]:
        a = ev.get("args", {})
        c = a.get("correlation")
        if c is None or ev.get("ph") != "X":
            continue
        if ev.get("cat") == "cuda_runtime":
            launch[c] = ev["ts"]
        elif ev.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset"):
            kern[c] = (ev["ts"], ev["dur"], ev["name"])
    delays = sorted((kern[c][0] - launch[c], kern[c][0]) for c in kern if c in launch)
    if delays:
        d = [x for x, _ in delays]
        print(f"launch->start delay over {len(d)} launches: median {d[len(d) // 2]:.0f} us, p10 {d[len(d) // 10]:.0f} us, "
              f"p90 {d[len(d) * 9 // 10]:.0f} us; launches that started < 20 us after their launch call: "
              f"{sum(1 for x in d if x < 20)} ({100.0 * sum(1 for x in d if x < 20) / len(d):.0f} %)")
    ks = sorted(kern.values())
    t_first, t_last = ks[0][0], max(k[0] + k[1] for k in ks)
    print(f"{n_steps} steps: device span {(t_last - t_first) / 1e3 / n_steps:.2f} ms per step, busy "
          f"{sum(k[1] for k in ks) / 1e3 / n_steps:.2f} ms per step")
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
    print(f"span {(t1 - t0) / 1e3 / n_steps:.2f} ms, kernel time {busy / 1e3 / n_steps:.2f} ms, {len(kernels) // n_steps} device activities (per step)")
    for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:32]:
        print(f"{us / 1e3 / n_steps:9.3f} ms {n // n_steps:6d} x {us / n:9.1f} us  {name}")
    # gaps: idle time between consecutive device activities (single stream view)
    ev = sorted(kernels, key=lambda e: e.time_range.start)
    gaps, end = [], ev[0].time_range.end
    for e in ev[1:]:
        if e.time_range.start > end:
            gaps.append((e.time_range.start - end, e.name[:60]))
        end = max(end, e.time_range.end)
    print(f"idle {sum(g for g, _ in gaps) / 1e3 / n_steps:.2f} ms per step in {len(gaps) // n_steps} gaps; largest:")
    for g, n in sorted(gaps, reverse=True)[:15]:
        print(f"   {g:9.1f} us before {n}")
    path.close()


if __name__ == "__main__":
    main()
