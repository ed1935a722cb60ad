#!/usr/bin/env python
"""Turn one gpurun measurement directory (gpurun_out/measure_<tag>/) into the tracked summary under profiles/:
   profiles/<tag>_bench.json        the bench line as printed
   profiles/<tag>_launches.md       ncu launch list aggregated per kernel (count, total us, share of the step)
   profiles/<tag>_ncu.md            key `ncu --set full` metrics of the captured GEMM / attention launches
Usage: python tools/summarize_profile.py <tag>"""
import collections
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("launch__registers_per_thread", "regs/thread"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU) pipe %"),
    ("launch__grid_size", "grid"),
    ("launch__waves_per_multiprocessor", "waves/SM"),
]


def short(name: str) -> str:
    name = re.sub(r"void\s+", "", name)
    name = re.sub(r"molly::|<unnamed>::|\(anonymous namespace\)::|unnamed>::", "", name)
    return name.split("(")[0][:90]


def launches_md(path: str) -> str:
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 10]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        v = {"ns": v / 1e3, "us": v, "ms": v * 1e3, "s": v * 1e6}.get(r[ui], v)
        a = agg.setdefault(short(r[ki]), [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values()) or 1.0
    out = ["| kernel | launches | total us | share |", "|---|---:|---:|---:|"]
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| `{k}` | {n} | {t:.1f} | {100 * t / tot:.1f} % |")
    out.append(f"| **total** | {sum(v[0] for v in agg.values())} | {tot:.1f} | 100 % |")
    return "\n".join(out)


def ncu_md(rep: str) -> str:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    if len(rows) < 3:
        return f"(could not read {rep})"
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    out = []
    for r in rows[2:]:
        out.append(f"**`{short(r[idx['Kernel Name']])}`**\n")
        out.append("| metric | value |")
        out.append("|---|---:|")
        for k, label in KEYS:
            if k in idx:
                out.append(f"| {label} (`{k}`) | {r[idx[k]]} {units[idx[k]]} |")
        out.append("")
    return "\n".join(out)


def traffic_json(rep: str) -> dict:
    """dram__bytes_read.sum + dram__bytes_write.sum per captured launch (bench.py's roofline.traffic)."""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    if len(rows) < 3:
        return {}
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    out = []
    for r in rows[2:]:
        tot = 0.0
        for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            tot += float(r[idx[k]].replace(",", "")) * scale.get(units[idx[k]], 1.0)
        out.append({"kernel": short(r[idx["Kernel Name"]]), "dram_bytes": tot,
                    "duration_us": float(r[idx["gpu__time_duration.sum"]].replace(",", "")) *
                    {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(units[idx["gpu__time_duration.sum"]], 1.0)})
    return {"launches": out}


def main():
    tag = sys.argv[1]
    src = os.path.join(ROOT, "gpurun_out", f"measure_{tag}")
    dst = os.path.join(ROOT, "profiles")
    os.makedirs(dst, exist_ok=True)
    bj = os.path.join(src, "bench.json")
    if os.path.isfile(bj):
        line = [l for l in open(bj).read().splitlines() if l.startswith("{")][-1]
        json.loads(line)
        open(os.path.join(dst, f"{tag}_bench.json"), "w").write(line + "\n")
    lc = os.path.join(src, "launches.csv")
    if os.path.isfile(lc):
        with open(os.path.join(dst, f"{tag}_launches.md"), "w") as f:
            f.write(f"# ncu launch list `{tag}` -- one bench step, `--metrics gpu__time_duration.sum --clock-control none`\n\n"
                    "Per-launch times are cold-cache and serialised: compare SHARES with the bench's live CUDA-event "
                    "numbers (`kernels` in the bench line), not absolutes.\n\n")
            f.write(launches_md(lc) + "\n")
    parts = []
    for name in ("prof_gemm", "prof_attn", "prof_ln"):
        rep = os.path.join(src, name + ".ncu-rep")
        if os.path.isfile(rep):
            parts.append(f"## {name}.ncu-rep (`ncu --set full --clock-control none --import-source on`)\n\n" + ncu_md(rep))
    if parts:
        open(os.path.join(dst, f"{tag}_ncu.md"), "w").write(f"# ncu --set full summary `{tag}`\n\n" + "\n".join(parts))
    rep = os.path.join(src, "prof_gemm.ncu-rep")
    if os.path.isfile(rep):
        tj = traffic_json(rep)
        tj["source"] = f"profiles/{tag}_ncu.md (ncu --set full, one NT-v2-500M encoder layer: QKV, attn-out, FFN1, FFN2)"
        json.dump(tj, open(os.path.join(dst, "ncu_traffic.json"), "w"), indent=1)
    print("wrote", sorted(f for f in os.listdir(dst) if f.startswith(tag)))


if __name__ == "__main__":
    main()
