#!/usr/bin/env python
"""Turn the raw ncu artefacts in gpurun_out/ into the small tracked summaries under profiles/.

    python tools/summarize_profiles.py <tag> [--launches gpurun_out/launches.csv]
                                             [--rep gpurun_out/prof_cg.ncu-rep] [--bench gpurun_out/bench.log]

Writes profiles/<tag>_launches.md (per-kernel launch count, total device time, share of the
step), profiles/<tag>_<rep>.md (the metrics DESIGN.md quotes + the most-sampled SASS lines
with their dominant stall reason) and copies the bench JSON line.
"""
import argparse
import collections
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "lts__t_bytes.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max",
]


def launches(path, out):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        name = r[ik].split("(")[0]
        name = name.replace("ials::<unnamed>::", "")[-70:]
        v = float(r[iv].replace(",", ""))
        scale = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3}.get(r[iu], 1.0)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v * scale
    tot = sum(a[1] for a in agg.values())
    with open(out, "w") as f:
        f.write(f"# ncu launch list ({os.path.basename(path)}): `--metrics gpu__time_duration.sum "
                "--clock-control none`\n\nPer-launch times are cold-cache and serialised: read the SHARES.\n\n")
        f.write("| kernel | launches | total us | share |\n|---|---:|---:|---:|\n")
        for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f"| `{k}` | {n} | {t:.1f} | {t / tot:.3f} |\n")
    print("wrote", out)


def ncu_csv(rep, page):
    res = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True)
    return list(csv.reader(io.StringIO(res.stdout)))


def report(rep, out, top=25):
    raw = ncu_csv(rep, "raw")
    hdr, units = raw[0], raw[1]
    with open(out, "w") as f:
        f.write(f"# ncu --set full summary of {os.path.basename(rep)}\n\n")
        for r in raw[2:]:
            f.write(f"## {r[hdr.index('Kernel Name')][:100]}  (launch id {r[hdr.index('ID')]})\n\n| metric | value | unit |\n|---|---:|---|\n")
            for k in KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    f.write(f"| {k} | {r[i]} | {units[i]} |\n")
            f.write("\n")
        src = ncu_csv(rep, "source")
        blocks, h = [], None
        for r in src:
            if r and r[0] == "Kernel Name":
                blocks.append([r[1] if len(r) > 1 else "", []])
            elif r and r[0] == "Address":
                h = r
            elif blocks and r:
                blocks[-1][1].append(r)
        if h is not None:
            i_s, i_src, i_ex = h.index("# Samples"), h.index("Source"), h.index("Instructions Executed")
            stalls = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
            for name, b in blocks[:1]:
                tot = sum(int(r[i_s]) for r in b) or 1
                f.write(f"## stall samples, first captured launch of `{name[:80]}`\n\n")
                agg = {s: sum(int(r[h.index(s)]) for r in b) for s in stalls}
                f.write("| stall reason | samples | share |\n|---|---:|---:|\n")
                for s, v in sorted(agg.items(), key=lambda x: -x[1])[:8]:
                    f.write(f"| {s} | {v} | {v / tot:.3f} |\n")
                f.write(f"\nMost-sampled SASS instructions (of {tot} samples):\n\n| # | samples | executed | SASS | top stall |\n|---|---:|---:|---|---|\n")
                order = sorted(range(len(b)), key=lambda i: -int(b[i][i_s]))[:top]
                for i in sorted(order):
                    r = b[i]
                    st = max(((int(r[h.index(s)]), s) for s in stalls))
                    f.write(f"| {i} | {r[i_s]} | {r[i_ex]} | `{r[i_src].strip()[:70]}` | {st[1]} ({st[0]}) |\n")
    print("wrote", out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("tag")
    ap.add_argument("--launches")
    ap.add_argument("--rep", action="append", default=[])
    ap.add_argument("--bench")
    a = ap.parse_args()
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    if a.launches:
        launches(a.launches, os.path.join(ROOT, "profiles", f"{a.tag}_launches.md"))
    for rep in a.rep:
        base = os.path.splitext(os.path.basename(rep))[0]
        report(rep, os.path.join(ROOT, "profiles", f"{a.tag}_{base}.md"))
    if a.bench:
        lines = [l for l in open(a.bench) if l.startswith("{")]
        if lines:
            with open(os.path.join(ROOT, "profiles", f"{a.tag}_bench.json"), "w") as f:
                f.write(json.dumps(json.loads(lines[-1]), indent=1) + "\n")
            print("wrote bench json")


if __name__ == "__main__":
    sys.exit(main())
