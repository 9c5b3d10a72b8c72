"""Summarise ncu CSV exports into the tables committed under profiles/.
    python tools/summarize_ncu.py launches <launches.csv> <out.md>     per-kernel totals / shares of a launch list
    python tools/summarize_ncu.py full <raw.csv> <out.md> [traffic.json]   key metrics of a --set full capture"""
import csv
import json
import re
import sys
from collections import OrderedDict


def short(name):
    name = re.sub(r"ditto::<unnamed>::|ditto::|void |\(anonymous namespace\)::", "", name)
    name = re.sub(r"^.*?unnamed>::", "", name)
    name = re.sub(r"\(CUtensorMap_st.*", "", name)
    name = re.sub(r"\(.*", "", name)
    return name.strip()


KEPI = {0: "store_f32", 1: "store_f32_resid", 2: "store_bf16", 3: "store_generic", 4: "geglu", 5: "qkv_rope"}


def pretty(name):
    n = short(name)
    m = re.match(r"tc_gemm_pair_kernel<\(int\)(\d)>|tc_gemm_pair_kernel<(\d)>", n)
    if m:
        return "tc_gemm_pair_kernel<%s>" % KEPI[int(m.group(1) or m.group(2))]
    m = re.match(r"tc_gemm_kernel<\(int\)(\d), \(bool\)(\d)>|tc_gemm_kernel<(\d), (\d)>", n)
    if m:
        e = int(m.group(1) or m.group(3)); kn = int(m.group(2) or m.group(4))
        return "tc_gemm_kernel<%s,%s>" % (KEPI[e], "B_kn" if kn else "B_nk")
    return n


def launches(path, out):
    rows = [r for r in csv.reader(open(path, errors="replace")) if r and not r[0].startswith("==")]
    hdr = rows[0]
    ki, mi, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    ui = hdr.index("Metric Unit")
    tot = OrderedDict()
    for r in rows[1:]:
        if len(r) <= vi or r[mi] != "gpu__time_duration.sum":
            continue
        v = float(r[vi].replace(",", ""))
        if r[ui] in ("ns", "nsecond"):
            v /= 1e3
        elif r[ui] in ("ms", "msecond"):
            v *= 1e3
        k = pretty(r[ki])
        a = tot.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
    total = sum(v[1] for v in tot.values())
    with open(out, "w") as f:
        f.write("| kernel | launches | total us | us / launch | share |\n|---|---:|---:|---:|---:|\n")
        for k, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {n} | {us:.1f} | {us / n:.1f} | {100 * us / total:.1f}% |\n")
        f.write(f"| **total** | {sum(v[0] for v in tot.values())} | {total:.1f} | | |\n")


WANT = [("gpu__time_duration.sum", "us"), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %"),
        ("dram__bytes_read.sum", "DRAM rd MB"), ("dram__bytes_write.sum", "DRAM wr MB"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"), ("launch__registers_per_thread", "regs"),
        ("launch__grid_size", "grid"), ("smsp__inst_executed.sum", "warp insts")]


def full(path, out, traffic=None):
    rows = list(csv.reader(open(path, errors="replace")))
    hdr, units, data = rows[0], rows[1], rows[2:]
    tr = {}
    with open(out, "w") as f:
        f.write("| # | kernel | " + " | ".join(w[1] for w in WANT) + " |\n|---|---|" + "---:|" * len(WANT) + "\n")
        for i, r in enumerate(data):
            cells = []
            for m, lab in WANT:
                if m not in hdr:
                    cells.append("-")
                    continue
                j = hdr.index(m)
                try:
                    v = float(r[j].replace(",", ""))
                except ValueError:
                    cells.append(r[j])
                    continue
                u = units[j]
                if "bytes" in m:
                    v *= {"Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3, "byte": 1e-6}.get(u, 1.0)
                if m == "gpu__time_duration.sum":
                    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(u, 1.0)
                cells.append(f"{v:.1f}" if abs(v) < 1e6 else f"{v:.3g}")
            name = pretty(r[hdr.index("Kernel Name")])
            f.write(f"| {i} | `{name}` | " + " | ".join(cells) + " |\n")
            try:
                rd = float(r[hdr.index("dram__bytes_read.sum")]); ru = units[hdr.index("dram__bytes_read.sum")]
                wr = float(r[hdr.index("dram__bytes_write.sum")]); wu = units[hdr.index("dram__bytes_write.sum")]
                sc = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}
                tr.setdefault(name, []).append(rd * sc.get(ru, 1.0) + wr * sc.get(wu, 1.0))
            except (ValueError, KeyError):
                pass
    if traffic:
        json.dump({k: sum(v) / len(v) for k, v in tr.items()}, open(traffic, "w"), indent=1)


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        full(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else None)
