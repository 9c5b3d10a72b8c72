"""Per-kernel census of the Blackwell-specific SASS mnemonics in libditto_b200.so (cuobjdump -sass), written as a markdown table:
    python tools/sass_census.py [out.md]
UTCHMMA = tcgen05.mma (bf16), .2CTA = cta_group::2; LDTM / STTM = tcgen05.ld / st; UTMALDG / UTMASTG = TMA tensor load / store;
UBLKCP = cp.async.bulk (shared -> cluster shared / global <-> shared); UTCBAR = tcgen05.commit; SYNCS = mbarrier ops;
HMMA = legacy mma.sync (must be 0); UCGABAR = cluster barrier; MUFU.EX2 = ex2.approx."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "ditto_tts_b200", "libditto_b200.so")
WANT = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "SYNCS", "UCGABAR", "MUFU.EX2", "HMMA", "STL/LDL"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    res = {}
    for m, d in zip(names, out):
        d = re.sub(r"\(anonymous namespace\)::|ditto::", "", d)
        d = re.sub(r"\(.*", "", d).replace("void ", "")
        res[m] = d
    return res


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else None
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    counts = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = counts.setdefault(m.group(1), collections.Counter())
            continue
        if cur is None:
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        base = op.split(".")[0]
        if base == "UTCHMMA":
            cur["UTCHMMA"] += 1
            if ".2CTA" in op:
                cur["UTCHMMA.2CTA"] += 1
        elif base in ("LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "SYNCS", "UCGABAR", "HMMA"):
            cur[base] += 1
        elif op.startswith("MUFU.EX2"):
            cur["MUFU.EX2"] += 1
        elif base in ("STL", "LDL"):
            cur["STL/LDL"] += 1
        cur["_total"] += 1
    names = demangle(list(counts))
    rows = []
    for k, c in counts.items():
        if c["UTCHMMA"] or c["UTMALDG"] or c["LDTM"] or c["UBLKCP"]:
            rows.append((names[k], c))
    rows.sort(key=lambda r: r[0])
    lines = ["| kernel | SASS instrs | " + " | ".join(WANT) + " |", "|---|---:|" + "---:|" * len(WANT)]
    for n, c in rows:
        lines.append(f"| `{n}` | {c['_total']} | " + " | ".join(str(c[w]) for w in WANT) + " |")
    tot = collections.Counter()
    for c in counts.values():
        tot.update(c)
    lines.append(f"| **whole library ({len(counts)} kernels)** | {tot['_total']} | " + " | ".join(str(tot[w]) for w in WANT) + " |")
    text = "\n".join(lines) + "\n"
    if out_path:
        with open(out_path, "w") as f:
            f.write("# SASS census of ditto_tts_b200/libditto_b200.so (`python tools/sass_census.py`, cuobjdump -sass, sm_100a)\n\n")
            f.write("Kernels that use the tensor cores / TMEM / TMA; the last row covers every kernel of the library.\n\n" + text)
    print(text)


if __name__ == "__main__":
    main()
