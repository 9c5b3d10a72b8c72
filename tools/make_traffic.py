"""profiles/traffic.json from an `ncu --set full` raw-page CSV of ONE launch of the dominant kernel at the bench's problem
size (tools/capture_traffic.sh): dram__bytes_read.sum + dram__bytes_write.sum, tied to the kernel source by its sha.
    python tools/make_traffic.py <raw.csv> [class=tc_gemm.glu] [note]"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (csrc_sha)

path = sys.argv[1]
cls = sys.argv[2] if len(sys.argv) > 2 else "tc_gemm.glu"
note = sys.argv[3] if len(sys.argv) > 3 else "one launch at the C4 size (M = 384 000 rows), tools/capture_traffic.sh"
rows = list(csv.reader(open(path, errors="replace")))
hdr, units, data = rows[0], rows[1], rows[2:]
sc = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}
tot, n = 0.0, 0
for r in data:
    b = 0.0
    for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        j = hdr.index(m)
        b += float(r[j].replace(",", "")) * sc.get(units[j], 1.0)
    tot += b
    n += 1
    print(r[hdr.index("Kernel Name")][:90], f"{b / 1e6:.1f} MB", r[hdr.index("gpu__time_duration.sum")], units[hdr.index("gpu__time_duration.sum")])
out = {"_source": f"{os.path.basename(path)}: ncu --set full --clock-control none, {note}; dram__bytes_read.sum + dram__bytes_write.sum "
                  f"per launch, bytes", "_csrc_sha": bench.csrc_sha(), cls: tot / max(n, 1)}
with open(os.path.join(ROOT, "profiles", "traffic.json"), "w") as f:
    json.dump(out, f, indent=1)
print(json.dumps(out, indent=1))
