"""Summarise an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` launch list of a
short bench.py run: per-kernel launch count, time, share, DRAM bytes; and the per-launch DRAM traffic of the tensor-core
GEMM kernel that bench.py reports as roofline.traffic.
usage: summarize_launches.py launches.csv out.txt out.json "<command>" <per_gpu_batch> <steps incl. warm-up>"""
import collections
import csv
import json
import re
import sys

src, out_txt, out_json, command, batch, steps = sys.argv[1:7]
precision = sys.argv[7] if len(sys.argv) > 7 else "bf16"
steps = int(steps)
rows = list(csv.reader(l for l in open(src) if l.startswith('"')))
ix = {h: i for i, h in enumerate(rows[0])}
per = collections.OrderedDict()
for r in rows[1:]:
    d = per.setdefault(r[ix["ID"]], {"name": r[ix["Kernel Name"]]})
    d[r[ix["Metric Name"]]] = float(r[ix["Metric Value"]].replace(",", ""))
unit = {r[ix["Metric Name"]]: r[ix["Metric Unit"]] for r in rows[1:]}
tscale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3}[unit["gpu__time_duration.sum"]]
bscale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def short(n):
    n = re.sub(r"^void ", "", n)
    return n.split("(")[0][:66]


agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for d in per.values():
    a = agg[short(d["name"])]
    a[0] += 1
    a[1] += d.get("gpu__time_duration.sum", 0.0) * tscale
    a[2] += d.get("dram__bytes_read.sum", 0.0) * bscale.get(unit.get("dram__bytes_read.sum", "byte"), 1.0)
    a[3] += d.get("dram__bytes_write.sum", 0.0) * bscale.get(unit.get("dram__bytes_write.sum", "byte"), 1.0)
total = sum(a[1] for a in agg.values())
lines = ["ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none : " + command,
         "(%d steps incl. warm-up + synthetic data generation by torch; per-launch times are cold-cache and serialised: compare SHARES)" % steps,
         "%d launches, %.2f ms kernel time" % (len(per), total), "",
         "%-66s %6s %10s %7s %10s %10s" % ("kernel", "n", "ms", "share", "rd GB", "wr GB")]
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    lines.append("%-66s %6d %10.3f %6.1f%% %10.2f %10.2f" % (k, a[0], a[1], 100 * a[1] / total, a[2] / 1e9, a[3] / 1e9))
open(out_txt, "w").write("\n".join(lines) + "\n")
gemm = [a for k, a in agg.items() if "conv_gemm_tc" in k]
ours = [a for k, a in agg.items() if k.startswith("sg::") or k.startswith("pair::") or "conv_gemm_tc" in k]
n = sum(a[0] for a in gemm)
byt = sum(a[2] + a[3] for a in gemm)
ms = sum(a[1] for a in gemm)
json.dump({"command": command, "per_gpu_batch": int(batch), "precision": precision, "gemm_launches_per_step": n / steps,
           "gemm_dram_bytes_per_step": byt / steps, "gemm_dram_bytes_per_launch": byt / max(n, 1),
           "gemm_ncu_ms_per_step": ms / steps,
           "gemm_share_of_engine_kernel_time": ms / max(sum(a[1] for a in ours), 1e-9),
           "source": out_txt + " (ncu dram__bytes_read.sum + dram__bytes_write.sum)",
           "note": "share = conv_gemm_tc time / (sg:: + pair:: kernel time) in the ncu launch list"}, open(out_json, "w"), indent=1)
print("\n".join(lines[:16]))
