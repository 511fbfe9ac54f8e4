"""ncu launch list (csv, --metrics gpu__time_duration.sum) -> per-kernel table of ONE bench step (markdown on stdout).
Usage: python tools/launch_breakdown.py gpurun_out/<file>.csv [first kernel name substring]"""
import csv
import sys
from collections import OrderedDict

path = sys.argv[1]
first = sys.argv[2] if len(sys.argv) > 2 else "fake_quantize_multi_fwd"
rows = [r for r in csv.reader(open(path)) if len(r) > 5]
hdr = rows[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
launches = []
for r in rows[1:]:
    v = float(r[iv].replace(",", ""))
    v = v / 1000.0 if r[iu] in ("ns", "nsecond") else v
    launches.append((r[ik], v))
starts = [i for i, (k, _) in enumerate(launches) if first in k]
lo, hi = starts[-2], starts[-1]  # the last complete step
step = launches[lo:hi]
agg = OrderedDict()
for k, v in step:
    name = k.split("(")[0].replace("void ", "").replace("gsb200::", "")
    if name.startswith("at::") or "elementwise" in name or "gemv" in name:
        name = "torch elementwise / small ops"
    a = agg.setdefault(name, [0.0, 0])
    a[0] += v
    a[1] += 1
total = sum(v for v, _ in agg.values())
print(f"{len(step)} kernel launches, {total:.1f} us of kernel time\n")
print("| time | share | launches | kernel |\n|---|---|---|---|")
for name, (v, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"| {v:.1f} us | {100 * v / total:.1f} % | {n} | `{name}` |")
