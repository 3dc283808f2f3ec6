"""Time-weighted tensor-pipe activity over the conv / GEMM launches of one train_step, from an ncu CSV with
gpu__time_duration.sum and sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active per launch
(tools/r02_profile_session2.sh). Usage: python tools/tensor_pipe_weighted.py <csv> <out.md>"""
import collections
import csv
import re
import sys


def main(path, out):
  lines = [l for l in open(path) if not l.startswith("==")]
  r = csv.reader(lines)
  hdr = next(r)
  ix = {h: i for i, h in enumerate(hdr)}
  per = collections.defaultdict(dict)
  names = {}
  for row in r:
    if len(row) < len(hdr):
      continue
    v = float(row[ix["Metric Value"]].replace(",", ""))
    unit = row[ix["Metric Unit"]]
    if row[ix["Metric Name"]].startswith("gpu__time"):
      v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(unit, 1.0)
    per[row[ix["ID"]]][row[ix["Metric Name"]]] = v
    names[row[ix["ID"]]] = re.sub(r"\(.*", "", row[ix["Kernel Name"]]).replace("void ", "").replace("xmc::", "")
  T, P = "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"
  agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
  for k, m in per.items():
    if T in m and P in m:
      a = agg[names[k]]
      a[0] += 1
      a[1] += m[T]
      a[2] += m[T] * m[P]
  tot_t = sum(a[1] for a in agg.values())
  tot_tp = sum(a[2] for a in agg.values())
  with open(out, "w") as f:
    f.write(f"# time-weighted tensor-pipe activity over the conv / GEMM launches of one train_step ({path})\n")
    f.write("# ncu --clock-control none, serialised cold-cache launches; weight = gpu__time_duration of the launch\n\n")
    f.write("| kernel | launches | us | tensor pipe % (time-weighted) |\n|---|---:|---:|---:|\n")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
      f.write(f"| {k} | {a[0]} | {a[1]:.0f} | {a[2] / a[1]:.1f} |\n")
    f.write(f"| **all** | {sum(a[0] for a in agg.values())} | {tot_t:.0f} | **{tot_tp / tot_t:.1f}** |\n")
  print(open(out).read())


if __name__ == "__main__":
  main(sys.argv[1], sys.argv[2])
