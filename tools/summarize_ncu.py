"""Turns ncu outputs brought back in gpurun_out/ into the small text summaries committed under profiles/.
  python tools/summarize_ncu.py launches <launches.csv> <out.md>
  python tools/summarize_ncu.py full <report.ncu-rep> <out.md>
  python tools/summarize_ncu.py traffic <metrics.csv> <out.json> [kernel regex]
      metrics.csv: ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv of one
      train_step; out.json: DRAM bytes per launch of the matching kernels (bench.py reports it as roofline.traffic)"""
import collections
import csv
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_active"]


def traffic(path, out, pattern="gemm_fwd_kernel|conv3x3_resident_kernel"):
  import json
  lines = [l for l in open(path) if not l.startswith("==")]
  r = csv.reader(lines)
  hdr = next(r)
  ix = {h: i for i, h in enumerate(hdr)}
  per = collections.defaultdict(dict)   # launch id -> metric -> value (bytes / ns)
  names = {}
  scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1.0, "us": 1e3, "ms": 1e6, "nsecond": 1.0,
           "usecond": 1e3, "msecond": 1e6}
  for row in r:
    if len(row) < len(hdr) or not re.search(pattern, row[ix["Kernel Name"]]):
      continue
    v = float(row[ix["Metric Value"]].replace(",", "")) * scale.get(row[ix["Metric Unit"]], 1.0)
    per[row[ix["ID"]]][row[ix["Metric Name"]]] = v
    names[row[ix["ID"]]] = re.sub(r"\(.*", "", row[ix["Kernel Name"]])
  n = len(per)
  rd = sum(m.get("dram__bytes_read.sum", 0.0) for m in per.values())
  wr = sum(m.get("dram__bytes_write.sum", 0.0) for m in per.values())
  ns = sum(m.get("gpu__time_duration.sum", 0.0) for m in per.values())
  json.dump({"source": path, "kernels": pattern, "launches": n, "dram_bytes_read": rd, "dram_bytes_write": wr,
             "dram_bytes_per_launch": (rd + wr) / max(n, 1), "gpu_time_ms_serialised": ns / 1e6,
             "note": "one train_step under ncu (--clock-control none): cold caches, serialised launches"},
            open(out, "w"), indent=1)
  print(open(out).read())


def launches(path, out):
  lines = [l for l in open(path) if not l.startswith("==")]
  r = csv.reader(lines)
  hdr = next(r)
  ix = {h: i for i, h in enumerate(hdr)}
  agg = collections.defaultdict(lambda: [0, 0.0])
  tot = 0.0
  for row in r:
    if len(row) < len(hdr):
      continue
    val = float(row[ix["Metric Value"]].replace(",", ""))
    unit = row[ix["Metric Unit"]]
    ms = val / 1e6 if unit.startswith("n") else (val / 1e3 if unit.startswith("u") else val)
    name = re.sub(r"\(.*", "", row[ix["Kernel Name"]])
    agg[name][0] += 1
    agg[name][1] += ms
    tot += ms
  with open(out, "w") as f:
    f.write(f"# ncu launch list ({path}): gpu__time_duration.sum per kernel, --clock-control none\n")
    f.write("# per-launch times are cold-cache and serialised: compare SHARES, not absolutes\n")
    f.write(f"total {tot:.3f} ms over {sum(v[0] for v in agg.values())} launches\n\n")
    f.write("| ms | share | launches | kernel |\n|---:|---:|---:|---|\n")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
      f.write(f"| {v[1]:.3f} | {100 * v[1] / tot:.1f}% | {v[0]} | {k[:100]} |\n")


def full(path, out):
  raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
  rows = list(csv.reader(raw.splitlines()))
  hdr, units = rows[0], rows[1]
  cols = [i for i, h in enumerate(hdr) if h in KEYS or h == "Kernel Name"]
  with open(out, "w") as f:
    f.write(f"# ncu --set full --clock-control none: {path}\n\n")
    for n, r in enumerate(rows[2:]):
      f.write(f"## launch {n}\n")
      for i in cols:
        f.write(f"- {hdr[i]} = {r[i]} {units[i]}\n")
      f.write("\n")


if __name__ == "__main__":
  {"launches": launches, "full": full, "traffic": traffic}[sys.argv[1]](*sys.argv[2:])
