"""Turns ncu outputs brought back in gpurun_out/ into the small text summaries committed under profiles/.
  python tools/summarize_ncu.py launches <launches.csv> <out.md>
  python tools/summarize_ncu.py full <report.ncu-rep> <out.md>"""
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
  {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
