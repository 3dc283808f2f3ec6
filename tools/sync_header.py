"""Developer aid: rewrites every prototype in include/xmc.h with the argument list of the matching `extern "C"`
definition in csrc/*.cu (comments in the header are kept). Also reports functions defined but not declared."""
import glob
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
hdr_path = os.path.join(ROOT, "include", "xmc.h")
hdr = open(hdr_path).read()
defs = {}
for f in glob.glob(os.path.join(ROOT, "xmcgan_image_generation_b200", "csrc", "*.cu")):
  src = open(f).read()
  for m in re.finditer(r'extern "C" (int|const char\*) (xmc_\w+)\(([^)]*)\)\s*\{', src):
    defs[m.group(2)] = (m.group(1), " ".join(m.group(3).split()))


def wrap(ret, name, args, width=118):
  head = f"{ret} {name}("
  parts = [a.strip() for a in args.split(",")] if args.strip() else ["void"]
  lines, cur = [], head
  indent = " " * len(head)
  for i, a in enumerate(parts):
    piece = a + ("," if i + 1 < len(parts) else ");")
    if len(cur) + len(piece) + (0 if cur.endswith("(") else 1) > width:
      lines.append(cur.rstrip())
      cur = indent + piece
    else:
      cur += ("" if cur.endswith("(") else " ") + piece
  lines.append(cur)
  return "\n".join(lines)


missing = []
for name, (ret, args) in sorted(defs.items()):
  pat = re.compile(r"\b(int|const char\*)\s+" + name + r"\s*\([^;{]*?\)\s*;", re.S)
  if not pat.search(hdr):
    missing.append(name)
    continue
  hdr = pat.sub(lambda m: wrap(ret, name, args), hdr, count=1)
open(hdr_path, "w").write(hdr)
print("not declared in the header:", missing)
