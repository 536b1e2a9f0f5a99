"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum]
--csv --log-file launches.csv <cmd>` launch list (one CSV row per metric per launch).
  python tools/launch_summary.py launches.csv "title" > launch_summary.md
The per-launch times are serialised and cold-cache: the kernels' SHARES of the total are what compares with
bench.py's own per-call breakdown, not the absolute times."""
import collections
import csv
import re
import sys

UNIT = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3,
        "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def short(name):
  name = re.sub(r"(\(anonymous namespace\)|<unnamed>)::", "", name)
  name = re.sub(r"\(.*$", "", name)            # drop the argument list
  return name.replace("|", "/")[:78]


def main(path, title):
  rows = [r for r in csv.reader(l for l in open(path, errors="replace") if l.startswith('"'))]
  head = rows[0]
  kn, mn, mu, mv, idc = (head.index(c) for c in ("Kernel Name", "Metric Name", "Metric Unit", "Metric Value", "ID"))
  ms, by, ids = collections.defaultdict(float), collections.defaultdict(float), collections.defaultdict(set)
  for r in rows[1:]:
    if len(r) != len(head):
      continue
    try:
      v = float(r[mv].replace(",", "")) * UNIT.get(r[mu], 1.0)
    except ValueError:
      continue
    k = short(r[kn])
    ids[k].add(r[idc])
    if r[mn] == "gpu__time_duration.sum":
      ms[k] += v
    elif r[mn].startswith("dram__bytes"):
      by[k] += v
  total = sum(ms.values())
  print(f"# {title}\n")
  print("(`gpu__time_duration.sum`" + (", `dram__bytes_read.sum + dram__bytes_write.sum`" if by else "")
        + "; serialised, cold-cache: compare SHARES)\n")
  print(f"total {total:.2f} ms over {sum(len(v) for v in ids.values())} launches, "
        f"`bv::` kernels {sum(v for k, v in ms.items() if 'bv::' in k) / total * 100:.1f} % of the time\n")
  print("| kernel | launches | time ms | share |" + (" DRAM GB | GB/s |" if by else ""))
  print("|---|---|---|---|" + ("---|---|" if by else ""))
  for k, v in sorted(ms.items(), key=lambda kv: -kv[1]):
    line = f"| {k} | {len(ids[k])} | {v:.2f} | {v / total * 100:.1f} % |"
    if by:
      line += f" {by[k] / 1e9:.2f} | {by[k] / 1e9 / (v * 1e-3):.0f} |"
    print(line)


if __name__ == "__main__":
  main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "ncu launch list")
