"""Digest of `ncu --set full` reports (run here, no GPU): one row per .ncu-rep with the metrics that
say what bounds the kernel.  python tools/ncu_digest.py gpurun_out/ncu > profiles/r02/ncu_summary.md"""
import csv
import glob
import io
import os
import subprocess
import sys

KEYS = {
    "gpu__time_duration.sum": "time_us",
    "dram__bytes_read.sum": "dram_rd",
    "dram__bytes_write.sum": "dram_wr",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pct",
    "sm__inst_executed_pipe_tensor.sum": "tensor_inst",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "occ_pct",
    "launch__registers_per_thread": "regs",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum": "smem_wavefronts",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
}


def to_bytes(v, unit):
  m = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
  return float(v) * m.get(unit, 1)


def main(d):
  rows = []
  for rep in sorted(glob.glob(os.path.join(d, "*.ncu-rep"))):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = list(csv.reader(io.StringIO(out)))
    if len(rd) < 3:
      continue
    head, units, vals = rd[0], rd[1], rd[2]
    rec = {"name": os.path.basename(rep)[:-8], "kernel": vals[head.index("Kernel Name")][:60]}
    for i, h in enumerate(head):
      if h in KEYS:
        v = vals[i].replace(",", "")
        try:
          v = float(v)
        except ValueError:
          continue
        if h.startswith("dram__bytes"):
          v = to_bytes(v, units[i])
        if h == "gpu__time_duration.sum":
          v = v * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(units[i], 1)
        rec[KEYS[h]] = v
    rows.append(rec)
  cols = ["name", "kernel", "time_us", "dram_GB/s", "dram_MB", "dram_pct", "tensor_pct", "issue_pct", "occ_pct",
          "regs", "grid", "block"]
  print("| " + " | ".join(cols) + " |")
  print("|" + "---|" * len(cols))
  for r in rows:
    by = r.get("dram_rd", 0) + r.get("dram_wr", 0)
    r["dram_MB"] = by / 1e6
    r["dram_GB/s"] = by / (r.get("time_us", 1) * 1e-6) / 1e9 if r.get("time_us") else 0
    print("| " + " | ".join(f"{r.get(c, ''):.1f}" if isinstance(r.get(c), float) else str(r.get(c, "")) for c in cols) + " |")


if __name__ == "__main__":
  main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/ncu")
