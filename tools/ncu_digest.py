"""Digest of an `ncu --page raw --csv` export (one row per profiled launch): per kernel, the LAST launch
(warm) with the metrics that say what bounds it.  python tools/ncu_digest.py zoo_raw.csv > ncu_summary.md"""
import csv
import sys

KEYS = {
    "gpu__time_duration.sum": "time_us",
    "dram__bytes_read.sum": "dram_rd",
    "dram__bytes_write.sum": "dram_wr",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pct",
    "sm__inst_executed_pipe_tensor_op_utcmma.sum": "utcmma_inst",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "occ_pct",
    "launch__registers_per_thread": "regs",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_pct",
    "l1tex__data_bank_reads.avg.pct_of_peak_sustained_elapsed": "smem_rd_pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed": "lsu_smem_pct",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active": "xu_pct",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active": "fma_pct",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active": "alu_pct",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": "st_long_sb",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio": "st_math_thr",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio": "st_mio_thr",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio": "st_barrier",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio": "st_short_sb",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio": "st_wait",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio": "st_no_inst",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio": "st_not_sel",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
}
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6,
        "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6}


def main(path):
  rd = list(csv.reader(open(path)))
  head, units, rows = rd[0], rd[1], rd[2:]
  kn = head.index("Kernel Name")
  last = {}
  for r in rows:
    if len(r) == len(head):
      last[r[kn]] = r                      # later launches overwrite earlier ones: keeps the warm one
  cols = ["kernel", "time_us", "dram_GB/s", "dram_MB", "dram_pct", "tensor_pct", "issue_pct", "xu_pct", "fma_pct",
          "alu_pct", "lsu_smem_pct", "occ_pct", "regs", "grid", "block", "st_long_sb", "st_short_sb", "st_math_thr",
          "st_mio_thr", "st_barrier", "st_wait", "st_no_inst", "st_not_sel"]
  print("| " + " | ".join(cols) + " |")
  print("|" + "---|" * len(cols))
  for name, r in sorted(last.items(), key=lambda kv: kv[0]):
    rec = {"kernel": name.replace("bv::(anonymous namespace)::", "").replace("|", "/")[:70]}
    for i, h in enumerate(head):
      if h in KEYS:
        try:
          v = float(r[i].replace(",", ""))
        except ValueError:
          continue
        rec[KEYS[h]] = v * UNIT.get(units[i], 1)
    by = rec.get("dram_rd", 0) + rec.get("dram_wr", 0)
    rec["dram_MB"] = by / 1e6
    rec["dram_GB/s"] = by / (rec["time_us"] * 1e-6) / 1e9 if rec.get("time_us") else 0.0
    print("| " + " | ".join(f"{rec[c]:.1f}" if isinstance(rec.get(c), float) else str(rec.get(c, "")) for c in cols) + " |")


if __name__ == "__main__":
  main(sys.argv[1])
