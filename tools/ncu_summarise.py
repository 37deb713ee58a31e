"""Write profiles/ncu_summary.json from a `ncu --set full` capture of the headline kernel.

  ncu -i gpurun_out/<name>.ncu-rep --page raw --csv > raw.csv     (here or on the GPU box)
  python tools/ncu_summarise.py raw.csv "<capture description>" [algorithmic bytes per launch]

bench.py reads `dram_bytes_per_launch` from the summary for `roofline.traffic`.
"""
import csv
import json
import os
import sys

KEEP = [
  "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
  "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
  "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
  "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
  "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
  "smsp__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
  "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
  "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__grid_size",
  "launch__block_size", "sm__cycles_elapsed.avg.per_second", "lts__t_sector_hit_rate.pct",
  "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
  "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
  "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
  "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
  "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
  "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
  "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
  "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
]
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def main():
  raw, capture = sys.argv[1], sys.argv[2]
  algorithmic = int(sys.argv[3]) if len(sys.argv) > 3 else 260 * 4096 * 16384
  rows = list(csv.reader(open(raw)))
  header, units = rows[0], rows[1]
  # the longest launch of the capture is the bank kernel
  col = {name: i for i, name in enumerate(header)}
  tcol = next(i for i, name in enumerate(header) if name.endswith("gpu__time_duration.sum"))
  launch = max(rows[2:], key=lambda r: float(r[tcol].replace(",", "")))

  def find(metric):
    for name, i in col.items():
      if name == metric or name.endswith("." + metric):
        return launch[i].replace(",", ""), units[i]
    return None, None

  metrics = {}
  for m in KEEP:
    value, unit = find(m)
    if value is not None:
      metrics[m] = {"value": value, "unit": unit}
  rd = float(metrics["dram__bytes_read.sum"]["value"]) * SCALE[metrics["dram__bytes_read.sum"]["unit"]]
  wr = float(metrics["dram__bytes_write.sum"]["value"]) * SCALE[metrics["dram__bytes_write.sum"]["unit"]]
  out = {
    "kernel": launch[col["Kernel Name"]] + " grid " + launch[col["Grid Size"]],
    "capture": capture,
    "dram_bytes_per_launch": rd + wr, "dram_bytes_read": rd, "dram_bytes_write": wr,
    "algorithmic_bytes_per_launch": algorithmic,
    "metrics": metrics,
  }
  path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "ncu_summary.json")
  with open(path, "w") as fh:
    json.dump(out, fh, indent=1)
  print("wrote", path, "traffic/algorithmic = %.3f" % ((rd + wr) / algorithmic))


if __name__ == "__main__":
  main()
