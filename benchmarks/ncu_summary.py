"""Turn `ncu --set full` reports (.ncu-rep) into the small per-kernel summary kept under profiles/
(metric rows, one column per captured kernel launch) -- the format bench.py's ncu_traffic() reads.

    python benchmarks/ncu_summary.py -o profiles/r02_ncu_full_x.csv --note "..." gpurun_out/a.ncu-rep [gpurun_out/b.ncu-rep ...]
"""
import argparse
import csv
import io
import subprocess

METRICS = """gpu__time_duration.sum dram__bytes_read.sum dram__bytes_write.sum
gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed sm__throughput.avg.pct_of_peak_sustained_elapsed
smsp__inst_executed.sum smsp__issue_active.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active l1tex__throughput.avg.pct_of_peak_sustained_active
l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed l1tex__data_pipe_lsu_wavefronts_mem_shared.sum
l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum lts__throughput.avg.pct_of_peak_sustained_elapsed lts__t_sector_hit_rate.pct
sm__warps_active.avg.pct_of_peak_sustained_active smsp__warps_eligible.avg.per_cycle_active launch__registers_per_thread
launch__grid_size launch__block_size launch__shared_mem_per_block_dynamic smsp__cycles_active.avg sm__cycles_elapsed.max""".split()
STALLS = ("barrier branch_resolving dispatch_stall drain lg_throttle long_scoreboard math_pipe_throttle membar mio_throttle misc "
          "no_instruction not_selected selected short_scoreboard sleeping tex_throttle wait").split()
METRICS += [f"smsp__average_warps_issue_stalled_{s}_per_issue_active.ratio" for s in STALLS]


def load(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    return [(dict(zip(hdr, r)), dict(zip(hdr, units))) for r in rows[2:]]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("-o", "--out", required=True)
    ap.add_argument("--note", default="ncu --set full --clock-control none")
    ap.add_argument("reports", nargs="+")
    a = ap.parse_args()
    launches = []
    for p in a.reports:
        launches += load(p)
    with open(a.out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["# " + a.note])
        names = [d["Kernel Name"] for d, _ in launches]
        w.writerow(["metric", "unit"] + names)
        w.writerow(["Kernel Name", ""] + names)
        for m in METRICS:
            if all(m not in d for d, _ in launches):
                continue
            unit = next((u.get(m, "") for d, u in launches if m in d), "")
            w.writerow([m, unit] + [d.get(m, "") for d, _ in launches])
    print(a.out, len(launches), "launches")


if __name__ == "__main__":
    main()
