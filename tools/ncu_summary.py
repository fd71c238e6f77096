"""Condense an ncu report (``ncu -i X.ncu-rep --page raw --csv``) into a small per-kernel table.

    python tools/ncu_summary.py gpurun_out/prof_r01a.ncu-rep profiles/r01_ncu_summary.md
"""
import csv
import io
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1%"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma%"),
    ("sm__inst_executed_pipe_xu.sum", "xu_inst"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "ld_sectors"),
    ("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "ld_requests"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum", "red_sectors"),
    ("l1tex__t_requests_pipe_lsu_mem_global_op_red.sum", "red_requests"),
    ("lts__t_sectors_op_red.sum", "l2_red_sectors"),
    ("l1tex__t_sector_hit_rate.pct", "l1_hit%"),
    ("lts__t_sector_hit_rate.pct", "l2_hit%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy%"),
    ("launch__registers_per_thread", "regs"),
    ("smsp__inst_executed.sum", "warp_inst"),
]


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    seen = {}
    for r in data:
        name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "")
        seen.setdefault(name, []).append(r)
    with open(out, "w") as f:
        f.write(f"# ncu --set full summary of `{rep}` (per kernel, mean over captured launches)\n\n")
        for name, rs in seen.items():
            f.write(f"## {name}  ({len(rs)} launches)\n\n| metric | value | unit |\n|---|---:|---|\n")
            for key, label in METRICS:
                if key not in idx:
                    continue
                vals = []
                for r in rs:
                    try:
                        vals.append(float(r[idx[key]].replace(",", "")))
                    except ValueError:
                        pass
                if vals:
                    f.write(f"| {label} (`{key}`) | {sum(vals) / len(vals):,.3f} | {units[idx[key]]} |\n")
            f.write("\n")
    print("wrote", out)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
