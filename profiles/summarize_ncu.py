"""Turns an ncu report (.ncu-rep, read with `ncu -i ... --page raw --csv`) into the markdown table kept under
profiles/.  Usage: python profiles/summarize_ncu.py <report.ncu-rep> [title] >> profiles/<file>.md"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm %peak"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64 pipe %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("smsp__inst_executed.sum", "warp inst"),
    ("l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum", "TMA load bytes"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_sb"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_sb"),
]


def main():
    rep = sys.argv[1]
    title = sys.argv[2] if len(sys.argv) > 2 else rep
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    print(f"\n### {title}\n")
    print("| kernel | " + " | ".join(n for _, n in KEYS) + " |")
    print("|---|" + "---|" * len(KEYS))
    for r in data:
        name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "").replace("tmb::", "")
        cells = []
        for k, _ in KEYS:
            if k in idx:
                v = r[idx[k]].replace(",", "")
                try:
                    f = float(v)
                    v = f"{f:.4g}"
                except ValueError:
                    pass
                cells.append(f"{v} {units[idx[k]]}".strip())
            else:
                cells.append("-")
        print(f"| {name} | " + " | ".join(cells) + " |")


if __name__ == "__main__":
    main()
