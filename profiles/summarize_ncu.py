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
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64 inst %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu inst %"),
    ("sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active", "adu inst %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu inst %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("smsp__inst_executed.sum", "warp inst"),
    ("l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum", "TMA load bytes"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
]
STALL = "smsp__pcsamp_warps_issue_stalled_"


def main():
    rep = sys.argv[1]
    title = sys.argv[2] if len(sys.argv) > 2 else rep
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    print(f"\n### {title}\n")
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith(STALL) and "not_issued" not in h]
    print("| kernel | " + " | ".join(n for _, n in KEYS) + " | top stall reasons (share of samples) |")
    print("|---|" + "---|" * (len(KEYS) + 1))
    traffic = {}
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
        st = []
        for i in stall_cols:
            try:
                st.append((float(r[i].replace(",", "")), hdr[i][len(STALL):]))
            except ValueError:
                pass
        tot = sum(x for x, _ in st) or 1.0
        top = ", ".join(f"{n} {100 * x / tot:.0f}%" for x, n in sorted(st, reverse=True)[:3])
        print(f"| {name} | " + " | ".join(cells) + f" | {top} |")
        try:
            def num(k):
                return float(r[idx[k]].replace(",", "")), units[idx[k]]
            def to_bytes(v, u):
                return v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
            rd, wr, du = num("dram__bytes_read.sum"), num("dram__bytes_write.sum"), num("gpu__time_duration.sum")
            t = traffic.setdefault(name, {"launches": 0, "dram_bytes": 0.0, "us": 0.0})
            t["launches"] += 1
            t["dram_bytes"] += to_bytes(*rd) + to_bytes(*wr)
            t["us"] += du[0] * {"us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6}.get(du[1], 1.0)
        except (KeyError, ValueError):
            pass
    if len(sys.argv) > 3:   # third argument: write the per-kernel DRAM traffic (bytes per launch) as JSON
        import json
        js = {k: {"launches": v["launches"], "dram_bytes_per_launch": v["dram_bytes"] / v["launches"],
                  "ncu_us_per_launch": v["us"] / v["launches"]} for k, v in traffic.items()}
        json.dump(js, open(sys.argv[3], "w"), indent=1)


if __name__ == "__main__":
    main()
