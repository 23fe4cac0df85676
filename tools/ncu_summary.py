"""Turns ncu-rep files into a compact markdown table (profiles/*.md).  Usage: ncu_summary.py out.md rep1 [rep2 ...]"""
import csv
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__grid_size", "grid"),
    ("sm__cycles_elapsed.max", "SM cycles"),
]


def main():
    out, reps = sys.argv[1], sys.argv[2:]
    lines = ["| kernel | " + " | ".join(n for _, n in WANT) + " |", "|---|" + "---|" * len(WANT)]
    for rep in reps:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        hdr, units = rows[0], rows[1]
        seen = set()
        for r in rows[2:]:
            name = r[hdr.index("Kernel Name")]
            short = name.split("(")[0].replace("void ", "").replace("b200::", "")
            key = (short, r[hdr.index("launch__grid_size")] if "launch__grid_size" in hdr else "")
            dur = r[hdr.index("gpu__time_duration.sum")]
            if (key, dur[:3]) in seen:
                continue
            seen.add((key, dur[:3]))
            cells = []
            for m, _ in WANT:
                if m in hdr:
                    i = hdr.index(m)
                    v = r[i]
                    try:
                        v = f"{float(v.replace(',', '')):.4g}"
                    except ValueError:
                        pass
                    cells.append(f"{v} {units[i]}".strip())
                else:
                    cells.append("-")
            lines.append(f"| `{short}` ({rep.split('/')[-1]}) | " + " | ".join(cells) + " |")
    open(out, "w").write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
