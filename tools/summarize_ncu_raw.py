"""Condense an `ncu --page raw --csv` export into one line per launch with the counters DESIGN.md cites."""
import csv
import sys

COLS = [("gpu__time_duration.sum", "us", 1e-3), ("dram__bytes_read.sum", "MB_rd", 1e-6), ("dram__bytes_write.sum", "MB_wr", 1e-6),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%", 1), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor%", 1),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%", 1), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%", 1),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%", 1), ("launch__registers_per_thread", "regs", 1),
        ("launch__grid_size", "grid", 1)]


def num(v):
    try:
        return float(v.replace(",", ""))
    except ValueError:
        return None


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print("# " + path)
    print(f"{'kernel':58s} " + " ".join(f"{n:>9s}" for _, n, _ in COLS))
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].replace("void ", "").replace("<unnamed>::", "")[:58]
        vals = []
        for key, _, scale in COLS:
            v = num(r[idx[key]]) if key in idx else None
            if v is not None and key == "gpu__time_duration.sum":
                u = units[idx[key]]
                v = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)
            elif v is not None and key.startswith("dram__bytes"):
                u = units[idx[key]]
                v = v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1e-6)
            vals.append("        -" if v is None else f"{v:9.1f}")
        print(f"{name:58s} " + " ".join(vals))


if __name__ == "__main__":
    for p in sys.argv[1:]:
        main(p)
