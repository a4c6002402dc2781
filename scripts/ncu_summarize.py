"""Turn `ncu -i rep --page raw --csv` into the markdown table kept under profiles/ (one row per captured launch).

usage: python scripts/ncu_summarize.py raw.csv [HBM_GBs=6574] [BF16_TFs=1457] > profiles/ncu_hot_kernels_rX.md"""
import csv
import re
import sys

raw = sys.argv[1]
hbm = float(sys.argv[2]) if len(sys.argv) > 2 else 6574.0
tfs = float(sys.argv[3]) if len(sys.argv) > 3 else 1457.0
rows = list(csv.reader(open(raw)))
h = rows[0]
units = rows[1]


def col(name):
    return h.index(name) if name in h else -1


def num(r, name, default=0.0):
    i = col(name)
    if i < 0 or i >= len(r):
        return default
    try:
        v = float(r[i].replace(",", ""))
    except ValueError:
        return default
    return v


def scaled(r, name):
    """value converted to base units using the unit row (Kbyte/Mbyte/Gbyte, us/ms/ns)."""
    i = col(name)
    if i < 0:
        return 0.0
    u = units[i].lower()
    v = num(r, name)
    mult = {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12, "ns": 1e-9, "us": 1e-6, "usecond": 1e-6,
            "ms": 1e-3, "msecond": 1e-3, "nsecond": 1e-9, "second": 1.0, "s": 1.0}
    return v * mult.get(u, 1.0)


print("| # | kernel | time (us) | DRAM rd+wr (MB) | DRAM GB/s | % HBM peak | tensor pipe % | SM busy % | L2 hit % | ach. occupancy % | regs | grid x block |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|")
for n, r in enumerate(rows[2:]):
    if len(r) < len(h) // 2:
        continue
    name = re.sub(r"\(.*", "", r[col("Kernel Name")]).replace("void ", "").replace("im::", "")
    t = scaled(r, "gpu__time_duration.sum")
    rd, wr = scaled(r, "dram__bytes_read.sum"), scaled(r, "dram__bytes_write.sum")
    gbs = (rd + wr) / t / 1e9 if t > 0 else 0.0
    tens = num(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
               num(r, "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed"))
    smb = num(r, "sm__throughput.avg.pct_of_peak_sustained_elapsed")
    l2 = num(r, "lts__t_sector_hit_rate.pct")
    occ = num(r, "sm__warps_active.avg.pct_of_peak_sustained_active")
    regs = int(num(r, "launch__registers_per_thread"))
    grid = r[col("Grid Size")].strip("()").split(",")[0:3] if col("Grid Size") >= 0 else ["?"]
    blk = r[col("Block Size")].strip("()").split(",")[0] if col("Block Size") >= 0 else "?"
    g = "x".join(x.strip() for x in grid if x.strip() not in ("1",)) or "1"
    print(f"| {n} | `{name}` | {t * 1e6:.1f} | {(rd + wr) / 1e6:.1f} | {gbs:.0f} | {100 * gbs / hbm:.0f} | {tens:.1f} | {smb:.1f} | "
          f"{l2:.0f} | {occ:.0f} | {regs} | {g}x{blk.strip()} |")
print(f"\nPeaks used: HBM {hbm:.0f} GB/s, bf16 {tfs:.0f} TF/s sustained (MEASURED_PEAKS.json).  Durations are ncu single-kernel replays "
      "(cold L2, serialised) — shares and pipe utilisation only; CUDA-event numbers live in bench_*.json.")
