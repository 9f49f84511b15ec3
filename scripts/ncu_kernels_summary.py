"""ncu --csv --metrics <scripts/ncu_metrics.txt> log  ->  per-kernel table (markdown on stdout, JSON to argv[2] if given).
Run in the build container on a file brought back in gpurun_out/."""
import collections
import csv
import json
import re
import sys

PEAK_GBS = 6548.2   # MEASURED_PEAKS.json hbm_gbs


def num(x):
    try:
        return float(x.replace(",", ""))
    except Exception:
        return None


def to_base(value, unit):
    """durations -> ns, bytes -> bytes"""
    scale = {"ns": 1.0, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "nsecond": 1.0, "second": 1e9, "s": 1e9,
             "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit)
    return value * scale if scale else value


def main():
    rows = list(csv.reader(open(sys.argv[1], errors="ignore")))
    hdr = None
    per = collections.OrderedDict()
    for r in rows:
        if "Kernel Name" in r and "Metric Name" in r:
            hdr = r
            continue
        if not hdr or len(r) != len(hdr):
            continue
        d = dict(zip(hdr, r))
        name = re.sub(r"\(.*", "", d["Kernel Name"])
        name = re.sub(r"^void ", "", name)
        v = num(d["Metric Value"])
        if v is None:
            continue
        per.setdefault(name, collections.defaultdict(list))[d["Metric Name"]].append(to_base(v, d["Metric Unit"]))
    out = []
    for name, m in per.items():
        def avg(k):
            return sum(m[k]) / len(m[k]) if m.get(k) else None
        t_ns = avg("gpu__time_duration.sum")
        rd, wr = avg("dram__bytes_read.sum"), avg("dram__bytes_write.sum")
        gbs = (rd + wr) / t_ns if (t_ns and rd is not None and wr is not None) else None
        out.append({"kernel": name, "launches": len(m.get("gpu__time_duration.sum", [])), "avg_us": t_ns / 1e3 if t_ns else None,
                    "dram_read_MB": rd / 1e6 if rd is not None else None, "dram_write_MB": wr / 1e6 if wr is not None else None,
                    "dram_GBs": gbs, "frac_of_measured_hbm_peak": gbs / PEAK_GBS if gbs else None,
                    "dram_throughput_pct": avg("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                    "l2_hit_pct": avg("lts__t_sector_hit_rate.pct"), "l2_throughput_pct": avg("lts__throughput.avg.pct_of_peak_sustained_elapsed"),
                    "sm_throughput_pct": avg("sm__throughput.avg.pct_of_peak_sustained_elapsed"),
                    "warps_active_pct": avg("sm__warps_active.avg.pct_of_peak_sustained_active"),
                    "issue_active_pct": avg("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                    "long_scoreboard_stall": avg("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
                    "regs": avg("launch__registers_per_thread"), "grid": avg("launch__grid_size"), "block": avg("launch__block_size")})
    out.sort(key=lambda d: -(d["avg_us"] or 0) * d["launches"])
    f = lambda x, p="%.1f": "-" if x is None else p % x
    print("| kernel | launches | avg µs | DRAM rd/wr MB | DRAM GB/s (of 6548 measured) | DRAM thr % | L2 hit % | L2 thr % | SM thr % | warps act % | issue act % | long-sb stall | regs | grid×block |")
    print("|---|---|---|---|---|---|---|---|---|---|---|---|---|---|")
    for d in out:
        print("| `%s` | %d | %s | %s / %s | %s (%s) | %s | %s | %s | %s | %s | %s | %s | %s | %s×%s |" % (
            d["kernel"][:70], d["launches"], f(d["avg_us"]), f(d["dram_read_MB"], "%.2f"), f(d["dram_write_MB"], "%.2f"), f(d["dram_GBs"], "%.0f"),
            f(d["frac_of_measured_hbm_peak"], "%.2f"), f(d["dram_throughput_pct"]), f(d["l2_hit_pct"]), f(d["l2_throughput_pct"]), f(d["sm_throughput_pct"]),
            f(d["warps_active_pct"]), f(d["issue_active_pct"]), f(d["long_scoreboard_stall"], "%.2f"), f(d["regs"], "%.0f"), f(d["grid"], "%.0f"), f(d["block"], "%.0f")))
    if len(sys.argv) > 2:
        json.dump(out, open(sys.argv[2], "w"), indent=1)


if __name__ == "__main__":
    main()
