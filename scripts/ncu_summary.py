"""Turns ncu outputs under gpurun_out/ into the committed summaries under profiles/ (run in the build container)."""
import collections
import csv
import json
import subprocess
import sys

METRICS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
           'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
           'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
           'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum', 'launch__registers_per_thread', 'launch__grid_size',
           'launch__block_size', 'launch__waves_per_multiprocessor', 'l1tex__m_xbar2l1tex_read_bytes.sum',
           'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
           'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
           'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
           'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
           'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio']


def launches(path):
    rows = list(csv.reader(open(path, errors='ignore')))
    hdr, data = None, []
    for r in rows:
        if 'Kernel Name' in r and 'Metric Value' in r:
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            data.append(dict(zip(hdr, r)))
    agg = collections.OrderedDict()
    for d in data:
        agg.setdefault(d['Kernel Name'].split('(')[0][-60:], []).append(float(d['Metric Value'].replace(',', '')))
    return agg


def raw(rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {'kernel': r[hdr.index('Kernel Name')][:80]}
        for m in METRICS:
            if m in hdr:
                d[m] = r[hdr.index(m)] + ' ' + units[hdr.index(m)]
        res.append(d)
    return res


if __name__ == '__main__':
    what = sys.argv[1]
    if what == 'launches':
        agg = launches(sys.argv[2])
        tot = sum(sum(v) for v in agg.values())
        print('| kernel | launches | avg us | share of listed time |\n|---|---|---|---|')
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            print('| `%s` | %d | %.1f | %.1f%% |' % (k, len(v), sum(v) / len(v) / 1e3, 100 * sum(v) / tot))
    else:
        print(json.dumps(raw(sys.argv[2]), indent=1))
