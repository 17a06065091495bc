"""Compact per-kernel table from `ncu -i X.ncu-rep --page raw --csv` (the .ncu-rep files themselves are tens of MB and
stay on the GPU box): duration, DRAM bytes and GB/s, DRAM / L2 / L1 / SM throughput %, tensor and fp64 pipe activity,
IPC, achieved warps, registers.  Usage: python tools/ncu_summary.py raw.csv [label ...] > profiles/ncu_r2_xxx.txt"""
import csv
import sys

COLS = [
    ('us', 'gpu__time_duration.sum'), ('rdMB', 'dram__bytes_read.sum'), ('wrMB', 'dram__bytes_write.sum'),
    ('dram%', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'), ('L2%', 'lts__throughput.avg.pct_of_peak_sustained_elapsed'),
    ('L1%', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed'), ('SM%', 'sm__throughput.avg.pct_of_peak_sustained_elapsed'),
    ('tensor%', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'),
    ('fp64%', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active'),
    ('ipc', 'sm__inst_executed.avg.per_cycle_active'), ('issue%', 'smsp__issue_active.avg.pct_of_peak_sustained_active'),
    ('warps', 'sm__warps_active.avg.per_cycle_active'), ('regs', 'launch__registers_per_thread'),
    ('grid', 'launch__grid_size'), ('block', 'launch__block_size'), ('L2hit%', 'lts__t_sector_hit_rate.pct'),
]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    labels = sys.argv[2:]
    print('# %-26s %s' % ('kernel', ' '.join('%8s' % c for c, _ in COLS)), ' GB/s(dram)')
    for k, r in enumerate(data):
        name = r[idx['Kernel Name']]
        short = name.split('(')[0].replace('void ', '').replace('dccn::', '')
        if len(short) > 40:
            short = short[:40]
        lab = labels[k] if k < len(labels) else short
        vals = []
        for c, key in COLS:
            v = r[idx[key]] if key in idx else ''
            try:
                f = float(v.replace(',', ''))
                if units[idx[key]] == 'ms' and c == 'us':
                    f *= 1e3
                vals.append('%8.1f' % f if abs(f) < 1e5 else '%8.3g' % f)
            except Exception:
                vals.append('%8s' % v[:8])
        try:
            us = float(r[idx['gpu__time_duration.sum']])
            if units[idx['gpu__time_duration.sum']] == 'ms':
                us *= 1e3
            scale = {'Mbyte': 1e6, 'Gbyte': 1e9, 'Kbyte': 1e3, 'byte': 1.0}
            rd = float(r[idx['dram__bytes_read.sum']]) * scale[units[idx['dram__bytes_read.sum']]]
            wr = float(r[idx['dram__bytes_write.sum']]) * scale[units[idx['dram__bytes_write.sum']]]
            gbs = '%8.0f' % ((rd + wr) / (us * 1e-6) / 1e9)
        except Exception:
            gbs = ''
        print('%-28s %s %s' % (lab[:28], ' '.join(vals), gbs))


if __name__ == '__main__':
    main()
