"""Condense an .ncu-rep (one or more kernels, --set full --import-source on) into a small CSV for profiles/:
key raw metrics per kernel + the stall-reason mix of the warp samples + the hottest source lines.
usage: python tools/ncu_summary.py rep.ncu-rep out.csv [note]"""
import collections
import csv
import io
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__waves_per_multiprocessor',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__t_sector_hit_rate.pct',
        'smsp__sass_inst_executed_op_local_ld.sum']


def run(args):
    return subprocess.run(['ncu', '-i'] + args, capture_output=True, text=True).stdout


def main():
    rep, out = sys.argv[1], sys.argv[2]
    note = sys.argv[3] if len(sys.argv) > 3 else ''
    raw = list(csv.reader(io.StringIO(run([rep, '--page', 'raw', '--csv']))))
    hdr, units = raw[0], raw[1]
    mix = list(csv.reader(io.StringIO(run([rep, '--page', 'source', '--csv', '--print-source', 'sass,cuda']))))
    lines, stall = collections.defaultdict(list), collections.Counter()
    kern = cur = h = None
    for r in mix:
        if not r:
            continue
        if r[0] == 'File Path':
            cur = r[1].split('/')[-1]
        elif r[0] == 'Function Name':
            kern = r[1]
        elif r[0] == 'Line No':
            h = {x: i for i, x in enumerate(r)}
        elif h is not None and r[0] not in ('', 'Kernel Name'):
            def num(v):
                try:
                    return int(v)
                except ValueError:
                    return 0
            lines[kern].append((num(r[h['# Samples']]), num(r[h['Instructions Executed']]), cur, r[0], r[1].strip()[:90]))
            for x, i in h.items():
                if x.startswith('stall_') and 'Not Issued' not in x:
                    stall[(kern, x)] += num(r[i])
    with open(out, 'w', newline='') as f:
        w = csv.writer(f)
        if note:
            w.writerow(['note', note])
        for row in raw[2:]:
            name = row[hdr.index('Kernel Name')]
            w.writerow(['kernel', name])
            for k in KEYS:
                if k in hdr:
                    w.writerow([k, units[hdr.index(k)], row[hdr.index(k)]])
            full = next((k for k in lines if name.split('<')[0].split(' ')[-1] in k), None)
            if full:
                tot = sum(a[0] for a in lines[full]) or 1
                for (k2, x), v in sorted(stall.items(), key=lambda t: -t[1]):
                    if k2 == full and v > .02 * tot:
                        w.writerow([x, '% of warp samples', f'{100 * v / tot:.1f}'])
                w.writerow(['warp_instructions_executed', '', sum(a[1] for a in lines[full])])
                for a in sorted(lines[full], reverse=True)[:12]:
                    w.writerow(['hot_line', f'{100 * a[0] / tot:.1f}% samples / {a[1]} inst', f'{a[2]}:{a[3]}', a[4]])


if __name__ == '__main__':
    main()
