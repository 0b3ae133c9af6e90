"""Summarise `ncu --set full` reports (.ncu-rep) into a markdown table: duration, DRAM bytes, issue rate, occupancy,
tensor-pipe activity, the warp-stall breakdown.  Needs only the ncu CLI (no GPU).

    python profiles/summarize_ncu.py gpurun_out/prof_*_r02.ncu-rep > profiles/ncu_full_r02.md
"""
import csv
import io
import subprocess
import sys

KEYS = [
    ('gpu__time_duration.sum', 'duration'),
    ('dram__bytes_read.sum', 'DRAM read'),
    ('dram__bytes_write.sum', 'DRAM write'),
    ('launch__grid_size', 'grid'),
    ('launch__block_size', 'block'),
    ('launch__registers_per_thread', 'regs/thread'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps active %'),
    ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue active %'),
    ('smsp__inst_executed.sum', 'warp instructions'),
    ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor pipe active %'),
    ('dram__throughput.avg.pct_of_peak_sustained_elapsed', 'DRAM throughput %'),
    ('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smem bank conflicts'),
]


def main():
    print('# ncu --set full summaries (one launch per kernel; cold-cache, serialised: read ratios, not absolutes)\n')
    for path in sys.argv[1:]:
        out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        if len(rows) < 3:
            print(f'## {path}: no data\n')
            continue
        hdr, units = rows[0], rows[1]
        for vals in rows[2:]:
            d = dict(zip(hdr, vals))
            u = dict(zip(hdr, units))
            print(f"## `{d.get('Kernel Name', '?')[:110]}`  ({path.split('/')[-1]})\n")
            print('| metric | value |')
            print('|---|---|')
            for k, name in KEYS:
                if d.get(k, '') != '':
                    print(f'| {name} (`{k}`) | {d[k]} {u.get(k, "")} |')
            stalls = [(h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), float(d[h]))
                      for h in hdr if 'issue_stalled' in h and h.endswith('per_issue_active.ratio') and d.get(h, '') not in ('', '0')]
            stalls.sort(key=lambda kv: -kv[1])
            print('| warps stalled per issue (top) | ' + ', '.join(f'{k} {v:.2f}' for k, v in stalls[:7]) + ' |')
            print()


if __name__ == '__main__':
    main()
