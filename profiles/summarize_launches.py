"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count / total / mean / share.

    python profiles/summarize_launches.py gpurun_out/launches.csv [--skip N] > profiles/launches_rNN.md

Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes (B200_PROFILING.md).
"""
import csv
import re
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    skip = int(sys.argv[sys.argv.index('--skip') + 1]) if '--skip' in sys.argv else 0
    rows = []
    with open(path, newline='') as f:
        lines = [l for l in f if not l.startswith('==')]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        val = float(r['Metric Value'].replace(',', ''))
        unit = r.get('Metric Unit', 'ns')
        ns = val * {'ns': 1, 'us': 1e3, 'ms': 1e6, 's': 1e9}.get(unit, 1)
        rows.append((int(r['ID']), r['Kernel Name'], ns))
    rows = rows[skip:]
    agg = defaultdict(lambda: [0, 0.0])
    for _, name, ns in rows:
        short = re.sub(r'\(.*', '', name)
        short = re.sub(r'void |at::native::|<unnamed>::|\(anonymous namespace\)::', '', short)[:90]
        agg[short][0] += 1
        agg[short][1] += ns
    total = sum(v[1] for v in agg.values())
    ours = sum(v[1] for k, v in agg.items() if re.match(r'k_[a-z_]+', k))
    print(f'launches: {len(rows)}  total device time: {total / 1e6:.3f} ms  (hand-written kernels: {ours / 1e6:.3f} ms, '
          f'{100 * ours / max(total, 1):.1f} %)\n')
    print('| kernel | launches | total ms | mean us | share |')
    print('|---|---:|---:|---:|---:|')
    for k, (c, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
        print(f'| `{k}` | {c} | {ns / 1e6:.3f} | {ns / c / 1e3:.2f} | {100 * ns / total:.1f} % |')


if __name__ == '__main__':
    main()
