"""Summarise an `ncu --page source --csv --print-source cuda,sass` export per source line.
usage: python scripts/ncu_lines.py prof_cs.csv [raw.csv]"""
import csv, collections, sys

def I(x):
    try:
        return int(x)
    except ValueError:
        return 0

def main():
    if len(sys.argv) > 2:
        rows = list(csv.reader(open(sys.argv[2])))
        h, r = rows[0], rows[-1]
        want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
                'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
                'smsp__inst_executed.sum', 'launch__registers_per_thread', 'lts__t_sector_hit_rate.pct',
                'smsp__issue_active.avg.pct_of_peak_sustained_active', 'launch__grid_size',
                'launch__shared_mem_per_block_dynamic']
        for i, n in enumerate(h):
            if n in want:
                print(n, rows[1][i], r[i])
            elif 'issue_stalled' in n and 'per_issue_active' in n and 'not_issued' not in n:
                try:
                    if float(r[i]) > 0.2:
                        print(n.replace('smsp__average_warps_issue_stalled_', 'stall ').replace('_per_issue_active.ratio', ''), r[i])
                except ValueError:
                    pass
    rows = list(csv.reader(open(sys.argv[1])))
    cur = None
    hdr = None
    agg = collections.defaultdict(lambda: [0, 0, collections.Counter()])
    src = {}
    for r in rows:
        if r and r[0] == 'File Path':
            cur = r[1].split('/')[-1]
            continue
        if r and r[0] == 'Function Name':
            continue
        if r and r[0] == 'Line No':
            hdr = r
            isamp = hdr.index('# Samples')
            iex = hdr.index('Instructions Executed')
            st = {n: i for i, n in enumerate(hdr) if n.startswith('stall_') and 'Not Issued' not in n}
            continue
        if not r or hdr is None:
            continue
        try:
            ln = int(r[0])
        except ValueError:
            continue
        if r[1]:
            src[(cur, ln)] = r[1]
        a = agg[(cur, ln)]
        a[0] += I(r[isamp])
        a[1] += I(r[iex])
        for n, i in st.items():
            if I(r[i]):
                a[2][n[6:]] += I(r[i])
    tot = sum(a[0] for a in agg.values())
    totex = sum(a[1] for a in agg.values())
    print('samples', tot, 'warp instructions', totex)
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:40]:
        print(f"{k[0]}:{k[1]}".ljust(28), f"{100 * a[0] / tot:5.1f}% inst {100 * a[1] / totex:5.1f}%",
              src.get(k, '').strip()[:64].ljust(64), dict(a[2].most_common(2)))

main()
