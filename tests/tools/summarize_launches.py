"""ncu launch list (csv, --metrics gpu__time_duration.sum[,dram__bytes_*]) -> per-kernel table of the LAST step.
usage: python tests/tools/summarize_launches.py launches.csv [first-kernel-of-a-step (default k_stem_split)]"""
import csv, collections, re, sys


def load(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith('==')]
    rows = collections.OrderedDict()
    for row in csv.DictReader(lines):
        i = int(row['ID'])
        k = re.sub(r'\(.*', '', row['Kernel Name']).replace('void ', '')
        k = re.sub(r'^.*::', '', k) if ('GLOBAL' in k or 'unnamed' in k or 'sdnms' in k) else k
        v = float(row['Metric Value'].replace(',', ''))
        if 'dram' in row['Metric Name']:
            v *= {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(row['Metric Unit'], 1)
        rows.setdefault(i, {'k': k})[row['Metric Name']] = v
    return rows


def main():
    rows = load(sys.argv[1])
    first = sys.argv[2] if len(sys.argv) > 2 else 'k_stem_split'
    ids = sorted(rows)
    starts = [i for i in ids if rows[i]['k'].startswith(first)]
    s = starts[-1]
    agg = collections.OrderedDict()
    for i in ids:
        if i < s: continue
        r = rows[i]; a = agg.setdefault(r['k'][:56], [0, 0.0, 0.0])
        a[0] += 1; a[1] += r.get('gpu__time_duration.sum', 0); a[2] += r.get('dram__bytes_read.sum', 0) + r.get('dram__bytes_write.sum', 0)
    tot = sum(a[1] for a in agg.values())
    print("last step: %d launches, %.1f us of kernel time (ncu: serialised, cold caches; shares are what matters)\n" % (sum(a[0] for a in agg.values()), tot / 1000))
    print("| kernel | launches | us | share | DRAM MB |\n|---|---|---|---|---|")
    for k, (n, t, b) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| `%s` | %d | %.1f | %.1f %% | %s |" % (k, n, t / 1000, 100 * t / tot, ("%.2f" % (b / 1e6)) if b else ""))


if __name__ == "__main__":
    main()
