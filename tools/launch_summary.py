#!/usr/bin/env python
"""Summarise an `ncu --metrics ... --csv` launch list: per-launch lines for the first wave and per-kernel totals."""
import collections, csv, sys
def main(path, limit=40):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    iK, iM, iV, iID = hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Value'), hdr.index('ID')
    L = collections.OrderedDict()
    for r in rows[1:]:
        L.setdefault(r[iID], {'k': r[iK].split('(')[0].replace('void ', '').replace('nrcu::', '')})[r[iM]] = float(r[iV].replace(',', ''))
    tot = collections.OrderedDict(); n = 0
    for id_, d in L.items():
        k = d['k']
        t = d.get('gpu__time_duration.sum', 0) / 1e3
        a = tot.setdefault(k, [0, 0.0, 0.0]); a[0] += 1; a[1] += t; a[2] += d.get('smsp__inst_executed.sum', 0)
        if n < limit and any(x in k for x in ('k_big', 'k_trace', 'k_shade', 'k_raygen', 'k_accum')):
            print(f"{id_:>4} {k[:18]:18s} {t:8.1f} us inst {d.get('smsp__inst_executed.sum',0)/1e6:7.1f}M thr/inst {d.get('smsp__thread_inst_executed_per_inst_executed.ratio',0):5.1f} "
                  f"issue {d.get('smsp__issue_active.avg.pct_of_peak_sustained_active',0):5.1f}% dram R {d.get('dram__bytes_read.sum',0)/1e6:7.1f} W {d.get('dram__bytes_write.sum',0)/1e6:7.1f} MB")
            n += 1
    print("---- totals over the captured launches")
    T = sum(a[1] for a in tot.values())
    for k, a in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[:28]:28s} launches {a[0]:4d}  time {a[1]:9.1f} us ({a[1]/T*100:5.1f}%)  inst {a[2]/1e6:8.1f}M")
if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
