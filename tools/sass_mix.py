"""Instruction mix and top stall sites of one kernel from an `ncu --page source --csv` dump
(sass view): python tools/sass_mix.py gpurun_out/sass_X.csv [ntop]"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 6 and r[0].startswith('0x')]
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 30


def I(s):
    try:
        return int(s)
    except ValueError:
        return 0


tot_samp = sum(I(r[2]) for r in rows)
tot_inst = sum(I(r[5]) for r in rows)
print('samples', tot_samp, 'warp-insts', tot_inst, 'sass lines', len(rows))
mix = collections.Counter()
smp = collections.Counter()
for r in rows:
    op = [o for o in r[1].split() if not o.startswith('@')][0].split('.')[0]
    mix[op] += I(r[5])
    smp[op] += I(r[2])
for op, c in mix.most_common(28):
    print('%-10s inst %5.1f%%  samples %5.1f%%' % (op, 100 * c / tot_inst, 100 * smp[op] / tot_samp))
print('--- top stall instrs')
for i in sorted(range(len(rows)), key=lambda i: -I(rows[i][2]))[:ntop]:
    print('%5d %-72s samp %6s exec %8s' % (i, rows[i][1].strip()[:72], rows[i][2], rows[i][5]))
