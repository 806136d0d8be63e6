"""Hot source lines of one kernel from `ncu -i rep --page source --print-source cuda,sass --csv --kernel-name regex:X`:
python tools/src_hot.py dump.csv [ntop]  -> per source line: share of stall samples, top stall reasons, instructions."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr, cur, out = None, None, []
for r in rows:
    if r and r[0] == 'File Path':
        cur = r[1].split('/')[-1]
    elif r and r[0] == 'Line No':
        hdr = r
    elif hdr and len(r) > 8 and r[2] == '-':
        out.append((cur, r))


def I(s):
    try:
        return int(s)
    except ValueError:
        return 0


stall_cols = [(i, h.replace('stall_', '')) for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
tot = sum(I(r[4]) for _, r in out)
agg = {}
for _, r in out:
    for i, h in stall_cols:
        agg[h] = agg.get(h, 0) + I(r[i])
print('samples', tot, ' by reason:', ', '.join('%s %.1f%%' % (h, 100 * v / tot) for h, v in sorted(agg.items(), key=lambda kv: -kv[1])[:10]))
for f, r in sorted(out, key=lambda o: -I(o[1][4]))[:ntop]:
    top = sorted(((I(r[i]), h) for i, h in stall_cols), reverse=True)[:3]
    print('%-15s %4s %-70s %5.2f%%  %-38s inst %d' % (f, r[0], r[1].strip()[:70], 100 * I(r[4]) / tot,
          ' '.join('%s:%d' % (h, v) for v, h in top if v), I(r[7])))
