"""Join an `ncu --page source --csv` export (SASS view) with `nvdisasm -g` line markers:
instructions executed and stall samples per CUDA source line.
usage: sass_by_line.py source.csv dis.txt kernel_substr [top]"""
import collections
import csv
import re
import sys

src_csv, dis, sub = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
# address -> line from the disassembly of the kernel
line_of = {}
cur = None
inside = False
for l in open(dis):
    if l.startswith('.text.') and l.rstrip().endswith(':'):
        inside = sub in l
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(\S.*);', l)
    if m:
        line_of[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
base = None
inst = collections.Counter()
samp = collections.Counter()
stall = collections.defaultdict(collections.Counter)
tot = 0
for r in rows[2:]:
    try:
        a = int(r[ix['Address']], 16) if not r[ix['Address']].isdigit() else int(r[ix['Address']])
        n = int(r[ix['Instructions Executed']])
    except ValueError:
        continue
    if base is None:
        base = a            # the export carries absolute device addresses; the first row is offset 0
    key = line_of.get(a - base)
    inst[key] += n
    tot += n
    try:
        samp[key] += int(r[ix['# Samples']])
    except ValueError:
        pass
    for h in ('stall_long_sb', 'stall_short_sb', 'stall_wait', 'stall_math', 'stall_mio', 'stall_not_selected',
              'stall_selected', 'stall_branch_resolving', 'stall_no_inst', 'stall_barrier', 'stall_dispatch'):
        try:
            stall[key][h] += int(r[ix[h]])
        except (ValueError, KeyError):
            pass
ts = sum(samp.values())
print('total inst %d, samples %d, mapped addrs %d' % (tot, ts, len(line_of)))
for key, n in inst.most_common(top):
    st = ', '.join('%s %d' % (h[6:], v) for h, v in stall[key].most_common(3))
    print('%-28s inst %10d %5.1f%%  samples %6d %5.1f%%  [%s]' % (key, n, 100.0 * n / tot, samp[key], 100.0 * samp[key] / max(ts, 1), st))
