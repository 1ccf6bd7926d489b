"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump by CUDA source line."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
cur_file, hdr = None, None
byline, byex, src = collections.Counter(), collections.Counter(), {}
stall = collections.defaultdict(collections.Counter)
KEYS = ('stall_long_sb', 'stall_no_inst', 'stall_wait', 'stall_short_sb', 'stall_barrier', 'stall_branch_resolving', 'stall_lg', 'stall_mio',
        'stall_math', 'stall_not_selected', 'stall_selected', 'stall_dispatch')
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path':
        cur_file = r[1].split('/')[-1]
        continue
    if len(r) > 2 and r[0] == 'Line No':
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr) - 2 or r[2] != '-':
        continue
    try:
        line, ns, ex = int(r[0]), int(r[hdr.index('# Samples')]), int(r[hdr.index('Instructions Executed')])
    except ValueError:
        continue
    key = (cur_file, line)
    src[key] = r[1]
    byline[key] += ns
    byex[key] += ex
    for k in KEYS:
        try:
            stall[key][k] += int(r[hdr.index(k)] or 0)
        except (ValueError, IndexError):
            pass
tot, totex = sum(byline.values()), sum(byex.values())
print('samples', tot, 'warp-instructions', totex)
allst = collections.Counter()
for k in stall.values():
    allst.update(k)
print('stalls:', ', '.join('%s=%d' % (k.replace('stall_', ''), v) for k, v in allst.most_common(8)))
for key, n in byline.most_common(top):
    st = stall[key]
    tops = ', '.join('%s=%d' % (k.replace('stall_', ''), v) for k, v in st.most_common(3))
    print('%5d %5.1f%% ex=%9d %5.1f%%  %s:%d  %s   [%s]' % (n, 100 * n / tot, byex[key], 100 * byex[key] / totex, key[0], key[1], src[key].strip()[:64], tops))
