"""Aggregate an `ncu --page source --csv` export by SASS segments between block barriers and list the hottest instructions.
Usage: python tools/ncu_segments.py src.csv [min_pct]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
hdr, data = rows[1], rows[2:]
ia, isamp, iex = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
tot_s = sum(int(r[isamp]) for r in data)
tot_e = sum(int(r[iex]) for r in data)
print('total samples', tot_s, 'warp-instructions', tot_e, 'sass lines', len(data))
seg, cur = [], {'s': 0, 'e': 0, 'start': 0, 'ops': {}, 'st': {}}
for i, r in enumerate(data):
    toks = r[ia].split()
    op = toks[1] if toks[0].startswith('@') else toks[0]
    cur['s'] += int(r[isamp]); cur['e'] += int(r[iex])
    for j in stall_cols:
        if r[j] not in ('', '0'):
            cur['st'][hdr[j]] = cur['st'].get(hdr[j], 0) + int(r[j])
    key = op.split('.')[0]
    if key in ('MUFU', 'UBLKCP', 'SYNCS', 'HMMA', 'STAS', 'LDGSTS', 'SHFL', 'CALL', 'LDG', 'STG', 'LD', 'ST'):
        k2 = op if key in ('MUFU', 'SYNCS', 'LDG', 'LD', 'ST', 'STG') else key
        cur['ops'][k2] = cur['ops'].get(k2, 0) + 1
    if op.startswith('BAR') or op.startswith('EXIT'):
        cur['end'] = i; seg.append(cur); cur = {'s': 0, 'e': 0, 'start': i + 1, 'ops': {}, 'st': {}}
cur['end'] = len(data) - 1; seg.append(cur)
for sg in seg:
    if sg['s'] * 100 > thr * tot_s:
        top = sorted(sg['st'].items(), key=lambda kv: -kv[1])[:4]
        tops = ' '.join(f"{k[6:]}:{100 * v / tot_s:.1f}" for k, v in top)
        print(f"[{sg['start']:5d}-{sg['end']:5d}] samples {100 * sg['s'] / tot_s:5.1f}%  inst {100 * sg['e'] / tot_e:5.1f}%  | {tops} | {sg['ops']}")
print('--- hottest instructions')
hot = sorted(range(len(data)), key=lambda i: -int(data[i][isamp]))[:40]
for i in sorted(hot):
    r = data[i]
    st = sorted(((int(r[j]), hdr[j][6:]) for j in stall_cols if r[j] not in ('', '0')), reverse=True)[:2]
    print(f"{i:5d} {100 * int(r[isamp]) / tot_s:5.2f}%  {r[ia].strip()[:70]:70s} {st}")
