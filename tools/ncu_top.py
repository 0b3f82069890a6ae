"""Top stall-sample SASS lines per kernel from an ncu report: ncu_top.py report.ncu-rep [n]"""
import csv, subprocess, sys
rep = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 14
txt = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
kern, hdr, cur = None, None, []
def flush():
    if not cur: return
    i_s = hdr.index('# Samples'); i_src = hdr.index('Source')
    tot = sum(int(r[i_s]) for r in cur)
    print('== %s  total samples %d' % (kern, tot))
    for r in sorted(cur, key=lambda r: -int(r[i_s]))[:n]:
        print('  %6d %5.1f%%  %s' % (int(r[i_s]), 100.0 * int(r[i_s]) / max(tot, 1), r[i_src].strip()[:110]))
for r in rows:
    if len(r) >= 2 and r[0] == 'Kernel Name':
        flush(); kern = r[1]; cur = []; hdr = None; continue
    if r and r[0] == 'Address':
        hdr = r; continue
    if hdr and len(r) == len(hdr):
        cur.append(r)
flush()
