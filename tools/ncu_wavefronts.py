"""Per-source-line shared-memory wavefronts / global L1 tag requests of one kernel in an ncu report.
usage: ncu_wavefronts.py report.ncu-rep kernel-regex [top]"""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--print-source', 'cuda,sass', '--csv', '--kernel-name',
                      'regex:' + kern, '--launch-count', '1'], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
cur = None; hdr = None; agg = {}
for r in rows:
    if len(r) == 2 and r[0] == 'File Path':
        cur = r[1].split('/')[-1]; continue
    if len(r) > 7 and r[0] == 'Line No':
        hdr = r
        iw, ii, ig, ie = hdr.index('L1 Wavefronts Shared'), hdr.index('L1 Wavefronts Shared Ideal'), hdr.index('L1 Tag Requests Global'), hdr.index('Instructions Executed')
        continue
    if hdr and len(r) > iw and r[0].isdigit() and r[2] == '-':
        def f(x):
            try: return float(x)
            except ValueError: return 0.0
        agg[(cur, int(r[0]))] = (f(r[iw]), f(r[ii]), f(r[ig]), f(r[ie]), r[1][:80])
tw = sum(v[0] for v in agg.values()); tg = sum(v[2] for v in agg.values())
print(f'total shared wavefronts {tw:.3e} (ideal {sum(v[1] for v in agg.values()):.3e}), global tag requests {tg:.3e}')
for k, v in sorted(agg.items(), key=lambda kv: -(kv[1][0] + kv[1][2]))[:top]:
    print(f'{v[0]/tw*100:5.1f}% shw (x{v[0]/max(v[1],1):4.1f} ideal) {v[2]/max(tg,1)*100:5.1f}% gtag  {k[0]}:{k[1]:<4} {v[4]}')
