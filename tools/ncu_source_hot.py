"""Hottest SASS instructions of one kernel in an ncu report captured with --set full --import-source on:
samples per instruction with the dominant stall reasons, plus the totals per stall reason.
    python tools/ncu_source_hot.py gpurun_out/x.ncu-rep regex:kernel_name [top_n] [out.txt]"""
import csv, io, subprocess, sys

rep, kern = sys.argv[1], sys.argv[2]
top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '-k', kern], stdout=subprocess.PIPE, text=True).stdout
lines = out.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
name = lines[start - 1].split('","')[1] if start > 0 else kern
end = next((i for i in range(start + 1, len(lines)) if lines[i].startswith('"Kernel Name"')), len(lines))   # first launch only
rows = list(csv.reader(io.StringIO('\n'.join(lines[start:end]))))
hdr = rows[0]
col = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith('stall_') and not h.endswith('_not_issued')]
samples = col['# Samples']
body = [r for r in rows[1:] if len(r) == len(hdr)]
total = sum(int(r[samples]) for r in body) or 1
text = ['# %s : %s' % (rep, name[:100]), '# %d instructions, %d warp-stall samples' % (len(body), total), '']
agg = {h: sum(int(r[col[h]] or 0) for r in body) for h in stall_cols}
text.append('stall reason totals (share of samples):')
for h, v in sorted(agg.items(), key=lambda kv: -kv[1])[:10]:
  text.append('  %-28s %6.1f %%' % (h[6:], 100.0 * v / total))
text.append('')
text.append('%6s %7s  %-58s %s' % ('line', 'share', 'instruction', 'top stall reasons'))
for r in sorted(body, key=lambda r: -int(r[samples]))[:top_n]:
  n = int(r[samples])
  why = sorted(((int(r[col[h]] or 0), h[6:]) for h in stall_cols), reverse=True)[:3]
  text.append('%6d %6.1f%%  %-58s %s' % (body.index(r), 100.0 * n / total, r[col['Source']].strip()[:58],
                                        ', '.join('%s %d' % (b, a) for a, b in why if a)))
text = '\n'.join(text)
print(text)
if len(sys.argv) > 4:
  open(sys.argv[4], 'w').write(text + '\n')
