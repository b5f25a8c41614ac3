"""Count the SASS mnemonics that show a kernel uses the Blackwell tensor / TMA paths (B200_PROFILING.md:
tcgen05.mma -> UTC*MMA, tcgen05.ld -> LDTM, cp.async.bulk.tensor -> UTMALDG, tcgen05.commit -> UTCBAR,
tcgen05.alloc -> UTCATOMSWS, mbarrier -> SYNCS) per kernel of libhsgb200.so.  Runs without a GPU.
    python tools/sass_mnemonics.py [profiles/r1_sass_mnemonics.txt]"""
import collections, os, re, subprocess, sys

LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'hsg_b200', 'libhsgb200.so')
PAT = re.compile(r'\b(UTC[A-Z]*MMA|UTMALDG|UTMASTG|UBLKCP|UTCBAR|LDTM|STTM|UTCATOMSWS|SYNCS|HMMA|LDGSTS)\b')

sass = subprocess.run(['cuobjdump', '-sass', LIB], stdout=subprocess.PIPE, text=True).stdout
counts, cur = collections.OrderedDict(), None
for line in sass.splitlines():
  m = re.search(r'Function : (\S+)', line)
  if m:
    cur = m.group(1)
    counts[cur] = collections.Counter()
  elif cur:
    for x in PAT.findall(line):
      counts[cur][x] += 1
rows = ['# cuobjdump -sass hsg_b200/libhsgb200.so (sm_100a): static instruction counts per kernel; kernels without any of',
        '# these mnemonics (the HBM-bound CUDA-core kernels) are omitted.  HMMA (legacy mma.sync) appears nowhere.']
for k, v in counts.items():
  if not v:
    continue
  name = subprocess.run(['c++filt', k], stdout=subprocess.PIPE, text=True).stdout.strip().split('(')[0]
  rows.append('%-64s %s' % (name[-64:], ' '.join('%s=%d' % kv for kv in sorted(v.items()))))
text = '\n'.join(rows)
print(text)
if len(sys.argv) > 1:
  open(sys.argv[1], 'w').write(text + '\n')
