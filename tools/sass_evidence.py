#!/usr/bin/env python
"""SASS evidence for profiles/: per kernel of libgims_b200.so the counts of the Blackwell-native instructions.
usage: python tools/sass_evidence.py > profiles/rNN_sass_tensor_kernels.txt"""
import collections
import re
import subprocess
import sys

LIB = sys.argv[1] if len(sys.argv) > 1 else 'gims_b200/lib/libgims_b200.so'
out = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(['c++filt', n], capture_output=True, text=True).stdout.strip()
kern, counts, total = None, collections.defaultdict(collections.Counter), collections.Counter()
WANT = ['UTCHMMA', 'LDTM', 'STTM', 'UTMALDG', 'UBLKCP', 'UTCBAR', 'SYNCS', 'HMMA', 'HGMMA', 'REDG', 'RED.']
for line in out.splitlines():
    m = re.match(r'\s*Function : (\S+)', line)
    if m:
        kern = demangle(m.group(1))
        kern = kern.replace('(anonymous namespace)::', '').replace('void ', '').replace('gims::', '')
        kern = re.sub(r'\(.*', '', kern)
        continue
    m = re.match(r'\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
    if m and kern:
        op = m.group(1)
        total[kern] += 1
        for w in WANT:
            if op.startswith(w) and not (w == 'HMMA' and op.startswith('HMMA') is False):
                counts[kern][w] += 1
print('cuobjdump -sass %s (sm_100a), per kernel: counts of the Blackwell-native instructions' % LIB)
print('(UTCHMMA = tcgen05.mma kind::tf32 / kind::f16, LDTM / STTM = tcgen05.ld / st, UTMALDG = cp.async.bulk.tensor, UBLKCP = cp.async.bulk,')
print(' UTCBAR = tcgen05.commit, SYNCS = mbarrier; HMMA / HGMMA (legacy / Hopper tensor paths) must be absent)\n')
for k in sorted(total):
    c = counts[k]
    if not (c['UTCHMMA'] or c['UTMALDG'] or c['UBLKCP'] or 'sinkhorn' in k):
        continue
    print('%-44s UTCHMMA %3d  LDTM %3d  STTM %3d  UTMALDG %3d  UBLKCP %2d  UTCBAR %3d  SYNCS %3d  HMMA %d  HGMMA %d  instructions %d' %
          (k[:44], c['UTCHMMA'], c['LDTM'], c['STTM'], c['UTMALDG'], c['UBLKCP'], c['UTCBAR'], c['SYNCS'], c['HMMA'], c['HGMMA'], total[k]))
