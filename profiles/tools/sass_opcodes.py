#!/usr/bin/env python
"""Per-kernel counts of the SASS opcodes that prove which hardware path a kernel uses
(B200_PROFILING.md): UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG = TMA tensor load,
UBLKCP = cp.async.bulk, HMMA = mma.sync, LDGSTS = cp.async, SHFL = warp shuffle.

    python profiles/tools/sass_opcodes.py [lib.so] > profiles/rNN_sass_opcodes.txt
"""
import collections
import re
import subprocess
import sys

OPS = ['UTCHMMA', 'LDTM', 'UTMALDG', 'UBLKCP', 'HMMA', 'LDGSTS', 'SHFL']
so = sys.argv[1] if len(sys.argv) > 1 else 'curla_b200/libcurla_b200.so'
sass = subprocess.run(['cuobjdump', '-sass', so], capture_output=True, text=True).stdout
counts, name = collections.defaultdict(collections.Counter), None
for line in sass.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        name = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
        name = name.replace('(anonymous namespace)::', '')
        name = re.sub(r'\(.*', '', name).replace('curla::', '').replace('void ', '')
        counts[name]
        continue
    if name:
        for op in re.findall(r'\b([A-Z][A-Z0-9]+)(?:\.[A-Z0-9_.]+)?\b', line):
            if op in OPS:
                counts[name][op] += 1
print('%-58s' % 'kernel' + ''.join('%9s' % o for o in OPS))
for k in sorted(counts):
    print('%-58s' % k[:58] + ''.join('%9d' % counts[k][o] for o in OPS))
