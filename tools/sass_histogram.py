"""SASS opcode evidence for the built library (no GPU needed): per kernel, the count of tensor-core / TMA / TMEM
instructions.  usage: python tools/sass_histogram.py [lib.so] > profiles/rNN_sass_histogram.md"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else 'convnet/pytorch_b200/libb200conv.so'
txt = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
KEYS = ['UTCHMMA', 'UTCHMMA.2CTA', 'UTMALDG', 'UTMALDG.*IM2COL', 'UTMASTG', 'LDTM', 'UTCBAR', 'UTCATOMSWS', 'SYNCS',
        'HMMA', 'LDGSTS', 'RED.E', 'ATOM']
per = collections.OrderedDict()
cur = None
for line in txt.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        cur = m.group(1)
        per[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.search(r'/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
    if not m:
        continue
    op = m.group(1)
    per[cur]['_total'] += 1
    for k in KEYS:
        if re.match(k.replace('.', r'\.').replace(r'\.*', '.*') + r'(\.|$)', op) or (k.endswith('IM2COL') and 'IM2COL' in op and op.startswith('UTMALDG')):
            per[cur][k] += 1


def demangle(n):
    m = re.match(r'_ZN4b200(\d+)(\w+)', n)
    if not m:
        return n
    return m.group(2)[:int(m.group(1))] + (' <' + n[-40:] + '>' if 'ILi' in n else '')


print('# SASS opcode histogram of `%s` (cuobjdump -sass, sm_100a)\n' % lib)
print('`UTCHMMA` = tcgen05.mma (`.2CTA` = cta_group::2), `UTMALDG`/`UTMASTG` = TMA tensor load / store (`IM2COL` = im2col '
      'mode), `LDTM` = tcgen05.ld (TMEM -> registers), `UTCBAR` = tcgen05.commit, `HMMA` = legacy mma.sync (must be 0).\n')
print('| kernel | instructions | ' + ' | '.join(KEYS) + ' |')
print('|---|---:|' + '---:|' * len(KEYS))
tot = collections.Counter()
for fn, c in per.items():
    if not any(c[k] for k in KEYS[:7]):
        continue
    print('| `%s` | %d | ' % (demangle(fn), c['_total']) + ' | '.join(str(c[k]) for k in KEYS) + ' |')
    tot.update(c)
print('| **all kernels with tensor/TMA instructions** | %d | ' % tot['_total'] + ' | '.join(str(tot[k]) for k in KEYS) + ' |')
allc = collections.Counter()
for c in per.values():
    allc.update(c)
print('\nWhole library: %d kernels, %d SASS instructions, HMMA (legacy tensor path) = %d.' % (len(per), allc['_total'], allc['HMMA']))
