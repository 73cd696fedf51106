"""Opcode histogram of an address range of a cuobjdump -sass listing:  sass_count.py LISTING LO HI [SKIPLO-SKIPHI ...]"""
import collections
import re
import sys

ALU = {'LOP3', 'SHF', 'IADD3', 'ISETP', 'SEL', 'VIADD', 'VIADDMNMX', 'VIMNMX', 'LEA', 'PLOP3', 'PRMT', 'IABS', 'MOV'}
L = open(sys.argv[1]).read().splitlines()
lo, hi = int(sys.argv[2], 16), int(sys.argv[3], 16)
skip = [tuple(int(x, 16) for x in a.split('-')) for a in sys.argv[4:]]
c = collections.Counter()
for l in L:
    m = re.match(r'\s+/\*([0-9a-f]{4,6})\*/\s+(@!?U?P\d\s+)?(\S+)', l)
    if not m:
        continue
    a = int(m.group(1), 16)
    if a < lo or a > hi or any(s <= a <= e for s, e in skip):
        continue
    c[m.group(3).split('.')[0]] += 1
print(sum(c.values()), 'ALU', sum(v for k, v in c.items() if k in ALU))
print(sorted(c.items(), key=lambda x: -x[1]))
