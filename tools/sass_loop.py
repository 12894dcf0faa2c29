#!/usr/bin/env python
"""Histogram of SASS opcodes inside the innermost loop of every kernel in a cuobjdump -sass
listing (loop = from the target of the last backward branch to that branch)."""
import collections, re, sys
txt = open(sys.argv[1]).read()
for f in re.split(r'\n\s*Function : ', txt)[1:]:
    name = f.split('\n')[0]
    ins = []
    for l in f.split('\n'):
        m = re.match(r'\s+/\*([0-9a-f]{4,5})\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)(.*?);', l)
        if m: ins.append((int(m.group(1), 16), m.group(3), m.group(4)))
    loops = []
    for a, op, rest in ins:
        if op.startswith('BRA'):
            t = re.search(r'0x([0-9a-f]+)', rest)
            if t and int(t.group(1), 16) < a: loops.append((int(t.group(1), 16), a))
    if not loops:
        print(name, 'no loop'); continue
    lo, hi = max(loops, key=lambda x: x[1] - x[0])
    c = collections.Counter(op for a, op, _ in ins if lo <= a <= hi)
    print(name, 'loop instrs', sum(c.values()), dict(c.most_common()))
