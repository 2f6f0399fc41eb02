#!/usr/bin/env python
"""List the loops of one kernel in a cuobjdump -sass dump with an opcode histogram each (static view of the hot loops).

    cuobjdump -sass strkit_b200/libstrkit_b200.so > /tmp/all.sass
    python tools/sass_loops.py /tmp/all.sass 'dp_packed_kernelILi10' [min_len]
"""
import collections
import re
import sys

path, pat = sys.argv[1], sys.argv[2]
min_len = int(sys.argv[3]) if len(sys.argv) > 3 else 20
lines = open(path).read().splitlines()
start = [i for i, l in enumerate(lines) if "Function :" in l and pat in l][0]
end = next((i for i in range(start + 1, len(lines)) if "Function :" in lines[i]), len(lines))
ins = []  # (addr, text)
rx = re.compile(r"/\*([0-9a-f]{4,})\*/\s+(.*?);")
for l in lines[start:end]:
    m = rx.search(l)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
addr_idx = {a: i for i, (a, _) in enumerate(ins)}
print(f"{pat}: {len(ins)} instructions")
for i, (a, t) in enumerate(ins):
    m = re.search(r"BRA(?:\.\w+)*\s+(?:!?U?P\d+,\s*)?0x([0-9a-f]+)", t)
    if not m:
        continue
    tgt = int(m.group(1), 16)
    if tgt <= a and tgt in addr_idx and i - addr_idx[tgt] + 1 >= min_len:
        j = addr_idx[tgt]
        ops = collections.Counter()
        for _, tt in ins[j:i + 1]:
            tt = re.sub(r"^@!?U?P\d+\s+", "", tt)
            ops[tt.split()[0]] += 1
        print(f"loop {tgt:#07x}..{a:#07x} len={i - j + 1:4d}  " + ", ".join(f"{k}:{v}" for k, v in ops.most_common(14)))
