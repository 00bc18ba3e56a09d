#!/usr/bin/env python
"""Print the SASS of the innermost loop(s) of a kernel that contain a given mnemonic, with an opcode histogram.
   usage: sass_loop.py <object or cubin> <kernel-substring> [mnemonic=SHFL.BFLY]"""
import collections, re, subprocess, sys
obj, kern = sys.argv[1], sys.argv[2]
needle = sys.argv[3] if len(sys.argv) > 3 else "SHFL.BFLY"
txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
cur, lines = None, []
for l in txt.splitlines():
    if "Function :" in l:
        cur = l
    elif cur and kern in cur:
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
        if m:
            lines.append((int(m.group(1), 16), m.group(2).strip()))
addr = {a: i for i, (a, _) in enumerate(lines)}
loops = []
for i, (a, ins) in enumerate(lines):
    m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d,\s*)?0x([0-9a-f]+)", ins)
    if m:
        t = int(m.group(1), 16)
        if t <= a and t in addr:
            loops.append((addr[t], i))
# innermost loops containing the needle
for s, e in loops:
    body = lines[s:e + 1]
    if not any(needle in ins for _, ins in body):
        continue
    if any(s2 >= s and e2 <= e and (s2, e2) != (s, e) and any(needle in ins for _, ins in lines[s2:e2 + 1]) for s2, e2 in loops):
        continue
    print(f"loop 0x{lines[s][0]:04x}..0x{lines[e][0]:04x}: {len(body)} instructions")
    h = collections.Counter()
    for _, ins in body:
        op = ins.split()[1] if ins.startswith("@") else ins.split()[0]
        h[op.split(".")[0]] += 1
    print("  " + ", ".join(f"{k} {v}" for k, v in h.most_common()))
    if "-v" in sys.argv:
        for a, ins in body:
            print(f"    {a:04x}  {ins}")
