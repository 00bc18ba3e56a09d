"""Summarise an `ncu --page source --csv --print-source sass` export: per kernel, total warp
instructions, and the address ranges (loops) where they are spent.
    python tools/ncu_sass.py gpurun_out/prof_k_sass.csv [kernel-substring] [--dump lo hi]
"""
import csv, sys, io
path = sys.argv[1]
want = sys.argv[2] if len(sys.argv) > 2 and not sys.argv[2].startswith("--") else ""
dump = None
if "--dump" in sys.argv:
    i = sys.argv.index("--dump"); dump = (int(sys.argv[i + 1], 16), int(sys.argv[i + 2], 16))
text = open(path).read()
blocks = text.split('"Kernel Name",')
seen = set()
for blk in blocks[1:]:
    lines = blk.split("\n")
    name = lines[0].strip().strip('",')[:60]
    if want and want not in name: continue
    if name in seen: continue
    seen.add(name)
    rd = csv.DictReader(io.StringIO("\n".join(lines[1:])))
    rows = []
    for r in rd:
        try:
            rows.append((int(r["Address"], 16) if r["Address"].startswith("0x") else int(r["Address"]), r["Source"], int(r["Instructions Executed"]), int(r["# Samples"] or 0), r))
        except Exception:
            pass
    if not rows: continue
    base = rows[0][0]
    tot = sum(r[2] for r in rows); samp = sum(r[3] for r in rows)
    print(f"== {name}: {len(rows)} SASS instr, {tot/1e6:.1f} M warp-instr executed, {samp} samples")
    # segment by execution count level (log buckets)
    seg_start = 0
    def flush(a, b):
        cnt = sum(rows[k][2] for k in range(a, b)); sm = sum(rows[k][3] for k in range(a, b))
        if cnt / max(tot, 1) > 0.01:
            print(f"  [{rows[a][0]-base:05x}-{rows[b-1][0]-base:05x}] {b-a:4d} instr  exec/instr {rows[a][2]/1e6:8.2f} M  share {cnt/tot*100:5.1f}%  samples {sm/max(samp,1)*100:5.1f}%")
    for k in range(1, len(rows) + 1):
        if k == len(rows) or not (0.7 < (rows[k][2] + 1) / (rows[seg_start][2] + 1) < 1.43):
            flush(seg_start, k); seg_start = k
    if dump:
        stall_cols = [c for c in rows[0][4].keys() if c.startswith("stall_") and "Not Issued" not in c]
        for a, src, cnt, sm, r in rows:
            if dump[0] <= a - base <= dump[1]:
                top = sorted(((int(r[c] or 0), c[6:]) for c in stall_cols), reverse=True)[:2]
                tops = " ".join(f"{n}:{v}" for v, n in top if v)
                print(f"   {a-base:05x} {cnt/1e6:7.2f}M s={sm:5d} wf={r.get('L1 Wavefronts Shared','')}/{r.get('L1 Wavefronts Shared Ideal','')} {src[:70]:70s} {tops}")
