#!/usr/bin/env python
"""Development helper: A/B builds of librodygs_b200.so that differ only in -D flags of chosen translation units.
    python tools/build_variant.py <name> [blend.cu:-DCH=48,-DPOOL=136,-DBWD_MINB=10] ...
writes rodygs_b200/_build/variants/lib_<name>.so; select it with RDG_LIB_PATH=<that path> (rodygs_b200/_lib.py)."""
import os, subprocess, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from rodygs_b200 import build as B

def main():
    name, specs = sys.argv[1], sys.argv[2:]
    B.build()
    vdir = os.path.join(B.OBJ_DIR, "variants")
    os.makedirs(vdir, exist_ok=True)
    over = {}
    for sp in specs:
        src, flags = sp.split(":", 1)
        over[src] = flags.split(",")
    objs = []
    for src, extra in B.SOURCES.items():
        obj = os.path.join(B.OBJ_DIR, src.replace(".cu", ".o"))
        if src in over:
            obj = os.path.join(vdir, f"{name}_{src.replace('.cu', '.o')}")
            cmd = [B._nvcc(), *B.ARCH, *B.COMMON, *extra, *over[src], "-Xptxas", "-v", "-c", os.path.join(B.CSRC, src), "-o", obj]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                print(r.stdout, r.stderr); raise SystemExit(1)
            for l in (r.stdout + r.stderr).splitlines():
                if "Used" in l or "spill" in l or "Compiling entry" in l:
                    print("   ", l.strip()[:150])
        objs.append(obj)
    out = os.path.join(vdir, f"lib_{name}.so")
    subprocess.run([B._nvcc(), *B.ARCH, "-shared", "-o", out, *objs, "-lcudart"], check=True)
    print(out)

if __name__ == "__main__":
    main()
