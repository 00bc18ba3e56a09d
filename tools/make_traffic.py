"""profiles/traffic.json from `ncu --set full --page raw --csv` dumps: DRAM bytes (read + write) per launch of
every kernel of a bench step, summed per stage the way bench.py's stage_ms groups them.
    python tools/make_traffic.py c4_iphone raw1.csv [raw2.csv ...]
"""
import csv
import json
import os
import sys

STAGE = {"preprocess_fwd_kernel": "preprocess_fwd", "tile_scan_kernel": "bin", "tile_place_kernel": "bin",
         "tile_sort_kernel": "bin", "blend_fwd_kernel": "blend_fwd", "tile_split_kernel": "blend_fwd", "local_pearson_finalize_kernel": "loss", "ssim_fwd_kernel": "loss", "ssim_bwd_kernel": "loss",
         "loss_finalize_kernel": "loss", "blend_bwd_kernel": "blend_bwd", "preprocess_bwd_kernel": "preprocess_bwd",
         "dtable_kernel": "preprocess_bwd", "dtable2_kernel": "preprocess_bwd"}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main():
    cfg, paths = sys.argv[1], sys.argv[2:]
    per_kernel = {}
    per_inst = {}
    for p in paths:
        rows = list(csv.reader(open(p)))
        hdr, units = rows[0], rows[1]
        idx = {h: i for i, h in enumerate(hdr)}
        for r in rows[2:]:
            name = r[idx["Kernel Name"]]
            base = name.split("(")[0].split("<")[0].replace("void ", "").strip()
            if base not in STAGE:
                continue
            tot = 0.0
            for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                tot += float(r[idx[m]].replace(",", "")) * UNIT[units[idx[m]]]
            key = name.split("(")[0].replace("void ", "").strip()
            per_kernel.setdefault(key, []).append(tot)
            if "smsp__inst_executed.sum" in idx:
                per_inst.setdefault(key, []).append(float(r[idx["smsp__inst_executed.sum"]].replace(",", "")))
    kern = {k: sum(v) / len(v) for k, v in per_kernel.items()}          # mean over the captured launches
    out_path0 = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "profiles", "traffic.json")
    if os.path.exists(out_path0):                                        # keep kernels an earlier capture recorded
        old = json.load(open(out_path0)).get(cfg + "_kernels", {})
        fresh = {k.split("<")[0] for k in kern}
        if "dtable2_kernel" in fresh:
            fresh.add("dtable_kernel")                                   # replaced by dtable2
        for k, b in old.items():
            if k.split("<")[0] not in fresh:
                kern[k] = float(b)
    stage = {}
    for k, b in kern.items():
        s = STAGE[k.split("<")[0]]
        stage[s] = stage.get(s, 0.0) + b
    out_path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "profiles", "traffic.json")
    data = json.load(open(out_path)) if os.path.exists(out_path) else {}
    data[cfg] = {s: int(b) for s, b in stage.items()}
    data[cfg + "_kernels"] = {k: int(b) for k, b in kern.items()}
    inst = dict(data.get(cfg + "_warp_inst", {}))                        # warp instructions per launch (issue roofline)
    inst.update({k: int(sum(v) / len(v)) for k, v in per_inst.items()})
    data[cfg + "_warp_inst"] = inst
    data["_note"] = ("dram__bytes_read.sum + dram__bytes_write.sum per launch (ncu --set full --clock-control none), summed over "
                     "the kernels of each bench stage; written by tools/make_traffic.py")
    json.dump(data, open(out_path, "w"), indent=1, sort_keys=True)
    print(json.dumps(data[cfg], indent=1))


if __name__ == "__main__":
    main()
