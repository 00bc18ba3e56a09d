"""CPU analysis of the blend workload at a bench config (design aid, not product code):
how many (warp, Gaussian) and (pixel, Gaussian) pairs the blend kernels visit versus how many
actually contribute.  Uses the oracle's preprocess + binning on the synthetic scene.

    python tools/blend_stats.py [config] [n_tiles_sampled]
"""
import math
import sys
import os

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import splat_oracle as so          # noqa: E402
from rodygs_b200 import synthetic              # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "c4_iphone"
n_sample = int(sys.argv[2]) if len(sys.argv) > 2 else 120
N, H, W, T, _ = synthetic.CONFIGS[cfg]
torch.manual_seed(0)
scene = synthetic.make_scene(N, H, W, T, seed=0)
cam = synthetic.make_camera(0, 8, H, W, T)
xyz = torch.cat([scene["static"]["xyz"], scene["dynamic"]["xyz"]])
scal = torch.exp(torch.cat([scene["static"]["scaling"], scene["dynamic"]["scaling"]]))
rot = torch.nn.functional.normalize(torch.cat([scene["static"]["rotation"], scene["dynamic"]["rotation"]]))
opa = torch.sigmoid(torch.cat([scene["static"]["opacity"], scene["dynamic"]["opacity"]]))
col = torch.rand(N, 3)
st = so.Settings(image_height=H, image_width=W, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy, bg=torch.zeros(3),
                 scale_modifier=1.0, projmatrix=cam.projection_matrix.t().contiguous(), sh_degree=0,
                 prefiltered=False, debug=False, enable_cov_grad=True, enable_sh_grad=True)
with torch.no_grad():
    pp = so.preprocess(xyz, scal, rot, opa, None, col, cam.world_view_transform.t().contiguous(), st)
    bn = so.bin_tiles(pp, H, W)
D = bn.keys.numel()
print(f"{cfg}: N={N} visible={int(pp.visible.sum())} D={D} D/tile={D / bn.ranges.shape[0]:.0f}")
gx = (W + 15) // 16
rng = np.random.default_rng(0)
nonempty = np.nonzero((bn.ranges[:, 1] > bn.ranges[:, 0]).numpy())[0]
tiles = rng.choice(nonempty, size=min(n_sample, len(nonempty)), replace=False)
xy = pp.xy.numpy(); con = pp.conic.numpy(); op = pp.opacity.numpy()
tot = dict(slots=0, slots_live=0, warp_pairs_bbox=0, warp_pairs_exact=0, lane_pairs_warp=0, hits=0, hits_live=0,
           sub44_pairs=0, steps32=0, steps64=0, steps128=0, steps256=0, chunks32=0, batches=0, warp_iters_live=0,
           warp_iters_any=0, gsteps256=0)
yy, xx = np.meshgrid(np.arange(16), np.arange(16), indexing="ij")
for t in tiles:
    lo, hi = bn.ranges[t].tolist()
    loc = bn.local[lo:hi].numpy()
    n = len(loc)
    tx, ty = t % gx, t // gx
    px = (tx * 16 + xx).astype(np.float32); py = (ty * 16 + yy).astype(np.float32)
    dx = xy[loc, 0][:, None, None] - px[None]; dy = xy[loc, 1][:, None, None] - py[None]
    A = con[loc, 0][:, None, None]; B = con[loc, 1][:, None, None]; Cc = con[loc, 2][:, None, None]
    power = -0.5 * (A * dx * dx + Cc * dy * dy) - B * dx * dy
    alpha = np.minimum(0.99, op[loc][:, None, None] * np.exp(np.minimum(power, 0)))
    hit = (power <= 0) & (alpha >= 1 / 255.0)                      # [n,16,16]
    # transmittance / early termination
    a_eff = np.where(hit, alpha, 0.0)
    Tcum = np.cumprod(1 - a_eff, axis=0)
    stopped = Tcum < 1e-4                                            # from the first Gaussian that would push T below
    live = ~stopped                                                  # Gaussian i is blended iff T after it >= 1e-4
    hit_live = hit & live
    # slots the tile walks: until every pixel is done
    done_all = stopped.all(axis=(1, 2))
    n_walk = int(np.argmax(done_all)) if done_all.any() else n
    n_walk = int(math.ceil(max(n_walk, 1) / 256) * 256) if n_walk < n else n
    tot["slots"] += n; tot["slots_live"] += min(n_walk, n)
    tot["hits"] += int(hit.sum()); tot["hits_live"] += int(hit_live.sum())
    # warp sub-tiles 8x4: [n, 4 rows of 4, 2 cols of 8]
    hw = hit.reshape(n, 4, 4, 2, 8).any(axis=(2, 4))                 # exact: any pixel of the sub-tile hit
    tot["warp_pairs_exact"] += int(hw.sum())
    hlw = hit_live.reshape(n, 4, 4, 2, 8)
    tot["warp_iters_any"] += int(hlw.any(axis=(2, 4)).sum())
    h44 = hit.reshape(n, 4, 4, 4, 4).any(axis=(2, 4))
    tot["sub44_pairs"] += int(h44.sum())
    # bbox cull as in rdg_warp_mask
    o = op[loc]
    tau = 2 * np.log(255 * np.maximum(o, 1e-9)) * 1.001 + 1e-3
    det = con[loc, 0] * con[loc, 2] - con[loc, 1] ** 2
    ex = np.sqrt(np.maximum(tau * con[loc, 2] / det, 0)) * 1.001 + 0.01
    ey = np.sqrt(np.maximum(tau * con[loc, 0] / det, 0)) * 1.001 + 0.01
    x0 = xy[loc, 0] - ex - tx * 16; x1 = xy[loc, 0] + ex - tx * 16
    y0 = xy[loc, 1] - ey - ty * 16; y1 = xy[loc, 1] + ey - ty * 16
    okx = np.stack([(x1 >= 0) & (x0 <= 7), (x1 >= 8) & (x0 <= 15)], 1)            # [n,2]
    oky = np.stack([(y1 >= 4 * s) & (y0 <= 4 * s + 3) for s in range(4)], 1)     # [n,4]
    wm = (oky[:, :, None] & okx[:, None, :]) & (o >= 1 / 255.0)[:, None, None]
    tot["warp_pairs_bbox"] += int(wm.sum())
    # per-lane step counts for the "each lane walks its own hit list" scheme (per warp, chunked)
    hl = hit_live.reshape(n, 4, 4, 2, 8).transpose(0, 1, 3, 2, 4).reshape(n, 8, 32)   # [n, warp, lane]
    for ck, key in ((32, "steps32"), (64, "steps64"), (128, "steps128"), (256, "steps256")):
        nb = math.ceil(n / ck)
        pad = np.zeros((nb * ck - n, 8, 32), bool)
        h = np.concatenate([hl, pad]).reshape(nb, ck, 8, 32)
        per_lane = h.sum(axis=1)                   # [nb, 8, 32]
        tot[key] += int(per_lane.max(axis=2).sum())
    # Gaussian-parallel steps: per warp per 256-batch, max over Gaussians of pixels hit in the sub-tile
    nb = math.ceil(n / 32)
    pad = np.zeros((nb * 32 - n, 8, 32), bool)
    h = np.concatenate([hl, pad]).reshape(nb, 32, 8, 32)
    per_g = h.sum(axis=3)                          # [nb, 32 gaussians, 8 warps]
    tot["gsteps256"] += int(per_g.max(axis=1).sum())
    tot["chunks32"] += nb * 8
    tot["batches"] += math.ceil(n / 256)

s = tot["slots"]
print(f"sampled tiles {len(tiles)}  slots {s}  (walked until all pixels done: {tot['slots_live'] / s:.3f})")
print(f"pixel pairs visited if no cull  : {s * 256}")
print(f"hits (alpha>=1/255)             : {tot['hits']}  = {tot['hits'] / (s * 256):.4f} of all pixel pairs; per slot {tot['hits'] / s:.1f} px")
print(f"hits that are blended (T live)  : {tot['hits_live']}  = {tot['hits_live'] / tot['hits']:.3f} of hits")
print(f"(warp,slot) pairs bbox mask     : {tot['warp_pairs_bbox'] / (8 * s):.3f}   exact: {tot['warp_pairs_exact'] / (8 * s):.3f}   with >=1 live hit: {tot['warp_iters_any'] / (8 * s):.3f}")
print(f"4x4 sub-tile pairs exact        : {tot['sub44_pairs'] / (16 * s):.3f}")
print(f"lane efficiency today (live hits / (bbox warp pairs*32)): {tot['hits_live'] / (tot['warp_pairs_bbox'] * 32):.3f}")
for key, ck in (("steps32", 32), ("steps64", 64), ("steps128", 128), ("steps256", 256)):
    print(f"own-list steps, chunk {ck:3d}: {tot[key]} per-warp steps = {tot[key] / (8 * s):.3f} per (warp,slot); lane eff {tot['hits_live'] / (tot[key] * 32):.3f}")
print(f"gaussian-parallel steps (32 gaussians/lane-group): {tot['gsteps256']} = {tot['gsteps256'] / (8 * s):.3f} per (warp,slot)")

# ---- sub-warp schemes: a warp walks k sub-lists in lock step (iterations = max over its sub-lists, per 256-batch)
def subwarp_iters(hit_arr, n, sh, sw, per_warp):
    """hit_arr [n,16,16] bool; sub-tiles of sh x sw pixels; `per_warp` sub-tiles share a warp (adjacent in x first)."""
    ny, nx = 16 // sh, 16 // sw
    sub = hit_arr.reshape(n, ny, sh, nx, sw).any(axis=(2, 4)).reshape(n, ny * nx)     # [n, subtiles] row-major
    nb = math.ceil(n / 256)
    pad = np.zeros((nb * 256 - n, ny * nx), bool)
    cnt = np.concatenate([sub, pad]).reshape(nb, 256, ny * nx).sum(axis=1)             # [nb, subtiles]
    grp = cnt.reshape(nb, (ny * nx) // per_warp, per_warp)
    return int(grp.max(axis=2).sum()), int(cnt.sum())

res = {}
for t in tiles:
    lo, hi = bn.ranges[t].tolist()
    loc = bn.local[lo:hi].numpy(); n = len(loc)
    tx, ty = t % gx, t // gx
    px = (tx * 16 + xx).astype(np.float32); py = (ty * 16 + yy).astype(np.float32)
    dx = xy[loc, 0][:, None, None] - px[None]; dy = xy[loc, 1][:, None, None] - py[None]
    A = con[loc, 0][:, None, None]; B = con[loc, 1][:, None, None]; Cc = con[loc, 2][:, None, None]
    power = -0.5 * (A * dx * dx + Cc * dy * dy) - B * dx * dy
    alpha = np.minimum(0.99, op[loc][:, None, None] * np.exp(np.minimum(power, 0)))
    hit = (power <= 0) & (alpha >= 1 / 255.0)
    for name, (sh, sw, pw) in {"8x4 full warp (sh=4,sw=8)": (4, 8, 1), "4x4 half warps": (4, 4, 2), "2x8 half warps": (2, 8, 2),
                               "4x2 quarter (sh=2,sw=4)": (2, 4, 4), "2x4 quarter (sh=4,sw=2)": (4, 2, 4),
                               "2x2 eighth": (2, 2, 8)}.items():
        it, pairs = subwarp_iters(hit, n, sh, sw, pw)
        r = res.setdefault(name, [0, 0, sh * sw * pw, 256 // (sh * sw * pw)])
        r[0] += it; r[1] += pairs
print("\nscheme: warp-iterations per slot (all warps of the tile), lane efficiency")
for name, (it, pairs, lanes, warps) in res.items():
    print(f"  {name:28s} warps/tile {warps:2d}  iters/slot {it / s:.3f}  (no-imbalance {pairs / s / (lanes // (16 * 16 // (warps * 0 + 1)) if False else 1):.0f} sub-pairs)  lane eff {tot['hits'] / (it * 32):.3f}")
