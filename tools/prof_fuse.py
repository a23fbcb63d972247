"""Fusion perf driver (resident inputs/outputs): config-3-shaped slab fused with bs_fuse_blocks.
usage: prof_fuse.py [grid=4] [zblocks=2] [blocks_per_call=64] [rot=0] [reps=3]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import bsgpu  # noqa: E402
from bsgpu import fusion as bf, synthetic  # noqa: E402

g = int(sys.argv[1]) if len(sys.argv) > 1 else 4
zb = int(sys.argv[2]) if len(sys.argv) > 2 else 2
per_call = int(sys.argv[3]) if len(sys.argv) > 3 else 64
rot = float(sys.argv[4]) if len(sys.argv) > 4 else 0.0
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
mode = sys.argv[6] if len(sys.argv) > 6 else "plain"      # plain | oblique (rot about (1,1,1)) | content
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev)
ctx = bsgpu.Context(0, stream=stream.cuda_stream)
out_n = 491 * (g - 1) + 576
out_n = (out_n // 256) * 256
tiles, models, tdims = synthetic.make_fusion_workload((g, g, g), 576, 491, dev, n_distinct=min(8, g ** 3), rot_deg=rot)
torch.cuda.synchronize()
nv = len(tiles)
if mode == "oblique":
    ax = np.array([1.0, 1.0, 1.0]) / np.sqrt(3.0)
    th = np.deg2rad(rot if rot else 0.5)
    K = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
    Rg = np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * (K @ K)
    c = np.array([288.0, 288.0, 288.0])
    models = [np.hstack([Rg, (np.asarray(m)[:, 3] + c - Rg @ c)[:, None]]) for m in synthetic.make_fusion_workload((g, g, g), 576, 491, dev, n_distinct=1)[1]]
regs = {i: models[i] for i in range(nv)}
vdims = {i: tdims for i in range(nv)}
handles = {i: ctx.volume_wrap(tiles[i], tdims, bsgpu.native.DTYPE_U16) for i in range(nv)}
blending = {i: bf.adjust_blending(models[i]) for i in range(nv)}
z0 = 384
grid = [b for b in bf.grid_create((out_n, out_n, out_n), (256, 256, 128), (128, 128, 128)) if z0 <= b[0][2] < z0 + 128 * zb]
nvox = sum(int(np.prod(b[1])) for b in grid)
out = torch.empty(nvox, dtype=torch.float32, device=dev)
params = ctx.fuse_params("AVG_BLEND_CONTENT" if mode == "content" else "AVG_BLEND")
chandles = {}
if mode == "content":
    for v in range(nv):
        key = tiles[v].data_ptr()
        if key not in chandles:
            chandles[key] = ctx.content_weights(handles[v], 20.0, 40.0)
allviews = ctx.make_views(dict(src_to_world=models[v], vol_handle=handles[v], blend_border=blending[v][0],
                               blend_range=blending[v][1], content_handle=chandles.get(tiles[v].data_ptr(), 0)) for v in range(nv))
calls = []
off = 0
for c0 in range(0, len(grid), per_call):
    chunk = grid[c0:c0 + per_call]
    ptrs = []
    for (o, s, _) in chunk:
        ptrs.append(out.data_ptr() + 4 * off)
        off += int(np.prod(s))
    calls.append(([b[0] for b in chunk], [b[1] for b in chunk], ptrs))


def step():
    for mins, sizes, ptrs in calls:
        ctx.fuse_blocks(allviews, mins, sizes, params, outs=ptrs)


step()
ctx.synchronize()
e0 = torch.cuda.Event(enable_timing=True)
e1 = torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record(stream)
for _ in range(reps):
    step()
e1.record(stream)
e1.synchronize()
ms = e0.elapsed_time(e1) / reps
print(f"grid {g}^3 out {out_n} slab {zb} x128: {len(grid)} blocks, {nvox/1e6:.0f} Mvox, {ms:.3f} ms/step, "
      f"{nvox/ms/1e6:.1f} Gvox/s, wall {1000*(time.perf_counter()-t0)/reps:.2f} ms, rot {rot}, mode {mode}")
ctx.profile_reset(); ctx.profile_enable(True); step(); ctx.profile_enable(False)
for tag in ("fuse_plan", "fuse"):
    tms, cnt = ctx.profile_get(tag)
    print(tag, tms, cnt)
print("checksum", float(out[::97].double().sum()))
ctx.close()
