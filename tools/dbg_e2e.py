"""where does the fusion e2e step spend its time? uploads only / fuse + D2H only / both"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bsgpu
from bsgpu import fusion as bf, synthetic
nat = bsgpu.native
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev)
ctx = bsgpu.Context(0, stream=stream.cuda_stream)
g, tile, stride, out_n = 4, 576, 491, 2048
nd = int(sys.argv[1]) if len(sys.argv) > 1 else 16
CH = int(sys.argv[2]) if len(sys.argv) > 2 else 64
tiles, models, tdims = synthetic.make_fusion_workload((g, g, g), tile, stride, dev, n_distinct=nd)
nv = len(tiles)
regs = {i: models[i] for i in range(nv)}
vdims = {i: tdims for i in range(nv)}
mine = list(range(nv))
blending = {i: bf.adjust_blending(models[i]) for i in mine}
grid = bf.grid_create((out_n,) * 3, (256, 256, 128), (128, 128, 128))
grid.sort(key=lambda b: (b[0][2], b[0][1], b[0][0]))
hosts = {}
for i in mine:
    k = tiles[i].data_ptr()
    if k not in hosts:
        hosts[k] = tiles[i].cpu().pin_memory()
order = sorted(mine, key=lambda v: models[v][2][3])
ring_n = 2 * CH
ring = torch.empty((ring_n, 256 * 256 * 128), dtype=torch.float32).pin_memory().numpy()
p = ctx.fuse_params("AVG_BLEND", 1, nat.DTYPE_F32)

def uploads():
    return {i: ctx.volume_upload_async(hosts[tiles[i].data_ptr()].numpy().view(np.uint16)) for i in order}

def fuse_all(hs, host_out=True):
    vd = {v: dict(src_to_world=models[v], vol_handle=hs[v], blend_border=blending[v][0], blend_range=blending[v][1]) for v in mine}
    slot = 0
    tt = []
    for c0 in range(0, len(grid), CH):
        t0 = time.perf_counter()
        chunk = grid[c0:c0 + CH]
        lo = tuple(min(b[0][d] for b in chunk) for d in range(3))
        hi = tuple(max(b[0][d] + b[1][d] - 1 for b in chunk) for d in range(3))
        vids = bf.find_overlapping_views(vdims, regs, lo, hi, mine)
        outs = []
        for (_, sz, _g) in chunk:
            outs.append(ring[slot % ring_n][:int(np.prod(sz))].reshape(sz[2], sz[1], sz[0]))
            slot += 1
        views = ctx.make_views(vd[v] for v in vids)
        t1 = time.perf_counter()
        ctx.fuse_blocks(views, [b[0] for b in chunk], [b[1] for b in chunk], p, outs=outs)
        tt.append((round(1e3 * (t1 - t0), 1), round(1e3 * (time.perf_counter() - t1), 1)))
    return tt

def T(fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize(); return 1e3 * (time.perf_counter() - t0), r

for rep in range(2):
    ms, hs = T(uploads)
    print(f"uploads only: {ms:.0f} ms ({len(order) * tile ** 3 * 2 / ms / 1e6:.1f} GB/s)")
    ms, tt = T(lambda: fuse_all(hs))
    print(f"fuse + D2H (resident tiles): {ms:.0f} ms ({out_n ** 3 * 4 / ms / 1e6:.1f} GB/s D2H); per call (host prep, call) ms: {tt[:4]} ...")
    for h in hs.values():
        ctx.volume_free(h)
    def both():
        hs = uploads()
        tt = fuse_all(hs)
        for h in hs.values():
            ctx.volume_free(h)
        return tt
    ms, tt = T(both)
    print(f"both: {ms:.0f} ms; per call: {tt}")
ctx.close()
