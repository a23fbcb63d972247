"""Small fixed workload for ncu captures: 2 phase-correlation pairs of 512^3 and one fusion
step over a 2x2x2 grid of 576^3 tiles (output 1024^3 in 256x256x128 super-blocks)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import bsgpu  # noqa: E402
from bsgpu import fusion as bf, synthetic  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "all"
dev = torch.device("cuda", 0)
ctx = bsgpu.Context(0)
if what in ("all", "pcm"):
    n = int(os.environ.get("PROF_N", "512"))
    a, b, s = synthetic.make_pcm_workload(2, n=n, device=dev)
    torch.cuda.synchronize()
    for _ in range(int(os.environ.get("PROF_REPS", "2"))):
        r = ctx.pcm_batch(a, b, None, [(n, n, n)] * 2, bsgpu.native.DTYPE_U16)
    print([x.shift_int for x in r], s)
if what in ("all", "fuse"):
    tiles, models, tdims = synthetic.make_fusion_workload((2, 2, 2), 576, 491, dev, n_distinct=8)
    torch.cuda.synchronize()
    regs = {i: models[i] for i in range(8)}
    vdims = {i: tdims for i in range(8)}
    handles = {i: ctx.volume_wrap(tiles[i], tdims, bsgpu.native.DTYPE_U16) for i in range(8)}
    out = torch.empty(256 * 256 * 128, dtype=torch.float32, device=dev)
    params = ctx.fuse_params("AVG_BLEND")
    grid = bf.grid_create((1024, 1024, 1024), (256, 256, 128), (128, 128, 128))
    for (o, sz, _) in grid[120:136]:
        vids = bf.find_overlapping_views(vdims, regs, o, tuple(o[d] + sz[d] - 1 for d in range(3)))
        views = [dict(src_to_world=models[v], vol_handle=handles[v], blend_border=bf.adjust_blending(models[v])[0],
                      blend_range=bf.adjust_blending(models[v])[1]) for v in vids]
        ctx.fuse_block(views, o, sz, params, out=out)
    ctx.synchronize()
    print("fused", float(out.sum()))
ctx.close()
