"""debug helper (test infrastructure, not collected by pytest): where do GPU fusion and the C oracle disagree on the
config-3 super-block?  python tests/dbg_fuse_vs_oracle.py"""
import sys
sys.path.insert(0, ".")
import numpy as np
import bsgpu
from oracle import c_fusion, fusion_oracle as fo
import tests.test_fusion_gpu as T

rot = float(sys.argv[1]) if len(sys.argv) > 1 else 0.0
tile, stride = 576, 491
bmin, bsize = (384, 384, 448), (256, 256, 128)
rng = np.random.default_rng(7)
a = np.deg2rad(rot)
R = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]])
views = []
for k in range(2):
    for j in range(2):
        for i in range(2):
            t = stride * np.array([i, j, k], dtype=np.float64) + rng.uniform(-2, 2, 3)
            c = np.array([tile / 2, tile / 2, 0.0])
            M = np.hstack([R, (t + c - R @ c)[:, None]])
            lo = np.array(bmin) - t - 12
            hi = np.array(bmin) + np.array(bsize) - t + 12
            region = [(int(lo[d]), int(hi[d])) for d in (2, 1, 0)]
            views.append((T._sparse_tile((tile,) * 3, region, 100 + 4 * k + 2 * j + i), M))
with bsgpu.Context(0) as ctx:
    got, want = T._run_c(ctx, views, bmin, bsize)
err = np.abs(got - want) / np.maximum(np.abs(want), 1.0)
bad = np.argwhere(err > 1e-4)
print("rot", rot, "n bad", len(bad), "max", err.max(), "median err", np.median(err))
ov = T._run.last[0]
for idx in bad[:25]:
    w = np.array([bmin[0] + idx[2], bmin[1] + idx[1], bmin[2] + idx[0]], dtype=np.float64)
    line = [tuple(int(v) for v in idx), float(got[tuple(idx)]), float(want[tuple(idx)])]
    for vi, v in enumerate(ov):
        inv = fo.invert_affine(v.src_to_world)
        s = inv[:, :3] @ w + inv[:, 3]
        dims = np.array(v.img.shape[::-1])
        if np.all(s >= -1) and np.all(s <= dims):
            line.append((vi, [round(float(x), 5) for x in s]))
    print(line)
hist = np.histogram(np.log10(np.maximum(err, 1e-12)), bins=[-12, -7, -6, -5, -4.5, -4, -3.5, -3, -2, 0])
print(hist)
