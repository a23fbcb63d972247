"""Known-answer tests of the DoG oracle (next row 8f-4)."""
import numpy as np

from oracle import dog_oracle as do


def _beads(shape=(40, 48, 56), centers=((20.3, 24.6, 17.2), (40.0, 10.0, 30.5), (8.7, 40.1, 8.0)), amp=3000.0, s=1.8, bg=200.0, seed=0):
    z, y, x = np.mgrid[0:shape[0], 0:shape[1], 0:shape[2]].astype(np.float64)
    img = np.full(shape, bg)
    for (cx, cy, cz) in centers:
        img += amp * np.exp(-((x - cx) ** 2 + (y - cy) ** 2 + (z - cz) ** 2) / (2 * s * s))
    img += np.random.default_rng(seed).normal(0, 5, shape)
    return np.clip(np.rint(img), 0, 65535).astype(np.uint16), centers


def test_sigmas_and_kernel():
    sa, sb, kinv = do.compute_sigmas(1.8)
    assert abs(sa - np.sqrt(1.8 ** 2 - 0.25)) < 1e-12 and abs(sb - np.sqrt((1.8 * 2 ** 0.25) ** 2 - 0.25)) < 1e-12
    assert abs(kinv - 1 / (2 ** 0.25 - 1)) < 1e-12
    k = do.gauss_kernel(sa)
    assert len(k) == 2 * (max(2, int(3 * sa + 0.5) + 1) - 1) + 1 and abs(k.sum() - 1) < 1e-6 and np.all(k == k[::-1])


def test_beads_are_found_with_subpixel_accuracy():
    img, centers = _beads()
    pts = do.detect(img, (0, 0, 0), img.shape[::-1], sigma=1.8, threshold=0.004, max_intensity=4000.0)
    assert len(pts) == len(centers)
    for c in centers:
        d = min(np.linalg.norm(np.subtract(p[0], c)) for p in pts)
        assert d < 0.15
    assert all(p[3] for p in pts) and all(p[1] > 0.004 for p in pts)
    # dark blobs are minima
    inv = (4000 - img.astype(np.int64)).clip(0).astype(np.uint16)
    assert len(do.detect(inv, (0, 0, 0), img.shape[::-1], threshold=0.004, max_intensity=4000.0, find_max=False, find_min=True)) == len(centers)


def test_block_grid_invariance():
    """A detection belongs to the block that contains its voxel: the union over a block grid == one whole-image call."""
    img, _ = _beads(seed=3)
    dims = img.shape[::-1]
    whole = do.detect(img, (0, 0, 0), dims, threshold=0.004, max_intensity=4000.0)
    parts = []
    for x0 in (0, 28):
        for y0 in (0, 24):
            parts += do.detect(img, (x0, y0, 0), (28, 24, dims[2]), threshold=0.004, max_intensity=4000.0)
    parts.sort(key=lambda p: (p[2][2], p[2][1], p[2][0]))
    assert [p[2] for p in parts] == [p[2] for p in whole]
    assert np.allclose([p[0] for p in parts], [p[0] for p in whole])
