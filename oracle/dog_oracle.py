"""CPU restatement of the reference's Difference-of-Gaussian interest-point detection for one block.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg, never by the
product.  PARITY UNPINNED: the arithmetic lives in net.preibisch:multiview-reconstruction 8.0.0
(DoGImgLib2.computeDoG / computeSigmas, imglib2 Gauss3, LocalExtrema, quadratic localisation), which is not under
/root/reference; this follows the call site src/main/java/net/preibisch/bigstitcher/spark/
SparkInterestPointDetection.java:469-566 (parameters :476-503, block + 1 px halo :397-424) and the published algorithm
as recalled (every recalled constant is named below and listed in PARITY_GAPS.md).
"""
from __future__ import annotations

import numpy as np
from scipy.ndimage import correlate1d

STEPS_PER_OCTAVE = 4          # DoGImgLib2 / DifferenceOfGaussian.computeK(4): k = 2^(1/4)
IMAGE_SIGMA = 0.5             # DifferenceOfGUI.defaultImageSigma*
INITIAL_THRESHOLD_DIV = 3.0   # candidates at |DoG| >= threshold / 3, kept after localisation at |value| >= threshold


def compute_sigmas(sigma: float):
    """DoGImgLib2.computeSigmas: the two blur sigmas (relative to the image's own 0.5) and 1 / (k - 1)."""
    k = 2.0 ** (1.0 / STEPS_PER_OCTAVE)
    s1, s2 = sigma, sigma * k
    return np.sqrt(s1 * s1 - IMAGE_SIGMA ** 2), np.sqrt(s2 * s2 - IMAGE_SIGMA ** 2), 1.0 / (k - 1.0)


def gauss_kernel(sigma: float) -> np.ndarray:
    """Gauss3: truncated, normalised; half kernel SIZE max(2, int(3 sigma + 0.5) + 1) (radius = size - 1)."""
    size = max(2, int(3.0 * sigma + 0.5) + 1)
    r = size - 1
    x = np.arange(-r, r + 1, dtype=np.float64)
    k = np.exp(-0.5 * (x / sigma) ** 2)
    return (k / k.sum()).astype(np.float32)


def dog_volume(img: np.ndarray, sigma: float, min_intensity: float, max_intensity: float) -> np.ndarray:
    """(G_sa * I' - G_sb * I') / (k - 1) on the whole [z,y,x] image, I' = (I - min) / (max - min), float32,
    Views.extendMirrorDouble borders (scipy mode 'reflect': d c b a | a b c d | d c b a)."""
    sa, sb, kinv = compute_sigmas(sigma)
    f = ((img.astype(np.float32) - np.float32(min_intensity)) * np.float32(1.0 / (max_intensity - min_intensity))).astype(np.float32)
    a, b = f, f
    for axis in (2, 1, 0):          # x, y, z
        a = correlate1d(a, gauss_kernel(sa), axis=axis, mode="reflect", output=np.float32)
        b = correlate1d(b, gauss_kernel(sb), axis=axis, mode="reflect", output=np.float32)
    return ((a - b) * np.float32(kinv)).astype(np.float32)


def detect(img: np.ndarray, interval_min_xyz, interval_size_xyz, sigma=1.8, threshold=0.008, min_intensity=0.0,
           max_intensity=65535.0, find_max=True, find_min=False, localization=True):
    """Detections inside the block, sorted by (z, y, x): [(loc_xyz, value, voxel_xyz, is_max)].  The DoG is evaluated
    on the (virtually infinite, mirror-extended) image, so a block's result does not depend on the block grid."""
    dog = dog_volume(img, sigma, min_intensity, max_intensity)
    pad = np.pad(dog, 1, mode="symmetric")       # neighbours of border voxels come from the mirror extension ...
    # ... of the IMAGE, not of the DoG: recompute the 1-px rim exactly
    ext = np.pad(img, 1 + 64, mode="symmetric")
    dog_ext = dog_volume(ext, sigma, min_intensity, max_intensity)[64:-64, 64:-64, 64:-64]
    pad = dog_ext
    x0, y0, z0 = (int(v) for v in interval_min_xyz)
    nx, ny, nz = (int(v) for v in interval_size_xyz)
    thr0 = threshold / INITIAL_THRESHOLD_DIV if localization else threshold
    out = []
    c = pad[z0 + 1:z0 + 1 + nz, y0 + 1:y0 + 1 + ny, x0 + 1:x0 + 1 + nx]
    cand = np.zeros(c.shape, bool)
    if find_max:
        cand |= c >= np.float32(thr0)
    if find_min:
        cand |= -c >= np.float32(thr0)
    for (kz, ky, kx) in np.argwhere(cand):
        z, y, x = z0 + kz, y0 + ky, x0 + kx
        nb = pad[z:z + 3, y:y + 3, x:x + 3].astype(np.float64)
        v = float(np.float32(nb[1, 1, 1]))
        others = np.delete(nb.ravel(), 13)
        is_max = find_max and v >= thr0 and not np.any(others > v)
        is_min = find_min and -v >= thr0 and not np.any(others < v)
        if not (is_max or is_min):
            continue
        d = np.zeros(3)
        val = v
        if localization:
            g = np.array([0.5 * (nb[1, 1, 2] - nb[1, 1, 0]), 0.5 * (nb[1, 2, 1] - nb[1, 0, 1]), 0.5 * (nb[2, 1, 1] - nb[0, 1, 1])])
            H = np.empty((3, 3))
            H[0, 0] = nb[1, 1, 2] - 2 * v + nb[1, 1, 0]
            H[1, 1] = nb[1, 2, 1] - 2 * v + nb[1, 0, 1]
            H[2, 2] = nb[2, 1, 1] - 2 * v + nb[0, 1, 1]
            H[0, 1] = H[1, 0] = 0.25 * (nb[1, 2, 2] - nb[1, 2, 0] - nb[1, 0, 2] + nb[1, 0, 0])
            H[0, 2] = H[2, 0] = 0.25 * (nb[2, 1, 2] - nb[2, 1, 0] - nb[0, 1, 2] + nb[0, 1, 0])
            H[1, 2] = H[2, 1] = 0.25 * (nb[2, 2, 1] - nb[2, 0, 1] - nb[0, 2, 1] + nb[0, 0, 1])
            det = np.linalg.det(H)
            if abs(det) >= 1e-30 and np.isfinite(det):
                d = np.clip(-np.linalg.solve(H, g), -0.5, 0.5)
                val = v + 0.5 * float(g @ d)
            if abs(val) < threshold:
                continue
        elif abs(v) < threshold:
            continue
        out.append(((x + d[0], y + d[1], z + d[2]), val, (x, y, z), bool(is_max)))
    out.sort(key=lambda p: (p[2][2], p[2][1], p[2][0]))
    return out
