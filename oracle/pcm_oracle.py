"""CPU oracle for hot path 1: phase correlation of one overlap-cropped tile pair.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` may be imported by the product
package; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline
legs use it, as the checker / the CPU arm.

PARITY UNPINNED.  The arithmetic restated here lives in Maven artefacts that are not
vendored under /root/reference (net.preibisch:BigStitcher:2.5.0 ->
net.imglib2.algorithm.phasecorrelation.{PhaseCorrelation2, PhaseCorrelation2Util,
PhaseCorrelationPeak2, FourNeighborhoodExtrema, BlendedExtendedMirroredRandomAccesible2},
pom.xml:107) and the reference has no golden vectors or assertion-bearing tests for this
path.  The restatement follows the reference's call site
(src/main/java/net/preibisch/bigstitcher/spark/SparkPairwiseStitching.java:194-303:
params.doSubpixel / params.peaksToCheck :200-202, computeStitching :247-255, "null ==
no shift found" :274-279) and the published upstream algorithm.
Every uncertain choice is a named constant below and is listed in PARITY_GAPS.md; the
pins are analytic known-answer tests (tests/test_pcm_oracle.py).

Array convention: numpy arrays are indexed [z, y, x] (C order), i.e. x is the fastest
axis exactly like imglib2's flat iteration order.  All public "dims"/"shift" triples
are in (x, y, z) order like the reference's long[] arrays.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np
import scipy.fft as sfft

# --- named constants for every recalled / uncertain upstream choice (PARITY_GAPS.md) ---
#: PairwiseStitching.getShift fills ``extension`` with 10 for every dimension (A.1 step 5).
DEFAULT_EXTENSION = 10
#: PhaseCorrelation2Util.normalizeInterval threshold: |c| < 1e-5 -> 0 (A.1 step 5, "(?)").
NORMALIZATION_THRESHOLD = 1e-5
#: PairwiseStitchingParameters default minOverlap, interpreted as a fraction of the crop's
#: voxel count (A.1 step 6, "(?)").
DEFAULT_MIN_OVERLAP = 0.25
#: peaksToCheck default of the CLI (SparkPairwiseStitching.java:79-80).
DEFAULT_PEAKS = 5
#: Local-maximum neighbourhood.  Upstream calls ``FourNeighborhoodExtrema.findMax`` on the
#: periodically extended PCM: a voxel is kept if NO axis neighbour (2n of them: +-1 along
#: each axis) is strictly larger.  (SURVEY A.1 recalls "3^n-1 neighbours, strict (?)"; the
#: class name pins the 2n-neighbourhood; listed in PARITY_GAPS.md.)
NEIGHBOURHOOD = "axis-2n-nonstrict"


def good_fft_size(n: int, even: bool = False) -> int:
    """Smallest 5-smooth length >= n (even if requested).

    Upstream pads to a mines-jtk "fast" length from an empirical cost table which is not
    reproducible here; the final shift is pad-size invariant (SURVEY.md section 7 hard
    part 1), so this build fixes its own policy: 2^a 3^b 5^c.
    """
    m = max(int(n), 2)
    while True:
        k = m
        for p in (2, 3, 5):
            while k % p == 0:
                k //= p
        if k == 1 and (not even or m % 2 == 0):
            return m
        m += 1


def extended_size(d: int, ext: int) -> int:
    """PhaseCorrelation2Util.getExtendedSize for equal-sized crops: d + 2*min(ext, d)."""
    return d + (2 * d if d < ext else 2 * ext)


def padded_dims(dims_xyz, extension=(DEFAULT_EXTENSION,) * 3):
    """Padded FFT size per axis (x must be even for the real-to-complex transform)."""
    return tuple(
        good_fft_size(extended_size(int(d), int(e)), even=(i == 0))
        for i, (d, e) in enumerate(zip(dims_xyz, extension))
    )


def _axis_profile(d: int, ext: int, P: int):
    """Source index and blending weight for every padded position along one axis.

    The crop is placed at offset ``e = min(ext, d)`` so the blended extension occupies
    [0, d + 2e) and zeros follow (both crops are placed identically, so the circular
    cross-power spectrum is the same as upstream's centred placement,
    FFTMethods.paddingIntervalCentered).  Outside the crop the value is the single-mirror
    extension (Views.extendMirrorSingle: -1 -> 1, d -> d-2) times a cosine fade
    0.5*(cos(pi*dist/e)+1), dist = distance in px to the nearest crop voxel
    (BlendedExtendedMirroredRandomAccesible2, A.1 step 5).
    """
    e = min(ext, d)
    idx = np.zeros(P, dtype=np.int64)
    w = np.zeros(P, dtype=np.float32)
    period = max(2 * d - 2, 1)
    for p in range(P):
        s = p - e  # coordinate in crop space
        if s < -e or s > d - 1 + e:
            continue
        if s < 0:
            dist = -s
        elif s > d - 1:
            dist = s - (d - 1)
        else:
            dist = 0
        if d == 1:
            m = 0
        else:
            m = s % period
            if m < 0:
                m += period
            if m >= d:
                m = period - m
        idx[p] = m
        w[p] = np.float32(0.5 * (math.cos(math.pi * dist / e) + 1.0)) if dist > 0 else np.float32(1.0)
    return idx, w


def blend_extend_pad(img: np.ndarray, extension, pdims_xyz) -> np.ndarray:
    """float32 padded volume [Pz, Py, Px]: blended mirrored extension, then zeros."""
    dz, dy, dx = img.shape
    ix, wx = _axis_profile(dx, extension[0], pdims_xyz[0])
    iy, wy = _axis_profile(dy, extension[1], pdims_xyz[1])
    iz, wz = _axis_profile(dz, extension[2], pdims_xyz[2])
    vol = img[np.ix_(iz, iy, ix)].astype(np.float32)
    # weight product order: ((1 * wx) * wy) * wz in float32 like the per-dimension loop
    wgt = (wx[None, None, :] * wy[None, :, None]) * wz[:, None, None]
    return vol * wgt


def calculate_pcm(img1: np.ndarray, img2: np.ndarray, extension=(DEFAULT_EXTENSION,) * 3,
                  workers: int = 1) -> np.ndarray:
    """PhaseCorrelation2.calculatePCM: float32 PCM of the padded size, [Pz, Py, Px].

    Forward transforms unnormalised, both spectra normalised to unit magnitude (elements
    with |c| < 1e-5 set to 0), second one conjugated, multiplied, inverse transform scaled
    by 1/N (FFT.complexToReal).  Single precision throughout (ComplexFloatType).
    """
    if img1.shape != img2.shape:
        raise ValueError("crops must have equal size (PairwiseStitching.getShift returns null otherwise)")
    dims_xyz = img1.shape[::-1]
    P = padded_dims(dims_xyz, extension)
    a = blend_extend_pad(img1, extension, P)
    b = blend_extend_pad(img2, extension, P)
    fa = sfft.rfftn(a, workers=workers)
    fb = sfft.rfftn(b, workers=workers)
    del a, b

    def _normalize(f):
        mag = np.abs(f)
        ok = mag >= np.float32(NORMALIZATION_THRESHOLD)
        out = np.zeros_like(f)
        np.divide(f, mag, out=out, where=ok)
        return out

    fa = _normalize(fa)
    fb = _normalize(fb)
    fa *= np.conj(fb)
    del fb
    pcm = sfft.irfftn(fa, s=(P[2], P[1], P[0]), workers=workers)
    return pcm.astype(np.float32, copy=False)


def find_peaks(pcm: np.ndarray, n_peaks: int):
    """FourNeighborhoodExtrema.findMax on the periodic PCM: the ``n_peaks`` largest voxels
    none of whose 6 axis neighbours is strictly larger.  Returns [(value, (x,y,z))], sorted
    by value descending; ties broken by ascending linear index (deterministic rule of this
    build; upstream's tie order depends on its thread split)."""
    ismax = np.ones(pcm.shape, dtype=bool)
    for ax in range(3):
        ismax &= pcm >= np.roll(pcm, 1, axis=ax)
        ismax &= pcm >= np.roll(pcm, -1, axis=ax)
    lin = np.flatnonzero(ismax)
    vals = pcm.ravel()[lin]
    if lin.size == 0:
        return []
    k = min(n_peaks, lin.size)
    # top-k by (value desc, index asc)
    order = np.lexsort((lin, -vals.astype(np.float64)))[:k]
    out = []
    for o in order:
        z, y, x = np.unravel_index(lin[o], pcm.shape)
        out.append((float(vals[o]), (int(x), int(y), int(z))))
    return out


def neighbourhood27(pcm: np.ndarray, loc_xyz) -> np.ndarray:
    """3x3x3 periodic neighbourhood, float32, indexed [dz+1, dy+1, dx+1]."""
    x, y, z = loc_xyz
    Pz, Py, Px = pcm.shape
    zi = [(z + d) % Pz for d in (-1, 0, 1)]
    yi = [(y + d) % Py for d in (-1, 0, 1)]
    xi = [(x + d) % Px for d in (-1, 0, 1)]
    return pcm[np.ix_(zi, yi, xi)].astype(np.float32)


def subpixel_offset(nb: np.ndarray):
    """imglib2 SubpixelLocalization.refinePeaks as driven by
    PhaseCorrelationPeak2.calculateSubpixelLocalization (no moves allowed, periodic
    extension): solve H*delta = -g with central-difference gradient/Hessian in double.
    ``nb`` is the 3x3x3 neighbourhood [z, y, x]; returns delta (x, y, z); zero offset if the
    Hessian is singular."""
    f = nb.astype(np.float64)
    c = f[1, 1, 1]
    g = np.array([
        (f[1, 1, 2] - f[1, 1, 0]) / 2.0,
        (f[1, 2, 1] - f[1, 0, 1]) / 2.0,
        (f[2, 1, 1] - f[0, 1, 1]) / 2.0,
    ])
    H = np.empty((3, 3))
    H[0, 0] = f[1, 1, 2] - 2 * c + f[1, 1, 0]
    H[1, 1] = f[1, 2, 1] - 2 * c + f[1, 0, 1]
    H[2, 2] = f[2, 1, 1] - 2 * c + f[0, 1, 1]
    H[0, 1] = H[1, 0] = (f[1, 2, 2] - f[1, 2, 0] - f[1, 0, 2] + f[1, 0, 0]) / 4.0
    H[0, 2] = H[2, 0] = (f[2, 1, 2] - f[2, 1, 0] - f[0, 1, 2] + f[0, 1, 0]) / 4.0
    H[1, 2] = H[2, 1] = (f[2, 2, 1] - f[2, 0, 1] - f[0, 2, 1] + f[0, 0, 1]) / 4.0
    return solve3(H, -g)


def solve3(H, rhs):
    """3x3 solve by Cramer's rule in double (same code path on the product's host side);
    returns zeros when det == 0 or non-finite."""
    a, b, c = H[0]
    d, e, f = H[1]
    g, h, i = H[2]
    det = a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g)
    if det == 0.0 or not math.isfinite(det):
        return (0.0, 0.0, 0.0)
    r0, r1, r2 = rhs
    dx = (r0 * (e * i - f * h) - b * (r1 * i - f * r2) + c * (r1 * h - e * r2)) / det
    dy = (a * (r1 * i - f * r2) - r0 * (d * i - f * g) + c * (d * r2 - r1 * g)) / det
    dz = (a * (e * r2 - r1 * h) - b * (d * r2 - r1 * g) + r0 * (d * h - e * g)) / det
    if not (math.isfinite(dx) and math.isfinite(dy) and math.isfinite(dz)):
        return (0.0, 0.0, 0.0)
    return (dx, dy, dz)


def expand_candidates(peak_xyz, pdims_xyz):
    """PhaseCorrelation2Util.expandPeakToPossibleShifts for equal-size crops (offset
    correction = 0): candidate i mirrors dimension d around the origin when bit d of i is
    0.  Returns the 8 integer shift triples in upstream's enumeration order."""
    out = []
    for i in range(8):
        s = list(peak_xyz)
        for d in range(3):
            if (i >> d) % 2 == 0:
                s[d] = s[d] + pdims_xyz[d] if s[d] < 0 else s[d] - pdims_xyz[d]
        out.append(tuple(s))
    return out


def overlap_intervals(dims_xyz, shift_xyz):
    """PhaseCorrelation2Util.getOverlapIntervals (equal-size images): returns
    (off1, off2, size) in xyz or None.  shift s means img1[p + s] <-> img2[p]."""
    off1, off2, size = [], [], []
    for d in range(3):
        s = shift_xyz[d]
        n = dims_xyz[d]
        if s >= 0:
            if s >= n:
                return None
            off1.append(s); off2.append(0); size.append(min(n - s, n))
        else:
            if s <= -n:
                return None
            off1.append(0); off2.append(-s); size.append(min(n + s, n))
    return tuple(off1), tuple(off2), tuple(size)


def pearson(a: np.ndarray, b: np.ndarray) -> float:
    """PhaseCorrelation2Util.getCorrelation: two-pass Pearson r in double; 0 when either
    variance sum is 0."""
    a = a.astype(np.float64).ravel()
    b = b.astype(np.float64).ravel()
    m1 = a.mean()
    m2 = b.mean()
    da = a - m1
    db = b - m2
    s11 = float(np.dot(da, da))
    s22 = float(np.dot(db, db))
    s12 = float(np.dot(da, db))
    if s11 == 0.0 or s22 == 0.0:
        return 0.0
    return s12 / math.sqrt(s11 * s22)


@dataclass
class PcmResult:
    found: bool
    shift_int: tuple = (0, 0, 0)
    shift_sub: tuple = (0.0, 0.0, 0.0)
    r: float = float("-inf")
    n_overlap_px: int = 0
    peak_index: tuple = (0, 0, 0)
    pcm_value: float = 0.0
    pad: tuple = (0, 0, 0)
    candidates: list = field(default_factory=list)  # (shift, r, npx) for diagnostics


def pcm_shift(img1: np.ndarray, img2: np.ndarray, peaks_to_check: int = DEFAULT_PEAKS,
              do_subpixel: bool = True, min_overlap_frac: float = DEFAULT_MIN_OVERLAP,
              extension=(DEFAULT_EXTENSION,) * 3, workers: int = 1) -> PcmResult:
    """calculatePCM + PhaseCorrelation2.getShift on two equal-size crops [z, y, x].

    Returns the best candidate by (r desc, overlap px desc, upstream enumeration order);
    ``found == False`` mirrors Java ``null`` (no candidate, or best r is -inf).
    """
    dims_xyz = tuple(int(v) for v in img1.shape[::-1])
    P = padded_dims(dims_xyz, extension)
    pcm = calculate_pcm(img1, img2, extension, workers=workers)
    n_px = dims_xyz[0] * dims_xyz[1] * dims_xyz[2]
    min_overlap_px = int(n_px * min_overlap_frac)  # (long) cast
    peaks = find_peaks(pcm, peaks_to_check)
    cands = []
    order = 0
    for val, loc in peaks:
        sub = subpixel_offset(neighbourhood27(pcm, loc)) if do_subpixel else (0.0, 0.0, 0.0)
        for s in expand_candidates(loc, P):
            ov = overlap_intervals(dims_xyz, s)
            r = float("-inf")
            npx = 0
            if ov is not None:
                o1, o2, sz = ov
                npx = sz[0] * sz[1] * sz[2]
                if npx < min_overlap_px:
                    npx = 0
                else:
                    a = img1[o1[2]:o1[2] + sz[2], o1[1]:o1[1] + sz[1], o1[0]:o1[0] + sz[0]]
                    b = img2[o2[2]:o2[2] + sz[2], o2[1]:o2[1] + sz[1], o2[0]:o2[0] + sz[0]]
                    r = pearson(a, b)
            cands.append((r, npx, order, s, sub, loc, val))
            order += 1
    if not cands:
        return PcmResult(False, pad=P)
    # Collections.sort(peaks, reverseOrder(byCrossCorr then nPixel)) is stable
    cands.sort(key=lambda c: (-c[0] if c[0] != float("-inf") else float("inf"), -c[1], c[2]))
    best = cands[0]
    diag = [(c[3], c[0], c[1]) for c in cands]
    if math.isinf(best[0]):
        return PcmResult(False, pad=P, candidates=diag)
    s = best[3]
    sub = best[4]
    return PcmResult(True, shift_int=s,
                     shift_sub=tuple(s[d] + sub[d] for d in range(3)) if do_subpixel else tuple(float(v) for v in s),
                     r=best[0], n_overlap_px=best[1], peak_index=best[5], pcm_value=best[6],
                     pad=P, candidates=diag)
