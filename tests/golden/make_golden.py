"""Regenerates tests/golden/*.json|npz from the oracle (run from the repo root:
``python tests/golden/make_golden.py``).  The reference cannot be executed in this container
(no JVM; arithmetic lives in un-vendored Maven artefacts), so these are regression pins of the
restated oracle on seeded inputs, committed so the GPU box can check against them."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import fusion_oracle as fo  # noqa: E402
from oracle import pcm_oracle as po  # noqa: E402
from tests import synth  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def pcm_cases():
    cases = []
    for shape, shift, seed in [((48, 56, 64), (3, -2, 1), 101), ((64, 64, 64), (-6, 7, 2), 102),
                               ((33, 45, 71), (4, 4, -5), 103), ((80, 40, 40), (0, -11, 9), 104)]:
        a, b = synth.shifted_pair(shape, shift, seed=seed)
        r = po.pcm_shift(a, b)
        cases.append(dict(shape=list(shape), shift=list(shift), seed=seed, found=bool(r.found),
                          shift_int=[int(v) for v in r.shift_int], shift_sub=[float(v) for v in r.shift_sub],
                          r=float(r.r), n_overlap_px=int(r.n_overlap_px), pad=list(r.pad)))
    return dict(generator="tests/golden/make_golden.py", cases=cases)


def fusion_case():
    G = synth.field((40, 60, 150), seed=7, sigma=1.5)
    views = []
    for i, t in enumerate([(0.0, 0.0, 0.0), (41.3, 2.6, -1.2), (83.1, -1.7, 1.4)]):
        vol = synth.tile_from(G, (2, 5, int(t[0]) + 3), (32, 44, 56), 70 + i, noise=5.0)
        M = synth.translation(t)
        border, rng = fo.adjust_blending(M)
        views.append(fo.View(vol, M, border, rng))
    out = fo.fuse_block(views, (-2, -3, -1), (144, 52, 36), fo.AVG_BLEND)
    return out


if __name__ == "__main__":
    with open(os.path.join(HERE, "pcm_golden.json"), "w") as f:
        json.dump(pcm_cases(), f, indent=1)
    np.savez_compressed(os.path.join(HERE, "fusion_golden.npz"), avg_blend=fusion_case())
    print("golden vectors written")
