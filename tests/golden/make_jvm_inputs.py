"""Seeded inputs for tests/golden/java/GoldenDump.java (the JVM side that pins the oracle).

    python tests/golden/make_jvm_inputs.py            # writes tests/golden/jvm_inputs/{manifest.json, *.raw}
    (on a machine with Maven)  cd tests/golden/java && mvn -q compile exec:java -Dexec.args="../jvm_inputs ../jvm"
    python -m pytest tests/test_jvm_golden.py          # compares oracle/ (and, with -m gpu, the CUDA path) with ../jvm

Raw files are little-endian, x-fastest ([z][y][x]); the manifest lists dims as {x, y, z}."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from tests import synth  # noqa: E402

OUT = os.path.join(HERE, "jvm_inputs")


def main():
    os.makedirs(OUT, exist_ok=True)
    man = {"pcm": [], "fusion": [], "dog": []}
    cases = [("int_64", (64, 64, 64), (3, -2, 1), 1, False), ("odd_sizes", (48, 80, 96), (-7, 5, 11), 3, False),
             ("subpixel", (64, 72, 80), (4.3, -6.25, 1.4), 5, True), ("identical", (40, 40, 40), (0, 0, 0), 7, False)]
    for name, shape, shift, seed, sub in cases:
        a, b = (synth.subpixel_pair if sub else synth.shifted_pair)(shape, shift, seed=seed)
        if name == "identical":
            b = a.copy()
        a.astype("<u2").tofile(os.path.join(OUT, f"pcm_{name}_a.raw"))
        b.astype("<u2").tofile(os.path.join(OUT, f"pcm_{name}_b.raw"))
        man["pcm"].append({"name": name, "dims": list(shape[::-1]), "a": f"pcm_{name}_a.raw", "b": f"pcm_{name}_b.raw",
                           "planted_shift": list(shift), "peaksToCheck": 5, "doSubpixel": True, "minOverlap": 0.25})
    rng = np.random.default_rng(11)
    G = synth.field((72, 80, 200), seed=21, sigma=1.5)
    for name, rot in (("translation", 0.0), ("rotated", 0.5)):
        views = []
        for i in range(3):
            off = np.array([i * 50, 3 * i, 2 * i], dtype=np.float64)
            vol = synth.tile_from(G, (int(off[2]), int(off[1]), int(off[0])), (48, 56, 64), 30 + i, noise=5.0)
            M = synth.translation(off + rng.uniform(-2, 2, 3))
            if rot:
                R = synth.rot_z(rot, center_xyz=(32, 28, 0))
                M = (np.vstack([M, [0, 0, 0, 1]]) @ np.vstack([R, [0, 0, 0, 1]]))[:3]
            fn = f"fusion_{name}_v{i}.raw"
            vol.astype("<u2").tofile(os.path.join(OUT, fn))
            views.append({"setup": i, "dims": [64, 56, 48], "file": fn, "model": [float(v) for v in M.ravel()]})
        man["fusion"].append({"name": name, "views": views, "block_min": [-3, -2, -1], "block_size": [170, 64, 52],
                              "fusion_types": ["AVG", "AVG_BLEND", "MAX_INTENSITY", "LOWEST_VIEWID_WINS", "HIGHEST_VIEWID_WINS",
                                               "CLOSEST_PIXEL_WINS"]})
    # next row 8f-4: DoG interest points of one bead block (DoGImgLib2.computeDoG, J/SparkInterestPointDetection.java:550-566)
    from tests.test_dog_oracle import _beads
    img, _ = _beads()
    img.astype("<u2").tofile(os.path.join(OUT, "dog_beads.raw"))
    for name, mn, sz, fmin in (("whole", (0, 0, 0), img.shape[::-1], False), ("inner", (10, 8, 4), (30, 32, 28), True)):
        man["dog"].append({"name": name, "file": "dog_beads.raw", "dims": list(img.shape[::-1]), "interval_min": list(mn),
                           "interval_size": [int(v) for v in sz], "sigma": 1.8, "threshold": 0.004, "localization": 1,
                           "findMin": fmin, "findMax": True, "minIntensity": 0.0, "maxIntensity": 4000.0})
    json.dump(man, open(os.path.join(OUT, "manifest.json"), "w"), indent=1)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
