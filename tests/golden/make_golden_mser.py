"""Golden MSER vectors produced by the REFERENCE's own sources compiled in place (oracle/_ref/libmods_ref.so ==
/root/reference/detectors/mser/* through oracle/build_ref.sh).  Run in the build container only:

    python tests/golden/make_golden_mser.py

Writes tests/golden/mser_vectors.npz:
  cat_regions / cat_keys      300x300 crop of build/examples/cat.png (crop stored in reference_vectors.npz: cat_gray)
  s_regions / s_keys          synthetic 320x240 blob image, seed 11, config_iter_mods_cviu.ini [MSER] parameters
  p_img / p_regions / p_keys  160x200 quantised plateau image (many equal-size merges, min_margin 1, min_size 8)
  n_img / n_regions           96x128 white-noise image, min_margin 2, min_size 5
region rows: polarity minI maxI threshold margin area border nruns cx cy sxx sxy syy   (getRLEExtrema, libExtrema.cpp:462)
key rows   : x y a11 a12 a21 a22 s response sub_type after DetectAffineRegions          (synth-detection.hpp:93)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))


def plateau_image(seed=5):
    rng = np.random.default_rng(seed)
    base = rng.random((40, 50))
    return (np.floor(np.kron(base, np.ones((4, 4))) * 12) * 20).astype(np.float32)


def noise_image(seed=6):
    return np.random.default_rng(seed).integers(0, 256, (96, 128)).astype(np.float32)


def main():
    import synth
    from oracle.pyoracle import Reference
    R = Reference()
    G = np.load(os.path.join(HERE, "reference_vectors.npz"))
    out = {}
    cat = G["cat_gray"]
    out.update(cat_regions=R.mser_regions(cat), cat_keys=R.mser_detect(cat))
    s = synth.blob_image(320, 240, seed=11)
    out.update(s_regions=R.mser_regions(s), s_keys=R.mser_detect(s))
    p = plateau_image()
    out.update(p_img=p.astype(np.uint8), p_regions=R.mser_regions(p, max_area=0.3, min_size=8, min_margin=1.0),
               p_keys=R.mser_detect(p, max_area=0.3, min_size=8, min_margin=1.0))
    n = noise_image()
    out.update(n_img=n.astype(np.uint8), n_regions=R.mser_regions(n, min_size=5, min_margin=2.0))
    np.savez_compressed(os.path.join(HERE, "mser_vectors.npz"), **out)
    print({k: getattr(v, "shape", v) for k, v in out.items()})


if __name__ == "__main__":
    main()
