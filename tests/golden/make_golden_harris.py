"""Golden vectors for the Harris flavour of the scale-space detector (DET_HARRIS: Harris measure of the blurred gradient products,
pyramid.cpp:283-305; [HarrisAffine] of config_iter_mods_cviu.ini:28-44) from the reference's own sources compiled in place (oracle/_ref):
raw keys in FixedTh and NotLessThanRegions modes and the identity-view pipeline of a 200x150 synthetic image.
Run in the build container only:  python tests/golden/make_golden_harris.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))


def main():
    import synth
    from oracle.pyoracle import HessParams, Reference
    R = Reference()
    im = synth.blob_image(200, 150, seed=7)
    out = {"image": im.astype(np.uint8)}
    hp = HessParams.harris()
    out["raw_fixed_th"] = R.hessaff_detect(im, hp, raw=True)
    hp.mode = 4; hp.reg_number = 80
    out["raw_not_less_80"] = R.hessaff_detect(im, hp, raw=True)
    v = R.view_pipeline(im, hp=HessParams.harris())
    out["view_det"], out["view_desc"] = v[0], v[2].astype(np.uint8)
    np.savez_compressed(os.path.join(HERE, "harris_vectors.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
