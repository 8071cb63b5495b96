"""Golden vectors for HalfRootSIFT from the reference's own sources compiled in place (oracle/_ref): orientations modulo pi
(DetectOrientation with doHalfSIFT = 1, synth-detection.cpp:801-808, 841-919) and the 64-entry HalfRootSIFT descriptor
(SIFTDescriptor::operator(), siftdesc.cpp:401-442) of a 200x150 synthetic image.  Run in the build container only:
    python tests/golden/make_golden_half.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))


def main():
    import synth
    from oracle.pyoracle import Reference
    R = Reference()
    im = synth.blob_image(200, 150, seed=5)
    kps = R.hessaff_detect(im)
    out = {"image": im.astype(np.uint8), "kps": kps}
    for maxA in (1, 5):
        out["oriented_half_a%d" % maxA] = R.detect_orientation(im, kps, maxAngles=maxA, doHalfSIFT=1)
    out["desc_half"] = R.describe(im, out["oriented_half_a1"], rootsift=3).astype(np.uint8)
    v = R.view_pipeline(im, detector=0, desc=(5.1962, 41, True, 7))
    out["view_det"], out["view_desc"] = v[0], v[2].astype(np.uint8)
    np.savez_compressed(os.path.join(HERE, "half_sift_vectors.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
