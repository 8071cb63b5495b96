"""Golden vectors for DSPSIFT (imagerepresentation.cpp:1547-1598: un-normalised plain-SIFT votes at numScales + 1 measurement-region sizes,
float sums, SIFTnorm on the float vector) from the reference's own sources compiled in place (oracle/_ref: DescribeRegions and
SIFTDescriptor are the reference's; the branch's glue is restated in oracle/ref_api.cpp: ref_describe_dsp).
Run in the build container only:  python tests/golden/make_golden_dsp.py"""
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
    im = synth.blob_image(240, 180, seed=17)
    keys = R.detect_orientation(im, R.hessaff_detect(im))
    out = {"image": im.astype(np.uint8), "keys": keys,
           "desc_default": R.describe_dsp(im, keys).astype(np.uint8),
           "desc_5_07_13_nophoto": R.describe_dsp(im, keys, numScales=5, startCoef=0.7, endCoef=1.3, photoNorm=False).astype(np.uint8)}
    np.savez_compressed(os.path.join(HERE, "dsp_vectors.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
