python - <<'P'
import numpy as np, sys
sys.path.insert(0, '.')
from mods_b200 import synth
W, H = 4096, 3072
A = synth.blob_image(W, H, seed=1, n_blobs=int(1.5e-3 * W * H))
(A.astype(np.int32) & 255).astype(np.uint8).tofile('/tmp/img.u8')
P
tools/micro/mtree_tiles.bin /tmp/img.u8 4096 3072 | tee gpurun_out/s3_tiles_all.log
