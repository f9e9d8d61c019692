"""TEST INFRASTRUCTURE.  Copies the reference's bundled test triple ``LHBDC/frames/{ref_1,current,ref_2}.png``
(1920x1080 RGB; the inputs of ``python encode_B.py`` / ``decode_B.py``, LHBDC/README.md, BASELINE.json configs[0]) into
``tests/golden/lhbdc_frames_1080p.npz`` as uint8 ``[3, 3, 1080, 1920]`` (order: ref_1 = x_before, current, ref_2 =
x_after) so that bench.py / the GPU tests can run config 1 on the GPU box, where /root/reference does not exist.

    python oracle/make_frames_fixture.py        # run in the build container
"""
import os
import sys

import numpy as np
from PIL import Image

REF = "/root/reference/LHBDC/frames"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "lhbdc_frames_1080p.npz")


def main():
    frames = []
    for name in ("ref_1.png", "current.png", "ref_2.png"):
        im = np.asarray(Image.open(os.path.join(REF, name)).convert("RGB"))
        assert im.shape == (1080, 1920, 3) and im.dtype == np.uint8, (name, im.shape, im.dtype)
        frames.append(im.transpose(2, 0, 1))
    arr = np.stack(frames, 0)
    np.savez_compressed(OUT, frames_u8=arr, names=np.array(["ref_1", "current", "ref_2"]))
    print(f"wrote {OUT}: {arr.shape} {os.path.getsize(OUT) / 1e6:.1f} MB")


if __name__ == "__main__":
    sys.exit(main())
