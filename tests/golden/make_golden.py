"""Regenerate tests/golden/*.npz from the CPU oracle (run from the repo root:
`python tests/golden/make_golden.py`). The reference ships no golden vectors for
a pixel cost (it has none), so these pin THIS repo's oracle against regressions;
the reference-derived known answers live in tests/test_oracle_vs_ref.py."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle  # noqa: E402
from unsynchronized_stereo_vision_proj325_b200 import _abi, synth  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = {
    # name: (W, H, C, shift, noise, params kwargs)
    "sad_gray_full": (96, 40, 1, 7, 2.0, dict(tmpl_w=16, tmpl_h=16, cost="sad")),
    "ssd_gray_d32": (96, 40, 1, 11, 2.0, dict(tmpl_w=8, tmpl_h=12, cost="ssd", search_min=2, search_max=32)),
    "zncc_bgr": (72, 36, 3, 5, 3.0, dict(tmpl_w=12, tmpl_h=10, cost="zncc", search_max=24, accept_threshold=0.5)),
    "ncc_gray_right": (80, 32, 1, -9, 1.0, dict(tmpl_w=16, tmpl_h=8, cost="ncc", camera_side=_abi.RIGHT_CAM,
                                                distance_kind=_abi.DIST_POWERLAW, search_max=40)),
    "sad_stride": (90, 50, 1, 13, 2.0, dict(tmpl_w=10, tmpl_h=6, cost="sad", stride_x=3, stride_y=4,
                                            distance_kind=_abi.DIST_POWERLAW, accept_threshold=0.05)),
}


def main():
    for name, (w, h, c, shift, noise, kw) in CASES.items():
        left, right = synth.make_pairs(2, w, h, c, shift=shift, noise_sigma=noise, seed=325)
        p = _abi.make_params(**kw)
        out = oracle.match_dense(left, right, p)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), left=np.ascontiguousarray(left), right=np.ascontiguousarray(right),
                            **{"out_" + k: v for k, v in out.items()})
        print(name, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
