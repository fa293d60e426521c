"""CPU: the C-ABI library loads, exports every symbol include/usv_b200.h declares,
its pure-host entry points agree with the oracle, and it refuses to run without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from unsynchronized_stereo_vision_proj325_b200 import _abi, api, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "usv_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(usv_[a-z0-9_]+)\s*\(", src)))


def test_exports_every_declared_symbol():
    L = api.lib()
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), n
    assert sorted(api.EXPORTS) == names
    assert L.usv_abi_version() == _abi.USV_ABI_VERSION


def test_struct_layouts_match_header():
    assert C.sizeof(_abi.Match) == 16 and _abi.Match.MatchValue.offset == 8
    assert C.sizeof(_abi.SearchParams) == 48 and _abi.SearchParams.accept_threshold.offset == 40
    assert C.sizeof(_abi.FrameDesc) == 24 and _abi.FrameDesc.frame_stride.offset == 16
    assert C.sizeof(_abi.Outputs) == 72 and _abi.Outputs.raw_cost_u16.offset == 56 and _abi.Outputs.resolved_disparity_u16.offset == 64


def test_no_cpu_fallback():
    """Without a CUDA device the product must refuse to run (no silent CPU path)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(api.UsvError):
        api.Context(0)


@pytest.mark.parametrize("kw", [
    dict(tmpl_w=16, tmpl_h=16), dict(tmpl_w=16, tmpl_h=16, search_max=128), dict(tmpl_w=8, tmpl_h=4, stride_x=3, stride_y=2, search_min=2, search_max=40),
    dict(tmpl_w=32, tmpl_h=32, camera_side=_abi.RIGHT_CAM, search_max=256),
])
def test_grid_dims_vs_oracle(oracle, kw):
    for (w, h) in ((640, 480), (1280, 720), (97, 61)):
        f = _abi.FrameDesc(w, h, 1, w, w * h)
        p = _abi.make_params(**kw)
        assert api.grid_dims(f, p) == oracle.grid_dims(f, p)


def test_baseline_config_counts():
    """C2 = 290 625 windows and 90 965 625 candidate evals per 640x480 pair."""
    f = _abi.FrameDesc(640, 480, 1, 640, 640 * 480)
    nx, ny, ev = api.grid_dims(f, _abi.make_params(tmpl_w=16, tmpl_h=16))
    assert (nx * ny, ev) == (290625, 90965625)
    _, _, ev = api.grid_dims(f, _abi.make_params(tmpl_w=16, tmpl_h=16, search_max=127))
    assert ev == 33420480 + 0  # D=128 variant (d in [0,127])


def test_pair_nearest_vs_oracle(oracle):
    rng = np.random.default_rng(325)
    for trial in range(60):
        nl, nr = int(rng.integers(0, 60)), int(rng.integers(0, 60))
        tl, _ = synth.make_timestamps(nl, seed=trial) if nl else (np.zeros(0), None)
        tr, _ = synth.make_timestamps(nr, phase=0.011, seed=1000 + trial) if nr else (np.zeros(0), None)
        if trial % 5 == 0 and nr > 3:  # exact ties and duplicate timestamps
            tr = np.round(tr, 2)
            tl = np.round(tl, 2)
        a = api.pair_nearest(tl, tr, 0.0167)
        b = oracle.pair_nearest(tl, tr, 0.0167)
        assert a[0].tolist() == b[0].tolist() and a[1].tolist() == b[1].tolist(), trial


def test_pair_nearest_properties():
    tl, _ = synth.make_timestamps(10000, seed=1)
    tr, _ = synth.make_timestamps(10000, phase=0.011, seed=2)
    li, ri = api.pair_nearest(tl, tr, 0.0167)
    assert len(li) > 9000
    assert (np.diff(li) > 0).all() and (np.diff(ri) > 0).all()  # each frame used once, order kept
    assert (np.abs(tl[li] - tr[ri]) <= 0.0167).all()
    with pytest.raises(api.UsvError):
        api.pair_nearest([0.0, 2.0, 1.0], [0.0], 1.0)  # not ascending
