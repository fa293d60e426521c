"""Synthetic frame-pair and timestamp generator.

Replaces the reference's live capture (two free-running USB webcams,
P/Main.cpp:876-879 capture + timestamp) with seeded, already-rectified pairs:
the right frame is the left frame shifted horizontally by a known number of
pixels, so the true disparity of every window is known.
"""
import numpy as np

DEFAULT_SEED = 325


def _lowpass(rng, n, h, w, c, cell=8):
    """Blocky low-frequency field in [0, 255], bilinear-free (nearest upsample)."""
    gh, gw = -(-h // cell) + 1, -(-w // cell) + 1
    coarse = rng.integers(0, 256, size=(n, gh, gw, c), dtype=np.uint8)
    return np.repeat(np.repeat(coarse, cell, axis=1), cell, axis=2)[:, :h, :w]


def make_pairs(n_pairs, width=640, height=480, channels=1, shift=37, noise_sigma=0.0,
               seed=DEFAULT_SEED, row_align=16):
    """Return (left, right) uint8 arrays [n, H, W] (C=1) or [n, H, W, C].

    left  = 1/2 uniform noise + 1/2 low-pass noise (textured and flat regions);
    right[y, x] = left[y, x + shift] (an object seen by the left camera at x
    appears at x' = x - shift in the right camera, i.e. LeftCam disparity =
    shift), fresh noise in the uncovered band, optional N(0, sigma) sensor
    noise. Rows are padded so that the row stride is a multiple of row_align;
    the returned arrays are views [:, :, :W].
    """
    rng = np.random.default_rng(seed)
    c = channels
    row_bytes = width * c
    stride = -(-row_bytes // row_align) * row_align
    fine = rng.integers(0, 256, size=(n_pairs, height, width, c), dtype=np.uint8)
    low = _lowpass(rng, n_pairs, height, width, c)
    left_v = ((fine.astype(np.uint16) + low.astype(np.uint16)) >> 1).astype(np.uint8)
    right_v = rng.integers(0, 256, size=(n_pairs, height, width, c), dtype=np.uint8)
    if shift >= 0:
        right_v[:, :, : width - shift] = left_v[:, :, shift:]
    else:
        right_v[:, :, -shift:] = left_v[:, :, : width + shift]
    if noise_sigma > 0:
        nz = rng.normal(0.0, noise_sigma, size=right_v.shape)
        right_v = np.clip(np.rint(right_v.astype(np.float64) + nz), 0, 255).astype(np.uint8)

    def pad(v):
        buf = np.zeros((n_pairs, height, stride), np.uint8)
        buf[:, :, :row_bytes] = v.reshape(n_pairs, height, row_bytes)
        view = buf[:, :, :row_bytes]
        return view if c == 1 else view.reshape(n_pairs, height, width, c)

    return pad(left_v), pad(right_v)


def make_timestamps(n_frames, fps=30.0, jitter_sigma=0.002, phase=0.0, drop_prob=0.01, seed=DEFAULT_SEED):
    """Ascending capture timestamps (seconds) of one free-running camera.

    nominal period 1/fps, per-frame jitter N(0, jitter_sigma), constant phase
    offset, random frame drops. Returns (timestamps, frame_ids).
    """
    rng = np.random.default_rng(seed)
    ids = np.arange(n_frames)
    t = ids / fps + phase + rng.normal(0.0, jitter_sigma, n_frames)
    keep = rng.random(n_frames) >= drop_prob
    t, ids = t[keep], ids[keep]
    order = np.argsort(t, kind="stable")
    return np.ascontiguousarray(t[order]), np.ascontiguousarray(ids[order])
