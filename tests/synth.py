"""Deterministic synthetic inputs for the parity tests (numpy only, no cv2).

Recipe of SURVEY.md section 8d / App. E.1: smoothed noise stretched to mean 128 / std 48,
a low-contrast band at the bottom (forces minThFAST retries and empty cells) and
hard-edged rectangles (strong corners, score ties).
"""
import numpy as np


def _gauss_blur(a, sigma):
    r = int(np.ceil(3 * sigma))
    x = np.arange(-r, r + 1, dtype=np.float64)
    k = np.exp(-0.5 * (x / sigma) ** 2)
    k /= k.sum()
    pad = np.pad(a.astype(np.float64), ((r, r), (r, r)), mode="reflect")
    tmp = np.zeros((pad.shape[0], a.shape[1]))
    for i, kv in enumerate(k):
        tmp += kv * pad[:, i:i + a.shape[1]]
    out = np.zeros(a.shape)
    for i, kv in enumerate(k):
        out += kv * tmp[i:i + a.shape[0], :]
    return out


def synth_frame(w, h, seed, sigma=2.5, flat_frac=0.25, nrect=20):
    rng = np.random.default_rng(seed)
    n = rng.integers(0, 256, (h, w), dtype=np.uint8)
    img = _gauss_blur(n, sigma)
    img = (img - img.mean()) / img.std() * 48 + 128
    y0 = int(h * (1 - flat_frac))
    img[y0:] = (img[y0:] - 128) * 0.12 + 128
    img = np.clip(np.rint(img), 0, 255).astype(np.uint8)
    for _ in range(nrect):
        rw = min(40, max(2, w - 1)); rh = min(40, max(2, h - 1))
        x = int(rng.integers(0, max(1, w - rw))); y = int(rng.integers(0, max(1, h - rh)))
        img[y:y + int(rng.integers(min(8, rh), rh)), x:x + int(rng.integers(min(8, rw), rw))] = int(rng.integers(0, 256))
    return np.ascontiguousarray(img)


def stereo_pair(w, h, seed):
    """Right image = left shifted by a per-row-block disparity of 5..60 px, plus +-2 grey noise."""
    left = synth_frame(w, h, seed)
    rng = np.random.default_rng(seed + 7919)
    right = np.empty_like(left)
    y = 0
    while y < h:
        bh = int(rng.integers(16, 64))
        d = int(rng.integers(5, 61))
        blk = left[y:y + bh]
        sh = np.empty_like(blk)
        sh[:, :w - d] = blk[:, d:]
        sh[:, w - d:] = blk[:, w - d - 1:w - d]
        right[y:y + bh] = sh
        y += bh
    noise = rng.integers(-2, 3, right.shape)
    right = np.clip(right.astype(np.int16) + noise, 0, 255).astype(np.uint8)
    return left, np.ascontiguousarray(right)
