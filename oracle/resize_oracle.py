"""TEST INFRASTRUCTURE (oracle): CPU restatement of `pad_and_resize_for_siglip` (reference scripts/utils_eef.py:44-77) in numpy.

The arithmetic lives in an un-vendored dependency, OpenCV's cv2.resize(..., interpolation=cv2.INTER_AREA) (opencv-python, unpinned
by the reference; 4.13.0 in this image).  Restated from its published algorithm (modules/imgproc/src/resize.cpp:
computeResizeAreaTab, ResizeArea_Invoker, ResizeAreaFast_Invoker) and PINNED: oracle/gen_golden_resize.py runs the reference
function with the real cv2 and stores inputs + outputs in tests/golden/resize_*.npz; tests/test_oracle_golden.py holds this
restatement to them bit for bit.  Only tests/ may import this file."""
import math

import numpy as np


def _area_tab(ssize: int, dsize: int, scale: float):
    """computeResizeAreaTab: (destination index, source index, fp32 weight) in the order OpenCV accumulates them"""
    tab = []
    for dx in range(dsize):
        fsx1 = dx * scale
        fsx2 = fsx1 + scale
        cell = min(scale, ssize - fsx1)
        sx1, sx2 = math.ceil(fsx1), math.floor(fsx2)
        sx2 = min(sx2, ssize - 1)
        sx1 = min(sx1, sx2)
        if sx1 - fsx1 > 1e-3:
            tab.append((dx, sx1 - 1, np.float32((sx1 - fsx1) / cell)))
        for sx in range(sx1, sx2):
            tab.append((dx, sx, np.float32(1.0 / cell)))
        if fsx2 - sx2 > 1e-3:
            tab.append((dx, sx2, np.float32(min(min(fsx2 - sx2, 1.0), cell) / cell)))
    return tab


def resize_area_square(img: np.ndarray, D: int) -> np.ndarray:
    """cv2.resize(img, (D, D), interpolation=cv2.INTER_AREA) for a square uint8 image with side >= D"""
    H, W, C = img.shape
    assert H == W and H >= D
    scale = W / D
    if W % D == 0:                                   # ResizeAreaFast: integer factor
        k = W // D
        s = img.reshape(D, k, D, k, C).astype(np.int64).sum(axis=(1, 3))
        if k == 2:
            return ((s + 2) >> 2).astype(np.uint8)
        v = s.astype(np.float32) * np.float32(1.0 / (k * k))
        return np.clip(np.rint(v), 0, 255).astype(np.uint8)
    tab = _area_tab(W, D, scale)
    S = img.astype(np.float32)
    out = np.zeros((D, D, C), np.uint8)
    rows = {}

    def hrow(sy):
        if sy not in rows:
            b = np.zeros((D, C), np.float32)
            for dx, sx, a in tab:
                b[dx] = b[dx] + S[sy, sx] * a          # separately rounded multiply and add, table order
            rows[sy] = b
        return rows[sy]

    sums, prev = None, -1
    for dy, sy, beta in tab:
        b = hrow(sy)
        if dy != prev:
            if prev >= 0:
                out[prev] = np.clip(np.rint(sums), 0, 255).astype(np.uint8)
            sums, prev = beta * b, dy
        else:
            sums = sums + beta * b
    out[prev] = np.clip(np.rint(sums), 0, 255).astype(np.uint8)
    return out


def pad_and_resize_for_siglip(image, target_size: int = 384):
    """scripts/utils_eef.py:44-77"""
    if image is None:
        return None
    h, w, c = image.shape
    m = max(h, w)
    sq = np.zeros((m, m, c), dtype=image.dtype)
    ph, pw = (m - h) // 2, (m - w) // 2
    sq[ph:ph + h, pw:pw + w, :] = image
    return resize_area_square(sq, target_size)
