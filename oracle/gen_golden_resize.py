"""Writes tests/golden/resize_*.npz from the UNMODIFIED reference function (scripts/utils_eef.py:44-77) and the real cv2:
small frames with every INTER_AREA down-scaling path (fractional factor, factor 2, factor 3, factor 1, H > W and W > H padding) as
full input / output arrays, and the deployment frame size (480 x 640 -> 384) as a seeded input with a SHA-256 of the output.
Run here (needs /root/reference and cv2):  python oracle/gen_golden_resize.py"""
import hashlib
import importlib.util
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = [("frac_w", 60, 80, 48), ("frac_h", 100, 77, 48), ("x2", 96, 80, 48), ("x3", 100, 144, 48), ("x1", 48, 40, 48), ("frac_big", 130, 201, 64)]
BIG = [("deploy_480x640", 480, 640, 384, 7), ("hd_720x1280", 720, 1280, 384, 8)]


def frame(h, w, seed):
    return np.random.default_rng(seed).integers(0, 256, (h, w, 3), dtype=np.uint8)


def main():
    spec = importlib.util.spec_from_file_location("utils_eef", "/root/reference/VLA/scripts/utils_eef.py")
    ref = importlib.util.module_from_spec(spec)
    sys.path.insert(0, "/root/reference/VLA")                # the reference file does `from docs.test_6drot import *`
    spec.loader.exec_module(ref)
    out = {}
    for i, (name, h, w, t) in enumerate(CASES):
        img = frame(h, w, 100 + i)
        out[f"{name}_in"] = img
        out[f"{name}_out"] = ref.pad_and_resize_for_siglip(img, t)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "resize_small.npz"), **out)
    dig = {}
    for name, h, w, t, seed in BIG:
        res = ref.pad_and_resize_for_siglip(frame(h, w, seed), t)
        dig[name] = np.frombuffer(hashlib.sha256(res.tobytes()).digest(), dtype=np.uint8)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "resize_digests.npz"), **dig)
    print("wrote", len(out) // 2, "small cases and", len(dig), "digests")


if __name__ == "__main__":
    main()
