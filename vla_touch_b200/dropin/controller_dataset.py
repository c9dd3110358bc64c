"""Shadows the reference's controller_dataset: dataset classes (HDF5 loading, out of scope here) are taken from the
reference's own module if it is importable further down sys.path; the two normalisation functions are the B200 ones."""
import importlib.util
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
for _p in sys.path:
    _f = os.path.join(_p or ".", "controller_dataset.py")
    if os.path.abspath(_p or ".") != _here and os.path.exists(_f):
        _spec = importlib.util.spec_from_file_location("_reference_controller_dataset", _f)
        _mod = importlib.util.module_from_spec(_spec)
        try:
            _spec.loader.exec_module(_mod)
            globals().update({k: v for k, v in vars(_mod).items() if not k.startswith("__")})
        except Exception:      # e.g. h5py missing: keep only the functions below
            pass
        break

from vla_touch_b200.controller_dataset import denormalize_actions, normalize_actions  # noqa: E402,F401
