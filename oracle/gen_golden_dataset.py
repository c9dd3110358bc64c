"""TEST INFRASTRUCTURE.  Writes tests/golden/dataset_*.npz from the UNMODIFIED reference `ControllerDataset` /
`ControllerDataModule` (VLA/residual_controller/controller_dataset.py:30-236, 386-476) run on seeded synthetic episodes.

h5py is not installed here, so the reference's `import h5py` gets a stand-in whose `File` serves the synthetic episode that
tests re-create from the same seeds (vla_touch_b200.synthetic.synth_episode); like a real h5py.Dataset its arrays hand out
COPIES on indexing (the reference divides what it reads in place).  Run here (needs /root/reference):
    python oracle/gen_golden_dataset.py"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vla_touch_b200.synthetic import synth_episode  # noqa: E402

# (episode number, frames, still frames, moving, dark) -- episode_10 / episode_2 check the natural sort, episode_5 never moves and is
# skipped, episode_3 is shorter than context + horizon: it yields no sample but still counts in the statistics
EPISODES = [(2, 40, 3, True, False), (10, 31, 0, True, False), (5, 25, 0, False, False), (7, 90, 6, True, True), (1, 22, 2, True, False),
            (3, 9, 0, True, False)]
CASES = {"h8": dict(context_frames=2, horizon=8, stride=1), "h16s3": dict(context_frames=2, horizon=16, stride=3),
         "h64c1": dict(context_frames=1, horizon=64, stride=2)}
IMAGE = 28
_STORE = {}


class _Dataset:
    def __init__(self, a):
        self._a, self.shape, self.dtype = a, a.shape, a.dtype

    def __getitem__(self, k):
        return np.array(self._a[k])

    def __len__(self):
        return self.shape[0]

    def __array__(self, dtype=None, copy=None):
        return np.array(self._a, dtype=dtype)


def _wrap(node):
    return {k: _wrap(v) for k, v in node.items()} if isinstance(node, dict) else _Dataset(node)


class _File:
    def __init__(self, path, mode="r"):
        self._root = _wrap(_STORE[os.path.basename(path)])

    def __getitem__(self, k):
        return self._root[k]

    def __enter__(self):
        return self

    def __exit__(self, *e):
        return False


def episode_name(num):
    return f"episode_{num}.h5"


def make_store():
    for k, (num, n, still, moving, dark) in enumerate(EPISODES):
        _STORE[episode_name(num)] = synth_episode(100 + k, n, IMAGE, still_frames=still, moving=moving, dark=dark)


def main():
    make_store()
    h5 = types.ModuleType("h5py")
    h5.File = _File
    sys.modules["h5py"] = h5
    sys.modules.setdefault("tqdm", types.ModuleType("tqdm"))
    sys.path.insert(0, "/root/reference/VLA")
    sys.path.insert(0, "/root/reference/VLA/residual_controller")
    import controller_dataset as ref
    import tempfile
    td = tempfile.mkdtemp()
    for name in _STORE:                      # the reference discovers *.h5 by walking the directory
        open(os.path.join(td, name), "wb").close()
    out = {}
    for tag, kw in CASES.items():
        ds = ref.ControllerDataset(td, use_images=True, image_size=IMAGE, **kw)
        out[f"{tag}.files"] = np.array([os.path.basename(p) for p in ds.file_paths])
        out[f"{tag}.episode_indices"] = np.array(ds.episode_indices, dtype=np.int64)
        for k, v in ds.stats.items():
            out[f"{tag}.stats.{k}"] = np.asarray(v)
        n = len(ds)
        picks = sorted(set([0, 1, n // 3, n // 2, n - 2, n - 1]))
        out[f"{tag}.picks"] = np.array(picks)
        for i in picks:
            item = ds[i]
            for k, v in item.items():
                if k.startswith("images"):
                    v = v[:, ::9, ::9]          # sub-sampled pixels keep the fixture small; values are u8 / 255 either way
                out[f"{tag}.item{i}.{k}"] = v.numpy()
        loader = torch.utils.data.DataLoader(ds, batch_size=5, shuffle=False, num_workers=0)
        batch = next(iter(loader))
        for k, v in batch.items():
            if not k.startswith("images"):
                out[f"{tag}.batch5.{k}"] = v.numpy()
        stats32 = {k: torch.tensor(v, dtype=torch.float32) for k, v in ds.stats.items()}      # as the trainer holds them, bridge_train.py:74
        out[f"{tag}.norm.expert"] = ref.normalize_actions(batch["expert_actions"], stats32, "expert").numpy()
        out[f"{tag}.norm.vla"] = ref.normalize_actions(batch["vla_actions"], stats32, "vla").numpy()
    np.random.seed(7)
    dm = ref.ControllerDataModule(td, batch_size=4, num_workers=0, context_frames=2, horizon=8, use_images=False, image_size=IMAGE, val_ratio=0.3)
    out["dm.train_files"] = np.array([os.path.basename(p) for p in dm.train_dataset.file_paths])
    out["dm.val_files"] = np.array([os.path.basename(p) for p in dm.val_dataset.file_paths])
    out["dm.lens"] = np.array([len(dm.train_dataset), len(dm.val_dataset)])
    for k, v in dm.stats.items():
        out[f"dm.stats.{k}"] = np.asarray(v)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "dataset_controller.npz"), **out)
    print("wrote", len(out), "arrays;", {t: len(out[f"{t}.episode_indices"]) for t in CASES})


if __name__ == "__main__":
    main()
