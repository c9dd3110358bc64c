"""normalize_actions / denormalize_actions with the reference's signature and error behaviour
(controller_dataset.py:303-384): affine map onto [-1, 1] over the min/max range padded by `padding_factor`.
On CUDA tensors the map runs in the AFFINE kernel of libvt_b200 (bit-identical to the PyTorch fp32 arithmetic:
same operation order, no FMA contraction).

ControllerDataset / ControllerDataModule (controller_dataset.py:30-236, 386-476) keep the reference's
constructor arguments, index mapping, item dictionaries and statistics, and read either the reference's HDF5 episodes (needs
h5py) or the pre-decoded `.vtep` shards of episode_store.py; `device_store()` puts the whole dataset into HBM (with a DinoV2
feature cache) and `EpisodeBatchSampler` is the DistributedSampler-equivalent over `episode_indices`."""
from __future__ import annotations

import fnmatch
import os
import re
from typing import Iterator, List

import numpy as np
import torch

from . import native as nv
from .plan import ptr


def _stats(stats, action_type, device):
    if action_type == 'expert':
        mins, maxs = stats['action_mins'], stats['action_maxs']
    elif action_type == 'vla':
        mins, maxs = stats['vla_mins'], stats['vla_maxs']
    else:
        raise ValueError(f"Unknown action_type: {action_type}")
    conv = lambda v: torch.as_tensor(v, dtype=torch.float32).to(device).contiguous()
    return conv(mins), conv(maxs)


def _affine(actions, stats, action_type, padding_factor, denorm):
    if not torch.is_tensor(actions):
        actions = torch.as_tensor(actions, dtype=torch.float32)
    if actions.device.type != "cuda":
        raise nv.NativeError("vla_touch_b200 runs on a CUDA (B200) device only: move `actions` to the GPU")
    mins, maxs = _stats(stats, action_type, actions.device)
    x = actions.to(torch.float32).contiguous()
    A = x.shape[-1]
    if mins.numel() != A:
        raise ValueError(f"stats have {mins.numel()} dims, actions have {A}")
    out = torch.empty_like(x)
    d = nv.AffineDesc()
    d.x, d.out, d.mins, d.maxs, d.rows, d.A, d.denorm, d.pad = ptr(x), ptr(out), ptr(mins), ptr(maxs), x.numel() // A, A, denorm, padding_factor
    prog = nv.Program()
    prog.add(d)
    prog.run()
    return out


def normalize_actions(actions, stats, action_type='expert', padding_factor=1.4):
    return _affine(actions, stats, action_type, padding_factor, 0)


def denormalize_actions(normalized_actions, stats, action_type='expert', padding_factor=1.4):
    return _affine(normalized_actions, stats, action_type, padding_factor, 1)


# --------------------------------------------------------------------------------------------------------------------
# ControllerDataset / ControllerDataModule (controller_dataset.py:17-236, 386-476)
# --------------------------------------------------------------------------------------------------------------------
EPISODE_PATTERNS = ("*.vtep", "*.h5")


def natural_sort_filenames(file_list):
    """episode_2 before episode_10 (controller_dataset.py:17-28); names without a number sort as 0, stably."""
    def number(name):
        m = re.search(r'episode_(\d+)', name)
        return int(m.group(1)) if m else 0
    return sorted(file_list, key=number)


def find_episode_files(data_dir: str) -> List[str]:
    """os.walk order, natural sort inside a directory (controller_dataset.py:60-64).  A directory that holds `.vtep` shards is
    read from those; otherwise from its `.h5` files."""
    paths = []
    for root, _, files in os.walk(data_dir):
        for pat in EPISODE_PATTERNS:
            hit = natural_sort_filenames(fnmatch.filter(files, pat))
            if hit:
                paths += [os.path.join(root, f) for f in hit]
                break
    return paths


class ControllerDataset(torch.utils.data.Dataset):
    """Samples = every `stride`-th start frame from the first frame where the end effector has moved by more than 1e-2 in any
    pose coordinate, with context_frames + horizon frames available (controller_dataset.py:72-96)."""

    def __init__(self, data_dir, file_paths=None, context_frames=2, horizon=8, use_images=False, image_size=384, stride=1):
        self.data_dir, self.context_frames, self.horizon = data_dir, context_frames, horizon
        self.use_images, self.image_size, self.stride = use_images, image_size, stride
        self.file_paths = find_episode_files(data_dir) if file_paths is None else file_paths
        self.create_index_mapping()
        self.stats = self.get_normalization_stats()

    def create_index_mapping(self):
        from .episode_store import open_episode
        self.episode_indices, self.total_samples = [], 0
        for file_idx, path in enumerate(self.file_paths):
            with open_episode(path) as f:
                poses = np.asarray(f['ee_poses'][:])
                moved = np.where(np.any(np.abs(poses - poses[0:1]) > 1e-2, axis=1))[0]
                if len(moved) == 0:
                    print(f"Warning: No movement detected in file {path}. Skipping.")
                    continue
                for start in range(moved[0], poses.shape[0] - (self.context_frames + self.horizon - 1), self.stride):
                    self.episode_indices.append((file_idx, start))
                    self.total_samples += 1

    def __len__(self):
        return self.total_samples

    def __getitem__(self, idx):
        from .episode_store import episode_qpos10, open_episode
        file_idx, start = self.episode_indices[idx]
        ctx, span = self.context_frames, self.context_frames + self.horizon
        with open_episode(self.file_paths[file_idx]) as f:
            qpos = episode_qpos10(f)[start:start + span]
            future = qpos[ctx:]
            future[:, -1] /= 255                      # a VIEW: the action rows of `states` carry the rescaled gripper too (:123-124)
            vla = f['vla_action'][start + ctx][:self.horizon]
            vla[:, -1] /= 255
            forces = f['gelsight_force']['forces'][start:start + span]
            disps = f['gelsight_force']['displacement'][start:start + span]
            if self.use_images:
                cam1 = np.array(f['camera1_resized'][start:start + ctx])
                cam2 = np.array(f['camera2_resized'][start:start + ctx])
        f32 = lambda a: torch.as_tensor(a, dtype=torch.float32)
        item = {'states': f32(qpos), 'vla_actions': f32(vla), 'expert_actions': f32(future), 'forces': f32(forces), 'disps': f32(disps)}
        if self.use_images:
            item['images_cam1'] = f32(cam1) / 255.0
            item['images_cam2'] = f32(cam2) / 255.0
        return item

    def get_normalization_stats(self):
        """min / max over all frames of all files of the expert actions (gripper / 255) and of the VLA chunks (:172-236)."""
        from .episode_store import episode_qpos10, open_episode
        dims = 10
        a_min, a_max = np.full(dims, np.inf), np.full(dims, -np.inf)
        v_min, v_max = np.full(dims, np.inf), np.full(dims, -np.inf)
        print(f"Computing normalization statistics across {len(self.file_paths)} files...")
        for path in self.file_paths:
            with open_episode(path) as f:
                expert = episode_qpos10(f)
                expert[:, -1] /= 255
                vla = f['vla_action'][:]
                vla[:, :, -1] /= 255
                a_min, a_max = np.minimum(a_min, np.min(expert, axis=0)), np.maximum(a_max, np.max(expert, axis=0))
                v_min, v_max = np.minimum(v_min, np.min(vla, axis=(0, 1))), np.maximum(v_max, np.max(vla, axis=(0, 1)))
        a_rng, v_rng = a_max - a_min, v_max - v_min
        a_rng[a_rng < 1e-6] = 1.0
        v_rng[v_rng < 1e-6] = 1.0
        return {'action_mins': a_min, 'action_maxs': a_max, 'vla_mins': v_min, 'vla_maxs': v_max, 'action_range': a_rng, 'vla_range': v_rng}

    # ---- B200 side ----
    def device_store(self, device="cuda", image_encoder=None, **kw):
        """The whole dataset in HBM (episode_store.DeviceEpisodeStore): store.gather(indices) replaces __getitem__ + collate +
        upload + normalize_actions; with `image_encoder` the camera frames become a DinoV2 feature cache."""
        from .episode_store import DeviceEpisodeStore
        return DeviceEpisodeStore(self, device=device, image_encoder=image_encoder, **kw)


class EpisodeBatchSampler:
    """Minibatches of sample numbers for one rank: a seeded permutation per epoch (identical on every rank), padded by wrapping to
    a multiple of world_size * batch_size unless drop_last, rank r takes every world_size-th entry -- torch's DistributedSampler
    partition followed by the reference's DataLoader batching (shuffle=True, drop_last=True for training, :451-459)."""

    def __init__(self, n_samples: int, batch_size: int, rank: int = 0, world_size: int = 1, shuffle: bool = True, seed: int = 0,
                 drop_last: bool = True):
        if not 0 <= rank < world_size:
            raise ValueError(f"rank {rank} outside world of {world_size}")
        self.n, self.batch_size, self.rank, self.world, self.shuffle, self.seed, self.drop_last = n_samples, batch_size, rank, world_size, shuffle, seed, drop_last
        self.epoch = 0

    def set_epoch(self, epoch: int) -> None:
        self.epoch = epoch

    def _rank_indices(self) -> np.ndarray:
        order = np.random.default_rng([self.seed, self.epoch]).permutation(self.n) if self.shuffle else np.arange(self.n)
        per_rank = self.n // self.world if self.drop_last else -(-self.n // self.world)
        total = per_rank * self.world
        if total > self.n and self.n > 0:
            order = np.concatenate([order, np.resize(order, total - self.n)])
        return order[:total][self.rank::self.world]

    def __len__(self) -> int:
        per_rank = self.n // self.world if self.drop_last else -(-self.n // self.world)
        return per_rank // self.batch_size if self.drop_last else -(-per_rank // self.batch_size)

    def __iter__(self) -> Iterator[np.ndarray]:
        idx = self._rank_indices()
        for i in range(len(self)):
            yield idx[i * self.batch_size:(i + 1) * self.batch_size]


class ControllerDataModule:
    """Train / validation split by FILES with numpy's global RNG (controller_dataset.py:428-443): seed numpy the same way and the
    split is the reference's.  `stats` comes from the training files (:446).  (The reference builds both datasets over ALL files first
    and then swaps their file lists, :431-443, so its `train_dataset.stats` attribute -- which nothing reads -- still holds the all-files
    statistics; here each dataset is built on its own files.)"""

    def __init__(self, data_dir, batch_size=32, num_workers=4, context_frames=2, horizon=8, use_images=True, image_size=384,
                 val_ratio=0.1, stride=1):
        self.data_dir, self.batch_size, self.num_workers = data_dir, batch_size, num_workers
        self.context_frames, self.horizon, self.use_images, self.image_size = context_frames, horizon, use_images, image_size
        self.val_ratio, self.stride = val_ratio, stride
        self.setup()

    def setup(self):
        print(f"loading dataset from {self.data_dir} ..")
        files = find_episode_files(self.data_dir)
        num_val = max(1, int(len(files) * self.val_ratio))
        val_idx = np.random.choice(len(files), num_val, replace=False)
        train_files = [files[i] for i in range(len(files)) if i not in val_idx]
        val_files = [files[i] for i in val_idx]
        kw = dict(data_dir=self.data_dir, context_frames=self.context_frames, horizon=self.horizon, use_images=self.use_images,
                  image_size=self.image_size, stride=self.stride)
        self.train_dataset = ControllerDataset(file_paths=train_files, **kw)
        self.val_dataset = ControllerDataset(file_paths=val_files, **kw)
        self.stats = self.train_dataset.stats

    def _loader(self, ds, shuffle, drop_last):
        return torch.utils.data.DataLoader(ds, batch_size=self.batch_size, shuffle=shuffle, num_workers=self.num_workers,
                                           pin_memory=True, drop_last=drop_last)

    def train_dataloader(self):
        return self._loader(self.train_dataset, True, True)

    def val_dataloader(self):
        return self._loader(self.val_dataset, False, False)
