"""Episode storage for controller training: a pre-decoded, memory-mappable shard per episode and an
HBM-resident store with a DinoV2 feature cache.

Reference: the trainer reads LZF-compressed HDF5 episodes through `ControllerDataset` (controller_dataset.py:30-236; schema
written by data/create_controller_dataset_episode.py:179-188 and data/franka_data/3_gelsight_data.py:57-61) with 4 DataLoader
workers: every sample re-opens a file, decompresses 2 x 2 camera frames of 442 KB, converts the whole episode's quaternions
through scipy (scripts/utils_eef.py:80-90) and slices a few rows.  At batch 256 per GPU that loader, not the GPU, sets the
step time; and because the DinoV2 encoder is frozen (visual_encoder.py:33-36) its output for a frame never changes.

* `.vtep` shard (`write_episode_shard` / `EpisodeShard`): the episode's arrays stored raw, 4096-byte aligned, behind a JSON
  index -- opened with np.memmap, sliced without decoding.  It also holds `qpos10`, the episode's
  `converted_ee_pose_with_gripper` (float64, computed once with the same scipy calls the reference makes per sample).
  `EpisodeShard` answers the same `f['ee_poses']`, `f['gelsight_force']['forces']`, ... look-ups as an h5py.File, so
  `ControllerDataset` runs unchanged on either (h5py itself is optional: it is not installed in this image).
* `DeviceEpisodeStore`: every stream of every episode concatenated over frames in device memory, camera frames as uint8 or --
  the point -- as cached DinoV2 features `[frames][camera][branch][D]` for BOTH normalisation branches of
  `_normalize_images` (visual_encoder.py:95-106), so that the reference's batch-global `mean < 0.5` predicate can still be
  evaluated per minibatch from per-frame means.  `gather(indices)` assembles a collated, normalised minibatch in one launch
  (csrc/vt_dataset.cuh, `vt_batch_gather`).  At cfg2 the frozen DinoV2 forward is 10.5 of the 19.2 ms training step; with the
  cache it is paid once per frame instead of once per sample per epoch.
"""
from __future__ import annotations

import ctypes
import json
import os
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

MAGIC = b"VTEP0001"
ALIGN = 4096
STREAMS = ("ee_poses", "gripper_pos", "vla_action", "gelsight_force/forces", "gelsight_force/displacement",
           "camera1_resized", "camera2_resized")


# --------------------------------------------------------------------------------------------------------------------
# pose conversion (scripts/utils_eef.py:80-100, docs/test_6drot.py:5-27,72-79)
# --------------------------------------------------------------------------------------------------------------------
def converted_ee_pose_with_gripper(epi) -> np.ndarray:
    """[N,7] position + xyzw quaternion, [N] gripper -> [N,10] = position, 6-D rotation (first two matrix columns), gripper.
    Follows the reference call for call, including its normalisation of the quaternion block by ONE Frobenius norm
    (docs/test_6drot.py:10; harmless: Rotation.from_quat normalises every row again) and the quaternion -> Euler -> matrix
    round trip."""
    from scipy.spatial.transform import Rotation as R
    poses = np.asarray(epi["ee_poses"][:])
    pos, quat = poses[:, :3], poses[:, 3:]
    quat = quat / np.linalg.norm(quat)
    euler = R.from_quat(quat).as_euler("xyz")
    rot = R.from_euler("xyz", euler).as_matrix()
    six = rot[:, :, :2].transpose(0, 2, 1).reshape(rot.shape[0], -1)
    grip = np.asarray(epi["gripper_pos"][:]).reshape(-1, 1)
    return np.concatenate((pos, six, grip), axis=-1)


# --------------------------------------------------------------------------------------------------------------------
# .vtep shards
# --------------------------------------------------------------------------------------------------------------------
def _get(epi, path: str):
    node = epi
    for k in path.split("/"):
        node = node[k]
    return node


def _has(epi, path: str) -> bool:
    try:
        _get(epi, path)
        return True
    except (KeyError, ValueError):
        return False


def write_episode_shard(epi, path: str, with_images: bool = True) -> str:
    """epi: an open h5py.File or any nested mapping with the reference's schema.  Writes `path` (.vtep) atomically."""
    arrays: List[Tuple[str, np.ndarray]] = []
    for name in STREAMS:
        if name.startswith("camera") and not with_images:
            continue
        if _has(epi, name):
            arrays.append((name, np.ascontiguousarray(_get(epi, name)[:])))
    arrays.append(("qpos10", np.ascontiguousarray(converted_ee_pose_with_gripper(epi))))
    index, off = [], 0
    for name, a in arrays:
        index.append({"name": name, "dtype": a.dtype.str, "shape": list(a.shape), "offset": off})
        off += (a.nbytes + ALIGN - 1) // ALIGN * ALIGN
    head = json.dumps({"arrays": index}).encode()
    data0 = (len(MAGIC) + 8 + len(head) + ALIGN - 1) // ALIGN * ALIGN
    tmp = path + ".tmp"
    with open(tmp, "wb") as f:
        f.write(MAGIC)
        f.write(np.uint64(len(head)).tobytes())
        f.write(head)
        for rec, (_, a) in zip(index, arrays):
            f.seek(data0 + rec["offset"])
            f.write(a.tobytes())
        f.truncate(data0 + off)
    os.replace(tmp, path)
    return path


class _Group(dict):
    pass


class _Array:
    """Read-only view of one stored array with h5py.Dataset's behaviour where the dataset code relies on it: indexing returns
    a fresh ndarray (the reference divides the result in place, controller_dataset.py:124,130,202-204)."""

    def __init__(self, mm: np.ndarray):
        self._a = mm
        self.shape, self.dtype = mm.shape, mm.dtype

    def __getitem__(self, key):
        return np.array(self._a[key])

    def __len__(self):
        return self.shape[0]

    def __array__(self, dtype=None, copy=None):
        return np.array(self._a, dtype=dtype)

    def raw(self) -> np.ndarray:
        return self._a


class EpisodeShard:
    """`with EpisodeShard(path) as f: f['ee_poses'][a:b]` -- the h5py.File look-ups the dataset makes, on a memory map."""

    def __init__(self, path: str, mode: str = "r"):
        if mode != "r":
            raise ValueError("EpisodeShard is read-only")
        with open(path, "rb") as f:
            if f.read(len(MAGIC)) != MAGIC:
                raise ValueError(f"{path}: not a .vtep episode shard")
            n = int(np.frombuffer(f.read(8), dtype=np.uint64)[0])
            head = json.loads(f.read(n).decode())
        data0 = (len(MAGIC) + 8 + n + ALIGN - 1) // ALIGN * ALIGN
        self.path, self._root = path, _Group()
        for rec in head["arrays"]:
            shape = tuple(rec["shape"])
            if int(np.prod(shape)) == 0:
                mm = np.zeros(shape, dtype=np.dtype(rec["dtype"]))
            else:
                mm = np.memmap(path, mode="r", dtype=np.dtype(rec["dtype"]), shape=shape, offset=data0 + rec["offset"])
            node, parts = self._root, rec["name"].split("/")
            for k in parts[:-1]:
                node = node.setdefault(k, _Group())
            node[parts[-1]] = _Array(mm)

    def __getitem__(self, key):
        return self._root[key]

    def __contains__(self, key):
        return key in self._root

    def keys(self):
        return self._root.keys()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


def open_episode(path: str):
    """.vtep -> EpisodeShard; .h5 / .hdf5 -> h5py.File (h5py is an optional dependency of the HDF5 side only)."""
    if path.endswith(".vtep"):
        return EpisodeShard(path)
    try:
        import h5py
    except ImportError as e:
        raise ImportError(f"reading {path} needs h5py, which is not installed: convert the episodes once with "
                          "vla_touch_b200.episode_store.convert_directory(...) where h5py is available, or install it") from e
    return h5py.File(path, "r")


def episode_qpos10(f) -> np.ndarray:
    """converted_ee_pose_with_gripper of an open episode: the stored copy of a shard, else computed."""
    if isinstance(f, EpisodeShard) and "qpos10" in f:
        return f["qpos10"][:]
    return converted_ee_pose_with_gripper(f)


def convert_directory(src_dir: str, dst_dir: str, with_images: bool = True) -> List[str]:
    """Every *.h5 under src_dir -> dst_dir/<same relative name>.vtep (needs h5py)."""
    out = []
    for root, _, files in os.walk(src_dir):
        for fn in sorted(files):
            if not fn.endswith(".h5"):
                continue
            rel = os.path.relpath(os.path.join(root, fn), src_dir)
            dst = os.path.join(dst_dir, os.path.splitext(rel)[0] + ".vtep")
            os.makedirs(os.path.dirname(dst) or ".", exist_ok=True)
            with open_episode(os.path.join(root, fn)) as f:
                out.append(write_episode_shard(f, dst, with_images))
    return out


# --------------------------------------------------------------------------------------------------------------------
# HBM-resident store
# --------------------------------------------------------------------------------------------------------------------
class DeviceEpisodeStore:
    """All episodes of a ControllerDataset, concatenated over frames, in device memory.

    dataset: vla_touch_b200.controller_dataset.ControllerDataset (its file list, index mapping, context_frames and horizon
    define the samples; sample i of the store IS sample i of the dataset).
    image_encoder: a DINOv2Encoder -> the camera frames are turned into the feature cache (both normalisation branches) and
    dropped; None -> no visual features (bridge_controller_no_visual).
    """

    def __init__(self, dataset, device="cuda", image_encoder=None, feature_chunk: int = 64, keep_displacements: bool = True):
        from . import native as nv
        nv.lib()
        if not torch.cuda.is_available():
            raise nv.NativeError("DeviceEpisodeStore needs a CUDA (B200) device; vla_touch_b200 has no host fallback")
        self.device = torch.device(device)
        self.context_frames, self.horizon = dataset.context_frames, dataset.horizon
        qpos, grip, vla, vla_last, forces, disps = [], [], [], [], [], []
        self.episode_offset: List[int] = []
        off = 0
        self._feat_parts, self._mean_parts = [], []
        self.D = 0
        for path in dataset.file_paths:
            with open_episode(path) as f:
                q = episode_qpos10(f)                                   # float64 [N,10]
                n = q.shape[0]
                self.episode_offset.append(off)
                off += n
                qpos.append(q.astype(np.float32))
                grip.append((q[:, -1] / 255).astype(np.float32))       # divided in the source precision, THEN cast (:124,:149-150)
                v = f["vla_action"][:]
                vl = v[:, :, -1].copy()
                vl /= 255                                               # in the stored dtype like the reference (:130)
                vla.append(v.astype(np.float32))
                vla_last.append(vl.astype(np.float32))
                forces.append(np.asarray(f["gelsight_force"]["forces"][:], dtype=np.float32).reshape(n, -1))
                if keep_displacements and "displacement" in f["gelsight_force"]:
                    disps.append(np.asarray(f["gelsight_force"]["displacement"][:], dtype=np.float32).reshape(n, -1))
                if image_encoder is not None:
                    self._cache_features(image_encoder, f["camera1_resized"], f["camera2_resized"], n, feature_chunk)
        self.frames = off
        up = lambda parts: torch.from_numpy(np.ascontiguousarray(np.concatenate(parts, axis=0))).to(self.device)
        self.qpos, self.grip_scaled, self.vla, self.vla_last_scaled, self.forces = up(qpos), up(grip), up(vla), up(vla_last), up(forces)
        self.disps = up(disps) if disps else None
        self.A, self.vla_T, self.Fd = self.qpos.shape[1], self.vla.shape[1], self.forces.shape[1]
        self.Dd = self.disps.shape[1] if self.disps is not None else 0
        self.feats = torch.cat(self._feat_parts, dim=0).contiguous() if self._feat_parts else None
        self.frame_mean = torch.cat(self._mean_parts, dim=0).contiguous() if self._mean_parts else None
        del self._feat_parts, self._mean_parts
        starts = np.array([self.episode_offset[fi] + s for fi, s in dataset.episode_indices], dtype=np.int64)
        self.sample_start = torch.from_numpy(starts).to(self.device)
        self._stats = None
        self.set_stats(dataset.stats)

    # -- feature cache --------------------------------------------------------------------------------------------
    def _cache_features(self, enc, cam1, cam2, n: int, chunk: int) -> None:
        """DINOv2Encoder.forward of every frame of both cameras under both outcomes of the `mean < 0.5` predicate.  The
        dataset hands the encoder float32 frames in [0, 1] (uint8 / 255, controller_dataset.py:167-168), for which the
        encoder's `max > 1` test is false; the uint8 frames go to the kernels with that division forced (flags[0] = 1: the
        same fp32 value, patchify's uint8 table is built with the reference's IEEE operations)."""
        from . import native as nv
        from .dino import DinoProgram
        from .plan import Plan
        H, W = cam1.shape[1], cam1.shape[2]
        self.D = enc.hidden_size
        feats = torch.empty((n, 2, 2, self.D), dtype=torch.float32, device=self.device)
        means = torch.empty((n, 2), dtype=torch.float64, device=self.device)
        progs = getattr(self, "_progs", None)
        if progs is None:
            progs = self._progs = {}
        for a in range(0, n, chunk):
            b = min(n, a + chunk)
            key = (b - a, H, W)
            if key not in progs:
                plan = Plan(self.device)
                progs[key] = (plan, DinoProgram(plan, enc.weights(), 2, b - a, H, W, torch.uint8, nv.LAYOUT_BHWC, host_flags=True))
            plan, prog = progs[key]
            for c, cam in enumerate((cam1, cam2)):
                frames = torch.from_numpy(np.ascontiguousarray(cam[a:b])).to(self.device)
                prog.img[c].copy_(frames)
                means[a:b, c] = frames.to(torch.float64).mean(dim=(1, 2, 3)) / 255.0
            for br in (0, 1):
                prog.flags.copy_(torch.tensor([[1, br, 0, 0]] * 2, dtype=torch.int32))
                plan.compile().run()
                feats[a:b, :, br] = prog.feat.permute(1, 0, 2)
        self._feat_parts.append(feats)
        self._mean_parts.append(means)

    def save_feature_cache(self, path: str, encoder_tag: str = "") -> str:
        """Writes the feature cache (features, per-frame means, frame count per episode, a caller-chosen tag naming the encoder
        weights) so that a later run can skip the DinoV2 pass: `DeviceEpisodeStore(ds)` + `load_feature_cache(path)`."""
        if self.feats is None:
            raise ValueError("this store holds no feature cache")
        ends = self.episode_offset[1:] + [self.frames]
        np.savez(path, feats=self.feats.cpu().numpy(), frame_mean=self.frame_mean.cpu().numpy(),
                 episode_frames=np.array([e - o for o, e in zip(self.episode_offset, ends)], dtype=np.int64), tag=np.array(encoder_tag))
        return path if path.endswith(".npz") else path + ".npz"

    def load_feature_cache(self, path: str, encoder_tag: str = "") -> None:
        """Rejects a cache whose episode layout or encoder tag differs from this store's (a stale cache must not train silently)."""
        z = np.load(path)
        ends = self.episode_offset[1:] + [self.frames]
        mine = np.array([e - o for o, e in zip(self.episode_offset, ends)], dtype=np.int64)
        if not np.array_equal(z["episode_frames"], mine) or z["feats"].shape[0] != self.frames:
            raise ValueError(f"{path}: feature cache was written for a different set of episodes")
        if str(z["tag"]) != encoder_tag:
            raise ValueError(f"{path}: feature cache was written for encoder '{z['tag']}', not '{encoder_tag}'")
        self.feats = torch.from_numpy(z["feats"]).to(self.device).contiguous()
        self.frame_mean = torch.from_numpy(z["frame_mean"]).to(self.device).contiguous()
        self.D = self.feats.shape[-1]

    # -- stats ----------------------------------------------------------------------------------------------------
    def set_stats(self, stats: Optional[Dict]) -> None:
        if stats is None:
            self._stats = None
            return
        conv = lambda k: torch.as_tensor(stats[k], dtype=torch.float32).to(self.device).contiguous()
        self._stats = {k: conv(k) for k in ("action_mins", "action_maxs", "vla_mins", "vla_maxs")}

    def __len__(self) -> int:
        return self.sample_start.numel()

    # -- minibatch ------------------------------------------------------------------------------------------------
    def gather(self, indices, padding_factor: float = 1.4, with_displacements: bool = True) -> Dict[str, torch.Tensor]:
        """indices: sample numbers (list / numpy / tensor).  Returns the collated batch of ControllerDataset items
        ('states', 'vla_actions', 'expert_actions', 'forces', 'disps') plus 'vla_act' / 'expert_act' (normalised chunks,
        bridge_train.py:135-136) and, with a feature cache, 'feat_cam1' / 'feat_cam2' ([B, D], what
        image_encoder.forward(images[:, -1]) returns for this batch) and 'branch'."""
        from . import native as nv
        if torch.is_tensor(indices) and indices.is_cuda:       # no host sync: torch's device-side index assert guards the range
            idx = indices.to(torch.int64)
        else:
            host = np.asarray(indices.cpu() if torch.is_tensor(indices) else indices, dtype=np.int64).reshape(-1)
            if host.size and (host.min() < 0 or host.max() >= len(self)):
                raise IndexError("sample index out of range")
            idx = torch.from_numpy(host).to(self.device, non_blocking=True)
        if idx.numel() == 0:
            raise ValueError("empty minibatch")
        start = self.sample_start[idx].contiguous()
        B, L, H, A = idx.numel(), self.context_frames + self.horizon, self.horizon, self.A
        new = lambda *s: torch.empty(s, dtype=torch.float32, device=self.device)
        out = {"states": new(B, L, A), "expert_actions": new(B, H, A), "vla_actions": new(B, H, A), "forces": new(B, L, self.Fd)}
        d = nv.BatchGatherDesc()
        d.qpos, d.grip_scaled, d.vla, d.vla_last_scaled = self.qpos.data_ptr(), self.grip_scaled.data_ptr(), self.vla.data_ptr(), self.vla_last_scaled.data_ptr()
        d.forces, d.start = self.forces.data_ptr(), start.data_ptr()
        d.B, d.A, d.vla_T, d.Fd, d.Dd, d.D = B, A, self.vla_T, self.Fd, self.Dd, self.D
        d.context_frames, d.horizon = self.context_frames, H
        d.states, d.expert_actions, d.vla_actions, d.forces_out = (out[k].data_ptr() for k in ("states", "expert_actions", "vla_actions", "forces"))
        if self.disps is not None and with_displacements:
            out["disps"] = new(B, L, self.Dd // 2, 2)
            d.disps, d.disps_out = self.disps.data_ptr(), out["disps"].data_ptr()
        if self.feats is not None:
            out["feat_cam1"], out["feat_cam2"] = new(B, self.D), new(B, self.D)
            out["branch"] = torch.empty(2, dtype=torch.int32, device=self.device)
            d.feats, d.frame_mean = self.feats.data_ptr(), self.frame_mean.data_ptr()
            d.feat_cam1, d.feat_cam2, d.branch = out["feat_cam1"].data_ptr(), out["feat_cam2"].data_ptr(), out["branch"].data_ptr()
        if self._stats is not None:
            out["expert_act"], out["vla_act"] = new(B, H, A), new(B, H, A)
            s = self._stats
            d.action_mins, d.action_maxs, d.vla_mins, d.vla_maxs = (s[k].data_ptr() for k in ("action_mins", "action_maxs", "vla_mins", "vla_maxs"))
            d.pad, d.expert_n, d.vla_n = padding_factor, out["expert_act"].data_ptr(), out["vla_act"].data_ptr()
        nv.check(nv.lib().vt_batch_gather(ctypes.byref(d), nv.current_stream_ptr()))
        self._keep = (start, d)
        return out


if __name__ == "__main__":      # python -m vla_touch_b200.episode_store <dir with *.h5> <dir for *.vtep> [--no-images]
    import sys
    if len(sys.argv) < 3:
        raise SystemExit("usage: python -m vla_touch_b200.episode_store SRC_DIR DST_DIR [--no-images]")
    done = convert_directory(sys.argv[1], sys.argv[2], with_images="--no-images" not in sys.argv)
    print(f"wrote {len(done)} shards under {sys.argv[2]}")
