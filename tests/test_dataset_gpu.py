"""SURVEY.md 8f row N2 on the B200: DeviceEpisodeStore.gather (csrc/vt_dataset.cuh through vt_batch_gather) against the collated
ControllerDataset items and the reference-written fixtures, bit for bit; the DinoV2 feature cache against the encoder run on the
batch's images; a training step fed from the store against the same step fed with images."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle.gen_golden_dataset import CASES, EPISODES, IMAGE  # noqa: E402  (constants only)
from vla_touch_b200 import controller_dataset as cd  # noqa: E402
from vla_touch_b200 import episode_store as es  # noqa: E402
from vla_touch_b200.synthetic import synth_episode  # noqa: E402

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
GOLD = np.load(os.path.join(ROOT, "tests", "golden", "dataset_controller.npz"))


def shards(dirname, image):
    for k, (num, n, still, moving, dark) in enumerate(EPISODES):
        epi = synth_episode(100 + k, n, image, still_frames=still, moving=moving, dark=dark)
        es.write_episode_shard(epi, os.path.join(dirname, f"episode_{num}.vtep"))
    return dirname


@pytest.fixture(scope="module")
def small_dir(tmp_path_factory):
    return shards(str(tmp_path_factory.mktemp("ep_small")), IMAGE)


@pytest.fixture(scope="module")
def big_dir(tmp_path_factory):
    return shards(str(tmp_path_factory.mktemp("ep_224")), 224)


def collate(ds, idx):
    items = [ds[int(i)] for i in idx]
    return {k: torch.stack([it[k] for it in items]) for k in items[0]}


@pytest.mark.parametrize("tag", list(CASES))
def test_gather_equals_the_collated_dataset_items(small_dir, tag):
    from vla_touch_b200.controller_dataset import normalize_actions
    ds = cd.ControllerDataset(small_dir, use_images=False, **CASES[tag])
    store = ds.device_store(DEV)
    assert len(store) == len(ds)
    g = np.random.default_rng(3)
    for idx in (np.arange(5), g.permutation(len(ds))[:min(len(ds), 37)], np.array([len(ds) - 1]), np.arange(len(ds))):
        got = store.gather(idx)
        want = collate(ds, idx)
        for k, v in want.items():
            assert got[k].shape == v.shape and torch.equal(got[k].cpu(), v), (tag, k)
        stats = {k: torch.as_tensor(v, dtype=torch.float32).to(DEV) for k, v in ds.stats.items()}
        assert torch.equal(got["expert_act"], normalize_actions(got["expert_actions"], stats, "expert"))
        assert torch.equal(got["vla_act"], normalize_actions(got["vla_actions"], stats, "vla"))
    first = store.gather(np.arange(5))                      # the reference's own loader + normalize_actions on the first 5 samples
    for k in ("states", "vla_actions", "expert_actions", "forces", "disps"):
        assert np.array_equal(first[k].cpu().numpy(), GOLD[f"{tag}.batch5.{k}"]), k
    assert np.array_equal(first["expert_act"].cpu().numpy(), GOLD[f"{tag}.norm.expert"])
    assert np.array_equal(first["vla_act"].cpu().numpy(), GOLD[f"{tag}.norm.vla"])


def test_gather_rejects_bad_requests(small_dir):
    ds = cd.ControllerDataset(small_dir, **CASES["h8"])
    store = ds.device_store(DEV)
    with pytest.raises(IndexError):
        store.gather([len(ds)])
    with pytest.raises(ValueError):
        store.gather([])
    store.set_stats(None)
    assert "vla_act" not in store.gather([0, 1])


def _encoder(layers=2):
    import vt_testutil as U
    from vla_touch_b200.visual_encoder import DINOv2Encoder
    return DINOv2Encoder("facebook/dinov2-small", device=DEV, state_dict=U.dino_sd(384, layers, 4))


def test_feature_cache_equals_the_encoder_on_the_batch(big_dir):
    """Cached features, with the branch picked from per-frame means, against DINOv2Encoder.forward on the float images the dataset
    hands out -- for a bright batch (ImageNet-normalised), a dark one (episode_7: passed through) and a mixed one."""
    ds = cd.ControllerDataset(big_dir, use_images=True, image_size=224, **CASES["h16s3"])
    enc = _encoder()
    store = ds.device_store(DEV, image_encoder=enc, feature_chunk=32)
    files = [os.path.basename(p) for p in ds.file_paths]
    by_file = {}
    for i, (fi, _) in enumerate(ds.episode_indices):
        by_file.setdefault(files[fi], []).append(i)
    dark, bright = by_file["episode_7.vtep"], by_file["episode_2.vtep"]
    worst = 0.0
    for name, idx, want_branch in (("bright", bright[:6], 1), ("dark", dark[:6], 0), ("mixed_bright", bright[:5] + dark[:2], None),
                                   ("mixed_dark", bright[:1] + dark[:6], None)):
        got = store.gather(idx)
        batch = collate(ds, idx)
        for cam, key in (("images_cam1", "feat_cam1"), ("images_cam2", "feat_cam2")):
            imgs = batch[cam][:, -1].to(DEV)
            ref = enc.forward(imgs)
            br = int(got["branch"][0 if cam == "images_cam1" else 1])
            assert br == (0 if float(imgs.mean()) < 0.5 else 1), (name, cam)
            if want_branch is not None:
                assert br == want_branch
            err = float((got[key] - ref).abs().max())
            worst = max(worst, err)
            assert err == 0.0, (name, cam, err)           # per-image results do not depend on the batch they were computed in
    print(f"feature cache vs encoder on the batch: worst |diff| {worst:.3g}")


def test_training_step_from_the_store_equals_the_step_from_images(big_dir):
    import vt_testutil as U
    from vla_touch_b200.trainer import DiffusionControllerTrainer
    ds = cd.ControllerDataset(big_dir, use_images=True, image_size=224, **CASES["h16s3"])
    A, Fd, T, seed = 10, 3, 16, 21
    c = dict(A=A, F=Fd, T=T, hidden=384, steps=10, seed=seed, dino=U.dino_sd(384, 2, seed), enc=U.enc_sd(2 * 384 + A + Fd, seed),
             stats={k: torch.as_tensor(v, dtype=torch.float32) for k, v in ds.stats.items()})
    idx = np.array([0, 3, 11, 20, 34, 7])
    B = len(idx)
    g = torch.Generator().manual_seed(5)
    step, z = torch.rand(B, generator=g).to(DEV), torch.randn(B, T, A, generator=g).to(DEV)
    ctl_a, ctl_b = U.make_controller(c, DEV), U.make_controller(c, DEV)
    for ctl in (ctl_a, ctl_b):
        ctl.diffusion_model.step_override, ctl.diffusion_model.z_override = step, z
    store = ds.device_store(DEV, image_encoder=ctl_a.image_encoder, feature_chunk=32)
    tr_a = DiffusionControllerTrainer(ctl_a, ds.stats, device=DEV)
    tr_b = DiffusionControllerTrainer(ctl_b, ds.stats, device=DEV)
    for _ in range(2):
        la = tr_a.train_step(store.gather(idx))
        lb = tr_b.train_step({k: v.to(DEV) for k, v in collate(ds, idx).items()})
        assert torch.isfinite(la["loss"])
        assert float((la["loss"] - lb["loss"]).abs()) <= 1e-4 * max(1.0, float(lb["loss"].abs()))
    for (n, pa), pb in zip(ctl_a.diffusion_model.net.named_parameters(), ctl_b.diffusion_model.net.parameters()):
        assert float((pa - pb).abs().max()) <= 1e-5 * max(1.0, float(pb.abs().max())), n
    for pa, pb in zip(ctl_a.state_encoder.parameters(), ctl_b.state_encoder.parameters()):
        assert float((pa - pb).abs().max()) <= 1e-5


def test_feature_cache_file_round_trip(big_dir, tmp_path):
    ds = cd.ControllerDataset(big_dir, use_images=True, image_size=224, **CASES["h64c1"])
    enc = _encoder()
    a = ds.device_store(DEV, image_encoder=enc, feature_chunk=64)
    path = a.save_feature_cache(str(tmp_path / "cache"), encoder_tag="synthetic-s2-seed4")
    b = ds.device_store(DEV)
    assert b.feats is None and "feat_cam1" not in b.gather([0])
    with pytest.raises(ValueError):
        b.load_feature_cache(path, encoder_tag="another encoder")
    b.load_feature_cache(path, encoder_tag="synthetic-s2-seed4")
    ga, gb = a.gather(np.arange(len(ds))), b.gather(np.arange(len(ds)))
    for k in ("feat_cam1", "feat_cam2", "branch", "states", "vla_act"):
        assert torch.equal(ga[k], gb[k]), k
    other = cd.ControllerDataset(big_dir, file_paths=ds.file_paths[:2], **CASES["h64c1"]).device_store(DEV)
    with pytest.raises(ValueError):
        other.load_feature_cache(path, encoder_tag="synthetic-s2-seed4")


def test_lstm_trainer_step_from_the_store(big_dir):
    """lstm_train.py:57-82,122-139 fed from the store: forces[:, ctx-1:-1], cached features into obs_encoder, loss finite and falling."""
    import vt_testutil as U
    from vla_touch_b200.lstm_step_controller import TactileLSTMController
    from vla_touch_b200.trainer import LSTMControllerTrainer
    ds = cd.ControllerDataset(big_dir, use_images=True, image_size=224, **CASES["h16s3"])
    lc = TactileLSTMController(state_dim=10, hidden_dim=256, device=DEV, force_dim=3, use_force=True, image_state_dict=U.dino_sd(384, 2, 4))
    store = ds.device_store(DEV, image_encoder=lc.image_encoder, feature_chunk=32)
    tr = LSTMControllerTrainer(lc, ds.stats, learning_rate=1e-3, device=DEV)
    idx = np.arange(12)
    losses = [float(tr.train_step(store.gather(idx))) for _ in range(6)]
    assert all(np.isfinite(losses)) and losses[-1] < losses[0], losses


def test_pinned_camera_batches_give_the_same_features():
    """visual_encoder.forward_two_cameras: pinned host batches (the DataLoader's pin_memory=True) take the side-stream upload path;
    features must equal the plain two-call path bit for bit, for the 5-D uint8 and the float layouts."""
    from vla_touch_b200.visual_encoder import forward_two_cameras
    enc = _encoder()
    g = torch.Generator().manual_seed(3)
    u8 = [torch.randint(90, 256, (6, 1, 224, 224, 3), generator=g, dtype=torch.uint8) for _ in range(2)]
    f32 = [t[:, 0].float() / 255.0 for t in u8]
    for a, b in (u8, f32):
        want = (enc.forward(a.to(DEV)), enc.forward(b.to(DEV)))
        for _ in range(2):                                  # twice: the second call reuses the side stream and freed blocks
            got = forward_two_cameras(enc, a.pin_memory(), b.pin_memory())
            assert torch.equal(got[0], want[0]) and torch.equal(got[1], want[1])
        plain = forward_two_cameras(enc, a, b)              # pageable host tensors: the plain path
        assert torch.equal(plain[0], want[0]) and torch.equal(plain[1], want[1])
