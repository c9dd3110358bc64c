"""SURVEY.md 8f row N2 on the host: ControllerDataset / ControllerDataModule / the .vtep shard format against fixtures written
by the UNMODIFIED reference classes (oracle/gen_golden_dataset.py), bit for bit; the data-parallel sampler under gloo."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.gen_golden_dataset import CASES, EPISODES, IMAGE  # noqa: E402  (constants only: nothing of the reference is imported)
from vla_touch_b200 import controller_dataset as cd  # noqa: E402
from vla_touch_b200 import episode_store as es  # noqa: E402
from vla_touch_b200.synthetic import synth_episode  # noqa: E402

GOLD = np.load(os.path.join(ROOT, "tests", "golden", "dataset_controller.npz"))


def write_shards(dirname, images=True):
    os.makedirs(dirname, exist_ok=True)
    for k, (num, n, still, moving, dark) in enumerate(EPISODES):
        epi = synth_episode(100 + k, n, IMAGE, still_frames=still, moving=moving, dark=dark)
        es.write_episode_shard(epi, os.path.join(dirname, f"episode_{num}.vtep"), with_images=images)
    return dirname


@pytest.fixture(scope="module")
def shard_dir(tmp_path_factory):
    return write_shards(str(tmp_path_factory.mktemp("episodes")))


def stems(paths):
    return [os.path.splitext(os.path.basename(str(p)))[0] for p in paths]


def test_shard_round_trip_and_h5py_like_reads(tmp_path):
    epi = synth_episode(5, 12, 14)
    p = es.write_episode_shard(epi, str(tmp_path / "episode_0.vtep"))
    with es.open_episode(p) as f:
        for name in es.STREAMS:
            node, src = f, epi
            for k in name.split("/"):
                node, src = node[k], src[k]
            assert node.shape == src.shape and node.dtype == src.dtype
            assert np.array_equal(node[:], src)
        a = f["vla_action"][3]
        a[:, -1] /= 255                                     # like an h5py read: a copy, the file is untouched
        assert np.array_equal(f["vla_action"][3], epi["vla_action"][3])
        assert np.array_equal(f["qpos10"][:], es.converted_ee_pose_with_gripper(epi))
    with pytest.raises(ValueError):
        open(tmp_path / "bad.vtep", "wb").write(b"not a shard")
        es.open_episode(str(tmp_path / "bad.vtep"))
    with pytest.raises(ImportError):
        es.open_episode(str(tmp_path / "episode_1.h5"))     # h5py is not in this image: a clear error, no silent substitute


@pytest.mark.parametrize("tag", list(CASES))
def test_dataset_reproduces_the_reference(shard_dir, tag):
    ds = cd.ControllerDataset(shard_dir, use_images=True, image_size=IMAGE, **CASES[tag])
    assert stems(ds.file_paths) == stems(GOLD[f"{tag}.files"])                       # natural sort: 1, 2, 5, 7, 10
    assert np.array_equal(np.array(ds.episode_indices, dtype=np.int64), GOLD[f"{tag}.episode_indices"])
    assert len(ds) == len(GOLD[f"{tag}.episode_indices"])
    used = {stems(ds.file_paths)[fi] for fi, _ in ds.episode_indices}
    assert "episode_5" not in used and "episode_3" not in used         # never moves / shorter than context + horizon: no samples
    for k, v in ds.stats.items():
        g = GOLD[f"{tag}.stats.{k}"]
        assert v.dtype == g.dtype and np.array_equal(v, g), k
    for i in GOLD[f"{tag}.picks"]:
        item = ds[int(i)]
        keys = {k.split(".")[-1] for k in GOLD.files if k.startswith(f"{tag}.item{i}.")}
        assert set(item) == keys
        for k, v in item.items():
            g = GOLD[f"{tag}.item{i}.{k}"]
            if k.startswith("images"):
                v = v[:, ::9, ::9]
            assert v.dtype == torch.float32 and np.array_equal(v.numpy(), g), (tag, i, k)
    ctx = CASES[tag]["context_frames"]
    item = ds[0]
    assert torch.equal(item["states"][ctx:], item["expert_actions"])                 # the reference's in-place /255 on a view


def test_data_module_split_and_stats(shard_dir):
    np.random.seed(7)
    dm = cd.ControllerDataModule(shard_dir, batch_size=4, num_workers=0, context_frames=2, horizon=8, use_images=False,
                                 image_size=IMAGE, val_ratio=0.3)
    assert stems(dm.train_dataset.file_paths) == stems(GOLD["dm.train_files"])
    assert stems(dm.val_dataset.file_paths) == stems(GOLD["dm.val_files"])
    assert [len(dm.train_dataset), len(dm.val_dataset)] == list(GOLD["dm.lens"])
    for k, v in dm.stats.items():
        assert np.array_equal(v, GOLD[f"dm.stats.{k}"]), k
    batch = next(iter(dm.train_dataloader()))
    assert batch["states"].shape == (4, 10, 10) and batch["vla_actions"].shape == (4, 8, 10)


def test_collated_batch_matches_the_reference_loader(shard_dir):
    ds = cd.ControllerDataset(shard_dir, use_images=False, **CASES["h8"])
    batch = next(iter(torch.utils.data.DataLoader(ds, batch_size=5, shuffle=False)))
    for k, v in batch.items():
        assert np.array_equal(v.numpy(), GOLD[f"h8.batch5.{k}"]), k


def test_sampler_partitions_the_samples():
    n, bs, world = 103, 8, 4
    per_rank = [np.concatenate(list(cd.EpisodeBatchSampler(n, bs, r, world, seed=3))) for r in range(world)]
    flat = np.concatenate(per_rank)
    assert all(len(p) == (n // world // bs) * bs for p in per_rank)
    assert len(set(flat.tolist())) == len(flat)                                      # disjoint across ranks
    keep = cd.EpisodeBatchSampler(n, bs, 0, world, shuffle=False, drop_last=False)
    assert len(keep) == -(-(-(-n // world)) // bs)
    full = np.concatenate([np.concatenate(list(cd.EpisodeBatchSampler(n, bs, r, world, shuffle=False, drop_last=False))) for r in range(world)])
    assert set(full.tolist()) == set(range(n))                                       # padded by wrapping: everything is covered
    s = cd.EpisodeBatchSampler(n, bs, 1, world, seed=3)
    a = np.concatenate(list(s))
    s.set_epoch(1)
    assert not np.array_equal(a, np.concatenate(list(s)))
    with pytest.raises(ValueError):
        cd.EpisodeBatchSampler(n, bs, 4, 4)


def _sampler_worker(rank, world, port, n, bs, q):
    import torch.distributed as dist
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    mine = torch.from_numpy(np.concatenate(list(cd.EpisodeBatchSampler(n, bs, rank, world, seed=11)))).long()
    got = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(got, mine)
    if rank == 0:
        q.put([g.tolist() for g in got])
    dist.destroy_process_group()


def test_sampler_under_gloo_world_2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_sampler_worker, args=(r, 2, port, 57, 4, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    a, b = got
    assert len(a) == len(b) == (57 // 2 // 4) * 4 and not set(a) & set(b)


def test_device_store_needs_the_gpu(shard_dir):
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from vla_touch_b200 import native as nv
    ds = cd.ControllerDataset(shard_dir, **CASES["h8"])
    with pytest.raises(nv.NativeError):
        ds.device_store()


def test_dataset_reads_hdf5_episodes_through_h5py(tmp_path, monkeypatch):
    """The `.h5` route (open_episode -> h5py.File) with the stand-in h5py the fixture generator uses for the reference (h5py is not in
    this image): same items, index mapping and statistics as the reference got from the very same `File` objects."""
    import types
    from oracle import gen_golden_dataset as gen
    gen.make_store()
    fake = types.ModuleType("h5py")
    fake.File = gen._File
    monkeypatch.setitem(sys.modules, "h5py", fake)
    for name in gen._STORE:
        open(tmp_path / name, "wb").close()
    ds = cd.ControllerDataset(str(tmp_path), use_images=True, image_size=IMAGE, **CASES["h16s3"])
    assert stems(ds.file_paths) == stems(GOLD["h16s3.files"]) and all(p.endswith(".h5") for p in ds.file_paths)
    assert np.array_equal(np.array(ds.episode_indices, dtype=np.int64), GOLD["h16s3.episode_indices"])
    for k, v in ds.stats.items():
        assert np.array_equal(v, GOLD[f"h16s3.stats.{k}"]), k
    for i in GOLD["h16s3.picks"]:
        for k, v in ds[int(i)].items():
            g = GOLD[f"h16s3.item{i}.{k}"]
            assert np.array_equal((v[:, ::9, ::9] if k.startswith("images") else v).numpy(), g), (i, k)
    # and the converter: .h5 -> .vtep shards that read back identically
    out = es.convert_directory(str(tmp_path), str(tmp_path / "shards"))
    assert len(out) == len(gen._STORE)
    ds2 = cd.ControllerDataset(str(tmp_path / "shards"), use_images=False, **CASES["h16s3"])
    assert ds2.episode_indices == ds.episode_indices
    assert torch.equal(ds2[3]["states"], ds[3]["states"]) and torch.equal(ds2[3]["vla_actions"], ds[3]["vla_actions"])
