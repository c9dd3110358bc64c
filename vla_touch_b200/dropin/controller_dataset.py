"""Shadows the reference's controller_dataset (controller_dataset.py:17-476): the dataset classes and the two normalisation
functions are the B200 package's (same names, arguments, items, statistics; `.vtep` shards, or `.h5` where h5py is installed)."""
from vla_touch_b200.controller_dataset import (ControllerDataModule, ControllerDataset, EpisodeBatchSampler,  # noqa: F401
                                               denormalize_actions, natural_sort_filenames, normalize_actions)
