"""mirror of cc/train/minibatch.py:14-88 (dataset tuples, MiniBatchState, make_dataloader).
The minibatch state machine restates tree_utils.tree_dataloader: one (re)shuffled permutation per
epoch, `n_minibatches` equally sized minibatches gathered along `axis`; the permutation stream is torch's, not
JAX threefry's, so minibatch ORDER differs from the reference for the same key."""
from __future__ import annotations

from typing import Callable, NamedTuple, Optional

import numpy as np
import torch

from ..utils import to_torch, tree_map, tree_shape


class MiniBatchState(NamedTuple):
    indices: torch.Tensor
    i: int
    bs: int
    n_minibatches: int
    minibatch_size: int
    key: int


class SupervisedDataset(NamedTuple):
    inputs: object
    targets: object


class SupervisedDatasetWithWeights(NamedTuple):
    inputs: object
    targets: object
    weights: object


class UnsupervisedDataset(NamedTuple):
    refss: object


class Dataloader(NamedTuple):
    minibatch_state: MiniBatchState
    update_fn: Callable


def _key_to_seed(key) -> int:
    if isinstance(key, (int, np.integer)):
        return int(key)
    arr = np.asarray(key).ravel()
    return int(arr[-1]) if arr.size else 1


def make_dataloader(dataset, key=1, n_minibatches: int = 1, axis: int = 0, reshuffle: bool = True,
                    tree_transform: Optional[Callable] = None, do_bootstrapping: bool = False, device="cuda") -> Dataloader:
    if not isinstance(dataset, (SupervisedDataset, SupervisedDatasetWithWeights, UnsupervisedDataset)):
        raise ValueError("`dataset` should be either a `SupervisedDataset`, `SupervisedDatasetWithWeights` or `UnsupervisedDataset`.")
    if axis != 0:
        raise NotImplementedError("minibatching along axis 0 only")
    bs = tree_shape(dataset)
    if isinstance(dataset, SupervisedDataset):
        dataset = SupervisedDatasetWithWeights(dataset.inputs, dataset.targets, np.ones((bs,), np.float32))
    dataset = to_torch(dataset, device)
    if bs % n_minibatches != 0:
        raise ValueError(f"batch size {bs} is not divisible by n_minibatches {n_minibatches}")
    mbs = bs // n_minibatches

    def permutation(seed):
        g = torch.Generator().manual_seed(seed)
        if do_bootstrapping:
            return torch.randint(0, bs, (bs,), generator=g)
        return torch.randperm(bs, generator=g) if reshuffle or seed == _key_to_seed(key) else torch.arange(bs)

    seed0 = _key_to_seed(key)
    state0 = MiniBatchState(permutation(seed0), 0, bs, n_minibatches, mbs, seed0)

    def update_fn(state: MiniBatchState):
        indices, i, seed = state.indices, state.i, state.key
        if i >= n_minibatches:  # new epoch
            seed = seed + 1
            indices = permutation(seed) if (reshuffle or do_bootstrapping) else indices
            i = 0
        idx = indices[i * mbs:(i + 1) * mbs].to(device)
        batch = tree_map(lambda a: a.index_select(0, idx), dataset)
        if tree_transform is not None:
            batch = tree_transform(seed, batch, mbs)
        return MiniBatchState(indices, i + 1, bs, n_minibatches, mbs, seed), batch

    return Dataloader(state0, update_fn)
