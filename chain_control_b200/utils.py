"""Metrics and pytree helpers mirroring cc/utils/utils.py (mse :53-55, rmse :66-67, mae :48-50,
weighted_mse :58-63, l1_norm/l2_norm :38-45) and the few tree_utils functions the hot path's callers use
(batch_concat, tree_batch).  Trees are (nested) dicts / lists / tuples of torch tensors."""
from __future__ import annotations

from collections import OrderedDict

import numpy as np
import torch


def tree_leaves(tree):
    if isinstance(tree, torch.Tensor) or isinstance(tree, np.ndarray) or np.isscalar(tree):
        return [tree]
    if isinstance(tree, dict):
        return [l for k in tree for l in tree_leaves(tree[k])]
    if isinstance(tree, (list, tuple)):
        return [l for t in tree for l in tree_leaves(t)]
    if tree is None:
        return []
    raise TypeError(f"unsupported pytree node {type(tree)}")


def tree_map(fn, tree):
    if isinstance(tree, dict):
        return type(tree)((k, tree_map(fn, v)) for k, v in tree.items())
    if isinstance(tree, tuple) and hasattr(tree, "_fields"):
        return type(tree)(*[tree_map(fn, t) for t in tree])
    if isinstance(tree, (list, tuple)):
        return type(tree)(tree_map(fn, t) for t in tree)
    if tree is None:
        return None
    return fn(tree)


def batch_concat(tree, num_batch_dims: int = 1):
    """tree_utils.batch_concat: ravel every leaf beyond its first `num_batch_dims` axes, concatenate (pytree order)."""
    leaves = [torch.as_tensor(l) for l in tree_leaves(tree)]
    flat = [l.reshape(tuple(l.shape[:num_batch_dims]) + (-1,)) for l in leaves]
    return flat[0] if len(flat) == 1 else torch.cat(flat, dim=-1)


def tree_shape(tree, axis: int = 0) -> int:
    return int(tree_leaves(tree)[0].shape[axis])


def to_torch(tree, device=None):
    return tree_map(lambda a: torch.as_tensor(np.asarray(a) if not isinstance(a, torch.Tensor) else a, dtype=torch.float32).to(device or "cuda"), tree)


def to_numpy(tree):
    return tree_map(lambda a: a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a), tree)


def l2_norm(vector):
    assert vector.ndim == 1
    return torch.sqrt(torch.sum(vector ** 2))


def l1_norm(vector):
    assert vector.ndim == 1
    return torch.sum(torch.abs(vector))


def mae(y, yhat):
    y, yhat = batch_concat(y, 0), batch_concat(yhat, 0)
    return torch.mean(torch.abs(y - yhat))


def mse(y, yhat):
    y, yhat = batch_concat(y, 0), batch_concat(yhat, 0)
    return torch.mean((y - yhat) ** 2)


def weighted_mse(y, yhat, weights):
    y, yhat = batch_concat(y, 1), batch_concat(yhat, 1)  # keep the batch axis
    se = (y - yhat) ** 2
    sse = torch.sum(weights.reshape(-1, 1) * se, dim=0)
    return torch.mean(sse)


def rmse(y, yhat):
    return torch.sqrt(mse(y, yhat))


def make_postprocess_fn(output_struct):
    """flat [..., O] -> the observation pytree (utils.py:29-35 uses ravel_pytree's inverse).
    output_struct: None (flat) or OrderedDict name -> size."""
    if output_struct is None:
        return lambda flat: flat

    def unravel(flat):
        out, off = OrderedDict(), 0
        for name, size in output_struct.items():
            out[name] = flat[..., off:off + size]
            off += size
        return out

    return unravel


def struct_of_specs(specs):
    """(total size, structure) of an int / shape tuple / object with .shape / (Ordered)dict of those."""
    def size(s):
        if isinstance(s, (int, np.integer)):
            return int(s)
        shape = getattr(s, "shape", s)
        return int(np.prod(shape)) if len(tuple(shape)) else 1

    if isinstance(specs, dict):
        st = OrderedDict((k, size(v)) for k, v in specs.items())
        return sum(st.values()), st
    return size(specs), None
