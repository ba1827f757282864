"""NumPy stand-ins for the third-party packages the reference imports, so that the reference's OWN source files
(/root/reference/cc/...) execute unmodified in this container and produce golden vectors for the hot path.

TEST INFRASTRUCTURE ONLY (used by tests/golden/make_golden.py, run once here; the fixtures it writes are committed).
Nothing in chain_control_b200/ imports this file.

jax / jaxlib / equinox==0.9.0 / optax / imt-tree-utils / dm_env / dm_control are not installable here (no network,
absent from /opt/wheelhouse).  The reference's hot path uses a small, well-defined part of their public API
(inventory: `grep -o "(jnp|jax|eqx|jtu|jrand)\\.[a-z_.]*"` over cc/core, cc/examples, cc/train/step_fn.py):

  jax.numpy   array asarray zeros mean sum abs sqrt squeeze moveaxis arctan tanh ... -> the numpy function of the same name
  jax.lax     scan                        -> Python loop, outputs stacked leaf-wise
  jax.random  PRNGKey split normal uniform -> numpy Generator streams keyed by the key's two words.  NOT threefry:
                                             every normal() draw is recorded (``NOISE_LOG``) so a fixture carries the
                                             noise the reference added as an explicit array
  jax.tree_util / equinox partition combine filter tree_at is_array Module nn.{Linear,Lambda,Sequential,Dropout}
                                          -> a minimal pytree registry with jax's flattening order (dict keys sorted,
                                             OrderedDict in insertion order, None = no leaves, Module = dataclass fields
                                             in declaration order, static fields in the treedef)
  equinox.filter_vmap                     -> Python loop over axis 0 of every array argument, outputs stacked
  equinox.filter_value_and_grad           -> COMPLEX-STEP differentiation (Im f(p + i h e_k) / h, h = 1e-30): exact to
                                             fp64 round-off for the analytic/piecewise-linear operations of the path,
                                             one evaluation of the reference's loss per parameter
  tree_utils  batch_concat tree_batch add_batch_dim tree_map_flat tree_shape tree_ndim -> restated from the package's
                                             documented behaviour (imt-tree-utils, setup.py:47 of the reference)
  optax       adam clip_by_global_norm chain apply (via eqx.apply_updates) -> restated from the published algorithms

Arithmetic runs in float64 / complex128 (the reference under jax would run float32): the fixtures are the
reference's algorithm in double precision, the same yardstick BASELINE.json's tolerances are written against.
``ravel_pytree``'s unflatten therefore does not cast back to the toy output's dtype.
"""
from __future__ import annotations

import abc
import dataclasses
import sys
import types
from collections import OrderedDict

import numpy as np

# --------------------------------------------------------------------------------------------- pytrees


class _Static:
    pass


def _is_namedtuple(x):
    return isinstance(x, tuple) and hasattr(x, "_fields")


def _module_fields(cls):
    dyn, stat = [], []
    for f in dataclasses.fields(cls):
        (stat if f.metadata.get("static", False) else dyn).append(f.name)
    return dyn, stat


def _children(node):
    """-> (children, rebuild) or None for a leaf.  Order = jax's flattening order."""
    if node is None:
        return [], lambda ch: None
    if isinstance(node, Module):
        dyn, stat = _module_fields(type(node))
        statv = {n: getattr(node, n) for n in stat}
        cls = type(node)

        def rebuild(ch, cls=cls, dyn=dyn, statv=statv):
            obj = object.__new__(cls)
            for n, v in zip(dyn, ch):
                object.__setattr__(obj, n, v)
            for n, v in statv.items():
                object.__setattr__(obj, n, v)
            return obj

        return [getattr(node, n) for n in dyn], rebuild
    if isinstance(node, OrderedDict):
        keys = list(node.keys())
        return [node[k] for k in keys], lambda ch, keys=keys: OrderedDict(zip(keys, ch))
    if isinstance(node, dict):
        keys = sorted(node.keys())
        return [node[k] for k in keys], lambda ch, keys=keys: dict(zip(keys, ch))
    if _is_namedtuple(node):
        cls = type(node)
        return list(node), lambda ch, cls=cls: cls(*ch)
    if isinstance(node, tuple):
        return list(node), lambda ch: tuple(ch)
    if isinstance(node, list):
        return list(node), lambda ch: list(ch)
    return None


def tree_map(f, tree, *rest, is_leaf=None):
    if is_leaf is not None and is_leaf(tree):
        return f(tree, *rest)
    c = _children(tree)
    if c is None:
        return f(tree, *rest)
    ch, rebuild = c
    rest_ch = []
    for r in rest:
        rc = _children(r)
        assert rc is not None and len(rc[0]) == len(ch), "tree_map: structures differ"
        rest_ch.append(rc[0])
    return rebuild([tree_map(f, x, *[rc[i] for rc in rest_ch], is_leaf=is_leaf) for i, x in enumerate(ch)])


def tree_leaves(tree, is_leaf=None):
    out = []
    tree_map(lambda x: out.append(x), tree, is_leaf=is_leaf)
    return out


class _TreeDef:
    def __init__(self, tree):
        self.skel = tree_map(lambda x: _Static, tree)
        self.num_leaves = len(tree_leaves(tree))

    def unflatten(self, leaves):
        it = iter(leaves)
        return tree_map(lambda _: next(it), self.skel)


def tree_flatten(tree, is_leaf=None):
    return tree_leaves(tree, is_leaf), _TreeDef(tree)


def tree_unflatten(treedef, leaves):
    return treedef.unflatten(leaves)


def _map_prefix(fn, spec, tree):
    """spec is a prefix of tree: a spec leaf (bool / callable) applies to every leaf below it."""
    c = _children(spec)
    if c is None:
        return tree_map(lambda leaf: fn(spec, leaf), tree)
    sch, _ = c
    tc = _children(tree)
    assert tc is not None and len(tc[0]) == len(sch), "filter_spec is not a prefix of the tree"
    return tc[1]([_map_prefix(fn, s, t) for s, t in zip(sch, tc[0])])


# --------------------------------------------------------------------------------------------- equinox


def is_array(x):
    return isinstance(x, (np.ndarray, np.generic))


def static_field(**kw):
    return dataclasses.field(metadata={"static": True}, **kw)


class _ModuleMeta(abc.ABCMeta):
    def __new__(mcs, name, bases, ns, **kw):
        cls = super().__new__(mcs, name, bases, ns, **kw)
        return dataclasses.dataclass(eq=False, repr=False)(cls)


class Module(metaclass=_ModuleMeta):
    """equinox.Module: a dataclass that is a pytree of its fields."""


def _pick(spec, leaf):
    return bool(spec(leaf)) if callable(spec) else bool(spec)


def partition(tree, filter_spec, replace=None):
    a = _map_prefix(lambda s, x: x if _pick(s, x) else replace, filter_spec, tree)
    b = _map_prefix(lambda s, x: replace if _pick(s, x) else x, filter_spec, tree)
    return a, b


def filter(tree, filter_spec, inverse=False, replace=None):  # noqa: A001
    a, b = partition(tree, filter_spec, replace)
    return b if inverse else a


def combine(*trees):
    def pick(*xs):
        for x in xs:
            if x is not None:
                return x
        return None

    return tree_map(pick, *trees, is_leaf=lambda x: x is None)


class _Box:
    def __init__(self, v):
        self.v = v


def tree_at(where, pytree, replace=None, replace_fn=None):
    boxed = tree_map(_Box, pytree)
    nodes = where(boxed)
    single = not isinstance(nodes, (tuple, list)) or isinstance(nodes, Module)
    if single:
        nodes, replace = (nodes,), (replace,)
    ids = {id(n): i for i, n in enumerate(nodes)}

    def unbox(t):
        return tree_map(lambda b: b.v, t)

    def walk(node):
        if id(node) in ids:
            r = replace[ids[id(node)]] if replace_fn is None else replace_fn(unbox(node))
            return r
        if isinstance(node, _Box):
            return node.v
        ch, rebuild = _children(node)
        return rebuild([walk(c) for c in ch])

    return walk(boxed)


def filter_jit(fun=None, **kw):
    return fun if fun is not None else (lambda f: f)


def filter_vmap(fun=None, **kw):
    if fun is None:
        return lambda f: filter_vmap(f, **kw)

    def vmapped(*args):
        arrs = [x for x in tree_leaves(args) if is_array(x) and np.ndim(x) > 0]
        n = arrs[0].shape[0]
        assert all(a.shape[0] == n for a in arrs)
        outs = []
        for b in range(n):
            sl = tree_map(lambda x: x[b] if (is_array(x) and np.ndim(x) > 0) else x, args)
            outs.append(fun(*sl))
        return tree_map(lambda *xs: np.stack(xs), *outs)

    return vmapped


CSTEP = 1e-30


def filter_value_and_grad(fun=None, *, arg=is_array, has_aux=False, **kw):
    """Complex-step derivative with respect to the leaves of the first argument that ``arg`` selects."""
    if fun is None:
        return lambda f: filter_value_and_grad(f, arg=arg, has_aux=has_aux)

    def vg(x, *args, **kwargs):
        diff, nondiff = partition(x, arg)
        leaves, treedef = tree_flatten(diff)
        out = fun(x, *args, **kwargs)
        grads = []
        for li, leaf in enumerate(leaves):
            g = np.zeros(np.shape(leaf), dtype=np.float64)
            flat = np.asarray(leaf, dtype=np.complex128).reshape(-1)
            for k in range(flat.size):
                pert = flat.copy()
                pert[k] += 1j * CSTEP
                new_leaves = list(leaves)
                new_leaves[li] = pert.reshape(np.shape(leaf))
                xk = combine(tree_unflatten(treedef, new_leaves), nondiff)
                ok = fun(xk, *args, **kwargs)
                val = ok[0] if has_aux else ok
                g.reshape(-1)[k] = np.imag(val) / CSTEP
            grads.append(g)
        return out, tree_unflatten(treedef, grads)

    return vg


def apply_updates(model, updates):
    return tree_map(lambda m, u: m if u is None else m + u, model, updates, is_leaf=lambda x: x is None)


def tree_inference(tree, value):
    return tree


class _Linear(Module):
    weight: np.ndarray
    bias: np.ndarray
    in_features: int = static_field()
    out_features: int = static_field()
    use_bias: bool = static_field()

    def __init__(self, in_features, out_features, use_bias=True, *, key):
        lim = 1.0 / np.sqrt(in_features)
        wk, bk = _split(key, 2)
        self.weight = _uniform(wk, (out_features, in_features), minval=-lim, maxval=lim)
        self.bias = _uniform(bk, (out_features,), minval=-lim, maxval=lim) if use_bias else None
        self.in_features, self.out_features, self.use_bias = in_features, out_features, use_bias

    def __call__(self, x, *, key=None):
        y = self.weight @ x
        return y + self.bias if self.bias is not None else y


class _Lambda(Module):
    fn: object

    def __call__(self, x, *, key=None):
        return self.fn(x)


class _Dropout(Module):
    p: float = 0.5

    def __call__(self, x, *, key=None):
        raise NotImplementedError("dropout is not implemented by the reference either (neural_ode_model.py:41-43)")


class _Sequential(Module):
    layers: tuple

    def __init__(self, layers):
        self.layers = tuple(layers)

    def __call__(self, x, *, key=None):
        for layer in self.layers:
            x = layer(x)
        return x

    def __getitem__(self, i):
        return self.layers[i]


# --------------------------------------------------------------------------------------------- jax.random

NOISE_LOG = []


def _PRNGKey(seed):
    return np.array([0, int(seed) & 0xFFFFFFFF], dtype=np.uint32)


def _gen(key):
    k = np.asarray(key, dtype=np.uint32).reshape(-1)
    return np.random.default_rng(np.random.SeedSequence([int(k[0]), int(k[1])]))


def _split(key, num=2):
    return _gen(key).integers(0, 2**32, size=(num, 2), dtype=np.uint32)


def _normal(key, shape=(), dtype=np.float64):
    # rounded to fp32-representable values so that an fp32 consumer of the recorded noise sees the same numbers
    v = np.float32(_gen(key).standard_normal(size=tuple(shape))).astype(np.float64)
    NOISE_LOG.append(np.array(v))
    return v


def _uniform(key, shape=(), dtype=np.float64, minval=0.0, maxval=1.0):
    return _gen(key).uniform(size=tuple(shape)) * (np.asarray(maxval) - np.asarray(minval)) + np.asarray(minval)


# --------------------------------------------------------------------------------------------- jax.numpy etc.


def _relu(x):
    x = np.asarray(x)
    return np.where(x.real > 0, x, 0 * x)


def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def _abs(x):
    x = np.asarray(x)
    return np.where(x.real < 0, -x, x) if np.iscomplexobj(x) else np.abs(x)


def _scan(f, init, xs, length=None):
    n = length if xs is None else tree_leaves(xs)[0].shape[0]
    carry, ys = init, []
    for i in range(n):
        x = None if xs is None else tree_map(lambda a: a[i], xs)
        carry, y = f(carry, x)
        ys.append(y)
    return carry, tree_map(lambda *v: np.stack(v), *ys)


def ravel_pytree(tree):
    leaves, treedef = tree_flatten(tree)
    shapes = [np.shape(x) for x in leaves]
    sizes = [int(np.prod(s)) for s in shapes]
    flat = np.concatenate([np.reshape(x, -1) for x in leaves]) if leaves else np.zeros((0,))

    def unravel(v):
        out, o = [], 0
        for s, n in zip(shapes, sizes):
            out.append(np.reshape(v[o:o + n], s))
            o += n
        return tree_unflatten(treedef, out)

    return flat, unravel


# --------------------------------------------------------------------------------------------- tree_utils


class _PyTreeMeta(type):
    def __getitem__(cls, item):
        return cls


class PyTree(metaclass=_PyTreeMeta):
    pass


def batch_concat(values, num_batch_dims=1):
    leaves = [np.asarray(x) for x in tree_leaves(values)]
    leaves = [np.reshape(x, x.shape[:num_batch_dims] + (-1,)) for x in leaves]
    return np.concatenate(leaves, axis=-1)


def tree_batch(trees, along_existing_first_axis=False, backend="numpy"):
    op = np.concatenate if along_existing_first_axis else np.stack
    return tree_map(lambda *xs: op([np.asarray(x) for x in xs], axis=0), *trees)


def add_batch_dim(tree):
    return tree_map(lambda x: np.asarray(x)[None], tree)


def tree_map_flat(tree, f, *args, **kwargs):
    flat, unravel = ravel_pytree(tree)
    return unravel(f(flat, *args, **kwargs))


def tree_shape(tree, axis=0):
    return np.shape(tree_leaves(tree)[0])[axis]


def tree_ndim(tree):
    return np.ndim(tree_leaves(tree)[0])


# --------------------------------------------------------------------------------------------- optax (restated)


class GradientTransformation:
    def __init__(self, init, update):
        self.init, self.update = init, update


def _leafmap(f, *trees):
    return tree_map(lambda *xs: None if xs[0] is None else f(*xs), *trees, is_leaf=lambda x: x is None)


def clip_by_global_norm(max_norm):
    def update(grads, state, params=None):
        gn = np.sqrt(sum(float(np.sum(np.square(g))) for g in tree_leaves(grads)))
        # optax.clip_by_global_norm: g * max_norm / max(norm, max_norm)
        return _leafmap(lambda g: g * (max_norm / max(gn, max_norm)), grads), state

    return GradientTransformation(lambda params: (), update)


def adam(learning_rate, b1=0.9, b2=0.999, eps=1e-8, eps_root=0.0):
    def init(params):
        z = _leafmap(lambda p: np.zeros_like(np.asarray(p, dtype=np.float64)), params)
        return {"count": 0, "mu": z, "nu": _leafmap(np.copy, z)}

    def update(grads, state, params=None):
        c = state["count"] + 1
        mu = _leafmap(lambda g, m: b1 * m + (1 - b1) * g, grads, state["mu"])
        nu = _leafmap(lambda g, v: b2 * v + (1 - b2) * g * g, grads, state["nu"])
        upd = _leafmap(lambda m, v: -learning_rate * (m / (1 - b1**c)) / (np.sqrt(v / (1 - b2**c) + eps_root) + eps),
                       mu, nu)
        return upd, {"count": c, "mu": mu, "nu": nu}

    return GradientTransformation(init, update)


def chain(*ts):
    def init(params):
        return tuple(t.init(params) for t in ts)

    def update(grads, state, params=None):
        new = []
        for t, s in zip(ts, state):
            grads, s = t.update(grads, s, params)
            new.append(s)
        return grads, tuple(new)

    return GradientTransformation(init, update)


# --------------------------------------------------------------------------------------------- dm_env.specs


class Array:
    def __init__(self, shape, dtype, name=None):
        self.shape, self.dtype, self.name = tuple(shape), np.dtype(dtype), name


class BoundedArray(Array):
    def __init__(self, shape, dtype, minimum, maximum, name=None):
        super().__init__(shape, dtype, name)
        self.minimum, self.maximum = minimum, maximum


# --------------------------------------------------------------------------------------------- install


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install():
    """Register the stand-ins in sys.modules (refuses to shadow a real jax)."""
    if "jax" in sys.modules and not getattr(sys.modules["jax"], "__refshim__", False):
        raise RuntimeError("a real jax is imported; the shim is only for images without it")
    jnp = _mod("jax.numpy", ndarray=np.ndarray, array=lambda x, dtype=None: np.array(x, dtype=dtype if dtype else None),
               asarray=np.asarray, zeros=np.zeros, ones=np.ones, mean=np.mean, sum=np.sum, abs=_abs, sqrt=np.sqrt,
               squeeze=np.squeeze, moveaxis=np.moveaxis, arctan=np.arctan, tanh=np.tanh, concatenate=np.concatenate,
               stack=np.stack, exp=np.exp, clip=np.clip, float32=np.float32, float64=np.float64, pi=np.pi,
               arange=np.arange, where=np.where, eye=np.eye, atleast_1d=np.atleast_1d, zeros_like=np.zeros_like)
    jrandom = _mod("jax.random", PRNGKey=_PRNGKey, split=_split, normal=_normal, uniform=_uniform)
    jtu = _mod("jax.tree_util", tree_map=tree_map, tree_flatten=tree_flatten, tree_unflatten=tree_unflatten,
               tree_leaves=tree_leaves)
    jnn = _mod("jax.nn", relu=_relu, sigmoid=_sigmoid, tanh=np.tanh)
    jlax = _mod("jax.lax", scan=_scan)
    jfu = _mod("jax.flatten_util", ravel_pytree=ravel_pytree)
    _mod("jax", __refshim__=True, numpy=jnp, random=jrandom, tree_util=jtu, nn=jnn, lax=jlax, flatten_util=jfu,
         Array=np.ndarray, tree_map=tree_map, vmap=filter_vmap, jit=filter_jit)
    enn = _mod("equinox.nn", Linear=_Linear, Lambda=_Lambda, Sequential=_Sequential, Dropout=_Dropout)
    _mod("equinox", Module=Module, nn=enn, is_array=is_array, partition=partition, combine=combine, filter=filter,
         tree_at=tree_at, filter_jit=filter_jit, filter_vmap=filter_vmap, filter_value_and_grad=filter_value_and_grad,
         apply_updates=apply_updates, static_field=static_field, tree_inference=tree_inference,
         tree_serialise_leaves=None, tree_deserialise_leaves=None)

    def _no(*a, **k):
        raise NotImplementedError("not on the hot path")

    _mod("tree_utils", PyTree=PyTree, batch_concat=batch_concat, tree_batch=tree_batch, add_batch_dim=add_batch_dim,
         tree_map_flat=tree_map_flat, tree_shape=tree_shape, tree_ndim=tree_ndim, tree_dataloader=_no,
         tree_insert_IMPURE=_no, tree_zeros_like=lambda t: tree_map(np.zeros_like, t), tree_slice=_no,
         tree_concat=_no, tree_indices=_no)
    _mod("optax", GradientTransformation=GradientTransformation, adam=adam, clip_by_global_norm=clip_by_global_norm,
         chain=chain, OptState=object)
    specs = _mod("dm_env.specs", Array=Array, BoundedArray=BoundedArray)
    _mod("dm_env", specs=specs)
    ctl = _mod("dm_control.rl.control", Environment=object)
    rl = _mod("dm_control.rl", control=ctl)
    _mod("dm_control", rl=rl)
