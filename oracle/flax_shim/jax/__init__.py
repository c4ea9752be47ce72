"""TEST INFRASTRUCTURE ONLY -- numpy stand-in for the slice of `jax` that vit_flax/vit.py and simple_vit.py use
(oracle/flax_shim/README.md).  Arrays are float64 numpy arrays (subclass `Array`, which adds the one jax.Array method
the reference calls: `.split`)."""
from __future__ import annotations

from . import numpy, random  # noqa: F401  (jax.numpy, jax.random)
from .numpy import Array  # noqa: F401


def tree_leaves(tree):
    """Leaves of a nested dict / list / tuple pytree, in key-sorted order like jax."""
    if isinstance(tree, dict):
        return [leaf for k in sorted(tree) for leaf in tree_leaves(tree[k])]
    if isinstance(tree, (list, tuple)):
        return [leaf for t in tree for leaf in tree_leaves(t)]
    return [] if tree is None else [tree]


def tree_map(f, tree, *rest):
    if isinstance(tree, dict):
        return {k: tree_map(f, tree[k], *(r[k] for r in rest)) for k in tree}
    if isinstance(tree, (list, tuple)):
        return type(tree)(tree_map(f, t, *(r[i] for r in rest)) for i, t in enumerate(tree))
    return None if tree is None else f(tree, *rest)


class _TreeUtil:
    tree_leaves = staticmethod(tree_leaves)
    tree_map = staticmethod(tree_map)


tree_util = _TreeUtil()
