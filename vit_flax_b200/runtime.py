"""Engine cache behind ``ViT.apply``: one ``Engine`` per (config, precision, device),
weights re-packed only when a different params object is passed."""
from __future__ import annotations

from typing import Dict, Tuple

from .engine import Engine
from .params import flatten_params

_engines: Dict[Tuple, Engine] = {}
_loaded: Dict[Tuple, Tuple] = {}


def _fingerprint(variables) -> Tuple:
    """Identity of the params object: ids of the container and of every leaf.  In-place
    mutation of a leaf is NOT detected -- pass ``reload=True`` to ``apply`` after one."""
    flat = flatten_params(variables)
    return (id(variables),) + tuple((k, id(v)) for k, v in sorted(flat.items()))


def get_engine(vit, channels: int, precision: str, device: int, batch: int, variables,
               reload: bool = False) -> Engine:
    key = (vit.image_size, vit.patch_size, vit.num_classes, vit.dim, vit.depth, vit.heads,
           vit.mlp_dim, vit.pool, channels, precision, device, float(vit.dropout), float(vit.emb_dropout))
    eng = _engines.get(key)
    if eng is None or eng.max_batch < batch:
        if eng is not None:
            eng.close()
            _loaded.pop(key, None)
        eng = Engine(image_size=vit.image_size, patch_size=vit.patch_size,
                     num_classes=vit.num_classes, dim=vit.dim, depth=vit.depth, heads=vit.heads,
                     mlp_dim=vit.mlp_dim, pool=vit.pool, channels=channels, precision=precision,
                     max_batch=batch, device=device, dropout=vit.dropout, emb_dropout=vit.emb_dropout)
        _engines[key] = eng
    fp = _fingerprint(variables)
    if reload or _loaded.get(key) != fp:
        eng.load_params(variables)
        _loaded[key] = fp
    return eng


def clear_cache() -> None:
    for e in _engines.values():
        e.close()
    _engines.clear()
    _loaded.clear()
