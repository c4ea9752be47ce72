"""TEST INFRASTRUCTURE ONLY -- `flax.linen` restated over numpy, enough to EXECUTE /root/reference/vit_flax/vit.py and
simple_vit.py unmodified (oracle/flax_shim/README.md).  Follows flax 0.5.0 (the version the reference's README cites):

* `Module`: subclasses become dataclasses with `parent` / `name` appended as keyword fields.  A module constructed while
  some module's method is running gets THAT module as its parent; without an explicit name it is auto-named
  `ClassName_N` (N counts that class among the parent's children, and only inside an `@compact` method).  A module that
  already has a parent keeps it when it is passed to another module as a field (flax adopts only parent-less modules).
* `init(rngs, *args)` runs the module with an empty variable collection and returns `{'params': tree}`;
  `apply(variables, *args, rngs=None)` runs it reading `variables['params']`.  `param(name, init_fn, *init_args)` calls
  `init_fn(rng, *init_args)` when initialising and checks the stored shape when applying.
* Layers: Dense, LayerNorm, Dropout, Sequential; functions gelu, softmax; initializers zeros / ones / lecun_normal.
"""
from __future__ import annotations

import dataclasses
import types
from typing import Any, Callable, Optional, Sequence

import numpy as np

import jax
from jax import numpy as jnp
from jax.numpy import wrap

__all__ = ["Module", "compact", "Dense", "LayerNorm", "Dropout", "Sequential", "gelu", "softmax", "initializers"]


class _Unspecified:
    def __repr__(self):
        return "<unspecified parent>"


_UNSPECIFIED = _Unspecified()
_module_stack: list = []       # modules whose methods are running, innermost last (flax: _context.module_stack)


class ScopeParamNotFoundError(KeyError):
    pass


class ScopeParamShapeError(ValueError):
    pass


class InvalidRngError(ValueError):
    pass


class NameInUseError(ValueError):
    pass


class _Run:
    """State of one top-level init / apply: the variable tree, the rngs and their per-name draw counters."""

    def __init__(self, params, rngs, initialising):
        self.params = params
        self.rngs = dict(rngs or {})
        self.initialising = initialising
        self.rng_counters = {}

    def subtree(self, path, create):
        node = self.params
        for name in path:
            if name not in node:
                if not create:
                    return None
                node[name] = {}
            node = node[name]
        return node


def _wrap_method(fn, is_compact):
    def wrapped(self, *args, **kwargs):
        if self._run is None:
            raise RuntimeError(f"{type(self).__name__}: modules run only inside .init() / .apply() (unbound module)")
        state = self._state
        if is_compact and state["compact_running"]:
            raise RuntimeError("nested call of the module's own compact method")
        _module_stack.append(self)
        if is_compact:
            state["compact_running"] = True
        try:
            return fn(self, *args, **kwargs)
        finally:
            _module_stack.pop()
            if is_compact:                    # flax: _state.reset() + scope.rewound() after a compact method
                state["compact_running"] = False
                state["autonames"] = {}
                state["children"] = set()
    wrapped.__name__ = fn.__name__
    wrapped.__doc__ = fn.__doc__
    wrapped._is_compact = is_compact
    wrapped._wrapped = True
    return wrapped


def compact(fn):
    fn._compact_marker = True
    return fn


class Module:
    def __init_subclass__(cls, **kwargs):
        super().__init_subclass__(**kwargs)
        ann = dict(cls.__dict__.get("__annotations__", {}))
        ann.pop("parent", None)
        ann.pop("name", None)
        ann["parent"] = Any                    # appended last, keyword defaults (flax: kw_only fields)
        ann["name"] = Optional[str]
        cls.__annotations__ = ann
        cls.parent = _UNSPECIFIED
        cls.name = None
        for attr, val in list(cls.__dict__.items()):
            if attr in ann:                    # a field default that happens to be a function (kernel_init, ...)
                continue
            if isinstance(val, types.FunctionType) and (attr == "__call__" or not attr.startswith("_")):
                setattr(cls, attr, _wrap_method(val, getattr(val, "_compact_marker", False)))
        dataclasses.dataclass(cls, eq=False, repr=False)

    # ---- construction: parent / name resolution (flax Module.__post_init__) ----
    def __post_init__(self):
        object.__setattr__(self, "_state", {"compact_running": False, "autonames": {}, "children": set()})
        object.__setattr__(self, "_run", None)
        object.__setattr__(self, "_path", ())
        if self.parent is _UNSPECIFIED:
            object.__setattr__(self, "parent", _module_stack[-1] if _module_stack else None)
        for f in dataclasses.fields(self):
            if f.name in ("parent", "name"):
                continue
            for sub in _modules_in(getattr(self, f.name)):
                if sub.parent is None:
                    raise NotImplementedError("flax_shim: adoption of a module constructed outside any module scope is "
                                              "not implemented (neither vit.py nor simple_vit.py does that)")
        p = self.parent
        if isinstance(p, Module):
            name = self.name
            if name is None:
                if not p._state["compact_running"]:
                    raise ValueError(f"{type(self).__name__} constructed without a name outside a compact method")
                prefix = type(self).__name__
                idx = p._state["autonames"].get(prefix, 0)
                p._state["autonames"][prefix] = idx + 1
                name = f"{prefix}_{idx}"
                object.__setattr__(self, "name", name)
            if name in p._state["children"]:
                raise NameInUseError(f"submodule name {name!r} already used in {type(p).__name__}")
            p._state["children"].add(name)
            object.__setattr__(self, "_path", p._path + (name,))
            object.__setattr__(self, "_run", p._run)
        elif p is not None:
            raise TypeError("parent must be a Module or None")

    def __repr__(self):
        return f"{type(self).__name__}(name={self.name!r})"

    # ---- variables and rngs ----
    def param(self, name, init_fn, *init_args):
        run = self._run
        if run.initialising:
            node = run.subtree(self._path, create=True)
            if name in node:
                value = node[name]
            else:
                value = np.asarray(init_fn(self.make_rng("params"), *init_args), dtype=np.float64)
                node[name] = value
            return wrap(np.asarray(value, dtype=np.float64))
        node = run.subtree(self._path, create=False)
        if node is None or name not in node:
            raise ScopeParamNotFoundError(f'no parameter named "{name}" in /{"/".join(self._path)}')
        value = np.asarray(node[name], dtype=np.float64)
        want = np.asarray(init_fn(jax.random.PRNGKey(0), *init_args)).shape      # flax: jax.eval_shape(init_fn, ...)
        if want != value.shape:
            raise ScopeParamShapeError(f'parameter "{name}" in /{"/".join(self._path)}: expected shape {want}, '
                                       f"got {value.shape}")
        return wrap(value)

    def make_rng(self, name):
        run = self._run
        if name not in run.rngs:
            raise InvalidRngError(f'{type(self).__name__} needs PRNG for "{name}"')
        count = run.rng_counters.get(name, 0)
        run.rng_counters[name] = count + 1
        key = jax.random.fold_in(jax.random.fold_in(run.rngs[name], "/".join(self._path)), count)
        key.draw_index = count               # the how-many-th draw of this stream in this run (golden script: the site)
        key.base = run.rngs[name]
        return key

    # ---- top level ----
    def _bound_clone(self, run):
        if isinstance(self.parent, Module):
            raise RuntimeError("init / apply are called on a top-level module")
        values = {f.name: getattr(self, f.name) for f in dataclasses.fields(self) if f.name not in ("parent", "name")}
        root = type(self)(**values, parent=None, name=self.name)
        object.__setattr__(root, "_run", run)
        return root

    def init(self, rngs, *args, **kwargs):
        if not isinstance(rngs, dict):
            rngs = {"params": rngs}
        run = _Run({}, rngs, initialising=True)
        self._bound_clone(run)(*args, **kwargs)
        return {"params": run.params}

    def init_with_output(self, rngs, *args, **kwargs):
        if not isinstance(rngs, dict):
            rngs = {"params": rngs}
        run = _Run({}, rngs, initialising=True)
        out = self._bound_clone(run)(*args, **kwargs)
        return out, {"params": run.params}

    def apply(self, variables, *args, rngs=None, **kwargs):
        params = variables["params"] if "params" in variables else {}
        if rngs is not None and not isinstance(rngs, dict):
            rngs = {"params": rngs}
        run = _Run(params, rngs, initialising=False)
        return self._bound_clone(run)(*args, **kwargs)


def _modules_in(value):
    if isinstance(value, Module):
        yield value
    elif isinstance(value, (list, tuple)):
        for v in value:
            yield from _modules_in(v)
    elif isinstance(value, dict):
        for v in value.values():
            yield from _modules_in(v)


# ---------------------------------------------------------------- initializers
class _Initializers:
    @staticmethod
    def zeros(key, shape, dtype=np.float64):
        return np.zeros(tuple(shape), dtype=np.float64)

    @staticmethod
    def ones(key, shape, dtype=np.float64):
        return np.ones(tuple(shape), dtype=np.float64)

    @staticmethod
    def lecun_normal():
        def init(key, shape, dtype=np.float64):
            fan_in = shape[-2] if len(shape) > 1 else shape[0]
            std = np.sqrt(1.0 / fan_in) / 0.87962566103423978       # variance_scaling(1, 'fan_in', 'truncated_normal')
            return np.asarray(jax.random.truncated_normal(key, -2.0, 2.0, tuple(shape))) * std
        return init


initializers = _Initializers()
_default_kernel_init = initializers.lecun_normal()


# ---------------------------------------------------------------- layers
class Dense(Module):
    features: int
    use_bias: bool = True
    dtype: Any = None
    param_dtype: Any = np.float32
    precision: Any = None
    kernel_init: Callable = _default_kernel_init
    bias_init: Callable = initializers.zeros

    @compact
    def __call__(self, inputs):
        inputs = jnp.asarray(inputs)
        kernel = self.param("kernel", self.kernel_init, (inputs.shape[-1], self.features))
        y = wrap(np.matmul(inputs, kernel))                         # lax.dot_general over the last axis
        if self.use_bias:
            bias = self.param("bias", self.bias_init, (self.features,))
            y = wrap(y + bias)
        return y


class LayerNorm(Module):
    epsilon: float = 1e-6
    dtype: Any = None
    param_dtype: Any = np.float32
    use_bias: bool = True
    use_scale: bool = True
    bias_init: Callable = initializers.zeros
    scale_init: Callable = initializers.ones

    @compact
    def __call__(self, x):
        x = jnp.asarray(x)
        mean = np.mean(x, axis=-1, keepdims=True)
        mean2 = np.mean(np.square(x), axis=-1, keepdims=True)
        var = np.maximum(0.0, mean2 - np.square(mean))              # flax _compute_stats
        mul = 1.0 / np.sqrt(var + self.epsilon)                     # lax.rsqrt(var + epsilon)
        features = x.shape[-1]
        if self.use_scale:
            mul = mul * self.param("scale", self.scale_init, (features,))
        y = (x - mean) * mul
        if self.use_bias:
            y = y + self.param("bias", self.bias_init, (features,))
        return wrap(np.asarray(y))


class Dropout(Module):
    rate: float
    broadcast_dims: Sequence[int] = ()
    deterministic: Optional[bool] = None

    @compact
    def __call__(self, inputs, deterministic=None):
        if self.deterministic is not None and deterministic is not None:
            raise ValueError("deterministic given both as attribute and as argument")
        deterministic = self.deterministic if deterministic is None else deterministic
        if deterministic is None:
            raise ValueError("deterministic must be given")
        if self.rate == 0.0:
            return inputs
        if self.rate == 1.0:
            return wrap(np.zeros_like(inputs))
        keep_prob = 1.0 - self.rate
        if deterministic:
            return inputs
        rng = self.make_rng("dropout")
        shape = list(inputs.shape)
        for d in self.broadcast_dims:
            shape[d] = 1
        mask = np.broadcast_to(np.asarray(jax.random.bernoulli(rng, keep_prob, tuple(shape))), inputs.shape)
        return wrap(np.where(mask, np.asarray(inputs) / keep_prob, 0.0))


class Sequential(Module):
    layers: Sequence[Callable]

    def __call__(self, *args, **kwargs):
        if not self.layers:
            raise ValueError(f"Empty Sequential module {self.name}.")
        outputs = self.layers[0](*args, **kwargs)
        for layer in self.layers[1:]:
            outputs = layer(outputs)
        return outputs


# ---------------------------------------------------------------- functions
def gelu(x, approximate=True):
    x = np.asarray(x)
    if approximate:
        c = np.sqrt(2.0 / np.pi)
        return wrap(0.5 * x * (1.0 + np.tanh(c * (x + 0.044715 * (x ** 3)))))
    from math import erf
    return wrap(0.5 * x * (1.0 + np.vectorize(erf)(x / np.sqrt(2.0))))


def softmax(x, axis=-1):
    x = np.asarray(x)
    unnormalized = np.exp(x - x.max(axis=axis, keepdims=True))
    return wrap(unnormalized / unnormalized.sum(axis=axis, keepdims=True))
