"""Checkpoint interop: the Flax msgpack layout, .npz and .safetensors."""
import msgpack
import numpy as np
import pytest

from vit_flax_b200 import flatten_params, init_params, load_params, perturb_params, save_params
from vit_flax_b200.checkpoint import msgpack_restore, msgpack_serialize
from _util import TINY


def _tree():
    return perturb_params(init_params(seed=3, **TINY), seed=4)


def _assert_same(a, b):
    fa, fb = flatten_params(a), flatten_params(b)
    assert sorted(fa) == sorted(fb)
    for k in fa:
        np.testing.assert_array_equal(np.asarray(fa[k], np.float32), fb[k])
        assert fb[k].dtype == np.float32 and fb[k].flags.c_contiguous


@pytest.mark.parametrize("ext", [".msgpack", ".npz", ".safetensors"])
def test_round_trip(tmp_path, ext):
    v = _tree()
    path = tmp_path / ("ckpt" + ext)
    save_params(v, path)
    _assert_same(v, load_params(path))


def test_reads_bytes_laid_out_like_flax_serialization():
    """Bytes built by hand from the published layout of flax.serialization (not with our writer):
    leaves are ExtType(1, packb((shape, dtype.name, C-order bytes))); bf16 leaves are widened."""
    import ml_dtypes
    k = np.arange(12, dtype=np.float32).reshape(3, 4)
    b = np.array([1.5, -2.0, 0.25, 8.0], dtype=ml_dtypes.bfloat16)

    def leaf(a):
        return msgpack.ExtType(1, msgpack.packb((a.shape, a.dtype.name, a.tobytes()), use_bin_type=True))

    blob = msgpack.packb({"params": {"Dense_0": {"kernel": leaf(k), "bias": leaf(b)}}}, use_bin_type=True)
    tree = load_params(blob)
    np.testing.assert_array_equal(tree["params"]["Dense_0"]["kernel"], k)
    np.testing.assert_array_equal(tree["params"]["Dense_0"]["bias"], b.astype(np.float32))


def test_chunked_arrays_are_reassembled():
    a = np.arange(10, dtype=np.float32).reshape(2, 5)
    node = {"__msgpack_chunked_array__": True, "shape": [2, 5],
            "chunks": {"0": a.ravel()[:4], "1": a.ravel()[4:8], "2": a.ravel()[8:]}}
    got = msgpack_restore(msgpack_serialize({"w": node["chunks"]["0"]}))   # plain leaf survives
    np.testing.assert_array_equal(got["w"], a.ravel()[:4])
    import vit_flax_b200.checkpoint as ck
    blob = msgpack.packb({"w": {"__msgpack_chunked_array__": True, "shape": [2, 5],
                                "chunks": {k: ck._ext_pack(v) for k, v in node["chunks"].items()}}},
                         use_bin_type=True)
    np.testing.assert_array_equal(load_params(blob)["w"], a)


def test_unknown_extension(tmp_path):
    with pytest.raises(ValueError, match="unknown checkpoint format"):
        load_params(tmp_path / "x.bin")


def test_chunked_arrays_use_the_flax_tuple_layout(monkeypatch):
    """Leaves above the chunk limit: flax writes ``shape`` and ``chunks`` through ``_tuple_to_dict``
    (``{'0': d0, '1': d1}``).  A hand-built chunk dict in that layout must load, our writer must emit it, and
    the older list form of ``shape`` stays readable."""
    from vit_flax_b200 import checkpoint
    a = np.arange(24, dtype=np.float32).reshape(4, 6)

    def leaf(x):
        return msgpack.ExtType(1, msgpack.packb((x.shape, x.dtype.name, x.tobytes()), use_bin_type=True))

    flat = a.ravel()
    flax_style = {"params": {"w": {"__msgpack_chunked_array__": True, "shape": {"0": 4, "1": 6},
                                   "chunks": {"0": leaf(flat[:10]), "1": leaf(flat[10:20]), "2": leaf(flat[20:])}}}}
    got = load_params(msgpack.packb(flax_style, use_bin_type=True))
    np.testing.assert_array_equal(got["params"]["w"], a)
    list_style = {"w": {"__msgpack_chunked_array__": True, "shape": [4, 6], "chunks": {"0": leaf(flat[:13]), "1": leaf(flat[13:])}}}
    np.testing.assert_array_equal(load_params(msgpack.packb(list_style, use_bin_type=True))["w"], a)
    monkeypatch.setattr(checkpoint, "_MAX_CHUNK", 40)            # 10 floats per chunk
    raw = msgpack.unpackb(msgpack_serialize({"w": a, "b": a[0]}), raw=False, strict_map_key=False)
    assert raw["w"]["__msgpack_chunked_array__"] is True and raw["w"]["shape"] == {"0": 4, "1": 6}
    assert sorted(raw["w"]["chunks"]) == ["0", "1", "2"] and not isinstance(raw["b"], dict)
    back = msgpack_restore(msgpack_serialize({"w": a, "b": a[0]}))
    np.testing.assert_array_equal(back["w"], a)
    np.testing.assert_array_equal(back["b"], a[0])
