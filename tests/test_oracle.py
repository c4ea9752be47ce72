"""The oracle against everything that can pin it without the reference:
documented known answers, committed golden vectors, an independent restatement, einops."""
import numpy as np
import pytest
import torch

from oracle import vit_numpy, vit_torch
from vit_flax_b200.params import count_params, init_params, perturb_params
from _util import C1, C2, TINY, TINY_MEAN, images_for, load_golden


def test_readme_known_answers():
    # README.md:34 -> (1, 1000); vit.py:195-197 prints the parameter count
    v = init_params(seed=1, **C1)
    y = vit_numpy.vit_forward(v, images_for(C1, 1), dtype=np.float32, **C1)
    assert y.shape == (1, 1000)
    assert count_params(**C1) == 54_622_184
    assert count_params(**C2) == 86_540_008      # no qkv bias (vit.py:68)


def test_zero_image_reference_init_gives_zero_logits():
    # every bias, cls and pos_embedding is zero-initialised (vit.py:142-144): LN(0) = 0
    v = init_params(seed=3, **TINY)
    y = vit_numpy.vit_forward(v, np.zeros((2, 32, 32, 3), np.float32), **TINY)
    assert np.all(y == 0.0)


def test_patch_permutation_invariance_without_pos_embedding():
    v = perturb_params(init_params(seed=4, **TINY), seed=5)
    v["params"]["pos_embedding"][:] = 0.0
    img = images_for(TINY, 1, seed=6)
    p = 8
    grid = img.reshape(1, 4, p, 4, p, 3)
    perm = np.random.default_rng(0).permutation(16)
    patches = grid.transpose(0, 1, 3, 2, 4, 5).reshape(1, 16, p, p, 3)[:, perm]
    img2 = patches.reshape(1, 4, 4, p, p, 3).transpose(0, 1, 3, 2, 4, 5).reshape(1, 32, 32, 3)
    y1 = vit_numpy.vit_forward(v, img, **TINY)
    y2 = vit_numpy.vit_forward(v, img2, **TINY)
    np.testing.assert_allclose(y1, y2, atol=1e-10)


def test_patchify_matches_einops():
    from einops import rearrange
    x = images_for(dict(image_size=(16, 24)), 2, seed=1)
    want = rearrange(x, "b (h p1) (w p2) c -> b (h w) (p1 p2 c)", p1=4, p2=8)   # vit.py:146
    np.testing.assert_array_equal(vit_numpy.patchify(x, 4, 8), want)


def test_softmax_and_layernorm_semantics():
    rng = np.random.default_rng(0)
    s = vit_numpy.softmax_last(rng.standard_normal((3, 5, 7)) * 30)
    np.testing.assert_allclose(s.sum(-1), 1.0, atol=1e-12)
    x = rng.standard_normal((4, 64)) * 3 + 5
    p = {"scale": rng.standard_normal(64), "bias": rng.standard_normal(64)}
    got = vit_numpy.layer_norm(x, p)
    want = torch.nn.functional.layer_norm(torch.tensor(x), (64,), torch.tensor(p["scale"]),
                                          torch.tensor(p["bias"]), eps=1e-6).numpy()
    np.testing.assert_allclose(got, want, atol=1e-9)
    g = vit_numpy.gelu_tanh(np.linspace(-6, 6, 101))
    gt = torch.nn.functional.gelu(torch.linspace(-6, 6, 101, dtype=torch.float64), approximate="tanh").numpy()
    np.testing.assert_allclose(g, gt, atol=1e-12)
    # erf-GELU is NOT what flax computes by default: make sure we did not restate that one
    ge = torch.nn.functional.gelu(torch.linspace(-6, 6, 101, dtype=torch.float64)).numpy()
    assert np.abs(g - ge).max() > 1e-4


@pytest.mark.parametrize("name,cfg,pool", [("tiny_cls.npz", TINY, "cls"), ("tiny_mean.npz", TINY_MEAN, "mean")])
def test_golden_tiny(name, cfg, pool):
    v, meta = load_golden(name)
    y = vit_numpy.vit_forward(v, meta["images"], pool=pool, **cfg)
    np.testing.assert_allclose(y, meta["logits"], atol=1e-12)
    yt = vit_torch.vit_forward(vit_torch.tree_to_torch(v), meta["images"], pool=pool, **cfg).numpy()
    np.testing.assert_allclose(yt, meta["logits"], atol=1e-5)      # independent restatement


def test_golden_tokens():
    v, meta = load_golden("tiny_cls.npz")
    _, tok = vit_numpy.vit_forward(v, meta["images"], return_tokens=True, **TINY)
    np.testing.assert_allclose(tok, meta["tokens"], atol=1e-12)


def test_golden_c1_from_seeds():
    _, meta = load_golden("c1_logits.npz")
    v = perturb_params(init_params(seed=int(meta["init_seed"]), **C1), seed=int(meta["perturb_seed"]))
    img = images_for(C1, 1, seed=int(meta["image_seed"]))
    yt = vit_torch.vit_forward(vit_torch.tree_to_torch(v), img, **C1).numpy()
    np.testing.assert_allclose(yt, meta["logits"], atol=1e-4)


def test_no_project_out_when_single_head_dim64():
    v = init_params(seed=0, **TINY_MEAN)                     # heads=1, dim=64 -> vit.py:65
    assert "Dense_1" not in v["params"]["Transformer_0"]["Attention_0"]


def test_philox_known_answer_and_dropout_semantics():
    """Philox4x32-10 known-answer vectors (Salmon et al., Random123 kat_vectors) pin the generator the
    kernels' dropout masks come from; flax nn.Dropout semantics: keep w.p. 1-rate, scale by 1/(1-rate)."""
    from oracle import philox
    # counter = (0,0,0,0), key = (0,0)  ->  6627e8d5 e169c58d bc57ac4c 9b00dbd8
    got = philox.philox4x32_10(np.array([0], np.uint64), 0, 0)[0]
    assert [int(v) for v in got] == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    # counter = ffffffff x4, key = ffffffff x2  ->  408f276d 41c83b0e a20bc7c6 6d5451fd
    # (our counter layout is (quad_lo, quad_hi, site, 0): c3 = 0, so use the all-zero vector above
    #  plus the pi-digits vector below with c3 = 0 replaced -- checked against a scalar reference)
    def scalar(c, k):
        c, k = list(c), list(k)
        for _ in range(10):
            p0, p1 = 0xD2511F53 * c[0], 0xCD9E8D57 * c[2]
            c = [(p1 >> 32) ^ c[1] ^ k[0], p1 & 0xFFFFFFFF, (p0 >> 32) ^ c[3] ^ k[1], p0 & 0xFFFFFFFF]
            k = [(k[0] + 0x9E3779B9) & 0xFFFFFFFF, (k[1] + 0xBB67AE85) & 0xFFFFFFFF]
        return c
    quad, site, key = 0x123456789ABCDEF, 77, 0xDEADBEEFCAFEF00D
    want = scalar([quad & 0xFFFFFFFF, quad >> 32, site, 0], [key & 0xFFFFFFFF, key >> 32])
    assert [int(v) for v in philox.philox4x32_10(np.array([quad], np.uint64), site, key)[0]] == want
    x = np.ones((64, 256), np.float64)
    y = philox.dropout(x, 0.25, 3, 42)
    kept = y != 0
    assert abs(kept.mean() - 0.75) < 0.02
    np.testing.assert_allclose(y[kept], 1.0 / 0.75, rtol=1e-6)
    np.testing.assert_array_equal(y, philox.dropout(x, 0.25, 3, 42))          # same key, same mask
    assert (philox.dropout(x, 0.25, 4, 42) != y).any() and (philox.dropout(x, 0.25, 3, 43) != y).any()


def test_oracle_matches_third_party_vit_with_mapped_weights():
    """Independent pin: HuggingFace ``transformers`` ViTForImageClassification (written by other people,
    from the same paper) configured like vit.py -- pre-norm blocks, tanh GELU (nn.gelu default), LayerNorm
    eps 1e-6, no qkv bias (vit.py:68), head_dim 64, cls pooling -- and loaded with a reference-layout
    params tree through the layout mapping below must give the oracle's logits.  This does not replace
    a run of the Flax reference (impossible here), but it rules out a shared misreading of the
    architecture in oracle/vit_numpy.py and oracle/vit_torch.py."""
    tr = pytest.importorskip("transformers")
    cfg = dict(image_size=32, patch_size=8, num_classes=10, dim=128, depth=2, heads=2, mlp_dim=256)   # inner = dim
    P, D, I = cfg["patch_size"], cfg["dim"], 64 * cfg["heads"]
    v = perturb_params(init_params(seed=11, **cfg), seed=12)
    p = v["params"]
    hf_cfg = tr.ViTConfig(hidden_size=D, num_hidden_layers=cfg["depth"], num_attention_heads=cfg["heads"],
                          intermediate_size=cfg["mlp_dim"], hidden_act="gelu_pytorch_tanh", layer_norm_eps=1e-6,
                          qkv_bias=False, image_size=cfg["image_size"], patch_size=P, num_labels=cfg["num_classes"],
                          hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    hf = tr.ViTForImageClassification(hf_cfg).double().eval()
    t = lambda a: torch.as_tensor(np.asarray(a, dtype=np.float64))
    sd = {
        "vit.embeddings.cls_token": t(p["cls"]),
        "vit.embeddings.position_embeddings": t(p["pos_embedding"]),
        # Dense_0 kernel [(p1 p2 c), D] (vit.py:146-147) -> conv weight [D, c, p1, p2]
        "vit.embeddings.patch_embeddings.projection.weight": t(p["Dense_0"]["kernel"]).reshape(P, P, 3, D).permute(3, 2, 0, 1),
        "vit.embeddings.patch_embeddings.projection.bias": t(p["Dense_0"]["bias"]),
        "vit.layernorm.weight": t(p["LayerNorm_0"]["scale"]), "vit.layernorm.bias": t(p["LayerNorm_0"]["bias"]),
        "classifier.weight": t(p["Dense_1"]["kernel"]).T, "classifier.bias": t(p["Dense_1"]["bias"]),
    }
    tp = p["Transformer_0"]
    for l in range(cfg["depth"]):
        pre = f"vit.encoder.layer.{l}."
        qkv = t(tp[f"Attention_{l}"]["Dense_0"]["kernel"])                     # [D, 3*inner], q | k | v (vit.py:69)
        for i, name in enumerate(("query", "key", "value")):
            sd[pre + f"attention.attention.{name}.weight"] = qkv[:, i * I:(i + 1) * I].T
        sd[pre + "attention.output.dense.weight"] = t(tp[f"Attention_{l}"]["Dense_1"]["kernel"]).T
        sd[pre + "attention.output.dense.bias"] = t(tp[f"Attention_{l}"]["Dense_1"]["bias"])
        sd[pre + "intermediate.dense.weight"] = t(tp[f"FeedForward_{l}"]["Dense_0"]["kernel"]).T
        sd[pre + "intermediate.dense.bias"] = t(tp[f"FeedForward_{l}"]["Dense_0"]["bias"])
        sd[pre + "output.dense.weight"] = t(tp[f"FeedForward_{l}"]["Dense_1"]["kernel"]).T
        sd[pre + "output.dense.bias"] = t(tp[f"FeedForward_{l}"]["Dense_1"]["bias"])
        for hf_name, ours in (("layernorm_before", f"PreNorm_{2 * l}"), ("layernorm_after", f"PreNorm_{2 * l + 1}")):
            sd[pre + hf_name + ".weight"] = t(tp[ours]["LayerNorm_0"]["scale"])
            sd[pre + hf_name + ".bias"] = t(tp[ours]["LayerNorm_0"]["bias"])
    missing, unexpected = hf.load_state_dict({k: x.contiguous() for k, x in sd.items()}, strict=False)
    assert not missing and not unexpected, (missing, unexpected)
    img = images_for(cfg, 3, seed=13)
    with torch.no_grad():
        want = hf(pixel_values=torch.as_tensor(img, dtype=torch.float64).permute(0, 3, 1, 2)).logits.numpy()
    got = vit_numpy.vit_forward(v, img, **cfg)
    assert np.abs(want).max() > 0.1                                            # a non-trivial comparison
    np.testing.assert_allclose(got, want, atol=1e-10)


def test_vjp_oracle_matches_finite_differences():
    """oracle/vit_torch.vit_vjp (float64 autograd through the restatement) is the checker of the CUDA
    backward pass; pin it with central differences on a few entries of every kind of leaf."""
    cfg = dict(image_size=16, patch_size=8, num_classes=8, dim=64, depth=2, heads=2, mlp_dim=96)
    v = perturb_params(init_params(seed=51, **cfg), seed=52)
    img = images_for(cfg, 2, seed=53)
    dl = np.random.default_rng(54).standard_normal((2, 8))
    for pool in ("cls", "mean"):
        _, g = vit_torch.vit_vjp(v, img, dl, pool=pool, **cfg)
        probes = [(("cls",), (0, 0, 5)), (("pos_embedding",), (0, 2, 7)), (("Dense_0", "kernel"), (17, 3)),
                  (("Transformer_0", "Attention_1", "Dense_0", "kernel"), (4, 100)),
                  (("Transformer_0", "PreNorm_1", "LayerNorm_0", "scale"), (9,)),
                  (("Transformer_0", "FeedForward_0", "Dense_0", "bias"), (11,)), (("Dense_1", "kernel"), (6, 2))]
        for path, idx in probes:
            import copy
            eps = 1e-5
            vals = []
            for sgn in (1, -1):
                w = copy.deepcopy(v)
                w64 = vit_torch.tree_to_torch(w, torch.float64)
                node = w64["params"]
                for k in path[:-1]:
                    node = node[k]
                node[path[-1]][idx] += sgn * eps
                vals.append(float((vit_torch.vit_forward(w64, img, pool=pool, **cfg).numpy() * dl).sum()))
            fd = (vals[0] - vals[1]) / (2 * eps)
            ref = g
            for k in path:
                ref = ref[k]
            assert abs(fd - ref[idx]) < 1e-6 * max(1.0, abs(fd)), (pool, path, fd, ref[idx])


def _round_kernels(tree, dt):
    """Every Dense kernel rounded to the 16-bit type `dt` and back (biases / LayerNorm affine / cls / pos stay fp32,
    as they do on the tensor-core path)."""
    if isinstance(tree, dict):
        return {k: (v.to(dt).to(torch.float32) if k == "kernel" else _round_kernels(v, dt)) for k, v in tree.items()}
    return tree


def test_bf16_operand_floor_on_vit_b16():
    """BASELINE.json's north star asks for bf16 logits within max-abs 2e-2 of the fp32 reference.  On the
    metric's own config (ViT-B/16, reference initialisers => unit-variance logits) that bound is not
    reachable by ANY bf16-operand implementation: rounding the Dense KERNELS ALONE to bf16 -- exact fp32
    arithmetic everywhere else -- already moves the logits by ~2.2e-2, and rounding every tensor-core operand
    (what a bf16 GEMM path must do) by ~3e-2.  The error is spread evenly over patch / qkv / out / ff1 / ff2 /
    head (no single GEMM to special-case).  fp16 operands (same tcgen05 rate, 3 more significand bits) stay
    8x below the bound, which is why fp16 is the shipped default and bf16 is held to this derived floor
    (tests/test_gpu_forward.py: error <= max(2e-2, 1.25 x this emulated error), and within 2e-2 of the
    same-rounding emulation)."""
    variables = perturb_params(init_params(seed=1, **C2), seed=2)
    img = images_for(C2, 8, seed=0)
    pt = vit_torch.tree_to_torch(variables)
    want = vit_torch.vit_forward(pt, img, **C2).numpy()
    err = {}
    for name, dt in (("bf16", torch.bfloat16), ("fp16", torch.float16)):
        w_only = vit_torch.vit_forward(_round_kernels(pt, dt), img, **C2).numpy()
        full = vit_torch.vit_forward(pt, img, operand_dtype=dt, **C2).numpy()
        err[name] = (float(np.abs(w_only - want).max()), float(np.abs(full - want).max()))
    print(f"[bf16 floor] ViT-B/16, 8 images: weights-only / every-operand rounding: bf16 {err['bf16']}, fp16 {err['fp16']}")
    assert 0.9 < want.std() < 1.1                 # unit-variance logits: the absolute tolerance is a relative one
    assert err["bf16"][0] > 1.8e-2                # weights alone (measured 2.19e-2)
    assert err["bf16"][1] > 2e-2                  # every operand (measured 3.03e-2): the north-star bound is out of reach
    assert err["fp16"][1] < 6e-3                  # fp16 operands (measured 3.2e-3)
