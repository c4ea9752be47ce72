"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference ViT forward.

Nothing under ``oracle/`` is part of the product path.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl
reference`` legs may import it, and there only as the checker or the timed
CPU baseline, never as a fallback for the CUDA path.

PARITY UNPINNED: the reference (``/root/reference/vit_flax/vit.py``) cannot be
imported here (jax / jaxlib / flax are absent and not installable) and ships no
tests, golden vectors or fixtures for this path.  The
oracle therefore restates ``vit.py`` under the published Flax/JAX semantics and
is cross-checked by two independent restatements (numpy float64 in
``vit_numpy.py``, torch-CPU float32 in ``vit_torch.py``), an einops check of the
patchify order, and the known answers the reference does document (output shape
``(1, 1000)`` -- README.md:34 -- and the parameter counts implied by
vit.py:142-165).  A third-party implementation of the same architecture
(HuggingFace ``transformers`` ViTForImageClassification, configured like vit.py
and loaded through an explicit weight-layout mapping) reproduces the oracle's
logits to 1e-10 in float64 (tests/test_oracle.py) -- independent evidence, not a
run of the reference.
"""
