"""TEST INFRASTRUCTURE ONLY -- numpy stand-in for `flax` (only `flax.linen`), see oracle/flax_shim/README.md."""
from . import linen  # noqa: F401
