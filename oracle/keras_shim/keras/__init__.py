"""TEST INFRASTRUCTURE — a torch-backed stand-in for the slice of the Keras 3 API that cpuimage/minSDTF uses.

Why: `keras` / `tensorflow` are not installable offline, so the reference's graphs cannot be executed as shipped.  With this
package first on `sys.path`, `/root/reference/stable_diffusion/*.py` imports and runs UNMODIFIED: its own model-building
code (diffusion_model.py, control_net.py, image_decoder.py, image_encoder.py, layers.py, text_encoder.py), its own
positional weight loader (ckpt_loader.load_weights_from_file over `model.weights` / `set_weights`), and its own
`generate_image` loop (stable_diffusion.py).  Only the primitives are restated here, from the documented Keras 3
semantics: Dense (kernel (in,out)), Conv2D (kernel HWIO, channels_last, 'valid'), ZeroPadding2D, GroupNormalization /
LayerNormalization (biased variance, rsqrt(var + eps)), UpSampling2D (nearest), Embedding, Concatenate, Activation,
softmax / einsum / reshape / transpose, and the bookkeeping that decides the ORDER of `model.weights`:
  * a layer's weights = its own variables, then those of its sub-layers in attribute-assignment order (lists included),
    sub-layers created in `build()` after those created in `__init__`;
  * a functional `Model`'s layers = the graph's operations sorted by depth from the outputs (deepest first), ties broken by
    depth-first traversal order from the outputs (keras/src/ops/function.py `map_graph`);
  * a `Sequential`'s layers = list order.
The positional loader only reproduces the oracle's name-keyed results if that emulation is right, which is what
tests/test_cpu_reference_harness.py checks (and `set_weights` refuses any shape mismatch).

Nothing outside tests/ and tools/make_golden_ref.py imports this.
"""
from . import activations, layers, ops, random, utils  # noqa: F401
from .core import Model, Sequential  # noqa: F401

__version__ = "3.shim"
