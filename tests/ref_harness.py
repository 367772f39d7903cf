"""TEST INFRASTRUCTURE — runs the reference's own code (/root/reference/stable_diffusion/*.py, UNMODIFIED) on the
torch-backed Keras stand-in of oracle/keras_shim.  Only usable where /root/reference exists (the authoring container);
the GPU box gets the outputs as committed fixtures (tests/golden/reference_run.npz, tools/make_golden_ref.py)."""
import contextlib
import io
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE = "/root/reference"
SHIM = os.path.join(ROOT, "oracle", "keras_shim")


def available():
    return os.path.isdir(os.path.join(REFERENCE, "stable_diffusion"))


def import_reference():
    """-> the reference's `stable_diffusion` package, imported with the shim standing in for `keras`."""
    if not available():
        raise RuntimeError("the reference tree is not present on this machine")
    if "keras" in sys.modules and not getattr(sys.modules["keras"], "__version__", "").endswith("shim"):
        raise RuntimeError("a real keras is already imported")
    for p in (REFERENCE, SHIM):
        if p not in sys.path:
            sys.path.insert(0, p)
    import stable_diffusion  # noqa: the reference package (ours is minsdtf_b200.stable_diffusion)
    assert os.path.dirname(stable_diffusion.__file__).startswith(REFERENCE)
    return stable_diffusion


def write_checkpoints(directory):
    """Seeded synthetic SD1.5 checkpoints in the reference's file formats -> dict of paths (+ the state dicts)."""
    import torch
    from minsdtf_b200 import synth
    os.makedirs(directory, exist_ok=True)
    sds = {"unet": synth.make_state_dict("unet"), "vae": synth.make_vae_state_dict(),
           "controlnet": synth.make_controlnet_state_dict(), "text_encoder": synth.make_state_dict("text_encoder")}
    paths = {k: os.path.join(directory, k + (".pth" if k == "controlnet" else ".safetensors")) for k in sds}
    for k, sd in sds.items():
        if not os.path.exists(paths[k]):
            if k == "controlnet":  # control_sd15_canny.pth is a torch pickle (control_net.py:33)
                torch.save(sd, paths[k])
            else:
                synth.save_safetensors(sd, paths[k])
    paths["vocab"] = os.path.join(directory, "bpe_vocab.txt.gz")
    if not os.path.exists(paths["vocab"]):
        synth.make_bpe_vocab(paths["vocab"])
    return paths, sds


class Recorder:
    """wraps a model: remembers the inputs of its last `predict_on_batch` (to read the final latent out of the loop)"""

    def __init__(self, model):
        self.model, self.last_input = model, None

    def predict_on_batch(self, x):
        self.last_input = x
        return self.model.predict_on_batch(x)


def quiet(fn, *args, **kwargs):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*args, **kwargs)


def reference_pipeline(paths, control=False, active_tcd=False, clip_skip=-1, lora_path=None, size=128):
    """The reference's `StableDiffusion`, constructed with local checkpoint paths so that nothing is downloaded."""
    ref = import_reference()
    from stable_diffusion.clip_tokenizer import SimpleTokenizer
    from stable_diffusion.stable_diffusion import StableDiffusion
    sd = quiet(StableDiffusion, img_height=size, img_width=size, clip_skip=clip_skip, unet_ckpt=paths["unet"],
               text_encoder_ckpt=paths["text_encoder"], vae_ckpt=paths["vae"], lora_path=lora_path,
               controlnet_path=paths["controlnet"] if control else None, active_tcd=active_tcd)
    sd._tokenizer = SimpleTokenizer(bpe_path=paths["vocab"])  # the default constructor downloads the vocabulary
    quiet(lambda: (sd.diffusion_model, sd.image_decoder, sd.text_encoder, sd.text_clip_embedding))
    sd._image_decoder = Recorder(sd._image_decoder)
    return sd
