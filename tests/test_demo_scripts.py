"""demos/sd_demo.py — the reference's demo scripts (text_to_image.py, image_to_image.py, inpaint.py, app.py pipelines)
pointed at the engine."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "demos"))


def test_cli_arguments_mirror_the_reference_pages():
    import sd_demo
    a = sd_demo.build_parser().parse_args(["inpaint", "a (cat:1.2)", "--image", "i.png", "--mask", "m.png", "--seed", "7"])
    assert (a.mode, a.steps, a.guidance_scale, a.strength, a.mask_blur) == ("inpaint", 25, 7.0, 0.8, 5)  # the pages' defaults
    calls = {}

    class Fake:
        def inpaint(self, **kw):
            calls.update(kw)
            return np.zeros((1, 8, 8, 3), np.uint8)

    sd_demo.run(Fake(), "inpaint", a.prompt, a.negative_prompt, a.steps, a.guidance_scale, a.seed, 1, a.image, a.strength, a.mask, a.mask_blur)
    assert calls["seed"] == 7 and calls["inpaint_mask"] == "m.png" and calls["mask_blur_strength"] == 5
    assert calls["num_steps"] == 25 and calls["unconditional_guidance_scale"] == 7.0 and calls["reference_image_strength"] == 0.8


@pytest.mark.gpu
def test_cli_generates_an_image_from_a_string_prompt(tmp_path):
    import sd_demo
    from PIL import Image
    out = str(tmp_path / "o.png")
    imgs = sd_demo.main(["txt2img", "a photo of an (astronaut:1.3)", "--synthetic", "--height", "128", "--width", "128", "--steps", "2",
                         "--seed", "3", "-o", out])
    assert imgs.shape == (1, 128, 128, 3) and np.array_equal(np.array(Image.open(out)), imgs[0])
