#!/usr/bin/env python
"""The reference's demo scripts (text_to_image.py, image_to_image.py, inpaint.py: one Gradio page each, and the three
pipelines of app.py) pointed at the B200 engine: the same `StableDiffusion` calls with the same arguments, as a command
line tool and — when `gradio` is installed — as the same three-tab page.

    python demos/sd_demo.py txt2img "a photo of an (astronaut:1.2)" --unet unet.safetensors --vae vae.safetensors \\
           --text-encoder text_encoder.safetensors --bpe-vocab bpe_simple_vocab_16e6.txt.gz -o out.png
    python demos/sd_demo.py img2img "prompt" --image in.png --strength 0.8 ...
    python demos/sd_demo.py inpaint "prompt" --image in.png --mask mask.png --mask-blur 5 ...
    python demos/sd_demo.py ui ...                       # Gradio page (reference: text_to_image.py:17-44 etc.)

`--synthetic` runs on seeded random-init weights and a synthetic vocabulary (no checkpoints are available offline).
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def build_parser():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("mode", choices=["txt2img", "img2img", "inpaint", "ui"])
    ap.add_argument("prompt", nargs="?", default="hello stable diffusion")
    ap.add_argument("--negative-prompt", default="")
    ap.add_argument("--steps", type=int, default=25)
    ap.add_argument("--guidance-scale", type=float, default=7.0)
    ap.add_argument("--seed", type=int, default=-1, help="-1: random, as the reference's pages")
    ap.add_argument("--batch-size", type=int, default=1)
    ap.add_argument("--height", type=int, default=512)
    ap.add_argument("--width", type=int, default=512)
    ap.add_argument("--image", help="reference image (img2img / inpaint)")
    ap.add_argument("--strength", type=float, default=0.8, help="reference_image_strength")
    ap.add_argument("--mask", help="inpaint mask (white = repaint)")
    ap.add_argument("--mask-blur", type=int, default=5)
    ap.add_argument("--control-image", help="ControlNet (canny) image")
    ap.add_argument("--embedding", help="textual-inversion file")
    ap.add_argument("--tcd", action="store_true", help="TCD scheduler (active_tcd=True)")
    ap.add_argument("--clip-skip", type=int, default=-1)
    for name in ("unet", "vae", "text-encoder", "controlnet", "lora", "bpe-vocab"):
        ap.add_argument(f"--{name}", default=None)
    ap.add_argument("--synthetic", action="store_true")
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("-o", "--output", default="output.png")
    return ap


def make_pipeline(a):
    from minsdtf_b200.stable_diffusion import StableDiffusion
    vocab = a.bpe_vocab
    if a.synthetic and vocab is None:
        import tempfile
        from minsdtf_b200 import synth
        vocab = synth.make_bpe_vocab(os.path.join(tempfile.gettempdir(), "sdtf_synthetic_bpe_vocab.txt.gz"))
    return StableDiffusion(img_height=a.height, img_width=a.width, clip_skip=a.clip_skip, unet_ckpt=a.unet, text_encoder_ckpt=a.text_encoder,
                           vae_ckpt=a.vae, lora_path=a.lora, controlnet_path=a.controlnet, active_tcd=a.tcd, device=a.device,
                           synthetic=a.synthetic, bpe_vocab=vocab)


def run(sd, mode, prompt, negative_prompt="", steps=25, guidance_scale=7.0, seed=-1, batch_size=1, image=None, strength=0.8,
        mask=None, mask_blur=5, control_image=None, embedding=None, callback=None):
    """the calls of the reference's inference_fn functions (text_to_image.py:6-14, image_to_image.py:6-16, inpaint.py:6-20)"""
    common = dict(prompt=prompt, negative_prompt=negative_prompt, batch_size=batch_size, num_steps=steps,
                  unconditional_guidance_scale=guidance_scale, seed=None if seed == -1 else seed, embedding=embedding,
                  control_net_image=control_image, callback=callback)
    if mode == "txt2img":
        return sd.text_to_image(**common)
    if mode == "img2img":
        return sd.image_to_image(reference_image=image, reference_image_strength=strength, **common)
    return sd.inpaint(reference_image=image, reference_image_strength=strength, inpaint_mask=mask, mask_blur_strength=mask_blur, **common)


def launch_ui(sd, a):
    import gradio as gr  # optional dependency, as in the reference's scripts
    with gr.Blocks() as app:
        for mode, label in (("txt2img", "Text2Image"), ("img2img", "Image2Image"), ("inpaint", "Inpaint")):
            with gr.Tab(label):
                prompt = gr.Textbox(label="prompt", value="hello stable diffusion")
                negative = gr.Textbox(label="negative prompt", value="")
                steps = gr.Slider(label="steps", value=25, minimum=1, maximum=100, step=1)
                scale = gr.Slider(label="guidance scale", value=7.0, minimum=0.0, maximum=100.0, step=0.01)
                seed = gr.Number(label="seed", value=-1, precision=0)
                inputs = [prompt, negative, steps, scale, seed]
                if mode != "txt2img":
                    strength = gr.Slider(label="denoise strength", value=0.8, minimum=0.0, maximum=1.0, step=0.01)
                    ref = gr.Image(width=a.width, height=a.height, label="Image 2 Image")
                    inputs += [ref, strength]
                if mode == "inpaint":
                    blur = gr.Slider(label="mask feathering strength", value=5, minimum=1, maximum=256, step=1)
                    msk = gr.Image(width=a.width, height=a.height, label="Inpaint Mask")
                    inputs += [msk, blur]
                out = gr.Image(width=a.width, height=a.height)

                def fn(p, n, st, sc, sd_, *rest, _mode=mode):
                    kw = {}
                    if _mode != "txt2img":
                        kw.update(image=rest[0], strength=rest[1])
                    if _mode == "inpaint":
                        kw.update(mask=rest[2], mask_blur=int(rest[3]))
                    return run(sd, _mode, p, n, int(st), sc, int(sd_), **kw)[0]

                gr.Button("inference").click(fn=fn, inputs=inputs, outputs=out)
    app.launch()


def main(argv=None):
    a = build_parser().parse_args(argv)
    sd = make_pipeline(a)
    if a.mode == "ui":
        return launch_ui(sd, a)
    if a.mode != "txt2img" and not a.image:
        raise SystemExit(f"{a.mode} needs --image")
    if a.mode == "inpaint" and not a.mask:
        raise SystemExit("inpaint needs --mask")
    ctrl = np.array(__import__("PIL.Image", fromlist=["Image"]).open(a.control_image).convert("RGB")) if a.control_image else None
    done = []
    imgs = run(sd, a.mode, a.prompt, a.negative_prompt, a.steps, a.guidance_scale, a.seed, a.batch_size, a.image, a.strength, a.mask,
               a.mask_blur, ctrl, a.embedding, callback=lambda i: (done.append(i), print(f"\rstep {i}", end="", flush=True)))
    print()
    from PIL import Image
    root, ext = os.path.splitext(a.output)
    for i, img in enumerate(imgs):
        path = a.output if len(imgs) == 1 else f"{root}_{i}{ext}"
        Image.fromarray(img).save(path)
        print("saved", path)
    return imgs


if __name__ == "__main__":
    main()
