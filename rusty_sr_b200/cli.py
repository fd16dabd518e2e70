"""Command line with the reference's flag surface (src/main.rs:33-116, upscale() :130-178):

    python -m rusty_sr_b200.cli INPUT_FILE OUTPUT_FILE [-p imagenet|imagenetlinear|anime|bilinear]
                                                       [-c FILE.rsr] [-d]

Image decode/encode is host I/O (the reference uses the `image` crate, main.rs:164,175); here PIL.
`-p bilinear` and `-d` (network.rs:111-138) run as small elementwise GPU kernels.
"""
from __future__ import annotations

import argparse
import sys

import numpy as np


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(prog="rusty_sr", description="A convolutional neural network trained to upscale images")
    ap.add_argument("INPUT_FILE", help="Sets the input image to upscale")
    ap.add_argument("OUTPUT_FILE", help="Sets the output file to write/overwrite (.png recommended)")
    g = ap.add_mutually_exclusive_group()
    g.add_argument("-p", "--parameters", choices=["imagenet", "imagenetlinear", "anime", "bilinear"],
                   help="Sets which built-in parameters to use with the neural net")
    g.add_argument("-c", "--custom", metavar="PARAMETER_FILE",
                   help="Sets a custom parameter file (.rsr) to use with the neural net")
    g.add_argument("-d", "--downsample", action="store_true",
                   help="Perform downsampling rather than upscaling")
    ap.add_argument("--device", type=int, default=0, help="CUDA device ordinal (extension)")
    args = ap.parse_args(argv)

    from PIL import Image
    import rusty_sr_b200 as R

    if args.downsample or args.parameters == "bilinear":
        img = np.asarray(Image.open(args.INPUT_FILE).convert("RGBA"))
        if args.downsample:
            print("Downsampling using average pooling of linear RGB values...", end="", flush=True)   # main.rs:140
            out = R.downsample(img, R.FACTOR, device=args.device)
        else:
            print("Upscaling using bilinear interpolation...", end="", flush=True)                    # main.rs:154
            out = R.bilinear(img, R.FACTOR, device=args.device)
        print(" Writing file...", end="", flush=True)
        Image.fromarray(out, "RGBA").save(args.OUTPUT_FILE)
        print(" Done")
        return 0

    name = args.parameters or "imagenet"  # default: main.rs:144
    if args.custom:
        print("Upscaling using custom neural net parameters...", end="", flush=True)
    else:
        label = {"imagenet": "imagenet", "imagenetlinear": "linear loss imagenet", "anime": "anime"}[name]
        print(f"Upscaling using {label} neural net parameters...", end="", flush=True)
    img = np.asarray(Image.open(args.INPUT_FILE).convert("RGBA"))  # alpha is discarded by img_to_data
    out = R.upscale(img, parameters=name, custom=args.custom, device=args.device)
    print(" Writing file...", end="", flush=True)
    Image.fromarray(out, "RGBA").save(args.OUTPUT_FILE)
    print(" Done")
    return 0


if __name__ == "__main__":
    sys.exit(main())
