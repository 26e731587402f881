#!/usr/bin/env python
"""Per-image parity + quality report (north_star: bit-exactness of coefficients and file bytes,
mismatch counts, decoded PSNR as defined by the reference's analyze_results.py).

Runs on a B200: every image goes BMP file -> jpeg_compression_app (our CLI, CUDA path) and, when
oracle/_ref exists, through the reference CLI; coefficients come from the device tap.
    python tools/parity_report.py [--out profiles/parity_report.md] [bmp files ...]
Without file arguments the synthetic shapes of BASELINE.json (1080p, 4K) and edge sizes are used."""
import argparse
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import jpeg_image_compression_b200 as jb  # noqa: E402
from jpeg_image_compression_b200.quality import analyze, coefficient_mismatches  # noqa: E402
from oracle.oracle import REF_APP, Oracle, write_bmp  # noqa: E402  (checker only)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("files", nargs="*")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    orc = Oracle()
    enc = jb.DeviceEncoder(0)
    app = os.path.join(ROOT, "jpeg_image_compression_b200", "jpeg_compression_app")
    rows = []
    with tempfile.TemporaryDirectory() as td:
        cases = [(os.path.basename(f), f) for f in args.files]
        if not cases:
            # BASELINE configs[0]: the reference's own assets at full size (copied next to oracle/_ref by build())
            adir = os.path.join(ROOT, "oracle", "_ref", "assets")
            for n in ("lena", "blackbuck", "greenland", "offset_sample"):
                if os.path.exists(os.path.join(adir, n + ".bmp")):
                    cases.append(("asset_" + n, os.path.join(adir, n + ".bmp")))
            for name, (w, h, seed, amp) in {"synth_1920x1080_amp20": (1920, 1080, 0, 20), "synth_3840x2160_amp20": (3840, 2160, 1, 20),
                                            "synth_3840x2160_amp0": (3840, 2160, 1, 0), "synth_3840x2160_amp64": (3840, 2160, 1, 64),
                                            "synth_762x1309_amp20": (762, 1309, 5, 20), "synth_1283x725_amp20": (1283, 725, 9, 20)}.items():
                p = os.path.join(td, name + ".bmp")
                write_bmp(p, orc.synth_rgb(w, h, seed, amp))
                cases.append((name, p))
        for name, bmp in cases:
            rgb = jb.loadBMPImage(bmp)
            ours, ref = os.path.join(td, name + ".jpg"), os.path.join(td, name + "_ref.jpg")
            subprocess.run([app, bmp, ours], check=True, stdout=subprocess.DEVNULL)
            identical = "n/a"
            if os.path.exists(REF_APP):
                subprocess.run([REF_APP, bmp, ref], check=True, stdout=subprocess.DEVNULL)
                identical = "yes" if open(ours, "rb").read() == open(ref, "rb").read() else "NO"
            enc.encode(rgb)
            nb = ((rgb.shape[0] + 7) // 8) * ((rgb.shape[1] + 7) // 8)
            mm = coefficient_mismatches(enc.coefficients(nb), orc.coefficients(rgb))
            q = analyze(bmp, ours)
            rows.append((name, f"{rgb.shape[1]}x{rgb.shape[0]}", identical, mm["mismatched"], mm["off_by_one"], q["file_size_comp"],
                         f"{q['compression_ratio']:.1f}", f"{q['bpp']:.3f}", f"{q['mse']:.2f}", f"{q['psnr']:.2f}"))
    lines = ["| image | size | file == reference build | coefficient mismatches | of which +-1 | bytes | CR | bpp | MSE | PSNR dB |",
             "|---|---|---|---|---|---|---|---|---|---|"] + ["| " + " | ".join(str(c) for c in r) + " |" for r in rows]
    text = "\n".join(lines)
    print(text)
    if args.out:
        with open(args.out, "w") as f:
            f.write("# Parity and quality per image (GPU CLI vs reference natural_c build; metrics as analyze_results.py)\n\n" + text + "\n")


if __name__ == "__main__":
    main()
