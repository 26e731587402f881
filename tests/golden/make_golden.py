#!/usr/bin/env python
"""Mint the golden fixtures from the UNMODIFIED reference build.

Run in the build container (needs /root/reference and oracle/_ref):

    make -C oracle && python tests/golden/make_golden.py

Every expected value written here comes from oracle/_ref/libnaturalc_ref.so, i.e.
the reference's own natural_c objects (loadBMPImage, convertBMPToJPEGGrayscale,
centerYImage, performDCT, quantizeImage, performZigZag, performRLE, encodeHuffman,
saveJPEGGrayscale) -- never from our restatement or from the CUDA path.

Outputs (committed):
    tests/golden/golden_v1.npz      small cases: input RGB + every stage output
    tests/golden/synth_hashes.json  SHA-256 of reference outputs on full-size synthetic inputs
    tests/golden/bmp/*.bmp          tiny BMP files + expected decoded RGB in the npz
"""
import hashlib
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.oracle import Oracle, Ref, write_bmp  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
ASSETS = "/root/reference/assets/input"


def sha(b) -> str:
    return hashlib.sha256(bytes(b)).hexdigest()


def small_cases(ref: Ref, orc: Oracle) -> dict:
    rng = np.random.default_rng(20261018)
    cases = {}
    lena = ref.load_bmp(f"{ASSETS}/lena.bmp")
    buck = ref.load_bmp(f"{ASSETS}/blackbuck.bmp")
    green = ref.load_bmp(f"{ASSETS}/greenland.bmp")
    offs = ref.load_bmp(f"{ASSETS}/offset_sample.bmp")
    cases["lena_crop256"] = lena[128:384, 128:384]
    cases["blackbuck_crop128"] = buck[200:328, 180:308]
    cases["greenland_corner250x205"] = green[-205:, -250:]          # both dims unaligned, 3W%16!=0
    cases["offset_crop320x213"] = offs[-213:, 400:720]              # height unaligned
    cases["one_pixel"] = np.array([[[200, 30, 90]]], np.uint8)
    cases["w7_h9"] = rng.integers(0, 256, (9, 7, 3), dtype=np.uint8)
    cases["w8_h8"] = rng.integers(0, 256, (8, 8, 3), dtype=np.uint8)
    cases["w17_h3"] = rng.integers(0, 256, (3, 17, 3), dtype=np.uint8)
    cases["w33_h41_smooth"] = orc.synth_rgb(33, 41, 7, 0)
    cases["black16"] = np.zeros((16, 16, 3), np.uint8)
    cases["white16"] = np.full((16, 24, 3), 255, np.uint8)
    cases["noise64"] = rng.integers(0, 256, (64, 64, 3), dtype=np.uint8)   # max symbols/block
    yy, xx = np.mgrid[0:32, 0:40]
    chk = (((xx + yy) & 1) * 255).astype(np.uint8)                          # Nyquist: coef 63 != 0 (no EOB)
    cases["checker_nyquist"] = np.repeat(chk[:, :, None], 3, 2)
    # a single high vertical frequency on a flat field: long zero runs -> ZRL symbols
    zrl = np.full((24, 24), 128.0)
    zrl += 120.0 * np.cos((2 * (yy[:24, :24] % 8) + 1) * 7 * np.pi / 16) * np.cos((2 * (xx[:24, :24] % 8) + 1) * 6 * np.pi / 16)
    cases["zrl_runs"] = np.repeat(np.clip(zrl, 0, 255).astype(np.uint8)[:, :, None], 3, 2)
    steps = (np.arange(48)[None, :] * 5 + np.arange(16)[:, None] * 3).astype(np.uint8)  # DC ties (sum%128==64)
    cases["dc_ramp"] = np.repeat(steps[:, :, None], 3, 2)
    cases["synth_200x120_amp20"] = orc.synth_rgb(200, 120, 1, 20)
    cases["synth_130x70_amp64"] = orc.synth_rgb(130, 70, 3, 64)
    return cases


def main() -> None:
    ref, orc = Ref(), Oracle()
    store = {}
    names = []
    for name, rgb in small_cases(ref, orc).items():
        rgb = np.ascontiguousarray(rgb, np.uint8)
        st = ref.stages(rgb)
        names.append(name)
        store[f"{name}/rgb"] = rgb
        store[f"{name}/y"] = st["y"]
        store[f"{name}/dct"] = st["dct"]
        store[f"{name}/zigzag"] = st["zigzag"]
        store[f"{name}/symbols"] = st["symbols"].view(np.uint8).reshape(-1, 6)[:, [0, 2, 3, 4]].copy()
        store[f"{name}/scan"] = np.frombuffer(st["scan"], np.uint8)
        # whole file through the reference orchestrator
        with tempfile.TemporaryDirectory() as td:
            bmp = os.path.join(td, "in.bmp")
            jpg = os.path.join(td, "out.jpg")
            write_bmp(bmp, rgb)
            loaded = ref.load_bmp(bmp)
            assert np.array_equal(loaded, rgb)
            devnull = os.open(os.devnull, os.O_WRONLY)
            saved = os.dup(1)
            os.dup2(devnull, 1)
            try:
                import ctypes as C
                from oracle.oracle import _Img
                img = _Img(rgb.shape[1], rgb.shape[0], rgb.ctypes.data)
                ok = ref.lib.saveJPEGGrayscale(jpg.encode(), C.byref(img))
                C.CDLL(None).fflush(None)
            finally:
                os.dup2(saved, 1)
                os.close(devnull)
            assert ok
            store[f"{name}/file"] = np.frombuffer(open(jpg, "rb").read(), np.uint8)
    store["names"] = np.array(names)

    # BMP reader fixtures: bottom-up, top-down, V5 header (bfOffBits=138), padded rows
    os.makedirs(os.path.join(HERE, "bmp"), exist_ok=True)
    rng = np.random.default_rng(7)
    bmp_cases = {
        "bottom_up_5x3": (rng.integers(0, 256, (3, 5, 3), dtype=np.uint8), False, 40),
        "top_down_6x4": (rng.integers(0, 256, (4, 6, 3), dtype=np.uint8), True, 40),
        "v5_header_9x5": (rng.integers(0, 256, (5, 9, 3), dtype=np.uint8), False, 124),
        "v4_header_8x2": (rng.integers(0, 256, (2, 8, 3), dtype=np.uint8), True, 108),
    }
    for name, (rgb, top_down, hsz) in bmp_cases.items():
        path = os.path.join(HERE, "bmp", name + ".bmp")
        write_bmp(path, rgb, top_down=top_down, header_size=hsz)
        got = ref.load_bmp(path)                       # what the reference loader returns
        store[f"bmp/{name}"] = got
    store["bmp_names"] = np.array(list(bmp_cases))
    np.savez_compressed(os.path.join(HERE, "golden_v1.npz"), **store)

    # full-size synthetic inputs: hashes of the reference outputs
    hashes = {}
    for (w, h, seed, amp) in [(1920, 1080, 0, 20), (1920, 1080, 64, 20), (3840, 2160, 1, 20),
                              (3840, 2160, 1, 0), (3840, 2160, 1, 64), (7680, 4320, 1, 20),
                              (762, 1309, 5, 20), (1283, 725, 9, 20)]:
        rgb = orc.synth_rgb(w, h, seed, amp)
        st = ref.stages(rgb)
        hashes[f"{w}x{h}_seed{seed}_amp{amp}"] = {
            "w": w, "h": h, "seed": seed, "amp": amp,
            "rgb_sha256": sha(rgb.tobytes()),
            "zigzag_sha256": sha(st["zigzag"].tobytes()),
            "scan_sha256": sha(st["scan"]),
            "scan_bytes": len(st["scan"]),
            "symbols": int(st["symbols"].size),
        }
        print(w, h, seed, amp, len(st["scan"]), st["symbols"].size, flush=True)
    # BASELINE config 5 is beyond the reference's int-indexed buffers: minted by the oracle's streaming twin
    big = orc.encode_scan_synth_banded(32768, 32768, 1, 20, 8, 0)
    hashes["32768x32768_seed1_amp20"] = {
        "w": 32768, "h": 32768, "seed": 1, "amp": 20, "scan_bytes": len(big), "scan_sha256": sha(big),
        "minted_by": "oracle streaming (banded) encode, orc_encode_scan_synth_banded: the stock reference cannot hold this "
                     "image (int-indexed buffers, SURVEY.md 7.3-F)"}
    del big
    # the -O2 reference twin must equal the reference built at its own flags (natural_c/Makefile:4, -g)
    ref_g_path = os.path.join(os.path.dirname(HERE), "..", "oracle", "_ref", "libnaturalc_ref_g.so")
    if os.path.exists(ref_g_path):
        ref_g = Ref(os.path.abspath(ref_g_path))
        for (w, h, seed, amp) in [(640, 360, 3, 20), (257, 131, 4, 64)]:
            rgb = orc.synth_rgb(w, h, seed, amp)
            a, b = ref.stages(rgb), ref_g.stages(rgb)
            assert bytes(a["scan"]) == bytes(b["scan"]) and np.array_equal(a["zigzag"], b["zigzag"]), "-O2 twin differs from the stock -g build"
    # the four reference assets (inputs are not committed; hashes are checked when
    # /root/reference is present, i.e. in the build container)
    for name in ["lena", "blackbuck", "greenland", "offset_sample"]:
        rgb = ref.load_bmp(f"{ASSETS}/{name}.bmp")
        st = ref.stages(rgb)
        hashes[f"asset_{name}"] = {
            "w": rgb.shape[1], "h": rgb.shape[0],
            "rgb_sha256": sha(rgb.tobytes()),
            "zigzag_sha256": sha(st["zigzag"].tobytes()),
            "scan_sha256": sha(st["scan"]), "scan_bytes": len(st["scan"]),
            "symbols": int(st["symbols"].size),
        }
    with open(os.path.join(HERE, "synth_hashes.json"), "w") as f:
        json.dump(hashes, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
