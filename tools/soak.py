"""Randomised parity soak on a B200: N images of random size / content family / alignment through the device API, the scan
bytes and the quantized coefficients compared with the oracle (oracle/: test infrastructure).
    python tools/soak.py [N=600] [seed=1]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import jpeg_image_compression_b200 as jb
from oracle.oracle import Oracle

n_img = int(sys.argv[1]) if len(sys.argv) > 1 else 600
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
orc = Oracle()
encs = {"tensor-core": jb.DeviceEncoder(0), "butterfly": jb.DeviceEncoder(0, dct_mode=2)}
encs["tensor-core"].set_concurrency(1)
bad = 0
px = 0
fam_count = {}
t0 = time.time()
for i in range(n_img):
    fam = ["synth", "noise", "flat+noise", "gradient", "extreme", "sparse-hi"][int(rng.integers(6))]
    w = int(rng.choice([rng.integers(1, 40), rng.integers(40, 700), rng.integers(700, 2600), 8 * rng.integers(1, 300), 256 * rng.integers(1, 9)]))
    h = int(rng.choice([rng.integers(1, 30), rng.integers(30, 300), 8 * rng.integers(1, 40)]))
    if fam == "synth":
        rgb = orc.synth_rgb(w, h, int(rng.integers(1 << 30)), int(rng.integers(0, 128)))
    elif fam == "noise":
        rgb = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    elif fam == "flat+noise":
        base = rng.integers(0, 256, (1, 1, 3))
        rgb = np.clip(base + rng.integers(-4, 5, (h, w, 3)), 0, 255).astype(np.uint8)
    elif fam == "gradient":
        yy, xx = np.mgrid[0:h, 0:w]
        rgb = np.stack([(xx * int(rng.integers(1, 7)) + yy) % 256, (yy * int(rng.integers(1, 7))) % 256, (xx + yy * 3) % 256], -1).astype(np.uint8)
    elif fam == "extreme":
        rgb = (rng.integers(0, 2, (h, w, 1), dtype=np.uint8) * 255).repeat(3, axis=2)
    else:                                   # mostly smooth with a few busy blocks: strips with and without a second plane
        rgb = orc.synth_rgb(w, h, int(rng.integers(1 << 30)), 3)
        for _ in range(int(rng.integers(0, 6))):
            y0, x0 = int(rng.integers(0, h)), int(rng.integers(0, w))
            rgb[y0:y0 + 8, x0:x0 + 8] = rng.integers(0, 256, rgb[y0:y0 + 8, x0:x0 + 8].shape, dtype=np.uint8)
    want = orc.encode_scan(rgb)
    want_zz = orc.coefficients(rgb)
    for name, e in encs.items():
        if name == "tensor-core":
            e.set_concurrency(int(rng.choice([1, 1, 4, 16])))
        got = e.encode(rgb)
        nb = ((w + 7) // 8) * ((h + 7) // 8)
        zz_bad = int((e.coefficients(nb) != want_zz).sum())
        if got != want or zz_bad:
            bad += 1
            print(f"MISMATCH {name} image {i}: {fam} {w}x{h} scan_equal={got == want} coefficient mismatches={zz_bad}", flush=True)
    px += w * h
    fam_count[fam] = fam_count.get(fam, 0) + 1
print(f"soak: {n_img} images ({px / 1e6:.1f} Mpixel; {fam_count}), both transforms, scan bytes and coefficients vs the oracle: "
      f"{bad} mismatches, {time.time() - t0:.0f} s")
sys.exit(1 if bad else 0)
