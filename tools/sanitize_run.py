"""Small, deterministic workout of every kernel path (single images, batch, files mode, stripes, stage API),
checked against the oracle.  Runs plainly, or under compute-sanitizer where that is available:
    compute-sanitizer --tool memcheck|racecheck|synccheck|initcheck python tools/sanitize_run.py
Checks the results against the oracle as it goes (sizes kept small: the tools slow kernels 10-100x)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import jpeg_image_compression_b200 as jb
from jpeg_image_compression_b200 import stripes
from oracle.oracle import Oracle

orc = Oracle()
enc = jb.DeviceEncoder(0)
rng = np.random.default_rng(7)
ok = True


def check(name, got, want):
    global ok
    good = got == want
    ok &= good
    print(f"{'ok ' if good else 'BAD'} {name}: {len(got)} bytes")


# single images: aligned, ragged, tiny, noise (window overflow -> retry path), many tiles (ticket path)
for (w, h, amp) in ((256, 64, 20), (203, 77, 30), (1, 1, 0), (7, 9, 64), (2048, 24, 20)):
    rgb = orc.synth_rgb(w, h, 3, amp)
    check(f"synth {w}x{h}", enc.encode(rgb), orc.encode_scan(rgb))
noise = rng.integers(0, 256, (48, 80, 3), dtype=np.uint8)
check("noise 80x48 (host entry, workspace retry)", jb.encode_scan_host(noise), orc.encode_scan(noise))
big = orc.synth_rgb(256, 8 * 8 * 80, 5, 20)           # 640 strips = 80 tiles on one 256-px column
check("tall 256x5120", enc.encode(big), orc.encode_scan(big))
# batch + files mode
imgs = np.stack([orc.synth_rgb(120, 50, s, 25) for s in range(5)])
for i, s in enumerate(enc.encode_batch(imgs)):
    check(f"batch[{i}]", s, orc.encode_scan(imgs[i]))
hdr = orc.jfif_header(120, 50)
for i, f in enumerate(enc.encode_batch_files(imgs)):
    check(f"files[{i}]", f, hdr + orc.encode_scan(imgs[i]) + b"\xff\xd9")
# stripes: 3 virtual ranks on one GPU
w, h = 300, 100
rgb = orc.synth_rgb(w, h, 11, 20)
world = 3
encs = [jb.DeviceEncoder(0) for _ in range(world)]
d = torch.from_numpy(rgb).cuda()
parts, scans = [], []
for r in range(world):
    y0, owned, halo = stripes.stripe_rows(h, world, r)
    parts.append(d[y0:y0 + owned + halo].contiguous())
    scans.append(torch.zeros(1 << 16, dtype=torch.uint8, device="cuda"))
check("stripes x3", stripes.encode_striped_local(encs, parts, w, h, scans), orc.encode_scan(rgb))
# stage API (one exact kernel per reference stage)
from jpeg_image_compression_b200 import stages
small = orc.synth_rgb(40, 24, 2, 30)
y = stages.convertBMPToJPEGGrayscale(small)
zz = stages.performZigZag(stages.quantizeImage(stages.performDCT(stages.centerYImage(y))))
sym = stages.performRLE(zz)
bw, bh = (40 + 7) // 8, (24 + 7) // 8
check("stage chain", bytes(stages.encodeHuffman(sym, bw * bh)), orc.encode_scan(small))
torch.cuda.synchronize()
print("ALL OK" if ok else "MISMATCH")
sys.exit(0 if ok else 1)
