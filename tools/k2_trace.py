"""Tuning aid: phase timeline of the fused entropy kernel (run with JPEGB200_K2_TRACE=1)."""
import os, sys
os.environ["JPEGB200_K2_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ.setdefault("JPEGB200_LIB", os.path.join(ROOT, "jpeg_image_compression_b200", "libjpegb200_trace.so"))   # make -C .../csrc trace
sys.path.insert(0, ROOT)
import numpy as np, torch
import jpeg_image_compression_b200 as jb
from jpeg_image_compression_b200._lib import check

enc = jb.DeviceEncoder(0)
w, h = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (3840, 2160)
d = enc.synth(w, h, 1, 1, 20)
for _ in range(5):
    enc.encode_device(d, w, h, 1)
torch.cuda.synchronize()
ntiles = (((w + 255) // 256) * ((h + 7) // 8) + 15) // 16      # K2_TILE_STRIPS = 16
tr = np.zeros((ntiles, 8), np.uint64)
check(enc.lib.jpegb200_encoder_read_trace(enc.handle, tr.ctypes.data, ntiles), "trace")
t = tr.astype(np.int64)
t0 = t[:, 0].min()
names = ["start", "prefix", "offsets", "(unused)", "assembled", "ffcount", "fflook", "written"]
print("tiles", ntiles, "kernel span us", (t[:, 7].max() - t0) / 1e3)
for i, n in enumerate(names):
    rel = (t[:, i] - t0) / 1e3
    print(f"{n:10s} min {rel.min():6.2f} median {np.median(rel):6.2f} max {rel.max():6.2f}")
t[:, 3] = t[:, 2]
for i in range(1, 8):
    d_ = (t[:, i] - t[:, i - 1]) / 1e3
    print(f"phase {names[i-1]:>9s}->{names[i]:<9s}: median {np.median(d_):5.2f} p90 {np.percentile(d_, 90):5.2f} max {d_.max():5.2f}")
span = (t[:, 7] - t[:, 0]) / 1e3
print(f"tile life: median {np.median(span):5.2f} p90 {np.percentile(span, 90):5.2f} max {span.max():5.2f}")
order = np.argsort(t[:, 0])
print("start order == tile order:", bool((order == np.arange(ntiles)).all()))
