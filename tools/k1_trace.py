"""Tuning aid: per-warp timeline of the fused block kernel (run with JPEGB200_K1_TRACE=1 set by this script)."""
import os, sys
os.environ["JPEGB200_K1_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ.setdefault("JPEGB200_LIB", os.path.join(ROOT, "jpeg_image_compression_b200", "libjpegb200_trace.so"))   # make -C .../csrc trace
sys.path.insert(0, ROOT)
import numpy as np, torch
import jpeg_image_compression_b200 as jb
from jpeg_image_compression_b200._lib import check

enc = jb.DeviceEncoder(0)
w, h = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (3840, 2160)
d = enc.synth(w, h, 1, 1, 20)
enc.set_profiling(True)
for _ in range(6):
    enc.encode_device(d, w, h, 1)
torch.cuda.synchronize()
print("kernel times (events):", enc.kernel_times())
strips = ((w + 255) // 256) * ((h + 7) // 8)
import ctypes
grid, wpc = ctypes.c_int(0), ctypes.c_int(0)
check(enc.lib.jpegb200_encoder_launch_shape(enc.handle, ctypes.byref(grid), ctypes.byref(wpc)), "launch_shape")
nwarps = grid.value * wpc.value
print("block kernel grid", grid.value, "x", wpc.value, "warps")
tr = np.zeros((4 * nwarps, 8), np.uint64)
check(enc.lib.jpegb200_encoder_read_k1_trace(enc.handle, tr.ctypes.data, 4 * nwarps), "trace")
t = tr[:nwarps].astype(np.int64)
ends = tr[nwarps:2 * nwarps].astype(np.int64)
detail = tr[2 * nwarps:3 * nwarps].astype(np.int64)
detail2 = tr[3 * nwarps:].astype(np.int64)
t0 = t[:, 0].min()
names = ["entry", "prologue", "tile0", "-", "-", "-", "exit"]
print("warps", nwarps, "strips", strips, "span us", (t[:, 6].max() - t0) / 1e3)
for i, n in enumerate(names):
    col = t[:, i]
    ok = col > 0
    if not ok.any():
        continue
    rel = (col[ok] - t0) / 1e3
    print(f"{n:12s} n={ok.sum():5d} min {rel.min():6.2f} p10 {np.percentile(rel,10):6.2f} median {np.median(rel):6.2f} p90 {np.percentile(rel,90):6.2f} max {rel.max():6.2f}")
nf = t[:, 7]
two_ = np.arange(nwarps) + nwarps < strips
for name, grp in (("two-strip", two_), ("one-strip", ~two_)):
    for lo, hi in ((0, 0), (1, 1), (2, 2), (3, 99)):
        sel = grp & (nf >= lo) & (nf <= hi)
        if sel.any():
            life = (t[sel, 6] - t[sel, 0]) / 1e3
            print(f"{name} warps with {lo}..{hi} exact-path coefficients: n={sel.sum():5d} life median {np.median(life):6.2f} p90 {np.percentile(life, 90):6.2f} max {life.max():6.2f}")
two = np.arange(nwarps) + nwarps < strips
life = (t[:, 6] - t[:, 0]) / 1e3
print(f"warp life: two-strip warps median {np.median(life[two]):.2f}  one-strip warps median {np.median(life[~two]) if (~two).any() else 0:.2f}")
prev = t[:, 2]                                   # arrival of the first tile
for it in range(8):
    col = ends[:, it]
    okk = col > 0
    if not okk.any():
        break
    dur = (col[okk] - prev[okk]) / 1e3
    print(f"strip {it}: n={okk.sum():5d} duration median {np.median(dur):5.2f} p90 {np.percentile(dur, 90):5.2f} max {dur.max():5.2f}  (ends at median {np.median((col[okk] - t0) / 1e3):6.2f})")
    prev = col

# phases of every warp's 4th tile (steady state on large inputs)
ok = (detail > 0).all(axis=1)
if ok.any():
    d = detail[ok]
    names = ["wait for pixels", "luma pass", "prefetch+arrive(+MMA issue)", "statistics", "wait for MMA", "TMEM loads + quantization", "flag pass + stores"]
    print(f"4th tile of {ok.sum()} warps, phase durations in us:")
    for i, n in enumerate(names):
        x = (d[:, i + 1] - d[:, i]) / 1e3
        print(f"  {n:30s} median {np.median(x):5.2f} p10 {np.percentile(x, 10):5.2f} p90 {np.percentile(x, 90):5.2f} mean {x.mean():5.2f}")
    x = (d[:, 7] - d[:, 0]) / 1e3
    print(f"  {'whole tile':30s} median {np.median(x):5.2f} mean {x.mean():5.2f}")

ok2 = ok & (detail2[:, :6] > 0).all(axis=1)
if ok2.any():
    d, e = detail[ok2], detail2[ok2]
    rows = [("luma done -> proxy fence done", e[:, 0] - d[:, 2]), ("tcgen05 fence + syncwarp", e[:, 1] - e[:, 0]), ("arrival (shared atomic)", e[:, 2] - e[:, 1]),
            ("MMA issue (last arriver only)", e[:, 3] - e[:, 2]), ("next strip: position + TMA issue", d[:, 3] - e[:, 3]),
            ("MMA done -> first TMEM chunk loaded", e[:, 4] - d[:, 5]), ("first chunk quantized + second loaded", e[:, 5] - e[:, 4])]
    print("inside the arrival and quantization phases:")
    for n, x in rows:
        x = x / 1e3
        print(f"  {n:40s} median {np.median(x):5.2f} p10 {np.percentile(x, 10):5.2f} p90 {np.percentile(x, 90):5.2f} mean {x.mean():5.2f}")
