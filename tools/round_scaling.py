"""Timing experiment: K1 / K1b / K2 event times versus image size (whole 'rounds' of strips per warp slot)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import jpeg_image_compression_b200 as jb

enc = jb.DeviceEncoder(0)
for (w, h) in [(2048, 1184), (2048, 2368), (2048, 4736), (2048, 9472), (2048, 18944), (3840, 2160), (7680, 4320)]:
    ring = [enc.synth(w, h, 1, s, 20) for s in range(1, 5)]
    for i in range(8):
        enc.encode_device(ring[i % 4], w, h, 1)
    torch.cuda.synchronize()
    enc.set_profiling(True); enc.kernel_times(reset=True)
    for i in range(32):
        enc.encode_device(ring[i % 4], w, h, 1)
    t = enc.kernel_times(reset=True); enc.set_profiling(False)
    strips = ((w + 255) // 256) * ((h + 7) // 8)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); ev0.record()
    for i in range(64):
        enc.encode_device(ring[i % 4], w, h, 1)
    ev1.record(); torch.cuda.synchronize()
    print(f"{w}x{h}: strips={strips} rounds={strips/2368:.2f} K1={1e3*t['ms'][0]/32:.1f}us K1b={1e3*t['ms'][5]/32:.1f}us K2={1e3*t['ms'][1]/32:.1f}us "
          f"stream_step={1e3*ev0.elapsed_time(ev1)/64:.1f}us  {w*h/1e6/(ev0.elapsed_time(ev1)/64*1e-3)/1e3:.0f} Gpx/s", flush=True)
