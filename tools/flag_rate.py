"""Tuning aid: how many coefficients take the reference-order re-evaluation path."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import jpeg_image_compression_b200 as jb
enc = jb.DeviceEncoder(0)
for (w, h, seed, amp) in [(3840, 2160, 1, 20), (3840, 2160, 1, 0), (3840, 2160, 1, 64), (1920, 1080, 0, 20)]:
    d = enc.synth(w, h, 1, seed, amp)
    enc.stats()
    enc.encode_device(d, w, h, 1)
    enc.status()
    st = enc.stats()
    print(f"{w}x{h} amp={amp}: blocks={st['blocks']} flagged={st['flagged_coefficients']} "
          f"per strip={st['flagged_coefficients'] / (st['blocks'] / 32):.2f} per coefficient={st['flagged_coefficients'] / (63 * st['blocks']):.2e}")
