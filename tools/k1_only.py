"""Tuning aid: run the encode a few times on one large image (steady state for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import jpeg_image_compression_b200 as jb
enc = jb.DeviceEncoder(0)
w, h = 7680, 4320
d = enc.synth(w, h, 1, 1, 20)
for _ in range(6):
    enc.encode_device(d, w, h, 1)
torch.cuda.synchronize()
enc.status()
print("ok")
