"""Encode one synthetic w x h image n times on cuda:0 (a small, fixed command line for ncu captures).
usage: python tools/encode_once.py W H N"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import jpeg_image_compression_b200 as jb

w, h, n = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
enc = jb.DeviceEncoder(0)
d = enc.synth(w, h, 1, 1, 20)
for _ in range(n):
    scan, offs = enc.encode_device(d, w, h, 1)
torch.cuda.synchronize()
enc.status()
print(w, h, "scan bytes", int(offs[1].item()))
