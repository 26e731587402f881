import os, subprocess, sys, tempfile
sys.path.insert(0, os.getcwd())
from oracle.oracle import Oracle, write_bmp
orc = Oracle()
td = tempfile.mkdtemp()
ins, want = [], {}
for i in range(12):
    w, h = 640 + 16 * i, 360 + 8 * i
    rgb = orc.synth_rgb(w, h, 100 + i, 20)
    p = os.path.join(td, f"img{i}.bmp")
    write_bmp(p, rgb)
    ins.append(p)
    want[f"img{i}.jpg"] = orc.encode_file_bytes(rgb)
out = os.path.join(td, "out"); os.mkdir(out)
r = subprocess.run(["jpeg_image_compression_b200/jpeg_compression_batch", out] + ins, capture_output=True, text=True)
print(r.stdout.strip(), r.stderr.strip()[-200:])
ok = all(open(os.path.join(out, k), "rb").read() == v for k, v in want.items())
print("batch CLI on all GPUs:", "byte-identical to the oracle" if ok and r.returncode == 0 else "MISMATCH")
