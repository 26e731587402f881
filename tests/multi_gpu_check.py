#!/usr/bin/env python
"""Multi-GPU parity check (run under torchrun on 2/4/8 B200s; not collected by pytest):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 tests/multi_gpu_check.py

Each rank encodes its MCU-row stripe of a synthetic 7680x4320 image (BASELINE configs[2]); the
stitched stream on rank 0 must hash to the reference build's output.  Then a batch of 1080p
images is sharded by image and checked against the reference hashes of images 0 and 64."""
import hashlib
import json
import os
import sys

if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
    os.environ["NCCL_DEBUG"] = "WARN"
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import jpeg_image_compression_b200 as jb  # noqa: E402
from jpeg_image_compression_b200.stripes import StripedEncoder, shard_range, stripe_rows  # noqa: E402


def gigapixel(enc, se, rank, world, dev):
    """BASELINE configs[4]: synthetic 32768x32768 image, MCU-row stripes over all ranks.  The stock
    reference cannot load it (int overflow), so the striped result is checked against
    the single-GPU encode of the same image (rank 0) and the first block rows against the oracle."""
    import time
    import numpy as np
    from oracle.oracle import Oracle
    w = h = 32768
    y0, owned, halo = stripe_rows(h, world, rank)
    full = enc.synth(w, h, 1, 1, 20)[0]
    stripe = full[y0:y0 + owned + halo].contiguous()
    scan = torch.empty(enc.scan_capacity(w, max(owned, 8), 1), dtype=torch.uint8, device=dev)
    se.encode(stripe, w, h, scan)                                   # warm-up: sizes the workspace
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    n = se.encode(stripe, w, h, scan)
    torch.cuda.synchronize()
    dist.barrier()
    dt = time.perf_counter() - t0
    stitched = se.gather(scan, n)
    good = True
    if rank == 0:
        del stripe
        solo = jb.DeviceEncoder(dev.index)
        out = torch.empty(solo.scan_capacity(w, h, 1), dtype=torch.uint8, device=dev)
        offs = torch.zeros(2, dtype=torch.int64, device=dev)
        solo.encode_device(full, w, h, 1, scan=out, offsets=offs)
        solo.status()
        t1 = time.perf_counter()
        solo.encode_device(full, w, h, 1, scan=out, offsets=offs)
        solo.status()
        dt1 = time.perf_counter() - t1
        m = int(offs[1].item())
        single = out[:m].cpu().numpy().tobytes()
        good = single == stitched
        # first 8 block rows of coefficients against the CPU oracle (coefficients are position independent)
        orc = Oracle()
        head = full[:64].cpu().numpy()
        nb = (w // 8) * 8
        good &= bool(np.array_equal(solo.coefficients(nb), orc.coefficients(head)))
        print(f"gigapixel 32768x32768 over {world} GPUs: {len(stitched)} scan bytes, "
              f"{'identical to the single-GPU encode, first rows identical to the oracle' if good else 'MISMATCH'}; "
              f"striped {dt * 1e3:.2f} ms ({w * h / dt / 1e9:.1f} Gpixel/s incl. exchange), single GPU {dt1 * 1e3:.2f} ms "
              f"({w * h / dt1 / 1e9:.1f} Gpixel/s)", flush=True)
        solo.close()
    return good


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    rank, world = dist.get_rank(), dist.get_world_size()
    hashes = json.load(open(os.path.join(ROOT, "tests", "golden", "synth_hashes.json")))
    enc = jb.DeviceEncoder(local)
    ok = True

    e = hashes["7680x4320_seed1_amp20"]
    w, h = e["w"], e["h"]
    y0, owned, halo = stripe_rows(h, world, rank)
    # every rank generates only its own rows (+ halo): the generator is a pure function of (x, y)
    full = enc.synth(w, h, 1, e["seed"], e["amp"])[0]          # simple: generate, then keep the stripe
    stripe = full[y0:y0 + owned + halo].contiguous()
    del full
    se = StripedEncoder(enc, device=dev)
    scan = torch.empty(enc.scan_capacity(w, max(owned, 8), 1), dtype=torch.uint8, device=dev)
    n = se.encode(stripe, w, h, scan)
    stitched = se.gather(scan, n)
    if rank == 0:
        good = len(stitched) == e["scan_bytes"] and hashlib.sha256(stitched).hexdigest() == e["scan_sha256"]
        print(f"stripes 7680x4320 over {world} GPUs: {'byte-identical to reference' if good else 'MISMATCH'}", flush=True)
        ok &= good

    # batch sharded by image: image i uses seed i; ranks own contiguous ranges
    total = 128
    b, en = shard_range(total, world, rank)
    d = enc.synth(1920, 1080, en - b, b, 20)
    s, o = enc.encode_device(d, 1920, 1080, en - b)
    enc.status()
    o = o[: en - b + 1].cpu().numpy()
    for i in (0, 64):
        if b <= i < en:
            ref = hashes[f"1920x1080_seed{i}_amp20"]
            data = s[int(o[i - b]): int(o[i - b + 1])].cpu().numpy().tobytes()
            good = len(data) == ref["scan_bytes"] and hashlib.sha256(data).hexdigest() == ref["scan_sha256"]
            print(f"batch image {i} on rank {rank}: {'byte-identical to reference' if good else 'MISMATCH'}", flush=True)
            ok &= good
    if "--giga" in sys.argv:
        ok &= gigapixel(enc, se, rank, world, dev)
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    enc.close()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
