"""CPU tests for the multi-GPU host logic: batch sharding, stripe planning, offset resolution,
and the full striped encode + stitch over torch.distributed (gloo, world_size 2..4) with the
oracle standing in for the CUDA backend.  Output must be byte-identical to the unsharded encode."""
import os
import socket

import numpy as np
import pytest

from jpeg_image_compression_b200.stripes import (StripedEncoder, dc_cost, encode_striped_local, resolve_offsets, shard_range,
                                                 stripe_plan, stripe_rows)
from stripe_backend import OracleStripeBackend


def test_shard_range_covers_everything():
    for n in (0, 1, 7, 8, 4096, 4099):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1
    assert shard_range(4096, 8, 3) == (1536, 2048)


def test_stripe_plan_matches_survey():
    assert [n for _, n in stripe_plan(4320, 8)] == [68, 68, 68, 68, 67, 67, 67, 67]
    assert [n for _, n in stripe_plan(4320, 2)] == [270, 270]
    assert [n for _, n in stripe_plan(32768, 8)] == [512] * 8
    assert stripe_rows(4320, 2, 0) == (0, 2160, 8) and stripe_rows(4320, 2, 1) == (2160, 2160, 0)
    assert stripe_rows(21, 2, 0) == (0, 16, 5) and stripe_rows(21, 2, 1) == (16, 5, 0)      # ragged bottom
    assert stripe_rows(9, 4, 2)[1:] == (0, 0) and stripe_rows(9, 4, 3)[1:] == (0, 0)          # more ranks than block rows


def test_dc_cost_table():
    assert [dc_cost(d) for d in (0, 1, -1, 2, 3, 4, 15, 16, 63, 64, 127, 255)] == [2, 4, 4, 5, 5, 6, 7, 8, 10, 12, 12, 14]


def test_resolve_offsets_chain():
    s = [{"first_dc": 5, "last_dc": -3, "bits_pred0": 100}, None, {"first_dc": 7, "last_dc": 9, "bits_pred0": 50}]
    plan = resolve_offsets(s)
    assert plan[0] == (0, 0) and plan[1] is None
    assert plan[2] == (-3, 100)                               # predictor = previous stripe's last DC
    s.append({"first_dc": 9, "last_dc": 0, "bits_pred0": 10})
    assert resolve_offsets(s)[3] == (9, 100 + 50 - dc_cost(7) + dc_cost(7 - (-3)))


@pytest.mark.parametrize("w,h,world,kind", [(64, 64, 2, "synth"), (70, 45, 2, "noise"), (33, 100, 3, "synth"),
                                            (200, 37, 4, "synth"), (16, 9, 4, "noise"), (8, 8, 2, "flat"),
                                            (40, 24, 3, "flat"), (257, 19, 2, "noise")])
def test_striped_local_equals_unsharded(oracle, w, h, world, kind):
    rng = np.random.default_rng(w * 1000 + h)
    if kind == "synth":
        rgb = oracle.synth_rgb(w, h, 3, 25)
    elif kind == "noise":
        rgb = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    else:
        rgb = np.full((h, w, 3), 77, np.uint8)                # 6-bit blocks: many stripes inside one byte
    backends = [OracleStripeBackend(oracle) for _ in range(world)]
    stripes, scans = [], []
    for r in range(world):
        y0, owned, halo = stripe_rows(h, world, r)
        stripes.append(rgb[y0:y0 + owned + halo])
        scans.append(np.zeros(4 * w * max(owned, 1) + 64, np.uint8))
    got = encode_striped_local(backends, stripes, w, h, scans)
    assert got == oracle.encode_scan(rgb)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _rank_main(rank, world, port, w, h, seed, result_path):
    import torch.distributed as dist
    from oracle.oracle import Oracle
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        oracle = Oracle()
        rgb = oracle.synth_rgb(w, h, seed, 30)               # every rank can regenerate the image; it reads only its rows
        y0, owned, halo = stripe_rows(h, world, rank)
        enc = StripedEncoder(OracleStripeBackend(oracle), device="cpu")
        scan = np.zeros(4 * w * max(owned, 1) + 64, np.uint8)
        n = enc.encode(rgb[y0:y0 + owned + halo], w, h, scan)
        stitched = enc.gather(scan, n)
        # the size-exact gather (sizes all-gathered as tensors, then send / recv of exactly each rank's bytes)
        import torch
        exact, total = enc.gather_exact(torch.from_numpy(scan), torch.tensor(n, dtype=torch.int64))
        if rank == 0:
            ok = stitched == oracle.encode_scan(rgb) and exact[:total].numpy().tobytes() == stitched
            with open(result_path, "w") as f:
                f.write("ok" if ok else f"mismatch: {len(stitched)} bytes")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,w,h", [(2, 96, 80), (3, 50, 61), (4, 24, 20)])
def test_striped_over_gloo(tmp_path, world, w, h):
    import torch.multiprocessing as mp
    result = str(tmp_path / "result.txt")
    mp.spawn(_rank_main, args=(world, _free_port(), w, h, 11, result), nprocs=world, join=True)
    assert open(result).read() == "ok"
