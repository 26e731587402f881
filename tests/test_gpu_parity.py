"""GPU parity tests (run on the B200 box): the CUDA path, called through the C ABI, against
  (1) the committed golden fixtures minted from the reference build,
  (2) the oracle on fresh seeded inputs,
  (3) SHA-256 of reference outputs on full-size synthetic images,
bit-exact at every tap: quantized zig-zag coefficients, per-block bit costs, scan bytes, file."""
import hashlib
import os
import subprocess

import numpy as np
import pytest

import jpeg_image_compression_b200 as jb

pytestmark = pytest.mark.gpu
ZIGZAG = np.array([0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6,
                   7, 14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38,
                   31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63])
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def enc():
    e = jb.DeviceEncoder(0)
    yield e
    e.close()


@pytest.fixture(scope="module")
def enc_exact():
    e = jb.DeviceEncoder(0, dct_mode=1)
    yield e
    e.close()


def _nblocks(rgb):
    h, w, _ = rgb.shape
    return ((h + 7) // 8) * ((w + 7) // 8)


def test_golden_scan_coefficients_bits(enc, oracle, golden, golden_names):
    for name in golden_names:
        rgb = golden[f"{name}/rgb"]
        scan = enc.encode(rgb)
        nb = _nblocks(rgb)
        zz = enc.coefficients(nb)
        mism = int((zz != golden[f"{name}/zigzag"]).sum())
        assert mism == 0, f"{name}: {mism} coefficient mismatches"
        assert np.array_equal(enc.block_bits(nb), oracle.block_bits(golden[f"{name}/zigzag"])), name
        assert scan == golden[f"{name}/scan"].tobytes(), name


def test_golden_exact_mode(enc_exact, golden, golden_names):
    enc_exact.stats()                                   # read-and-reset the flagged counter
    for name in golden_names:
        rgb = golden[f"{name}/rgb"]
        assert enc_exact.encode(rgb) == golden[f"{name}/scan"].tobytes(), name
        assert np.array_equal(enc_exact.coefficients(_nblocks(rgb)), golden[f"{name}/zigzag"]), name
        st = enc_exact.stats()
        assert st["flagged_coefficients"] == 63 * st["blocks"], name


def test_golden_fused_host_entry_and_file(golden, golden_names, tmp_path):
    for name in golden_names:
        rgb = golden[f"{name}/rgb"]
        scan, first = jb.encode_scan_host(rgb, with_first_block=True)
        assert scan == golden[f"{name}/scan"].tobytes(), name
        raster = np.zeros(64, np.int16)
        raster[ZIGZAG] = golden[f"{name}/zigzag"][0]
        assert np.array_equal(first.reshape(64), raster), name
        out = str(tmp_path / f"{name}.jpg")
        assert jb.saveJPEGGrayscale(out, rgb)
        assert open(out, "rb").read() == golden[f"{name}/file"].tobytes(), name


def test_random_sizes_vs_oracle(enc, oracle):
    rng = np.random.default_rng(2026)
    dims = [(1, 1), (2, 9), (8, 8), (9, 9), (255, 8), (256, 8), (257, 8), (263, 17), (511, 33), (762, 40),
            (1280, 24), (1283, 9), (64, 300), (5, 1031)]
    for (w, h) in dims:
        for kind in range(3):
            if kind == 0:
                rgb = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
            elif kind == 1:
                rgb = oracle.synth_rgb(w, h, int(rng.integers(1 << 31)), int(rng.integers(0, 64)))
            else:
                base = rng.integers(0, 256, (1, 1, 3), dtype=np.uint8)
                rgb = np.clip(base.astype(np.int32) + rng.integers(-3, 4, (h, w, 3)), 0, 255).astype(np.uint8)
            scan = enc.encode(rgb)
            zz = enc.coefficients(_nblocks(rgb))
            ref_zz = oracle.coefficients(rgb)
            assert int((zz != ref_zz).sum()) == 0, (w, h, kind)
            assert scan == oracle.encode_scan(rgb), (w, h, kind)


def test_extreme_blocks(enc, oracle):
    """Saturated basis-function sign patterns (largest possible coefficients), flats, ties."""
    blocks = []
    r = np.arange(8)
    for u in range(8):
        for v in range(8):
            pat = np.sign(np.outer(np.cos((2 * r + 1) * u * np.pi / 16), np.cos((2 * r + 1) * v * np.pi / 16)))
            blocks.append(np.where(pat >= 0, 255, 0))
            blocks.append(np.where(pat >= 0, 0, 255))
    for v in (0, 1, 127, 128, 129, 254, 255):
        blocks.append(np.full((8, 8), v))
    img = np.concatenate([np.concatenate(blocks[i:i + 15], 1) for i in range(0, 135, 15)], 0).astype(np.uint8)
    rgb = np.repeat(img[:, :, None], 3, 2)
    assert enc.encode(rgb) == oracle.encode_scan(rgb)
    assert np.array_equal(enc.coefficients(_nblocks(rgb)), oracle.coefficients(rgb))


def test_fast_mode_equals_exact_mode_soak(enc, enc_exact):
    """Randomised soak: guard-band path == reference-order path, coefficient for coefficient."""
    total_flagged = total_blocks = 0
    enc.stats()
    for seed, amp in [(11, 5), (12, 20), (13, 40), (14, 64), (15, 127)]:
        d = enc.synth(1024, 1024, 1, seed, amp)
        s1, o1 = enc.encode_device(d, 1024, 1024, 1)
        enc.status()
        a = enc.coefficients(128 * 128)
        n1 = int(o1[1].item())
        b1 = s1[:n1].cpu().numpy().tobytes()
        st = enc.stats()
        total_flagged += st["flagged_coefficients"]
        total_blocks += st["blocks"]
        s2, o2 = enc_exact.encode_device(d, 1024, 1024, 1)
        enc_exact.status()
        b = enc_exact.coefficients(128 * 128)
        assert int((a != b).sum()) == 0, (seed, amp)
        assert b1 == s2[: int(o2[1].item())].cpu().numpy().tobytes()
    assert total_flagged < 0.01 * 63 * total_blocks          # the fallback must stay rare


def test_encode_host_pinned_buffers(enc, oracle):
    import torch
    rgb = oracle.synth_rgb(640, 360, 21, 20)
    pin_in = torch.from_numpy(rgb).pin_memory()
    pin_out = torch.empty(enc.scan_capacity(640, 360, 1), dtype=torch.uint8).pin_memory()
    n = enc.encode_host(pin_in, 640, 360, pin_out)
    assert pin_out[:n].numpy().tobytes() == oracle.encode_scan(rgb)
    # pageable numpy buffers take the same path
    out = np.empty(enc.scan_capacity(640, 360, 1), np.uint8)
    n = enc.encode_host(rgb, 640, 360, out)
    assert out[:n].tobytes() == oracle.encode_scan(rgb)


def test_kernel_timing_taps(enc):
    d = enc.synth(1024, 512, 1, 3, 20)
    enc.set_profiling(True)
    enc.kernel_times(reset=True)
    for _ in range(3):
        enc.encode_device(d, 1024, 512, 1)
    t = enc.kernel_times(reset=True)
    enc.set_profiling(False)
    assert t["calls"][:2] == [3, 3] and all(ms > 0 for ms in t["ms"][:2])


def test_cuda_graph_replay_is_correct(enc, oracle):
    """The three-kernel encode is graph-capturable: the merge kernel's look-back state is reset by K1 itself."""
    import torch
    rgbs = [oracle.synth_rgb(512, 256, s, 25) for s in (1, 2)]
    d = [torch.from_numpy(r).cuda() for r in rgbs]
    cap = enc.scan_capacity(512, 256, 1)
    outs = [(torch.empty(cap, dtype=torch.uint8, device="cuda"), torch.zeros(2, dtype=torch.int64, device="cuda")) for _ in d]
    for di, (s, o) in zip(d, outs):                       # warm-up: allocates the workspace
        enc.encode_device(di, 512, 256, 1, scan=s, offsets=o)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for di, (s, o) in zip(d, outs):
            enc.encode_device(di, 512, 256, 1, scan=s, offsets=o)
    for _ in range(3):
        for s, o in outs:
            s.zero_(); o.zero_()
        g.replay()
        torch.cuda.synchronize()
        enc.status()
        for r, (s, o) in zip(rgbs, outs):
            assert s[: int(o[1].item())].cpu().numpy().tobytes() == oracle.encode_scan(r)


def test_device_synth_equals_host_synth(enc):
    d = enc.synth(333, 77, 3, seed0=41, amp=20).cpu().numpy()
    for i in range(3):
        assert np.array_equal(d[i], jb.synth_rgb(333, 77, 41 + i, 20))


def test_full_size_reference_hashes(enc, synth_hashes):
    for key in ["1920x1080_seed0_amp20", "3840x2160_seed1_amp20", "3840x2160_seed1_amp0", "3840x2160_seed1_amp64",
                "762x1309_seed5_amp20", "1283x725_seed9_amp20", "7680x4320_seed1_amp20"]:
        e = synth_hashes[key]
        d = enc.synth(e["w"], e["h"], 1, e["seed"], e["amp"])
        assert hashlib.sha256(d.cpu().numpy().tobytes()).hexdigest() == e["rgb_sha256"], key
        scan, offs = enc.encode_device(d, e["w"], e["h"], 1)
        enc.status()
        n = int(offs[1].item())
        assert n == e["scan_bytes"], key
        assert hashlib.sha256(scan[:n].cpu().numpy().tobytes()).hexdigest() == e["scan_sha256"], key
        nb = ((e["w"] + 7) // 8) * ((e["h"] + 7) // 8)
        assert hashlib.sha256(enc.coefficients(nb).tobytes()).hexdigest() == e["zigzag_sha256"], key


def test_batch_equals_single(enc, oracle):
    imgs = np.stack([oracle.synth_rgb(200, 120, s, 20) for s in range(7)])
    scans = enc.encode_batch(imgs)
    for i in range(7):
        assert scans[i] == oracle.encode_scan(imgs[i]), i
    # ragged geometry + noise in a batch
    rng = np.random.default_rng(4)
    imgs = rng.integers(0, 256, (5, 37, 51, 3), dtype=np.uint8)
    scans = enc.encode_batch(imgs)
    for i in range(5):
        assert scans[i] == oracle.encode_scan(imgs[i]), i


def test_batch_1080p_hashes(enc, synth_hashes):
    d = enc.synth(1920, 1080, 66, 0, 20)
    scan, offs = enc.encode_device(d, 1920, 1080, 66)
    enc.status()
    offs = offs[:67].cpu().numpy()
    for i in (0, 64):
        e = synth_hashes[f"1920x1080_seed{i}_amp20"]
        b = scan[int(offs[i]): int(offs[i + 1])].cpu().numpy().tobytes()
        assert len(b) == e["scan_bytes"] and hashlib.sha256(b).hexdigest() == e["scan_sha256"], i


def test_workspace_overflow_is_reported_not_silent(oracle):
    import torch
    e = jb.DeviceEncoder(0, bytes_per_block=2)
    rng = np.random.default_rng(1)
    rgb = rng.integers(0, 256, (64, 64, 3), dtype=np.uint8)
    # the device API never hides an overflow: the launch succeeds, the status call reports it
    e.encode_device(torch.from_numpy(rgb).cuda(), 64, 64, 1)
    with pytest.raises(jb.JpegB200Error):
        e.status()
    # the convenience entries (Python encode / encode_batch, C jpegb200_encode_scan, the batch CLI) retry with the
    # worst-case workspace and still match
    assert e.encode(rgb) == oracle.encode_scan(rgb)
    assert e.bytes_per_block == 184
    e.close()
    assert jb.encode_scan_host(rgb) == oracle.encode_scan(rgb)


def test_stage_api_against_goldens(golden, golden_names, oracle):
    from oracle.oracle import SYMBOL_DTYPE as ORC_SYM
    for name in golden_names:
        rgb = golden[f"{name}/rgb"]
        y = jb.convertBMPToJPEGGrayscale(rgb)
        assert np.array_equal(y, golden[f"{name}/y"]), name
        c = jb.centerYImage(y)
        assert np.array_equal(c, oracle.level_shift(golden[f"{name}/y"])), name
        d = jb.performDCT(c)
        assert np.array_equal(d.view(np.uint32), golden[f"{name}/dct"].view(np.uint32)), name
        q = jb.quantizeImage(d)
        z = jb.performZigZag(q)
        assert np.array_equal(z, golden[f"{name}/zigzag"]), name
        s = jb.performRLE(z)
        g = golden[f"{name}/symbols"].astype(np.uint16)
        assert np.array_equal(s["symbol"], g[:, 0]) and np.array_equal(s["code"], g[:, 1] | (g[:, 2] << 8)) \
            and np.array_equal(s["codeBits"], g[:, 3]), name
        assert jb.encodeHuffman(s, z.shape[0]) == golden[f"{name}/scan"].tobytes(), name
    blk = np.arange(-32, 32, dtype=np.int8).reshape(8, 8)
    assert np.array_equal(jb.computeDCTBlock(blk).view(np.uint32), oracle.fdct_block(blk).view(np.uint32))


def test_stage_api_null_and_block_cap(oracle):
    assert jb.convertBMPToJPEGGrayscale(None) is None and jb.performDCT(None) is None
    rgb = oracle.synth_rgb(96, 64, 9, 30)
    zz = oracle.coefficients(rgb)
    sym = jb.performRLE(zz)
    # totalBlocks smaller than the symbol stream: the reference stops after that many blocks
    for nb in (1, 5, zz.shape[0]):
        o_sym = oracle.rle(zz)
        assert jb.encodeHuffman(sym, nb) == oracle.huffman(o_sym, nb).tobytes(), nb


def test_cli_end_to_end(golden, tmp_path, ref):
    from oracle.oracle import REF_APP, write_bmp
    app = os.path.join(ROOT, "jpeg_image_compression_b200", "jpeg_compression_app")
    for name in ["lena_crop256", "greenland_corner250x205", "one_pixel", "noise64", "checker_nyquist"]:   # the last two: block 0 reaches beyond zig-zag position 31 (second coefficient plane in the first-block print)
        rgb = golden[f"{name}/rgb"]
        bmp, out, out_ref = (str(tmp_path / f"{name}{s}") for s in (".bmp", ".jpg", "_ref.jpg"))
        write_bmp(bmp, rgb)
        r = subprocess.run([app, bmp, out], capture_output=True, text=True)
        assert r.returncode == 0 and r.stdout.endswith("Save is sucesfull")
        assert open(out, "rb").read() == golden[f"{name}/file"].tobytes(), name
        if os.path.exists(REF_APP):                     # identical stdout chatter, file name aside
            rr = subprocess.run([REF_APP, bmp, out_ref], capture_output=True, text=True)
            assert rr.stdout.replace(out_ref, "X") == r.stdout.replace(out, "X")
            assert open(out_ref, "rb").read() == open(out, "rb").read()


def test_reference_orchestrator_over_library(golden, tmp_path):
    """INTEGRATION.md option A, executed: the reference's UNMODIFIED main.c + src/io/bmp_handler.c + src/io/jpeg_handler.c
    (saveJPEGGrayscale, natural_c/src/io/jpeg_handler.c:119-282) linked against libjpegb200.so instead of its own
    src/core/*.c -- the seven stage calls at jpeg_handler.c:133-201 run as CUDA kernels.  File bytes and stdout must
    equal those of the reference binary."""
    from oracle.oracle import REF_APP, write_bmp
    app = os.path.join(ROOT, "oracle", "_ref", "jpeg_compression_app_optA")
    if not os.path.exists(app):
        pytest.skip("oracle/_ref/jpeg_compression_app_optA not built (needs /root/reference at build time)")
    for name in ["lena_crop256", "greenland_corner250x205", "one_pixel", "noise64", "checker_nyquist"]:   # the last two: block 0 reaches beyond zig-zag position 31 (second coefficient plane in the first-block print)
        rgb = golden[f"{name}/rgb"]
        bmp, out, out_ref = (str(tmp_path / f"{name}{s}") for s in (".bmp", "_a.jpg", "_ref.jpg"))
        write_bmp(bmp, rgb)
        r = subprocess.run([app, bmp, out], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        assert open(out, "rb").read() == golden[f"{name}/file"].tobytes(), name
        if os.path.exists(REF_APP):
            rr = subprocess.run([REF_APP, bmp, out_ref], capture_output=True, text=True)
            assert rr.stdout.replace(out_ref, "X") == r.stdout.replace(out, "X"), name
    # the library in use really is ours: the binary has no core stage of its own
    syms = subprocess.run(["nm", "-D", "--undefined-only", app], capture_output=True, text=True).stdout
    for fn in ("convertBMPToJPEGGrayscale", "centerYImage", "performDCT", "quantizeImage", "performZigZag", "performRLE", "encodeHuffman"):
        assert fn in syms, fn


def test_full_size_assets(enc, synth_hashes, tmp_path):
    """BASELINE configs[0]: the reference's own assets/input/*.bmp at full size (copied next to oracle/_ref by build();
    not committed), through the BMP loader + fused path and through the in-place BMP ingest, against the hashes minted
    from the reference build."""
    adir = os.path.join(ROOT, "oracle", "_ref", "assets")
    names = [n for n in ("lena", "blackbuck", "greenland", "offset_sample") if os.path.exists(os.path.join(adir, n + ".bmp"))]
    if not names:
        pytest.skip("asset BMPs not present (oracle/_ref/assets)")
    app = os.path.join(ROOT, "jpeg_image_compression_b200", "jpeg_compression_app")
    for name in names:
        e = synth_hashes["asset_" + name]
        path = os.path.join(adir, name + ".bmp")
        out = str(tmp_path / (name + ".jpg"))
        r = subprocess.run([app, path, out], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        data = open(out, "rb").read()
        scan = data[328:-2]
        assert len(scan) == e["scan_bytes"] and hashlib.sha256(scan).hexdigest() == e["scan_sha256"], name
        assert enc.encode_bmp_to_jpeg(open(path, "rb").read()) == data, name


# ---- MCU-row stripes (several ranks emulated on one GPU: one encoder handle per rank) ------------

def _striped_on_one_gpu(encoders, rgb_t, w, h):
    import torch
    from jpeg_image_compression_b200.stripes import encode_striped_local, stripe_rows
    world = len(encoders)
    stripes, scans = [], []
    for r in range(world):
        y0, owned, halo = stripe_rows(h, world, r)
        stripes.append(rgb_t[y0:y0 + owned + halo].contiguous())
        scans.append(torch.empty(encoders[0].scan_capacity(w, max(owned, 8), 1), dtype=torch.uint8, device="cuda"))
    return encode_striped_local(encoders, stripes, w, h, scans)


@pytest.fixture(scope="module")
def rank_encoders():
    es = [jb.DeviceEncoder(0) for _ in range(8)]
    yield es
    for e in es:
        e.close()


def test_stripes_equal_unsharded_small(rank_encoders, oracle):
    import torch
    rng = np.random.default_rng(8)
    cases = [(64, 64, 2, "synth"), (70, 45, 2, "noise"), (33, 100, 3, "synth"), (200, 37, 4, "synth"), (16, 9, 4, "noise"),
             (8, 8, 2, "flat"), (40, 24, 3, "flat"), (257, 19, 2, "noise"), (1283, 725, 8, "synth"), (300, 300, 5, "flat")]
    for (w, h, world, kind) in cases:
        if kind == "synth":
            rgb = oracle.synth_rgb(w, h, 3, 25)
        elif kind == "noise":
            rgb = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        else:
            rgb = np.full((h, w, 3), 77, np.uint8)            # 6-bit blocks: stripe boundaries inside bytes
        got = _striped_on_one_gpu(rank_encoders[:world], torch.from_numpy(rgb).cuda(), w, h)
        assert got == oracle.encode_scan(rgb), (w, h, world, kind)


def test_stripes_8k_hash(rank_encoders, enc, synth_hashes):
    e = synth_hashes["7680x4320_seed1_amp20"]
    d = enc.synth(e["w"], e["h"], 1, e["seed"], e["amp"])[0]
    for world in (2, 4, 8):
        got = _striped_on_one_gpu(rank_encoders[:world], d, e["w"], e["h"])
        assert len(got) == e["scan_bytes"] and hashlib.sha256(got).hexdigest() == e["scan_sha256"], world


def _striped_on_one_gpu_device_path(encoders, rgb_t, w, h):
    """The device-resident stripe path: summaries reduced and resolved by kernels, nothing returns to the host
    between analyze, the (here: emulated) all-gather and the merge."""
    import torch
    from jpeg_image_compression_b200.stripes import stripe_rows
    world = len(encoders)
    summaries = torch.zeros((world, 2), dtype=torch.int64, device="cuda")      # the all-gather buffer
    infos = torch.zeros((world, 2), dtype=torch.int64, device="cuda")
    stripes, scans = [], []
    for r in range(world):
        y0, owned, halo = stripe_rows(h, world, r)
        stripes.append(rgb_t[y0:y0 + owned + halo].contiguous())
        scans.append(torch.empty(encoders[0].scan_capacity(w, max(owned, 8), 1), dtype=torch.uint8, device="cuda"))
        if owned:
            encoders[r].stripe_analyze_device(stripes[r], w, owned, halo, summaries[r])
    for r in range(world):
        if stripe_rows(h, world, r)[1]:
            encoders[r].stripe_encode_device(summaries, world, r, scans[r], infos[r])
    torch.cuda.synchronize()
    for e in encoders:
        e.status()
    sizes = infos[:, 1].tolist()
    return b"".join(scans[r][: sizes[r]].cpu().numpy().tobytes() for r in range(world))


def test_stripes_device_path(rank_encoders, enc, oracle, synth_hashes):
    import torch
    rng = np.random.default_rng(9)
    for (w, h, world, kind) in [(64, 64, 2, "synth"), (70, 45, 2, "noise"), (200, 37, 4, "synth"), (16, 9, 4, "noise"),
                                (40, 24, 3, "flat"), (1283, 725, 8, "synth")]:
        if kind == "synth":
            rgb = oracle.synth_rgb(w, h, 3, 25)
        elif kind == "noise":
            rgb = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        else:
            rgb = np.full((h, w, 3), 77, np.uint8)
        got = _striped_on_one_gpu_device_path(rank_encoders[:world], torch.from_numpy(rgb).cuda(), w, h)
        assert got == oracle.encode_scan(rgb), (w, h, world, kind)
    e = synth_hashes["7680x4320_seed1_amp20"]
    d = enc.synth(e["w"], e["h"], 1, e["seed"], e["amp"])[0]
    got = _striped_on_one_gpu_device_path(rank_encoders[:8], d, e["w"], e["h"])
    assert len(got) == e["scan_bytes"] and hashlib.sha256(got).hexdigest() == e["scan_sha256"]


def test_dense_tiles_and_ff_bytes(enc, oracle):
    """Adversarial densities: saturated random blocks (hundreds of bits per block, many 0xFF bytes)."""
    rng = np.random.default_rng(77)
    rgb = (rng.integers(0, 2, (264, 520, 1), dtype=np.uint8) * 255).repeat(3, 2)
    e = jb.DeviceEncoder(0, bytes_per_block=184)
    assert e.encode(rgb) == oracle.encode_scan(rgb)
    e.close()


def test_misaligned_base_and_strided_batch(enc, oracle):
    """Input pointers at any byte alignment and batches with padding between images."""
    import torch
    rng = np.random.default_rng(31)
    w, h, n = 101, 37, 3
    imgs = rng.integers(0, 256, (n, h, w, 3), dtype=np.uint8)
    nbytes = w * h * 3
    for shift, pad in [(1, 0), (3, 5), (7, 13), (0, 64)]:
        stride = nbytes + pad
        buf = torch.zeros(shift + n * stride + 64, dtype=torch.uint8, device="cuda")
        for i in range(n):
            buf[shift + i * stride: shift + i * stride + nbytes] = torch.from_numpy(imgs[i].reshape(-1)).cuda()
        view = buf[shift:]
        scan, offs = enc.encode_device(view, w, h, n, image_stride=stride)
        enc.status()
        offs = offs[: n + 1].cpu().numpy()
        for i in range(n):
            got = scan[int(offs[i]): int(offs[i + 1])].cpu().numpy().tobytes()
            assert got == oracle.encode_scan(imgs[i]), (shift, pad, i)


def test_extreme_aspect_ratios(enc, oracle):
    rng = np.random.default_rng(32)
    for (w, h) in [(1, 700), (700, 1), (8, 4099), (4099, 8), (2049, 9), (15, 15)]:
        rgb = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        assert enc.encode(rgb) == oracle.encode_scan(rgb), (w, h)
        smooth = oracle.synth_rgb(w, h, 5, 3)
        assert enc.encode(smooth) == oracle.encode_scan(smooth), (w, h)


def test_many_tiles_cross_group_checkpoint(enc, oracle):
    """More than 1024 K2 tiles in one image exercises the look-back group checkpoints."""
    w, h = 256, 8 * 16 * 1030 + 8                     # 1 strip per block row, 16 strips per tile (K2_TILE_STRIPS) -> 1031 tiles
    rgb = oracle.synth_rgb(w, h, 77, 30)
    assert enc.encode(rgb) == oracle.encode_scan(rgb)


def test_bmp_in_place_ingest_equals_reference_path(enc, oracle, golden, tmp_path):
    """Device-side BMP ingest (bottom-up / top-down rows, padded pitch, BGR, V4/V5 headers) gives the
    same JPEG file as loadBMPImage + saveJPEGGrayscale of the reference build."""
    from oracle.oracle import write_bmp
    rng = np.random.default_rng(12)
    cases = [(golden["lena_crop256/rgb"], False, 40), (golden["greenland_corner250x205/rgb"], False, 40),
             (golden["offset_crop320x213/rgb"], False, 124), (golden["one_pixel/rgb"], True, 40),
             (rng.integers(0, 256, (37, 101, 3), dtype=np.uint8), True, 108), (oracle.synth_rgb(1283, 725, 9, 20), False, 40),
             (rng.integers(0, 256, (9, 7, 3), dtype=np.uint8), False, 40), (oracle.synth_rgb(258, 64, 2, 30), True, 40)]
    for i, (rgb, top_down, hsz) in enumerate(cases):
        p = str(tmp_path / f"c{i}.bmp")
        write_bmp(p, rgb, top_down=top_down, header_size=hsz)
        got = enc.encode_bmp_to_jpeg(open(p, "rb").read())
        assert got == oracle.encode_file_bytes(rgb), (i, rgb.shape, top_down, hsz)
    for name in golden["bmp_names"]:
        data = open(os.path.join(ROOT, "tests", "golden", "bmp", f"{name}.bmp"), "rb").read()
        assert enc.encode_bmp_to_jpeg(data) == oracle.encode_file_bytes(golden[f"bmp/{name}"]), name
    # rejected like the reference loader: wrong magic, 8-bit, compressed, truncated
    bad = bytearray(open(str(tmp_path / "c0.bmp"), "rb").read())
    for patch in ((0, b"P"), (28, b"\x08"), (30, b"\x01")):
        b2 = bytearray(bad)
        b2[patch[0]:patch[0] + 1] = patch[1]
        with pytest.raises(jb.JpegB200Error):
            enc.encode_bmp_to_jpeg(bytes(b2))
    with pytest.raises(jb.JpegB200Error):
        enc.encode_bmp_to_jpeg(bytes(bad[:-5]))


def test_batch_cli(golden, oracle, tmp_path):
    from oracle.oracle import write_bmp
    app = os.path.join(ROOT, "jpeg_image_compression_b200", "jpeg_compression_batch")
    names = ["lena_crop256", "greenland_corner250x205", "noise64", "one_pixel", "synth_200x120_amp20"]
    ins = []
    for n in names:
        p = str(tmp_path / f"{n}.bmp")
        write_bmp(p, golden[f"{n}/rgb"], top_down=(n == "noise64"))
        ins.append(p)
    out_dir = tmp_path / "out"
    out_dir.mkdir()
    r = subprocess.run([app, str(out_dir)] + ins + [str(tmp_path / "missing.bmp")], capture_output=True, text=True)
    assert r.returncode == 1 and "Encoded 5 of 6 files" in r.stdout and "Unable to open file" in r.stderr
    for n in names:
        assert open(out_dir / f"{n}.jpg", "rb").read() == golden[f"{n}/file"].tobytes(), n


def test_two_streams_many_tiles_no_deadlock(synth_hashes):
    """Two encoder handles on two streams, images with more K2 tiles than resident CTAs: the fused
    entropy kernels of both streams run side by side, each only partly resident.  Tiles are handed out
    by an atomic counter, so neither grid can wait on a tile that has not been taken by a running CTA."""
    import torch
    e = synth_hashes["7680x4320_seed1_amp20"]
    encs = [jb.DeviceEncoder(0), jb.DeviceEncoder(0)]
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    d = encs[0].synth(e["w"], e["h"], 1, e["seed"], e["amp"])
    torch.cuda.synchronize()
    outs = []
    for k in range(6):
        for enc_, st in zip(encs, streams):
            with torch.cuda.stream(st):
                cap = enc_.scan_capacity(e["w"], e["h"], 1)
                s = torch.empty(cap, dtype=torch.uint8, device="cuda")
                o = torch.zeros(2, dtype=torch.int64, device="cuda")
                enc_.encode_device(d, e["w"], e["h"], 1, scan=s, offsets=o)
                outs.append((s, o))
    torch.cuda.synchronize()
    for enc_ in encs:
        enc_.status()
    for s, o in outs:
        n = int(o[1].item())
        assert n == e["scan_bytes"] and hashlib.sha256(s[:n].cpu().numpy().tobytes()).hexdigest() == e["scan_sha256"]
    # a taller image: 2700 merge tiles per encode, more than the ~1000 CTAs that can be resident
    w2, h2 = 7680, 11520
    d2 = encs[0].synth(w2, h2, 1, 3, 20)
    s0, o0 = encs[0].encode_device(d2, w2, h2, 1)
    encs[0].status()
    want = s0[: int(o0[1].item())].clone()
    outs = []
    for k in range(3):
        for enc_, st in zip(encs, streams):
            with torch.cuda.stream(st):
                s = torch.empty(enc_.scan_capacity(w2, h2, 1), dtype=torch.uint8, device="cuda")
                o = torch.zeros(2, dtype=torch.int64, device="cuda")
                enc_.encode_device(d2, w2, h2, 1, scan=s, offsets=o)
                outs.append((s, o))
    torch.cuda.synchronize()
    for enc_ in encs:
        enc_.status()
    for s, o in outs:
        n = int(o[1].item())
        assert n == want.numel() and torch.equal(s[:n], want)
    for enc_ in encs:
        enc_.close()


def test_files_framed_on_device(enc, oracle, golden, golden_names):
    """jpegb200_encode_batch_files_device: header + scan + EOI assembled on the device equal the files
    the reference writes (golden /file fixtures), for single images and inside a batch."""
    for name in golden_names:
        rgb = golden[f"{name}/rgb"]
        files = enc.encode_batch_files(rgb[None])
        assert files[0] == golden[f"{name}/file"].tobytes(), name
    imgs = np.stack([oracle.synth_rgb(203, 77, s, 30) for s in range(9)])          # ragged size, odd offsets
    files = enc.encode_batch_files(imgs)
    header = oracle.jfif_header(203, 77)
    for i in range(9):
        assert files[i] == header + oracle.encode_scan(imgs[i]) + b"\xff\xd9", i
    # too small an output buffer is reported, not overrun
    import torch
    d = torch.from_numpy(imgs).cuda()
    small = torch.zeros(1000, dtype=torch.uint8, device="cuda")
    offs = torch.zeros(10, dtype=torch.int64, device="cuda")
    enc.encode_files_device(d, 203, 77, 9, files=small, offsets=offs)
    with pytest.raises(jb.JpegB200Error):
        enc.status()


def test_gigapixel_equals_streaming_oracle(enc, oracle):
    """BASELINE config 5 (32768x32768, beyond the reference's int-indexed buffers): the single-GPU encode of
    the device-generated image against the oracle's streaming (banded) encode of the same synthetic image --
    87 MB of scan, byte for byte (SHA-256)."""
    import torch
    w = h = 32768
    d = enc.synth(w, h, 1, 1, 20)
    scan, offs = enc.encode_device(d, w, h, 1)
    enc.status()
    n = int(offs[1].item())
    got = scan[:n].cpu().numpy().tobytes()
    del d
    torch.cuda.empty_cache()
    want = oracle.encode_scan_synth_banded(w, h, 1, 20, 8, 0)
    assert n == len(want)
    assert hashlib.sha256(got).hexdigest() == hashlib.sha256(want).hexdigest()


def test_tensor_map_path_edges(enc, oracle):
    """Inputs with a 16-byte aligned base and pitch take K1's tensor-map TMA path (one copy per 8 x 768-byte
    tile): ragged heights (zero-filled rows below the image must not leak: the luma pass clamps the row),
    a partial last strip in each block row, single-strip and many-strip widths, and a batch of them."""
    for (w, h) in ((320, 13), (320, 77), (256, 7), (1024, 61), (16, 9), (2048, 23), (3072, 8)):
        assert (3 * w) % 16 == 0
        rgb = oracle.synth_rgb(w, h, 5, 40)
        assert enc.encode(rgb) == oracle.encode_scan(rgb), (w, h)
    rng = np.random.default_rng(11)
    imgs = rng.integers(0, 256, (6, 21, 336, 3), dtype=np.uint8)          # stride 21168 = 16 * 1323
    scans = enc.encode_batch(imgs)
    for i in range(6):
        assert scans[i] == oracle.encode_scan(imgs[i]), i


@pytest.mark.gpu
@pytest.mark.parametrize("handles", [2, 8, 16, 64])
def test_concurrency_hint_does_not_change_bytes(handles):
    """jpegb200_encoder_set_concurrency only reshapes the launches (fewer CTAs, more work per CTA): same scan bytes
    for a single image, a ragged one and a batch."""
    import torch
    base = jb.DeviceEncoder(0)
    hinted = jb.DeviceEncoder(0)
    hinted.set_concurrency(handles)
    for (w, h, n) in [(3840, 2160, 1), (1001, 777, 1), (640, 480, 5)]:
        rgb = base.synth(w, h, n, 7, 20)
        a_scan, a_off = base.encode_device(rgb, w, h, n)
        torch.cuda.synchronize()
        a_off = a_off.cpu().numpy().copy()
        a = a_scan[: int(a_off[n])].cpu().numpy().tobytes()
        b_scan, b_off = hinted.encode_device(rgb, w, h, n)
        torch.cuda.synchronize()
        b_off = b_off.cpu().numpy()
        assert (a_off[: n + 1] == b_off[: n + 1]).all()
        assert a == b_scan[: int(b_off[n])].cpu().numpy().tobytes()


def test_luma_pass_row_alignments(enc, oracle):
    """K1's lane-per-block luma pass reads 8-byte units: rows whose 16-byte phase is 0 or 8 take it (pitch = 8 mod 16,
    or a base shifted by 8), every other phase takes the funnel-shift path -- both against the oracle, full strips and
    a ragged last strip in each row."""
    import torch
    rng = np.random.default_rng(77)
    for (w, h) in [(264, 17), (520, 9), (776, 24), (260, 11), (268, 16), (1032, 8)]:
        rgb = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        assert enc.encode(rgb) == oracle.encode_scan(rgb), (w, h)
    w, h = 512, 19
    rgb = oracle.synth_rgb(w, h, 9, 30)
    want = oracle.encode_scan(rgb)
    for shift in (4, 8, 12, 16, 2):
        buf = torch.zeros(shift + w * h * 3 + 64, dtype=torch.uint8, device="cuda")
        buf[shift: shift + w * h * 3] = torch.from_numpy(rgb.reshape(-1)).cuda()
        scan, offs = enc.encode_device(buf[shift:], w, h, 1)
        enc.status()
        n = int(offs[1].item())
        assert scan[:n].cpu().numpy().tobytes() == want, shift


def test_butterfly_transform_mode_equals_default(enc, oracle):
    """dct_mode 2 (register butterfly transform, kept for comparison) produces the same coefficients and bytes as the
    tensor-core transform: full strips, ragged strips, a batch, a dense image."""
    import torch
    bf = jb.DeviceEncoder(0, dct_mode=2)
    rng = np.random.default_rng(5)
    for (w, h) in [(1024, 64), (264, 17), (1001, 77), (16, 9)]:
        rgb = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        assert bf.encode(rgb) == oracle.encode_scan(rgb), (w, h)
        assert int((bf.coefficients(_nblocks(rgb)) != oracle.coefficients(rgb)).sum()) == 0, (w, h)
    d = enc.synth(1920, 1080, 3, 21, 20)
    a_scan, a_off = enc.encode_device(d, 1920, 1080, 3)
    b_scan, b_off = bf.encode_device(d, 1920, 1080, 3)
    torch.cuda.synchronize()
    n = int(a_off[3].item())
    assert n == int(b_off[3].item())
    assert torch.equal(a_scan[:n], b_scan[:n])
    bf.close()


def test_workspace_guard_bands_intact(oracle, monkeypatch):
    """compute-sanitizer is not available on the GPU pool, so out-of-bounds WRITES are hunted with guard bands:
    JPEGB200_GUARD=1 allocates every workspace buffer at exactly the size the launch asks for, between two 4 KB bands
    of 0xA5 (jpegb200_encoder_check_guards), and the caller-owned output buffers get the same treatment here.  Single
    images (ragged, tiny, tall, dense with the 184 B/block retry), batch, files mode, stripes, hinted launch shapes."""
    import torch
    monkeypatch.setenv("JPEGB200_GUARD", "1")
    encs = [jb.DeviceEncoder(0) for _ in range(3)]
    g = encs[0]
    rng = np.random.default_rng(123)
    singles = [oracle.synth_rgb(203, 77, 3, 30), oracle.synth_rgb(1, 1, 3, 0), oracle.synth_rgb(2048, 24, 3, 20),
               oracle.synth_rgb(256, 8 * 8 * 40, 5, 20), rng.integers(0, 256, (48, 264, 3), dtype=np.uint8),
               oracle.synth_rgb(1001, 333, 4, 64)]
    for rgb in singles:
        assert g.encode(rgb) == oracle.encode_scan(rgb), rgb.shape
    imgs = np.stack([oracle.synth_rgb(120, 50, s, 25) for s in range(5)])
    for i, s in enumerate(g.encode_batch(imgs)):
        assert s == oracle.encode_scan(imgs[i]), i
    hdr = oracle.jfif_header(120, 50)
    for i, f in enumerate(g.encode_batch_files(imgs)):
        assert f == hdr + oracle.encode_scan(imgs[i]) + b"\xff\xd9", i
    # caller-owned device buffers between canaries
    w, h, n = 640, 101, 3
    d = g.synth(w, h, n, 5, 20)
    cap = g.scan_capacity(w, h, n)
    raw = torch.full((cap + 8192,), 0x5A, dtype=torch.uint8, device="cuda")
    raw_off = torch.full((n + 1 + 128,), 0x5A5A5A5A, dtype=torch.int64, device="cuda")
    g.set_concurrency(16)
    g.encode_device(d, w, h, n, scan=raw[4096: 4096 + cap], offsets=raw_off[64: 64 + n + 1])
    g.status()
    g.set_concurrency(1)
    assert bool((raw[:4096] == 0x5A).all()) and bool((raw[4096 + cap:] == 0x5A).all())
    assert bool((raw_off[:64] == 0x5A5A5A5A).all()) and bool((raw_off[64 + n + 1:] == 0x5A5A5A5A).all())
    offs = raw_off[64: 64 + n + 1].cpu().numpy()
    host = d.cpu().numpy().reshape(n, h, w, 3)
    for i in range(n):
        assert raw[4096 + int(offs[i]): 4096 + int(offs[i + 1])].cpu().numpy().tobytes() == oracle.encode_scan(host[i]), i
    # stripes: three ranks emulated on one GPU
    rgb = oracle.synth_rgb(300, 203, 3, 25)
    assert _striped_on_one_gpu(encs, torch.from_numpy(rgb).cuda(), 300, 203) == oracle.encode_scan(rgb)
    for e in encs:
        bad, guarded = e.check_guards()
        assert guarded >= 8, guarded
        assert bad == 0, bad
        e.close()
