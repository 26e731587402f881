#!/usr/bin/env python
"""bench.py -- throughput of the natural_c JPEG encode hot path on B200 (BASELINE.json metric:
Mpixel/s BMP->JPEG encode; achieved HBM GB/s vs peak).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

A "step" is one pass of the hot path (RGB pixels -> stuffed JPEG scan bytes) over one synthetic
input.  Headline workload (SURVEY.md section 8d generator, seed/amp as named there):
    uhd4k      (default, BASELINE configs[1]) one 3840x2160 image per step per GPU; N>1: images sharded by
               rank, no data-path collective (weak scaling)
    batch1080p a batch of 1920x1080 images per step, sharded by image across ranks
The timed region is K steps, every one of them replayed from a CUDA graph (the graph holds exactly K steps, or a
divisor of K), repeated R times back to back until the region is >= 50 ms (`repeats` in the line; ms_per_step is
the mean over R*K steps), bracketed by barrier + synchronize, timed with CUDA events on the launching stream,
max over ranks.  Inputs rotate through a ring of distinct images larger than L2 so every step reads its pixels
from HBM.

`config.extra` carries the sharded configurations of BASELINE.json, measured in the same run at the same N:
    batch1080p_4096   configs[3]: 4096 images in total, split by image across the N ranks (strong scaling)
    stripes_8k        configs[2]: one 7680x4320 image in MCU-row stripes over the N ranks, incl. exchange and gather
    stripes_giga      configs[4]: one 32768x32768 image in MCU-row stripes over the N ranks
(--no-extras skips them.)

--impl reference times the reference's own CPU implementation (oracle/_ref, the unmodified natural_c objects) on
the host cores: per step every host thread encodes one whole 3840x2160 image through the stage chain of
saveJPEGGrayscale; see cpu_baseline in the printed line.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
# stdout must carry exactly ONE JSON line, but native libraries (e.g. NCCL's version banner) write to
# file descriptor 1 too: park the real stdout, point fd 1 at stderr for the whole run, and emit the
# result line on the parked descriptor at the end.
_REAL_STDOUT = None


def park_stdout() -> None:
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict) -> None:
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, (json.dumps(line) + "\n").encode())


METRIC = "Mpixel/s BMP->JPEG encode (natural_c hot path)"
UNIT = "Mpixel/s"
W4K, H4K = 3840, 2160
BAND_ROWS = 128            # reference arm: one 3840x128 band per thread per step


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


# --------------------------------------------------------------------------------------------
# clocks during the timed region (NVML polled from a thread; the region can be short)
class ClockSampler:
    def __init__(self, index: int):
        self.index, self.samples, self.reasons, self._stop = index, [], set(), threading.Event()
        self.max_mhz, self.thread, self.ok = None, None, False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def _poll(self):
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while True:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            if self._stop.wait(0.005):
                break

    def start(self):
        if self.ok:
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()

    def stop(self) -> dict:
        if not self.ok:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable"]}
        self._stop.set()
        self.thread.join()
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# --------------------------------------------------------------------------------------------
# reference CPU implementation (oracle/_ref) -- the one place bench.py may execute oracle/
def _ref_encode_band_fn():
    import ctypes as C
    import numpy as np
    from oracle.oracle import Oracle, Ref, _Img, have_ref
    orc = Oracle()
    if have_ref():
        ref = Ref()
        L = ref.lib

        def encode(rgb: np.ndarray) -> int:
            """the reference's own stage chain, exactly as saveJPEGGrayscale runs it
            (natural_c/src/io/jpeg_handler.c:133-201), host RGB in, scan bytes out"""
            h, w, _ = rgb.shape
            bmp = _Img(w, h, rgb.ctypes.data)
            y = L.convertBMPToJPEGGrayscale(C.byref(bmp))
            c = L.centerYImage(y)
            d = L.performDCT(c)
            q = L.quantizeImage(d)
            z = L.performZigZag(q)
            r = L.performRLE(z)
            b = L.encodeHuffman(r, z.contents.totalBlocks)
            n = b.contents.size
            L.freeJpegEncoderBuffer(b); L.freeRLEData(r); L.freeZigZagData(z); L.freeQuantizedImage(q)
            L.freeDCTImage(d); L.freeCenteredYImage(c); L.freeYImage(y)
            return n
        # warm the reference's lazily initialised (non thread-safe) Huffman tables once
        encode(np.zeros((8, 8, 3), np.uint8))
        return encode, "reference", orc

    def encode_port(rgb):
        return len(orc.encode_scan(rgb))
    return encode_port, "port", orc


def _ref_encode_fn_from(path):
    """encode(rgb) -> scan size through the stage chain of the reference objects in the shared library at `path`."""
    import ctypes as C
    from oracle.oracle import Ref, _Img
    if not os.path.exists(path):
        return None
    L = Ref(path).lib

    def encode(rgb):
        h, w, _ = rgb.shape
        bmp = _Img(w, h, rgb.ctypes.data)
        y = L.convertBMPToJPEGGrayscale(C.byref(bmp))
        c = L.centerYImage(y)
        d = L.performDCT(c)
        q = L.quantizeImage(d)
        z = L.performZigZag(q)
        r = L.performRLE(z)
        b = L.encodeHuffman(r, z.contents.totalBlocks)
        n = b.contents.size
        L.freeJpegEncoderBuffer(b); L.freeRLEData(r); L.freeZigZagData(z); L.freeQuantizedImage(q)
        L.freeDCTImage(d); L.freeCenteredYImage(c); L.freeYImage(y)
        return n
    return encode


WORKLOAD_4K = "uhd4k: synthetic 3840x2160 24-bit RGB, single image encode (BASELINE configs[1]); N>1: images sharded by rank"


def time_reference(steps: int, warmup: int, threads: int):
    """Each step: `threads` host threads each encode one WHOLE 3840x2160 image of the 4K workload (distinct seeds)
    with the reference's CPU code, stage by stage as saveJPEGGrayscale does (ctypes releases the GIL).
    Returns Mpixel/s, ms per step, the kind and the pixels per step."""
    from concurrent.futures import ThreadPoolExecutor
    encode, kind, orc = _ref_encode_band_fn()
    nimg = min(threads, 4)                                   # distinct images; more would only cost host memory
    images = [orc.synth_rgb(W4K, H4K, 1 + i, 20) for i in range(nimg)]
    px_per_step = threads * W4K * H4K
    with ThreadPoolExecutor(threads) as pool:
        def one_step(k):
            list(pool.map(encode, [images[(k + t) % nimg] for t in range(threads)]))
        for k in range(warmup):
            one_step(k)
        t0 = time.perf_counter()
        for k in range(steps):
            one_step(warmup + k)
        dt = time.perf_counter() - t0
    # the stock behaviour beside it: one thread, one whole image at a time
    t1 = time.perf_counter()
    n1 = 0
    while n1 < 3:
        encode(images[n1 % nimg])
        n1 += 1
    single = W4K * H4K * n1 / (time.perf_counter() - t1) / 1e6
    return px_per_step * steps / dt / 1e6, dt / steps * 1e3, kind, px_per_step, single


def run_reference(args, rank: int, world: int):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    val, ms, kind, px, single = time_reference(args.steps, args.warmup, threads)
    sample = (f"per step each of {threads} host threads encodes one whole synthetic 3840x2160 image (amp=20) with the "
              f"reference natural_c core stages in the order of saveJPEGGrayscale (-O2 build of the unmodified sources); "
              f"the stock single-threaded program does {single:.1f} Mpixel/s on this host")
    line = {
        "impl": "reference", "metric": METRIC, "value": round(val, 3), "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms, 4), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD_4K, "seed": 1, "amp": 20, "images_per_step": threads, "host_threads": threads},
        "cpu_baseline": {"value": round(val, 3), "unit": UNIT, "cores": threads, "kind": kind, "sample": sample,
                         "stock_single_thread": {"value": round(single, 3), "unit": UNIT, "cores": 1}},
        "e2e": {"value": round(val, 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# --------------------------------------------------------------------------------------------
def _graph_len(steps: int) -> int:
    """Length of the CUDA graph: exactly `steps` if that is small, else the largest divisor of `steps` in [8, 64]
    (so that the timed region is a whole number of replays and no step runs eagerly), else 32 + a remainder graph."""
    if steps <= 64:
        return steps
    for L in range(64, 7, -1):
        if steps % L == 0:
            return L
    return 32


def _static_traffic(workload: str):
    """DRAM bytes of the block kernel per launch from the committed ncu capture of this workload (profiles/)."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "k1_dram_traffic.json")))
        e = d.get(workload)
        if e:
            return e.get("dram_bytes_per_launch"), f"static_ncu ({e.get('capture', '?')}, commit {e.get('commit', '?')})"
    except Exception:
        pass
    return None, None


def run_ours(args, rank: int, local_rank: int, world: int):
    import hashlib
    import numpy as np
    import torch
    import torch.distributed as dist
    import jpeg_image_compression_b200 as jb
    from jpeg_image_compression_b200.stripes import StripedEncoder, shard_range, stripe_rows

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (libjpegb200 has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # one encoder handle (workspace) per in-flight image: consecutive encodes are independent, so with several
    # handles on several streams the three kernels of different images overlap
    nstreams = max(1, args.streams)
    encs = [jb.DeviceEncoder(local_rank) for _ in range(nstreams)]
    enc = encs[0]

    if args.workload == "uhd4k":
        w, h, per_step = W4K, H4K, 1
        ring = nstreams * ((8 + nstreams - 1) // nstreams)
        wl_name = WORKLOAD_4K
    elif args.workload == "batch1080p":
        w, h, per_step, ring = 1920, 1080, args.batch, 2
        wl_name = f"batch1080p: {args.batch} synthetic 1920x1080 images per step per GPU, sharded by image"
    else:
        raise SystemExit(f"unknown workload {args.workload}")
    px_step = w * h * per_step
    # ring of distinct inputs, larger than the 126 MB L2, generated on the device
    seed0 = 1 + rank * 100000
    inputs = [enc.synth(w, h, per_step, seed0 + i * per_step, 20) for i in range(ring)]
    ring_bytes = sum(t.numel() for t in inputs)
    cap = enc.scan_capacity(w, h, per_step)
    outs = [(torch.empty(cap, dtype=torch.uint8, device=dev), torch.zeros(per_step + 1, dtype=torch.int64, device=dev))
            for _ in range(ring)]

    def step(i, e=None):
        k = i % ring
        (e or enc).encode_device(inputs[k], w, h, per_step, scan=outs[k][0], offsets=outs[k][1])

    # ---- warm-up (also sizes the workspaces), launch count ------------------------------------------
    for e in encs:
        for i in range(max(args.warmup, ring)):
            step(i, e)
    torch.cuda.synchronize()
    for e in encs:
        e.status()
    launches_per_step = enc.stats()["kernel_launches"]
    scan_bytes = [int(o[per_step].item()) for _, o in outs]

    # ---- CUDA graphs that together hold exactly K steps ---------------------------------------------
    def capture(n_streams, nsteps, handles=None):
        handles = handles or encs
        side = [torch.cuda.Stream() for _ in range(n_streams)]
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            main = torch.cuda.current_stream()
            if n_streams == 1:
                for i in range(nsteps):
                    step(i)
            else:
                for s_ in side:
                    s_.wait_stream(main)
                for j, s_ in enumerate(side):
                    with torch.cuda.stream(s_):
                        for i in range(j, nsteps, n_streams):
                            step(i, handles[j])
                for s_ in side:
                    main.wait_stream(s_)
        return g

    use_graph = not args.no_graph
    glen = _graph_len(args.steps)
    full, rest = divmod(args.steps, glen)

    def make_runner(n_streams, handles=None):
        if not use_graph:
            def eager(n=args.steps):
                for i in range(n):
                    step(i, (handles or encs)[i % n_streams] if n_streams > 1 else None)
            return eager
        g_main = capture(n_streams, glen, handles)
        g_rest = capture(n_streams, rest, handles) if rest else None

        def run_k():
            for _ in range(full):
                g_main.replay()
            if g_rest is not None:
                g_rest.replay()
        return run_k

    run_k = make_runner(nstreams)
    run_k()
    torch.cuda.synchronize()
    for e in encs:
        e.status()

    def timed(fn, repeats):
        """R x K steps between two events; returns total ms."""
        for _ in range(max(1, (args.warmup + args.steps - 1) // args.steps)):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sync_all()
        e0.record()
        for _ in range(repeats):
            fn()
        e1.record()
        sync_all()
        return e0.elapsed_time(e1)

    # ---- device-resident timing: R x K steps, >= 50 ms ---------------------------------------------
    pilot = max_over_ranks(timed(run_k, 1))
    repeats = max(1, min(100000, int(np.ceil(50.0 / max(pilot, 1e-3)))))
    sampler = ClockSampler(local_rank)
    sampler.start()
    ms_total = max_over_ranks(timed(run_k, repeats))
    clocks = sampler.stop()
    for e in encs:
        e.status()
    timed_steps = repeats * args.steps
    value = px_step * timed_steps * world / (ms_total * 1e-3) / 1e6
    single_stream_ms = None
    if use_graph and nstreams > 1:                          # for the record: strictly serial encodes
        run_1 = make_runner(1)
        single_stream_ms = timed(run_1, repeats) / timed_steps

    # ---- the same workload with the library's concurrency hint: 16 handles, each sized for its share of the SMs ----
    shared_device = None
    if args.workload == "uhd4k" and use_graph and not args.no_shared_device:
        n_sh = 16
        sh_encs = [jb.DeviceEncoder(local_rank) for _ in range(n_sh)]
        for e in sh_encs:
            e.set_concurrency(n_sh)
            for i in range(ring):
                step(i, e)
        torch.cuda.synchronize()
        for e in sh_encs:
            e.status()
        sh_ring = [bytes(outs[k][0][: int(outs[k][1][per_step].item())].cpu().numpy().tobytes()) for k in range(min(ring, 2))]
        for k in range(min(ring, 2)):                       # the hint must not change a byte
            step(k, enc)
        torch.cuda.synchronize()
        same = all(sh_ring[k] == outs[k][0][: int(outs[k][1][per_step].item())].cpu().numpy().tobytes() for k in range(min(ring, 2)))
        sh_steps = 15 * n_sh                                # one graph of 15 images per handle: the pipeline drains at graph ends
        g_sh = capture(n_sh, sh_steps, sh_encs)
        run_sh = g_sh.replay
        r_sh = max(1, int(np.ceil(50.0 / max(max_over_ranks(timed(run_sh, 1)), 1e-3))))
        ms_sh = max_over_ranks(timed(run_sh, r_sh))
        for e in sh_encs:
            e.status()
        shared_device = {"value": round(px_step * sh_steps * r_sh * world / (ms_sh * 1e-3) / 1e6, 1), "unit": UNIT,
                         "encoder_streams": n_sh, "concurrency_hint": n_sh, "steps": sh_steps * r_sh, "graph_steps": sh_steps,
                         "bytes_equal_to_default_launch_shape": bool(same),
                         "note": "jpegb200_encoder_set_concurrency(16): every launch sized for ~0.3 of the SMs; "
                                 "the headline value keeps the default (whole-device) launch shape"}
        del run_sh, g_sh, sh_encs

    # ---- sensitivity rows of SURVEY.md 8(d): the same workload at amp=0 (smooth) and amp=64 (entropy-heavy) ----
    sensitivity = None
    if args.workload == "uhd4k" and world == 1 and use_graph and not args.no_sensitivity:
        sensitivity = {}
        main_inputs = inputs
        for amp in (0, 64):
            inputs = [enc.synth(w, h, per_step, seed0 + i * per_step, amp) for i in range(ring)]
            for e in encs:
                for i in range(ring):
                    step(i, e)
            torch.cuda.synchronize()
            for e in encs:
                e.status()
            sb = sum(int(o[per_step].item()) for _, o in outs) / ring
            run_s = make_runner(nstreams)
            r_s = max(1, repeats // 4)
            ms = timed(run_s, r_s)
            for e in encs:
                e.status()
            sensitivity[f"amp{amp}"] = {"value": round(px_step * r_s * args.steps / (ms * 1e-3) / 1e6, 1), "unit": UNIT,
                                        "scan_bytes_per_step": int(sb), "steps": r_s * args.steps}
        inputs = main_inputs
        for i in range(ring):
            step(i)
        torch.cuda.synchronize()

    # ---- per-kernel time of the dominant kernel (cudaEvents on the launching stream) -----------
    enc.set_profiling(True)
    enc.kernel_times(reset=True)
    prof_steps = min(max(args.steps, 8), 64)
    for i in range(prof_steps):
        step(i)
    kt = enc.kernel_times(reset=True)
    enc.set_profiling(False)
    k1_ms = kt["ms"][0] / max(kt["calls"][0], 1)
    step_ms_sum = sum(kt["ms"]) / prof_steps
    mean_scan = sum(scan_bytes) / len(scan_bytes)
    algo_bytes = 3.0 * px_step + mean_scan                 # A = 3*W*H + S per launch
    peak, peak_src = measured_peak_gbs()
    achieved = algo_bytes / (k1_ms * 1e-3) / 1e9
    traffic, traffic_src = _static_traffic(args.workload)
    if traffic is not None and args.workload == "batch1080p":
        traffic = int(traffic * per_step)                  # the capture is per image
    roofline = {"bound": "hbm", "kernel": "k_fused_blocks", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4), "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": int(algo_bytes), "kernel_ms": round(k1_ms, 5),
                "kernel_share_of_step": round(k1_ms / step_ms_sum, 3) if step_ms_sum > 0 else None,
                "per_kernel_ms": {n: round(m / max(c, 1), 5) for n, m, c in zip(kt["names"], kt["ms"], kt["calls"]) if c},
                "step_frac": round(algo_bytes / (ms_total / timed_steps * 1e-3) / 1e9 / peak, 4),
                # SURVEY.md 8(d) judges the block kernel on its ACTUAL DRAM bytes (ncu) per second; the capture is static
                "traffic_frac": round(traffic / (k1_ms * 1e-3) / 1e9 / peak, 4) if traffic else None}

    # ---- end to end: pinned host buffers -> C-ABI call -> pinned host buffers --------------------
    e2e = None
    if args.workload == "uhd4k":
        # one host thread per encoder handle, each with its own stream: the H2D copy of one image
        # overlaps the kernels and the D2H copy of the others (the C call releases the GIL)
        nthr = min(len(encs), 4)                           # 4 host threads saturate the PCIe link; more only add contention
        pin_in = [inputs[i].cpu().pin_memory() for i in range(min(ring, 4))]
        pin_out = [torch.empty(cap, dtype=torch.uint8).pin_memory() for _ in range(nthr)]
        streams = [torch.cuda.Stream() for _ in range(nthr)]
        e2e_steps = max(nthr, min(max(args.steps, 16), 64))
        e2e_steps -= e2e_steps % nthr
        d2h_total = [0] * nthr

        def worker(j, n):
            torch.cuda.set_device(local_rank)
            with torch.cuda.stream(streams[j]):
                for i in range(n):
                    d2h_total[j] += encs[j].encode_host(pin_in[(i * nthr + j) % len(pin_in)], w, h, pin_out[j])

        def copy_worker(j, n):                              # the same H2D traffic without any encode: the platform's ceiling
            torch.cuda.set_device(local_rank)
            with torch.cuda.stream(streams[j]):
                for i in range(n):
                    inputs[j % ring].copy_(pin_in[(i * nthr + j) % len(pin_in)].view_as(inputs[j % ring]), non_blocking=True)
                streams[j].synchronize()

        def run_threads(fn, n_per_thread):
            ts = [threading.Thread(target=fn, args=(j, n_per_thread)) for j in range(nthr)]
            for t_ in ts:
                t_.start()
            for t_ in ts:
                t_.join()

        def wall(fn, n_per_thread):
            sync_all()
            t0 = time.perf_counter()
            run_threads(fn, n_per_thread)
            torch.cuda.synchronize()
            return max_over_ranks(time.perf_counter() - t0)

        run_threads(worker, 2)
        torch.cuda.synchronize()
        d2h_total = [0] * nthr
        dt = wall(worker, e2e_steps // nthr)
        d2h = int(sum(d2h_total) / e2e_steps) + 16
        dt_copy = wall(copy_worker, e2e_steps // nthr)
        for i in range(ring):                               # the copy test overwrote device inputs: regenerate
            inputs[i] = enc.synth(w, h, per_step, seed0 + i * per_step, 20)
        e2e = {"value": round(px_step * e2e_steps * world / dt / 1e6, 1), "unit": UNIT,
               "h2d_bytes_per_step": 3 * px_step, "d2h_bytes_per_step": d2h,
               "steps": e2e_steps, "host_threads": nthr,
               "h2d_gbs_per_rank": round(3 * px_step * e2e_steps / dt / 1e9, 2),
               "h2d_copy_only_gbs_per_rank": round(3 * px_step * e2e_steps / dt_copy / 1e9, 2),
               "api": "jpegb200_encode_host (C ABI, pinned host buffers, H2D + 3 kernels + D2H per call)"}
    else:
        # batch: pinned host batch -> device -> encode -> D2H of offsets + scan bytes
        pin = inputs[0].cpu().pin_memory()
        dbuf = torch.empty_like(inputs[0])
        pin_out = torch.empty(cap, dtype=torch.uint8).pin_memory()
        e2e_steps = max(2, min(args.steps, 5))
        t0 = None
        for i in range(e2e_steps + 1):
            if i == 1:
                torch.cuda.synchronize()
                t0 = time.perf_counter()
            dbuf.copy_(pin, non_blocking=True)
            s_, o = enc.encode_device(dbuf, w, h, per_step, scan=outs[0][0], offsets=outs[0][1])
            n = int(o[per_step].item())
            pin_out[:n].copy_(s_[:n], non_blocking=True)
            torch.cuda.synchronize()
        dt = max_over_ranks(time.perf_counter() - t0)
        e2e = {"value": round(px_step * e2e_steps * world / dt / 1e6, 1), "unit": UNIT,
               "h2d_bytes_per_step": 3 * px_step, "d2h_bytes_per_step": n + 8 * (per_step + 1), "steps": e2e_steps,
               "api": "DeviceEncoder.encode_device with pinned H2D/D2H around jpegb200_encode_batch_device"}
        del pin, dbuf, pin_out

    # ---- the sharded configurations of BASELINE.json at this N (config.extra) -------------------------
    extra = None
    if not args.no_extras:
        extra = {}
        hashes = {}
        try:
            hashes = json.load(open(os.path.join(ROOT, "tests", "golden", "synth_hashes.json")))
        except Exception:
            pass
        del inputs, outs
        torch.cuda.empty_cache()

        # configs[3]: 4096 x 1080p in total, split by image (strong scaling); sub-batches of <= 512 images per launch
        total_images = args.extra_images
        lo, hi = shard_range(total_images, world, rank)
        sub = 512
        chunks = [(b, min(b + sub, hi)) for b in range(lo, hi, sub)]
        bw_, bh_ = 1920, 1080
        cins = [enc.synth(bw_, bh_, e_ - b_, 1 + b_, 20) for b_, e_ in chunks]
        ccap = enc.scan_capacity(bw_, bh_, sub)
        cscan = torch.empty(ccap, dtype=torch.uint8, device=dev)
        coffs = torch.zeros(sub + 1, dtype=torch.int64, device=dev)

        def batch_step():
            for t_, (b_, e_) in zip(cins, chunks):
                enc.encode_device(t_, bw_, bh_, e_ - b_, scan=cscan, offsets=coffs)

        batch_step()
        torch.cuda.synchronize()
        enc.status()
        nb_steps = 3
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sync_all()
        e0.record()
        for _ in range(nb_steps):
            batch_step()
        e1.record()
        sync_all()
        enc.status()
        bms = max_over_ranks(e0.elapsed_time(e1)) / nb_steps
        verified = None
        key = "1920x1080_seed64_amp20"                       # image 63 of the batch (seed 1 + 63): rank 0 owns it
        if rank == 0 and key in hashes and chunks and chunks[0][0] == 0 and chunks[0][1] > 63:
            enc.encode_device(cins[0], bw_, bh_, chunks[0][1] - chunks[0][0], scan=cscan, offsets=coffs)
            enc.status()
            o63, o64 = int(coffs[63].item()), int(coffs[64].item())
            ok = (o64 - o63) == hashes[key]["scan_bytes"] and \
                hashlib.sha256(cscan[o63:o64].cpu().numpy().tobytes()).hexdigest() == hashes[key]["scan_sha256"]
            verified = "image 63 byte-identical to the reference build (sha256)" if ok else "MISMATCH"
        extra["batch1080p_4096"] = {"config": "BASELINE configs[3]", "scaling": "strong", "images_total": total_images,
                                    "images_this_rank": hi - lo, "ms_per_batch": round(bms, 4), "steps": nb_steps,
                                    "value": round(total_images * bw_ * bh_ / (bms * 1e-3) / 1e6, 1), "unit": UNIT,
                                    "verified": verified}
        del cins, cscan, coffs
        torch.cuda.empty_cache()

        # configs[2] / configs[4]: MCU-row stripes of ONE image over the N ranks, exchange on the device
        def stripes(name, sw, sh, hash_key, iters):
            y0, owned, halo = stripe_rows(sh, world, rank)
            full_img = enc.synth(sw, sh, 1, 1, 20)[0]
            stripe = full_img[y0:y0 + owned + halo].contiguous() if world > 1 else full_img
            del full_img
            torch.cuda.empty_cache()
            scan = torch.empty(enc.scan_capacity(sw, max(owned, 8), 1), dtype=torch.uint8, device=dev)
            out = {"config": name, "width": sw, "height": sh, "ranks": world}
            if world == 1:
                offs = torch.zeros(2, dtype=torch.int64, device=dev)

                def one():
                    enc.encode_device(stripe, sw, sh, 1, scan=scan, offsets=offs)
                gather = None
            else:
                se = StripedEncoder(enc, device=dev)
                summ = torch.zeros((world, 2), dtype=torch.int64, device=dev)
                info = torch.zeros(2, dtype=torch.int64, device=dev)

                def one():
                    se.encode_on_device(stripe, sw, sh, scan, summaries=summ, info=info)

                def gather():
                    return se.gather_exact(scan, info[1])
            one()
            torch.cuda.synchronize()
            enc.status()
            # `iters` images in 5 chunks: the mean over all of them is the reported figure, the best chunk is kept beside it
            # (the 16-byte all-gather has a latency tail that a 60 us image feels)
            chunk = max(1, iters // 5)
            evs = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
            sync_all()
            evs[0].record()
            for c_ in range(5):
                for _ in range(chunk):
                    one()
                evs[c_ + 1].record()
            sync_all()
            enc.status()
            iters = 5 * chunk
            ms = max_over_ranks(evs[0].elapsed_time(evs[5])) / iters
            best = min(max_over_ranks(evs[c_].elapsed_time(evs[c_ + 1])) / chunk for c_ in range(5))
            out.update({"ms_per_image": round(ms, 4), "ms_per_image_best_chunk": round(best, 4),
                        "value": round(sw * sh / (ms * 1e-3) / 1e6, 1), "unit": UNIT, "steps": iters,
                        "timed": "analyze + all-gather of the boundary summaries + merge (device-resident, no host round trip)"
                        if world > 1 else "single-GPU encode"})
            if world > 1:
                gather()                                      # first use sets up the NCCL point-to-point connections
                sync_all()
                t0 = time.perf_counter()
                stitched, total = gather()
                torch.cuda.synchronize()
                out["gather_to_rank0_ms"] = round(max_over_ranks(time.perf_counter() - t0) * 1e3, 3)
                out["scan_bytes"] = total
                data = stitched[:total] if rank == 0 else None
            else:
                total = int(offs[1].item())
                out["scan_bytes"] = total
                data = scan[:total]
            if rank == 0 and hash_key in hashes:
                ok = total == hashes[hash_key]["scan_bytes"] and \
                    hashlib.sha256(data.cpu().numpy().tobytes()).hexdigest() == hashes[hash_key]["scan_sha256"]
                out["verified"] = "byte-identical to the reference build (sha256)" if ok else "MISMATCH"
            return out

        extra["stripes_8k"] = stripes("BASELINE configs[2]", 7680, 4320, "7680x4320_seed1_amp20", 50)
        if not args.no_giga:
            extra["stripes_giga"] = stripes("BASELINE configs[4]", 32768, 32768, "32768x32768_seed1_amp20", 5)

    # ---- CPU baseline beside it (rank 0, N=1 only; bounded sample) -------------------------------
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        encode, kind, orc = _ref_encode_band_fn()
        full_ = orc.synth_rgb(W4K, H4K, 1, 20)
        t0 = time.perf_counter()
        n_img = 0
        while True:
            encode(full_)
            n_img += 1
            if time.perf_counter() - t0 > 8.0 or n_img >= 30:       # a bounded sample
                break
        dt = time.perf_counter() - t0
        cpu_baseline = {"value": round(W4K * H4K * n_img / dt / 1e6, 3), "unit": UNIT, "cores": 1, "kind": kind,
                        "sample": f"{n_img} full synthetic 3840x2160 images (seed=1, amp=20) through the reference's seven "
                                  f"core stages, single thread (the reference is single-threaded), {dt:.1f} s"}
        # the stock "-g" build of natural_c/Makefile beside the -O2 twin, and all cores over disjoint 1080p images
        try:
            enc_g = _ref_encode_fn_from(os.path.join(ROOT, "oracle", "_ref", "libnaturalc_ref_g.so"))
            if enc_g is not None:
                enc_g(np.zeros((8, 8, 3), np.uint8))          # the reference initialises its Huffman tables lazily
                t0 = time.perf_counter()
                enc_g(full_)
                cpu_baseline["stock_g_build"] = {"value": round(W4K * H4K / (time.perf_counter() - t0) / 1e6, 3), "unit": UNIT, "cores": 1,
                                                 "sample": "one 3840x2160 image, reference sources at natural_c/Makefile's own flags (-g)"}
            from concurrent.futures import ThreadPoolExecutor
            thr = os.cpu_count() or 1
            imgs = [orc.synth_rgb(1920, 1080, 1 + i, 20) for i in range(min(thr, 8))]
            with ThreadPoolExecutor(thr) as pool:
                list(pool.map(encode, [imgs[i % len(imgs)] for i in range(thr)]))
                t0 = time.perf_counter()
                rounds = 2
                for _ in range(rounds):
                    list(pool.map(encode, [imgs[i % len(imgs)] for i in range(thr)]))
                dt = time.perf_counter() - t0
            cpu_baseline["all_cores_1080p"] = {"value": round(1920 * 1080 * thr * rounds / dt / 1e6, 3), "unit": UNIT, "cores": thr,
                                               "sample": f"{rounds} rounds of {thr} threads, one whole 1920x1080 image each (BASELINE configs[3] on the CPU)"}
        except Exception as ex:                               # the extra rows are best effort
            cpu_baseline["note"] = f"extra CPU rows skipped: {ex}"

    if rank == 0:
        line = {
            "metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms_total / timed_steps, 5), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "repeats": repeats, "timed_steps": timed_steps,
            "config": {"workload": wl_name, "width": w, "height": h, "images_per_step_per_gpu": per_step, "seed": seed0,
                       "amp": 20, "l2": f"inputs rotate through a ring of {ring} distinct device buffers "
                                         f"({ring_bytes / 1e6:.0f} MB > 126 MB L2), no flush needed",
                       "cuda_graph": bool(use_graph), "graph_steps": glen if use_graph else None,
                       "scan_bytes_per_step": int(mean_scan), "encoder_streams": nstreams,
                       "single_stream_ms_per_step": round(single_stream_ms, 5) if single_stream_ms else None,
                       "shared_device": shared_device, "sensitivity": sensitivity, "extra": extra},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches_per_step * timed_steps),
            "roofline": roofline, "cpu_baseline": cpu_baseline,
        }
        emit(line)
    for e in encs:
        e.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    park_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None, help="timed steps (default: 20000 for uhd4k, 200 for batch1080p, 100 for --impl reference)")
    ap.add_argument("--warmup", type=int, default=None, help="untimed warm-up steps (default: 200 / 5 / 5)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="uhd4k", choices=["uhd4k", "batch1080p"])
    ap.add_argument("--batch", type=int, default=512, help="images per step per GPU for batch1080p")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--streams", type=int, default=8, help="encoder handles / streams with independent images in flight")
    ap.add_argument("--no-extras", action="store_true", help="skip config.extra (the sharded BASELINE configs)")
    ap.add_argument("--no-giga", action="store_true", help="skip the 32768x32768 stripe row of config.extra")
    ap.add_argument("--extra-images", type=int, default=4096, help="total images of the strong-scaling batch row")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-shared-device", action="store_true", help="skip the 16-handle concurrency-hint row")
    ap.add_argument("--no-sensitivity", action="store_true", help="skip the amp=0 / amp=64 rows")
    args = ap.parse_args()
    if args.impl == "reference":
        dsteps, dwarm = 20, 3                               # the CPU arm: whole 4K images on every core, ~0.5 s per step
    elif args.workload == "batch1080p":
        dsteps, dwarm = 200, 5
    else:
        dsteps, dwarm = 2000, 200                           # repeated until the timed region is >= 50 ms
    args.steps = dsteps if args.steps is None else args.steps
    args.warmup = max(dwarm if args.warmup is None else args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
