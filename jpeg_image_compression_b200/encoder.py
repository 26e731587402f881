"""Device-resident encoder: torch tensors for HBM buffers and streams, the C ABI for the work."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import Batch, Stats, StripeSummary, check, load_library


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise _lib.JpegB200Error("no CUDA device: libjpegb200 has no CPU fallback")
    return torch


class DeviceEncoder:
    """One encoder handle (workspace + tables) on one GPU."""

    def __init__(self, device: int = 0, dct_mode: int = 0, bytes_per_block: int = 24):
        self.torch = _torch()
        self.lib = load_library()
        self.device = device
        self.handle = self.lib.jpegb200_encoder_create(device)
        if not self.handle:
            raise _lib.JpegB200Error("jpegb200_encoder_create: " + _lib.last_error())
        check(self.lib.jpegb200_encoder_set_dct_mode(self.handle, dct_mode), "set_dct_mode")
        check(self.lib.jpegb200_encoder_set_bytes_per_block(self.handle, bytes_per_block), "set_bytes_per_block")
        self.bytes_per_block = bytes_per_block
        self._scan = None
        self._offsets = None

    def set_concurrency(self, handles: int) -> None:
        """Launch-shape hint (include/jpegb200.h): `handles` encoders are kept busy side by side on this GPU."""
        check(self.lib.jpegb200_encoder_set_concurrency(self.handle, int(handles)), "set_concurrency")

    def check_guards(self):
        """(corrupted guard bytes, guarded buffers) -- see JPEGB200_GUARD in include/jpegb200.h."""
        bad, n = C.c_uint64(0), C.c_int(0)
        check(self.lib.jpegb200_encoder_check_guards(self.handle, C.byref(bad), C.byref(n)), "check_guards")
        return int(bad.value), int(n.value)

    def close(self):
        if self.handle:
            self.lib.jpegb200_encoder_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        return C.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)

    # ---- buffers -------------------------------------------------------------------
    def scan_capacity(self, w: int, h: int, count: int) -> int:
        nb = ((w + 7) // 8) * ((h + 7) // 8)
        return count * (nb * self.bytes_per_block + 64) * 2

    def _ensure_out(self, capacity: int, count: int):
        t = self.torch
        if self._scan is None or self._scan.numel() < capacity:
            self._scan = t.empty(capacity, dtype=t.uint8, device=f"cuda:{self.device}")
        if self._offsets is None or self._offsets.numel() < count + 1:
            self._offsets = t.zeros(count + 1, dtype=t.int64, device=f"cuda:{self.device}")
        return self._scan, self._offsets

    # ---- encode --------------------------------------------------------------------
    def encode_device(self, d_rgb, w: int, h: int, count: int = 1, image_stride: int = 0, scan=None, offsets=None):
        """Launch the encode of `count` images resident in HBM (async on the current stream).
        Returns (scan tensor, offsets tensor[count+1]) -- device tensors."""
        if scan is None or offsets is None:
            scan, offsets = self._ensure_out(self.scan_capacity(w, h, count), count)
        b = Batch(d_rgb.data_ptr(), w, h, count, image_stride)
        check(self.lib.jpegb200_encode_batch_device(self.handle, C.byref(b), scan.data_ptr(), scan.numel(),
                                                    offsets.data_ptr(), self._stream()), "encode_batch_device")
        return scan, offsets

    def encode_files_device(self, d_rgb, w: int, h: int, count: int = 1, image_stride: int = 0, files=None, offsets=None):
        """Like encode_device, but the output tensor holds complete JFIF files (header + scan + EOI per
        image, back to back; offsets[count+1]) -- jpegb200_encode_batch_files_device."""
        if files is None or offsets is None:
            files, offsets = self._ensure_out(self.scan_capacity(w, h, count) + 330 * count, count)
        b = Batch(d_rgb.data_ptr(), w, h, count, image_stride)
        check(self.lib.jpegb200_encode_batch_files_device(self.handle, C.byref(b), files.data_ptr(), files.numel(),
                                                          offsets.data_ptr(), self._stream()), "encode_batch_files_device")
        return files, offsets

    def encode_batch_files(self, rgbs: np.ndarray) -> list:
        """Host RGB (n,h,w,3) -> list of complete .jpg file images, framed on the device."""
        t = self.torch
        rgbs = np.ascontiguousarray(rgbs, np.uint8)
        n, h, w, _ = rgbs.shape
        d = t.from_numpy(rgbs).to(f"cuda:{self.device}")
        files, offsets = self.encode_files_device(d, w, h, n)
        self.status()
        offs = offsets[: n + 1].cpu().numpy()
        data = files[: int(offs[n])].cpu().numpy().tobytes()
        return [data[int(offs[i]): int(offs[i + 1])] for i in range(n)]

    def status(self):
        check(self.lib.jpegb200_encoder_status(self.handle, self._stream()), "encoder_status")

    def stats(self) -> dict:
        s = Stats()
        check(self.lib.jpegb200_encoder_stats(self.handle, C.byref(s)), "encoder_stats")
        return {"blocks": s.blocks, "flagged_coefficients": s.flagged_coefficients,
                "kernel_launches": s.kernel_launches, "packed_bytes": s.packed_bytes}

    def encode(self, rgb: np.ndarray) -> bytes:
        """Host RGB (h,w,3) -> scan bytes through the device API (H2D, kernels, D2H)."""
        scans = self.encode_batch(rgb[None])
        return scans[0]

    def _grow_workspace(self):
        """A dense image overflowed the packed-bits workspace or the scan buffer sized from bytes_per_block: switch to
        the worst case (184 bytes per block covers every input), like the C entries jpegb200_encode_scan / the batch CLI."""
        self.bytes_per_block = 184
        check(self.lib.jpegb200_encoder_set_bytes_per_block(self.handle, 184), "set_bytes_per_block")
        self._scan = None

    def encode_batch(self, rgbs: np.ndarray) -> list:
        t = self.torch
        rgbs = np.ascontiguousarray(rgbs, np.uint8)
        n, h, w, _ = rgbs.shape
        d = t.from_numpy(rgbs).to(f"cuda:{self.device}")
        scan, offsets = self.encode_device(d, w, h, n)
        try:
            self.status()
        except _lib.JpegB200Error:
            if self.bytes_per_block >= 184:
                raise
            self._grow_workspace()
            scan, offsets = self.encode_device(d, w, h, n)
            self.status()
        offs = offsets[: n + 1].cpu().numpy()
        data = scan[: int(offs[n])].cpu().numpy().tobytes()
        return [data[int(offs[i]): int(offs[i + 1])] for i in range(n)]

    def encode_host(self, host_rgb, w: int, h: int, host_scan) -> int:
        """jpegb200_encode_host: host RGB buffer -> host scan buffer (H2D + kernels + D2H inside).
        Buffers are numpy arrays or (preferably pinned) CPU torch tensors.  Returns the byte count."""
        def ptr(x):
            return x.data_ptr() if hasattr(x, "data_ptr") else x.ctypes.data
        cap = host_scan.numel() if hasattr(host_scan, "numel") else host_scan.size
        n = C.c_uint64(0)
        check(self.lib.jpegb200_encode_host(self.handle, ptr(host_rgb), w, h, ptr(host_scan), cap, C.byref(n),
                                            self._stream()), "encode_host")
        return int(n.value)

    def encode_bmp_to_jpeg(self, bmp_bytes, out=None) -> bytes:
        """BMP file image (bytes / uint8 array / pinned tensor) -> complete JPEG file bytes.  The BMP pixel
        array is read in place on the device (no host-side BGR->RGB / flip pass)."""
        if isinstance(bmp_bytes, (bytes, bytearray)):
            bmp_bytes = np.frombuffer(bytes(bmp_bytes), np.uint8)
        n_in = bmp_bytes.numel() if hasattr(bmp_bytes, "numel") else bmp_bytes.size
        ptr = bmp_bytes.data_ptr() if hasattr(bmp_bytes, "data_ptr") else bmp_bytes.ctypes.data
        if out is None:
            out = np.empty(n_in + 1024, np.uint8)
        optr = out.data_ptr() if hasattr(out, "data_ptr") else out.ctypes.data
        cap = out.numel() if hasattr(out, "numel") else out.size
        n, w, h = C.c_uint64(0), C.c_int(0), C.c_int(0)
        check(self.lib.jpegb200_encode_bmp_to_jpeg_host(self.handle, ptr, n_in, optr, cap, C.byref(n), C.byref(w), C.byref(h),
                                                        self._stream()), "encode_bmp_to_jpeg_host")
        res = out[: n.value]
        return res.numpy().tobytes() if hasattr(res, "numpy") else res.tobytes()

    def set_profiling(self, on: bool):
        check(self.lib.jpegb200_encoder_set_profiling(self.handle, int(on)), "set_profiling")

    def kernel_times(self, reset: bool = False) -> dict:
        ms = (C.c_double * 8)()
        calls = (C.c_uint64 * 8)()
        check(self.lib.jpegb200_encoder_kernel_times(self.handle, ms, calls, int(reset)), "kernel_times")
        return {"ms": list(ms), "calls": [int(c) for c in calls],
                "names": ["fused_block", "merge_stuff", "batch_layout", "batch_compact", "frame_files", "strip_entropy", "", ""]}

    def coefficients(self, nblocks: int) -> np.ndarray:
        out = np.empty((nblocks, 64), np.int16)
        check(self.lib.jpegb200_encoder_read_coefficients(self.handle, out.ctypes.data, nblocks), "read_coefficients")
        return out

    def block_bits(self, nblocks: int) -> np.ndarray:
        out = np.empty(nblocks, np.uint32)
        check(self.lib.jpegb200_encoder_read_block_bits(self.handle, out.ctypes.data, nblocks), "read_block_bits")
        return out

    # ---- stripes -------------------------------------------------------------------
    def stripe_analyze(self, d_rgb, w: int, stripe_h: int, halo_rows: int = 0) -> dict:
        """K1 over the stripe's rows (+ halo rows of the next stripe) -> boundary summary."""
        s = StripeSummary()
        check(self.lib.jpegb200_stripe_analyze(self.handle, d_rgb.data_ptr(), w, stripe_h, halo_rows, C.byref(s),
                                               self._stream()), "stripe_analyze")
        return {"first_dc": s.first_dc, "last_dc": s.last_dc, "bits_pred0": s.bits_pred0}

    def stripe_encode(self, dc_pred: int, bit_begin: int, scan) -> int:
        """K2 at the stripe's true predictor and global bit offset -> number of stuffed bytes."""
        n = C.c_uint64(0)
        check(self.lib.jpegb200_stripe_encode(self.handle, dc_pred, bit_begin, scan.data_ptr(), scan.numel(),
                                              C.byref(n), self._stream()), "stripe_encode")
        return int(n.value)

    def stripe_analyze_device(self, d_rgb, w: int, stripe_h: int, halo_rows: int, d_summary):
        """Like stripe_analyze, but the 16-byte summary stays on the device (d_summary: int64[2] CUDA tensor or a
        slice of the all-gather buffer).  Asynchronous on the current stream."""
        check(self.lib.jpegb200_stripe_analyze_device(self.handle, d_rgb.data_ptr(), w, stripe_h, halo_rows,
                                                      d_summary.data_ptr(), self._stream()), "stripe_analyze_device")

    def stripe_encode_device(self, d_all, world: int, rank: int, scan, d_info):
        """Resolve this rank's predictor / bit offset from all ranks' summaries (d_all: int64[world, 2] on the
        device) and run the merge kernel; d_info (int64[2] on the device) receives [0, stuffed bytes]."""
        check(self.lib.jpegb200_stripe_encode_device(self.handle, d_all.data_ptr(), world, rank, scan.data_ptr(), scan.numel(),
                                                     d_info.data_ptr(), self._stream()), "stripe_encode_device")

    # ---- synthetic inputs on the device --------------------------------------------
    def synth(self, w: int, h: int, count: int = 1, seed0: int = 1, amp: int = 20):
        t = self.torch
        out = t.empty((count, h, w, 3), dtype=t.uint8, device=f"cuda:{self.device}")
        check(self.lib.jpegb200_synth_rgb_device(out.data_ptr(), w, h, count, 0, seed0, amp, self._stream()), "synth")
        return out
