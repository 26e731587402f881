"""ctypes bindings for the CHECKERS -- test infrastructure only.

* ``Oracle``  -> oracle/liboracle.so, our plain-C restatement (jpeg_oracle.c).
* ``Ref``     -> oracle/_ref/libnaturalc_ref.so, the unmodified reference objects
                compiled by oracle/Makefile from /root/reference/natural_c (present
                in the build container; the prebuilt .so travels to the GPU box).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product package never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "liboracle.so")
REF_SO = os.path.join(HERE, "_ref", "libnaturalc_ref.so")
REF_APP = os.path.join(HERE, "_ref", "jpeg_compression_app")


def build(quiet: bool = True) -> None:
    """Compile liboracle.so (and oracle/_ref when /root/reference is present)."""
    subprocess.run(["make", "-C", HERE], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


def pad8(n: int) -> int:
    return (n + 7) & ~7


SYMBOL_DTYPE = np.dtype({"names": ["symbol", "amplitude", "nbits"],
                         "formats": [np.uint8, np.uint16, np.uint8],
                         "offsets": [0, 2, 4], "itemsize": 6})


def _u8p(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint8))


class Oracle:
    """Restated algorithm (see jpeg_oracle.h for the reference line citations)."""

    def __init__(self, path: str = ORACLE_SO):
        if not os.path.exists(path):
            build()
        self.lib = L = C.CDLL(path)
        vp, sz = C.c_void_p, C.c_size_t
        L.orc_luma_pad.argtypes = [vp, C.c_int, C.c_int, vp]
        L.orc_level_shift.argtypes = [vp, sz, vp]
        L.orc_fdct_block.argtypes = [vp, vp]
        L.orc_fdct_image.argtypes = [vp, C.c_int, C.c_int, vp]
        L.orc_quantize.argtypes = [vp, C.c_int, C.c_int, vp]
        L.orc_zigzag.argtypes = [vp, C.c_int, C.c_int, vp]
        L.orc_rle.argtypes = [vp, sz, vp]
        L.orc_rle.restype = sz
        L.orc_huffman.argtypes = [vp, sz, sz, vp]
        L.orc_huffman.restype = sz
        L.orc_block_bits.argtypes = [vp, C.c_int16]
        L.orc_block_bits.restype = C.c_uint32
        L.orc_coefficients.argtypes = [vp, C.c_int, C.c_int, vp]
        L.orc_coefficients.restype = sz
        L.orc_encode_scan.argtypes = [vp, C.c_int, C.c_int, C.POINTER(vp)]
        L.orc_encode_scan.restype = sz
        L.orc_free.argtypes = [vp]
        L.orc_encode_scan_synth_banded.argtypes = [C.c_int, C.c_int, C.c_uint32, C.c_int, C.c_int, C.c_int, C.POINTER(vp)]
        L.orc_encode_scan_synth_banded.restype = sz
        L.orc_jfif_header.argtypes = [C.c_int, C.c_int, vp]
        L.orc_jfif_header.restype = sz
        L.orc_synth_rgb.argtypes = [C.c_int, C.c_int, C.c_uint32, C.c_int, vp]

    # -- stages ------------------------------------------------------------
    def luma_pad(self, rgb: np.ndarray) -> np.ndarray:
        h, w, _ = rgb.shape
        rgb = np.ascontiguousarray(rgb, np.uint8)
        y = np.empty((pad8(h), pad8(w)), np.uint8)
        self.lib.orc_luma_pad(rgb.ctypes.data, w, h, y.ctypes.data)
        return y

    def level_shift(self, y: np.ndarray) -> np.ndarray:
        y = np.ascontiguousarray(y, np.uint8)
        out = np.empty(y.shape, np.int8)
        self.lib.orc_level_shift(y.ctypes.data, y.size, out.ctypes.data)
        return out

    def fdct_block(self, blk: np.ndarray) -> np.ndarray:
        blk = np.ascontiguousarray(blk, np.int8).reshape(64)
        out = np.empty(64, np.float32)
        self.lib.orc_fdct_block(blk.ctypes.data, out.ctypes.data)
        return out.reshape(8, 8)

    def fdct_image(self, centered: np.ndarray) -> np.ndarray:
        centered = np.ascontiguousarray(centered, np.int8)
        hp, wp = centered.shape
        out = np.empty((hp, wp), np.float32)
        self.lib.orc_fdct_image(centered.ctypes.data, wp, hp, out.ctypes.data)
        return out

    def quantize(self, coef: np.ndarray) -> np.ndarray:
        coef = np.ascontiguousarray(coef, np.float32)
        hp, wp = coef.shape
        out = np.empty((hp, wp), np.int16)
        self.lib.orc_quantize(coef.ctypes.data, wp, hp, out.ctypes.data)
        return out

    def zigzag(self, q: np.ndarray) -> np.ndarray:
        q = np.ascontiguousarray(q, np.int16)
        hp, wp = q.shape
        out = np.empty(((hp // 8) * (wp // 8), 64), np.int16)
        self.lib.orc_zigzag(q.ctypes.data, wp, hp, out.ctypes.data)
        return out

    def rle(self, zz: np.ndarray) -> np.ndarray:
        zz = np.ascontiguousarray(zz, np.int16).reshape(-1, 64)
        n = self.lib.orc_rle(zz.ctypes.data, zz.shape[0], None)
        out = np.zeros(n, SYMBOL_DTYPE)
        self.lib.orc_rle(zz.ctypes.data, zz.shape[0], out.ctypes.data)
        return out

    def huffman(self, sym: np.ndarray, nblocks: int) -> np.ndarray:
        sym = np.ascontiguousarray(sym)
        assert sym.dtype == SYMBOL_DTYPE
        n = self.lib.orc_huffman(sym.ctypes.data, sym.size, nblocks, None)
        out = np.empty(n, np.uint8)
        self.lib.orc_huffman(sym.ctypes.data, sym.size, nblocks, out.ctypes.data)
        return out

    def block_bits(self, zz: np.ndarray) -> np.ndarray:
        """Per-block Huffman bit cost with the image-wide DC chain (pred 0 first)."""
        zz = np.ascontiguousarray(zz, np.int16).reshape(-1, 64)
        out = np.empty(zz.shape[0], np.uint32)
        prev = 0
        for b in range(zz.shape[0]):
            out[b] = self.lib.orc_block_bits(zz[b].ctypes.data, prev)
            prev = int(zz[b, 0])
        return out

    # -- composed ------------------------------------------------------------
    def coefficients(self, rgb: np.ndarray) -> np.ndarray:
        h, w, _ = rgb.shape
        rgb = np.ascontiguousarray(rgb, np.uint8)
        nb = (pad8(h) // 8) * (pad8(w) // 8)
        zz = np.empty((nb, 64), np.int16)
        got = self.lib.orc_coefficients(rgb.ctypes.data, w, h, zz.ctypes.data)
        assert got == nb
        return zz

    def encode_scan(self, rgb: np.ndarray) -> bytes:
        h, w, _ = rgb.shape
        rgb = np.ascontiguousarray(rgb, np.uint8)
        p = C.c_void_p()
        n = self.lib.orc_encode_scan(rgb.ctypes.data, w, h, C.byref(p))
        try:
            return C.string_at(p.value, n) if n else b""
        finally:
            self.lib.orc_free(p)

    def encode_scan_synth_banded(self, w: int, h: int, seed: int = 1, amp: int = 20, band_block_rows: int = 8,
                                 threads: int = 0) -> bytes:
        """Streaming encode of the synthetic image of any size (also beyond the reference's 715 Mpixel
        index limit): orc_encode_scan_synth_banded.  threads=0: all host cores."""
        import os
        p = C.c_void_p()
        n = self.lib.orc_encode_scan_synth_banded(w, h, seed, amp, band_block_rows, threads or (os.cpu_count() or 1), C.byref(p))
        try:
            if not p.value:
                raise MemoryError("orc_encode_scan_synth_banded failed")
            return C.string_at(p.value, n)
        finally:
            self.lib.orc_free(p)

    def jfif_header(self, w: int, h: int) -> bytes:
        buf = (C.c_uint8 * 328)()
        n = self.lib.orc_jfif_header(w, h, buf)
        return bytes(buf[:n])

    def encode_file_bytes(self, rgb: np.ndarray) -> bytes:
        h, w, _ = rgb.shape
        return self.jfif_header(w, h) + self.encode_scan(rgb) + b"\xff\xd9"

    def synth_rgb(self, w: int, h: int, seed: int = 1, amp: int = 20) -> np.ndarray:
        out = np.empty((h, w, 3), np.uint8)
        self.lib.orc_synth_rgb(w, h, seed, amp, out.ctypes.data)
        return out


# ---------------------------------------------------------------------------
# the unmodified reference objects


class _Img(C.Structure):          # BMPImage / YImage / CenteredYImage / DCTImage / QuantizedImage
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("data", C.c_void_p)]


class _ZigZag(C.Structure):       # include/zigzag.h:7-12
    _fields_ = [("numBlocksW", C.c_int), ("numBlocksH", C.c_int), ("totalBlocks", C.c_int),
                ("data", C.c_void_p)]


class _Rle(C.Structure):          # include/rle.h:17-21
    _fields_ = [("data", C.c_void_p), ("count", C.c_size_t), ("capacity", C.c_size_t)]


class _Buf(C.Structure):          # include/huffman.h:9-13
    _fields_ = [("data", C.c_void_p), ("size", C.c_size_t), ("capacity", C.c_size_t)]


def have_ref() -> bool:
    return os.path.exists(REF_SO)


class Ref:
    """The reference's own stage functions, called through its own struct API."""

    def __init__(self, path: str = REF_SO):
        self.lib = L = C.CDLL(path)
        P = C.POINTER
        L.loadBMPImage.argtypes = [C.c_char_p]
        L.loadBMPImage.restype = P(_Img)
        L.freeBMPImage.argtypes = [P(_Img)]
        L.convertBMPToJPEGGrayscale.argtypes = [P(_Img)]
        L.convertBMPToJPEGGrayscale.restype = P(_Img)
        L.centerYImage.argtypes = [P(_Img)]
        L.centerYImage.restype = P(_Img)
        L.performDCT.argtypes = [P(_Img)]
        L.performDCT.restype = P(_Img)
        L.computeDCTBlock.argtypes = [C.c_void_p, C.c_void_p]
        L.quantizeImage.argtypes = [P(_Img)]
        L.quantizeImage.restype = P(_Img)
        L.performZigZag.argtypes = [P(_Img)]
        L.performZigZag.restype = P(_ZigZag)
        L.performRLE.argtypes = [P(_ZigZag)]
        L.performRLE.restype = P(_Rle)
        L.encodeHuffman.argtypes = [P(_Rle), C.c_int]
        L.encodeHuffman.restype = P(_Buf)
        for name, t in (("freeYImage", _Img), ("freeCenteredYImage", _Img), ("freeDCTImage", _Img),
                        ("freeQuantizedImage", _Img), ("freeZigZagData", _ZigZag),
                        ("freeRLEData", _Rle), ("freeJpegEncoderBuffer", _Buf)):
            getattr(L, name).argtypes = [P(t)]
        L.saveJPEGGrayscale.argtypes = [C.c_char_p, P(_Img)]
        L.saveJPEGGrayscale.restype = C.c_bool

    def load_bmp(self, path: str) -> np.ndarray | None:
        img = self.lib.loadBMPImage(path.encode())
        if not img:
            return None
        w, h = img.contents.width, img.contents.height
        out = np.frombuffer(C.string_at(img.contents.data, w * h * 3), np.uint8).reshape(h, w, 3).copy()
        self.lib.freeBMPImage(img)
        return out

    def fdct_block(self, blk: np.ndarray) -> np.ndarray:
        blk = np.ascontiguousarray(blk, np.int8).reshape(64)
        out = np.empty(64, np.float32)
        self.lib.computeDCTBlock(blk.ctypes.data, out.ctypes.data)
        return out.reshape(8, 8)

    def stages(self, rgb: np.ndarray, upto: str = "scan") -> dict:
        """Run the reference stage chain, returning copies of every intermediate."""
        L = self.lib
        h, w, _ = rgb.shape
        rgb = np.ascontiguousarray(rgb, np.uint8)
        bmp = _Img(w, h, rgb.ctypes.data)
        out = {}
        y = L.convertBMPToJPEGGrayscale(C.byref(bmp))
        wp, hp = y.contents.width, y.contents.height
        n = wp * hp
        out["y"] = np.frombuffer(C.string_at(y.contents.data, n), np.uint8).reshape(hp, wp).copy()
        c = L.centerYImage(y)
        out["centered"] = np.frombuffer(C.string_at(c.contents.data, n), np.int8).reshape(hp, wp).copy()
        d = L.performDCT(c)
        out["dct"] = np.frombuffer(C.string_at(d.contents.data, n * 4), np.float32).reshape(hp, wp).copy()
        q = L.quantizeImage(d)
        out["quant"] = np.frombuffer(C.string_at(q.contents.data, n * 2), np.int16).reshape(hp, wp).copy()
        z = L.performZigZag(q)
        nb = z.contents.totalBlocks
        out["zigzag"] = np.frombuffer(C.string_at(z.contents.data, nb * 128), np.int16).reshape(nb, 64).copy()
        out["blocks_w"], out["blocks_h"] = z.contents.numBlocksW, z.contents.numBlocksH
        if upto == "scan":
            r = L.performRLE(z)
            ns = r.contents.count
            out["symbols"] = np.frombuffer(C.string_at(r.contents.data, ns * 6), SYMBOL_DTYPE).copy()
            b = L.encodeHuffman(r, nb)
            out["scan"] = C.string_at(b.contents.data, b.contents.size) if b.contents.size else b""
            L.freeJpegEncoderBuffer(b)
            L.freeRLEData(r)
        L.freeZigZagData(z)
        L.freeQuantizedImage(q)
        L.freeDCTImage(d)
        L.freeCenteredYImage(c)
        L.freeYImage(y)
        return out

    def encode_scan(self, rgb: np.ndarray) -> bytes:
        return self.stages(rgb)["scan"]


def write_bmp(path: str, rgb: np.ndarray, top_down: bool = False, header_size: int = 40) -> None:
    """Write a 24-bit uncompressed BMP that the reference loader accepts
    (io/bmp_handler.c:30-49,75,88,115-117).  header_size 40 (BITMAPINFOHEADER),
    108 (V4) or 124 (V5) only changes biSize / bfOffBits; extra bytes are zero."""
    h, w, _ = rgb.shape
    pitch = (w * 3 + 3) & ~3
    rows = rgb[:, :, ::-1]                       # RGB -> BGR
    if not top_down:
        rows = rows[::-1]
    body = np.zeros((h, pitch), np.uint8)
    body[:, : w * 3] = rows.reshape(h, w * 3)
    off = 14 + header_size
    hdr = bytearray(off)
    hdr[0:2] = b"BM"
    hdr[2:6] = (off + body.size).to_bytes(4, "little")
    hdr[10:14] = off.to_bytes(4, "little")
    hdr[14:18] = header_size.to_bytes(4, "little")
    hdr[18:22] = w.to_bytes(4, "little", signed=True)
    hdr[22:26] = (-h if top_down else h).to_bytes(4, "little", signed=True)
    hdr[26:28] = (1).to_bytes(2, "little")
    hdr[28:30] = (24).to_bytes(2, "little")
    hdr[34:38] = body.size.to_bytes(4, "little")
    hdr[38:42] = (2835).to_bytes(4, "little")
    hdr[42:46] = (2835).to_bytes(4, "little")
    with open(path, "wb") as f:
        f.write(bytes(hdr))
        f.write(body.tobytes())
