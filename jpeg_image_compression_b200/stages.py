"""Host-side mirror of the reference's operator interface for the hot path.

Same function names and argument meaning as natural_c/include/*.h; every function goes
through the C ABI of libjpegb200.so (host buffers in, host buffers out, kernels on the GPU)
and returns numpy copies.  ``None`` in / failure => ``None`` out, like the reference's NULL.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import BMPImage, PlaneImage, RLEData, ZigZagData, load_library

SYMBOL_DTYPE = np.dtype({"names": ["symbol", "code", "codeBits"], "formats": [np.uint8, np.uint16, np.uint8],
                         "offsets": [0, 2, 4], "itemsize": 6})       # RLESymbol, include/rle.h:8-14


def _plane(arr: np.ndarray) -> PlaneImage:
    h, w = arr.shape
    return PlaneImage(w, h, arr.ctypes.data)


def _copy(ptr, nbytes, dtype, shape):
    return np.frombuffer(C.string_at(ptr, nbytes), dtype).reshape(shape).copy()


def loadBMPImage(path: str):
    """io/bmp_handler.c:15 -> (h, w, 3) uint8 RGB top-down, or None."""
    L = load_library()
    img = L.loadBMPImage(path.encode())
    if not img:
        return None
    w, h = img.contents.width, img.contents.height
    out = _copy(img.contents.data, w * h * 3, np.uint8, (h, w, 3))
    L.freeBMPImage(img)
    return out


def convertBMPToJPEGGrayscale(rgb):
    if rgb is None:
        return None
    L = load_library()
    rgb = np.ascontiguousarray(rgb, np.uint8)
    h, w, _ = rgb.shape
    bmp = BMPImage(w, h, rgb.ctypes.data)
    y = L.convertBMPToJPEGGrayscale(C.byref(bmp))
    if not y:
        return None
    out = _copy(y.contents.data, y.contents.width * y.contents.height, np.uint8, (y.contents.height, y.contents.width))
    L.freeYImage(y)
    return out


def centerYImage(y):
    if y is None:
        return None
    L = load_library()
    y = np.ascontiguousarray(y, np.uint8)
    src = _plane(y)
    c = L.centerYImage(C.byref(src))
    if not c:
        return None
    out = _copy(c.contents.data, y.size, np.int8, y.shape)
    L.freeCenteredYImage(c)
    return out


def computeDCTBlock(block):
    L = load_library()
    block = np.ascontiguousarray(block, np.int8).reshape(8, 8)
    out = np.zeros((8, 8), np.float32)
    L.computeDCTBlock(block.ctypes.data, out.ctypes.data)
    return out


def performDCT(centered):
    if centered is None:
        return None
    L = load_library()
    centered = np.ascontiguousarray(centered, np.int8)
    src = _plane(centered)
    d = L.performDCT(C.byref(src))
    if not d:
        return None
    out = _copy(d.contents.data, centered.size * 4, np.float32, centered.shape)
    L.freeDCTImage(d)
    return out


def quantizeImage(dct):
    if dct is None:
        return None
    L = load_library()
    dct = np.ascontiguousarray(dct, np.float32)
    src = _plane(dct)
    q = L.quantizeImage(C.byref(src))
    if not q:
        return None
    out = _copy(q.contents.data, dct.size * 2, np.int16, dct.shape)
    L.freeQuantizedImage(q)
    return out


def performZigZag(quant):
    if quant is None:
        return None
    L = load_library()
    quant = np.ascontiguousarray(quant, np.int16)
    src = _plane(quant)
    z = L.performZigZag(C.byref(src))
    if not z:
        return None
    nb = z.contents.totalBlocks
    out = _copy(z.contents.data, nb * 128, np.int16, (nb, 64))
    L.freeZigZagData(z)
    return out


def performRLE(zigzag, blocks_w: int = 0, blocks_h: int = 0):
    if zigzag is None:
        return None
    L = load_library()
    zigzag = np.ascontiguousarray(zigzag, np.int16).reshape(-1, 64)
    nb = zigzag.shape[0]
    src = ZigZagData(blocks_w or nb, blocks_h or 1, nb, zigzag.ctypes.data)
    r = L.performRLE(C.byref(src))
    if not r:
        return None
    out = np.frombuffer(C.string_at(r.contents.data, r.contents.count * 6), SYMBOL_DTYPE).copy()
    L.freeRLEData(r)
    return out


def encodeHuffman(symbols, total_blocks: int):
    if symbols is None:
        return None
    L = load_library()
    symbols = np.ascontiguousarray(symbols)
    assert symbols.dtype.itemsize == 6
    src = RLEData(symbols.ctypes.data if symbols.size else None, symbols.size, symbols.size)
    b = L.encodeHuffman(C.byref(src), int(total_blocks))
    if not b:
        return None
    out = C.string_at(b.contents.data, b.contents.size) if b.contents.size else b""
    L.freeJpegEncoderBuffer(b)
    return out


def encode_scan_host(rgb, with_first_block: bool = False):
    """The fused entry (jpegb200_encode_scan): host RGB -> stuffed scan bytes."""
    L = load_library()
    rgb = np.ascontiguousarray(rgb, np.uint8)
    h, w, _ = rgb.shape
    bmp = BMPImage(w, h, rgb.ctypes.data)
    first = np.zeros(64, np.int16)
    b = L.jpegb200_encode_scan_dbg(C.byref(bmp), first.ctypes.data)
    if not b:
        raise _lib.JpegB200Error("jpegb200_encode_scan failed: " + _lib.last_error())
    out = C.string_at(b.contents.data, b.contents.size) if b.contents.size else b""
    L.freeJpegEncoderBuffer(b)
    return (out, first.reshape(8, 8)) if with_first_block else out


def jfif_header(w: int, h: int) -> bytes:
    L = load_library()
    buf = (C.c_uint8 * 328)()
    n = L.jpegb200_jfif_header(w, h, buf)
    return bytes(buf[:n])


def saveJPEGGrayscale(path: str, rgb) -> bool:
    """io/jpeg_handler.c:119: writes the complete .jpg (headers + scan + EOI)."""
    L = load_library()
    rgb = np.ascontiguousarray(rgb, np.uint8)
    h, w, _ = rgb.shape
    bmp = BMPImage(w, h, rgb.ctypes.data)
    return bool(L.saveJPEGGrayscale(path.encode(), C.byref(bmp)))
