"""ctypes loader for libjpegb200.so (built in-tree by csrc/Makefile)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.environ.get("JPEGB200_LIB") or os.path.join(HERE, "libjpegb200.so")   # override: tracing build
_lib = None


class JpegB200Error(RuntimeError):
    pass


def library_path() -> str:
    return _LIB_PATH


def build_library(quiet: bool = True) -> str:
    """Compile the CUDA extension for sm_100a (nvcc cross-compiles without a GPU)."""
    subprocess.run(["make", "-C", os.path.join(HERE, "csrc")], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)
    return _LIB_PATH


# ---- reference-compatible structs (include/jpegb200.h) ---------------------------------
class BMPImage(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("data", C.c_void_p)]


class PlaneImage(C.Structure):      # YImage / CenteredYImage / DCTImage / QuantizedImage
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("data", C.c_void_p)]


class ZigZagData(C.Structure):
    _fields_ = [("numBlocksW", C.c_int), ("numBlocksH", C.c_int), ("totalBlocks", C.c_int), ("data", C.c_void_p)]


class RLEData(C.Structure):
    _fields_ = [("data", C.c_void_p), ("count", C.c_size_t), ("capacity", C.c_size_t)]


class JpegEncoderBuffer(C.Structure):
    _fields_ = [("data", C.c_void_p), ("size", C.c_size_t), ("capacity", C.c_size_t)]


class Batch(C.Structure):
    _fields_ = [("d_rgb", C.c_void_p), ("width", C.c_int32), ("height", C.c_int32), ("count", C.c_int32),
                ("image_stride", C.c_uint64)]


class Stats(C.Structure):
    _fields_ = [("blocks", C.c_uint64), ("flagged_coefficients", C.c_uint64), ("kernel_launches", C.c_uint64),
                ("packed_bytes", C.c_uint64)]


class StripeSummary(C.Structure):
    _fields_ = [("first_dc", C.c_int16), ("last_dc", C.c_int16), ("reserved", C.c_uint32), ("bits_pred0", C.c_uint64)]


EXPORTED_FUNCTIONS = [
    # reference core-stage API
    "convertBMPToJPEGGrayscale", "centerYImage", "freeCenteredYImage", "computeDCTBlock", "performDCT",
    "freeDCTImage", "quantizeImage", "freeQuantizedImage", "performZigZag", "freeZigZagData", "performRLE",
    "freeRLEData", "encodeHuffman", "freeJpegEncoderBuffer",
    # host I/O
    "loadBMPImage", "freeBMPImage", "saveJPEGGrayscale", "freeYImage",
    # fused + device API
    "jpegb200_encode_scan", "jpegb200_encode_scan_dbg", "jpegb200_jfif_header", "jpegb200_device_count",
    "jpegb200_last_error", "jpegb200_encoder_create", "jpegb200_encoder_destroy", "jpegb200_encoder_set_dct_mode", "jpegb200_encoder_set_concurrency", "jpegb200_encoder_check_guards",
    "jpegb200_encoder_set_bytes_per_block", "jpegb200_encode_batch_device", "jpegb200_encode_batch_files_device", "jpegb200_encoder_status",
    "jpegb200_encoder_stats", "jpegb200_encoder_read_coefficients", "jpegb200_encoder_read_block_bits", "jpegb200_encoder_read_trace", "jpegb200_encoder_read_k1_trace", "jpegb200_encoder_launch_shape",
    "jpegb200_encoder_set_profiling", "jpegb200_encoder_kernel_times", "jpegb200_encode_host",
    "jpegb200_encode_bmp_to_jpeg_host",
    "jpegb200_stripe_analyze", "jpegb200_stripe_encode", "jpegb200_stripe_analyze_device", "jpegb200_stripe_encode_device",
    "jpegb200_synth_rgb_device",
]
EXPORTED_DATA = ["std_luminance_quant_tbl", "std_dc_luminance_nrcodes", "std_dc_luminance_values",
                 "std_ac_luminance_nrcodes", "std_ac_luminance_values"]


def load_library():
    """Load libjpegb200.so and declare signatures.  Raises if the extension is missing:
    there is deliberately no fallback implementation."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise JpegB200Error(f"{_LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                            "(there is no CPU fallback)")
    L = C.CDLL(_LIB_PATH)
    P, vp, u64 = C.POINTER, C.c_void_p, C.c_uint64
    L.convertBMPToJPEGGrayscale.argtypes = [P(BMPImage)]
    L.convertBMPToJPEGGrayscale.restype = P(PlaneImage)
    L.centerYImage.argtypes = [P(PlaneImage)]
    L.centerYImage.restype = P(PlaneImage)
    L.computeDCTBlock.argtypes = [vp, vp]
    L.computeDCTBlock.restype = None
    L.performDCT.argtypes = [P(PlaneImage)]
    L.performDCT.restype = P(PlaneImage)
    L.quantizeImage.argtypes = [P(PlaneImage)]
    L.quantizeImage.restype = P(PlaneImage)
    L.performZigZag.argtypes = [P(PlaneImage)]
    L.performZigZag.restype = P(ZigZagData)
    L.performRLE.argtypes = [P(ZigZagData)]
    L.performRLE.restype = P(RLEData)
    L.encodeHuffman.argtypes = [P(RLEData), C.c_int]
    L.encodeHuffman.restype = P(JpegEncoderBuffer)
    for name, t in (("freeYImage", PlaneImage), ("freeCenteredYImage", PlaneImage), ("freeDCTImage", PlaneImage),
                    ("freeQuantizedImage", PlaneImage), ("freeZigZagData", ZigZagData), ("freeRLEData", RLEData),
                    ("freeJpegEncoderBuffer", JpegEncoderBuffer), ("freeBMPImage", BMPImage)):
        getattr(L, name).argtypes = [P(t)]
        getattr(L, name).restype = None
    L.loadBMPImage.argtypes = [C.c_char_p]
    L.loadBMPImage.restype = P(BMPImage)
    L.saveJPEGGrayscale.argtypes = [C.c_char_p, P(BMPImage)]
    L.saveJPEGGrayscale.restype = C.c_bool
    L.jpegb200_encode_scan.argtypes = [P(BMPImage)]
    L.jpegb200_encode_scan.restype = P(JpegEncoderBuffer)
    L.jpegb200_encode_scan_dbg.argtypes = [P(BMPImage), vp]
    L.jpegb200_encode_scan_dbg.restype = P(JpegEncoderBuffer)
    L.jpegb200_jfif_header.argtypes = [C.c_int, C.c_int, vp]
    L.jpegb200_jfif_header.restype = C.c_size_t
    L.jpegb200_device_count.restype = C.c_int
    L.jpegb200_last_error.restype = C.c_char_p
    L.jpegb200_encoder_create.argtypes = [C.c_int]
    L.jpegb200_encoder_create.restype = vp
    L.jpegb200_encoder_destroy.argtypes = [vp]
    L.jpegb200_encoder_destroy.restype = None
    L.jpegb200_encoder_set_dct_mode.argtypes = [vp, C.c_int]
    L.jpegb200_encoder_set_concurrency.argtypes = [vp, C.c_int]
    L.jpegb200_encoder_check_guards.argtypes = [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_int)]
    L.jpegb200_encoder_set_bytes_per_block.argtypes = [vp, C.c_int]
    L.jpegb200_encode_batch_device.argtypes = [vp, P(Batch), vp, u64, vp, vp]
    L.jpegb200_encode_batch_files_device.argtypes = [vp, P(Batch), vp, u64, vp, vp]
    L.jpegb200_encoder_status.argtypes = [vp, vp]
    L.jpegb200_encoder_stats.argtypes = [vp, P(Stats)]
    L.jpegb200_encoder_read_coefficients.argtypes = [vp, vp, u64]
    L.jpegb200_encoder_read_block_bits.argtypes = [vp, vp, u64]
    L.jpegb200_encoder_read_trace.argtypes = [vp, vp, u64]
    L.jpegb200_encoder_read_k1_trace.argtypes = [vp, vp, u64]
    L.jpegb200_encoder_launch_shape.argtypes = [vp, P(C.c_int), P(C.c_int)]
    L.jpegb200_encoder_set_profiling.argtypes = [vp, C.c_int]
    L.jpegb200_encoder_kernel_times.argtypes = [vp, P(C.c_double), P(u64), C.c_int]
    L.jpegb200_encode_host.argtypes = [vp, vp, C.c_int, C.c_int, vp, u64, P(u64), vp]
    L.jpegb200_encode_bmp_to_jpeg_host.argtypes = [vp, vp, u64, vp, u64, P(u64), P(C.c_int), P(C.c_int), vp]
    L.jpegb200_stripe_analyze.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, P(StripeSummary), vp]
    L.jpegb200_stripe_encode.argtypes = [vp, C.c_int16, u64, vp, u64, P(u64), vp]
    L.jpegb200_stripe_analyze_device.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, vp, vp]
    L.jpegb200_stripe_encode_device.argtypes = [vp, vp, C.c_int, C.c_int, vp, u64, vp, vp]
    L.jpegb200_synth_rgb_device.argtypes = [vp, C.c_int, C.c_int, C.c_int, u64, C.c_uint32, C.c_int, vp]
    _lib = L
    return L


def last_error() -> str:
    return (load_library().jpegb200_last_error() or b"").decode()


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise JpegB200Error(f"{what} failed (code {rc}): {last_error()}")
