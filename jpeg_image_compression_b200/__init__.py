"""jpeg_image_compression_b200 -- B200 (sm_100a) implementation of the natural_c grayscale
JPEG encode hot path of strbac-damjan/jpeg-image-compression.

The product is ``libjpegb200.so`` (hand-written CUDA kernels behind the C ABI declared in
``include/jpegb200.h``) plus the drop-in CLI ``jpeg_compression_app``.  This Python package
is the thin host layer used by the tests and the benchmark: ctypes bindings, torch tensors
for device memory / streams, and torch.distributed for the multi-GPU exchanges.

There is no CPU fallback: importing works anywhere, but every compute entry point raises
``JpegB200Error`` unless the CUDA library loads and a B200 is present.
"""
from ._lib import JpegB200Error, build_library, library_path, load_library  # noqa: F401
from .stages import (  # noqa: F401
    centerYImage, computeDCTBlock, convertBMPToJPEGGrayscale, encodeHuffman, encode_scan_host,
    jfif_header, loadBMPImage, performDCT, performRLE, performZigZag, quantizeImage, saveJPEGGrayscale,
)
from .encoder import DeviceEncoder  # noqa: F401
from .synth import synth_rgb  # noqa: F401
