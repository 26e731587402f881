"""Oracle-based stand-in for the CUDA stripe backend (tests only).

Implements the same two-phase contract as DeviceEncoder.stripe_analyze / stripe_encode with the
CPU oracle, bit by bit, so the host-side stitching logic (jpeg_image_compression_b200/stripes.py)
can be exercised over gloo without a GPU."""
import numpy as np

from jpeg_image_compression_b200.stripes import dc_cost


def _unstuff(data: bytes) -> bytes:
    out = bytearray()
    i = 0
    while i < len(data):
        out.append(data[i])
        if data[i] == 0xFF:
            i += 1                      # skip the stuffed 0x00
        i += 1
    return bytes(out)


class OracleStripeBackend:
    def __init__(self, oracle):
        self.o = oracle

    def stripe_analyze(self, rows: np.ndarray, width: int, owned: int, halo: int) -> dict:
        assert rows.shape[0] == owned + halo and rows.shape[1] == width
        zz = self.o.coefficients(rows)                       # owned + halo block rows, bottom rows replicated
        bw = (width + 7) // 8
        self.nb_owned = bw * ((owned + 7) // 8)
        self.zz = zz
        bits = int(self.o.block_bits(zz[: self.nb_owned]).sum())
        return {"first_dc": int(zz[0, 0]), "last_dc": int(zz[self.nb_owned - 1, 0]), "bits_pred0": bits}

    def _bits(self, zz: np.ndarray, pred: int) -> str:
        """MSB-first bit string of blocks `zz` with the DC chain starting from `pred`."""
        lead = np.zeros((1, 64), np.int16)
        lead[0, 0] = pred                                    # dummy block: DC = pred, no AC
        allz = np.concatenate([lead, zz])
        nbits = int(self.o.block_bits(allz).sum())
        raw = _unstuff(self.o.huffman(self.o.rle(allz), allz.shape[0]).tobytes())
        s = "".join(f"{b:08b}" for b in raw)[:nbits]
        return s[dc_cost(pred) + 4:]                         # drop the dummy block: DC symbol + EOB '1010'

    def stripe_encode(self, pred: int, bit_begin: int, scan) -> int:
        own = self._bits(self.zz[: self.nb_owned], pred)
        halo_zz = self.zz[self.nb_owned:]
        tail = self._bits(halo_zz[:2], int(self.zz[self.nb_owned - 1, 0])) if len(halo_zz) else ""
        begin, end = bit_begin, bit_begin + len(own)
        b0, b1 = (begin + 7) // 8, (end + 7) // 8            # bytes whose first bit lies in [begin, end)
        stream = "0" * (begin - (begin // 8) * 8) + own + tail
        stream += "0" * 16                                   # zero padding of the final byte (huffman.c:65-81)
        base = begin // 8
        out = bytearray()
        for byte in range(b0, b1):
            v = int(stream[(byte - base) * 8:(byte - base) * 8 + 8], 2)
            out.append(v)
            if v == 0xFF:
                out.append(0)
        scan[: len(out)] = np.frombuffer(bytes(out), np.uint8)
        return len(out)
