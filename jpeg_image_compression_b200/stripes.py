"""Multi-GPU sharding of the encode path.

Two modes, one process per GPU:

* batch: images are independent (each has its own DC chain and bit stream), so a batch is cut
  into contiguous image ranges per rank -- no data-path collective at all (`shard_range`).

* MCU-row stripes of ONE image: rank r owns a contiguous range of 8-pixel block rows.  The only
  cross-stripe state of the reference's serial entropy coder is the DC predictor (rle.c:59-70)
  and the running bit offset (huffman.c:35-62), so one tiny all-gather suffices:

      analyze  (local)   fused block kernel -> {first_dc, last_dc, bits_pred0}
      all-gather         4 int64 per rank
      encode   (local)   scan + pack + stuff at the true predictor and global bit phase
      gather             stuffed byte counts, then the byte ranges themselves

  A byte of the stream is emitted by the rank that holds its first bit; to finish the byte its
  stream ends in, a rank also transforms the first block row of the next stripe (the "halo",
  at most 8 pixel rows that it reads from the source image itself), so no bits are exchanged.
  The concatenation of the ranks' outputs is byte-identical to the single-GPU encode.

The functions here are pure host logic over a small backend interface
(`stripe_analyze` / `stripe_encode`), which `DeviceEncoder` implements with the CUDA library;
the CPU tests drive the same logic over gloo with an oracle-based stand-in backend.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

# DC luminance code lengths per size class (canonical codes of the standard table,
# natural_c/src/core/jpeg_tables.c:14-20) -- only the LENGTH is needed to place bits.
_DC_COUNTS = (0, 1, 5, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0)
_DC_VALUES = (0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11)


def _dc_code_lengths() -> List[int]:
    out, vi = [0] * 16, 0
    for length in range(1, 17):
        for _ in range(_DC_COUNTS[length - 1]):
            out[_DC_VALUES[vi]] = length
            vi += 1
    return out


_DC_LEN = _dc_code_lengths()


def dc_cost(diff: int) -> int:
    """Bits of one DC-difference symbol: Huffman code of the size class + `size` amplitude bits
    (rle.c:9-22,72-76; huffman.c:145-153)."""
    size = abs(int(diff)).bit_length()
    return _DC_LEN[size] + size


def shard_range(n_items: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous [begin, end) share of n_items for `rank` (batch mode: images by rank)."""
    base, rem = divmod(n_items, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def stripe_plan(height: int, world: int) -> List[Tuple[int, int]]:
    """Per rank (first block row, number of block rows); 540 rows over 8 ranks ->
    68,68,68,68,67,67,67,67.  Ranks beyond the number of block rows get (x, 0)."""
    bh = (height + 7) // 8
    return [(b, e - b) for b, e in (shard_range(bh, world, r) for r in range(world))]


def stripe_rows(height: int, world: int, rank: int) -> Tuple[int, int, int]:
    """(first pixel row, pixel rows owned, halo pixel rows) of `rank`'s stripe.  The rank must
    hold rows [y0, y0 + owned + halo) of the source image contiguously."""
    row0, nrows = stripe_plan(height, world)[rank]
    y0 = row0 * 8
    if nrows == 0:
        return min(y0, height), 0, 0
    owned = min(height, (row0 + nrows) * 8) - y0
    halo = min(8, height - (y0 + owned))
    return y0, owned, max(halo, 0)


def resolve_offsets(summaries: Sequence[Optional[dict]]) -> List[Optional[Tuple[int, int]]]:
    """From every rank's analyze() summary (None for ranks without rows) derive each rank's
    (dc_predictor, global bit offset).  bits_pred0 assumed predictor 0 for the stripe's first
    block; the true cost replaces that one DC symbol."""
    out: List[Optional[Tuple[int, int]]] = []
    bit, pred = 0, 0
    for s in summaries:
        if s is None:
            out.append(None)
            continue
        out.append((pred, bit))
        bit += int(s["bits_pred0"]) - dc_cost(int(s["first_dc"])) + dc_cost(int(s["first_dc"]) - pred)
        pred = int(s["last_dc"])
    return out


def encode_striped_local(backends: Sequence, stripes: Sequence, width: int, height: int, scans: Sequence) -> bytes:
    """Single-process emulation of an N-rank striped encode (used on one GPU and in CPU tests).
    backends[r]: object with stripe_analyze / stripe_encode; stripes[r]: that rank's pixel rows
    (owned + halo) in whatever form the backend takes; scans[r]: its output buffer."""
    world = len(backends)
    summaries = []
    for r in range(world):
        _, owned, halo = stripe_rows(height, world, r)
        summaries.append(backends[r].stripe_analyze(stripes[r], width, owned, halo) if owned else None)
    plan = resolve_offsets(summaries)
    parts = []
    for r in range(world):
        if plan[r] is None:
            continue
        n = backends[r].stripe_encode(plan[r][0], plan[r][1], scans[r])
        parts.append(_to_bytes(scans[r], n))
    return b"".join(parts)


def _to_bytes(buf, n: int) -> bytes:
    if hasattr(buf, "cpu"):            # torch tensor (device or host)
        return buf[:n].cpu().numpy().tobytes()
    return bytes(buf[:n])


class StripedEncoder:
    """One rank of a striped encode over torch.distributed (NCCL on GPUs, gloo in CPU tests)."""

    def __init__(self, backend, group=None, device="cpu"):
        import torch.distributed as dist
        self.backend, self.group, self.device = backend, group, device
        self.dist = dist
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)

    def _all_gather_ints(self, values: Sequence[int]) -> List[List[int]]:
        import torch
        mine = torch.tensor(list(values), dtype=torch.int64, device=self.device)
        out = torch.empty(self.world * len(values), dtype=torch.int64, device=self.device)
        self.dist.all_gather_into_tensor(out, mine, group=self.group)
        return out.cpu().view(self.world, len(values)).tolist()

    def encode(self, stripe, width: int, height: int, scan) -> int:
        """Encode this rank's stripe (rows owned + halo, see stripe_rows) into `scan`.
        Returns the number of stuffed bytes this rank produced."""
        _, owned, halo = stripe_rows(height, self.world, self.rank)
        s = self.backend.stripe_analyze(stripe, width, owned, halo) if owned else None
        row = [1, s["first_dc"], s["last_dc"], s["bits_pred0"]] if s else [0, 0, 0, 0]
        table = self._all_gather_ints(row)
        summaries = [({"first_dc": t[1], "last_dc": t[2], "bits_pred0": t[3]} if t[0] else None) for t in table]
        mine = resolve_offsets(summaries)[self.rank]
        if mine is None:
            return 0
        return self.backend.stripe_encode(mine[0], mine[1], scan)

    def encode_on_device(self, stripe, width: int, height: int, scan, summaries=None, info=None):
        """The same encode with no host round trip between analyze, exchange and merge (CUDA backends only):
        the boundary summaries are reduced on the device, all-gathered as device tensors (NCCL) and resolved by a
        one-warp kernel.  Returns `info`, an int64[2] device tensor whose element 1 is this rank's stuffed byte count
        (read it after synchronising)."""
        import torch
        dev = scan.device
        if summaries is None:
            summaries = torch.zeros((self.world, 2), dtype=torch.int64, device=dev)
        if info is None:
            info = torch.zeros(2, dtype=torch.int64, device=dev)
        mine = torch.zeros(2, dtype=torch.int64, device=dev)
        _, owned, halo = stripe_rows(height, self.world, self.rank)
        if owned:
            self.backend.stripe_analyze_device(stripe, width, owned, halo, mine)
        self.dist.all_gather_into_tensor(summaries.view(-1), mine, group=self.group)
        if owned:
            self.backend.stripe_encode_device(summaries, self.world, self.rank, scan, info)
        else:
            info.zero_()
        return info

    def gather_exact(self, scan, nbytes_dev, out=None):
        """Stitch on rank 0 with size-exact transfers: sizes are all-gathered on the device (one host read), then
        every rank sends exactly its bytes and rank 0 receives them at their offsets.  Returns (tensor, total) on
        rank 0, (None, total) elsewhere."""
        import torch
        dev = scan.device
        sizes_d = torch.empty(self.world, dtype=torch.int64, device=dev)
        self.dist.all_gather_into_tensor(sizes_d, nbytes_dev.reshape(1), group=self.group)
        sizes = sizes_d.tolist()                                  # the one host synchronisation of the striped encode
        total = int(sum(sizes))
        if self.rank == 0:
            if out is None or out.numel() < total:
                out = torch.empty(max(total, 1), dtype=torch.uint8, device=dev)
            if sizes[0]:
                out[: sizes[0]].copy_(scan[: sizes[0]])
            reqs, off = [], sizes[0]
            for r in range(1, self.world):
                if sizes[r]:
                    reqs.append(self.dist.irecv(out[off: off + sizes[r]], src=r, group=self.group))
                off += sizes[r]
            for q in reqs:
                q.wait()
            return out, total
        if sizes[self.rank]:
            self.dist.send(scan[: sizes[self.rank]], dst=0, group=self.group)
        return None, total

    def gather(self, scan, nbytes: int) -> Optional[bytes]:
        """Stitch: every rank's stuffed bytes, in rank order, on rank 0 (None elsewhere)."""
        import torch
        sizes = [t[0] for t in self._all_gather_ints([nbytes])]
        cap = max(max(sizes), 1)
        mine = torch.zeros(cap, dtype=torch.uint8, device=self.device)
        if nbytes:
            src = scan[:nbytes] if isinstance(scan, torch.Tensor) else torch.frombuffer(bytearray(bytes(scan[:nbytes])), dtype=torch.uint8)
            mine[:nbytes] = src.to(self.device)
        out = torch.empty(self.world * cap, dtype=torch.uint8, device=self.device)
        self.dist.all_gather_into_tensor(out, mine, group=self.group)
        if self.rank != 0:
            return None
        out = out.cpu().view(self.world, cap)
        return b"".join(out[r, : sizes[r]].numpy().tobytes() for r in range(self.world))
