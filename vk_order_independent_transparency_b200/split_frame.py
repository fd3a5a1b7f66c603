"""Sort-first split-frame rendering across the GPUs of one box.

Every rank owns the interleaved row strips `k % world == rank` of the output image (strips of `stripRows` rows), renders
only the triangles that touch them, and the resolved BGRA8 strips are exchanged with ONE all-gather over NCCL/NVLink.
The reference has no multi-GPU mode; pixels are independent in every technique, so the split is exact whenever the
technique is deterministic (for the linked list each band has its own node pool).
"""
import numpy as np

DEFAULT_STRIP_ROWS = 32


def band_rows(height, band_count, band_index, strip_rows=DEFAULT_STRIP_ROWS):
    """Global output rows owned by a band, in local order (mirrors tileRowOwner / oit_local_row_to_global)."""
    rows = np.arange(height)
    return rows[(rows // strip_rows) % band_count == band_index]


def max_band_rows(height, band_count, strip_rows=DEFAULT_STRIP_ROWS):
    return max(len(band_rows(height, band_count, b, strip_rows)) for b in range(band_count))


def gather_permutation(height, band_count, strip_rows=DEFAULT_STRIP_ROWS):
    """perm[y] = row index inside the gathered [band_count, pad, width] buffer that holds output row y."""
    pad = max_band_rows(height, band_count, strip_rows)
    perm = np.empty(height, np.int64)
    for b in range(band_count):
        rows = band_rows(height, band_count, b, strip_rows)
        perm[rows] = b * pad + np.arange(len(rows))
    return perm, pad


def assemble(bands, height, width, strip_rows=DEFAULT_STRIP_ROWS):
    """bands[b] = uint32 array [rows_b(+padding), width] -> full [height, width] frame (numpy or torch)."""
    n = len(bands)
    first = bands[0]
    if isinstance(first, np.ndarray):
        out = np.empty((height, width), first.dtype)
        for b in range(n):
            rows = band_rows(height, n, b, strip_rows)
            out[rows] = bands[b][:len(rows)]
        return out
    import torch
    out = torch.empty((height, width), dtype=first.dtype, device=first.device)
    for b in range(n):
        rows = torch.as_tensor(band_rows(height, n, b, strip_rows), device=first.device)
        out[rows] = bands[b][:len(rows)]
    return out


class BandGather:
    """The band gather with everything precomputed: one in-place all-gather of the padded strips + one row gather.

    `slot` is this rank's slice of the gather buffer; point the renderer's output at it (or copy the strips into it) and
    call `gather()`: it returns the full [height, width] frame on every rank.  All work is enqueued on the current torch
    stream, so run it under `torch.cuda.stream(<the renderer's stream>)` to keep render and gather ordered."""

    def __init__(self, height, width, rank, world, device, strip_rows=DEFAULT_STRIP_ROWS, dtype=None, group=None):
        import torch
        self.height, self.width, self.rank, self.world, self.group = height, width, rank, world, group
        perm, self.pad = gather_permutation(height, world, strip_rows)
        self.perm = torch.as_tensor(perm, device=device)
        self.local_rows = len(band_rows(height, world, rank, strip_rows))
        self.buf = torch.zeros((world, self.pad, width), dtype=dtype or torch.int32, device=device)
        self.slot = self.buf[rank]
        self.frame = torch.empty((height, width), dtype=self.buf.dtype, device=device)

    def gather(self, local=None):
        import torch
        import torch.distributed as dist
        if local is not None:
            self.slot[:local.shape[0]].copy_(local)
        if self.world > 1:
            # in place: the input is this rank's slice of the output
            dist.all_gather_into_tensor(self.buf.view(-1), self.slot.reshape(-1), group=self.group)
        torch.index_select(self.buf.view(self.world * self.pad, self.width), 0, self.perm, out=self.frame)
        return self.frame


def gather_frame(local_rows, height, width, rank, world, strip_rows=DEFAULT_STRIP_ROWS, group=None, gather_buf=None):
    """One-shot convenience wrapper around BandGather (plans are rebuilt every call: use BandGather in a frame loop)."""
    g = BandGather(height, width, rank, world, local_rows.device, strip_rows, local_rows.dtype, group)
    return g.gather(local_rows)
