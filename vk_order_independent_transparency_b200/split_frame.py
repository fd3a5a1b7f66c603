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


def gather_frame(local_rows, height, width, rank, world, strip_rows=DEFAULT_STRIP_ROWS, group=None, gather_buf=None):
    """The band gather: all-gather of the (padded) strips of every rank, then the interleave.

    local_rows: torch tensor [localRows, width] (int32/uint32 BGRA8 words) on this rank's device (cuda -> NCCL, cpu -> gloo).
    Returns the full [height, width] frame on every rank."""
    import torch
    import torch.distributed as dist

    pad = max_band_rows(height, world, strip_rows)
    if gather_buf is None:
        gather_buf = torch.empty((world, pad, width), dtype=local_rows.dtype, device=local_rows.device)
    mine = gather_buf[rank]
    mine[:local_rows.shape[0]].copy_(local_rows)
    if world > 1:
        dist.all_gather_into_tensor(gather_buf.view(-1), mine.reshape(-1).clone(), group=group)
    return assemble([gather_buf[b] for b in range(world)], height, width, strip_rows)
