"""CPU, world_size 2 over gloo: the band partition and the band gather of the split-frame mode."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vk_order_independent_transparency_b200 import split_frame as SF  # noqa: E402


@pytest.mark.parametrize("height,bands,strip", [(720, 1, 32), (720, 2, 32), (2160, 8, 32), (513, 4, 16), (40, 8, 32), (1080, 3, 64)])
def test_bands_partition_the_frame(height, bands, strip):
    rows = [SF.band_rows(height, bands, b, strip) for b in range(bands)]
    allr = np.sort(np.concatenate(rows))
    assert np.array_equal(allr, np.arange(height))          # every row exactly once
    for r in rows:
        assert np.all(np.diff(r) > 0)
    frame = np.random.default_rng(0).integers(0, 2**32, (height, 37), dtype=np.uint32)
    back = SF.assemble([frame[r] for r in rows], height, 37, strip)
    assert np.array_equal(back, frame)


def _worker(rank, world, port, height, width, strip, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    full = torch.from_numpy(np.random.default_rng(7).integers(0, 2**31 - 1, (height, width), dtype=np.int64).astype(np.int32))
    mine = full[torch.as_tensor(SF.band_rows(height, world, rank, strip))]
    got = SF.gather_frame(mine, height, width, rank, world, strip)
    q.put((rank, bool(torch.equal(got, full))))
    dist.destroy_process_group()


@pytest.mark.parametrize("height", [96, 100])
def test_band_gather_world2_gloo(height):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, height, 24, 32, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
    assert res == [(0, True), (1, True)]
