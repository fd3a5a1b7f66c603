"""Multi-GPU check of the split-frame exchange inside the library (run under torchrun, one rank per GPU): every rank
renders its strips, the frame is exchanged (peer-memory stores from the frame kernel, or the NCCL all-gather), and the
result on every rank must equal a single-band render.  usage: torchrun ... tools/check_band_gather.py [peers|nccl|both]"""
import faulthandler
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vk_order_independent_transparency_b200 as oit  # noqa: E402

faulthandler.dump_traceback_later(240, exit=True)
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
which = sys.argv[1] if len(sys.argv) > 1 else "both"
modes = ["peers", "nccl"] if which == "both" else [which]
W, H = 1280, 720
ok = True
CASES = [dict(algorithm=1, aaType=4), dict(algorithm=3, aaType=0), dict(algorithm=6, aaType=1), dict(algorithm=4, aaType=3),
         dict(algorithm=1, aaType=0, percentTransparent=0), dict(algorithm=5, aaType=2, percentTransparent=50)]
for mode in modes:
    for kw in CASES:
        st = oit.State(**kw)
        verts, idx, ipo = oit.generate_scene(st)
        ubo = oit.default_camera(W, H)
        s = oit.Sample(st, W, H, device=local, bandCount=world, bandIndex=rank)
        s.setScene(verts, idx, ipo)
        if mode == "peers":
            if not s.enableBandPeers(dist):
                print(f"[rank {rank}] peer memory unavailable: {s.L.oit_last_error(s.h)}", flush=True)
                ok = False
                s.close()
                break
        else:
            s.enableBandGather(dist)
        for i in range(4):
            s.onRender(ubo)
        got = s.readFrame()
        full = oit.Sample(st, W, H, device=local)
        full.setScene(verts, idx, ipo)
        full.onRender(ubo)
        want = full.readColor()
        same = bool(np.array_equal(got, want))
        print(f"[rank {rank}] {mode} {kw}: exchanged frame == single-band frame: {same}", flush=True)
        ok &= same
        s.close()
        full.close()
    # frame time of the headline workload with this exchange (host clock over 20 frames, max over ranks)
    st = oit.State(algorithm=1, aaType=4)
    verts, idx, ipo = oit.generate_scene(st)
    s = oit.Sample(st, 3840, 2160, device=local, bandCount=world, bandIndex=rank)
    s.setScene(verts, idx, ipo)
    ubo = oit.default_camera(3840, 2160)
    if (s.enableBandPeers(dist) if mode == "peers" else (s.enableBandGather(dist), True)[1]):
        for _ in range(5):
            s.onRender(ubo)
        s.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        for _ in range(20):
            s.onRender(ubo)
        s.synchronize()   # oit_render only enqueues
        ms = torch.tensor([(time.perf_counter() - t0) / 20 * 1e3], device="cuda")
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(f"{mode}: 4K 8xMSAA linked list, {world} bands: {ms.item():.3f} ms / frame (host clock, incl. sync)", flush=True)
    s.close()
t = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print("BAND EXCHANGE", "OK" if t.item() else "FAILED", flush=True)
dist.destroy_process_group()
sys.exit(0 if t.item() else 1)
