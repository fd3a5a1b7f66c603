"""Multi-GPU check of the in-library band gather (run under torchrun, one rank per GPU):
every rank renders its strips, the frame graph gathers them, and the result must equal a single-band render."""
import faulthandler
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vk_order_independent_transparency_b200 as oit  # noqa: E402

faulthandler.dump_traceback_later(60, exit=True)
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
W, H = 1280, 720
ok = True
for alg, aa in ((1, 4), (3, 0), (6, 1)):
    st = oit.State(algorithm=alg, aaType=aa)
    verts, idx, ipo = oit.generate_scene(st)
    ubo = oit.default_camera(W, H)
    s = oit.Sample(st, W, H, device=local, bandCount=world, bandIndex=rank)
    s.setScene(verts, idx, ipo)
    print(f"[rank {rank}] enabling gather", flush=True)
    s.enableBandGather(dist)
    for i in range(4):
        s.onRender(ubo)
        print(f"[rank {rank}] frame {i} done", flush=True)
    got = s.readFrame()
    full = oit.Sample(st, W, H, device=local)
    full.setScene(verts, idx, ipo)
    full.onRender(ubo)
    want = full.readColor()
    same = bool(np.array_equal(got, want))
    print(f"[rank {rank}] alg {alg} aa {aa}: gathered frame == single-band frame: {same}", flush=True)
    ok &= same
    s.close()
    full.close()
t = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print("BAND GATHER", "OK" if t.item() else "FAILED", flush=True)
dist.destroy_process_group()
sys.exit(0 if t.item() else 1)
