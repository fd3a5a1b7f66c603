#!/usr/bin/env python
"""bench.py -- transparent fragments/s and ms/frame of the OIT hot path on N B200s (see BASELINE.json / DESIGN.md).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload headline|config1..config5] [--impl reference]

A step = one frame of the hot path (vertex stage + binning, clears, colour pass(es), composite, resolve [, band gather]).
`value` is measured with the scene resident in HBM; `e2e` is the same frame through the C ABI with HOST buffers
(the scene's sphere table + UBO uploaded from pinned memory and the resolved frame read back, every step; the step that
uploads the flattened mesh instead is reported as `e2e.flattened_mesh`).  N>1 is sort-first split frame:
every rank renders its interleaved row strips and the frame kernel stores the resolved pixels into every rank's frame
buffer over NVLink peer memory (fallback / --nccl-gather: ONE ncclAllGather at the end of the frame graph).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (W, H, State kwargs, description)
    "headline": (3840, 2160, dict(algorithm=1, aaType=4), "default sphere scene (1024 spheres, subdiv 16), Linked List, OIT_LAYERS=8, N=10, 8x MSAA coverage masks, 3840x2160"),
    "config1": (1280, 720, dict(algorithm=1), "default scene, Linked List, no AA, 1280x720"),
    "config2": (1920, 1080, dict(algorithm=3), "default scene, Loop64 OIT_LAYERS=8, no AA, 1920x1080"),
    "config3": (1920, 1080, dict(algorithm=4, aaType=2), "default scene, Spinlock, 4x MSAA per-sample, 1920x1080"),
    "config4": (3840, 2160, dict(algorithm=6, aaType=4), "default scene, WBOIT, 8x MSAA, 3840x2160"),
    "config5": (3840, 2160, dict(algorithm=1, numObjects=100000, linkedListAllocatedPerElement=128), "100k spheres, Linked List N=128, no AA, 3840x2160, split frame"),
}
TECH_TABLE = [(a, 0) for a in range(7)]  # per-technique table at 3840x2160, no AA, default parameters


KERNEL_SOURCES = ["oit_raster_ll.cu", "oit_raster_q.cu", "oit_raster.cu", "oit_raster_common.cuh", "oit_fragment.cuh", "oit_fused.cuh", "oit_device.cuh",
                  "oit_internal.h", "oit_clip.cuh"]


def kernel_source_hash():
    """sha256 (first 16 hex digits) over the sources of the frame kernel: ties an ncu capture to the code it was taken from."""
    import hashlib
    h = hashlib.sha256()
    for name in KERNEL_SOURCES:
        with open(os.path.join(ROOT, "vk_order_independent_transparency_b200", "csrc", name), "rb") as f:
            h.update(f.read())
    return h.hexdigest()[:16]


def measured_traffic(workload):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed ncu capture
    (profiles/traffic.json, written by tools/make_traffic.py from the .ncu-rep).  A capture of OTHER kernel sources than the
    ones in this tree is refused: the number would silently be stale."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            e = json.load(f)[workload]
        if e.get("kernel_source_sha16") != kernel_source_hash():
            return None, f"profiles/traffic.json[{workload}] was captured from kernel sources {e.get('kernel_source_sha16')}, this tree is {kernel_source_hash()}: refused"
        return int(e["dram_bytes_per_launch"]), f"ncu --set full capture {e.get('capture')}: dram__bytes_read.sum + dram__bytes_write.sum, 1 launch"
    except Exception as ex:
        return None, f"no capture for this workload ({type(ex).__name__})"


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20", "-i", str(self.index)],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if not self.proc:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        if not sm:
            # the timed region was shorter than the polling period: one direct query right after it (the GPU is still warm)
            try:
                f = [x.strip() for x in subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                                       capture_output=True, text=True, timeout=10).stdout.strip().split(",")]
                sm, mx = [float(f[1])], [float(f[2])]
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}
        return out


def algorithmic_bytes(oit, st, stats, W, H, nVerts, nIndices):
    """SURVEY 8(d) per-unit figures x the units of one frame, per stage (DESIGN.md 'Roofline accounting')."""
    S = st.msaa
    ss = st.supersample
    P = W * ss * H * ss
    Pa = P * (S if st.sampleShading else 1)
    Ps = P * S
    L = st.oitLayers
    cov = st.coverageShading()
    E = 16 if cov else 8
    F, Fst, Ftb = stats["fragments"], stats["fragmentsStored"], stats["fragmentsTail"]
    cbar = S / 2 if (cov or (st.algorithm == oit.OIT_WEIGHTED and S > 1)) else 1  # mean covered samples of a tail / WBOIT fragment (estimate)
    geom = 40 * nVerts + 4 * nIndices
    a = st.algorithm
    if a == oit.OIT_LINKEDLIST:
        clear, color, comp = 4 * Pa, Fst * 24 + Ftb * 8 * cbar, 4 * Pa + 16 * Fst + 8 * Ps
    elif a == oit.OIT_LOOP64:
        clear, color, comp = 8 * L * Pa, F * 16 + Ftb * 8 * cbar, 8 * min(Fst, L * Pa) + 8 * Ps
    elif a == oit.OIT_LOOP:
        clear, color, comp = 4 * L * Pa, F * 8 + F * 8 + geom, 8 * Fst + 8 * Ps
    elif a == oit.OIT_SIMPLE:
        clear, color, comp = 4 * Pa, F * 8 + Fst * E, 4 * Pa + E * Fst + 8 * Ps
    elif a in (oit.OIT_SPINLOCK, oit.OIT_INTERLOCK):
        clear, color, comp = (12 if a == oit.OIT_SPINLOCK else 8) * Pa, F * 12 + Fst * E, 4 * Pa + E * min(Fst, L * Pa) + 8 * Ps
    else:
        clear, color, comp = 10 * Ps, F * cbar * 20, 10 * Ps + 8 * Ps
    clear += 4 * Ps                      # colour clear (render-pass clear of m_colorImage)
    color += geom
    resolve = 4 * Ps + 4 * W * H
    return {"clear": clear, "color": color, "composite": comp, "resolve": resolve}


def run_ours(args):
    import torch
    import vk_order_independent_transparency_b200 as oit
    from vk_order_independent_transparency_b200 import split_frame as SF

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        os.environ["NCCL_DEBUG"] = "WARN"  # keep NCCL's version banner off stdout: the contract is ONE JSON line
        dist.init_process_group("nccl", device_id=dev)

    W, H, kw, desc = WORKLOADS[args.workload]
    st = oit.State(**kw)
    verts, idx, ipo = oit.generate_scene(st)
    # pinned host copies (e2e leg) and the device-resident scene (value leg)
    hverts = torch.from_numpy(verts).pin_memory()
    hidx = torch.from_numpy(idx.view(np.int32)).pin_memory()
    dverts, didx = hverts.to(dev), hidx.to(dev)
    ubo = oit.default_camera(W, H)
    s = oit.Sample(st, W, H, device=local, bandCount=world, bandIndex=rank, stripRows=args.strip_rows)
    s.setSceneDevice(dverts.data_ptr(), verts.shape[0], didx.data_ptr(), idx.size, ipo, keepalive=(dverts, didx))
    stream = torch.cuda.ExternalStream(s.L.oit_stream(s.h), device=dev)
    fin_dev = torch.as_tensor(s.device_array(oit.BUF_FINAL, "<i4"), device=dev).view(-1)[: s.localRows * W].view(s.localRows, W)
    band_gather, exchange = None, None
    if world > 1:
        if args.torch_gather:
            band_gather = SF.BandGather(H, W, rank, world, dev, args.strip_rows, torch.int32)   # gather driven from Python
        elif not args.nccl_gather and s.enableBandPeers(dist):
            exchange = "peer memory: the frame kernel stores every resolved pixel into all bands' frame buffers over NVLink (CUDA IPC), 2 flag rounds per frame inside the frame graph"
        else:
            s.enableBandGather(dist)   # the band gather as ONE ncclAllGather + interleave at the end of the frame graph
            exchange = "in-library ncclAllGather inside the frame graph"
            fin_dev =torch.as_tensor(s.device_array(oit.BUF_FINAL, "<i4"), device=dev).view(-1)[: s.localRows * W].view(s.localRows, W)
    hfinal = torch.empty((s.localRows, W), dtype=torch.int32).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def gather():
        # the band gather: ONE NCCL all-gather of the resolved strips + the row interleave, on the renderer's stream
        with torch.cuda.stream(stream):
            band_gather.gather(fin_dev)

    def step_resident():
        s.onRender(ubo)                      # render (+ in-library band exchange): enqueued on the library's stream; up to 4 frames in flight
        if band_gather is not None:
            gather()

    def step_e2e():
        s.setScene(hverts.numpy(), hidx.numpy().view(np.uint32), ipo)   # H2D of the step's inputs from pinned memory
        s.onRender(ubo)
        if band_gather is not None:
            gather()
        s.readColor(hfinal.numpy().view(np.uint32))                      # D2H of the step's result

    def timed(step, steps, warmup, sample_clocks=False):
        # (the clock sampler polls every 20 ms: it starts before the warm-up so that a short timed region -- 40 frames of
        # 0.35 ms at 8 GPUs -- still falls between samples taken under load; samples of the warm-up are load samples too)
        sampler = ClockSampler(local) if sample_clocks else None
        if sampler:
            sampler.start()
        for _ in range(warmup):
            step()
        s.synchronize()   # oit_render is asynchronous: completes the warm-up frames (and their one-off buffer growth)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        stage = {"geometry": 0.0, "clear": 0.0, "color": 0.0, "composite": 0.0, "resolve": 0.0, "opaque": 0.0, "exchange_wait": 0.0}
        launches = 0
        e0.record(stream)
        t0 = time.perf_counter()
        for _ in range(steps):
            step()
        e1.record(stream)
        barrier()
        wall_ms = (time.perf_counter() - t0) * 1e3
        ms = max(e0.elapsed_time(e1), 0.0)
        clocks = sampler.stop() if sampler else None
        # per-stage breakdown (the library's own CUDA events) from a few extra, untimed frames
        for _ in range(min(steps, 5)):
            step()
            sst = s.stats()
            for k, n in (("geometry", "msGeometry"), ("clear", "msClear"), ("color", "msColor"), ("composite", "msComposite"), ("resolve", "msResolve"),
                         ("opaque", "msOpaque"), ("exchange_wait", "msExchangeWait")):
                stage[k] += sst[n] / min(steps, 5) * steps
        launches = sst["kernelLaunches"] * steps
        t = torch.tensor([ms, wall_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t[0].item(), t[1].item(), {k: v / steps for k, v in stage.items()}, launches, clocks, sst

    ms_total, wall_total, stage_ms, launches, clocks, last = timed(step_resident, args.steps, args.warmup, sample_clocks=True)
    gather_ok, gather_note = None, None
    if world > 1 and band_gather is None:
        # untimed check of the in-library band exchange: the frame EVERY rank holds after the exchange must equal the frame
        # a single band (one GPU rendering all rows) produces -- rendered here, on this rank's GPU, for the comparison
        frame = torch.as_tensor(s.frameDevice(), device=dev).clone()
        full = oit.Sample(st, W, H, device=local)
        full.setSceneDevice(dverts.data_ptr(), verts.shape[0], didx.data_ptr(), idx.size, ipo, keepalive=(dverts, didx))
        full.onRender(ubo)
        full.synchronize()
        want = torch.as_tensor(full.device_array(oit.BUF_FINAL, "<i4"), device=dev).view(-1)[: H * W].view(H, W)
        same = bool(torch.equal(frame.view(H, W), want))
        full_tail = full.stats()["fragmentsTail"]
        full.close()
        okt = torch.tensor([1 if same else 0], device=dev)
        dist.all_reduce(okt, op=dist.ReduceOp.MIN)
        gather_ok = bool(okt.item())
        if full_tail > 0 and st.algorithm == oit.OIT_LINKEDLIST:
            # the linked-list pool overflows: WHICH fragments are tail-blended depends on the allocation order, and every band
            # has its own pool (SURVEY 8(e) "semantic caveat"), so the frame is not defined bit for bit -- not a verdict on
            # the exchange (tests/test_gpu_fullsize.py::test_config5_full_scene_overflow_regime states what IS defined)
            gather_ok = None
            gather_note = "not applicable: the linked-list pool overflows, so the frame is not defined bit for bit (per-band pools)"

    # split frame: colour-pass time, wait time and fragment count of EVERY band (rank order)
    per_band = None
    if world > 1:
        mine = torch.tensor([stage_ms["color"], stage_ms["exchange_wait"], stage_ms["geometry"], float(last["fragments"])], dtype=torch.float64, device=dev)
        allb = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allb, mine)
        per_band = {"color_ms": [round(t[0].item(), 4) for t in allb], "exchange_wait_ms": [round(t[1].item(), 4) for t in allb],
                    "geometry_ms": [round(t[2].item(), 4) for t in allb], "fragments": [int(t[3].item()) for t in allb],
                    "note": "per-stage times come from frames rendered one at a time after the timed loop (library events); the wait is READY + DONE"}
    F_local = last["fragments"]
    Ft = torch.tensor([F_local, last["fragmentsStored"], last["fragmentsTail"]], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(Ft)
    F, Fst, Ftb = (int(x) for x in Ft.tolist())
    # render and band gather are enqueued on ONE stream (the library's), so the CUDA events bracket the whole step on
    # the device; the maximum over ranks is the frame time (the wall clock of the loop is kept for reference)
    frame_ms = ms_total / args.steps
    e_ms_total, e_wall_total, _, _, _, _ = timed(step_e2e, max(3, args.steps // 2), 3)
    e_steps = max(3, args.steps // 2)
    e2e_ms = max(e_ms_total, e_wall_total) / e_steps  # the D2H read-back is synchronous: wall clock covers it
    # The end-to-end step with the instanced scene input: the step's inputs are what the sample's scene IS -- the
    # 32-byte-per-sphere table (centre, radius, colour) -- and the UBO, from pinned host memory; the mesh is flattened on the
    # device inside the step.  This is the `e2e` of the line (it scales with the bands: 32 KB go up instead of 34.8 MB per
    # rank); the step that uploads the flattened 40-byte-per-vertex mesh like initScene is reported next to it.
    sph_table = oit.generate_spheres(st)
    hsph = torch.from_numpy(sph_table).pin_memory()

    def step_e2e_instanced():
        s.setSceneSpheres(hsph.numpy())
        s.onRender(ubo)
        if band_gather is not None:
            gather()
        s.readColor(hfinal.numpy().view(np.uint32))

    i_ms_total, i_wall_total, _, _, _, _ = timed(step_e2e_instanced, e_steps, 3)
    e2e_inst_ms = max(i_ms_total, i_wall_total) / e_steps
    s.setSceneDevice(dverts.data_ptr(), verts.shape[0], didx.data_ptr(), idx.size, ipo, keepalive=(dverts, didx))

    peak, peak_src = peaks()
    bytes_stage = algorithmic_bytes(oit, st, {"fragments": F_local, "fragmentsStored": last["fragmentsStored"], "fragmentsTail": last["fragmentsTail"]},
                                    W, s.localRows, verts.shape[0], idx.size)
    # oit_render's fast path (what the library decides in oit_render): one fused kernel whenever something transparent is drawn
    fused = os.environ.get("OIT_B200_NO_FUSE") is None and st.percentTransparent > 0 and st.numObjects > 0
    if fused:
        # oit_render's fused frame kernel: colour pass + composite + resolve of a tile in one launch; the colour samples
        # stay in shared memory, so the algorithmic bytes of the three reference stages are charged to this one kernel
        bytes_stage["color"] += bytes_stage["composite"] + bytes_stage["resolve"]
        bytes_stage["composite"] = bytes_stage["resolve"] = 0
    dom = max(("color", "composite", "clear", "resolve"), key=lambda k: stage_ms[k])
    kernel_name = {"color": "k_raster (fused colour pass + composite + resolve)" if fused else "k_raster (fragment insert)",
                   "composite": "k_composite", "clear": "k_fill32", "resolve": "k_resolve"}[dom]
    achieved = bytes_stage[dom] / (stage_ms[dom] * 1e-3) / 1e9 if stage_ms[dom] > 0 else 0.0
    frame_bytes = sum(bytes_stage.values())
    per_stage = {k: {"ms": round(stage_ms[k], 4), "alg_bytes": int(bytes_stage.get(k, 0)),
                     "gbs": round(bytes_stage.get(k, 0) / (stage_ms[k] * 1e-3) / 1e9, 1) if stage_ms[k] > 0 and k in bytes_stage else None}
                 for k in stage_ms}

    traffic, traffic_src = measured_traffic(args.workload) if (world == 1 and dom == "color") else (None, "N>1 or another dominant stage")
    out = None
    if rank == 0:
        out = {
            "metric": "transparent fragments/s", "value": F / (frame_ms * 1e-3), "unit": "fragments/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": frame_ms, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "u32/f32 (rgba8 + f32 depth A-buffer words, fp32 blend)",
            "data": "synthetic (the sample's seeded sphere cloud, generated on the host)",
            "config": {"workload": f"{args.workload}: {desc}", "fragments_per_frame": F, "fragments_stored": Fst, "fragments_tail": Ftb,
                       "width": W, "height": H, "parallelism": f"split-frame x{world}, {args.strip_rows}-row interleaved strips" if world > 1 else "single GPU",
                       "band_gather": None if world == 1 else ("torch.distributed all_gather" if band_gather is not None else exchange),
                       "band_gather_verified": gather_ok, "band_gather_note": gather_note, "bands": per_band,
                       "l2_policy": "working set (A-buffer + colour samples) is larger than L2; no explicit flush"},
            "ms_per_frame": frame_ms, "wall_ms_per_frame": wall_total / args.steps, "stages": per_stage, "gpu_launches": int(launches),
            "e2e": {"value": F / (e2e_inst_ms * 1e-3), "unit": "fragments/s", "ms_per_step": e2e_inst_ms,
                    "h2d_bytes_per_step": int(sph_table.nbytes + 224), "d2h_bytes_per_step": int(s.localRows * W * 4),
                    "note": "oit_set_scene_spheres + oit_render + oit_read_color every step: the scene's 32 B/sphere table and the UBO go up from pinned "
                            "host memory, the mesh is flattened on the device, this band's resolved strips are read back",
                    "flattened_mesh": {"value": F / (e2e_ms * 1e-3), "ms_per_step": e2e_ms, "h2d_bytes_per_step": int(verts.nbytes + idx.nbytes + 224),
                                       "note": "oit_set_scene instead: the flattened mesh (40 B/vertex + indices, what initScene uploads once) goes up every step"}},
            "roofline": {"bound": "hbm", "kernel": kernel_name, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "stage": dom, "alg_bytes_per_launch": int(bytes_stage[dom]),
                         "frame": {"alg_bytes": int(frame_bytes), "achieved": frame_bytes / (frame_ms * 1e-3) / 1e9,
                                   "frac": frame_bytes / (frame_ms * 1e-3) / 1e9 / peak}},
            "clocks": clocks,
        }
    # per-technique table @3840x2160 (the metric is quoted per technique), N=1 only
    if rank == 0 and world == 1 and not args.no_table:
        table = {}
        for alg, aa in TECH_TABLE:
            stt = oit.State(algorithm=alg, aaType=aa)
            t = oit.Sample(stt, 3840, 2160, device=local)
            t.setSceneDevice(dverts.data_ptr(), verts.shape[0], didx.data_ptr(), idx.size, ipo, keepalive=(dverts, didx))
            for _ in range(3):
                t.onRender(oit.default_camera(3840, 2160))
            ms = []
            for _ in range(5):
                t.onRender(oit.default_camera(3840, 2160))
                ms.append(t.stats()["msFrame"])
            ts = t.stats()
            table[oit.ALGORITHM_NAMES[alg]] = {"ms_per_frame": float(np.mean(ms)), "fragments": ts["fragments"],
                                               "fragments_per_s": ts["fragments"] / (np.mean(ms) * 1e-3)}
            t.close()
        out["per_technique_4k_noaa"] = table
    # CPU baseline: the oracle on this box's host cores, one frame of the same workload (rank 0, N=1 only)
    if rank == 0 and world == 1 and not args.no_cpu:
        s.onRender(ubo)
        got = s.readColor().copy()
        cb, want, Fo = cpu_baseline(args.workload, 1, 0, want_final=True)
        out["cpu_baseline"] = cb
        # the oracle frame the baseline just rendered is the parity check of the number-bearing configuration
        if want is not None:
            out["parity"] = {"checker": "oracle/liboit_oracle.so, same scene + UBO", "pixels": int(want.size), "pixels_differ": int((got != want).sum()),
                             "fragments_equal": bool(Fo == F), "fragments_oracle": int(Fo)}
        else:
            out["parity"] = {"checker": "not applicable: the CPU baseline of this workload is a bounded sample (first 20000 spheres); "
                                        "parity of config 5 is tests/test_gpu_fullsize.py (cfg5_*)"}
    s.close()
    if world > 1:
        dist.destroy_process_group()
    if rank == 0:
        emit(out)


def cpu_baseline(workload, steps, warmup, want_final=False):
    """The oracle (CPU restatement of the reference path; the reference itself needs Vulkan and cannot run here),
    band-parallel over all host cores."""
    from oracle import oracle_py as O
    W, H, kw, desc = WORKLOADS[workload]
    cores = os.cpu_count() or 1
    cfg = O.make_config(width=W, height=H, **kw)
    bounded = ""
    if kw.get("numObjects", 1024) > 20000:
        cfg.numObjects = 20000  # bounded sample of config 5: the first 20k spheres
        bounded = " (bounded sample: first 20000 spheres)"
    verts, idx, ipo = O.generate_scene(cfg)
    o = O.Oracle(cfg, threads=cores)
    o.set_scene(verts, idx, ipo)
    sd = O.camera(W, H)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        o.render(sd)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
        if sum(times) > 150:
            break
    F = o.stats["fragments"]
    final = o.final.copy() if (want_final and not bounded) else None
    o.close()
    t = float(np.mean(times))
    cb = {"value": F / t, "unit": "fragments/s", "cores": cores, "kind": "port", "ms_per_frame": t * 1e3, "frames_timed": len(times),
            "sample": f"{len(times)} full frame(s) of {workload}{bounded}, {F} fragments each, oracle/liboit_oracle.so with {cores} OpenMP threads"}
    return (cb, final, F) if want_final else cb


def probe_vulkan():
    """is there a Vulkan loader + a headless ICD on this box, i.e. could the reference's own path run here?
    Looked up every time (ICD manifests, ldconfig, the reference binary); the outcome goes into the reference arm's line."""
    import glob
    import shutil
    icds = []
    for d in ("/usr/share/vulkan/icd.d", "/etc/vulkan/icd.d", "/usr/local/share/vulkan/icd.d", os.path.expanduser("~/.local/share/vulkan/icd.d")):
        icds += sorted(glob.glob(os.path.join(d, "*.json")))
    for var in ("VK_ICD_FILENAMES", "VK_DRIVER_FILES"):
        if os.environ.get(var):
            icds += os.environ[var].split(":")
    loader = None
    try:
        out = subprocess.run(["ldconfig", "-p"], capture_output=True, text=True, timeout=10).stdout
        hits = [ln.split("=>")[-1].strip() for ln in out.splitlines() if "libvulkan.so" in ln]
        loader = hits[0] if hits else None
    except Exception:
        pass
    binary = shutil.which("vk_order_independent_transparency")
    return {"icd_manifests": icds, "loader": loader, "vulkaninfo": shutil.which("vulkaninfo"), "reference_binary": binary,
            "reference_runnable": bool(icds and loader and binary)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cb = cpu_baseline(args.workload, args.steps, min(args.warmup, 1))
    W, H, kw, desc = WORKLOADS[args.workload]
    vk = probe_vulkan()
    emit(({
        "impl": "reference", "metric": "transparent fragments/s", "value": cb["value"], "unit": "fragments/s", "n_gpus": args.gpus,
        "steps": cb["frames_timed"], "warmup": min(args.warmup, 1), "ms_per_step": cb["ms_per_frame"], "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "u32/f32", "data": "synthetic (the sample's seeded sphere cloud)",
        "config": {"workload": f"{args.workload}: {desc}", "vulkan_probe": vk,
                   "note": ("a Vulkan loader and ICD are present, but " if (vk["icd_manifests"] and vk["loader"]) else "no Vulkan loader + ICD on this box, and ")
                   + "the reference binary is not built (it needs nvpro_core2, shaderc, GLFW; no network): this arm is the reference's CPU restatement (the oracle) on the host cores"},
        "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": "fragments/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


_REAL_STDOUT = None


def emit(obj):
    """The ONE JSON line of the contract, on the real stdout."""
    sys.stdout.flush()
    if _REAL_STDOUT is not None:
        os.dup2(_REAL_STDOUT, 1)
    print(json.dumps(obj), flush=True)


if __name__ == "__main__":
    # libraries (NCCL's version banner, torchrun notices) write to the C-level stdout: route everything except the
    # final JSON line to stderr
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="headline", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--strip-rows", type=int, default=32)
    ap.add_argument("--nccl-gather", action="store_true", help="N>1: exchange the strips with the in-library ncclAllGather instead of peer-memory stores")
    ap.add_argument("--torch-gather", action="store_true", help="N>1: drive the band gather from Python (torch.distributed) instead of the library's frame graph")
    ap.add_argument("--no-table", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3) if a.impl == "ours" else a.warmup
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
