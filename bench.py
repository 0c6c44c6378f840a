#!/usr/bin/env python
"""bench.py — Mray/s + fps of one Renderer::run frame at 1080p on the procedural 1 M-triangle / 1 k-emissive
scene (BASELINE.json configs[2]), on N B200s of one node.

  python bench.py [--gpus N] [--steps K] [--warmup W]            this repo's CUDA path
  python bench.py --impl reference [...]                          the reference algorithm on the host CPU (the oracle;
                                                                  the Vulkan reference cannot run here)

A step = one frame = the reference's 12 dispatches (direct_stage, indirect_stage, denoise_direct x4,
denoise_indirect x5, compose).  `value` = rays actually issued (ClosestHit + AnyHit, counted on the device) by all
ranks / wall time of the K timed frames, inputs resident in HBM.  `e2e` = the same through eid_renderer_render_host
with HOST buffers: camera + RtxState uploaded and both result images downloaded to pinned memory every frame.
N > 1 (torchrun): rank r traces row band r (direct + indirect stage), ONE exchange step all-gathers the pre-denoise
buffers over NCCL, every rank then denoises + composes the full frame (SURVEY.md §8e alternative).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W, H = 1920, 1080
MAX_DEPTH = 3
RESTIR_STATE = 3     # eTemporal: the reference's default (sample_example.hpp:165) and the headline configuration; --restir changes it
WORKLOAD = "C3: procedural closed room, 707x707-quad noise height-field floor (999,698 tris) + 10 wall tris + 1,000 emissive tris, " \
           "1920x1080, ReSTIR DI+GI temporal, M=4, maxDepth 3, MIS, denoise on (K1..K5), static camera"


WORKLOADS = {   # BASELINE.json configs[2..4]; the default / headline is c3
    "c3": dict(size=(1920, 1080), quads=707, light_quads=500, light_seed=566, text=WORKLOAD),
    "c4": dict(size=(3840, 2160), quads=707, light_quads=500, light_seed=566,
               text=WORKLOAD.replace("C3:", "C4:").replace("1920x1080", "3840x2160")),
    "c5": dict(size=(1920, 1080), quads=2236, light_quads=5000, light_seed=567,
               text="C5: same generator at 2236x2236 quads (9,999,402 floor tris) + 10,000 emissive tris, 1920x1080, full pipeline"),
}
_ACTIVE = "c3"


def scene_arrays(quick=False):
    from eidola_b200 import scenes
    if quick:
        return scenes.heightfield_room(quads=96, n_light_quads=50)
    c = WORKLOADS[_ACTIVE]
    return scenes.heightfield_room(quads=c["quads"], n_light_quads=c["light_quads"], light_seed=c["light_seed"])


def frame_state(info, frame, w=W, h=H):
    from eidola_b200 import abi
    return abi.default_rtx_state(
        w, h, environmentProb=0.0, time=1000 + 16 * frame, maxDepth=MAX_DEPTH, ReSTIRState=RESTIR_STATE,
        fireflyClampThreshold=float(np.float32(4 * np.pi)), envMapLuminIntegInv=float(np.float32(1 / np.pi)),
        lightLuminIntegInv=float(np.float32(1.0) / (np.float32(info.trigLightWeight) + np.float32(info.puncLightWeight))))


ENV = (0.25, 0.25, 0.25)

# Algorithmic screen-space bytes per pixel of each stage (N = W*H, indirect stages per N/4)
SCREEN_BYTES_PER_PX = {"direct_stage": 124.0, "indirect_stage": 51.0, "denoise_direct": 192.0, "denoise_indirect": 60.0, "compose": 68.0}
NODE_BYTES, TRI_BYTES, HIT_GATHER_BYTES = 64, 48, 108 + 80


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons while the timed region runs (B200_PROFILING.md recipe).  The timed region of the headline workload
    is ~60 ms, shorter than nvidia-smi's start-up: the sampler is started BEFORE the warm-up frames, polls every 20 ms, and every row carries the
    host time at which it arrived; finish(t0, t1) keeps the rows of the timed region [t0, t1] (and says so), else the rows taken under the same
    load just around it (warm-up frames before, end-to-end frames after)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.rows, self.stop_flag, self.index = [], False, index
        self.proc = None

    def run(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, bufsize=1)
            for line in self.proc.stdout:
                self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))
                if self.stop_flag:
                    break
        except Exception:
            pass

    def wait_first(self, timeout=3.0):
        t0 = time.perf_counter()
        while not self.rows and time.perf_counter() - t0 < timeout and self.is_alive():
            time.sleep(0.01)

    def finish(self, t0=None, t1=None):
        self.stop_flag = True
        if self.proc is not None:
            try:
                self.proc.terminate()
            except Exception:
                pass
        self.join(timeout=2)

        def parse(rows):
            sm, mx, reasons = [], 0, set()
            for _, r in rows:
                try:
                    sm.append(float(r[0]))
                    mx = max(mx, float(r[1]))
                    for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                        if v.lower().startswith("active"):
                            reasons.add(name)
                except Exception:
                    continue
            return sm, mx, reasons
        window = "timed region"
        rows = [r for r in self.rows if t0 is not None and t0 <= r[0] <= t1 + 0.02]
        sm, mx, reasons = parse(rows)
        if not sm and t0 is not None:      # none landed inside: the rows taken under the same load around it (warm-up before, e2e frames after)
            window = "under load around the timed region (warm-up / end-to-end frames)"
            sm, mx, reasons = parse([r for r in self.rows if t0 - 1.0 <= r[0] <= t1 + 1.0])
        if not sm:
            window = "whole run"
            sm, mx, reasons = parse(self.rows)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm), "window": window}


# ncu DRAM / L2 bytes per frame and stage come from a committed capture of THIS workload (tools/ncu_traffic.py writes the file from
# `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,gpu__time_duration.sum` of one bench frame) together
# with a hash of the kernel sources it was taken from; the line says whether that hash still matches the sources being run.
def kernel_source_hash():
    import hashlib
    h = hashlib.sha256()
    d = os.path.join(ROOT, "cis-565-final-vr-raytracer_b200", "csrc")
    for fn in sorted(os.listdir(d)):
        if fn.endswith((".cu", ".cuh", ".h")):
            h.update(fn.encode())
            h.update(open(os.path.join(d, fn), "rb").read())
    return h.hexdigest()[:16]


def load_traffic_profile(workload):
    p = os.path.join(ROOT, "profiles", "r02_traffic_%s.json" % workload)
    if not os.path.exists(p):
        return None
    try:
        d = json.load(open(p))
        d["file"] = os.path.relpath(p, ROOT)
        d["matches_sources"] = d.get("kernel_source_hash") == kernel_source_hash()
        return d
    except Exception:
        return None


# what the counters say bounds each stage (profiles/README.md, round 2 captures)
STAGE_LIMITER = {"direct_stage": "instruction issue + L1/L2 latency of the BVH walk (DRAM ~5 % of peak): not HBM-bound",
                 "indirect_stage": "latency of dependent BVH node fetches in small ray queues (long-scoreboard stalls): not HBM-bound",
                 "denoise_direct": "fp32 + MUFU issue (25 taps x 3 exponentials per pixel and pass), shared-memory tiles: not HBM-bound",
                 "denoise_indirect": "fp32 + MUFU issue, as denoise_direct at quarter resolution",
                 "compose": "HBM streaming"}


def measured_peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle on the host cores
# ------------------------------------------------------------------------------------------------------------------
def oracle_run(arrays, w, h, frames, warm):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as ol
    osc = ol.OracleScene()
    t0 = time.time()
    osc.load_arrays(arrays)
    load_s = time.time() - t0
    orr = ol.OracleRenderer(osc, (w, h))
    orr.set_env_constant(ENV)
    osc.update_camera(w, h)
    info = osc.info()
    rays, secs, per_kernel = 0, 0.0, np.zeros(5)
    for f in range(warm + frames):
        osc.update_camera(w, h)
        st = frame_state(info, f, w, h)
        t0 = time.time()
        orr.run(st, f)
        dt = time.time() - t0
        if f >= warm:
            s = orr.stats()
            rays += s.closestHitRays + s.anyHitRays
            secs += dt
            per_kernel += np.array(s.kernelMs[:])
    return dict(mrays=rays / secs / 1e6, ms_per_frame=1e3 * secs / frames, cores=ol.lib().orc_num_threads(), load_s=load_s,
                kernel_ms=(per_kernel / frames).tolist(), rays_per_frame=rays / frames)


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # torchrun exports OMP_NUM_THREADS=1 to its ranks; the reference arm always uses every host core it is allowed to run on
    nthreads = host_threads()
    os.environ["OMP_NUM_THREADS"] = str(nthreads)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as ol
    ol.lib().orc_set_num_threads(nthreads)
    fw, fh = WORKLOADS[_ACTIVE]["size"]
    sw, sh = (fw, fh) if not args.quick else (W // 8, H // 8)
    arrays = scene_arrays(args.quick)
    r = oracle_run(arrays, sw, sh, args.steps, args.warmup)
    sample = "%d full frames of the workload at %dx%d (the stated configuration), %d host threads" % (args.steps, sw, sh, r["cores"])
    line = {
        "impl": "reference", "metric": "Mray/s (ClosestHit+AnyHit rays per second, full Renderer::run frame)", "value": r["mrays"], "unit": "Mray/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_frame"], "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOADS[_ACTIVE]["text"], "width": sw, "height": sh,
                   "reference_arm": "CPU oracle (C++/OpenMP restatement of the reference shaders; the Vulkan app cannot run here)", "sample": sample},
        "cpu_baseline": {"value": r["mrays"], "unit": "Mray/s", "cores": r["cores"], "kind": "port", "sample": sample},
        "e2e": {"value": r["mrays"], "unit": "Mray/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "fps": 1e3 / r["ms_per_frame"], "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------------
# CUDA arm
# ------------------------------------------------------------------------------------------------------------------
class DevBuf:
    """Minimal __cuda_array_interface__ wrapper so torch can alias library-owned device memory (NCCL plumbing)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3, "strides": None}


def run_cuda(args):
    import torch
    import eidola_b200 as eid
    from eidola_b200 import abi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        # NCCL prints its version banner on stdout at communicator creation; keep stdout clean for the ONE JSON line by
        # pointing fd 1 at stderr until the first collective has run
        sys.stdout.flush()
        _saved_stdout = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        _t = torch.zeros(1, device=torch.device("cuda", local))
        dist.all_reduce(_t)
        torch.cuda.synchronize()
        sys.stdout.flush()
        os.dup2(_saved_stdout, 1)
        os.close(_saved_stdout)
    if not torch.cuda.is_available() or eid.lib().eid_device_count() < 1:
        raise SystemExit("bench.py: no CUDA device - this framework has no CPU fallback (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # a dedicated (non-default) torch stream: its handle is non-null, so the library enqueues on it and the
    # torch.cuda.Event pair below brackets exactly the kernels of the timed frames
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0

    w, h = WORKLOADS[_ACTIVE]["size"] if not args.quick else (W // 4, H // 4 // 16 * 16)
    arrays = scene_arrays(args.quick)
    scene = eid.Scene(local)
    scene.load_arrays(arrays)
    accel = eid.AccelStructure()
    accel.create(scene)
    ainfo = accel.info()
    info = scene.info()

    # band partition (eid_group_layout): rows per rank rounded up to 8, allocation padded to N equal bands;
    # stage pipeline (eid_group_pipeline_layout): ranks per stage, bands per stage, allocation padded for the widest split
    pipe = world > 1 and args.mgpu == "pipeline"
    stages = tuple(int(x) for x in args.stages.split(",")) if args.stages else (0, 0, 0)
    lay = eid.Group.pipeline_layout(h, world, rank, stages) if pipe else None
    alloc_h = lay.paddedHeight if pipe else eid.Group.layout(h, world)[2]
    rr = eid.Renderer()
    rr.create((w, alloc_h), scene, accel, stream=stream.cuda_stream)
    rr.set_env_constant(ENV)
    rr.set_overlap(not args.no_overlap)
    rr.set_pipeline(args.frames_in_flight)
    rr.set_denoise_rows(args.denoise_rows)
    rr.set_denoise_tiles({"tma": 1, "cpasync": 2, "legacy": 0}[args.denoiser], args.tile_rows)
    rr.set_wavefront({"wavefront": 1, "wavefront-serial": 2, "mega": 0}[args.k2], args.trace_blocks)
    grp = None
    shm = None
    if world > 1:
        # the library owns the NCCL communicator; the host only carries rank 0's 128-byte id to the other ranks
        idt = torch.zeros(128, dtype=torch.uint8, device=dev)
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(eid.Group.random_id() if pipe else eid.Group.unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        grp = eid.Group()
        if pipe:
            # stage pipeline: direct | indirect | post ranks joined by peer copies over NVLink (CUDA IPC mappings), no collective
            grp.create_pipeline(rr, rank, world, bytes(idt.cpu().numpy().tobytes()), h, stages)
        else:
            grp.create(rr, rank, world, bytes(idt.cpu().numpy().tobytes()))
        grp.set_mode(post_sharded=(args.post == "sharded"), history={"never": 0, "always": 1, "auto": 2}[args.history], gather_final=True)
        # ONE host image pair set shared by all ranks (POSIX shared memory, page-locked in every process): each rank delivers its own band
        shm_path = "/dev/shm/eidola_bench_%s" % os.environ.get("MASTER_PORT", "0")
        nbytes = 4 * w * h * 16
        if rank == 0:
            with open(shm_path, "wb") as f:
                f.truncate(nbytes)
        dist.barrier()
        shm = np.memmap(shm_path, dtype=np.uint8, mode="r+", shape=(nbytes,))
        rc = torch.cuda.cudart().cudaHostRegister(shm.ctypes.data, nbytes, 0)
        assert int(rc) == 0, "cudaHostRegister failed: %s" % rc

    cam0 = arrays.camera
    scene.update_camera(w, h)

    def step(frame, e2e_bufs=None):
        if args.orbit:
            a = np.deg2rad(args.orbit * frame)
            e = np.array(cam0["eye"], np.float64)
            scene.set_lookat((e[0] * np.cos(a) - e[2] * np.sin(a), e[1], e[0] * np.sin(a) + e[2] * np.cos(a)), cam0["center"], cam0["up"], np.rad2deg(cam0["yfov"]))
        scene.update_camera(w, h)
        st = frame_state(info, frame, w, h)
        if world == 1:
            if e2e_bufs is None:
                rr.run(st, frame)
            else:
                # public host-buffer API, pipelined: camera + RtxState go up, both result images come down to pinned memory on a
                # copy stream while the next frame renders (two buffer pairs alternate); timed() waits for the last copy
                pair = e2e_bufs[frame & 1]
                rr.render_host_async(scene.get_camera(), st, frame, pair[0], pair[1])
        elif e2e_bufs is None:
            grp.run(st, frame)                          # whole multi-GPU frame inside the library (exchanges A, B, C)
        else:
            pair = e2e_bufs[frame & 1]                  # every rank delivers ITS band over its own PCIe link; no exchange C
            grp.render_host_async(scene.get_camera(), st, frame, pair[0], pair[1])

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    host_ms = []

    def timed(nsteps, first_frame, e2e_bufs=None, profiling=0):
        rr.set_profiling(profiling)
        s0 = rr.stats()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        kms = np.zeros(5)
        e0.record(stream)
        th0 = time.perf_counter()
        for k in range(nsteps):
            step(first_frame + k, e2e_bufs)
            if profiling:
                kms += np.array(rr.stats().kernelMs[:])     # syncs; only used in the separate per-kernel pass
        host_ms.append(1e3 * (time.perf_counter() - th0) / nsteps)   # host time to ENQUEUE one frame (no synchronisation inside the loop)
        if e2e_bufs is not None and world == 1:
            rr.wait_host()                           # the last frame's device->host copies are inside the timed region
        if e2e_bufs is not None and world > 1:
            grp.wait_host()                          # ... on every rank in the multi-GPU path
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        s1 = rr.stats()
        rays = (s1.totalClosestHitRays + s1.totalAnyHitRays) - (s0.totalClosestHitRays + s0.totalAnyHitRays)
        if dist is not None:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
            rt = torch.tensor([rays], device=dev, dtype=torch.int64)
            dist.all_reduce(rt, op=dist.ReduceOp.SUM)
            rays = int(rt.item())
        return ms, rays, kms / max(1, nsteps)

    frame = 0
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
        sampler.wait_first()                         # nvidia-smi is up and streaming before any frame runs
    for _ in range(max(3, args.warmup)):
        step(frame)
        frame += 1
    t_clk0 = time.perf_counter()
    ms, rays, _ = timed(args.steps, frame)
    t_clk1 = time.perf_counter()
    frame += args.steps
    value = rays / (ms * 1e-3) / 1e6

    # end to end: host buffers, H2D of the per-frame inputs + D2H of both result images inside the timed region
    if world == 1:
        pinned_t = [[torch.empty((h, w, 4), dtype=torch.float32).pin_memory() for _ in range(2)] for _ in range(2)]
        pinned = [[t.data_ptr() for t in pair] for pair in pinned_t]
    else:
        img = w * h * 16
        pinned = [[shm.ctypes.data + (2 * k + j) * img for j in range(2)] for k in range(2)]
    for _ in range(2):
        step(frame, pinned)
        frame += 1
    e2e_steps = max(3, min(args.steps, 16))
    ems, erays, _ = timed(e2e_steps, frame, pinned)
    frame += e2e_steps
    e2e_value = erays / (ems * 1e-3) / 1e6
    clocks = sampler.finish(t_clk0, t_clk1) if sampler else None     # rows of the device-timed region; stopped here, after the end-to-end frames

    # per-kernel pass (CUDA events inside the library around every stage, same frames/state sequence) + visit counters
    ksteps = max(3, min(args.steps, 16))
    rr.set_overlap(False)                       # isolated per-stage times: strict K1..K5 order on one stream
    _, _, kms = timed(ksteps, frame, None, profiling=1)
    rr.set_overlap(not args.no_overlap)
    frame += ksteps
    rank_kms = None
    if dist is not None:                        # every rank's stage times (the frame is as slow as the slowest rank)
        tk = torch.tensor(kms, device=dev, dtype=torch.float64)
        allk = [torch.zeros_like(tk) for _ in range(world)]
        dist.all_gather(allk, tk)
        rank_kms = [[round(float(v), 4) for v in t.cpu().numpy()] for t in allk]
    # latency of ONE frame in isolation (enqueue to the last rank's last kernel): equals the frame time for row bands, the sum of the
    # stages plus two hand-overs for the stage pipeline
    lat = []
    for _ in range(3):
        lms, _, _ = timed(1, frame)
        lat.append(lms)
        frame += 1
    rr.set_profiling(2)
    step(frame)
    vs = rr.stats()
    frame += 1
    rr.set_profiling(0)
    # checksum of the last composed frame (rows of the rendered size): equal for every N when the sharded frame is bit-identical
    import zlib
    crc = 0
    pinfo = None
    if pipe:
        # the composed frame is spread over the post ranks: each writes its band into the shared host images, rank 0 checksums the whole
        gi = grp.info()
        grp.sync()
        imgs = shm[:2 * w * h * 16].view(np.float32).reshape(2, h, w, 4)
        if gi.stages & abi.STAGE_POST:
            y0, y1 = gi.y0, min(gi.y1, h)
            for k, which in enumerate((abi.BUF_DIRECT, abi.BUF_INDIRECT)):
                imgs[k, y0:y1] = rr.read(which).reshape(alloc_h, w, 4)[y0:y1]
        tp = torch.tensor([gi.peerCopies, gi.peerBytes] + [int(v) for v in vs.kernelLaunches[:]], device=dev, dtype=torch.int64)
        dist.all_reduce(tp, op=dist.ReduceOp.SUM)
        dist.barrier()
        if rank == 0:
            crc = zlib.crc32(np.ascontiguousarray(imgs).tobytes(), 0)
            pinfo = {"stages": [gi.nDirect, gi.nIndirect, gi.nPost], "peer_copies_per_frame": int(tp[0].item()) / frame,
                     "peer_MB_per_frame": int(tp[1].item()) / frame / 1e6, "stream_mem_ops": bool(gi.streamMemOps),
                     "launches_per_frame_all_ranks": [int(v) for v in tp[2:].cpu().numpy()]}
    elif rank == 0:
        for which in (abi.BUF_DIRECT, abi.BUF_INDIRECT):
            img = rr.read(which).reshape(alloc_h, w, 4)[:h]
            crc = zlib.crc32(np.ascontiguousarray(img).tobytes(), crc)

    line = None
    if rank == 0:
        n_px = w * h
        peak, peak_src = measured_peak_hbm()
        names = abi.KERNEL_NAMES
        launches = [max(1, int(v)) for v in (pinfo["launches_per_frame_all_ranks"] if pipe else vs.kernelLaunches[:])]   # 1, 10, 1 prep + 4 passes, 5 passes, 1
        screen = {k: SCREEN_BYTES_PER_PX[k] * n_px for k in names}
        band_frac = 1.0 / world                                       # this rank's share of the trace stages ...
        post_frac = band_frac if args.post == "sharded" else 1.0      # ... and of denoise + compose (mode B: per band; mode A: replicated)
        fracs = [band_frac, band_frac, post_frac, post_frac, post_frac]
        if pipe:
            # every stage runs on its own ranks: report the slowest rank of each stage and that rank's share of the rows
            kms = np.max(np.array(rank_kms), axis=0)
            nD, nI, nP = pinfo["stages"]
            fracs = [1.0 / nD, 1.0 / max(nI, 1), 1.0 / nP, 1.0 / nP, 1.0 / nP]
        lpr = launches                                                # launches of ONE rank per frame and stage
        if pipe:
            lpr = [max(1, launches[i] // n) for i, n in enumerate((nD, max(nI, 1), nP, nP, nP))]
        dom = int(np.argmax(kms))
        tot_rays = max(1, vs.closestHitRays + vs.anyHitRays)
        # BVH fetches counted by the STATS kernels: served by L1 / L2 (the scene is cache-resident at 1 M triangles), reported apart
        trace_bytes = vs.nodeVisits * NODE_BYTES + vs.triangleTests * TRI_BYTES + (vs.closestHitRays * HIT_GATHER_BYTES)
        k1_rays = min(vs.closestHitRays, int(n_px * fracs[0])) + vs.primaryHits
        cache_served = {names[0]: trace_bytes * k1_rays / tot_rays, names[1]: trace_bytes * (tot_rays - k1_rays) / tot_rays}
        prof = load_traffic_profile(_ACTIVE) if (world == 1 and not args.quick) else None
        kernels = {}
        for i, k in enumerate(names):
            algo = screen[k] * fracs[i]                                # compulsory screen-space bytes of SURVEY 8(d), this rank's rows
            t = kms[i] * 1e-3
            e = {"ms_per_frame": float(kms[i]), "launches": lpr[i], "algorithmic_MB_per_frame": algo / 1e6,
                 "algorithmic_MB_per_launch": algo / lpr[i] / 1e6,
                 "achieved_GBps": (algo / t / 1e9) if t > 0 else None, "frac_of_hbm_peak": (algo / t / 1e9 / peak) if t > 0 else None,
                 "limiter": STAGE_LIMITER[k]}
            if k in cache_served:
                e["cache_served_bvh_MB_per_frame"] = cache_served[k] / 1e6
                e["cache_served_bvh_GBps"] = (cache_served[k] / t / 1e9) if t > 0 else None
            if prof and k in prof.get("stages", {}):
                q = prof["stages"][k]
                e["ncu_dram_MB_per_frame"] = q["dram_bytes"] / 1e6
                e["ncu_dram_frac_of_hbm_peak"] = (q["dram_bytes"] / t / 1e9 / peak) if t > 0 else None
                e["ncu_l2_GBps"] = (q["lts_bytes"] / t / 1e9) if t > 0 else None
                e["ncu_traffic_over_algorithmic"] = q["dram_bytes"] / algo if algo > 0 else None
            kernels[k] = e
        kd = kernels[names[dom]]
        dom_launch_ms = kms[dom] / lpr[dom]
        achieved = kd["algorithmic_MB_per_launch"] * 1e6 / (dom_launch_ms * 1e-3) / 1e9 if dom_launch_ms > 0 else 0.0
        traffic = None
        if prof and names[dom] in prof.get("stages", {}):
            traffic = prof["stages"][names[dom]]["dram_bytes"] / lpr[dom]
        line = {
            "metric": "Mray/s (ClosestHit+AnyHit rays per second, full Renderer::run frame)", "value": value, "unit": "Mray/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOADS[_ACTIVE]["text"] if not args.quick else "QUICK smoke variant (not a benchmark number)", "width": w, "height": h,
                       "triangles": int(ainfo.triangleCount), "emissive_triangles": int(info.trigLightCount), "maxDepth": MAX_DEPTH,
                       "ReSTIRState": ["none", "ris", "spatial", "temporal", "spatiotemporal"][RESTIR_STATE],
                       "parallelism": ("eid_group stage pipeline: %d direct | %d indirect | %d post ranks (row bands per stage, one rank per GPU), frames handed over "
                                       "by peer copies over NVLink into the consumer's buffers (CUDA IPC mappings) + stream-ordered flags, no collective; "
                                       "reservoir history: %s" % (pinfo["stages"][0], pinfo["stages"][1], pinfo["stages"][2], args.history)) if pipe else
                                      ("eid_group: %d row bands of %d rows (multiples of 8), one rank per GPU; exchange A (G-buffer + direct image, behind "
                                       "indirect_stage) and B (quarter-res indirect image): one NCCL launch each; %s; reservoir history: %s" % (
                                           world, grp.info().bandRows, "denoise+compose per band, exchange C: all-gather of the composed images (1 NCCL launch)"
                                           if args.post == "sharded" else "denoise+compose replicated on every rank", args.history)) if world > 1 else "single GPU",
                       "camera": ("orbit %.2f deg/frame" % args.orbit) if args.orbit else "static",
                       "l2": "inputs larger than L2: each frame streams ~1.0 GB of screen-space buffers + ~0.14 GB of BVH/triangles/vertices (L2 = 126 MB)",
                       "bvh": {"nodes": int(ainfo.nodeCount), "node_MB": ainfo.nodeBytes / 1e6, "tri_MB": ainfo.triBytes / 1e6,
                               "height": int(ainfo.maxDepth), "build_ms": float(ainfo.buildMs),
                               "build": "binned SAH on the host threads + refit / 4-wide collapse on the GPU (EID_ACCEL_FAST_TRACE)" if ainfo.fastTrace
                               else "Morton LBVH on the GPU (EID_ACCEL_FAST_BUILD)"}},
            "fps": 1e3 / (ms / args.steps),
            "rays_per_frame": rays / args.steps,
            "e2e": {"value": e2e_value, "unit": "Mray/s", "h2d_bytes_per_step": C.sizeof(abi.SceneCamera) + C.sizeof(abi.RtxState),
                    "d2h_bytes_per_step": 2 * w * h * 16, "pipelined": "D2H of frame f overlaps the kernels of frame f+1 (copy stream, 2 pinned buffer pairs)" if world == 1 else ("every post rank copies ITS band of the composed images into one shared pinned host image pair over its own PCIe link (eid_group_render_host_async)" if pipe else "every rank copies ITS band of the composed images into one shared pinned host image pair over its own PCIe link (eid_group_render_host_async); no exchange C"),
                    "ms_per_step": ems / e2e_steps, "fps": 1e3 / (ems / e2e_steps), "steps": e2e_steps},
            "gpu_launches": int(sum(launches)) * args.steps, "launches_per_frame": launches,
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": names[dom], "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic,
                         "traffic_source": (("%s (ncu, kernel sources %s)" % (prof["file"], "unchanged since the capture" if prof["matches_sources"] else "CHANGED since the capture")) if prof else None),
                         "peak_source": peak_src, "limiter": STAGE_LIMITER[names[dom]],
                         "dram_frac": kd.get("ncu_dram_frac_of_hbm_peak"), "l2_GBps": kd.get("ncu_l2_GBps"),
                         "cache_served_bvh_GBps": kd.get("cache_served_bvh_GBps"),
                         "note": "achieved = compulsory screen-space bytes of SURVEY 8(d) per launch (each needed element read once, each output "
                                 "written once) / CUDA-event time of the dominant kernel, against the measured HBM copy peak; the kernel is NOT "
                                 "HBM-bound (see limiter): its BVH fetches are served by L1/L2 and are reported apart as cache_served_bvh_GBps"},
            "kernels": kernels,
            "kernels_note": "per-stage times are measured in a separate pass with the stages serialised on one stream; the timed frames "
                            + ("run K3 concurrently with K2+K4 on a second stream, so their sum exceeds ms_per_step" if not args.no_overlap else "are serialised too"),
            "exchange1_ms": float(vs.exchangeMs) if world > 1 else None,
            "image_crc32": "%08x" % crc, "frames_rendered": frame,
            "host_enqueue_ms_per_frame": {"device_timed": host_ms[0], "e2e": host_ms[1]},
            "stage_ms_per_rank": rank_kms,
            "nccl_launches_per_frame": 0 if (pipe or world == 1) else (3 if args.post == "sharded" else 2),
            "pipeline": pinfo,
            "frame_latency_ms": float(min(lat)),
            "visits_per_ray": {"nodes": vs.nodeVisits / tot_rays, "triangles": vs.triangleTests / tot_rays},
        }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # reference algorithm on the host cores, bounded sample of the same workload (reported, not a target)
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_lib as ol
        ol.lib().orc_set_num_threads(host_threads())
        r = oracle_run(arrays, w, h, 3, 1)
        line["cpu_baseline"] = {"value": r["mrays"], "unit": "Mray/s", "cores": r["cores"], "kind": "port",
                                "sample": "3 full frames (after 1 warm-up) of the same workload at %dx%d, %d host threads; oracle BVH build %.1fs not included" % (w, h, r["cores"], r["load_s"]),
                                "ms_per_frame_at_sample": r["ms_per_frame"]}
    elif rank == 0:
        line["cpu_baseline"] = None
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=32)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--impl", default="eidola", choices=["eidola", "reference"])
    ap.add_argument("--quick", action="store_true", help="tiny scene/resolution (plumbing check, not a benchmark)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--denoise-rows", type=int, default=2, choices=(1, 2, 4), help="A-Trous pixels per thread (rows of one column sharing tap rows)")
    ap.add_argument("--denoiser", choices=["tma", "cpasync", "legacy"], default="tma",
                    help="A-Trous passes: shared-memory tile kernel fed by TMA (default) / by cp.async, or the round-1 L1-served kernel")
    ap.add_argument("--tile-rows", type=int, default=2, choices=(2, 4), help="tile kernel: lattice rows per thread")
    ap.add_argument("--k2", choices=["wavefront", "wavefront-serial", "mega"], default="wavefront",
                    help="form of indirect_stage: ray queues + persistent dynamic-fetch traversal (default) or one thread per pixel")
    ap.add_argument("--trace-blocks", type=int, default=0, help="grid of the persistent traversal kernel in 128-thread blocks (0 = library default)")
    ap.add_argument("--no-overlap", action="store_true", help="strict K1..K5 order on one stream (default: K3 runs beside K2/K4 on a second stream)")
    ap.add_argument("--frames-in-flight", type=int, default=1, choices=(1, 2),
                    help="eid_renderer_set_pipeline: 2 lets direct_stage of frame f+1 overlap indirect_stage/denoise/compose of frame f (measured +2 %%: the stages compete for the register file, profiles/README.md)")
    ap.add_argument("--orbit", type=float, default=0.0, help="degrees per frame the camera orbits the scene centre (SURVEY 8(d): 0.5); default static")
    ap.add_argument("--history", default="auto", choices=["never", "always", "auto"],
                    help="N>1: how last frame's reservoirs cross band edges (eid_group_set_mode): auto = gathered when the camera moved")
    ap.add_argument("--post", default="sharded", choices=["sharded", "replicated"],
                    help="N>1 only. sharded (mode B): each rank denoises/composes its band, 2 exchange steps; replicated (mode A): "
                         "one exchange step, every rank post-processes the full frame")
    ap.add_argument("--mgpu", default="pipeline", choices=["pipeline", "bands"],
                    help="N>1: pipeline = direct | indirect | post ranks, one frame behind each other (throughput set by the slowest stage; default); "
                         "bands = every rank renders one row band of the same frame (lowest latency)")
    ap.add_argument("--stages", default="", help="N>1, --mgpu pipeline: ranks per stage as d,i,p (default: the library's split, e.g. 3,3,2 at N=8)")
    ap.add_argument("--restir", default="temporal", choices=["none", "ris", "spatial", "temporal", "spatiotemporal"],
                    help="RtxState.ReSTIRState (default temporal = the reference's default and the headline configuration)")
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS), help="c3 = headline (BASELINE.json metric); c4 = 4K; c5 = 10 M triangles")
    args = ap.parse_args()
    global _ACTIVE, RESTIR_STATE
    _ACTIVE = args.workload
    RESTIR_STATE = ["none", "ris", "spatial", "temporal", "spatiotemporal"].index(args.restir)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()
