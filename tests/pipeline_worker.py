"""One rank of an eid_group stage pipeline (csrc/pipeline.cu), run as its own process by tests/test_gpu_pipeline.py and tools/.

    python tests/pipeline_worker.py <rank> <world> <id hex> <out.npz> <json config>

config: {"scene": "small_room", "size": [w, h], "frames": n, "stages": [d, i, p], "history": 0|1|2, "orbit": bool, "device": k,
         "restir": int, "host": bool}
Every rank renders the same frame sequence through eid_group_run (or eid_group_render_host_async with "host") and writes what it
owns after every frame: direct ranks the G-buffer / motion / direct reservoirs of their band, indirect ranks the indirect reservoirs,
post ranks their band of the two composed images.  All ranks may share ONE GPU (CUDA IPC works between processes on a device)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def orbit_camera(psc, cam, f):
    ang = np.deg2rad(1.5 * f)
    e = np.array(cam["eye"], np.float64)
    psc.set_lookat((e[0] * np.cos(ang) - e[2] * np.sin(ang), e[1] + 0.15 * f, e[0] * np.sin(ang) + e[2] * np.cos(ang)), cam["center"], cam["up"],
                   np.rad2deg(cam["yfov"]))


def owned_tables():
    from eidola_b200 import abi
    owned = {   # name -> (buffer, bytes per row of a width-w allocation, quarter-res rows?)
        "gbuffer": (abi.BUF_THIS_GBUFFER, lambda w: w * 16, False), "motion": (abi.BUF_MOTION, lambda w: w * 4, False),
        "direct_resv": (abi.BUF_THIS_DIRECT_RESV, lambda w: w * 36, False),
        "indirect_resv": (abi.BUF_THIS_INDIRECT_RESV, lambda w: (w // 2) * 76, True),
        "direct": (abi.BUF_DIRECT, lambda w: w * 16, False), "indirect": (abi.BUF_INDIRECT, lambda w: w * 16, False)}
    stage_of = {"gbuffer": abi.STAGE_DIRECT, "motion": abi.STAGE_DIRECT, "direct_resv": abi.STAGE_DIRECT, "indirect_resv": abi.STAGE_INDIRECT,
                "direct": abi.STAGE_POST, "indirect": abi.STAGE_POST}
    return owned, stage_of


def main():
    rank, world, id_hex, out_path, cfg = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], sys.argv[4], json.loads(sys.argv[5])
    import eidola_b200 as eid
    from eidola_b200 import abi, scenes
    import common
    global OWNED, STAGE_OF
    OWNED, STAGE_OF = owned_tables()
    dev = int(cfg.get("device", 0))
    w, h = cfg["size"]
    arrays = getattr(scenes, cfg.get("scene", "small_room"))()
    psc = eid.Scene(dev)
    psc.load_arrays(arrays)
    acc = eid.AccelStructure()
    acc.create(psc)
    stages = tuple(cfg.get("stages", (0, 0, 0)))
    lay = eid.Group.pipeline_layout(h, world, rank, stages)
    rr = eid.Renderer()
    rr.create((w, lay.paddedHeight), psc, acc)
    rr.set_env_constant(common.ENV)
    rr.set_strict_math(bool(cfg.get("strict", True)))
    grp = eid.Group()
    grp.create_pipeline(rr, rank, world, bytes.fromhex(id_hex), h, stages)
    grp.set_mode(history=int(cfg.get("history", 2)))
    info = psc.info()
    psc.update_camera(w, h)
    out = {}
    host = None
    if cfg.get("host"):
        host = [np.zeros((h, w, 4), np.float32) for _ in range(4)]        # two pairs; page-locking is not required for correctness
    y0, y1 = lay.y0, min(lay.y1, h)
    dband = ((h + world - 1) // world + 7) // 8 * 8
    for f in range(int(cfg["frames"])):
        if cfg.get("orbit"):
            orbit_camera(psc, arrays.camera, f)
        psc.update_camera(w, h)
        st = common.frame_state(w, h, info, f, ReSTIRState=int(cfg.get("restir", abi.eTemporal)), **cfg.get("state", {}))
        if host is None:
            grp.run(st, f)
        else:
            pair = host[2 * (f & 1):2 * (f & 1) + 2]
            grp.render_host_async(psc.get_camera(), st, f, pair[0].ctypes.data, pair[1].ctypes.data)
        if cfg.get("lockstep", True) or f == int(cfg["frames"]) - 1:
            if host is not None:
                grp.wait_host()
            grp.sync()
            for name, (which, row_bytes, half) in OWNED.items():
                if host is not None and name in ("direct", "indirect"):
                    # host delivery: EVERY rank copies the rows of its delivery band (rank k of n: rows k B .. (k + 1) B, B = ceil8(ceil(h / n))) to the host
                    a, b = min(rank * dband, h), min((rank + 1) * dband, h)
                    out["%s_%d" % (name, f)] = pair[0 if name == "direct" else 1][a:b].view(np.uint8).reshape(-1).copy()
                    out["rows_%s" % name] = np.array([a, b])
                    continue
                if not (lay.stages & STAGE_OF[name]):
                    continue
                a, b = (y0 // 2, y1 // 2) if half else (y0, y1)
                rb = row_bytes(w)
                out["%s_%d" % (name, f)] = rr.read(which).view(np.uint8).reshape(-1)[a * rb:b * rb].copy()
                out["rows_%s" % name] = np.array([a, b])
    s = rr.stats()
    gi = grp.info()
    out["meta"] = np.array([lay.stages, y0, y1, s.totalClosestHitRays, s.totalAnyHitRays, gi.peerCopies, gi.peerBytes, gi.streamMemOps], np.int64)
    np.savez(out_path, **out)
    grp.destroy()


if __name__ == "__main__":
    main()
