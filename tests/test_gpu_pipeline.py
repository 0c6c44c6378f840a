"""eid_group as a stage pipeline (csrc/pipeline.cu): N processes — here all on cuda:0, CUDA IPC works between processes of one device —
render a frame sequence as direct | indirect | post ranks; everything a rank owns must equal the single-GPU frame (eid_renderer_run)
bit for bit: G-buffer, motion, reservoirs, both composed images, and the ray counts."""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest

import eidola_b200 as eid
from eidola_b200 import abi, scenes

import common
import pipeline_worker as pw

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def run_ranks(world, cfg, timeout=420):
    gid = eid.Group.random_id().hex()
    with tempfile.TemporaryDirectory() as tmp:
        procs = []
        for rank in range(world):
            out = os.path.join(tmp, "rank%d.npz" % rank)
            env = dict(os.environ, EID_PIPE_TIMEOUT="240")
            procs.append((out, subprocess.Popen([sys.executable, os.path.join(HERE, "pipeline_worker.py"), str(rank), str(world), gid, out, json.dumps(cfg)],
                                                stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env)))
        results, fail = [], []
        for rank, (out, p) in enumerate(procs):
            try:
                log, _ = p.communicate(timeout=timeout)
            except subprocess.TimeoutExpired:
                for _, q in procs:
                    q.kill()
                log, _ = p.communicate()
                fail.append("rank %d timed out:\n%s" % (rank, log[-2000:]))
                continue
            if p.returncode != 0:
                fail.append("rank %d exited with %d:\n%s" % (rank, p.returncode, log[-2000:]))
                continue
            with np.load(out) as z:
                results.append({k: z[k] for k in z.files})
        assert not fail, "\n".join(fail)
        return results


def single_gpu_frames(cfg):
    """The same frame sequence through eid_renderer_run on one renderer: per frame, the byte image of every buffer a rank may own."""
    w, h = cfg["size"]
    arrays = getattr(scenes, cfg.get("scene", "small_room"))()
    psc = eid.Scene(0)
    psc.load_arrays(arrays)
    acc = eid.AccelStructure()
    acc.create(psc)
    rr = eid.Renderer()
    rr.create((w, h), psc, acc)
    rr.set_env_constant(common.ENV)
    rr.set_strict_math(True)
    info = psc.info()
    psc.update_camera(w, h)
    owned, _ = pw.owned_tables()
    frames = []
    for f in range(int(cfg["frames"])):
        if cfg.get("orbit"):
            pw.orbit_camera(psc, arrays.camera, f)
        psc.update_camera(w, h)
        st = common.frame_state(w, h, info, f, ReSTIRState=int(cfg.get("restir", abi.eTemporal)), **cfg.get("state", {}))
        rr.run(st, f)
        rr.sync()
        frames.append({name: rr.read(which).view(np.uint8).reshape(-1).copy() for name, (which, _, _) in owned.items()})
    s = rr.stats()
    return frames, (s.totalClosestHitRays, s.totalAnyHitRays)


CASES = [
    # world, stages (0,0,0 = default split), extra
    pytest.param(2, (0, 0, 0), dict(frames=3), id="n2_1-0-1_static_lockstep"),
    pytest.param(3, (0, 0, 0), dict(frames=6, lockstep=False), id="n3_1-1-1_static_free_running"),
    pytest.param(4, (0, 0, 0), dict(frames=4, orbit=True, history=2), id="n4_2-1-1_orbit_auto_history"),
    pytest.param(5, (0, 0, 0), dict(frames=6, orbit=True, history=1, lockstep=False), id="n5_2-2-1_orbit_free_running"),
    pytest.param(4, (1, 1, 2), dict(frames=4, orbit=True, history=2, host=True), id="n4_1-1-2_sharded_post_host_delivery"),
    pytest.param(8, (0, 0, 0), dict(frames=6, lockstep=False, host=True), id="n8_3-3-2_static_free_running_host_delivery"),
    pytest.param(3, (1, 1, 1), dict(frames=3, restir=abi.eSpatiotemporal, orbit=True), id="n3_spatiotemporal"),
]


@pytest.mark.parametrize("world,stages,extra", CASES)
def test_stage_pipeline_equals_single_gpu(world, stages, extra):
    cfg = dict(scene="small_room", size=[256, 144], stages=list(stages), history=2)
    cfg.update(extra)
    w, h = cfg["size"]
    want, want_rays = single_gpu_frames(cfg)
    ranks = run_ranks(world, cfg)
    owned, stage_of = pw.owned_tables()
    closest = sum(int(r["meta"][3]) for r in ranks)
    anyhit = sum(int(r["meta"][4]) for r in ranks)
    assert (closest, anyhit) == want_rays, "the ranks together traced %s rays, one GPU %s" % ((closest, anyhit), want_rays)
    covered = {name: 0 for name in owned}
    checked = 0
    for rank, r in enumerate(ranks):
        role, y0, y1 = int(r["meta"][0]), int(r["meta"][1]), int(r["meta"][2])
        for key, got in r.items():
            if key == "meta" or key.startswith("rows_"):
                continue
            name, f = key.rsplit("_", 1)
            which, row_bytes, half = owned[name]
            a, b = (int(v) for v in r["rows_" + name])
            rb = row_bytes(w)
            ref = want[int(f)][name][a * rb:b * rb]
            assert got.tobytes() == ref.tobytes(), "rank %d (stages %d, rows %d..%d): %s of frame %s differs from the single-GPU frame in %d bytes" % (
                rank, role, a, b, name, f, int((got != ref).sum()))
            checked += 1
            if int(f) == int(cfg["frames"]) - 1:
                covered[name] += b - a
        if not (role & abi.STAGE_POST):
            assert int(r["meta"][5]) > 0 and int(r["meta"][6]) > 0     # every producer wrote its rows into its consumers' buffers (peer copies)
    # the ranks of a stage cover the whole frame between them
    for name, rows in covered.items():
        assert rows == (h // 2 if owned[name][2] else h), "%s: %d rows covered" % (name, rows)
    assert checked >= 6


def test_pipeline_rejects_bad_arguments():
    arrays = scenes.cornell_scene()
    psc = eid.Scene(0)
    psc.load_arrays(arrays)
    acc = eid.AccelStructure()
    acc.create(psc)
    r = eid.Renderer()
    r.create((64, 64), psc, acc)
    gid = eid.Group.random_id()
    with pytest.raises(eid.EidolaError):
        eid.Group().create_pipeline(r, 0, 1, gid, 64)                 # a pipeline needs at least two ranks
    with pytest.raises(eid.EidolaError):
        eid.Group().create_pipeline(r, 0, 4, gid, 64, (1, 1, 1))      # stages do not add up
    with pytest.raises(eid.EidolaError):
        eid.Group().create_pipeline(r, 0, 3, gid, 256)                # allocation below the padded height
