"""CPU (gloo, world_size 2) test of the N>1 host logic: band partition, equal-size padded chunks and the in-place
all-gather stitch give every rank the same frame a single process renders.  The trace/post stages are played by the
CPU oracle here (no GPU in this container); the GPU twin is test_gpu_parity.test_band_sharded_trace_equals_full_frame."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
W, H = 96, 72            # 72 rows / 2 ranks = 36 -> bands of 40 rows (a band edge in the middle of a quarter-res tile row), padded allocation 80 rows


def _worker(rank, world, port, out, restir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import eidola_b200 as eid
    from eidola_b200 import abi, scenes, sharding
    import common
    import oracle_lib as ol
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ol.lib().orc_set_num_threads(2)
    alloc_h = sharding.padded_height(H, world)
    y0, y1 = sharding.band_range(rank, world, H)
    osc = ol.OracleScene()
    osc.load_arrays(scenes.cornell_scene())
    orr = ol.OracleRenderer(osc, (W, alloc_h))
    orr.set_env_constant(common.ENV)
    osc.update_camera(W, H)
    info = osc.info()
    for f in range(2):
        osc.update_camera(W, H)
        st = common.frame_state(W, H, info, f, maxDepth=2, ReSTIRState=restir)
        orr.run_trace(st, f, y0, min(y1, H))     # spatial reuse: the band's direct stage also primes one halo row on each side
        # the exchange step: pre-denoise buffers, full-res rows for G-buffer / direct, half-res rows for the indirect temp
        for which in (abi.BUF_THIS_GBUFFER, abi.BUF_DIRECT, abi.BUF_DENOISE_IND_A):
            buf = orr.read(which).view(np.uint8).copy()
            t = torch.from_numpy(buf)
            if which == abi.BUF_DENOISE_IND_A:      # half-res rows with the full-res pitch: rows [y0/2, y1/2)
                rows = t.view(alloc_h, W * 16)[:alloc_h // 2].reshape(-1)
                sharding.all_gather_bands(dist, rows, rank, world, inplace=False)
            else:
                sharding.all_gather_bands(dist, t, rank, world, inplace=False)
            orr.write(which, buf)
        orr.run_post(st, f)
    snap = {k: orr.read(w).view(np.uint8).copy() for k, w in (("direct", abi.BUF_DIRECT), ("indirect", abi.BUF_INDIRECT),
                                                             ("gbuffer", abi.BUF_THIS_GBUFFER))}
    np.savez(os.path.join(out, "rank%d.npz" % rank), **snap)
    dist.barrier()
    dist.destroy_process_group()


def test_band_partition_properties():
    from eidola_b200 import sharding
    for h in (1080, 2160, 72, 17, 1):
        for world in (1, 2, 4, 8):
            b = sharding.band_rows(h, world)
            assert b % 8 == 0 and b * world >= h
            rows = [sharding.band_range(r, world, h) for r in range(world)]
            assert rows[0][0] == 0 and all(a[1] == c[0] for a, c in zip(rows, rows[1:]))
            assert rows[-1][1] == sharding.padded_height(h, world) or world == 1
            assert all((y0 % 8) == 0 for y0, _ in rows)
    assert sharding.band_rows(1080, 8) == 136 and sharding.padded_height(1080, 8) == 1088


@pytest.mark.parametrize("restir", [3, 4])      # eTemporal (default), eSpatiotemporal (neighbour reads cross the band edge)
def test_two_rank_gloo_band_exchange_matches_single_process(tmp_path, restir):
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path), restir), nprocs=2, join=True)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from eidola_b200 import abi, scenes, sharding
    import common
    import oracle_lib as ol
    alloc_h = sharding.padded_height(H, 2)
    osc = ol.OracleScene()
    osc.load_arrays(scenes.cornell_scene())
    orr = ol.OracleRenderer(osc, (W, alloc_h))
    orr.set_env_constant(common.ENV)
    osc.update_camera(W, H)
    info = osc.info()
    for f in range(2):
        osc.update_camera(W, H)
        orr.run(common.frame_state(W, H, info, f, maxDepth=2, ReSTIRState=restir), f)
    want = {k: orr.read(w).view(np.uint8) for k, w in (("direct", abi.BUF_DIRECT), ("indirect", abi.BUF_INDIRECT), ("gbuffer", abi.BUF_THIS_GBUFFER))}
    for r in range(2):
        got = np.load(os.path.join(str(tmp_path), "rank%d.npz" % r))
        for k in want:
            assert got[k].tobytes() == want[k].tobytes(), "rank %d: %s differs from the single-process frame" % (r, k)
