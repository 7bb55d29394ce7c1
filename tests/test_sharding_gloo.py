"""Multi-process (world_size 2, gloo, CPU) test of the N>1 path's host logic: the shards are disjoint, cover the problem,
keep every landmark on one rank, and their problem structure (ktk_get_structure, host-only) is the restriction of the
global structure."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, result_q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from kontiki_b200 import _lib, sharding, synthetic as syn
        cfg = syn.make_config("H1", scale=0.004)
        sh = sharding.shard_config(cfg, rank, world)
        # ---- cover / disjointness via all_gather of row masks
        for name, n in (("gyro", len(cfg["gyro"]["t"])), ("accel", len(cfg["accel"]["t"])), ("cam", len(cfg["cam"]["lm_idx"]))):
            mask = torch.zeros(n, dtype=torch.int32)
            rows = sh[name]["rows"] if name != "cam" else sh["cam"]["rows_sel"]
            mask[torch.from_numpy(rows)] = 1
            dist.all_reduce(mask)
            assert bool((mask == 1).all()), f"{name}: rows not covered exactly once"
        # ---- landmarks are rank-local
        lm_mask = torch.zeros(len(cfg["cam"]["rho"]), dtype=torch.int32)
        lm_mask[torch.from_numpy(np.unique(sh["cam"]["lm_idx"]).astype(np.int64))] = 1
        dist.all_reduce(lm_mask)
        assert bool((lm_mask <= 1).all())
        # ---- balance
        counts = torch.tensor([len(sh["gyro"]["t"]), len(sh["accel"]["t"]), len(sh["cam"]["lm_idx"])], dtype=torch.float64)
        mx, mn = counts.clone(), counts.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(mn, op=dist.ReduceOp.MIN)
        assert bool(((mx - mn) <= torch.tensor([1.0, 1.0, 0.1 * len(cfg["cam"]["lm_idx"])], dtype=torch.float64)).all())
        # ---- structure of the shard == restriction of the global structure (host-only handle, no GPU)
        def structure(c):
            p = _lib.Problem(-1)
            p.set_se3_spline(c["dt"], 0.0, len(c["knots"]))
            g = p.add_gyroscope(_lib.make_sensor(), c["gyro"]["t"], c["gyro"]["y"])
            cm = c["cam"]
            k = p.add_static_rs(_lib.make_camera(cm["rows"], cm["cols"], cm["readout"], cm["K"]), cm["obs_uv"], cm["obs_t0"], cm["ref_uv"], cm["ref_t0"], cm["lm_idx"])
            return p.get_structure(g, 4)[0], p.get_structure(k, 24)[0]
        g_all, c_all = structure(cfg)
        g_sh, c_sh = structure(sh)
        assert (g_sh == g_all[sh["gyro"]["rows"]]).all() and (c_sh == c_all[sh["cam"]["rows_sel"]]).all()
        # ---- the exchange a Gauss-Newton consumer needs: sum over ranks of a per-knot accumulator (gloo stands in for NCCL)
        acc = torch.zeros(len(cfg["knots"]), dtype=torch.float64)
        acc.index_add_(0, torch.from_numpy(g_sh.reshape(-1).astype(np.int64)), torch.ones(g_sh.size, dtype=torch.float64))
        dist.all_reduce(acc)
        ref = np.bincount(g_all.reshape(-1), minlength=len(cfg["knots"])).astype(float)
        assert np.array_equal(acc.numpy(), ref)
        result_q.put((rank, "ok"))
    except Exception as e:     # noqa: BLE001
        result_q.put((rank, f"{type(e).__name__}: {e}"))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_shards_world_size_2_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(0, "ok"), (1, "ok")], results
