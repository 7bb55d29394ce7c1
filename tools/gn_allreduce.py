#!/usr/bin/env python
"""BASELINE.json configs[3] ("Full VI-SfM ... J^T J allreduce at 1/2/4/8 x B200"): the Gauss-Newton products of a problem
whose measurements are sharded over the ranks, with ONE NCCL all-reduce of a parameter-sized fp64 vector per product.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/gn_allreduce.py [--scale 1.0]

Checks (rank 0 also evaluates the UNSHARDED problem on its GPU): cost, gradient J^T r and (J^T J) v of the sharded run equal
the single-GPU ones to 1e-10 relative; reports the time per evaluation and per product (max over ranks, CUDA events).
"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kontiki_b200 import _lib, gn, sharding, synthetic as syn        # noqa: E402
from kontiki_b200.estimator import _se3_plus_jacobian                 # noqa: E402


def build(cfg, device):
    p = _lib.Problem(device)
    p.set_se3_spline(cfg["dt"], cfg["t0"], len(cfg["knots"]))
    imu = _lib.make_sensor()
    p.add_gyroscope(imu, cfg["gyro"]["t"], cfg["gyro"]["y"], cfg["gyro"]["weight"])
    p.add_accelerometer(imu, cfg["accel"]["t"], cfg["accel"]["y"], cfg["accel"]["weight"])
    c = cfg["cam"]
    p.add_static_rs(_lib.make_camera(c["rows"], c["cols"], c["readout"], c["K"]), c["obs_uv"], c["obs_t0"], c["ref_uv"], c["ref_t0"], c["lm_idx"], c["weight"], c["huber_c"])
    ne = gn.DeviceNormalEquations(p, False, len(cfg["knots"]), 0, len(c["rho"]), device)
    ne.set_point(cfg["knots"].reshape(-1), c["rho"], _se3_plus_jacobian(cfg["knots"]), None)
    return p, ne


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--iters", type=int, default=20)
    a = ap.parse_args()
    saved = os.dup(1)
    os.dup2(2, 1)
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cfg = syn.make_config("C4", scale=a.scale)           # 5k knots, 200k IMU + 500k static RS at scale 1
    shard = sharding.shard_config(cfg, rank, world)
    p, ne = build(shard, local)
    v = torch.from_numpy(np.random.default_rng(0).normal(size=ne.n_loc)).to(ne.dev)

    def timed(fn, iters):
        fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            out = fn()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / iters], dtype=torch.float64, device=ne.dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return out, float(ms.item())

    cost, ms_eval = timed(ne.evaluate, a.iters)
    g, ms_grad = timed(ne.gradient, a.iters)
    Hv, ms_hv = timed(lambda: ne.hessian_apply(v), a.iters)
    line = dict(config="C4 (SE3 5k knots, 200k IMU + 500k static RS) x scale %g" % a.scale, n_gpus=world, rows_total=syn.num_measurements(cfg),
                parameters=ne.n_loc, allreduce_bytes_per_product=ne.n_amb * 8, ms_evaluate=ms_eval, ms_gradient=ms_grad, ms_jtj_apply=ms_hv,
                jtj_products_per_s=1e3 / ms_hv)
    if rank == 0:
        # reference: the unsharded problem on this GPU (no exchange)
        saved_dist = gn._dist
        gn._dist = lambda: None
        p0, ne0 = build(cfg, local)
        cost0 = ne0.evaluate()
        g0, Hv0 = ne0.gradient(), ne0.hessian_apply(v)
        gn._dist = saved_dist
        line["rel_err_cost"] = abs(cost - cost0) / abs(cost0)
        line["rel_err_gradient"] = float((g - g0).abs().max() / g0.abs().max())
        line["rel_err_jtj_apply"] = float((Hv - Hv0).abs().max() / Hv0.abs().max())
        line["ok"] = bool(max(line["rel_err_cost"], line["rel_err_gradient"], line["rel_err_jtj_apply"]) < 1e-10)
        sys.stdout.flush()
        os.dup2(saved, 1)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
