"""integration/batched_backend.h -- the reference-side binding INTEGRATION.md describes (a ceres::EvaluationCallback around ktk_evaluate plus thin
ceres::CostFunction objects that copy their row) -- compiled with g++ against integration/ceres_stub (Ceres is not installed here; the stub declares
the two interfaces with Ceres' signatures) and libkontiki_b200.so.  CPU: it compiles and links against every symbol it uses.  GPU: the C++ driver
integration/example_main.cc evaluates a small problem through the binding exactly as ceres::Solve's residual-block loop would, and what Ceres would
have received equals what the Python binding returns for the same problem."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "tests", "_build")
EXE = os.path.join(BUILD, "example_main")


def _compile():
    from kontiki_b200 import build as kb
    kb.build()
    os.makedirs(BUILD, exist_ok=True)
    libdir = os.path.dirname(kb.LIB_PATH)
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "integration", "ceres_stub"), "-I" + os.path.join(ROOT, "include"),
           "-I" + os.path.join(ROOT, "integration"), os.path.join(ROOT, "integration", "example_main.cc"), "-o", EXE, "-L" + libdir, "-lkontiki_b200",
           "-Wl,-rpath," + libdir, "-Wl,-rpath,/usr/local/cuda/lib64", "-L/usr/local/cuda/lib64"]
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    return EXE


def test_reference_side_binding_compiles_and_links():
    exe = _compile()
    assert os.path.exists(exe)
    res = subprocess.run([exe], capture_output=True, text=True)      # no arguments: usage, exit code 2 (no CUDA call is made)
    assert res.returncode == 2 and "usage" in res.stderr


@pytest.mark.gpu
def test_binding_hands_ceres_what_the_python_binding_returns(tmp_path):
    from kontiki_b200 import _lib, synthetic as syn
    exe = _compile()
    cfg = syn.make_config("H1", scale=0.001)      # 50 gyro + 500 camera rows on 5k knots
    c, g = cfg["cam"], cfg["gyro"]
    ng, nc, n_lm = len(g["t"]), len(c["lm_idx"]), len(c["rho"])
    head = np.array([len(cfg["knots"]), cfg["dt"], cfg["t0"], ng, nc, n_lm, c["rows"], c["cols"], c["readout"]] + list(np.asarray(c["K"], float).reshape(-1)))
    parts = [head, cfg["knots"].reshape(-1), g["t"], g["y"].reshape(-1), g["weight"], c["obs_uv"].reshape(-1), c["obs_t0"], c["ref_uv"].reshape(-1), c["ref_t0"],
             c["lm_idx"].astype(float), c["weight"], c["huber_c"], c["rho"]]
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    np.concatenate([np.asarray(p, float).reshape(-1) for p in parts]).tofile(fin)
    res = subprocess.run([exe, fin, fout], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    out = np.fromfile(fout)
    # the same problem through the Python binding (compat_zero_dB does not touch gyroscope / camera rows)
    p = _lib.Problem(0)
    p.set_se3_spline(cfg["dt"], cfg["t0"], len(cfg["knots"]), compat_zero_dB=True)
    gg = p.add_gyroscope(_lib.make_sensor(), g["t"], g["y"], g["weight"])
    gc = p.add_static_rs(_lib.make_camera(c["rows"], c["cols"], c["readout"], c["K"]), c["obs_uv"], c["obs_t0"], c["ref_uv"], c["ref_t0"], c["lm_idx"], c["weight"], c["huber_c"])
    outs = p.evaluate(cfg["knots"], c["rho"], _lib.EVAL_RESIDUALS | _lib.EVAL_JACOBIANS)      # no KTK_EVAL_ROBUST: the loss stays with Ceres
    gy = out[:88 * ng].reshape(ng, 88)
    assert np.array_equal(gy[:, 0].astype(int), outs[gg]["i0"]) and np.array_equal(gy[:, 1:4], outs[gg]["r"])
    assert np.array_equal(gy[:, 4:].reshape(ng, 4, 3, 7), outs[gg]["J"])
    cap = 24
    cy = out[88 * ng:].reshape(nc, 1 + cap + 2 + cap * 14 + 2)
    ids, nids = p.get_structure(gc, cap=cap)
    assert np.array_equal(cy[:, 0].astype(int), nids) and np.array_equal(cy[:, 1:1 + cap].astype(int), ids)
    assert np.array_equal(cy[:, 1 + cap:3 + cap], outs[gc]["r"])
    Js = p.expand_static_rs(gc, ids, outs[gc]["J"], outs[gc]["i0"], outs[gc]["i0_b"])      # (n, cap, 2, 7)
    assert np.array_equal(cy[:, 3 + cap:3 + cap + cap * 14].reshape(nc, cap, 2, 7), Js)
    assert np.array_equal(cy[:, -2:], outs[gc]["J"][:, 112:114])
