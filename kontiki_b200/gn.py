"""Gauss-Newton / Levenberg-Marquardt step on the device, matrix-free.

The reference hands the normal equations to Ceres (SPARSE_SCHUR, cpplib/include/kontiki/trajectory_estimator.h:38-64).
Here the Jacobian rows never leave the GPU: every product with J or J^T is a pass of libkontiki_b200's k_j_apply /
k_jt_apply kernels over the packed rows, the LM system (J^T J + D/radius) delta = -J^T r is solved by preconditioned
conjugate gradients on device vectors (torch is only the vector plumbing), and with measurements sharded over several GPUs
the one exchange per product is an all-reduce of a parameter-sized vector (NCCL over NVLink when torch.distributed is
initialised with the nccl backend; knots and inverse depths are replicated, rows are not).
"""
import numpy as np
import torch

from . import _lib


def _dist():
    try:
        import torch.distributed as dist
        return dist if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1 else None
    except Exception:      # noqa: BLE001
        return None


class DeviceNormalEquations:
    """Residuals / Jacobian rows of one ktk problem held in device memory + the products a GN solver needs.

    Local (tangent) coordinates: SE3 knots 6 (Plus = T exp(delta)), R3 knots 3, SO3 knots 3 (Plus = q_delta q), rho 1.
    """

    def __init__(self, problem, split, n_a, n_b, n_rho, device, robust=True):
        self.p, self.split, self.n_a, self.n_b, self.n_rho = problem, split, n_a, n_b, n_rho
        self.dev = torch.device("cuda", device)
        self.flags = _lib.EVAL_RESIDUALS | _lib.EVAL_JACOBIANS | _lib.EVAL_DEVICE_ORDER | (_lib.EVAL_ROBUST if robust else 0)   # rows stay on the device: device order
        self.n_amb = problem.num_parameters(n_rho)
        self.n_loc = (3 * n_a + 3 * n_b if split else 6 * n_a) + n_rho
        self.outs, self._keep, self.u = [], [], []
        for g in range(problem.num_groups):
            n, cam = problem.group_size(g), problem.group_kind(g) in (_lib.STATIC_RS, _lib.NEWTON_RS)
            r = torch.zeros((n, 2 if cam else (1 if problem.group_kind(g) == _lib.ORIENTATION else 3)), dtype=torch.float64, device=self.dev)
            J = torch.zeros((n, problem.group_row_size(g)), dtype=torch.float64, device=self.dev)
            idx = [torch.zeros(n, dtype=torch.int32, device=self.dev) for _ in range(4)]
            self._keep.append((r, J, idx))
            self.outs.append(dict(r=r.data_ptr(), J=J.data_ptr(), i0=idx[0].data_ptr(), i0_b=idx[1].data_ptr(), i0_c=idx[2].data_ptr(), i0_d=idx[3].data_ptr()))
            self.u.append(torch.zeros_like(r))
        self.problem_stream = torch.cuda.current_stream(self.dev)
        problem.set_stream(self.problem_stream.cuda_stream)
        self.free = torch.ones(self.n_loc, dtype=torch.float64, device=self.dev)      # 0 for constant (locked) parameters
        self._huber_sq = {}

    def set_huber(self, group, huber_c):
        """Huber constants of a camera group (caller order), for the cost 1/2 sum rho(s); rows live in device order."""
        order = self.p.get_row_order(group)
        self._huber_sq[group] = torch.from_numpy(np.asarray(huber_c, np.float64)[order] ** 2).to(self.dev)

    # ---- parameter point ------------------------------------------------------------------------------------------------
    def set_point(self, knots_flat, rho, P_a, P_b):
        """knots_flat: ambient [knots | ...] as for ktk_evaluate; P_a / P_b: d Plus / d delta of the knots (numpy, may be None)."""
        self.knots = torch.from_numpy(np.ascontiguousarray(knots_flat, np.float64)).to(self.dev)
        self.rho = torch.from_numpy(np.ascontiguousarray(rho if rho is not None else np.zeros(0), np.float64)).to(self.dev)
        self.P_a = None if P_a is None else torch.from_numpy(np.ascontiguousarray(P_a)).to(self.dev)
        self.P_b = None if P_b is None else torch.from_numpy(np.ascontiguousarray(P_b)).to(self.dev)

    def evaluate(self):
        """One batched residual + Jacobian evaluation; returns the cost 1/2 sum |r|^2 summed over ranks."""
        self.p.evaluate_device(self.knots.data_ptr(), self.rho.data_ptr() if self.n_rho else 0, self.n_rho, self.flags, self.outs)
        self.p.synchronize()
        # Ceres' cost is 1/2 sum rho(s); the rows carry the corrected residual, |r_c|^2 = a sqrt(s) in Huber's linear region (estimator._cost)
        cost = torch.zeros((), dtype=torch.float64, device=self.dev)
        for g, k in enumerate(self._keep):
            s = (k[0] ** 2).sum(1)
            a2 = self._huber_sq.get(g)
            if a2 is not None and (self.flags & _lib.EVAL_ROBUST):
                s = torch.where(s > a2, 2.0 * s - a2, s)
            cost = cost + 0.5 * s.sum()
        cost = cost.reshape(1).clone()
        d = _dist()
        if d is not None:
            d.all_reduce(cost)
        return float(cost.item())

    # ---- ambient <-> local ------------------------------------------------------------------------------------------------
    def _to_ambient(self, v):
        na, nb = self.n_a, self.n_b
        if not self.split:
            vk = torch.einsum("nad,nd->na", self.P_a, v[:6 * na].view(na, 6)).reshape(-1)
            return torch.cat([vk, v[6 * na:]])
        vb = torch.einsum("nad,nd->na", self.P_b, v[3 * na:3 * na + 3 * nb].view(nb, 3)).reshape(-1)
        return torch.cat([v[:3 * na], vb, v[3 * na + 3 * nb:]])

    def _to_local(self, y):
        na, nb = self.n_a, self.n_b
        if not self.split:
            yk = torch.einsum("nad,na->nd", self.P_a, y[:7 * na].view(na, 7)).reshape(-1)
            return torch.cat([yk, y[7 * na:]])
        yb = torch.einsum("nad,na->nd", self.P_b, y[3 * na:3 * na + 4 * nb].view(nb, 4)).reshape(-1)
        return torch.cat([y[:3 * na], yb, y[3 * na + 4 * nb:]])

    def _reduce(self, y):
        d = _dist()
        if d is not None:
            d.all_reduce(y)           # the ONE exchange of the sharded problem: a parameter-sized fp64 vector
        return y

    # ---- products ---------------------------------------------------------------------------------------------------------
    def gradient(self):
        """g = P^T J^T r (local coordinates), summed over ranks."""
        y = torch.zeros(self.n_amb, dtype=torch.float64, device=self.dev)
        self.p.jt_apply(self.outs, [k[0].data_ptr() for k in self._keep], y.data_ptr(), _lib.EVAL_DEVICE_ORDER)
        return self._to_local(self._reduce(y)) * self.free

    def hessian_apply(self, v):
        """(P^T J^T J P) v."""
        va = self._to_ambient(v * self.free).contiguous()
        self.p.j_apply(self.outs, va.data_ptr(), [u.data_ptr() for u in self.u], _lib.EVAL_DEVICE_ORDER)
        y = torch.zeros(self.n_amb, dtype=torch.float64, device=self.dev)
        self.p.jt_apply(self.outs, [u.data_ptr() for u in self.u], y.data_ptr(), _lib.EVAL_DEVICE_ORDER)
        return self._to_local(self._reduce(y)) * self.free

    def hessian_diagonal(self):
        y = torch.zeros(self.n_loc, dtype=torch.float64, device=self.dev)
        self.p.jtj_diagonal_local(self.outs, 0 if self.P_a is None else self.P_a.data_ptr(), 0 if self.P_b is None else self.P_b.data_ptr(), y.data_ptr(), _lib.EVAL_DEVICE_ORDER)
        return self._reduce(y) * self.free


def pcg(apply_A, b, Minv, tol=1e-10, max_iter=500):
    """Jacobi-preconditioned conjugate gradients on device vectors; returns (x, iterations)."""
    x = torch.zeros_like(b)
    r = b.clone()
    z = Minv * r
    p = z.clone()
    rz = torch.dot(r, z)
    b_norm = torch.linalg.vector_norm(b)
    if float(b_norm) == 0.0:
        return x, 0
    it = 0
    for it in range(1, max_iter + 1):
        Ap = apply_A(p)
        alpha = rz / torch.dot(p, Ap)
        x += alpha * p
        r -= alpha * Ap
        if float(torch.linalg.vector_norm(r)) <= tol * float(b_norm):
            break
        z = Minv * r
        rz_new = torch.dot(r, z)
        p = z + (rz_new / rz) * p
        rz = rz_new
    return x, it
