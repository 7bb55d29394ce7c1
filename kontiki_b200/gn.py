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

    def evaluate(self, cost=True):
        """One batched residual + Jacobian evaluation; returns the cost 1/2 sum rho(s) summed over ranks (cost=False: enqueue only, no
        host synchronisation)."""
        self.p.evaluate_device(self.knots.data_ptr(), self.rho.data_ptr() if self.n_rho else 0, self.n_rho, self.flags, self.outs)
        if not cost:
            return None
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


class _DevArray:
    """A raw device pointer as something torch.as_tensor() can wrap without a copy."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f8", "data": (int(ptr), False), "version": 2}


class DeviceSchurSolver:
    """Levenberg-Marquardt step on the device (csrc/gn_device.cuh behind the ktk_gn_* calls of the C ABI): rho eliminated (implicit Schur
    complement), block-Jacobi preconditioned CG on the knots with its scalars in device memory, Plus() on the device, every reduction a
    fixed-order gather.  This class is the plumbing: it owns the parameter point on the device, enqueues the calls, and -- when
    torch.distributed is initialised (rows sharded by kontiki_b200/sharding.py: every landmark's rows on ONE rank) -- all-reduces the
    parameter-sized buffers between them (NCCL over NVLink).  Host round trips: one for the cost, one per `check_every` CG iterations,
    one for the model decrease."""

    def __init__(self, problem, split, n_a, n_b, n_rho, device, lm_locked=None, lock_a=False, lock_b=False, hubers=None, robust=True):
        self.p, self.split, self.n_a, self.n_b, self.n_rho = problem, split, n_a, n_b, n_rho
        self.dev = torch.device("cuda", device)
        self.flags = _lib.EVAL_RESIDUALS | _lib.EVAL_JACOBIANS | _lib.EVAL_DEVICE_ORDER | (_lib.EVAL_ROBUST if robust else 0)
        self.lm_locked, self.lock_a, self.lock_b, self.hubers = lm_locked, lock_a, lock_b, hubers
        self.outs, self._keep = [], []
        for g in range(problem.num_groups):
            n, cam = problem.group_size(g), problem.group_kind(g) in (_lib.STATIC_RS, _lib.NEWTON_RS)
            r = torch.zeros((n, 2 if cam else (1 if problem.group_kind(g) == _lib.ORIENTATION else 3)), dtype=torch.float64, device=self.dev)
            J = torch.zeros((n, problem.group_row_size(g)), dtype=torch.float64, device=self.dev)
            idx = [torch.zeros(n, dtype=torch.int32, device=self.dev) for _ in range(4)]
            self._keep.append((r, J, idx))
            self.outs.append(dict(r=r.data_ptr(), J=J.data_ptr(), i0=idx[0].data_ptr(), i0_b=idx[1].data_ptr(), i0_c=idx[2].data_ptr(), i0_d=idx[3].data_ptr()))
        self.stream = torch.cuda.current_stream(self.dev)
        problem.set_stream(self.stream.cuda_stream)
        self.prepared = False
        self.knots = self.rho = self.knots_new = self.rho_new = None
        self._buf = {}

    # ---- point ------------------------------------------------------------------------------------------------------------------
    def set_point(self, knots_flat, rho):
        self.knots = torch.from_numpy(np.ascontiguousarray(knots_flat, np.float64)).to(self.dev)
        self.rho = torch.from_numpy(np.ascontiguousarray(rho if rho is not None else np.zeros(0), np.float64)).to(self.dev)
        self.knots_new, self.rho_new = torch.empty_like(self.knots), torch.empty_like(self.rho)

    def point(self):
        """(knots_flat, rho) of the current point as numpy arrays (one device -> host copy)."""
        return self.knots.cpu().numpy(), self.rho.cpu().numpy()

    def buf(self, name):
        t = self._buf.get(name)
        if t is None:
            ptr, n = self.p.gn_buffer(name)
            t = self._buf[name] = torch.as_tensor(_DevArray(ptr, n), device=self.dev) if n > 0 and ptr else torch.zeros(0, dtype=torch.float64, device=self.dev)
        return t

    def _reduce(self, *names):
        d = _dist()
        if d is not None:
            for n in names:
                t = self.buf(n)
                if t.numel():
                    d.all_reduce(t)

    # ---- steps ------------------------------------------------------------------------------------------------------------------
    def evaluate(self, at_new=False, cost=True):
        """Residuals + Jacobian rows at the current (or the candidate) point; returns the cost 1/2 sum rho(s) over all ranks
        (cost=False: enqueue only, no host synchronisation)."""
        k, r = (self.knots_new, self.rho_new) if at_new else (self.knots, self.rho)
        self.p.evaluate_device(k.data_ptr(), r.data_ptr() if self.n_rho else 0, self.n_rho, self.flags, self.outs)
        if not cost and self.prepared:
            return None
        if not self.prepared:
            self.p.synchronize()                      # raises where a measurement left the trajectory
            self.p.gn_prepare(self.flags, self.outs, self.n_rho, self.lm_locked, self.lock_a, self.lock_b, self.hubers)
            self.prepared, self._buf = True, {}
        self.p.gn_call("cost")
        scal = self.buf("scal")
        cost = scal[12:13].clone()
        d = _dist()
        if d is not None:
            d.all_reduce(cost)
        self.p.synchronize()
        return float(cost.item())

    def linearize(self, radius):
        """Normal equations at the current point (the rows of the last evaluate() must be the current point's); returns max |gradient|."""
        self.linearize_local()
        self._reduce("lin")               # [c | grho | blocks_a | blocks_b | z_a | z_b] in ONE collective
        gmax = self.linearize_rhs(radius)
        self._reduce("qq")
        return gmax

    # the two rank-local halves of linearize(), callable on their own (bench.py captures each stretch between two collectives in a CUDA graph)
    def linearize_local(self):
        import ctypes as C
        self.p.gn_call("linearize_local", C.c_void_p(self.knots.data_ptr()), C.c_void_p(self.rho.data_ptr()) if self.n_rho else None)
        self.p.gn_call("gradient_local")

    def linearize_rhs(self, radius):
        import ctypes as C
        gmax = torch.zeros(1, dtype=torch.float64, device=self.dev)
        for name, locked in (("z_a", self.lock_a), ("z_b", self.lock_b)):
            t = self.buf(name)
            if t.numel() and not locked:
                gmax = torch.maximum(gmax, t.abs().max().reshape(1))
        g = self.buf("grho")
        if g.numel():
            free = torch.ones_like(g) if self.lm_locked is None else torch.from_numpy(1.0 - np.asarray(self.lm_locked, np.float64)).to(self.dev)
            gmax = torch.maximum(gmax, (g * free).abs().max().reshape(1))
        self.p.gn_call("linearize_rhs", C.c_double(radius))
        return gmax

    def solve(self, radius, tol=1e-6, max_iter=300, check_every=8):
        """CG on the reduced system; returns (iterations, |r|/|b|)."""
        import ctypes as C
        self.p.gn_call("pcg_begin", C.c_double(radius), C.c_double(tol), C.c_int32(max_iter))
        it, done, rel = 0, False, 1.0
        while not done:
            for _ in range(check_every):
                self.p.gn_call("product")
                self._reduce("qq")                # [q_a | q_b]
                self.p.gn_call("pcg_update")
            it, done, rel = self.p.gn_pcg_status()
        return it, rel

    def finish(self):
        """delta_rho, the model decrease -(delta^T g + 1/2 delta^T J^T J delta) and |delta|; retracts into the candidate point."""
        import ctypes as C
        self.p.gn_call("finish_local")
        if _dist() is not None:
            self.p.gn_call("finish_mask")
            self._reduce("drho")
        self.p.gn_call("model_local")
        scal = self.buf("scal")
        sums = scal[9:11].clone()
        d = _dist()
        if d is not None:
            d.all_reduce(sums)
        dr = self.buf("drho")
        step2 = scal[11:12] + ((dr * dr).sum().reshape(1) if dr.numel() else 0.0)
        self.p.gn_call("retract", C.c_void_p(self.knots.data_ptr()), C.c_void_p(self.rho.data_ptr()) if self.n_rho else None, C.c_void_p(self.knots_new.data_ptr()),
                       C.c_void_p(self.rho_new.data_ptr()) if self.n_rho else None)
        vals = torch.cat([sums, step2]).cpu().numpy()        # the one host round trip of this phase
        return -(float(vals[0]) + 0.5 * float(vals[1])), float(np.sqrt(vals[2]))

    def accept(self):
        self.knots, self.knots_new = self.knots_new, self.knots
        self.rho, self.rho_new = self.rho_new, self.rho


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
