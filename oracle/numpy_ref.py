"""ORACLE — TEST INFRASTRUCTURE ONLY.  Second, independently written restatement (NumPy).

Written from the reference's behaviour, not from oracle/stomp_oracle.cpp, and vectorised differently
(whole-array expressions instead of ordered loops), so that a misreading of the reference in one of the
two restatements shows up as a disagreement in tests/test_oracle_cross.py.  (The C++ restatement is, in addition,
pinned against the reference's own compiled code: oracle/ref, tests/test_reference_pin.py.)

Citations are relative to /root/reference/src/planners/.
"""
from __future__ import annotations

import numpy as np

PAD = 6            # stomp/include/stomp/StompUtils.hpp:57 TRAJECTORY_PADDING
RULES = np.array([  # StompUtils.hpp:60-66
    [0, 0, 0, 1, 0, 0, 0],
    [0, 0, -1, 1, 0, 0, 0],
    [0, -1 / 12.0, 16 / 12.0, -30 / 12.0, 16 / 12.0, -1 / 12.0, 0],
    [0, 1 / 12.0, -17 / 12.0, 46 / 12.0, -46 / 12.0, 17 / 12.0, -1 / 12.0],
])


def diff_matrix(n, order, dt):
    """stomp/src/StompUtils.cpp:6-23"""
    m = np.zeros((n, n))
    mult = 1.0 / dt ** order
    for i in range(n):
        for j in range(-3, 4):
            m[i, min(max(i + j, 0), n - 1)] += mult * RULES[order][j + 3]
    return m


class Policy:
    """stomp/src/CovariantMovementPrimitive.cpp"""

    def __init__(self, T, D, duration, initial_all, weights=(0.0, 0.0, 1.0, 0.0)):
        self.T, self.D, self.N = T, D, T + 2 * PAD
        self.dt = duration / (T + 1)                                         # :204
        self.w = np.tile(np.asarray(weights, dtype=np.float64), (self.N, 1))  # derivative_costs_[d] (same for all d)
        self.Dm = [diff_matrix(self.N, r, self.dt) for r in range(4)]       # :292-301
        self.R_all = sum(self.dt * (self.Dm[r].T @ np.diag(self.w[:, r]) @ self.Dm[r]) for r in range(4))  # :258-266
        self.R = self.R_all[PAD:PAD + T, PAD:PAD + T].copy()
        self.Rinv = np.linalg.inv(self.R)                                    # :273 (fullPivLu there)
        self.params_all = np.array(initial_all, dtype=np.float64)            # [D][N]
        self.linear()

    def linear(self):
        """:136-172"""
        T, N = self.T, self.N
        head, tail = self.params_all[:, :PAD], self.params_all[:, N - PAD:]
        lin = head @ self.R_all[:PAD, PAD:PAD + T] + tail @ self.R_all[N - PAD:, PAD:PAD + T]
        lin *= 2.0
        lin += -self.dt * 2.0 * self.params_all[:, PAD:PAD + T] * self.w[PAD:PAD + T, 0]
        self.lin = lin

    def to_min_control_cost(self):
        """:174-189"""
        self.params_all[:, PAD:PAD + self.T] = -0.5 * (self.Rinv @ self.lin.T).T
        self.mincc = self.params_all[:, PAD:PAD + self.T].copy()

    def set_min_control_cost(self, params_all):
        """:191-200"""
        self.mincc = np.array(params_all)[:, PAD:PAD + self.T].copy()

    @property
    def params(self):
        return self.params_all[:, PAD:PAD + self.T]

    def control_costs(self, x_free, weight):
        """:327-412.  x_free [..., D, T] = parameters + projected noise -> [..., D, T]"""
        T, N = self.T, self.N
        lead = x_free.shape[:-2]
        x = np.broadcast_to(self.params_all, lead + self.params_all.shape).copy()
        x[..., PAD:PAD + T] = x_free
        c_all = np.zeros_like(x)
        for r in range(4):
            Ax = (x @ self.Dm[r].T) * np.sqrt(self.w[:, r])
            c_all += self.dt * weight * Ax * Ax
        cc = c_all[..., PAD:PAD + T].copy()
        cc[..., 0] += c_all[..., :PAD].sum(-1)
        cc[..., T - 1] += c_all[..., N - PAD:].sum(-1)
        return cc


# ---- sphere / SDF task --------------------------------------------------------------------------

def _rot_axis(axis, q):
    a = np.asarray(axis, dtype=np.float64)
    K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    s, c = np.sin(q)[..., None, None], np.cos(q)[..., None, None]
    return np.eye(3) + s * K + (1 - c) * (K @ K)


def _rpy_matrix(rpy):
    r, p, y = rpy
    cx, sx, cy, sy, cz, sz = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def sphere_centres(chain, spheres, q):
    """q [..., D] -> centres [..., S, 3]; textbook FK with libm sin/cos"""
    q = np.asarray(q, dtype=np.float64)
    lead = q.shape[:-1]
    out = np.zeros(lead + (len(spheres.link), 3))
    R = np.broadcast_to(np.eye(3), lead + (3, 3)).copy()
    p = np.zeros(lead + (3,))
    for d in range(q.shape[-1]):
        if chain.parent[d] < 0:
            R = np.broadcast_to(np.eye(3), lead + (3, 3)).copy()
            p = np.zeros(lead + (3,))
        p = p + R @ chain.origin_xyz[d]
        R = R @ _rpy_matrix(chain.origin_rpy[d])
        if chain.prismatic[d]:
            p = p + q[..., d, None] * (R @ chain.axis[d])
        else:
            R = R @ _rot_axis(chain.axis[d], q[..., d])
        for s in np.nonzero(spheres.link == d)[0]:
            out[..., s, :] = p + R @ spheres.xyz[s]
    return out


def collides(chain, spheres, sdf, q, return_margin=False):
    """q [..., D] -> bool [...]: any sphere with sdf(centre) - r < 0 (nearest voxel, clamped)"""
    c = sphere_centres(chain, spheres, q)
    f = (c - sdf.origin) * (1.0 / sdf.voxel)
    idx = np.clip(np.floor(f), 0, np.asarray(sdf.dims) - 1).astype(np.int64)
    val = sdf.grid[idx[..., 2], idx[..., 1], idx[..., 0]].astype(np.float64) - spheres.radius
    hit = (val < 0.0).any(-1)
    if return_margin:
        # distance of every coordinate to the nearest voxel face, in voxels (verdict-stability check)
        # (coordinates that are exactly integral come from exact arithmetic, e.g. spheres on the base
        # axis, and are exact in every implementation: they are left out)
        frac = np.abs(f - np.rint(f))
        frac = frac[frac > 0.0]
        return hit, (frac.min() if frac.size else 1.0), np.abs(val).min()
    return hit


class NumpyStomp:
    """stomp/src/PolicyImprovement.cpp + stomp/src/Stomp.cpp + wrappers/stomp/OptimizationTask.cpp with
    injected post-Cholesky noise."""

    def __init__(self, problem, *, min_rollouts, max_rollouts, per_iteration, noise_stddev, noise_decay=None,
                 noise_min_stddev=None, use_noise_adaptation=True, control_cost_weight=0.001, duration=5.0,
                 use_cumulative_costs=True, query=None):
        self.pb = problem
        self.T, self.D = problem.num_time_steps, problem.chain.num_dimensions
        T, D = self.T, self.D
        s, g = problem.start, problem.goal
        if s.ndim == 2:
            s, g = s[query or 0], g[query or 0]
        init = np.zeros((D, T + 2 * PAD))                                    # OptimizationTask.cpp:46-66
        init[:, :PAD] = s[:, None]
        init[:, PAD + T:] = g[:, None]
        inc = (g - s) / (T - 1)
        init[:, PAD:PAD + T] = s[:, None] + np.arange(T)[None, :] * inc[:, None]
        self.policy = Policy(T, D, duration, init)
        self.policy.to_min_control_cost()                                    # :108-119
        self.min_r, self.max_r, self.per_it = min_rollouts, max_rollouts, per_iteration
        self.sigma0 = np.asarray(noise_stddev, dtype=np.float64)
        self.decay = np.ones(D) if noise_decay is None else np.asarray(noise_decay, dtype=np.float64)
        self.sigma_min = np.full(D, 0.01) if noise_min_stddev is None else np.asarray(noise_min_stddev, dtype=np.float64)
        self.adapt, self.ccw, self.cum = use_noise_adaptation, control_cost_weight, use_cumulative_costs
        self.h = 10.0                                                        # PolicyImprovement.cpp:55
        self.n = 0
        self.adapted_valid = False
        self.sigma = np.ones(D)
        self.noiseless = None
        # per-rollout storage, index 0..n-1
        self.theta_noisy = np.zeros((0, D, T)); self.theta_proj = np.zeros((0, D, T))
        self.noise = np.zeros((0, D, T)); self.state = np.zeros((0, T)); self.total = np.zeros(0)

    def state_costs(self, theta):
        """theta [K][D][T] -> cost [K][T] in {0,1}; validity [K] = last timestep (OptimizationTask.cpp:183-204)"""
        q = np.moveaxis(theta, -2, -1)
        hit = collides(self.pb.chain, self.pb.spheres, self.pb.sdf, q)
        return hit.astype(np.float64), ~hit[..., -1]

    def iterate(self, it, unit_noise):
        """Stomp::runSingleIteration (Stomp.cpp:274-301); unit_noise [gen][D][T]"""
        pol, T, D = self.policy, self.T, self.D
        theta = pol.params.copy()
        if not self.adapted_valid:                                           # PolicyImprovement.cpp:162-163
            self.sigma = self.sigma0 * self.decay ** (it - 1)                # Stomp.cpp:179
        # bookkeeping :170-186
        prev = self.n
        gen = self.per_it
        reused = prev
        if prev + gen < self.min_r:
            gen = self.min_r - prev
        if prev + gen > self.max_r:
            reused = prev - (prev + gen - self.max_r)
        assert unit_noise.shape == (gen, D, T)
        if reused > 0:                                                       # :188-255
            lo, hi = self.total[:prev].min(), self.total[:prev].max()
            den = max(hi - lo, 1e-8)
            w = np.exp(-self.h * (self.total[:prev] - lo) / den)
            order = sorted(range(prev), key=lambda r: (-w[r], r))[:reused]
            k_proj = self.theta_proj[order]
            k_state = self.state[order]
            k_noise = k_proj - theta                                         # noise re-based on the new parameters
        # new rollouts :258-286
        l1 = self.ccw
        l2 = 1.0 / (self.sigma * self.sigma)
        new_sd = 1.0 / np.sqrt(l1 + l2)
        p1, p2 = l1 / (l1 + l2), l2 / (l1 + l2)
        noisy = (p1[:, None] * pol.mincc + p2[:, None] * theta)[None] + new_sd[None, :, None] * unit_noise
        noisy = np.clip(noisy, self.pb.chain.lower[None, :, None], self.pb.chain.upper[None, :, None])  # filter :85-106
        g_noise = noisy - theta
        g_proj = theta + g_noise
        g_state, self.gen_validity = self.state_costs(noisy)
        parts_proj, parts_noise, parts_state = [g_proj], [g_noise], [g_state]
        parts_noisy = [noisy]
        if reused > 0:
            parts_proj.append(k_proj); parts_noise.append(k_noise); parts_state.append(k_state)
            parts_noisy.append(theta + k_noise)
        if self.noiseless is not None:                                       # :304-308
            parts_proj.append(self.noiseless["theta"][None]); parts_noise.append(np.zeros((1, D, T)))
            parts_state.append(self.noiseless["state"][None]); parts_noisy.append(self.noiseless["theta"][None])
        self.theta_proj = np.concatenate(parts_proj); self.noise = np.concatenate(parts_noise)
        self.state = np.concatenate(parts_state); self.theta_noisy = np.concatenate(parts_noisy)
        self.n = n = len(self.state)
        self.gen = gen
        # costs :442-495.  the control cost is evaluated on parameters_ + noise_projected_
        base = np.broadcast_to(theta, (n, D, T)).copy()
        if self.noiseless is not None:
            base[-1] = self.noiseless["theta"]
        self.control = pol.control_costs(base + self.noise, self.ccw)
        S = self.state.sum(-1)
        Cd = self.control.sum(-1)
        self.full_costs = S[:, None] + Cd
        self.total = S + Cd.sum(-1)
        tot = self.state[:, None, :] + self.control
        self.cumulative = np.broadcast_to(tot.sum(-1, keepdims=True), tot.shape).copy() if self.cum else tot
        # probabilities :497-582 (min / max over all rollouts AND timesteps)
        lo = self.cumulative.min(axis=(0, 2), keepdims=True); hi = self.cumulative.max(axis=(0, 2), keepdims=True)
        den = np.maximum(hi - lo, 1e-8)
        p = np.exp(-self.h * (self.cumulative - lo) / den)
        self.prob = p / p.sum(0, keepdims=True)
        lo, hi = self.full_costs.min(0), self.full_costs.max(0)
        den = np.maximum(hi - lo, 1e-8)
        pf = np.exp(-self.h * (self.full_costs - lo) / den)
        self.full_prob = pf / pf.sum(0)
        # update :584-711
        self.updates = (self.noise * self.prob).sum(0)
        if self.adapt:
            q = np.einsum("rdt,tu,rdu->rd", self.noise, pol.R, self.noise)
            frob = np.sqrt((self.full_prob * q).sum(0) / (self.full_prob.sum(0) * T))
            self.sigma = np.maximum(0.8 * self.sigma + 0.2 * frob, self.sigma_min)
            self.adapted_valid = True
        pol.params_all[:, PAD:PAD + T] += self.updates                      # CovariantMovementPrimitive.cpp:476-479
        # noiseless rollout Stomp.cpp:253-272, PolicyImprovement.cpp:401-419
        th = pol.params.copy()
        st, val = self.state_costs(th[None])
        cc = pol.control_costs(th, self.ccw)
        self.noiseless = dict(theta=th, state=st[0], control=cc, total=st[0].sum() + cc.sum(), valid=bool(val[0]))
        return self.noiseless["total"]
