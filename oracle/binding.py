"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes binding of oracle/liboracle.so (the CPU restatement of the reference's STOMP loop).  Imported
only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs; the
product package never imports it.  Pinned against the reference's own code (oracle/ref, oracle/ref_binding.py,
tests/test_reference_pin.py, tests/golden): see oracle/stomp_oracle.hpp.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
MAX_DIMS = 32


class OracleConfig(C.Structure):
    _fields_ = [
        ("num_time_steps", C.c_int32), ("num_dimensions", C.c_int32),
        ("min_rollouts", C.c_int32), ("max_rollouts", C.c_int32),
        ("num_rollouts_per_iteration", C.c_int32), ("num_iterations", C.c_int32),
        ("movement_duration", C.c_double), ("control_cost_weight", C.c_double),
        ("min_cost_improvement", C.c_double),
        ("noise_stddev", C.c_double * MAX_DIMS), ("noise_decay", C.c_double * MAX_DIMS),
        ("noise_min_stddev", C.c_double * MAX_DIMS),
        ("use_noise_adaptation", C.c_int32), ("use_openmp", C.c_int32),
        ("use_cumulative_costs", C.c_int32), ("use_projection", C.c_int32),
        ("per_timestep_minmax", C.c_int32), ("dense_control_costs", C.c_int32),
        ("seed", C.c_uint64),
    ]


def build(force: bool = False) -> str:
    """Compile liboracle.so with the committed Makefile (g++ only; seconds)."""
    if force or not os.path.exists(_LIB_PATH) or any(
            os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(_LIB_PATH)
            for f in ("stomp_oracle.cpp", "stomp_oracle.hpp", "oracle_capi.cpp", "kinematics_spec.hpp", "Makefile")):
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        dp, ip, u8p, vp = C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_uint8), C.c_void_p
        L.oracle_create.restype = vp
        L.oracle_create.argtypes = [C.POINTER(OracleConfig)]
        L.oracle_destroy.argtypes = [vp]
        L.oracle_set_chain.argtypes = [vp, C.c_int, dp, dp, dp, ip, ip, dp, dp]
        L.oracle_set_spheres.argtypes = [vp, C.c_int, ip, dp, dp]
        L.oracle_set_sdf.argtypes = [vp, ip, dp, C.c_double, C.POINTER(C.c_float)]
        L.oracle_set_self_collision.argtypes = [vp, C.c_int, ip]
        L.oracle_set_cost_extras.argtypes = [vp, C.c_int, C.c_double, C.c_double, C.c_int, dp, dp, C.c_double]
        L.oracle_set_start_goal.argtypes = [vp, dp, dp]
        L.oracle_set_initial_trajectory.argtypes = [vp, dp]
        L.oracle_get_policy.argtypes = [vp, dp, dp, dp, dp, dp, dp]
        L.oracle_get_movement_dt.restype = C.c_double
        L.oracle_get_movement_dt.argtypes = [vp]
        L.oracle_begin_solve.argtypes = [vp]
        L.oracle_set_cholesky.argtypes = [vp, dp]
        L.oracle_iterate.argtypes = [vp, C.c_int, dp, dp]
        L.oracle_finish_solve.argtypes = [vp, dp, ip]
        L.oracle_solve.argtypes = [vp, C.c_int, C.c_int, dp, ip, dp, dp]
        L.oracle_num_rollouts.argtypes = [vp, ip, ip]
        L.oracle_get_rollout_field.argtypes = [vp, C.c_int, dp]
        L.oracle_get_rollout_validity.argtypes = [vp, u8p]
        L.oracle_get_updates.argtypes = [vp, dp]
        L.oracle_get_parameters.argtypes = [vp, dp]
        L.oracle_get_stddevs.argtypes = [vp, dp]
        L.oracle_get_noiseless.argtypes = [vp, dp, ip, dp, dp, dp]
        L.oracle_sincos.argtypes = [C.c_double, dp, dp]
        L.oracle_sincos.restype = None
        L.oracle_sphere_centres.argtypes = [vp, dp, dp]
        L.oracle_state_costs.argtypes = [vp, dp, C.c_int, dp, u8p, u8p, C.c_int]
        L.oracle_control_costs.argtypes = [vp, dp, dp, C.c_int, C.c_double, dp, C.c_int]
        L.oracle_full_piv_lu_inverse.argtypes = [dp, C.c_int, dp]
        L.oracle_llt_lower.argtypes = [dp, C.c_int, dp]
        L.oracle_max_threads.restype = C.c_int
        L.oracle_set_num_threads.restype = C.c_int
        L.oracle_set_num_threads.argtypes = [C.c_int]
        L.oracle_set_sdf_primitives.argtypes = [vp, ip, dp, C.c_double, C.c_int, ip, dp, dp]
        L.oracle_build_sdf_primitives.argtypes = [ip, dp, C.c_double, C.c_int, ip, dp, dp, C.POINTER(C.c_float)]
        L.oracle_build_sdf_occupancy.argtypes = [ip, C.c_double, u8p, C.POINTER(C.c_float)]
        L.oracle_voxelise_scene.argtypes = [ip, dp, C.c_double, C.c_int, dp, C.c_int, C.c_int, dp, dp, u8p, u8p]
        _lib = L
    return _lib


def _dp(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_int32))


def _c64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


FIELDS = {"parameters_noise": 0, "noise": 1, "control_costs": 2, "probabilities": 3, "cumulative_costs": 4,
          "total_costs": 5, "state_costs": 6, "full_probabilities": 7, "full_costs": 8, "total_cost": 9,
          "parameters_noise_projected": 10, "noise_projected": 11}


class Oracle:
    """The reference's StompPlanner / Stomp / PolicyImprovement path on the CPU."""

    def __init__(self, *, num_time_steps, num_dimensions, min_rollouts, max_rollouts, num_rollouts_per_iteration,
                 num_iterations=30, movement_duration=5.0, control_cost_weight=0.001, min_cost_improvement=0.01,
                 noise_stddev=None, noise_decay=None, noise_min_stddev=None, use_noise_adaptation=True,
                 use_openmp=False, use_cumulative_costs=True, use_projection=False, per_timestep_minmax=False,
                 dense_control_costs=False, seed=42):
        D = num_dimensions
        cfg = OracleConfig()
        cfg.num_time_steps, cfg.num_dimensions = num_time_steps, D
        cfg.min_rollouts, cfg.max_rollouts = min_rollouts, max_rollouts
        cfg.num_rollouts_per_iteration, cfg.num_iterations = num_rollouts_per_iteration, num_iterations
        cfg.movement_duration, cfg.control_cost_weight = movement_duration, control_cost_weight
        cfg.min_cost_improvement = min_cost_improvement
        for i in range(D):
            cfg.noise_stddev[i] = float(noise_stddev[i]) if noise_stddev is not None else 0.1
            cfg.noise_decay[i] = float(noise_decay[i]) if noise_decay is not None else 1.0
            cfg.noise_min_stddev[i] = float(noise_min_stddev[i]) if noise_min_stddev is not None else 0.01
        cfg.use_noise_adaptation, cfg.use_openmp = int(use_noise_adaptation), int(use_openmp)
        cfg.use_cumulative_costs, cfg.use_projection = int(use_cumulative_costs), int(use_projection)
        cfg.per_timestep_minmax, cfg.dense_control_costs = int(per_timestep_minmax), int(dense_control_costs)
        cfg.seed = seed
        self.cfg = cfg
        self.T, self.D, self.N = num_time_steps, D, num_time_steps + 12
        self.h = lib().oracle_create(C.byref(cfg))
        if not self.h:
            raise ValueError("oracle_create rejected the configuration")
        self._keep = []
        self.S = 0

    def __del__(self):
        if getattr(self, "h", None):
            lib().oracle_destroy(self.h)
            self.h = None

    # -- scene ------------------------------------------------------------------------------------
    def set_chain(self, chain):
        a = [_c64(chain.origin_xyz), _c64(chain.origin_rpy), _c64(chain.axis)]
        par = np.ascontiguousarray(chain.parent, dtype=np.int32)
        pri = np.ascontiguousarray(chain.prismatic, dtype=np.int32)
        lo, up = _c64(chain.lower), _c64(chain.upper)
        rc = lib().oracle_set_chain(self.h, self.D, _dp(a[0]), _dp(a[1]), _dp(a[2]), _ip(par), _ip(pri), _dp(lo), _dp(up))
        assert rc == 0, rc

    def set_spheres(self, spheres):
        link = np.ascontiguousarray(spheres.link, dtype=np.int32)
        xyz, rad = _c64(spheres.xyz), _c64(spheres.radius)
        rc = lib().oracle_set_spheres(self.h, len(link), _ip(link), _dp(xyz), _dp(rad))
        assert rc == 0, rc
        self.S = len(link)

    def set_self_collision(self, pairs):
        """pairs: [n][2] sphere indices checked against each other; empty switches the check off"""
        pr = np.ascontiguousarray(pairs, dtype=np.int32).reshape(-1, 2)
        rc = lib().oracle_set_self_collision(self.h, len(pr), _ip(pr))
        assert rc == 0, rc

    def set_cost_extras(self, smooth=None, joint_constraint=None):
        """smooth: (margin, weight) or None; joint_constraint: (value [D], tolerance [D], weight) or None"""
        sm = smooth or (0.0, 1.0)
        if joint_constraint is not None:
            v, tol, w = _c64(joint_constraint[0]), _c64(joint_constraint[1]), float(joint_constraint[2])
            rc = lib().oracle_set_cost_extras(self.h, int(smooth is not None), float(sm[0]), float(sm[1]), 1, _dp(v), _dp(tol), w)
        else:
            rc = lib().oracle_set_cost_extras(self.h, int(smooth is not None), float(sm[0]), float(sm[1]), 0, None, None, 1.0)
        assert rc == 0, rc

    def set_sdf(self, sdf, analytic=None):
        """A host grid is used as it is.  A lazy Sdf (grid None, primitives listed) is either built here (OpenMP) or, for
        `analytic` (default: grids of >= 2^28 voxels), evaluated from the primitives at the voxel each lookup hits."""
        if sdf.grid is None:
            kind, centre, size = sdf.primitive_arrays()
            dims = np.ascontiguousarray(sdf.dims, dtype=np.int32)
            org = _c64(sdf.origin)
            count = int(dims[0]) * int(dims[1]) * int(dims[2])
            if analytic is None:
                analytic = count >= (1 << 28)
            if analytic:
                self._keep.append((kind, centre, size))
                rc = lib().oracle_set_sdf_primitives(self.h, _ip(dims), _dp(org), float(sdf.voxel), len(kind), _ip(kind),
                                                     _dp(centre), _dp(size))
                assert rc == 0, rc
                return
            sdf = type(sdf)(dims=sdf.dims, origin=sdf.origin, voxel=sdf.voxel,
                            grid=build_sdf_primitives(dims, org, sdf.voxel, kind, centre, size), obstacles=sdf.obstacles)
        grid = np.ascontiguousarray(sdf.grid, dtype=np.float32)
        dims = np.ascontiguousarray(sdf.dims, dtype=np.int32)
        org = _c64(sdf.origin)
        self._keep.append(grid)   # the oracle does not copy the grid
        rc = lib().oracle_set_sdf(self.h, _ip(dims), _dp(org), float(sdf.voxel), grid.ctypes.data_as(C.POINTER(C.c_float)))
        assert rc == 0, rc

    def set_problem(self, problem, query=None):
        self.set_chain(problem.chain)
        self.set_spheres(problem.spheres)
        self.set_sdf(problem.sdf)
        s, g = problem.start, problem.goal
        if s.ndim == 2:
            s, g = s[query or 0], g[query or 0]
        self.set_start_goal(s, g)

    def set_start_goal(self, start, goal):
        s, g = _c64(start), _c64(goal)
        assert lib().oracle_set_start_goal(self.h, _dp(s), _dp(g)) == 0

    def set_initial_trajectory(self, traj):
        t = _c64(traj)
        assert t.shape == (self.D, self.T)
        assert lib().oracle_set_initial_trajectory(self.h, _dp(t)) == 0

    def policy(self):
        T, D, N = self.T, self.D, self.N
        out = dict(R=np.empty((T, T)), Rinv=np.empty((T, T)), L=np.empty((T, T)), params_all=np.empty((D, N)),
                   mincc=np.empty((D, T)), linear=np.empty((D, T)))
        rc = lib().oracle_get_policy(self.h, _dp(out["R"]), _dp(out["Rinv"]), _dp(out["L"]), _dp(out["params_all"]),
                                     _dp(out["mincc"]), _dp(out["linear"]))
        assert rc == 0
        out["dt"] = lib().oracle_get_movement_dt(self.h)
        return out

    # -- loop -------------------------------------------------------------------------------------
    def begin_solve(self):
        assert lib().oracle_begin_solve(self.h) == 0

    def set_cholesky(self, L):
        Lc = _c64(L)
        assert lib().oracle_set_cholesky(self.h, _dp(Lc)) == 0

    def iterate(self, iteration, noise=None, epsilon=None):
        n = None if noise is None else _c64(noise)
        e = None if epsilon is None else _c64(epsilon)
        rc = lib().oracle_iterate(self.h, iteration, _dp(n), _dp(e))
        assert rc >= 0
        return bool(rc)

    def finish_solve(self):
        sol = np.empty((self.D, self.T))
        it = C.c_int32(0)
        ok = lib().oracle_finish_solve(self.h, _dp(sol), C.byref(it))
        return bool(ok), sol, it.value

    def solve(self, max_iterations, honour_stop=True):
        sol = np.empty((self.D, self.T))
        it = C.c_int32(0)
        sec, setup = C.c_double(0), C.c_double(0)
        ok = lib().oracle_solve(self.h, max_iterations, int(honour_stop), _dp(sol), C.byref(it), C.byref(sec), C.byref(setup))
        assert ok >= 0
        return dict(found=bool(ok), solution=sol, iterations=it.value, seconds=sec.value, setup_seconds=setup.value)

    def num_rollouts(self):
        a, b = C.c_int32(0), C.c_int32(0)
        assert lib().oracle_num_rollouts(self.h, C.byref(a), C.byref(b)) == 0
        return a.value, b.value

    def field(self, name):
        n, _ = self.num_rollouts()
        fid = FIELDS[name]
        shape = {6: (n, self.T), 7: (n, self.D), 8: (n, self.D), 9: (n,)}.get(fid, (n, self.D, self.T))
        out = np.empty(shape)
        assert lib().oracle_get_rollout_field(self.h, fid, _dp(out)) == 0
        return out

    def rollout_validity(self):
        _, g = self.num_rollouts()
        out = np.empty(g, dtype=np.uint8)
        assert lib().oracle_get_rollout_validity(self.h, out.ctypes.data_as(C.POINTER(C.c_uint8))) == 0
        return out

    def updates(self):
        out = np.empty((self.D, self.T))
        assert lib().oracle_get_updates(self.h, _dp(out)) == 0
        return out

    def parameters(self):
        out = np.empty((self.D, self.T))
        assert lib().oracle_get_parameters(self.h, _dp(out)) == 0
        return out

    def stddevs(self):
        out = np.empty(self.D)
        assert lib().oracle_get_stddevs(self.h, _dp(out)) == 0
        return out

    def noiseless(self):
        tc, valid, best = C.c_double(0), C.c_int32(0), C.c_double(0)
        sc, cc = np.empty(self.T), np.empty((self.D, self.T))
        assert lib().oracle_get_noiseless(self.h, C.byref(tc), C.byref(valid), _dp(sc), _dp(cc), C.byref(best)) == 0
        return dict(total_cost=tc.value, valid=bool(valid.value), state_costs=sc, control_costs=cc, best_cost=best.value)

    # -- kernel-level checkers --------------------------------------------------------------------
    def sphere_centres(self, q):
        qc = _c64(q)
        out = np.empty((self.S, 3))
        lib().oracle_sphere_centres(self.h, _dp(qc), _dp(out))
        return out

    def state_costs(self, theta, threads=1):
        th = _c64(theta)
        K = th.shape[0]
        assert th.shape == (K, self.D, self.T)
        costs = np.empty((K, self.T))
        verdict = np.empty((K, self.T), dtype=np.uint8)
        validity = np.empty(K, dtype=np.uint8)
        u8 = C.POINTER(C.c_uint8)
        lib().oracle_state_costs(self.h, _dp(th), K, _dp(costs), verdict.ctypes.data_as(u8), validity.ctypes.data_as(u8), threads)
        return costs, verdict, validity

    def control_costs(self, parameters, noise, weight, dense_form=False):
        p, n = _c64(parameters), _c64(noise)
        K = n.shape[0]
        out = np.empty((K, self.D, self.T))
        assert lib().oracle_control_costs(self.h, _dp(p), _dp(n), K, float(weight), _dp(out), int(dense_form)) == 0
        return out


def build_sdf_primitives(dims, origin, voxel, kind, centre, size):
    """float32 [nz][ny][nx]: exact signed distance of the primitive union at the voxel centres (OpenMP)."""
    dims = np.ascontiguousarray(dims, dtype=np.int32)
    org, kind = _c64(origin), np.ascontiguousarray(kind, dtype=np.int32)
    centre, size = _c64(centre).reshape(-1, 3), _c64(size).reshape(-1, 3)
    out = np.empty((int(dims[2]), int(dims[1]), int(dims[0])), dtype=np.float32)
    rc = lib().oracle_build_sdf_primitives(_ip(dims), _dp(org), float(voxel), len(kind), _ip(kind), _dp(centre), _dp(size),
                                           out.ctypes.data_as(C.POINTER(C.c_float)))
    assert rc == 0
    return out


def voxelise_scene(dims, origin, voxel, triangles=None, solid=False, leaf_centres=None, leaf_sizes=None, occupied=None):
    """Occupancy [nz][ny][nx] (uint8) of a mesh (triangles [n][3][3]), octomap leaves and / or a given occupancy grid: the
    CPU statement of stomp_b200_build_sdf_scene's voxelisation."""
    dims = np.ascontiguousarray(dims, dtype=np.int32)
    org = _c64(origin)
    tri = _c64(triangles).reshape(-1, 9) if triangles is not None else np.zeros((0, 9))
    lc = _c64(leaf_centres).reshape(-1, 3) if leaf_centres is not None else np.zeros((0, 3))
    ls = _c64(leaf_sizes).reshape(-1) if leaf_sizes is not None else np.zeros(0)
    occ_in = np.ascontiguousarray(occupied, dtype=np.uint8) if occupied is not None else None
    out = np.zeros((int(dims[2]), int(dims[1]), int(dims[0])), dtype=np.uint8)
    u8 = C.POINTER(C.c_uint8)
    rc = lib().oracle_voxelise_scene(_ip(dims), _dp(org), float(voxel), len(tri), _dp(tri) if len(tri) else None, int(solid),
                                     len(ls), _dp(lc) if len(ls) else None, _dp(ls) if len(ls) else None,
                                     occ_in.ctypes.data_as(u8) if occ_in is not None else None, out.ctypes.data_as(u8))
    assert rc == 0, rc
    return out


def build_sdf_occupancy(occupied, voxel):
    """float32 [nz][ny][nx]: signed Euclidean distance transform of an occupancy grid (centre to centre, metres)."""
    occ = np.ascontiguousarray(occupied, dtype=np.uint8)
    dims = np.array(occ.shape[::-1], dtype=np.int32)
    out = np.empty(occ.shape, dtype=np.float32)
    rc = lib().oracle_build_sdf_occupancy(_ip(dims), float(voxel), occ.ctypes.data_as(C.POINTER(C.c_uint8)),
                                          out.ctypes.data_as(C.POINTER(C.c_float)))
    assert rc == 0
    return out


def det_sincos(x):
    s, c = C.c_double(0), C.c_double(0)
    lib().oracle_sincos(float(x), C.byref(s), C.byref(c))
    return s.value, c.value


def full_piv_lu_inverse(A):
    a = _c64(A)
    out = np.empty_like(a)
    lib().oracle_full_piv_lu_inverse(_dp(a), a.shape[0], _dp(out))
    return out


def llt_lower(A):
    a = _c64(A)
    out = np.empty_like(a)
    lib().oracle_llt_lower(_dp(a), a.shape[0], _dp(out))
    return out


def max_threads():
    return lib().oracle_max_threads()


def host_cores():
    """Cores this process may run on (the CPU arm uses all of them, whatever OMP_NUM_THREADS the launcher exported)."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def set_num_threads(n):
    """OpenMP thread count of the oracle (and of oracle/_ref, which shares the OpenMP runtime); returns the count in effect."""
    return lib().oracle_set_num_threads(int(n))
