"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes binding of oracle/_ref/libstomp_ref.so: the reference's own, unmodified STOMP core
(/root/reference/src/planners/stomp/src/*.cpp) compiled where it lies against the Eigen / Boost stand-ins of
oracle/ref/shim, behind the driver oracle/ref/ref_driver.cpp.  It exists to pin the CPU restatement (oracle/
stomp_oracle.cpp): tests/test_reference_pin.py compares the two field by field, and tests/golden/make_golden.py
writes the vectors that keep the restatement pinned on machines without /root/reference.

The library can only be built where /root/reference exists (`make -C oracle/ref`); `available()` says whether it is
there.  Nothing of the product imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from . import binding as ob

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_ref", "libstomp_ref.so")
REFERENCE_ROOT = "/root/reference/src/planners/stomp"


def build() -> bool:
    """(Re)build the library when the reference sources are present; returns whether the library exists afterwards."""
    if os.path.isdir(REFERENCE_ROOT):
        subprocess.run(["make", "-C", os.path.join(_HERE, "ref"), "-s"], check=True)
    return os.path.exists(_LIB_PATH)


def available() -> bool:
    """make decides whether the library is stale (it links the oracle's sources too); without the reference
    sources — on the GPU box — the prebuilt file is used as it travelled"""
    return build()


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError("oracle/_ref/libstomp_ref.so is missing and /root/reference is not here to build it")
        L = C.CDLL(_LIB_PATH)
        dp, ip, u8p, vp = C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_uint8), C.c_void_p
        L.ref_create.restype = vp
        L.ref_create.argtypes = [vp]
        L.ref_destroy.argtypes = [vp]
        L.ref_set_start_goal.argtypes = [vp, dp, dp]
        L.ref_set_initial_trajectory.argtypes = [vp, dp]
        L.ref_get_policy.argtypes = [vp, dp, dp, dp, dp, dp, dp]
        L.ref_begin_solve.argtypes = [vp]
        L.ref_set_cost_cumulation.argtypes = [vp, C.c_int]
        L.ref_set_projection.argtypes = [vp, C.c_int]
        L.ref_next_num_generated.argtypes = [vp]
        L.ref_iterate.argtypes = [vp, C.c_int, dp, C.c_int]
        L.ref_get_unit_noise.argtypes = [vp, dp]
        L.ref_num_rollouts.argtypes = [vp, ip, ip]
        L.ref_get_rollout_field.argtypes = [vp, C.c_int, dp]
        L.ref_get_rollout_validity.argtypes = [vp, u8p, C.c_int]
        L.ref_get_updates.argtypes = [vp, dp]
        L.ref_get_parameters.argtypes = [vp, dp]
        L.ref_get_stddevs.argtypes = [vp, dp]
        L.ref_get_noiseless.argtypes = [vp, dp, ip, dp, dp, dp]
        L.ref_finish_solve.argtypes = [vp, dp, ip]
        L.ref_solve.argtypes = [vp, C.c_int, C.c_int, dp]
        _lib = L
    return _lib


_dp, _c64 = ob._dp, ob._c64


class Reference:
    """The reference's stomp::Stomp / PolicyImprovement / CovariantMovementPrimitive, configured like `oracle`
    (an oracle.binding.Oracle whose chain, spheres, SDF and joint limits it borrows for the state verdicts)."""

    def __init__(self, oracle: ob.Oracle):
        self.oracle = oracle                       # keeps the scene alive
        self.T, self.D, self.N = oracle.T, oracle.D, oracle.N
        self.h = lib().ref_create(oracle.h)
        self._gen = 0

    def __del__(self):
        if getattr(self, "h", None):
            lib().ref_destroy(self.h)
            self.h = None

    def set_start_goal(self, start, goal):
        s, g = _c64(start), _c64(goal)
        assert lib().ref_set_start_goal(self.h, _dp(s), _dp(g)) == 0

    def set_initial_trajectory(self, traj):
        t = _c64(traj)
        assert t.shape == (self.D, self.T)
        assert lib().ref_set_initial_trajectory(self.h, _dp(t)) == 0

    def policy(self):
        T, D, N = self.T, self.D, self.N
        out = dict(R=np.empty((T, T)), Rinv=np.empty((T, T)), L=np.empty((T, T)), params_all=np.empty((D, N)),
                   mincc=np.empty((D, T)), linear=np.empty((D, T)))
        assert lib().ref_get_policy(self.h, _dp(out["R"]), _dp(out["Rinv"]), _dp(out["L"]), _dp(out["params_all"]),
                                    _dp(out["mincc"]), _dp(out["linear"])) == 0
        return out

    def begin_solve(self):
        assert lib().ref_begin_solve(self.h) == 0

    def set_cost_cumulation(self, use_cumulative_costs: bool):
        """stomp::Stomp::setCostCumulation; call after begin_solve (a new stomp::Stomp is made per solve)."""
        assert lib().ref_set_cost_cumulation(self.h, int(use_cumulative_costs)) == 0

    def set_projection(self, use_projection: bool):
        """Flips PolicyImprovement::use_projection_ (no setter in the reference) and recomputes the projection matrices."""
        assert lib().ref_set_projection(self.h, int(use_projection)) == 0

    def next_num_generated(self):
        return lib().ref_next_num_generated(self.h)

    def iterate(self, iteration, epsilon):
        """epsilon [G][D][T]: the standard normals the reference's generators return, in rollout-major layout.
        Returns (stop, unit) with unit = L * eps exactly as MultivariateGaussian::sample formed it."""
        e = _c64(epsilon)
        G = e.shape[0]
        assert e.shape == (G, self.D, self.T)
        rc = lib().ref_iterate(self.h, iteration, _dp(e), G)
        assert rc >= 0, f"ref_iterate: {rc} (-3: the reference drew a different number of normals)"
        unit = np.empty((G, self.D, self.T))
        lib().ref_get_unit_noise(self.h, _dp(unit))
        self._gen = G
        return bool(rc), unit

    def num_rollouts(self):
        a, b = C.c_int32(0), C.c_int32(0)
        assert lib().ref_num_rollouts(self.h, C.byref(a), C.byref(b)) == 0
        return a.value, b.value

    def field(self, name):
        n, _ = self.num_rollouts()
        fid = ob.FIELDS[name]
        shape = {6: (n, self.T), 7: (n, self.D), 8: (n, self.D), 9: (n,)}.get(fid, (n, self.D, self.T))
        out = np.empty(shape)
        assert lib().ref_get_rollout_field(self.h, fid, _dp(out)) == 0
        return out

    def rollout_validity(self):
        out = np.empty(self._gen, dtype=np.uint8)
        lib().ref_get_rollout_validity(self.h, out.ctypes.data_as(C.POINTER(C.c_uint8)), self._gen)
        return out

    def updates(self):
        out = np.empty((self.D, self.T))
        assert lib().ref_get_updates(self.h, _dp(out)) == 0
        return out

    def parameters(self):
        out = np.empty((self.D, self.T))
        assert lib().ref_get_parameters(self.h, _dp(out)) == 0
        return out

    def stddevs(self):
        out = np.empty(self.D)
        assert lib().ref_get_stddevs(self.h, _dp(out)) == 0
        return out

    def noiseless(self):
        tc, valid, best = C.c_double(0), C.c_int32(0), C.c_double(0)
        sc, cc = np.empty(self.T), np.empty((self.D, self.T))
        assert lib().ref_get_noiseless(self.h, C.byref(tc), C.byref(valid), _dp(sc), _dp(cc), C.byref(best)) == 0
        return dict(total_cost=tc.value, valid=bool(valid.value), state_costs=sc, control_costs=cc, best_cost=best.value)

    def finish_solve(self):
        sol, it = np.empty((self.D, self.T)), C.c_int32(0)
        status = lib().ref_finish_solve(self.h, _dp(sol), C.byref(it))
        return dict(solution=sol, iterations=it.value, found=bool(status))

    def solve(self, iterations, honour_stop=False):
        sec = C.c_double(0)
        n = lib().ref_solve(self.h, iterations, int(honour_stop), C.byref(sec))
        return dict(iterations=n, seconds=sec.value)
