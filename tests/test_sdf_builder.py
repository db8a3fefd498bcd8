"""Environment -> distance field: the CUDA builders behind stomp_b200_build_sdf_primitives /
stomp_b200_build_sdf_occupancy against the oracle's builders (bit for bit), and the oracle's builders against
independent statements of the same fields (NumPy formulas, scipy's exact EDT)."""
import numpy as np
import pytest

from motion_planners_b200 import binding, problems as P
from oracle import binding as ob


def _primitives(seed=5, n=12):
    rng = np.random.default_rng(seed)
    obs = []
    for i in range(n):
        kind = int(rng.integers(0, 2))
        obs.append((kind, rng.uniform(-1.2, 1.2, 3), rng.uniform(0.05, 0.4, 3)))
    return obs


def test_oracle_primitive_field_equals_the_numpy_formulas():
    obs = _primitives()
    ref = P.make_sdf(48, obs, lazy=False)
    lazy = P.make_sdf(48, obs, lazy=True)
    assert lazy.grid is None and len(lazy.obstacles) == len(obs)
    got = ob.build_sdf_primitives(lazy.dims, lazy.origin, lazy.voxel, *lazy.primitive_arrays())
    assert got.dtype == np.float32 and got.shape == (48, 48, 48)
    assert np.array_equal(got.view(np.uint32), ref.grid.view(np.uint32))
    assert (got < 0).any() and (got > 0).any()


def test_oracle_analytic_field_equals_the_built_grid(small_problem):
    """oracle_set_sdf_primitives evaluates the field at the voxel a lookup hits: same verdicts as with the built grid."""
    pb = small_problem
    T, D = pb.num_time_steps, pb.chain.num_dimensions
    lazy = P.Sdf(dims=pb.sdf.dims, origin=pb.sdf.origin, voxel=pb.sdf.voxel, grid=None, obstacles=pb.sdf.obstacles)
    rng = np.random.default_rng(2)
    theta = rng.uniform(pb.chain.lower[None, :, None], pb.chain.upper[None, :, None], (48, D, T))
    out = []
    for sdf, analytic in ((pb.sdf, None), (lazy, True), (lazy, False)):
        o = ob.Oracle(num_time_steps=T, num_dimensions=D, min_rollouts=2, max_rollouts=2, num_rollouts_per_iteration=2,
                      noise_stddev=pb.noise_stddev)
        o.set_chain(pb.chain); o.set_spheres(pb.spheres); o.set_sdf(sdf, analytic=analytic)
        out.append(o.state_costs(theta)[1])
    np.testing.assert_array_equal(out[0], out[1])
    np.testing.assert_array_equal(out[0], out[2])
    assert 0.02 < out[0].mean() < 0.98


def test_oracle_distance_transform_equals_scipy():
    ndimage = pytest.importorskip("scipy.ndimage")
    rng = np.random.default_rng(9)
    occ = np.zeros((20, 28, 36), dtype=np.uint8)        # [nz][ny][nx], ragged on purpose
    occ[4:9, 10:20, 5:12] = 1
    occ[14:18, 2:6, 20:33] = 1
    occ[rng.integers(0, 20, 30), rng.integers(0, 28, 30), rng.integers(0, 36, 30)] = 1
    h = 0.037
    got = ob.build_sdf_occupancy(occ, h)
    ref = (ndimage.distance_transform_edt(occ == 0) - ndimage.distance_transform_edt(occ != 0)) * h
    np.testing.assert_allclose(got, ref.astype(np.float32), rtol=1e-6, atol=1e-7)
    assert np.all((got < 0) == (occ != 0))


@pytest.mark.gpu
@pytest.mark.parametrize("dims", [(64, 64, 64), (40, 56, 72), (130, 3, 1)])
def test_cuda_primitive_field_is_bit_identical_to_the_oracle(dims):
    obs = _primitives(seed=dims[0])
    kind = np.array([o[0] for o in obs], dtype=np.int32)
    centre = np.array([o[1] for o in obs]); size = np.array([o[2] for o in obs])
    origin = np.array([-1.5, -1.4, -1.3]); voxel = 3.0 / max(dims)
    ref = ob.build_sdf_primitives(np.array(dims, dtype=np.int32), origin, voxel, kind, centre, size)
    e = binding.Engine(num_time_steps=10, num_dimensions=7, min_rollouts=2, max_rollouts=2, num_rollouts_per_iteration=2)
    e.build_sdf_primitives(dims, origin, voxel, kind, centre, size)
    got, org, vox = e.get_sdf()
    assert got.shape == ref.shape and vox == voxel and np.array_equal(org, origin)
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
    # an empty world: +inf everywhere, as the oracle says
    e.build_sdf_primitives(dims, origin, voxel, kind[:0], centre[:0], size[:0])
    assert np.all(np.isposinf(e.get_sdf()[0]))
    e.close()


@pytest.mark.gpu
def test_cuda_distance_transform_is_bit_identical_to_the_oracle():
    rng = np.random.default_rng(10)
    occ = (rng.random((33, 47, 70)) < 0.01).astype(np.uint8)
    occ[10:20, 20:30, 30:50] = 1
    h = 0.02
    ref = ob.build_sdf_occupancy(occ, h)
    e = binding.Engine(num_time_steps=10, num_dimensions=7, min_rollouts=2, max_rollouts=2, num_rollouts_per_iteration=2)
    e.build_sdf_occupancy(occ, (-0.7, -0.5, -0.3), h)
    got, org, vox = e.get_sdf()
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
    # the transform feeds the planner like any other field: verdicts against the oracle holding the same grid
    pb = P.single_arm_problem(K=4, T=10, sdf_n=64)
    e.set_chain(pb.chain); e.set_spheres(pb.spheres)
    o = ob.Oracle(num_time_steps=10, num_dimensions=7, min_rollouts=2, max_rollouts=2, num_rollouts_per_iteration=2,
                  noise_stddev=pb.noise_stddev)
    o.set_chain(pb.chain); o.set_spheres(pb.spheres)
    o.set_sdf(P.Sdf(dims=np.array(occ.shape[::-1], dtype=np.int32), origin=np.array([-0.7, -0.5, -0.3]), voxel=h, grid=ref))
    theta = rng.uniform(pb.chain.lower[None, :, None], pb.chain.upper[None, :, None], (32, 7, 10))
    np.testing.assert_array_equal(e.evaluate_states(theta)[1], o.state_costs(theta)[1])
    e.close()


@pytest.mark.gpu
def test_lazy_problem_builds_its_field_on_the_device(medium_problem):
    pb = medium_problem
    lazy = P.Problem(pb.chain, pb.spheres, P.Sdf(pb.sdf.dims, pb.sdf.origin, pb.sdf.voxel, None, pb.sdf.obstacles),
                     pb.start, pb.goal, pb.noise_stddev, pb.num_time_steps, pb.num_rollouts)
    e = binding.engine_for_problem(lazy)
    got, _, _ = e.get_sdf()
    assert np.array_equal(got.view(np.uint32), pb.sdf.grid.view(np.uint32))
    e.close()
