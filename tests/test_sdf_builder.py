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


# ---- meshes, octomap leaves, cylinders ------------------------------------------------------------------------------
def _icosphere(radius, centre, subdivisions=2):
    """Closed triangle mesh of a sphere [n][3][3] (subdivided icosahedron)."""
    t = (1.0 + 5.0 ** 0.5) / 2.0
    v = np.array([[-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0], [0, -1, t], [0, 1, t], [0, -1, -t], [0, 1, -t],
                  [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]], dtype=np.float64)
    f = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4), (11, 10, 2), (10, 7, 6), (7, 1, 8),
         (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5), (2, 4, 11), (6, 2, 10), (8, 6, 7), (9, 8, 1)]
    tris = np.array([[v[a], v[b], v[c]] for a, b, c in f])
    for _ in range(subdivisions):
        a, b, c = tris[:, 0], tris[:, 1], tris[:, 2]
        ab, bc, ca = (a + b) / 2, (b + c) / 2, (c + a) / 2
        tris = np.concatenate([np.stack([a, ab, ca], 1), np.stack([b, bc, ab], 1), np.stack([c, ca, bc], 1), np.stack([ab, bc, ca], 1)])
    tris = tris / np.linalg.norm(tris, axis=2, keepdims=True) * radius
    return tris + np.asarray(centre)


def _box_mesh(lo, hi):
    lo, hi = np.asarray(lo, dtype=np.float64), np.asarray(hi, dtype=np.float64)
    c = np.array([[lo[0] if i & 1 == 0 else hi[0], lo[1] if i & 2 == 0 else hi[1], lo[2] if i & 4 == 0 else hi[2]] for i in range(8)])
    quads = [(0, 1, 3, 2), (4, 6, 7, 5), (0, 4, 5, 1), (2, 3, 7, 6), (0, 2, 6, 4), (1, 5, 7, 3)]
    return np.array([[c[a], c[b], c[cc]] for a, b, cc, d in quads] + [[c[a], c[cc], c[d]] for a, b, cc, d in quads])


def test_oracle_cylinder_field_equals_the_numpy_formula():
    dims = np.array([40, 36, 44], dtype=np.int32); origin = np.array([-1.0, -0.9, -1.1]); h = 0.05
    kind = np.array([2, 0], dtype=np.int32)
    centre = np.array([[0.1, -0.05, 0.2], [-0.6, 0.5, -0.4]]); size = np.array([[0.3, 0.45, 0.0], [0.2, 0.0, 0.0]])
    got = ob.build_sdf_primitives(dims, origin, h, kind, centre, size)
    z, y, x = np.meshgrid(*(origin[2 - a] + (np.arange(dims[2 - a]) + 0.5) * h for a in range(3)), indexing="ij")
    dx, dy, dz = x - 0.1, y + 0.05, z - 0.2
    qr, qh = np.sqrt(dx * dx + dy * dy) - 0.3, np.abs(dz) - 0.45
    cyl = np.sqrt(np.maximum(qr, 0) ** 2 + np.maximum(qh, 0) ** 2) + np.minimum(np.maximum(qr, qh), 0)
    sph = np.sqrt((x + 0.6) ** 2 + (y - 0.5) ** 2 + (z + 0.4) ** 2) - 0.2
    np.testing.assert_allclose(got, np.minimum(cyl, sph).astype(np.float32), rtol=1e-6, atol=1e-7)
    assert (got < 0).any()


def test_oracle_mesh_voxelisation_of_a_sphere_and_a_box():
    """Conservative surface voxelisation + interior fill: the occupied set of a closed mesh contains every voxel whose
    centre is inside the solid and nothing farther than one voxel diagonal outside it."""
    dims = np.array([48, 40, 44], dtype=np.int32); origin = np.array([-0.6, -0.5, -0.55]); h = 0.025
    tris = _icosphere(0.31, (0.02, -0.01, 0.03), subdivisions=3)
    shell = ob.voxelise_scene(dims, origin, h, triangles=tris, solid=False)
    solid = ob.voxelise_scene(dims, origin, h, triangles=tris, solid=True)
    z, y, x = np.meshgrid(*(origin[2 - a] + (np.arange(dims[2 - a]) + 0.5) * h for a in range(3)), indexing="ij")
    r = np.sqrt((x - 0.02) ** 2 + (y + 0.01) ** 2 + (z - 0.03) ** 2)
    assert shell.sum() < solid.sum()
    assert np.all(solid[r < 0.31 - 0.005] == 1)                      # inside the solid (the icosphere is inscribed: small slack)
    assert np.all(solid[r > 0.31 + h * 3 ** 0.5 * 0.5 + 1e-9] == 0)   # no voxel farther than half a diagonal from the surface
    assert np.all(shell[np.abs(r - 0.31) > h * 3 ** 0.5 * 0.5 + 0.006] == 0)
    # a box whose faces lie exactly on voxel boundaries: the closed-set rule also takes the voxels that only touch it
    bx = ob.voxelise_scene(dims, origin, h, triangles=_box_mesh((-0.1, -0.2, -0.05), (0.15, 0.1, 0.2)), solid=True)
    inside = (x > -0.1) & (x < 0.15) & (y > -0.2) & (y < 0.1) & (z > -0.05) & (z < 0.2)
    assert np.all(bx[inside] == 1)
    grown = (x > -0.1 - h) & (x < 0.15 + h) & (y > -0.2 - h) & (y < 0.1 + h) & (z > -0.05 - h) & (z < 0.2 + h)
    assert np.all(bx[~grown] == 0)
    # octomap leaves: cubes whose centres' voxels are exactly those with centre inside
    lv = ob.voxelise_scene(dims, origin, h, leaf_centres=[[0.0, 0.0, 0.0], [0.3, 0.2, -0.3]], leaf_sizes=[0.1, 0.2])
    ref = ((np.abs(x) <= 0.05) & (np.abs(y) <= 0.05) & (np.abs(z) <= 0.05)) | ((np.abs(x - 0.3) <= 0.1) & (np.abs(y - 0.2) <= 0.1) & (np.abs(z + 0.3) <= 0.1))
    assert lv.sum() > 0 and np.abs(lv.astype(int) - ref.astype(int)).sum() <= 0.02 * ref.sum()    # boundary ties only


@pytest.mark.gpu
@pytest.mark.parametrize("solid", [False, True])
def test_cuda_scene_builder_is_bit_identical_to_the_oracle(solid):
    """stomp_b200_build_sdf_scene: mesh + octomap leaves + occupancy, voxelised and transformed on the device, against the
    oracle's voxelisation followed by its distance transform."""
    rng = np.random.default_rng(12)
    dims = np.array([56, 44, 50], dtype=np.int32); origin = np.array([-0.7, -0.55, -0.6]); h = 0.025
    tris = np.concatenate([_icosphere(0.22, (0.1, -0.05, 0.02), 2), _box_mesh((-0.5, -0.4, -0.45), (-0.2, -0.15, 0.1)),
                           _icosphere(0.3, (0.6, 0.5, 0.55), 1)])          # the last one sticks out of the grid
    leaves_c = rng.uniform(-0.6, 0.6, (40, 3)); leaves_s = rng.choice([0.025, 0.05, 0.1], 40)
    occ_in = (rng.random((50, 44, 56)) < 0.002).astype(np.uint8)
    occ_ref = ob.voxelise_scene(dims, origin, h, triangles=tris, solid=solid, leaf_centres=leaves_c, leaf_sizes=leaves_s, occupied=occ_in)
    ref = ob.build_sdf_occupancy(occ_ref, h)
    e = binding.Engine(num_time_steps=10, num_dimensions=7, min_rollouts=2, max_rollouts=2, num_rollouts_per_iteration=2)
    e.build_sdf_scene(dims, origin, h, triangles=tris, solid=solid, leaf_centres=leaves_c, leaf_sizes=leaves_s, occupied=occ_in)
    got, org, vox = e.get_sdf()
    assert np.array_equal((got < 0), occ_ref != 0)
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
    # each part on its own
    for kw in (dict(triangles=tris, solid=solid), dict(leaf_centres=leaves_c, leaf_sizes=leaves_s), dict(occupied=occ_in)):
        e.build_sdf_scene(dims, origin, h, **kw)
        assert np.array_equal(e.get_sdf()[0] < 0, ob.voxelise_scene(dims, origin, h, **kw) != 0)
    e.close()


@pytest.mark.gpu
def test_cuda_cylinder_primitive_is_bit_identical_to_the_oracle():
    dims = np.array([40, 36, 44], dtype=np.int32); origin = np.array([-1.0, -0.9, -1.1]); h = 0.05
    kind = np.array([2, 1, 0, 2], dtype=np.int32)
    centre = np.array([[0.1, -0.05, 0.2], [-0.5, 0.4, 0.1], [0.5, 0.5, -0.5], [0.0, 0.3, -0.6]])
    size = np.array([[0.3, 0.45, 0.0], [0.2, 0.1, 0.3], [0.25, 0, 0], [0.1, 0.2, 0.0]])
    ref = ob.build_sdf_primitives(dims, origin, h, kind, centre, size)
    e = binding.Engine(num_time_steps=10, num_dimensions=7, min_rollouts=2, max_rollouts=2, num_rollouts_per_iteration=2)
    e.build_sdf_primitives(dims, origin, h, kind, centre, size)
    assert np.array_equal(e.get_sdf()[0].view(np.uint32), ref.view(np.uint32))
    e.close()
