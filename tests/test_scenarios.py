import numpy as np

from multi_agent_pkgs_b200 import scenarios as sc


def test_batch_shapes_and_determinism():
    a = sc.config2_circle(n_swarms=2).make_batch()
    b = sc.config2_circle(n_swarms=2).make_batch()
    assert a.x0.shape == (20, 9) and a.ref.shape == (20, 10, 6) and a.poly_A.shape == (20, 4, 18, 3)
    assert a.all_pos.shape == (20, 11, 3) and a.prev_self_pos.shape == (20, 11, 3)
    for k in ("x0", "ref", "poly_A", "poly_b", "poly_rows", "all_pos"):
        assert np.array_equal(getattr(a, k), getattr(b, k))
    assert (a.nbr_begin[:10] == 0).all() and (a.nbr_end[10:] == 20).all()  # swarms are independent groups


def test_polytopes_contain_their_seed_and_have_reference_shape():
    b = sc.config2_circle().make_batch()
    for i in range(b.n):
        polys = b.polys_of(i)
        assert 1 <= len(polys) <= 4
        A, d = polys[0]
        assert 6 <= len(d) <= 18
        assert np.all(A @ b.x0[i, :3] - d <= 1e-9)  # first cell is seeded at the agent
        assert np.array_equal(A[-6:], np.array([[1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1]], float))
        assert np.all(np.abs(A) == np.round(np.abs(A)))  # small-integer normals (convex_decomp.cpp:341-357)


def test_reference_velocity_points_backwards():
    path = np.array([[0.0, 0, 1], [10.0, 0, 1]])
    ref = sc.sample_path(path, 5.0, 10, 0.1)
    assert np.allclose(np.diff(ref[:, 0]), 0.5) and np.allclose(ref[:-1, 3], -5.0)  # agent_class.cpp:1533-1538


def test_batch_roundtrip(tmp_path):
    b = sc.config1_single_agent().make_batch()
    b.save(tmp_path / "b.npz")
    c = sc.Batch.load(tmp_path / "b.npz")
    assert np.array_equal(b.poly_b, c.poly_b) and c.params["n_hor"] == 10 and c.params["r_x"][0] == 100.0


def test_large_configs_generate():
    s4 = sc.config4_circle256()
    assert s4.n == 256 and np.allclose(np.linalg.norm(s4.state[:, :2], axis=1), 60, atol=1.0)
    s5 = sc.config5_random(n_rob=256, side=50.0)
    d = np.linalg.norm(s5.state[:, None, :2] - s5.state[None, :, :2], axis=-1) + np.eye(256) * 9
    assert d.min() >= 1.0
