"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.

Bars (BASELINE.json north_star): correspondence index sets bit-exact (exact NN, (d2_f32, index) ordering, same FP32
distance arithmetic); poses within 1e-5 rad / 1e-4 m; f64 per-point quantities (normals, label vectors) bit-exact or
<= 1e-12; reduced quantities (cost, gradient, J^T J) <= 1e-9 relative.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROT_TOL, TRANS_TOL = 1e-5, 1e-4  # rad, m — north_star pose tolerance


@pytest.fixture(scope="module")
def room(pkg):
    return pkg.synth.room_pair(seed=100, n_points=10_000)


@pytest.fixture(scope="module")
def kitti_small(pkg):
    return pkg.synth.kitti_pair(pair=1, n_points=20_000, n_rings=32, n_az=700)


# ------------------------------------------------------------------------------------------------ kNN
@pytest.mark.parametrize("k", [1, 4, 20])
def test_knn_bit_exact_room(sicp, oracle, room, k):
    tgt = sicp.Cloud(room["tgt_xyz"])
    idx, d2 = sicp.knn(tgt, room["src_xyz"], k)
    ridx, rd2 = oracle.knn(room["tgt_xyz"], room["src_xyz"], k)
    assert np.array_equal(idx, ridx)
    assert np.array_equal(d2, rd2)


@pytest.mark.parametrize("k", [1, 4, 20, 7, 32])
def test_knn_bit_exact_lidar_with_pose(sicp, oracle, kitti_small, k):
    p = kitti_small
    tgt = sicp.Cloud(p["tgt_xyz"])
    idx, d2 = sicp.knn(tgt, p["src_xyz"], k, pose7=p["T_gt"])
    q = oracle.transform_points(p["T_gt"], p["src_xyz"])
    ridx, rd2 = oracle.knn(p["tgt_xyz"], q, k)
    assert np.array_equal(idx, ridx)
    assert np.array_equal(d2, rd2)


def test_knn_vs_bruteforce_definition(sicp, oracle, kitti_small):
    p = kitti_small
    tgt = sicp.Cloud(p["tgt_xyz"])
    q = p["src_xyz"][:777]
    idx, d2 = sicp.knn(tgt, q, 4)
    ridx, rd2 = oracle.knn(p["tgt_xyz"], q, 4, brute=True)
    assert np.array_equal(idx, ridx) and np.array_equal(d2, rd2)


def test_knn_ties_lowest_index(sicp, oracle):
    # integer lattice => many exactly equal distances; duplicated points => zero-distance ties
    g = np.stack(np.meshgrid(np.arange(12), np.arange(12), np.arange(6), indexing="ij"), -1).reshape(-1, 3).astype(np.float32)
    rng = np.random.default_rng(3)
    tgt_pts = np.concatenate([g, g[rng.integers(0, len(g), 200)]])
    rng.shuffle(tgt_pts)
    q = rng.integers(0, 12, size=(500, 3)).astype(np.float32) + 0.5 * rng.integers(0, 2, size=(500, 3)).astype(np.float32)
    tgt = sicp.Cloud(tgt_pts)
    for k in (1, 4, 20):
        idx, d2 = sicp.knn(tgt, q, k)
        ridx, rd2 = oracle.knn(tgt_pts, q, k, brute=True)
        assert np.array_equal(d2, rd2)
        assert np.array_equal(idx, ridx)


@pytest.mark.parametrize("n", [1, 3, 4, 5, 31, 32, 33, 257])
def test_knn_tiny_and_ragged_targets(sicp, oracle, n):
    rng = np.random.default_rng(n)
    tgt_pts = rng.normal(size=(n, 3)).astype(np.float32)
    q = rng.normal(size=(70, 3)).astype(np.float32)
    tgt = sicp.Cloud(tgt_pts)
    for k in (1, 4, 20):
        idx, d2 = sicp.knn(tgt, q, k)
        ridx, rd2 = oracle.knn(tgt_pts, q, k, brute=True)
        assert np.array_equal(idx, ridx)
        assert np.array_equal(d2[ridx >= 0], rd2[ridx >= 0])
        assert np.all(np.isinf(d2[ridx < 0]))


def test_knn_per_class(sicp, oracle, room):
    p = room
    tgt = sicp.Cloud(p["tgt_xyz"], p["tgt_labels"], layout=sicp.CLOUD_PER_CLASS)
    idx, d2 = sicp.knn(tgt, p["src_xyz"], 1, q_labels=p["src_labels"])
    for lab in np.unique(p["src_labels"]):
        qs = np.nonzero(p["src_labels"] == lab)[0]
        ts = np.nonzero(p["tgt_labels"] == lab)[0]
        if len(ts) == 0:
            assert np.all(idx[qs] == -1)
            continue
        ridx, rd2 = oracle.knn(p["tgt_xyz"][ts], p["src_xyz"][qs], 1)
        assert np.array_equal(idx[qs, 0], ts[ridx[:, 0]])
        assert np.array_equal(d2[qs], rd2)


def test_class_order_first_appearance(sicp, oracle, room):
    c = sicp.Cloud(room["src_xyz"], room["src_labels"], layout=sicp.CLOUD_PER_CLASS)
    labs, sizes = c.classes()
    rl, rs, _ = oracle.label_split(room["src_labels"])
    assert np.array_equal(labs, rl)
    assert np.array_equal(sizes, np.diff(rs))


# ------------------------------------------------------------------------------------------------ covariances / E-step inputs
def test_covariance_neighbours_normals_label_vectors(sicp, oracle, kitti_small):
    p = kitti_small
    c = sicp.Cloud(p["tgt_xyz"], p["tgt_labels"])
    c.precompute(20, 1e-3, p["cm"])
    ref = oracle.covariances(p["tgt_xyz"], 20, 1e-3, labels=p["tgt_labels"], N=p["N"], want_nn=True)
    assert np.array_equal(c.self_neighbours(), ref["nn"])          # neighbour sets, in order: bit-exact
    nrm = c.normals()
    assert np.array_equal(nrm, ref["normals"])                      # same Jacobi-SVD operation sequence: bit-exact
    cov = c.covariances()
    assert np.max(np.abs(cov - ref["cov"])) <= 1e-12               # I-(1-eps)nn^T vs sum_k v_k u_k u_k^T
    assert np.array_equal(c.label_distributions(), ref["dist"])    # repeated 1/k additions: bit-exact
    a_ref = ref["dist"] @ p["cm"]
    assert np.max(np.abs(c.label_vectors() - a_ref)) <= 1e-14


def test_self_neighbours_with_exact_ties(sicp, oracle):
    # integer lattice + duplicated points: whole shells of exactly equal distances straddle the 20th neighbour, so the
    # covariance search has to break ties by original index
    g = np.stack(np.meshgrid(np.arange(10), np.arange(10), np.arange(5), indexing="ij"), -1).reshape(-1, 3).astype(np.float32)
    rng = np.random.default_rng(11)
    pts = np.concatenate([g, g[rng.integers(0, len(g), 150)]])
    rng.shuffle(pts)
    for k in (20, 12):
        c = sicp.Cloud(pts)
        c.precompute(k, 1e-3)
        ridx, _ = oracle.knn(pts, pts, k, brute=True)
        assert np.array_equal(c.self_neighbours(), ridx)
    flat = np.concatenate([g[:, :2], np.zeros((len(g), 1), np.float32)], 1)  # every point 5 times: ties of 5, 20, 45 ... candidates
    c = sicp.Cloud(flat)
    c.precompute(20, 1e-3)
    ridx, _ = oracle.knn(flat, flat, 20, brute=True)
    assert np.array_equal(c.self_neighbours(), ridx)


def test_covariance_per_class(sicp, oracle, room):
    p = room
    c = sicp.Cloud(p["src_xyz"], p["src_labels"], layout=sicp.CLOUD_PER_CLASS)
    c.precompute(20, 1e-3)
    ref = oracle.covariances_per_class(p["src_xyz"], p["src_labels"], 20, 1e-3)
    assert np.array_equal(c.normals(), ref["normals"])
    assert np.max(np.abs(c.covariances() - ref["cov"])) <= 1e-12


def test_covariance_fewer_points_than_k(sicp, oracle):
    rng = np.random.default_rng(5)
    xyz = rng.normal(size=(13, 3)).astype(np.float32)  # divisor stays k=20 (semantic_point_cloud.hpp:60-64)
    c = sicp.Cloud(xyz)
    c.precompute(20, 1e-3)
    ref = oracle.covariances(xyz, 20, 1e-3)
    assert np.array_equal(c.normals(), ref["normals"])


# ------------------------------------------------------------------------------------------------ one pass
def _algo_setup(sicp, p, algo):
    if algo == "gicp":
        return sicp.ALGO_GICP, sicp.Cloud(p["src_xyz"]), sicp.Cloud(p["tgt_xyz"]), sicp.default_options(sicp.ALGO_GICP)
    if algo == "em":
        return (sicp.ALGO_EM, sicp.Cloud(p["src_xyz"], p["src_labels"]), sicp.Cloud(p["tgt_xyz"], p["tgt_labels"]),
                sicp.default_options(sicp.ALGO_EM, cm=p["cm"]))
    return (sicp.ALGO_SEMANTIC, sicp.Cloud(p["src_xyz"], p["src_labels"], layout=sicp.CLOUD_PER_CLASS),
            sicp.Cloud(p["tgt_xyz"], p["tgt_labels"], layout=sicp.CLOUD_PER_CLASS), sicp.default_options(sicp.ALGO_SEMANTIC))


def _oracle_align(oracle, p, algo):
    if algo == "gicp":
        return oracle.align_gicp(p["src_xyz"], p["tgt_xyz"], p["init"])
    if algo == "em":
        return oracle.align_em(p["src_xyz"], p["src_labels"], p["tgt_xyz"], p["tgt_labels"], p["cm"], p["init"])
    return oracle.align_semantic(p["src_xyz"], p["src_labels"], p["tgt_xyz"], p["tgt_labels"], p["init"])


@pytest.mark.parametrize("algo", ["gicp", "em", "semantic"])
def test_first_pass_correspondences_and_weights(sicp, oracle, room, algo):
    a, src, tgt, opts = _algo_setup(sicp, room, algo)
    idx, w, d2 = sicp.correspondences(a, src, tgt, opts, room["init"])
    ref = _oracle_align(oracle, room, algo)
    assert np.array_equal(idx, ref["corr0"])                        # bit-exact index sets (incl. gate / class rules)
    m = ref["corr0"] >= 0
    assert np.array_equal(d2[m], ref["d20"][m])
    assert np.max(np.abs(w - ref["w0"])) <= 1e-12 * max(1.0, np.max(np.abs(ref["w0"])))


@pytest.mark.parametrize("algo,loss", [("gicp", 0), ("em", 2), ("semantic", 1)])
def test_evaluate_cost_gradient_hessian(sicp, oracle, room, algo, loss):
    p = room
    a, src, tgt, opts = _algo_setup(sicp, p, algo)
    idx, w, _ = sicp.correspondences(a, src, tgt, opts, p["init"])
    if algo == "semantic":  # SemanticPointCloud covariances: neighbours restricted to the point's class (semantic_point_cloud.hpp:25-84)
        scov = oracle.covariances_per_class(p["src_xyz"], p["src_labels"], 20, 1e-3)["cov"]
        tcov = oracle.covariances_per_class(p["tgt_xyz"], p["tgt_labels"], 20, 1e-3)["cov"]
    else:
        scov = oracle.covariances(p["src_xyz"], 20, 1e-3)["cov"]
        tcov = oracle.covariances(p["tgt_xyz"], 20, 1e-3)["cov"]
    kc = idx.shape[1]
    s_idx = np.repeat(np.arange(src.n), kc)[idx.ravel() >= 0]
    t_idx = idx.ravel()[idx.ravel() >= 0]
    ww = w.ravel()[idx.ravel() >= 0]
    rng = np.random.default_rng(1)
    for trial in range(3):
        x = oracle.se3_exp(rng.normal(scale=0.02, size=6)) if trial else p["init"]
        cost, g, H = sicp.evaluate(a, src, tgt, opts, p["init"], x)
        rc, rg, rH = oracle.eval_problem(p["src_xyz"], scov, p["tgt_xyz"], tcov, s_idx, t_idx, ww, loss, x)
        assert abs(cost - rc) <= 1e-9 * abs(rc)
        assert np.max(np.abs(g - rg)) <= 1e-9 * np.max(np.abs(rg))
        assert np.max(np.abs(H - rH)) <= 1e-9 * np.max(np.abs(rH))


# ------------------------------------------------------------------------------------------------ full registrations
@pytest.mark.parametrize("algo", ["gicp", "em", "semantic"])
def test_register_pose_parity_room(sicp, oracle, pkg, room, algo):
    a, src, tgt, opts = _algo_setup(sicp, room, algo)
    res = sicp.register(a, src, tgt, opts, room["init"])
    ref = _oracle_align(oracle, room, algo)
    rot, trans = pkg.synth.pose_error(res["pose"], ref["pose"])
    assert rot < ROT_TOL and trans < TRANS_TOL, (rot, trans)
    assert res["outer_iter"] == ref["outer_iter"]
    n = ref["outer_iter"]
    for i in range(n):  # every intermediate pose agrees, so the discrete correspondence sequence is the same
        r, t = pkg.synth.pose_error(res["pass_pose"][i], ref["pass_pose"][i])
        assert r < ROT_TOL and t < TRANS_TOL, (i, r, t)
    assert list(res["pass_lm_iters"]) == list(ref["pass_lm_iters"])


@pytest.mark.parametrize("algo", ["gicp", "em"])
def test_register_pose_parity_lidar(sicp, oracle, pkg, kitti_small, algo):
    p = kitti_small
    a, src, tgt, opts = _algo_setup(sicp, p, algo)
    res = sicp.register(a, src, tgt, opts, p["init"])
    ref = _oracle_align(oracle, p, algo)
    rot, trans = pkg.synth.pose_error(res["pose"], ref["pose"])
    assert rot < ROT_TOL and trans < TRANS_TOL, (rot, trans)
    assert res["outer_iter"] == ref["outer_iter"]


def test_register_batch_matches_single(sicp, pkg, room):
    a, src, tgt, opts = _algo_setup(sicp, room, "em")
    single = sicp.register(a, src, tgt, opts, room["init"])
    inits = np.stack([room["init"]] * 5)
    batch = sicp.register_batch(a, [src] * 5, [tgt] * 5, opts, inits)
    for b in batch:
        assert np.array_equal(b["pose"], batch[0]["pose"])  # fixed-order reductions: bit-identical within a batch
        # a batch runs its solves on half-size grids (two per SM), so the summation grouping differs from a single run
        rot, trans = pkg.synth.pose_error(b["pose"], single["pose"])
        assert rot < 1e-9 and trans < 1e-9
        assert b["outer_iter"] == single["outer_iter"]
    again = sicp.register_batch(a, [src] * 5, [tgt] * 5, opts, inits)
    assert all(np.array_equal(x["pose"], y["pose"]) for x, y in zip(batch, again))  # run-to-run deterministic


def test_fused_labels(sicp, oracle, room):
    p = room
    a, src, tgt, opts = _algo_setup(sicp, p, "em")
    got = sicp.fused_labels(src, tgt, opts, p["T_gt"])
    ref = oracle.fused_labels(p["src_xyz"], p["src_labels"], p["tgt_xyz"], p["tgt_labels"], p["cm"], p["T_gt"])
    assert np.mean(got == ref) > 0.9995
    # every mismatch must be an exact tie of the arg-max (f64 sums whose order differs by <= 1e-15): recompute the class
    # scores of the mismatching points from the GPU's own label vectors and 4-NN lists
    mism = np.nonzero(got != ref)[0]
    if len(mism):
        idx, d2 = sicp.knn(tgt, p["src_xyz"], 4, pose7=p["T_gt"])
        a_s, a_t = src.label_vectors(), tgt.label_vectors()
        unchecked = 0
        for i in mism:
            keep = (idx[i] >= 0) & (d2[i] < 250)
            if np.any(d2[i][keep] >= 1.0):  # a far candidate could be gated by the bool-Probability rule: not recomputed here
                unchecked += 1
                continue
            sc = (a_t[idx[i][keep]] * a_s[i]).sum(0)
            top = np.sort(sc)
            assert top[-1] - top[-2] <= 1e-12 * max(top[-1], 1e-300), (i, top[-2:])
            assert sc[got[i] - 1] >= top[-1] * (1 - 1e-12)
        assert unchecked <= 2


def test_probability_gate_zeroes_weights(sicp, oracle, room):
    """GICPCostFunction::Probability -> bool (gicp_cost_function.h:75-87, em_icp.hpp:108): with the tiny epsilon of
    exec/scenenet_eval.cc:174 the Gaussian density underflows to exactly 0 for candidates far off the target plane (here
    an initial pose 0.9 m too high), and those weights must come out as exact zeros — the same ones as the oracle's."""
    p = room
    eps = 1e-6
    init = np.array([0, 0, 0, 1.0, 0, 0, 0.9])
    src, tgt = sicp.Cloud(p["src_xyz"], p["src_labels"]), sicp.Cloud(p["tgt_xyz"], p["tgt_labels"])
    opts = sicp.default_options(sicp.ALGO_EM, cm=p["cm"], epsilon=eps)
    idx, w, d2 = sicp.correspondences(sicp.ALGO_EM, src, tgt, opts, init)
    ref = oracle.align_em(p["src_xyz"], p["src_labels"], p["tgt_xyz"], p["tgt_labels"], p["cm"], init, eps=eps)
    assert np.array_equal(idx, ref["corr0"])
    zero_gpu, zero_ref = (idx >= 0) & (w == 0.0), (ref["corr0"] >= 0) & (ref["w0"] == 0.0)
    assert zero_gpu.sum() > 1000                     # the underflow branch is really taken (label-compat alone is > 0 here)
    assert (idx >= 0).sum() - zero_gpu.sum() > 100   # ... and not for everything
    assert np.array_equal(zero_gpu, zero_ref)
    assert np.max(np.abs(w - ref["w0"])) <= 1e-12 * max(1.0, np.max(np.abs(ref["w0"])))


def test_errors(sicp, room):
    lab0 = room["src_labels"].copy()
    lab0[5] = 0
    c = sicp.Cloud(room["src_xyz"], lab0)
    with pytest.raises(sicp.SicpError):
        c.precompute(20, 1e-3, room["cm"])  # label 0 is out of range (em_icp.hpp:301)
    with pytest.raises(sicp.SicpError):
        sicp.Cloud(room["src_xyz"], None, layout=sicp.CLOUD_PER_CLASS)
    g = sicp.Cloud(room["src_xyz"])
    with pytest.raises(sicp.SicpError):
        g.normals()  # precompute has not run
    with pytest.raises(sicp.SicpError, match="unit quaternion"):
        sicp.register(sicp.ALGO_GICP, g, g, sicp.default_options(sicp.ALGO_GICP), np.array([0, 0, 0, 2.0, 0, 0, 0]))
    with pytest.raises(sicp.SicpError, match="unit quaternion"):
        sicp.register(sicp.ALGO_GICP, g, g, sicp.default_options(sicp.ALGO_GICP), np.array([0, 0, 0, 1.0, np.nan, 0, 0]))
    # every C-ABI entry point that takes a pose validates it (finite, unit quaternion)
    bad = np.array([0, 0, 0, 2.0, 0, 0, 0])
    lab = sicp.Cloud(room["src_xyz"], room["src_labels"])
    eopts = sicp.default_options(sicp.ALGO_EM, cm=room["cm"])
    gopts = sicp.default_options(sicp.ALGO_GICP)
    for call in (lambda: sicp.correspondences(sicp.ALGO_GICP, g, g, gopts, bad),
                 lambda: sicp.evaluate(sicp.ALGO_GICP, g, g, gopts, room["init"], bad),
                 lambda: sicp.evaluate(sicp.ALGO_GICP, g, g, gopts, bad, room["init"]),
                 lambda: sicp.fused_labels(lab, lab, eopts, bad),
                 lambda: sicp.knn(g, room["src_xyz"][:10], 1, pose7=bad),
                 lambda: sicp.label_agreement(lab, lab, 11, pose7=bad),
                 lambda: g.transform_f32(bad)):
        with pytest.raises(sicp.SicpError, match="unit quaternion"):
            call()


# ------------------------------------------------------------------------------------------------ degenerate inputs
def test_degenerate_clouds_knn_and_covariances(sicp, oracle):
    rng = np.random.default_rng(11)
    cases = {
        "identical": np.tile(np.array([[1.5, -2.0, 0.25]], dtype=np.float32), (300, 1)),       # zero-extent bounding box
        "collinear": np.outer(np.linspace(0, 50, 400), [1, 2, -1]).astype(np.float32),          # rank-1 covariances
        "coplanar_far": (np.c_[rng.uniform(-3, 3, (500, 2)), np.zeros(500)] + np.array([8000.0, -9000.0, 50.0])).astype(np.float32),  # f32 cancellation at 1e4 m
        "lattice": np.stack(np.meshgrid(np.arange(8), np.arange(8), np.arange(8), indexing="ij"), -1).reshape(-1, 3).astype(np.float32),
    }
    for name, xyz in cases.items():
        c = sicp.Cloud(xyz)
        q = xyz[:: max(1, len(xyz) // 97)]
        for k in (1, 4, 20):
            idx, d2 = sicp.knn(c, q, k)
            ridx, rd2 = oracle.knn(xyz, q, k, brute=True)
            assert np.array_equal(idx, ridx), (name, k)
            assert np.array_equal(d2, rd2), (name, k)
        c.precompute(20, 1e-3)
        ref = oracle.covariances(xyz, 20, 1e-3, want_nn=True)
        assert np.array_equal(c.self_neighbours(), ref["nn"]), name
        assert np.array_equal(c.normals(), ref["normals"]), name   # same Jacobi sweeps even on rank-deficient input


def test_empty_problem_and_tiny_clouds(sicp, oracle, pkg):
    rng = np.random.default_rng(0)
    src = rng.normal(size=(50, 3)).astype(np.float32)
    tgt = src + np.float32(1000.0)                                  # every correspondence fails the 250 m^2 gate
    init = oracle.se3_exp([0.1, 0.2, 0.3, 0.01, 0.02, 0.03])
    res = sicp.register(sicp.ALGO_GICP, sicp.Cloud(src), sicp.Cloud(tgt), sicp.default_options(sicp.ALGO_GICP), init)
    ref = oracle.align_gicp(src, tgt, init)
    assert np.array_equal(res["pose"], init) and np.array_equal(ref["pose"], init)   # empty problem leaves the pose unchanged
    assert res["outer_iter"] == ref["outer_iter"] == 1 and res["n_corr_last"] == 0
    # fewer target points than correspondences asked for (EM wants 4): the missing ones are reported as -1
    tiny = sicp.Cloud(src[:3])
    idx, d2 = sicp.knn(tiny, src[:10], 4)
    assert np.all(idx[:, 3] == -1) and np.all(np.isinf(d2[:, 3])) and np.all(idx[:, :3] >= 0)
    # zero-point cloud: creation and queries are legal, nothing is found
    empty = sicp.Cloud(np.zeros((0, 3), dtype=np.float32))
    idx, d2 = sicp.knn(empty, src[:5], 1)
    assert np.all(idx == -1)


def test_semantic_class_rules(sicp, oracle, pkg):
    """semantic_icp.hpp:50-51: a class is used only if the target has it and the source class has > 400 points."""
    p = pkg.synth.room_pair(seed=21, n_points=6000)
    sl, tl = p["src_labels"].copy(), p["tgt_labels"].copy()
    tl[tl == 3] = 4                     # class 3 absent from the target
    small = np.nonzero(sl == 5)[0]
    sl[small[350:]] = 6                 # class 5 shrinks to 350 source points (<= 400)
    src = sicp.Cloud(p["src_xyz"], sl, layout=sicp.CLOUD_PER_CLASS)
    tgt = sicp.Cloud(p["tgt_xyz"], tl, layout=sicp.CLOUD_PER_CLASS)
    opts = sicp.default_options(sicp.ALGO_SEMANTIC)
    idx, w, d2 = sicp.correspondences(sicp.ALGO_SEMANTIC, src, tgt, opts, p["init"])
    ref = oracle.align_semantic(p["src_xyz"], sl, p["tgt_xyz"], tl, p["init"])
    assert np.array_equal(idx, ref["corr0"])
    assert np.all(idx[sl == 3] == -1) and np.all(idx[sl == 5] == -1)
    res = sicp.register(sicp.ALGO_SEMANTIC, src, tgt, opts, p["init"])
    rot, trans = pkg.synth.pose_error(res["pose"], ref["pose"])
    assert rot < ROT_TOL and trans < TRANS_TOL and res["outer_iter"] == ref["outer_iter"]


def test_final_cloud_transform_is_float_matrix_math(sicp, oracle, room):
    """gicp.hpp:166-171: `trans.matrix().cast<float>()` applied by pcl::transformPointCloud in float arithmetic."""
    c = sicp.Cloud(room["src_xyz"])
    pose = room["T_gt"]
    got = c.transform_f32(pose)
    M = oracle.se3_matrix(pose).astype(np.float32)  # Eigen toRotationMatrix op sequence, then the float cast
    x = room["src_xyz"]
    exp = np.stack([((M[r, 0] * x[:, 0] + M[r, 1] * x[:, 1]) + M[r, 2] * x[:, 2]) + M[r, 3] for r in range(3)], axis=1)
    assert got.dtype == np.float32 and np.array_equal(got, exp.astype(np.float32))


def test_class_order_with_large_sparse_labels(sicp, oracle):
    rng = np.random.default_rng(9)
    xyz = rng.normal(size=(2000, 3)).astype(np.float32)
    pool = np.array([4_000_000_000, 7, 65_536, 65_535, 123_456, 0], dtype=np.uint32)   # direct-table and hash-map labels mixed
    labels = pool[rng.integers(0, len(pool), size=2000)]
    c = sicp.Cloud(xyz, labels, layout=sicp.CLOUD_PER_CLASS)
    labs, sizes = c.classes()
    rl, rs, order = oracle.label_split(labels)
    assert np.array_equal(labs, rl) and np.array_equal(sizes, np.diff(rs))
    idx, d2 = sicp.knn(c, xyz[:300], 1, q_labels=labels[:300])
    for lab in pool:
        qs = np.nonzero(labels[:300] == lab)[0]
        ts = np.nonzero(labels == lab)[0]
        ridx, rd2 = oracle.knn(xyz[ts], xyz[qs], 1)
        assert np.array_equal(idx[qs, 0], ts[ridx[:, 0]]) and np.array_equal(d2[qs], rd2)


def test_class_partition_on_device(sicp, oracle, room):
    """pcl_2_semantic.h:24-35 on the device: first-appearance class order, class sizes, too many classes."""
    import torch

    p = room
    dx = torch.from_numpy(np.ascontiguousarray(p["src_xyz"])).cuda()
    dl = torch.from_numpy(p["src_labels"].astype(np.int32)).cuda()
    torch.cuda.synchronize()
    c = sicp.Cloud.from_device(dx.data_ptr(), dl.data_ptr(), len(p["src_xyz"]), layout=sicp.CLOUD_PER_CLASS)
    labs, sizes = c.classes()
    rl, rs, _ = oracle.label_split(p["src_labels"])
    assert np.array_equal(labs, rl) and np.array_equal(sizes, np.diff(rs))
    c.precompute(20, 1e-3)
    ref = oracle.covariances_per_class(p["src_xyz"], p["src_labels"], 20, 1e-3)
    assert np.array_equal(c.normals(), ref["normals"])            # device-input cloud == host-input cloud == oracle
    rng = np.random.default_rng(2)
    xyz = rng.normal(size=(3000, 3)).astype(np.float32)
    ok = sicp.Cloud(xyz, (np.arange(3000) % 128 + 5).astype(np.uint32), layout=sicp.CLOUD_PER_CLASS)   # exactly 128 classes
    assert len(ok.classes()[0]) == 128
    for nclass in (129, 700):                                       # over the class limit / over the table size
        with pytest.raises(sicp.SicpError, match="at most 128"):
            sicp.Cloud(xyz, (np.arange(3000) % nclass).astype(np.uint32) * 977, layout=sicp.CLOUD_PER_CLASS)


def test_device_label_ranges_are_checked_when_the_registration_completes(sicp, pkg, room):
    """Clouds created from DEVICE labels: the 1..N check of EM-ICP (em_icp.hpp:301 indexes label - 1) reads the range back
    with the registration instead of synchronising before it; results equal those of host-created clouds, bad labels still
    fail (lone and batch), and a direct precompute call keeps the immediate check."""
    import torch

    p = room
    opts = sicp.default_options(sicp.ALGO_EM, cm=p["cm"])
    n = len(p["src_xyz"])

    def dev_cloud(xyz, lab):
        dx = torch.from_numpy(np.ascontiguousarray(xyz)).cuda()
        dl = torch.from_numpy(np.ascontiguousarray(lab).astype(np.int32)).cuda()
        torch.cuda.synchronize()
        return sicp.Cloud.from_device(dx.data_ptr(), dl.data_ptr(), len(xyz)), (dx, dl)

    s, keep_s = dev_cloud(p["src_xyz"], p["src_labels"])
    t, keep_t = dev_cloud(p["tgt_xyz"], p["tgt_labels"])
    ref = sicp.register(sicp.ALGO_EM, sicp.Cloud(p["src_xyz"], p["src_labels"]), sicp.Cloud(p["tgt_xyz"], p["tgt_labels"]), opts, p["init"])
    got = sicp.register(sicp.ALGO_EM, s, t, opts, p["init"])
    assert np.array_equal(got["pose"], ref["pose"]) and got["outer_iter"] == ref["outer_iter"]
    for bad_value in (0, p["N"] + 1):
        lab = p["src_labels"].copy()
        lab[n // 2] = bad_value
        b, keep_b = dev_cloud(p["src_xyz"], lab)
        with pytest.raises(sicp.SicpError, match="1..N"):
            sicp.register(sicp.ALGO_EM, b, t, opts, p["init"])
        b2, keep_b2 = dev_cloud(p["src_xyz"], lab)
        with pytest.raises(sicp.SicpError, match="1..N"):
            sicp.register_batch(sicp.ALGO_EM, [s, b2, s], [t, t, t], opts, np.stack([p["init"]] * 3))
        b3, keep_b3 = dev_cloud(p["src_xyz"], lab)
        with pytest.raises(sicp.SicpError, match="1..N"):
            b3.precompute(20, 1e-3, p["cm"])
    again = sicp.register(sicp.ALGO_EM, s, t, opts, p["init"])  # the library is usable after a failed batch
    assert np.array_equal(again["pose"], ref["pose"])
