"""GPU parity at the sizes BASELINE.json's configs name (all through the C ABI, oracle as the checker), plus the
size-independent properties the domain offers (sortedness, self-neighbour, idempotence of a converged pose, determinism,
batch == single).  configs[0] (10k room pair) is covered by tests/test_gpu_parity.py."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROT_TOL, TRANS_TOL = 1e-5, 1e-4  # rad, m — north_star pose tolerance


@pytest.fixture(scope="module")
def kitti_full(pkg):
    return pkg.synth.kitti_pair(pair=0)  # 120,000 pts/scan, N=20


def _em(sicp, p):
    src, tgt = sicp.Cloud(p["src_xyz"], p["src_labels"]), sicp.Cloud(p["tgt_xyz"], p["tgt_labels"])
    return src, tgt, sicp.default_options(sicp.ALGO_EM, cm=p["cm"])


# ------------------------------------------------------------------------------------------------ configs[1]: KITTI-shape, 120k, N=20, EM
def test_config2_kitti_full_size_em_parity(sicp, oracle, pkg, kitti_full):
    p = kitti_full
    src, tgt, opts = _em(sicp, p)
    ref = oracle.align_em(p["src_xyz"], p["src_labels"], p["tgt_xyz"], p["tgt_labels"], p["cm"], p["init"])
    idx, w, d2 = sicp.correspondences(sicp.ALGO_EM, src, tgt, opts, p["init"])
    assert np.array_equal(idx, ref["corr0"])                      # 480,000 correspondence slots: bit-exact
    m = ref["corr0"] >= 0
    assert np.array_equal(d2[m], ref["d20"][m])
    assert np.max(np.abs(w - ref["w0"])) <= 1e-12 * max(1.0, np.max(np.abs(ref["w0"])))
    res = sicp.register(sicp.ALGO_EM, src, tgt, opts, p["init"])
    rot, trans = pkg.synth.pose_error(res["pose"], ref["pose"])
    assert rot < ROT_TOL and trans < TRANS_TOL, (rot, trans)
    assert res["outer_iter"] == ref["outer_iter"]
    assert list(res["pass_lm_iters"]) == list(ref["pass_lm_iters"])


def test_config2_pass_result_is_a_stationary_point_scipy(sicp, pkg):
    """Independent check of the M-step (not the oracle, not Ceres' rules): the pose a pass ends with must be a minimiser of
    that pass's effective objective  1/2 sum_i w_i * 9 log(1 + sqrt(r_i^2 + eps)/9),  r_i = d^T (C_t + R C_s R^T)^-1 d
    (SURVEY.md Appendix B.2), built here in numpy from explicit 3x3 covariances and inverses over the GPU's own
    correspondences of that pass.  scipy.optimize.least_squares restarted AT that pose must not move it."""
    from scipy.optimize import least_squares
    from scipy.spatial.transform import Rotation

    p = pkg.synth.kitti_pair(pair=0, n_points=60_000)  # 240,000 residual slots: the numpy objective stays within seconds
    src, tgt, opts = _em(sicp, p)
    res = sicp.register(sicp.ALGO_EM, src, tgt, opts, p["init"])
    n = res["outer_iter"]
    before = res["pass_pose"][n - 2] if n >= 2 else p["init"]
    after = res["pass_pose"][n - 1]
    idx, w, d2 = sicp.correspondences(sicp.ALGO_EM, src, tgt, opts, before)  # correspondences + E-step weights of the last pass
    keep = (idx >= 0) & (w > 0)
    si, ci = np.nonzero(keep)
    ti, ww = idx[si, ci], w[si, ci]
    eps = 1e-3
    ns, nt = src.normals()[si], tgt.normals()[ti]
    eye = np.eye(3)[None]
    Cs = eye - (1 - eps) * ns[:, :, None] * ns[:, None, :]
    Ct = eye - (1 - eps) * nt[:, :, None] * nt[:, None, :]
    ps, pt = p["src_xyz"][si].astype(np.float64), p["tgt_xyz"][ti].astype(np.float64)
    R0 = Rotation.from_quat(after[:4]).as_matrix()
    t0 = after[4:]

    def residuals(delta):  # T = T_after * exp(delta) with delta = (upsilon, omega)  (local_parameterization_se3.h:22)
        dR = Rotation.from_rotvec(delta[3:]).as_matrix()
        om = delta[3:]
        th = np.linalg.norm(om)
        O = np.array([[0, -om[2], om[1]], [om[2], 0, -om[0]], [-om[1], om[0], 0.0]])
        V = np.eye(3) + 0.5 * O + O @ O / 6.0 if th < 1e-6 else np.eye(3) + (1 - np.cos(th)) / th**2 * O + (th - np.sin(th)) / th**3 * (O @ O)
        R, t = R0 @ dR, t0 + R0 @ (V @ delta[:3])
        d = pt - (ps @ R.T + t)
        M = np.linalg.inv(Ct + np.einsum("ij,njk,lk->nil", R, Cs, R))
        r = np.einsum("ni,nij,nj->n", d, M, d)
        rho = ww * 9.0 * np.log1p(np.sqrt(r * r + np.finfo(float).eps) / 9.0)  # ScaledLoss(Cauchy(3), w) o SQLoss
        return np.sqrt(rho)

    f0 = residuals(np.zeros(6))
    sol = least_squares(residuals, np.zeros(6), method="trf", jac="2-point", x_scale=1.0, xtol=1e-12, ftol=1e-15, gtol=1e-15, max_nfev=12)
    assert np.linalg.norm(sol.x[3:]) < 0.3 * ROT_TOL and np.linalg.norm(sol.x[:3]) < 0.3 * TRANS_TOL, sol.x
    assert 0.5 * (f0 @ f0) - sol.cost <= 1e-9 * 0.5 * (f0 @ f0)


@pytest.mark.parametrize("k", [1, 4, 20])
def test_config2_knn_full_size_bit_exact_and_sorted(sicp, oracle, kitti_full, k):
    p = kitti_full
    tgt = sicp.Cloud(p["tgt_xyz"])
    idx, d2 = sicp.knn(tgt, p["src_xyz"], k, pose7=p["T_gt"])
    q = oracle.transform_points(p["T_gt"], p["src_xyz"])
    ridx, rd2 = oracle.knn(p["tgt_xyz"], q, k)
    assert np.array_equal(idx, ridx) and np.array_equal(d2, rd2)
    # size-independent properties: ascending (d2, index) order, no repeated neighbour
    assert np.all(np.diff(d2, axis=1) >= 0)
    ties = np.diff(d2, axis=1) == 0
    assert np.all(np.diff(idx, axis=1)[ties] > 0)
    if k > 1:
        s = np.sort(idx, axis=1)
        assert np.all(np.diff(s, axis=1) > 0)


def test_config2_self_neighbour_and_idempotence(sicp, pkg, kitti_full):
    p = kitti_full
    src, tgt, opts = _em(sicp, p)
    src.precompute(20, 1e-3, p["cm"])
    nn = src.self_neighbours()
    assert np.array_equal(nn[:, 0], np.arange(src.n, dtype=np.int32))  # every point is its own nearest neighbour (d2 = 0)
    assert np.allclose(np.linalg.norm(src.normals(), axis=1), 1.0, atol=1e-12)
    a = sicp.register(sicp.ALGO_EM, src, tgt, opts, p["init"])
    b = sicp.register(sicp.ALGO_EM, src, tgt, opts, a["pose"])     # restart at the converged pose
    rot, trans = pkg.synth.pose_error(a["pose"], b["pose"])
    assert b["outer_iter"] == 1 and rot < 3.2e-3 and trans < 3.2e-3  # one pass, step below the sqrt(1e-5) stop threshold
    c = sicp.register(sicp.ALGO_EM, src, tgt, opts, p["init"])
    assert np.array_equal(a["pose"], c["pose"])                     # fixed-order reductions: run-to-run bit-identical


# ------------------------------------------------------------------------------------------------ configs[2]: NYU-shape, 307,200 pts, N=40
@pytest.fixture(scope="module")
def nyu_full(pkg):
    return pkg.synth.nyu_pair(pair=0)


def test_config3_nyu_semantic_icp_parity(sicp, oracle, pkg, nyu_full):
    p = nyu_full
    src = sicp.Cloud(p["src_xyz"], p["src_labels"], layout=sicp.CLOUD_PER_CLASS)
    tgt = sicp.Cloud(p["tgt_xyz"], p["tgt_labels"], layout=sicp.CLOUD_PER_CLASS)
    opts = sicp.default_options(sicp.ALGO_SEMANTIC)
    ref = oracle.align_semantic(p["src_xyz"], p["src_labels"], p["tgt_xyz"], p["tgt_labels"], p["init"])
    idx, w, d2 = sicp.correspondences(sicp.ALGO_SEMANTIC, src, tgt, opts, p["init"])
    assert np.array_equal(idx, ref["corr0"])                      # per-class 1-NN incl. the >400-point class rule
    res = sicp.register(sicp.ALGO_SEMANTIC, src, tgt, opts, p["init"])
    rot, trans = pkg.synth.pose_error(res["pose"], ref["pose"])
    assert rot < ROT_TOL and trans < TRANS_TOL, (rot, trans)
    assert res["outer_iter"] == ref["outer_iter"]


def test_config3_nyu_em_parity(sicp, oracle, pkg, nyu_full):
    p = nyu_full
    src, tgt, opts = _em(sicp, p)
    ref = oracle.align_em(p["src_xyz"], p["src_labels"], p["tgt_xyz"], p["tgt_labels"], p["cm"], p["init"])
    idx, w, d2 = sicp.correspondences(sicp.ALGO_EM, src, tgt, opts, p["init"])
    assert np.array_equal(idx, ref["corr0"])                      # 1,228,800 slots, N = 40 label vectors
    assert np.max(np.abs(w - ref["w0"])) <= 1e-12 * max(1.0, np.max(np.abs(ref["w0"])))
    res = sicp.register(sicp.ALGO_EM, src, tgt, opts, p["init"])
    rot, trans = pkg.synth.pose_error(res["pose"], ref["pose"])
    assert rot < ROT_TOL and trans < TRANS_TOL, (rot, trans)
    assert res["outer_iter"] == ref["outer_iter"]


# ------------------------------------------------------------------------------------------------ configs[3]: odometry sequence, sharded batch
def test_config4_sequence_batch(sicp, oracle, pkg):
    frames, poses, cm = pkg.synth.kitti_sequence(6, n_points=30_000, n_rings=32, n_az=1000)
    clouds = [sicp.Cloud(x, l) for x, l in frames]
    opts = sicp.default_options(sicp.ALGO_EM, cm=cm)
    n_pairs = len(frames) - 1
    ident = np.tile(np.array([0, 0, 0, 1, 0, 0, 0], dtype=np.float64), (n_pairs, 1))

    def register_fn(lo, hi):  # pair i: source = frame i+1, target = frame i
        return sicp.register_batch(sicp.ALGO_EM, [clouds[i + 1] for i in range(lo, hi)], [clouds[i] for i in range(lo, hi)], opts, ident[lo:hi])

    rec, (lo, hi) = pkg.shard.register_sharded(register_fn, n_pairs, 1, rank=0, world=1)
    assert (lo, hi) == (0, n_pairs) and rec.shape == (n_pairs, pkg.shard.RECORD)
    out = pkg.shard.from_records(rec)
    for i in (0, n_pairs - 1):  # oracle on the first and last pair
        ref = oracle.align_em(frames[i + 1][0], frames[i + 1][1], frames[i][0], frames[i][1], cm, ident[0])
        rot, trans = pkg.synth.pose_error(out[i]["pose"], ref["pose"])
        assert rot < ROT_TOL and trans < TRANS_TOL, (i, rot, trans)
        assert out[i]["outer_iter"] == ref["outer_iter"]
    for i in range(n_pairs):    # and the odometry itself is right (ground truth within scan noise)
        gt = pkg.synth.relative_pose(poses[i], poses[i + 1])
        rot, trans = pkg.synth.pose_error(out[i]["pose"], gt)
        assert rot < 5e-3 and trans < 5e-2, (i, rot, trans)


# ------------------------------------------------------------------------------------------------ configs[4]: initial-pose sweep sharing one pair's clouds
def test_config5_pose_sweep_batch(sicp, oracle, pkg):
    p = pkg.synth.kitti_pair(pair=3, n_points=30_000, n_rings=32, n_az=1000)
    src, tgt, opts = _em(sicp, p)
    n_inits = 24
    inits = pkg.synth.sweep_inits(p["T_gt"], n_inits, pair=3, max_angle_deg=4.0, max_trans=0.8)
    batch = sicp.register_batch(sicp.ALGO_EM, [src] * n_inits, [tgt] * n_inits, opts, inits)  # clouds, covariances, label vectors built once
    ok = 0
    for j in (0, 5, 23):  # singles: same trajectory (half-size LM grids in a batch: <= 1e-9, see test_register_batch_matches_single)
        one = sicp.register(sicp.ALGO_EM, src, tgt, opts, inits[j])
        rot, trans = pkg.synth.pose_error(batch[j]["pose"], one["pose"])
        assert rot < 1e-8 and trans < 1e-8 and batch[j]["outer_iter"] == one["outer_iter"], (j, rot, trans)
    ref = oracle.align_em(p["src_xyz"], p["src_labels"], p["tgt_xyz"], p["tgt_labels"], p["cm"], inits[5])
    rot, trans = pkg.synth.pose_error(batch[5]["pose"], ref["pose"])
    assert rot < ROT_TOL and trans < TRANS_TOL and batch[5]["outer_iter"] == ref["outer_iter"]
    for b in batch:
        rot, trans = pkg.synth.pose_error(b["pose"], p["T_gt"])
        ok += rot < 5e-3 and trans < 5e-2
    assert ok >= 0.75 * n_inits  # the convergence basin: most small perturbations come back to the ground truth


def test_paired_solves_match_unpaired(sicp, pkg, monkeypatch):
    """SICP_PAIR=1: two registrations share one LM launch per pass (lm_pair_kernel: the blocks alternate between the two
    solves, each with its own controller block).  Same pass counts and LM iteration counts as the unpaired batch, poses
    equal far inside the tolerance, an odd batch size handled (the last job runs against an already-converged dummy), and
    bit-identical results for identical inputs."""
    pairs = [pkg.synth.kitti_pair(pair=i, n_points=20_000, n_rings=32, n_az=700) for i in range(5)]
    opts = sicp.default_options(sicp.ALGO_EM, cm=pairs[0]["cm"])
    cl = [(sicp.Cloud(q["src_xyz"], q["src_labels"]), sicp.Cloud(q["tgt_xyz"], q["tgt_labels"])) for q in pairs]
    inits = np.stack([q["init"] for q in pairs])
    monkeypatch.setenv("SICP_PAIR", "0")
    plain = sicp.register_batch(sicp.ALGO_EM, [c[0] for c in cl], [c[1] for c in cl], opts, inits)
    monkeypatch.setenv("SICP_PAIR", "1")
    paired = sicp.register_batch(sicp.ALGO_EM, [c[0] for c in cl], [c[1] for c in cl], opts, inits)
    for a, b in zip(plain, paired):
        # different block -> partial-sum grouping than the unpaired launch: equal to rounding, not to the bit (compared on
        # the 7 pose numbers; synth.pose_error's acos has a floor of sqrt(2 ulp) = 2.1e-8 rad)
        assert np.max(np.abs(np.asarray(a["pose"]) - np.asarray(b["pose"]))) < 1e-9, (a["pose"], b["pose"])
        assert a["outer_iter"] == b["outer_iter"] and list(a["pass_lm_iters"]) == list(b["pass_lm_iters"])
    same = sicp.register_batch(sicp.ALGO_EM, [cl[0][0]] * 3, [cl[0][1]] * 3, opts, inits[:1].repeat(3, 0))
    assert all(np.array_equal(x["pose"], same[0]["pose"]) for x in same)  # pair members and the odd job: same block order, same bits
    for algo, mk in ((sicp.ALGO_GICP, lambda q: (sicp.Cloud(q["src_xyz"]), sicp.Cloud(q["tgt_xyz"]))),):
        g = [mk(q) for q in pairs[:4]]
        gopts = sicp.default_options(algo)
        monkeypatch.setenv("SICP_PAIR", "0")
        a4 = sicp.register_batch(algo, [c[0] for c in g], [c[1] for c in g], gopts, inits[:4])
        monkeypatch.setenv("SICP_PAIR", "1")
        b4 = sicp.register_batch(algo, [c[0] for c in g], [c[1] for c in g], gopts, inits[:4])
        for a, b in zip(a4, b4):
            assert np.max(np.abs(np.asarray(a["pose"]) - np.asarray(b["pose"]))) < 1e-9 and a["outer_iter"] == b["outer_iter"]
