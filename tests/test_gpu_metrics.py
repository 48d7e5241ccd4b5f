"""SURVEY.md §8(f) rows 2-3 — the evaluation steps either side of the registration path, device vs oracle:
label agreement through 1-NN (exec/roc_metrics.h, exec/nyu_metrics.h), SE(3) errors (exec/kitti_metrics.h), range
filter (exec/filter_range.h)."""
import numpy as np
import pytest


@pytest.mark.gpu
def test_pose_errors_match_oracle(sicp, oracle):
    rng = np.random.default_rng(5)
    gt = np.stack([oracle.se3_exp(rng.normal(size=6) * s) for s in (1e-6, 1e-3, 0.1, 1.0, 2.5)])
    est = np.stack([oracle.se3_plus(g, rng.normal(size=6) * 1e-2) for g in gt])
    got = sicp.pose_errors(gt, est)
    ref = np.stack([oracle.pose_errors(g, e) for g, e in zip(gt, est)])
    assert np.allclose(got, ref, rtol=1e-9, atol=1e-18)
    assert np.allclose(sicp.pose_errors(gt, gt), 0.0, atol=1e-25)


@pytest.mark.gpu
def test_label_agreement_matches_oracle(sicp, oracle, pkg):
    p = pkg.synth.kitti_pair(pair=2, n_points=40_000, n_rings=32, n_az=1300)
    N = p["N"] + 1                                      # labels are 1..N; the reference's matrix is indexed by label
    src, tgt = sicp.Cloud(p["src_xyz"], p["src_labels"]), sicp.Cloud(p["tgt_xyz"], p["tgt_labels"])
    for pose in (None, p["T_gt"]):
        got = sicp.label_agreement(src, tgt, N, pose7=pose, want_pairs=True)
        ref = oracle.label_agreement(p["src_xyz"], p["src_labels"], p["tgt_xyz"], p["tgt_labels"], N, pose7=pose)
        assert np.array_equal(got["confusion"], ref["confusion"])          # integer counts: exact
        assert np.array_equal(got["pairs"], ref["pairs"])                  # (label_source, label_target) per point, gate included
        assert got["inliers"] == ref["inliers"] and got["total"] == ref["total"]
        assert abs(got["dist"] - ref["dist"]) <= 1e-11 * ref["dist"]       # same float sqrt values, different double summation order
    aligned = sicp.label_agreement(src, tgt, N, pose7=p["T_gt"])
    raw = sicp.label_agreement(src, tgt, N)
    assert aligned["inliers"] / aligned["total"] > raw["inliers"] / raw["total"]    # alignment improves label agreement
    again = sicp.label_agreement(src, tgt, N, pose7=p["T_gt"])
    assert again["dist"] == aligned["dist"]                                          # fixed-order sums: run-to-run identical
    with pytest.raises(sicp.SicpError):
        sicp.label_agreement(src, tgt, 5)                                            # labels >= n_labels: error, not an OOB write


@pytest.mark.gpu
def test_filter_range_matches_oracle(sicp, oracle, pkg):
    p = pkg.synth.kitti_pair(pair=4, n_points=50_000, n_rings=32, n_az=1600)
    for rng_m in (40.0, 5.0, 1e-3, 1e6):               # exec/kitti_eval.cc:124-127 uses 40 m
        keep = sicp.filter_range(p["src_xyz"], rng_m)
        ref = oracle.filter_range(p["src_xyz"], rng_m)
        assert np.array_equal(keep, ref)
    edge = np.array([[3, 4, 0], [3, 4, 1e-3], [0, 0, 5], [np.nextafter(np.float32(5), np.float32(6)), 0, 0]], dtype=np.float32)
    assert list(sicp.filter_range(edge, 5.0)) == list(oracle.filter_range(edge, 5.0)) == [0, 2]   # `>` is strict: r == range stays
    assert len(sicp.filter_range(np.zeros((0, 3), dtype=np.float32), 40.0)) == 0


@pytest.mark.gpu
def test_iterative_mean_and_pose_fusion(sicp, oracle, pkg):
    """§8(f) row 4 on the device (sicp_iterative_mean, sicp_pose_fusion) against the oracle's restatement of
    impl/semantic_icp.hpp:169-265."""
    from scipy.spatial.transform import Rotation

    rng = np.random.default_rng(8)
    base = np.concatenate([Rotation.from_rotvec([0.1, -0.2, 0.05]).as_quat(), [1.0, 2.0, 0.5]])
    for n, spread in ((1, 0.0), (9, 0.05), (40, 0.3)):  # 40 > one warp of residual blocks
        poses = np.array([oracle.se3_plus(base, rng.normal(scale=spread, size=6)) for _ in range(n)])
        got, ok = sicp.iterative_mean(poses, 100)
        ref, rok = oracle.iterative_mean(poses, 100)
        assert ok == rok and np.max(np.abs(got - ref)) < 1e-12
        covs = np.array([np.diag(rng.uniform(0.5, 2, 6)) * 1e-4 + 1e-6 for _ in range(n)])
        fused, it = sicp.pose_fusion(poses, covs, base)
        rfused, rit = oracle.pose_fusion(poses, covs, base)
        # both minimise the same objective with finite-difference Jacobians (rounded differently: FMA vs none) down to
        # tolerances of 1e-14, where the cost is flat: the minimisers agree far inside the pose tolerance, not to the bit
        rot, trans = pkg.synth.pose_error(fused, rfused)
        assert rot < 1e-6 and trans < 1e-6, (rot, trans)
        assert (it == 0) == (n == 1)
    got, ok = sicp.iterative_mean(np.array([base, oracle.se3_plus(base, np.array([3.0, 0, 0, 0, 0, 2.5]))]), 1)
    ref, rok = oracle.iterative_mean(np.array([base, oracle.se3_plus(base, np.array([3.0, 0, 0, 0, 0, 2.5]))]), 1)
    assert ok == rok and np.max(np.abs(got - ref)) < 1e-12  # step limit reached: the last iterate is returned
    with pytest.raises(sicp.SicpError):
        sicp.iterative_mean(np.array([[0, 0, 0, 2.0, 0, 0, 0]]), 5)
