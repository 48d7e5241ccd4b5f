"""Seeded synthetic workloads for the registration hot path.

Every generator returns float32 xyz, uint32 labels in 1..N, the confusion matrix handed to
EM-ICP and the ground-truth source->target pose as [qx,qy,qz,qw,tx,ty,tz] (Sophus::SE3d::data()
order, reference gicp_cost_function.h:64-70).  All generators add Gaussian noise so exact
distance ties are vanishingly rare (the exact-kNN contract is (d2_f32, index) lexicographic).

Host-side numpy only; nothing here is on the measured path.
"""
from __future__ import annotations

import numpy as np

# ----------------------------------------------------------------------------- SE(3) helpers


def quat_from_axis_angle(axis, angle):
    axis = np.asarray(axis, dtype=np.float64)
    axis = axis / np.linalg.norm(axis)
    s = np.sin(angle / 2.0)
    return np.array([axis[0] * s, axis[1] * s, axis[2] * s, np.cos(angle / 2.0)])


def quat_to_R(q):
    x, y, z, w = q
    return np.array(
        [
            [1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
            [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
            [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)],
        ]
    )


def R_to_quat(R):
    w = np.sqrt(max(0.0, 1 + R[0, 0] + R[1, 1] + R[2, 2])) / 2
    if w > 1e-6:
        x = (R[2, 1] - R[1, 2]) / (4 * w)
        y = (R[0, 2] - R[2, 0]) / (4 * w)
        z = (R[1, 0] - R[0, 1]) / (4 * w)
    else:  # 180 deg: not produced by these generators
        x = np.sqrt(max(0.0, 1 + R[0, 0] - R[1, 1] - R[2, 2])) / 2
        y = np.sqrt(max(0.0, 1 - R[0, 0] + R[1, 1] - R[2, 2])) / 2
        z = np.sqrt(max(0.0, 1 - R[0, 0] - R[1, 1] + R[2, 2])) / 2
    q = np.array([x, y, z, w])
    return q / np.linalg.norm(q)


def pose7(R, t):
    return np.concatenate([R_to_quat(np.asarray(R)), np.asarray(t, dtype=np.float64)])


def pose7_matrix(p):
    M = np.eye(4)
    M[:3, :3] = quat_to_R(p[:4])
    M[:3, 3] = p[4:]
    return M


def identity_pose():
    return np.array([0, 0, 0, 1, 0, 0, 0], dtype=np.float64)


def random_unit(rng):
    v = rng.normal(size=3)
    return v / np.linalg.norm(v)


def pose_error(p_a, p_b):
    """(rotation angle [rad], translation distance [m]) between two pose7."""
    Ra, Rb = quat_to_R(p_a[:4]), quat_to_R(p_b[:4])
    c = (np.trace(Ra.T @ Rb) - 1) / 2
    return float(np.arccos(np.clip(c, -1, 1))), float(np.linalg.norm(p_a[4:] - p_b[4:]))


def confusion_matrix(N, diag=0.8):
    """CM_true = diag*I + (1-diag)/N (row = true class, column = observed class)."""
    return diag * np.eye(N) + (1.0 - diag) / N * np.ones((N, N))


def observe_labels(rng, true_labels, cm):
    """observed label ~ Categorical(cm[true-1]); labels are 1-based (em_icp.hpp:301)."""
    N = cm.shape[0]
    cdf = np.cumsum(cm, axis=1)
    cdf[:, -1] = 1.0
    u = rng.random(true_labels.shape[0])
    obs = (u[:, None] > cdf[true_labels - 1]).sum(axis=1)
    return (np.minimum(obs, N - 1) + 1).astype(np.uint32)


# ----------------------------------------------------------------------------- scene primitives


class Scene:
    """Static world made of planes (bounded), axis-aligned boxes and vertical cylinders.
    Each primitive carries a true semantic class in 1..N."""

    def __init__(self):
        self.planes = []  # (point, normal, bounds(lo,hi) or None, cls)
        self.boxes = []  # (lo, hi, cls)
        self.cyls = []  # (cx, cy, r, z0, z1, cls)
        self.ground_bands = None  # optional: classes by |y| for the z=0 plane

    def cast(self, origin, dirs, max_range):
        """Nearest hit along rays origin + s*dirs. Returns (s, cls) with s=inf for misses."""
        n = dirs.shape[0]
        best = np.full(n, np.inf)
        cls = np.zeros(n, dtype=np.int64)
        o = np.asarray(origin, dtype=np.float64)
        for (p0, nrm, bounds, c) in self.planes:
            denom = dirs @ nrm
            with np.errstate(divide="ignore", invalid="ignore"):
                s = ((p0 - o) @ nrm) / denom
            ok = (np.abs(denom) > 1e-12) & (s > 1e-6) & (s < best)
            if bounds is not None:
                hit = o + np.where(ok, s, 0.0)[:, None] * dirs  # rays parallel to the plane have s = inf/nan: they are not `ok`
                lo, hi = bounds
                ok &= np.all(hit >= lo - 1e-9, axis=1) & np.all(hit <= hi + 1e-9, axis=1)
            best = np.where(ok, s, best)
            if callable(c):
                hit = o + np.where(ok, s, 0.0)[:, None] * dirs
                cls = np.where(ok, c(hit), cls)
            else:
                cls = np.where(ok, c, cls)
        for (lo, hi, c) in self.boxes:
            with np.errstate(divide="ignore", invalid="ignore"):
                t1 = (lo - o) / dirs
                t2 = (hi - o) / dirs
            lo3, hi3 = np.minimum(t1, t2), np.maximum(t1, t2)  # NaN (0/0: ray parallel to a slab it starts on) is ignored, like nanmax/nanmin
            tmin = np.fmax(np.fmax(lo3[:, 0], lo3[:, 1]), lo3[:, 2])
            tmax = np.fmin(np.fmin(hi3[:, 0], hi3[:, 1]), hi3[:, 2])
            ok = (tmax >= tmin) & (tmin > 1e-6) & (tmin < best)
            best = np.where(ok, tmin, best)
            cls = np.where(ok, c, cls)
        for (cx, cy, r, z0, z1, c) in self.cyls:
            ox, oy = o[0] - cx, o[1] - cy
            a = dirs[:, 0] ** 2 + dirs[:, 1] ** 2
            b = 2 * (ox * dirs[:, 0] + oy * dirs[:, 1])
            cc = ox * ox + oy * oy - r * r
            disc = b * b - 4 * a * cc
            with np.errstate(divide="ignore", invalid="ignore"):
                s = (-b - np.sqrt(np.maximum(disc, 0))) / (2 * a)
            z = o[2] + s * dirs[:, 2]
            ok = (disc > 0) & (a > 1e-12) & (s > 1e-6) & (z >= z0) & (z <= z1) & (s < best)
            best = np.where(ok, s, best)
            cls = np.where(ok, c, cls)
        miss = ~(best < max_range)
        best = np.where(miss, np.inf, best)
        return best, cls


# ----------------------------------------------------------------------------- C2: KITTI-shaped LiDAR


def kitti_scene(rng, N=20):
    sc = Scene()
    yl, yr = rng.uniform(6, 12), -rng.uniform(6, 12)

    def ground_cls(hit):
        ay = np.abs(hit[:, 1])
        return np.where(ay < 3.5, 1, np.where(ay < 5.5, 2, 3))

    sc.planes.append((np.zeros(3), np.array([0.0, 0.0, 1.0]), None, ground_cls))
    sc.planes.append((np.array([0, yl, 0.0]), np.array([0.0, -1.0, 0.0]), (np.array([-200, yl - 1, 0.0]), np.array([200, yl + 1, 9.0])), 4))
    sc.planes.append((np.array([0, yr, 0.0]), np.array([0.0, 1.0, 0.0]), (np.array([-200, yr - 1, 0.0]), np.array([200, yr + 1, 9.0])), 5))
    for b in range(20):  # cars
        cx, cy = rng.uniform(-45, 45), rng.uniform(yr + 1.5, yl - 1.5)
        if abs(cx) < 4 and abs(cy) < 2.5:
            cx += 8.0
        l, w, h = rng.uniform(3.5, 4.8), rng.uniform(1.6, 2.0), rng.uniform(1.4, 1.9)
        if rng.random() < 0.5:
            l, w = w, l
        sc.boxes.append((np.array([cx - l / 2, cy - w / 2, 0.0]), np.array([cx + l / 2, cy + w / 2, h]), 6 + (b % 5)))
    for c in range(30):  # poles / trunks
        cx = rng.uniform(-50, 50)
        cy = rng.choice([rng.uniform(3.8, yl - 0.5), rng.uniform(yr + 0.5, -3.8)])
        sc.cyls.append((cx, cy, rng.uniform(0.12, 0.4), 0.0, rng.uniform(3, 8), 11 + (c % min(10, max(1, N - 10)))))
    return sc


def velodyne_dirs(n_rings=64, n_az=1875, az_jitter=None):
    elev = np.deg2rad(np.linspace(2.0, -24.8, n_rings))
    az = np.linspace(-np.pi, np.pi, n_az, endpoint=False)
    E, A = np.meshgrid(elev, az, indexing="ij")
    if az_jitter is not None:
        A = A + az_jitter
    d = np.stack([np.cos(E) * np.cos(A), np.cos(E) * np.sin(A), np.sin(E)], axis=-1).reshape(-1, 3)
    return d


def lidar_scan(rng, scene, sensor_R, sensor_t, n_points, max_range=80.0, sigma=0.02, n_rings=64, n_az=1875):
    """Ray-cast one scan in the sensor frame; rays without a hit are dropped, then the scan is
    topped up by re-casting with jittered azimuth until exactly n_points remain."""
    pts, cls = [], []
    have = 0
    first = True
    while have < n_points:
        jit = None if first else rng.uniform(-np.pi, np.pi, size=(n_rings, n_az)) * 1.0
        first = False
        d_s = velodyne_dirs(n_rings, n_az, jit)
        d_w = d_s @ sensor_R.T
        s, c = scene.cast(sensor_t, d_w, max_range)
        ok = np.isfinite(s)
        s = s[ok] + rng.normal(0, sigma, size=int(ok.sum()))
        p = d_s[ok] * s[:, None]
        pts.append(p)
        cls.append(c[ok])
        have += p.shape[0]
    pts = np.concatenate(pts)[:n_points]
    cls = np.concatenate(cls)[:n_points]
    return pts.astype(np.float32), cls.astype(np.int64)


def cached(name, *args, **kw):
    """Disk cache of a generator call (identical output, keyed by name + arguments): synthetic scans take seconds each on
    the host, and bench.py / tools generate the same seeded pairs again and again.  SICP_SYNTH_CACHE=0 disables it."""
    import hashlib
    import os
    import pickle

    root = os.environ.get("SICP_SYNTH_CACHE", "/tmp/sicp_synth_cache")
    fn = globals()[name]
    if root in ("", "0"):
        return fn(*args, **kw)
    key = hashlib.sha1(repr((name, args, sorted(kw.items()))).encode()).hexdigest()[:20]
    path = os.path.join(root, f"{name}_{key}.pkl")
    try:
        with open(path, "rb") as f:
            return pickle.load(f)
    except Exception:
        pass
    out = fn(*args, **kw)
    try:
        os.makedirs(root, exist_ok=True)
        tmp = f"{path}.{os.getpid()}.tmp"
        with open(tmp, "wb") as f:
            pickle.dump(out, f, protocol=4)
        os.replace(tmp, path)
    except Exception:
        pass
    return out


def _yaw(psi):
    c, s = np.cos(psi), np.sin(psi)
    return np.array([[c, -s, 0], [s, c, 0], [0, 0, 1.0]])


def kitti_pair(pair=0, n_points=120_000, N=20, seed_base=2000, n_rings=64, n_az=1875, step=1.0):
    """C2: target scan at the origin, source scan `step` m further along +x with yaw N(0,1deg).
    Returns dict(src_xyz, src_labels, tgt_xyz, tgt_labels, cm, T_gt, init, N)."""
    rng = np.random.default_rng(seed_base + pair)
    sc = kitti_scene(rng, N)
    cm = confusion_matrix(N, 0.8)
    h = 1.73
    Rt, tt = np.eye(3), np.array([0, 0, h])
    psi = np.deg2rad(rng.normal(0, 1.0))
    Rs, ts = _yaw(psi), np.array([step, 0, h])
    txyz, tcls = lidar_scan(rng, sc, Rt, tt, n_points, n_rings=n_rings, n_az=n_az)
    sxyz, scls = lidar_scan(rng, sc, Rs, ts, n_points, n_rings=n_rings, n_az=n_az)
    # p_t = Rt^T (Rs p_s + ts - tt)
    T_gt = pose7(Rt.T @ Rs, Rt.T @ (ts - tt))
    return dict(
        src_xyz=sxyz, src_labels=observe_labels(rng, scls, cm), tgt_xyz=txyz, tgt_labels=observe_labels(rng, tcls, cm),
        cm=cm, T_gt=T_gt, init=identity_pose(), N=N,
    )


def kitti_sequence(n_frames, n_points=120_000, N=20, seed=4000, n_rings=64, n_az=1875):
    """C4: n_frames consecutive scans along a gently curving path through one scene.
    Returns (frames[list of (xyz, labels)], poses[list of (R,t)], cm)."""
    rng = np.random.default_rng(seed)
    sc = kitti_scene(rng, N)
    cm = confusion_matrix(N, 0.8)
    frames, poses = [], []
    x, y, psi = -0.5 * n_frames, 0.0, 0.0
    for f in range(n_frames):
        frng = np.random.default_rng(seed + 1 + f)
        R, t = _yaw(psi), np.array([x, y, 1.73])
        xyz, cls = lidar_scan(frng, sc, R, t, n_points, n_rings=n_rings, n_az=n_az)
        frames.append((xyz, observe_labels(frng, cls, cm)))
        poses.append((R, t))
        psi += np.deg2rad(0.3) * np.sin(f / 40.0)
        x += np.cos(psi) * 1.0
        y += np.sin(psi) * 1.0
    return frames, poses, cm


# ----------------------------------------------------------------------------- C4: long odometry sequence
# kitti_sequence() above keeps ONE 100 m scene, which a 1,000-frame drive leaves after ~80 frames (the rest would be
# ground + parallel facades only: degenerate for any ICP).  The long sequence tiles independently seeded scenes every
# 100 m along +x; a frame sees the tiles within sensor range.  Frames are generated independently (parallel, cacheable).
_TILE = 100.0


def _tile_scene(tile, N, seed):
    sc = kitti_scene(np.random.default_rng(seed + 7919 * (tile + 1000)), N)
    dx = np.array([_TILE * tile, 0.0, 0.0])
    out = Scene()
    out.boxes = [(lo + dx, hi + dx, c) for (lo, hi, c) in sc.boxes]
    out.cyls = [(cx + dx[0], cy, r, z0, z1, c) for (cx, cy, r, z0, z1, c) in sc.cyls]
    return out, sc.planes


def kitti_long_poses(n_frames):
    """Sensor poses (R, t) of the long sequence: 1 m per frame along a lane that weaves +-1.5 m (heading within +-1.5 deg),
    so the drive stays between the facades for any length."""
    poses = []
    for f in range(n_frames):
        x = 1.0 * f
        y = 1.5 * np.sin(x / 60.0)
        psi = np.arctan(1.5 / 60.0 * np.cos(x / 60.0))
        poses.append((_yaw(psi), np.array([x, y, 1.73])))
    return poses


def kitti_long_frame(f, n_frames=1001, n_points=120_000, N=20, seed=4000, n_rings=64, n_az=1875):
    """Frame f of the C4 sequence: dict(xyz, labels, R, t, cm)."""
    R, t = kitti_long_poses(n_frames)[f]
    k0 = int(np.floor(t[0] / _TILE))
    sc = Scene()
    planes = None
    for tile in (k0 - 1, k0, k0 + 1):
        ts, pl = _tile_scene(tile, N, seed)
        sc.boxes += ts.boxes
        sc.cyls += ts.cyls
        if tile == k0:
            planes = pl
    # ground + the two facades of the current tile, the facades shifted to the sensor's x (they span +-200 m)
    sc.planes = [planes[0]] + [(p0 + np.array([_TILE * k0, 0, 0]), n, (b[0] + np.array([_TILE * k0, 0, 0]), b[1] + np.array([_TILE * k0, 0, 0])), c)
                               for (p0, n, b, c) in planes[1:]]
    frng = np.random.default_rng(seed + 1 + f)
    cm = confusion_matrix(N, 0.8)
    xyz, cls = lidar_scan(frng, sc, R, t, n_points, n_rings=n_rings, n_az=n_az)
    return dict(xyz=xyz, labels=observe_labels(frng, cls, cm), R=R, t=t, cm=cm)


def relative_pose(pose_t, pose_s):
    (Rt, tt), (Rs, ts) = pose_t, pose_s
    return pose7(Rt.T @ Rs, Rt.T @ (ts - tt))


# ----------------------------------------------------------------------------- C1: box room, surface samples


def _room_surfaces(rng, size=(10.0, 8.0, 3.0), n_boxes=6):
    """list of (origin, u, v, cls_id): rectangles origin + a*u + b*v, a,b in [0,1]."""
    L, W, H = size
    surf = []
    rects = [
        ((0, 0, 0), (L, 0, 0), (0, W, 0)), ((0, 0, H), (L, 0, 0), (0, W, 0)),
        ((0, 0, 0), (L, 0, 0), (0, 0, H)), ((0, W, 0), (L, 0, 0), (0, 0, H)),
        ((0, 0, 0), (0, W, 0), (0, 0, H)), ((L, 0, 0), (0, W, 0), (0, 0, H)),
    ]
    for r in rects:
        surf.append(tuple(np.array(a, dtype=np.float64) for a in r))
    for _ in range(n_boxes):
        s = rng.uniform(0.4, 1.5, size=3)
        o = np.array([rng.uniform(0.5, L - 2), rng.uniform(0.5, W - 2), 0.0])
        ex, ey, ez = np.array([s[0], 0, 0]), np.array([0, s[1], 0]), np.array([0, 0, s[2]])
        surf += [(o + ez, ex, ey), (o, ex, ez), (o + ey, ex, ez), (o, ey, ez), (o + ex, ey, ez)]
    return surf


def _sample_surfaces(rng, surf, n, sigma):
    areas = np.array([np.linalg.norm(np.cross(u, v)) for (_, u, v) in surf])
    sid = rng.choice(len(surf), size=n, p=areas / areas.sum())
    a, b = rng.random(n), rng.random(n)
    O = np.stack([s[0] for s in surf])[sid]
    U = np.stack([s[1] for s in surf])[sid]
    V = np.stack([s[2] for s in surf])[sid]
    p = O + a[:, None] * U + b[:, None] * V + rng.normal(0, sigma, size=(n, 3))
    return p, sid


def room_pair(seed=100, n_points=10_000, N=11, max_angle_deg=5.0, max_trans=0.3, sigma=0.005):
    """C1 (test_icp-shape): box room 10x8x3 m + 6 boxes, uniform surface samples, sigma=5 mm;
    target = T_gt * independently resampled source surfaces."""
    rng = np.random.default_rng(seed)
    surf = _room_surfaces(rng)
    cm = confusion_matrix(N, 0.85)
    ps, sids = _sample_surfaces(rng, surf, n_points, sigma)
    pt, sidt = _sample_surfaces(rng, surf, n_points, sigma)
    centre = np.array([5.0, 4.0, 1.5])
    ps, pt = ps - centre, pt - centre
    q = quat_from_axis_angle(random_unit(rng), np.deg2rad(rng.uniform(0, max_angle_deg)))
    R = quat_to_R(q)
    t = random_unit(rng) * rng.uniform(0, max_trans)
    pt = pt @ R.T + t
    T_gt = pose7(R, t)
    sl = observe_labels(rng, (sids % N) + 1, cm)
    tl = observe_labels(rng, (sidt % N) + 1, cm)
    return dict(src_xyz=ps.astype(np.float32), src_labels=sl, tgt_xyz=pt.astype(np.float32), tgt_labels=tl, cm=cm,
                T_gt=T_gt, init=identity_pose(), N=N)


# ----------------------------------------------------------------------------- C3: NYU-shaped RGB-D


def nyu_pair(pair=0, width=640, height=480, N=40, seed_base=3000):
    """C3: 640x480 pinhole depth image of a box room + 12 boxes, back-projected; depth noise 1.2mm*z^2."""
    rng = np.random.default_rng(seed_base + pair)
    fx = fy = 518.86 * width / 640.0
    cx, cy = 325.58 * width / 640.0, 253.74 * height / 480.0
    sc = Scene()
    L, W, H = 7.0, 6.0, 3.0
    lo, hi = np.array([-L / 2, -W / 2, 0.0]), np.array([L / 2, W / 2, H])
    walls = [
        (np.array([0, 0, 0.0]), np.array([0, 0, 1.0]), 1), (np.array([0, 0, H]), np.array([0, 0, -1.0]), 2),
        (np.array([lo[0], 0, 0]), np.array([1.0, 0, 0]), 3), (np.array([hi[0], 0, 0]), np.array([-1.0, 0, 0]), 4),
        (np.array([0, lo[1], 0]), np.array([0, 1.0, 0]), 5), (np.array([0, hi[1], 0]), np.array([0, -1.0, 0]), 6),
    ]
    for p0, nrm, c in walls:
        sc.planes.append((p0, nrm, (lo - 1e-6, hi + 1e-6), c))
    for b in range(12):
        s = rng.uniform(0.3, 1.2, size=3)
        o = np.array([rng.uniform(lo[0] + 0.2, hi[0] - 1.4), rng.uniform(lo[1] + 0.2, hi[1] - 1.4), 0.0])
        sc.boxes.append((o, o + s, 7 + (b % max(1, N - 6))))
    cm = confusion_matrix(N, 0.8)
    u, v = np.meshgrid(np.arange(width), np.arange(height))
    d_cam = np.stack([(u - cx) / fx, (v - cy) / fy, np.ones_like(u, dtype=np.float64)], axis=-1).reshape(-1, 3)
    # camera looks along world +x from near one wall; camera frame: z forward, x right, y down
    C0 = np.array([[0, 0, 1.0], [-1.0, 0, 0], [0, -1.0, 0]])  # columns: cam axes in world

    def shoot(Rw, tw):
        d_w = d_cam @ Rw.T
        nrm = np.linalg.norm(d_w, axis=1)
        s, c = sc.cast(tw, d_w / nrm[:, None], 50.0)
        z = s / nrm  # depth along optical axis
        ok = np.isfinite(z) & (z > 0.5) & (z < 8.0)
        z = np.where(ok, z, 8.0)
        c = np.where(ok, c, 1)
        z = z + rng.normal(0, 1.0, size=z.shape) * 0.0012 * z * z
        return (d_cam * z[:, None]).astype(np.float32), c.astype(np.int64)

    Rt, tt = C0, np.array([lo[0] + 0.6, 0.3, 1.4])
    q = quat_from_axis_angle(random_unit(rng), np.deg2rad(rng.uniform(0, 4.0)))
    Rs = C0 @ quat_to_R(q)
    ts = tt + random_unit(rng) * rng.uniform(0, 0.15)
    txyz, tcls = shoot(Rt, tt)
    sxyz, scls = shoot(Rs, ts)
    T_gt = pose7(Rt.T @ Rs, Rt.T @ (ts - tt))
    return dict(src_xyz=sxyz, src_labels=observe_labels(rng, scls, cm), tgt_xyz=txyz, tgt_labels=observe_labels(rng, tcls, cm),
                cm=cm, T_gt=T_gt, init=identity_pose(), N=N)


# ----------------------------------------------------------------------------- C5: initial-pose sweep


def sweep_inits(T_gt, n_inits, pair=0, seed_base=6000, max_angle_deg=15.0, max_trans=3.0):
    """C5: init = T_gt * exp(xi): rotation U(0,15deg) about a uniform axis, translation U(0,3m)."""
    Mgt = pose7_matrix(T_gt)
    out = np.zeros((n_inits, 7))
    for j in range(n_inits):
        rng = np.random.default_rng(seed_base + pair * 4096 + j)
        q = quat_from_axis_angle(random_unit(rng), np.deg2rad(rng.uniform(0, max_angle_deg)))
        D = np.eye(4)
        D[:3, :3] = quat_to_R(q)
        D[:3, 3] = random_unit(rng) * rng.uniform(0, max_trans)
        M = Mgt @ D
        out[j] = pose7(M[:3, :3], M[:3, 3])
    return out
