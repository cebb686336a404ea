"""Seeded synthetic laser workloads for the projective 2D registration path.

Workload shape follows SURVEY.md 8(d): Stage-like rooms (axis-aligned walls, box obstacles,
a few oblique segments, one circle -- in the spirit of
/root/reference/srrg2_laser_slam_2d/apps/synthetic_scene_generator.cpp:36-56), a Hokuyo
UTM-30LX-shaped sensor (1081 beams over 270 deg), range noise N(0, 0.01 m), 2 % dropouts,
normals from a sliding-window line fit (window 0.3 m, >= 5 points: the values of
/root/reference/configurations/stage_segway_double_config_LASER_0.json:711-719) oriented
towards the sensor.  Invalid beams are far points (1e6, 0, 0, 0), which the projector's
range_max gate rejects under any SE(2) pose, so every cloud has exactly n_beams points.

All random draws come from numpy's PCG64 on the host (deterministic for a given seed); the
ray casting runs in float64 torch ops on whatever device is asked for.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch

HOKUYO_BEAMS = 1081
HOKUYO_FOV = 4.71239  # 270 deg
FAR_POINT = (1.0e6, 0.0, 0.0, 0.0)


@dataclass
class ScanPairs:
    """CSR batch of (fixed, moving) clouds; points are float32 rows (x, y, nx, ny)."""

    fixed_pts: np.ndarray      # [n_pairs * n_beams, 4] float32
    fixed_off: np.ndarray      # [n_pairs + 1] int32
    moving_pts: np.ndarray     # [n_pairs * n_beams, 4] float32
    moving_off: np.ndarray     # [n_pairs + 1] int32
    gt_xyt: np.ndarray         # [n_pairs, 3] float32 ground-truth moving_in_fixed
    init_xyt: np.ndarray       # [n_pairs, 3] float32 initial guess

    @property
    def n_pairs(self) -> int:
        return len(self.fixed_off) - 1


def _v2t(xyt: np.ndarray) -> np.ndarray:
    c, s = np.cos(xyt[..., 2]), np.sin(xyt[..., 2])
    T = np.zeros(xyt.shape[:-1] + (3, 3))
    T[..., 0, 0], T[..., 0, 1], T[..., 0, 2] = c, -s, xyt[..., 0]
    T[..., 1, 0], T[..., 1, 1], T[..., 1, 2] = s, c, xyt[..., 1]
    T[..., 2, 2] = 1.0
    return T


def _t2v(T: np.ndarray) -> np.ndarray:
    return np.stack([T[..., 0, 2], T[..., 1, 2], np.arctan2(T[..., 1, 0], T[..., 0, 0])], -1)


def _make_worlds(rng: np.random.Generator, n_worlds: int, n_boxes: int = 6, n_oblique: int = 3):
    """Returns segments [W, S, 4] (ax, ay, bx, by) and circles [W, 1, 3] (cx, cy, r)."""
    segs = []
    half_w = rng.uniform(3.0, 6.0, n_worlds)
    half_h = rng.uniform(3.0, 6.0, n_worlds)

    def rect(cx, cy, hw, hh):
        x0, x1, y0, y1 = cx - hw, cx + hw, cy - hh, cy + hh
        return [np.stack([x0, y0, x1, y0], -1), np.stack([x1, y0, x1, y1], -1),
                np.stack([x1, y1, x0, y1], -1), np.stack([x0, y1, x0, y0], -1)]

    zeros = np.zeros(n_worlds)
    segs += rect(zeros, zeros, half_w, half_h)
    for _ in range(n_boxes):
        # box obstacles hugging the outer part of the room so the centre stays free
        ang = rng.uniform(-math.pi, math.pi, n_worlds)
        rad = rng.uniform(0.55, 0.9, n_worlds)
        cx, cy = rad * half_w * np.cos(ang), rad * half_h * np.sin(ang)
        segs += rect(cx, cy, rng.uniform(0.25, 1.0, n_worlds), rng.uniform(0.25, 1.0, n_worlds))
    for _ in range(n_oblique):
        ang = rng.uniform(-math.pi, math.pi, n_worlds)
        rad = rng.uniform(0.5, 0.95, n_worlds)
        cx, cy = rad * half_w * np.cos(ang), rad * half_h * np.sin(ang)
        d = rng.uniform(-math.pi, math.pi, n_worlds)
        ln = rng.uniform(0.5, 1.5, n_worlds)
        segs.append(np.stack([cx - ln * np.cos(d), cy - ln * np.sin(d),
                              cx + ln * np.cos(d), cy + ln * np.sin(d)], -1))
    ang = rng.uniform(-math.pi, math.pi, n_worlds)
    rad = rng.uniform(0.5, 0.8, n_worlds)
    circles = np.stack([rad * half_w * np.cos(ang), rad * half_h * np.sin(ang),
                        rng.uniform(0.3, 0.8, n_worlds)], -1)[:, None, :]
    return np.stack(segs, 1), circles


def _raycast(origin_pose: torch.Tensor, beam_angles: torch.Tensor, segs: torch.Tensor,
             circles: torch.Tensor) -> torch.Tensor:
    """origin_pose [B,3] (x,y,theta), beam_angles [N], segs [B,S,4], circles [B,K,3] -> ranges [B,N]."""
    a = origin_pose[:, 2:3] + beam_angles[None, :]
    dx, dy = torch.cos(a)[:, :, None], torch.sin(a)[:, :, None]          # [B,N,1]
    ox, oy = origin_pose[:, 0, None, None], origin_pose[:, 1, None, None]  # [B,1,1]
    ax, ay = segs[:, None, :, 0], segs[:, None, :, 1]                      # [B,1,S]
    ex, ey = segs[:, None, :, 2] - ax, segs[:, None, :, 3] - ay
    den = dx * ey - dy * ex
    wx, wy = ax - ox, ay - oy
    t = (wx * ey - wy * ex) / den
    u = (wx * dy - wy * dx) / den
    ok = (den.abs() > 1e-12) & (t > 1e-6) & (u >= 0.0) & (u <= 1.0)
    t = torch.where(ok, t, torch.full_like(t, float("inf")))
    r = t.min(dim=2).values
    fx, fy = ox - circles[:, None, :, 0], oy - circles[:, None, :, 1]
    bq = fx * dx + fy * dy
    cq = fx * fx + fy * fy - circles[:, None, :, 2] ** 2
    disc = bq * bq - cq
    tc = -bq - torch.sqrt(disc.clamp_min(0.0))
    okc = (disc >= 0.0) & (tc > 1e-6)
    tc = torch.where(okc, tc, torch.full_like(tc, float("inf")))
    return torch.minimum(r, tc.min(dim=2).values)


def _normals(pts: torch.Tensor, valid: torch.Tensor, window: int = 12, radius: float = 0.3,
             min_points: int = 5):
    """Sliding-window line-fit normals along the beam axis. pts [B,N,2] float64, valid [B,N]."""
    B, N, _ = pts.shape
    x, y = pts[..., 0], pts[..., 1]
    cnt = torch.zeros_like(x)
    sx, sy, sxx, sxy, syy = (torch.zeros_like(x) for _ in range(5))
    for k in range(-window, window + 1):
        lo, hi = max(0, -k), min(N, N - k)       # i in [lo,hi) pairs with j = i + k
        xi, yi = x[:, lo:hi], y[:, lo:hi]
        xj, yj = x[:, lo + k:hi + k], y[:, lo + k:hi + k]
        m = valid[:, lo:hi] & valid[:, lo + k:hi + k] & ((xj - xi) ** 2 + (yj - yi) ** 2 <= radius * radius)
        m = m.to(x.dtype)
        # accumulate relative coordinates for numerical sanity
        rx, ry = (xj - xi) * m, (yj - yi) * m
        cnt[:, lo:hi] += m
        sx[:, lo:hi] += rx
        sy[:, lo:hi] += ry
        sxx[:, lo:hi] += rx * rx
        sxy[:, lo:hi] += rx * ry
        syy[:, lo:hi] += ry * ry
    n = cnt.clamp_min(1.0)
    mx, my = sx / n, sy / n
    cxx, cxy, cyy = sxx / n - mx * mx, sxy / n - mx * my, syy / n - my * my
    phi = 0.5 * torch.atan2(2.0 * cxy, cxx - cyy)     # principal (tangent) direction
    nx, ny = -torch.sin(phi), torch.cos(phi)
    flip = (nx * x + ny * y) > 0.0                   # orient towards the sensor (origin)
    nx, ny = torch.where(flip, -nx, nx), torch.where(flip, -ny, ny)
    good = valid & (cnt >= min_points)
    return torch.stack([nx, ny], -1), good


def _scan_clouds(pose: torch.Tensor, segs, circles, beam_angles, noise: torch.Tensor,
                 dropout: torch.Tensor, range_max: float):
    r = _raycast(pose, beam_angles, segs, circles) + noise
    valid = torch.isfinite(r) & (r > 0.05) & (r < range_max) & (~dropout)
    r = torch.where(valid, r, torch.zeros_like(r))
    pts = torch.stack([r * torch.cos(beam_angles)[None, :], r * torch.sin(beam_angles)[None, :]], -1)
    nrm, good = _normals(pts, valid)
    cloud = torch.cat([pts, nrm], -1)
    far = torch.tensor(FAR_POINT, dtype=cloud.dtype, device=cloud.device)
    return torch.where(good[..., None], cloud, far.expand_as(cloud))


def make_scan_pairs(n_pairs: int, n_beams: int = HOKUYO_BEAMS, seed: int = 0xC0FFEE,
                    fov: float = HOKUYO_FOV, motion_xy: float = 0.05, motion_theta: float = 0.05,
                    init_noise_xy: float = 0.0, init_noise_theta: float = 0.0,
                    range_noise: float = 0.01, dropout: float = 0.02, range_max: float = 20.0,
                    device: str = "cpu", chunk: int = 128) -> ScanPairs:
    """Tracking-shaped batch: fixed = scan at pose P, moving = scan at pose P*delta, ground truth
    moving_in_fixed = delta with delta uniform in +-motion (apps/synthetic_scene_generator.cpp:168-178).
    init guess = identity when init_noise_* = 0 (tracking), else delta * eps (loop closure)."""
    rng = np.random.default_rng(seed)
    segs_np, circ_np = _make_worlds(rng, n_pairs)
    P = np.stack([rng.uniform(-1.0, 1.0, n_pairs), rng.uniform(-1.0, 1.0, n_pairs),
                  rng.uniform(-math.pi, math.pi, n_pairs)], -1)
    delta = np.stack([rng.uniform(-motion_xy, motion_xy, n_pairs),
                      rng.uniform(-motion_xy, motion_xy, n_pairs),
                      rng.uniform(-motion_theta, motion_theta, n_pairs)], -1)
    eps = np.stack([rng.uniform(-init_noise_xy, init_noise_xy, n_pairs),
                    rng.uniform(-init_noise_xy, init_noise_xy, n_pairs),
                    rng.uniform(-init_noise_theta, init_noise_theta, n_pairs)], -1)
    Pm = _t2v(_v2t(P) @ _v2t(delta))
    if init_noise_xy > 0.0 or init_noise_theta > 0.0:
        init = _t2v(_v2t(delta) @ _v2t(eps))
    else:
        init = np.zeros_like(delta)
    noise = rng.normal(0.0, range_noise, (2, n_pairs, n_beams))
    drop = rng.uniform(0.0, 1.0, (2, n_pairs, n_beams)) < dropout
    beam = np.linspace(-0.5 * fov, 0.5 * fov, n_beams)

    dev = torch.device(device)
    beam_t = torch.from_numpy(beam).to(dev)
    out_f = np.empty((n_pairs, n_beams, 4), np.float32)
    out_m = np.empty((n_pairs, n_beams, 4), np.float32)
    for lo in range(0, n_pairs, chunk):
        hi = min(n_pairs, lo + chunk)
        segs = torch.from_numpy(segs_np[lo:hi]).to(dev)
        circ = torch.from_numpy(circ_np[lo:hi]).to(dev)
        for which, poses, out in ((0, P, out_f), (1, Pm, out_m)):
            cloud = _scan_clouds(torch.from_numpy(poses[lo:hi]).to(dev), segs, circ, beam_t,
                                 torch.from_numpy(noise[which, lo:hi]).to(dev),
                                 torch.from_numpy(drop[which, lo:hi]).to(dev), range_max)
            out[lo:hi] = cloud.to(torch.float32).cpu().numpy()
    off = (np.arange(n_pairs + 1, dtype=np.int64) * n_beams).astype(np.int32)
    return ScanPairs(out_f.reshape(-1, 4), off, out_m.reshape(-1, 4), off.copy(),
                     delta.astype(np.float32), init.astype(np.float32))


@dataclass
class RawScanPairs:
    """Raw LaserMessage-shaped input: ranges of the scan at pose P (fixed) and at pose P*delta (moving)."""

    fixed_ranges: np.ndarray   # [n_pairs, n_beams] float32; no-return / dropped beams read `no_return`
    moving_ranges: np.ndarray  # [n_pairs, n_beams] float32
    angle_min: float           # LaserMessage::angle_min / angle_max for the symmetric sensor matrix [1/res, n/2]
    angle_max: float
    gt_xyt: np.ndarray         # [n_pairs, 3] float32 ground-truth moving_in_fixed
    init_xyt: np.ndarray       # [n_pairs, 3] float32


def make_raw_scans(n_pairs: int, n_beams: int = HOKUYO_BEAMS, seed: int = 0xC0FFEE, fov: float = HOKUYO_FOV,
                   motion_xy: float = 0.05, motion_theta: float = 0.05, range_noise: float = 0.01,
                   dropout: float = 0.02, range_max: float = 20.0, no_return: float = 65.0,
                   device: str = "cpu", chunk: int = 128) -> RawScanPairs:
    """The same worlds / poses / noise model as make_scan_pairs, but stopping at the sensor: float32 ranges, the
    input of RawDataPreprocessorProjective2D.  Beam c looks along (c - n/2) * fov / n, the direction the
    reference's sensor matrix assigns to it (raw_data_preprocessor_projective_2d.cpp:87-90)."""
    rng = np.random.default_rng(seed)
    segs_np, circ_np = _make_worlds(rng, n_pairs)
    P = np.stack([rng.uniform(-1.0, 1.0, n_pairs), rng.uniform(-1.0, 1.0, n_pairs),
                  rng.uniform(-math.pi, math.pi, n_pairs)], -1)
    delta = np.stack([rng.uniform(-motion_xy, motion_xy, n_pairs),
                      rng.uniform(-motion_xy, motion_xy, n_pairs),
                      rng.uniform(-motion_theta, motion_theta, n_pairs)], -1)
    Pm = _t2v(_v2t(P) @ _v2t(delta))
    noise = rng.normal(0.0, range_noise, (2, n_pairs, n_beams))
    drop = rng.uniform(0.0, 1.0, (2, n_pairs, n_beams)) < dropout
    beam = (np.arange(n_beams) - 0.5 * n_beams) * (fov / n_beams)
    dev = torch.device(device)
    beam_t = torch.from_numpy(beam).to(dev)
    out = np.empty((2, n_pairs, n_beams), np.float32)
    for lo in range(0, n_pairs, chunk):
        hi = min(n_pairs, lo + chunk)
        segs = torch.from_numpy(segs_np[lo:hi]).to(dev)
        circ = torch.from_numpy(circ_np[lo:hi]).to(dev)
        for which, poses in ((0, P), (1, Pm)):
            r = _raycast(torch.from_numpy(poses[lo:hi]).to(dev), beam_t, segs, circ)
            r = r + torch.from_numpy(noise[which, lo:hi]).to(dev)
            bad = ~torch.isfinite(r) | (r <= 0.05) | (r >= range_max) | torch.from_numpy(drop[which, lo:hi]).to(dev)
            r = torch.where(bad, torch.full_like(r, no_return), r)
            out[which, lo:hi] = r.to(torch.float32).cpu().numpy()
    return RawScanPairs(out[0], out[1], -0.5 * fov, 0.5 * fov, delta.astype(np.float32),
                        np.zeros_like(delta, dtype=np.float32))


@dataclass
class ScanSequence:
    """One robot driving through one world: raw scans and ground-truth poses (world frame)."""

    ranges: np.ndarray     # [n_frames, n_beams] float32
    poses: np.ndarray      # [n_frames, 3] float64 (x, y, theta) of the sensor = robot
    angle_min: float
    angle_max: float


def make_scan_sequence(n_frames: int, n_beams: int = 721, seed: int = 1, fov: float = HOKUYO_FOV,
                       step_xy: float = 0.04, step_theta: float = 0.02, range_noise: float = 0.005,
                       dropout: float = 0.01, range_max: float = 20.0, no_return: float = 65.0) -> ScanSequence:
    """A smooth random drive inside one seeded room (frame-to-frame motion of the size the reference's generator
    uses, apps/synthetic_scene_generator.cpp:168-178), scanned at every pose: the input of a tracker run."""
    rng = np.random.default_rng(seed)
    segs_np, circ_np = _make_worlds(rng, 1)
    poses = np.zeros((n_frames, 3))
    poses[0] = (rng.uniform(-0.5, 0.5), rng.uniform(-0.5, 0.5), rng.uniform(-math.pi, math.pi))
    vel = np.array([step_xy, 0.0, step_theta]) * rng.uniform(0.5, 1.0, 3)
    for f in range(1, n_frames):
        vel = 0.9 * vel + 0.1 * np.array([step_xy, 0.0, step_theta]) * rng.uniform(-1.0, 1.5, 3)
        poses[f] = _t2v(_v2t(poses[f - 1]) @ _v2t(vel))
    beam = (np.arange(n_beams) - 0.5 * n_beams) * (fov / n_beams)
    segs = torch.from_numpy(np.repeat(segs_np, n_frames, 0))
    circ = torch.from_numpy(np.repeat(circ_np, n_frames, 0))
    r = _raycast(torch.from_numpy(poses), torch.from_numpy(beam), segs, circ)
    r = r + torch.from_numpy(rng.normal(0.0, range_noise, (n_frames, n_beams)))
    bad = ~torch.isfinite(r) | (r <= 0.05) | (r >= range_max) | torch.from_numpy(rng.uniform(0, 1, (n_frames, n_beams)) < dropout)
    r = torch.where(bad, torch.full_like(r, no_return), r)
    return ScanSequence(r.to(torch.float32).numpy(), poses, -0.5 * fov, 0.5 * fov)


def reference_demo_scene(n_points: int = 1024) -> np.ndarray:
    """The deterministic world of apps/synthetic_scene_generator.cpp:36-56: a circle (r = 3.5 m,
    2*n_points points) plus a 2 m + 3 m corner placed at (2, 0, pi/4).  The reference leaves the
    normals to a later NormalComputator; here they are the analytic ones (circle: inward radial;
    corner legs: the leg's left normal), which the projective path needs."""
    i = np.arange(2 * n_points)
    ang = (i * np.float32(2 * math.pi / (2 * n_points))).astype(np.float32)
    circle = np.stack([np.float32(3.5) * np.cos(ang), np.float32(3.5) * np.sin(ang),
                       -np.cos(ang), -np.sin(ang)], -1).astype(np.float32)
    step = np.float32((2.0 + 3.0) / n_points)
    n0 = int(np.float32(2.0) / step)
    n1 = n_points - n0
    leg0 = np.stack([step * np.arange(n0), np.zeros(n0), np.zeros(n0), np.ones(n0)], -1)
    leg1 = np.stack([np.zeros(n1 - 1), -step * np.arange(1, n1), np.ones(n1 - 1), np.zeros(n1 - 1)], -1)
    corner = np.concatenate([leg0, leg1]).astype(np.float32)
    c, s = np.float32(math.cos(math.pi * 0.25)), np.float32(math.sin(math.pi * 0.25))
    x = c * corner[:, 0] - s * corner[:, 1] + np.float32(2.0)
    y = s * corner[:, 0] + c * corner[:, 1]
    nx = c * corner[:, 2] - s * corner[:, 3]
    ny = s * corner[:, 2] + c * corner[:, 3]
    corner = np.stack([x, y, nx, ny], -1).astype(np.float32)
    return np.concatenate([circle, corner]).astype(np.float32)


@dataclass
class MultiSensorPairs:
    """MULTI.json-shaped batch: one robot with several rangefinders.  Slice s aligns the shared moving cloud
    (local-map surrogate, robot frame) onto fixed set s (the scan of sensor s, SENSOR frame)."""

    fixed_pts: list            # per sensor [n_pairs * n_beams, 4] float32
    fixed_off: list            # per sensor [n_pairs + 1] int32
    moving_pts: np.ndarray     # [n_pairs * n_sensors * n_beams, 4] float32 (robot frame of the moving pose)
    moving_off: np.ndarray     # [n_pairs + 1] int32
    sensors: np.ndarray        # [n_sensors, 3] sensor_in_robot
    gt_xyt: np.ndarray         # [n_pairs, 3] ground-truth moving_in_fixed
    init_xyt: np.ndarray       # [n_pairs, 3]
    odom_xyt: np.ndarray       # [n_pairs, 3] odometry's (noisy) prediction of moving_in_fixed: the prior slice's z

    @property
    def n_pairs(self) -> int:
        return len(self.moving_off) - 1


def make_multi_sensor_pairs(n_pairs: int, sensors=((0.2, 0.0, 0.0), (-0.2, 0.0, math.pi)), n_beams: int = 721,
                            seed: int = 0xD0C, fov: float = HOKUYO_FOV, motion_xy: float = 0.05,
                            motion_theta: float = 0.05, odom_noise_xy: float = 0.01, odom_noise_theta: float = 0.01,
                            range_noise: float = 0.01, dropout: float = 0.02, range_max: float = 20.0,
                            device: str = "cpu", chunk: int = 128) -> MultiSensorPairs:
    """Two (or more) rangefinders on one robot (MULTI.json:160-188,372-422): fixed set s = scan of sensor s at robot
    pose P in the sensor frame; moving = the scans of all sensors at robot pose P*delta moved into that robot
    frame (what the scene clipper hands the aligner); ground truth moving_in_fixed = delta; initial guess =
    identity; odom = delta with noise (the odometry prior's measurement)."""
    rng = np.random.default_rng(seed)
    S = np.asarray(sensors, np.float64)
    n_s = len(S)
    segs_np, circ_np = _make_worlds(rng, n_pairs)
    P = np.stack([rng.uniform(-1.0, 1.0, n_pairs), rng.uniform(-1.0, 1.0, n_pairs),
                  rng.uniform(-math.pi, math.pi, n_pairs)], -1)
    delta = np.stack([rng.uniform(-motion_xy, motion_xy, n_pairs), rng.uniform(-motion_xy, motion_xy, n_pairs),
                      rng.uniform(-motion_theta, motion_theta, n_pairs)], -1)
    odom = delta + np.stack([rng.uniform(-odom_noise_xy, odom_noise_xy, n_pairs),
                             rng.uniform(-odom_noise_xy, odom_noise_xy, n_pairs),
                             rng.uniform(-odom_noise_theta, odom_noise_theta, n_pairs)], -1)
    Pm = _t2v(_v2t(P) @ _v2t(delta))
    noise = rng.normal(0.0, range_noise, (2, n_s, n_pairs, n_beams))
    drop = rng.uniform(0.0, 1.0, (2, n_s, n_pairs, n_beams)) < dropout
    beam = np.linspace(-0.5 * fov, 0.5 * fov, n_beams)
    dev = torch.device(device)
    beam_t = torch.from_numpy(beam).to(dev)
    fixed = [np.empty((n_pairs, n_beams, 4), np.float32) for _ in range(n_s)]
    moving = np.empty((n_pairs, n_s, n_beams, 4), np.float32)
    for lo in range(0, n_pairs, chunk):
        hi = min(n_pairs, lo + chunk)
        segs = torch.from_numpy(segs_np[lo:hi]).to(dev)
        circ = torch.from_numpy(circ_np[lo:hi]).to(dev)
        for s in range(n_s):
            for which, poses in ((0, P), (1, Pm)):
                sensor_pose = _t2v(_v2t(poses[lo:hi]) @ _v2t(S[s]))
                cloud = _scan_clouds(torch.from_numpy(sensor_pose).to(dev), segs, circ, beam_t,
                                     torch.from_numpy(noise[which, s, lo:hi]).to(dev),
                                     torch.from_numpy(drop[which, s, lo:hi]).to(dev), range_max).cpu().numpy()
                if which == 0:
                    fixed[s][lo:hi] = cloud.astype(np.float32)
                else:  # sensor frame -> robot frame
                    c, sn = math.cos(S[s, 2]), math.sin(S[s, 2])
                    far = cloud[..., 0] > 1.0e5
                    out = np.empty_like(cloud)
                    out[..., 0] = c * cloud[..., 0] - sn * cloud[..., 1] + S[s, 0]
                    out[..., 1] = sn * cloud[..., 0] + c * cloud[..., 1] + S[s, 1]
                    out[..., 2] = c * cloud[..., 2] - sn * cloud[..., 3]
                    out[..., 3] = sn * cloud[..., 2] + c * cloud[..., 3]
                    out[far] = np.asarray(FAR_POINT)
                    moving[lo:hi, s] = out.astype(np.float32)
    off = (np.arange(n_pairs + 1, dtype=np.int64) * n_beams).astype(np.int32)
    moff = (np.arange(n_pairs + 1, dtype=np.int64) * n_beams * n_s).astype(np.int32)
    return MultiSensorPairs([f.reshape(-1, 4) for f in fixed], [off.copy() for _ in range(n_s)],
                            moving.reshape(-1, 4), moff, S.astype(np.float32), delta.astype(np.float32),
                            np.zeros_like(delta, dtype=np.float32), odom.astype(np.float32))
