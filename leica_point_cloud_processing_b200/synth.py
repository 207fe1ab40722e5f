"""Synthetic clouds for the BASELINE.json configurations (there is no dataset access): an "aircraft panel"
surface - a patch of a cylinder with raised stiffeners, stringers and a rivet row so that all six degrees of
freedom are observable - sampled independently for target (CAD-like) and source (scan-like, with range noise and
a known rigid offset), plus FOD blobs for the cloud-difference configuration.  SURVEY.md section 8(d)."""
import numpy as np


def _trapezoid(x, half_top, flank):
    """1 on |x| <= half_top, linear to 0 over `flank`."""
    return np.clip((half_top + flank - np.abs(x)) / flank, 0.0, 1.0)


def panel_height(u, s, length, width):
    """Radial relief (metres) over the base cylinder at axial coordinate u and arc coordinate s."""
    d = np.zeros_like(u)
    # 5 circumferential stiffener ribs, 20 mm x 20 mm profile
    for k in range(5):
        uk = length * (k + 0.5) / 5.0
        d = np.maximum(d, 0.020 * _trapezoid(u - uk, 0.010, 0.004))
    # 3 longitudinal stringers, 15 mm high
    for k in range(3):
        sk = width * (k + 0.5) / 3.0
        d = np.maximum(d, 0.015 * _trapezoid(s - sk, 0.008, 0.004))
    # a row of rivet bumps (10 mm radius caps, 4 mm high) every 100 mm along u
    s0 = width * 0.41
    ur = np.mod(u, 0.1) - 0.05
    r2 = ur * ur + (s - s0) ** 2
    d = np.maximum(d, 0.004 * np.clip(1.0 - r2 / (0.010 ** 2), 0.0, 1.0))
    return d


def panel_points(n, seed, length=4.0, width=2.0, radius=3.0, noise_sigma=0.0):
    """n samples of the panel (uniform in the (u, s) chart), float32 [n, 3]; optional Gaussian range noise
    along the cylinder's radial direction."""
    rng = np.random.default_rng(seed)
    u = rng.random(n) * length
    s = rng.random(n) * width
    d = panel_height(u, s, length, width)
    if noise_sigma > 0:
        d = d + rng.normal(0.0, noise_sigma, n)
    ang = (s - 0.5 * width) / radius
    r = radius + d
    pts = np.stack([u, r * np.sin(ang), r * np.cos(ang) - radius], axis=1)
    return pts.astype(np.float32)


def rigid_about(center, axis, angle_rad, translation):
    """4x4 float64: rotation by angle about `axis` through `center`, then translation."""
    axis = np.asarray(axis, np.float64)
    axis = axis / np.linalg.norm(axis)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    R = np.eye(3) + np.sin(angle_rad) * K + (1 - np.cos(angle_rad)) * (K @ K)
    T = np.eye(4)
    T[:3, :3] = R
    T[:3, 3] = np.asarray(center) - R @ np.asarray(center) + np.asarray(translation)
    return T


def apply_rigid(T, pts):
    """float32 points through a float64 4x4 (generator-side only; not the engine's float transform)."""
    p = pts.astype(np.float64)
    return (p @ T[:3, :3].T + T[:3, 3]).astype(np.float32)


def make_pair(n_source, n_target, length=4.0, width=2.0, radius=3.0, angle_deg=5.0, offset_m=0.02,
              noise_sigma=1e-3, seed_target=1234, seed_source=5678):
    """Config 2 / 3 of BASELINE.json: target = n_target clean samples; source = n_source independent noisy samples
    moved by the inverse of T* (angle_deg about (1,1,1)/sqrt(3) through the centroid + offset_m translation).
    Returns (source, target, T_star) with T_star the transform GICP should recover (source -> target)."""
    target = panel_points(n_target, seed_target, length, width, radius)
    scan = panel_points(n_source, seed_source, length, width, radius, noise_sigma=noise_sigma)
    center = np.array([0.5 * length, 0.0, 0.0])
    t = np.array([2.0, -1.0, 2.0]) / 3.0 * offset_m
    T_star = rigid_about(center, (1.0, 1.0, 1.0), np.deg2rad(angle_deg), t)
    source = apply_rigid(np.linalg.inv(T_star), scan)
    return source, target, T_star


def fuselage_dims(n_points):
    """Panel size for a given point count at ~2.2 mm spacing per sqrt(area / n) (config 3: 10 M -> 12 m x 4 m)."""
    if n_points >= 5_000_000:
        return 12.0, 4.0
    return 4.0, 2.0


def add_fod_blobs(cloud, n_blobs=20, seed=999, length=4.0, width=2.0, radius=3.0, blob_radius=0.015):
    """Config 4: append `n_blobs` spherical FOD blobs (200-2000 points each, centres 10-50 mm above the surface).
    Returns (cloud_with_blobs, is_fod boolean mask)."""
    rng = np.random.default_rng(seed)
    blobs = []
    for _ in range(n_blobs):
        u = rng.random() * length
        s = rng.random() * width
        lift = 0.010 + 0.040 * rng.random()
        d = float(panel_height(np.array([u]), np.array([s]), length, width)[0]) + lift
        ang = (s - 0.5 * width) / radius
        c = np.array([u, (radius + d) * np.sin(ang), (radius + d) * np.cos(ang) - radius])
        m = int(rng.integers(200, 2001))
        v = rng.normal(size=(m, 3))
        v /= np.linalg.norm(v, axis=1, keepdims=True)
        blobs.append(c + blob_radius * v)
    fod = np.concatenate(blobs).astype(np.float32)
    out = np.concatenate([cloud, fod])
    mask = np.zeros(len(out), bool)
    mask[len(cloud):] = True
    return out, mask


def rotation_error_rad(Ta, Tb):
    """Angle of Ra^T Rb, from atan2 of the skew part and the trace (well conditioned near zero, unlike acos)."""
    Ra, Rb = np.asarray(Ta, np.float64)[:3, :3], np.asarray(Tb, np.float64)[:3, :3]
    M = Ra.T @ Rb
    s = 0.5 * np.linalg.norm([M[2, 1] - M[1, 2], M[0, 2] - M[2, 0], M[1, 0] - M[0, 1]])
    c = 0.5 * (np.trace(M) - 1.0)
    return float(np.arctan2(s, c))


def translation_error(Ta, Tb):
    return float(np.linalg.norm(np.asarray(Ta, np.float64)[:3, 3] - np.asarray(Tb, np.float64)[:3, 3]))
