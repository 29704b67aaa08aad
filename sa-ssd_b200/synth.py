"""Synthetic KITTI-shaped Velodyne clouds (no dataset is available offline).

Ray-cast model of an HDL-64E fixed by SURVEY.md §8d: 64 beams with elevation
``linspace(+2°, −24.8°)``, azimuth step 0.1728°, field of view ±fov about +x
(the reference feeds camera-FOV-cropped ``velodyne_reduced`` clouds,
mmdet/datasets/kitti.py:58), sensor at the origin, ground plane z = −1.73 m,
12 car-sized boxes, two side walls, first hit per ray, 1 cm range noise,
uniform intensity, shuffled, float32.  ``seed`` is the frame index.
"""
import numpy as np

GROUND_Z = -1.73
CAR_SIZE = (1.6, 3.9, 1.56)  # w, l, h (configs/car_cfg.py anchor size)


def _ray_aabb(dirs, lo, hi):
    """Slab test for rays from the origin.  dirs [R,3]; lo/hi [3].
    Returns t of the first hit (inf if none)."""
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = 1.0 / dirs
        t0 = lo[None, :] * inv
        t1 = hi[None, :] * inv
    tmin = np.minimum(t0, t1)
    tmax = np.maximum(t0, t1)
    tmin = np.where(np.isnan(tmin), -np.inf, tmin)
    tmax = np.where(np.isnan(tmax), np.inf, tmax)
    tn = tmin.max(axis=1)
    tf = tmax.min(axis=1)
    hit = (tf >= tn) & (tf > 0)
    t = np.where(tn > 0, tn, tf)
    return np.where(hit, t, np.inf)


def synth_cloud(seed, fov_deg=28.0, az_step_deg=0.1728, n_cars=12):
    """Return points [N,4] float32 (x, y, z, intensity) for frame ``seed``.

    fov_deg=28 gives ≈19.6–20.0 k in-range points (the "~20 k" config);
    fov_deg=45 gives ≈32 k (more than 20 000 voxels → exercises the
    max_voxels truncation); fov_deg=180 ≈ a full sweep (≈120 k).
    """
    rng = np.random.default_rng(int(seed))
    elev = np.deg2rad(np.linspace(2.0, -24.8, 64))
    n_az = int(round(2.0 * fov_deg / az_step_deg))
    az = np.deg2rad(-fov_deg + az_step_deg * np.arange(n_az))
    ce, se = np.cos(elev)[:, None], np.sin(elev)[:, None]
    dirs = np.stack([ce * np.cos(az)[None, :], ce * np.sin(az)[None, :],
                     np.broadcast_to(se, (64, n_az))], axis=-1).reshape(-1, 3)

    t = np.full(dirs.shape[0], np.inf)
    # ground plane
    with np.errstate(divide="ignore"):
        tg = GROUND_Z / dirs[:, 2]
    t = np.minimum(t, np.where((dirs[:, 2] < 0) & (tg > 0), tg, np.inf))
    # cars
    cx = rng.uniform(5.0, 60.0, n_cars)
    cy = rng.uniform(-20.0, 20.0, n_cars)
    swap = rng.random(n_cars) < 0.5
    for i in range(n_cars):
        w, l, h = CAR_SIZE
        dx, dy = (l, w) if swap[i] else (w, l)
        lo = np.array([cx[i] - dx / 2, cy[i] - dy / 2, GROUND_Z])
        hi = np.array([cx[i] + dx / 2, cy[i] + dy / 2, GROUND_Z + h])
        t = np.minimum(t, _ray_aabb(dirs, lo, hi))
    # two walls, 1 m thick, x in [5,70], up to z = 1.5
    wy = rng.uniform(8.0, 15.0, 2)
    for sgn, y0 in zip((1.0, -1.0), wy):
        ylo, yhi = sorted((sgn * y0, sgn * (y0 + 1.0)))
        t = np.minimum(t, _ray_aabb(dirs, np.array([5.0, ylo, GROUND_Z]),
                                    np.array([70.0, yhi, 1.5])))
    keep = np.isfinite(t) & (t <= 120.0)
    t = t[keep] + rng.normal(0.0, 0.01, int(keep.sum()))
    pts = dirs[keep] * t[:, None]
    inten = rng.random(pts.shape[0])
    out = np.concatenate([pts, inten[:, None]], axis=1)
    rng.shuffle(out, axis=0)
    return np.ascontiguousarray(out.astype(np.float32))


def density_sweep_params():
    """(label, fov_deg, az_step_deg) for the 5 k–120 k sweep of BASELINE
    config 5; counts are approximate in-range points."""
    return [("5k", 28.0, 0.6912), ("10k", 28.0, 0.3456), ("20k", 28.0, 0.1728),
            ("40k", 56.0, 0.1728), ("80k", 112.0, 0.1728), ("120k", 180.0, 0.1728)]
