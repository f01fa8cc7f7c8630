"""Synthetic scenes, poses and "trained-like" field weights with fixed seeds (SURVEY.md
section 8d).  There is no dataset or simulator in this environment, so every measurement and
parity test runs on these: an indoor-like 12.8 m cube with 0.1 m occupancy cells (outer walls +
random axis-aligned boxes), a camera inside it, and a radiance field whose hash features are
U(-1, 1) instead of tcnn's U(-1e-4, 1e-4) -- with the default init the density is ~e^-1
everywhere, every alpha is below alpha_thre and nothing is ever composited (SURVEY.md section 7,
hard part 5).  `density_gain` scales the density row of the base network's output layer so the
density field is strongly heterogeneous (opaque blobs and free space), like a trained scene.
"""
import math

import numpy as np
import torch

ROI_AABB = [-6.4, -0.2, -6.4, 6.4, 12.6, 6.4]


def make_occupancy(resolution=128, n_boxes=48, seed=1, device="cpu"):
    """[1, R, R, R] bool: one-cell outer walls plus random boxes (about 5-10 % occupied)."""
    g = torch.Generator().manual_seed(seed)
    R = resolution
    occ = torch.zeros((R, R, R), dtype=torch.bool)
    occ[0], occ[-1] = True, True
    occ[:, 0], occ[:, -1] = True, True
    occ[:, :, 0], occ[:, :, -1] = True, True
    for _ in range(n_boxes):
        size = torch.randint(3, max(4, R // 6), (3,), generator=g)
        lo = torch.stack([torch.randint(1, R - 1 - int(s), (1,), generator=g)[0] for s in size])
        occ[lo[0]:lo[0] + size[0], lo[1]:lo[1] + size[1], lo[2]:lo[2] + size[2]] = True
    # keep a free corridor around the camera height so poses are never inside a box
    c = R // 2
    occ[c - 6:c + 6, 10:22, c - 6:c + 6] = False
    return occ[None].to(device)


def init_trained_like(field, seed=2, density_gain=6.0):
    """In-place "trained-like" random init of an NGPRadianceField (documented in BASELINE.md)."""
    g = torch.Generator().manual_seed(seed)
    dev = field.mlp_base.params.device
    with torch.no_grad():
        field.to("cpu")
        field.reset_parameters(grid_range=1.0, generator=g)
        # density row = row 0 of the base output matrix [16 x neurons]
        o = sum(a * b for a, b in field._base_dims[:-1])
        n_in = field._base_dims[-1][1]
        field.mlp_base.params[o:o + n_in] *= density_gain
        field.to(dev)
    return field


def quat_xyzw_to_matrix(q):
    """scipy Rotation.from_quat(...).as_matrix() for one xyzw quaternion (habitat_to_data.py:445-451)."""
    x, y, z, w = (float(v) for v in q)
    n = math.sqrt(x * x + y * y + z * z + w * w)
    x, y, z, w = x / n, y / n, z / n, w / n
    return np.array([
        [1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
        [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
        [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)],
    ], dtype=np.float64)


def _poses(x, z, yaw, height):
    poses = np.zeros((len(x), 7))
    poses[:, 0], poses[:, 1], poses[:, 2] = x, height, z
    poses[:, 4] = np.sin(yaw / 2)
    poses[:, 6] = np.cos(yaw / 2)
    return poses


def make_poses(n_views, seed=3, aabb=ROI_AABB, margin=1.0, height=1.5):
    """[n_views, 7] float64 poses (x, y, z, qx, qy, qz, qw), SURVEY.md section 8(d)-3: positions uniform in the
    aabb shrunk by `margin` metres in x and z, at y = `height`; yaw uniform in [0, 2 pi) about +y (planner pose
    format, planning/planning_funcs.py:222-399).  Cameras may start inside an occupied cell: such views march
    samples from the near plane on, as the reference would."""
    rng = np.random.default_rng(seed)
    x = rng.uniform(aabb[0] + margin, aabb[3] - margin, n_views)
    z = rng.uniform(aabb[2] + margin, aabb[5] - margin, n_views)
    yaw = rng.uniform(0, 2 * np.pi, n_views)
    return _poses(x, z, yaw, height)


def make_poses_corridor(n_views, seed=3, half_width=0.5, height=1.5):
    """Poses inside the free corridor make_occupancy() keeps around the scene centre (x, z uniform in
    [-half_width, half_width]): the round-1 distribution, kept for the golden fixtures (tests/golden/) and the
    small-scene tests whose scenario needs a camera in free space."""
    rng = np.random.default_rng(seed)
    x = rng.uniform(-half_width, half_width, n_views)
    z = rng.uniform(-half_width, half_width, n_views)
    yaw = rng.uniform(0, 2 * np.pi, n_views)
    return _poses(x, z, yaw, height)


def pose_to_matrix(pose7):
    m = np.eye(4)
    m[:3, :3] = quat_xyzw_to_matrix(pose7[3:])
    m[:3, 3] = pose7[:3]
    return m
