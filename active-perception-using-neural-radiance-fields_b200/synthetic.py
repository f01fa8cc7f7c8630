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


class TrainingSet:
    """Synthetic posed rgb / depth / semantic images and the reference's training-batch sampler (SURVEY.md section
    8(d)-4): `n_images` poses, W x H images with random rgb U8, depth U(0.5, 8) f32, labels randint(0, C) i64
    (seed 4), all resident on `device`.  ``fetch(num_rays)`` mirrors Dataset.fetch_data + preprocess in training
    mode (perception/data_proc/habitat_to_data.py:184-272): ONE random image per batch, `num_rays` random pixels of
    it, pinhole OpenGL rays, random background colour."""

    def __init__(self, n_images=40, width=320, height=240, focal=160.0, n_classes=29, seed=4, device="cpu"):
        g = torch.Generator().manual_seed(seed)
        self.width, self.height, self.size = width, height, n_images
        self.device = torch.device(device)
        self.images = torch.randint(0, 256, (n_images, height, width, 3), generator=g, dtype=torch.uint8).to(self.device)
        self.depths = (torch.rand((n_images, height, width), generator=g) * 7.5 + 0.5).to(self.device)
        self.semantics = torch.randint(0, n_classes, (n_images, height, width), generator=g).to(self.device)
        poses = make_poses(n_images, seed=seed)
        c2w = np.stack([pose_to_matrix(p)[:3] for p in poses]).astype(np.float32)
        self.camtoworlds = torch.from_numpy(c2w).to(self.device)
        self.K = torch.tensor([[focal, 0, width / 2.0], [0, focal, height / 2.0], [0, 0, 1]], device=self.device)
        self.gen = torch.Generator(device=self.device).manual_seed(seed + 1)

    def fetch_from_host(self, host, num_rays):
        """The same batch drawn from HOST copies of the images (pinned memory -> device every step): the end-to-end
        variant of the training benchmark."""
        g = getattr(self, "_host_gen", None)
        if g is None:
            g = self._host_gen = torch.Generator().manual_seed(12345)
        image_id = torch.randint(0, self.size, (1,), generator=g)
        x = torch.randint(0, self.width, (num_rays,), generator=g)
        y = torch.randint(0, self.height, (num_rays,), generator=g)
        pix = host["images"][image_id, y, x].pin_memory().to(self.device, non_blocking=True)
        dep = host["depths"][image_id, y, x].pin_memory().to(self.device, non_blocking=True)
        sem = host["semantics"][image_id, y, x].pin_memory().to(self.device, non_blocking=True)
        return self._batch(image_id.to(self.device), x.to(self.device), y.to(self.device), pix / 255.0, dep, sem)

    def fetch(self, num_rays):
        dev, g = self.device, self.gen
        image_id = torch.randint(0, self.size, (1,), device=dev, generator=g)
        x = torch.randint(0, self.width, (num_rays,), device=dev, generator=g)
        y = torch.randint(0, self.height, (num_rays,), device=dev, generator=g)
        rgb = self.images[image_id, y, x] / 255.0
        dep = self.depths[image_id, y, x]
        sem = self.semantics[image_id, y, x]
        return self._batch(image_id, x, y, rgb, dep, sem)

    def _batch(self, image_id, x, y, rgb, dep, sem):
        from .render import Rays

        dev, g = self.device, self.gen
        num_rays = x.shape[0]
        c2w = self.camtoworlds[image_id]
        camera_dirs = torch.nn.functional.pad(torch.stack([(x - self.K[0, 2] + 0.5) / self.K[0, 0],
                                                           (y - self.K[1, 2] + 0.5) / self.K[1, 1] * -1.0], dim=1),
                                              (0, 1), value=-1.0)
        directions = (camera_dirs[:, None, :] * c2w[:, :3, :3]).sum(dim=-1)
        origins = torch.broadcast_to(c2w[:, :3, -1], directions.shape).contiguous()
        viewdirs = directions / torch.linalg.norm(directions, dim=-1, keepdims=True)
        return {"pixels": rgb.reshape(num_rays, 3), "dep": dep.reshape(num_rays), "sem": sem.reshape(num_rays),
                "rays": Rays(origins=origins, viewdirs=viewdirs), "color_bkgd": torch.rand(3, device=dev, generator=g)}
