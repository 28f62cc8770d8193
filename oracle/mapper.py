"""Oracle for stage B: ``Semantic_Mapping.forward`` (depth + semantic masks -> ego voxel splat -> height
projections -> ego map -> rotate/translate -> max-fuse into the local map).

TEST INFRASTRUCTURE ONLY.  Restates nav/agent/mapping.py:52-179 and its helpers op by op in float32 torch
(CPU), batched over environments (the reference is hard-wired to one):
  * get_point_cloud_from_z_t      nav/agent/utils/depth_utils.py:129-155
  * transform_camera_view_t       nav/agent/utils/depth_utils.py:158-176  (elevation 0 => identity rotation)
  * transform_pose_t              nav/agent/utils/depth_utils.py:179-195  (shift_loc angle pi/2 => identity)
  * splat_feat_nd                 nav/agent/utils/depth_utils.py:198-252  (round after EACH of the 8 corners)
  * get_grid                      nav/agent/utils/model.py:7-43
Pinned: tests/golden/make_semmap_golden.py runs the unmodified reference module beside this file on the
same seeded inputs and requires bit-equality of all four outputs before writing the fixtures.
"""
import itertools
from types import SimpleNamespace

import numpy as np
import torch
import torch.nn.functional as F


def default_args(**kw):
    """Defaults of nav/arguments.py that shape stage B (lines 44-51, 58-60, 74-84)."""
    a = dict(frame_height=120, frame_width=160, map_resolution=5, map_size_cm=4800, global_downscaling=2,
             vision_range=100, hfov=79.0, du_scale=1, cat_pred_threshold=5.0, exp_pred_threshold=1.0,
             map_pred_threshold=0.1, num_sem_categories=10, camera_height=0.88, device=torch.device('cpu'))
    a.update(kw)
    return SimpleNamespace(**a)


class Geometry:
    """Constants computed in Semantic_Mapping.__init__ (mapping.py:12-50)."""

    def __init__(self, args):
        self.h, self.w = args.frame_height, args.frame_width
        self.res = args.map_resolution
        self.map_cells = (args.map_size_cm // args.global_downscaling) // args.map_resolution  # 480
        self.vr = args.vision_range
        self.max_h = int(360 / self.res)
        self.min_h = int(-40 / self.res)
        self.nz = self.max_h - self.min_h
        self.agent_height = args.camera_height * 100.
        self.shift_x = self.vr * self.res // 2
        self.xc = (self.w - 1.) / 2.
        self.zc = (self.h - 1.) / 2.
        self.f = (self.w / 2.) / np.tan(np.deg2rad(args.hfov / 2.))
        self.nf = 1 + args.num_sem_categories
        self.min_z = int(25 / self.res - self.min_h)
        self.max_z = int((self.agent_height + 1) / self.res - self.min_h)
        self.cat_thr, self.exp_thr, self.map_thr = args.cat_pred_threshold, args.exp_pred_threshold, args.map_pred_threshold
        self.num_sem = args.num_sem_categories


def normalised_coords(depth, g):
    """depth [B,h,w] (cm) -> coords [B,3,h*w] in the splat's [-1,1] convention (mapping.py:57-88)."""
    B = depth.shape[0]
    gx = torch.arange(g.w)[None, None, :].expand(B, g.h, g.w)
    gz = torch.arange(g.h - 1, -1, -1)[None, :, None].expand(B, g.h, g.w)
    X = (gx - g.xc) * depth / g.f
    Z = (gz - g.zc) * depth / g.f
    Y = depth
    Z = Z + g.agent_height          # transform_camera_view_t, identity rotation
    X = X + g.shift_x               # transform_pose_t, identity rotation; Y += 0
    X = X / g.res
    Y = Y / g.res
    X = (X - g.vr // 2.) / g.vr * 2.
    Y = (Y - g.vr // 2.) / g.vr * 2.
    Z = Z / g.res
    Z = (Z - (g.max_h + g.min_h) // 2.) / (g.max_h - g.min_h) * 2.
    return torch.stack((X, Y, Z), 1).reshape(B, 3, g.h * g.w).float()


def stair_mask_(coords, feat):
    """mapping.py:90-97, applied per environment, in place on coords."""
    for b in range(coords.shape[0]):
        zs = coords[b, 2, :]
        my = zs[(zs > -1) & (zs < 1)] * 2 + 1.6
        if len(my) > 0 and torch.quantile(my, 0.03) > 0.2 and torch.sum((my > 0.2) & (my < 0.7)) > 0.2 * len(my):
            below_floor = coords[b, 2, :] * 2 + 1.6 < 0.7
            no_toilet = feat[b, 1 + 4] == 0
            coords[b][:, below_floor & no_toilet] = 99999
    return coords


def splat_feat_nd(grid_dims, feat, coords):
    """depth_utils.py:198-252 on a zero grid; returns [B,F,*grid_dims]."""
    B, Fc, _ = feat.shape
    n_dims = len(grid_dims)
    grid_flat = torch.zeros(B, Fc, int(np.prod(grid_dims)))
    pos_dim, wts_dim = [], []
    for d in range(n_dims):
        pos = coords[:, [d], :] * grid_dims[d] / 2 + grid_dims[d] / 2
        pos_d, wts_d = [], []
        for ix in [0, 1]:
            pos_ix = torch.floor(pos) + ix
            safe_ix = ((pos_ix > 0) & (pos_ix < grid_dims[d])).type(pos.dtype)
            wts_ix = (1 - torch.abs(pos - pos_ix)) * safe_ix
            pos_d.append(pos_ix * safe_ix)
            wts_d.append(wts_ix)
        pos_dim.append(pos_d)
        wts_dim.append(wts_d)
    for ix_d in itertools.product(*[[0, 1]] * n_dims):
        wts = torch.ones_like(wts_dim[0][0])
        index = torch.zeros_like(wts_dim[0][0])
        for d in range(n_dims):
            index = index * grid_dims[d] + pos_dim[d][ix_d[d]]
            wts = wts * wts_dim[d][ix_d[d]]
        grid_flat.scatter_add_(2, index.long().expand(-1, Fc, -1), feat * wts)
        grid_flat = torch.round(grid_flat)
    return grid_flat.view(B, Fc, *grid_dims)


def ego_map(obs, g):
    """obs [B,4+S,h,w] -> (fp_map_pred [B,1,vr,vr], agent_view [B,4+S,cells,cells]) (mapping.py:57-139)."""
    B, c = obs.shape[0], obs.shape[1]
    coords = normalised_coords(obs[:, 3], g)
    feat = torch.ones(B, g.nf, g.h * g.w)
    feat[:, 1:, :] = obs[:, 4:].reshape(B, c - 4, g.h * g.w)
    coords = stair_mask_(coords, feat)
    voxels = splat_feat_nd((g.vr, g.vr, g.nz), feat, coords).transpose(2, 3)
    agent_height_proj = voxels[..., g.min_z:g.max_z].sum(4)
    all_height_proj = voxels.sum(4)
    special = (1 + 5, 1 + 2) if g.num_sem <= 16 else (1 + 3, 1 + 9, 1 + 14)
    for ch in special:
        agent_height_proj[:, ch] = all_height_proj[:, ch]
    fp_map_pred = torch.clamp(agent_height_proj[:, 0:1] / g.map_thr, min=0.0, max=1.0)
    fp_exp_pred = torch.clamp(all_height_proj[:, 0:1] / g.exp_thr, min=0.0, max=1.0)
    n = g.map_cells
    agent_view = torch.zeros(B, c, n, n)
    x1 = n // 2 - g.vr // 2
    y1 = n // 2
    agent_view[:, 0:1, y1:y1 + g.vr, x1:x1 + g.vr] = fp_map_pred
    agent_view[:, 1:2, y1:y1 + g.vr, x1:x1 + g.vr] = fp_exp_pred
    agent_view[:, 4:, y1:y1 + g.vr, x1:x1 + g.vr] = torch.clamp(agent_height_proj[:, 1:] / g.cat_thr, min=0.0, max=1.0)
    return fp_map_pred, agent_view


def new_pose_(pose, rel):
    """get_new_pose_batch, mapping.py:143-160; mutates pose [B,3] in place (x m, y m, theta deg)."""
    pose[:, 1] += rel[:, 0] * torch.sin(pose[:, 2] / 57.29577951308232) + rel[:, 1] * torch.cos(pose[:, 2] / 57.29577951308232)
    pose[:, 0] += rel[:, 0] * torch.cos(pose[:, 2] / 57.29577951308232) - rel[:, 1] * torch.sin(pose[:, 2] / 57.29577951308232)
    pose[:, 2] += rel[:, 2] * 57.29577951308232
    pose[:, 2] = torch.fmod(pose[:, 2] - 180.0, 360.0) + 180.0
    pose[:, 2] = torch.fmod(pose[:, 2] + 180.0, 360.0) - 180.0
    return pose


def get_grid(pose, size):
    """model.py:7-43 (affine_grid default align_corners=False)."""
    x, y, t = pose[:, 0], pose[:, 1], pose[:, 2]
    t = t * np.pi / 180.
    cos_t, sin_t = t.cos(), t.sin()
    zeros, ones = torch.zeros_like(x), torch.ones_like(x)
    theta1 = torch.stack([torch.stack([cos_t, -sin_t, zeros], 1), torch.stack([sin_t, cos_t, zeros], 1)], 1)
    theta2 = torch.stack([torch.stack([ones, -zeros, x], 1), torch.stack([zeros, ones, y], 1)], 1)
    return F.affine_grid(theta1, torch.Size(size), align_corners=False), F.affine_grid(theta2, torch.Size(size), align_corners=False)


def forward(obs, pose_obs, maps_last, poses_last, args=None):
    """Batched Semantic_Mapping.forward: obs [B,4+S,h,w], pose_obs [B,3], maps_last [B,4+S,n,n],
    poses_last [B,3] (mutated in place, as the reference mutates its view of poses_last).
    Returns (fp_map_pred [B,vr,vr], map_pred [B,4+S,n,n], pose_pred [B,3], current_poses [B,3]);
    pose_pred and current_poses alias poses_last like in the reference."""
    g = Geometry(args or default_args())
    with torch.no_grad():
        fp_map_pred, agent_view = ego_map(obs.float(), g)
        current = new_pose_(poses_last, pose_obs)
        st = current.clone()
        half = g.map_cells // 2
        st[:, :2] = - (st[:, :2] * 100.0 / g.res - half) / half
        st[:, 2] = 90. - st[:, 2]
        rot, trans = get_grid(st, agent_view.size())
        rotated = F.grid_sample(agent_view, rot, align_corners=True)
        translated = F.grid_sample(rotated, trans, align_corners=True)
        map_pred = torch.max(maps_last, translated)
    return fp_map_pred[:, 0], map_pred, poses_last, current


# ---------------------------------------------------------------------------------------------------
def synth_obs(seed, args=None, scene="room", sem_density=0.1):
    """Mapper observation [14,120,160] as the reference preprocessing produces it
    (agent_helper.py:183-217): depth in cm = 50 + d*450 with invalid pixels at 45050, semantic channels
    are small non-negative overlap counts.  Scenes: 'room' (floor + walls), 'stairs' (low steps that
    trigger the stair mask), 'wall' (flat wall at 0.6 m: worst-case points per voxel column), 'wall_near' (flat wall in front of
    the minimum depth: every pixel clips to exactly 50 cm, more than 2 048 points per voxel column and coordinates that sit
    exactly on a cell boundary), 'empty'."""
    a = args or default_args()
    rng = np.random.default_rng(seed)
    H, W = a.frame_height, a.frame_width
    g = Geometry(a)
    v = np.arange(H)[:, None].astype(np.float64)
    u = np.arange(W)[None, :].astype(np.float64)
    elev = ((H - 1 - v) - g.zc) / g.f  # tan of the ray elevation
    if scene == "empty":
        d_cm = np.full((H, W), 45050.0)
    else:
        floor_z = {"room": -88.0, "stairs": -60.0, "wall": -88.0, "wall_near": -88.0}[scene]
        wall = {"room": 250.0 + 120.0 * np.sin(u / W * np.pi * rng.uniform(0.5, 2.0)) + rng.uniform(0, 100),
                "stairs": np.full((1, W), 400.0), "wall": np.full((1, W), 60.0), "wall_near": np.full((1, W), 40.0)}[scene]
        with np.errstate(divide="ignore", invalid="ignore"):
            d_floor = np.where(elev < 0, floor_z / elev, np.inf)
        d_cm = np.minimum(d_floor, wall)
        d_cm = d_cm + rng.normal(0, 1.0, (H, W))
        d_cm = np.clip(d_cm, 50.0, 500.0)
        invalid = rng.random((H, W)) < 0.03
        d_cm[invalid] = 45050.0
        d_cm[d_cm >= 495.5] = 45050.0  # "too far" pixels (depth > 0.99 normalised)
    obs = np.zeros((4 + a.num_sem_categories, H, W), np.float32)
    obs[:3] = rng.integers(0, 256, (3, H, W)).astype(np.float32)
    obs[3] = d_cm.astype(np.float32)
    for c in range(a.num_sem_categories - 1):
        if rng.random() < 3 * sem_density:
            h, w = int(rng.integers(8, 50)), int(rng.integers(8, 60))
            y0, x0 = int(rng.integers(0, H - h)), int(rng.integers(0, W - w))
            obs[4 + c, y0:y0 + h, x0:x0 + w] += 1.0
            if rng.random() < 0.3:  # overlapping instances give counts of 2
                obs[4 + c, y0:y0 + h // 2, x0:x0 + w // 2] += 1.0
    return obs


def synth_state(seed, args=None):
    """(pose_delta [3], maps_last [14,480,480], poses_last [3]) as the agent loop would hold them."""
    a = args or default_args()
    rng = np.random.default_rng(seed + 7919)
    n = (a.map_size_cm // a.global_downscaling) // a.map_resolution
    c = 4 + a.num_sem_categories
    maps = np.zeros((c, n, n), np.float32)
    yy, xx = np.mgrid[0:n, 0:n]
    maps[1] = (((yy - n / 2) ** 2 + (xx - n / 2) ** 2) < (n * 0.2) ** 2).astype(np.float32)
    maps[0] = ((rng.random((n, n)) < 0.05) * maps[1]).astype(np.float32)
    for ch in range(4, c):
        if rng.random() < 0.5:
            y0, x0 = int(rng.integers(0, n - 30)), int(rng.integers(0, n - 30))
            maps[ch, y0:y0 + 20, x0:x0 + 25] = rng.integers(1, 6) / 5.0
    pose_delta = np.array([rng.uniform(0, 0.25), 0.0, rng.choice([0.0, np.deg2rad(30), -np.deg2rad(30)])], np.float32)
    half_m = a.map_size_cm / 100.0 / a.global_downscaling / 2.0
    poses = np.array([half_m + rng.uniform(-2, 2), half_m + rng.uniform(-2, 2), rng.uniform(-180, 180)], np.float32)
    return pose_delta, maps, poses
