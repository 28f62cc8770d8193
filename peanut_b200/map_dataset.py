"""The on-disk map-sequence format either side of the hot path - SURVEY.md section 8(f), N4.

Writer side, ``nav/collect_maps.py:52, 67-87``: during an episode the agent's ``full_map`` is sampled at the steps
``range(25, 525, 25)``, scaled by 255 and truncated to uint8, stacked to ``[20, 4 + num_sem_categories, full_w, full_h]`` and
written with ``np.savez_compressed(..., maps=seq)`` if anything semantic was seen and more than 4000 cell-values of the
explored channel are set.  Reader side, ``prediction/train_prediction_model.py:30-97, 125-150``: ``LoadMapFromFile`` turns
time step ``t_idx`` of a file into the network input (HWC float32 in [0, 1]) and the prediction target (the LAST time step's
goal-category channels where the input's explored channel is still empty); ``SemMapDataset.load_annotations`` enumerates
ten time steps per file.

File I/O is host-side code (numpy); the two byte-level transforms either side of it run on the device when the map lives
there: ``quantize_full_map`` of a CUDA tensor (the device-resident ``full_map`` of ``peanut_b200.map_state.MapState``) is the
``pn_map_quantize`` kernel, so a quarter of the bytes cross to the host, and ``DeviceMapSequence`` keeps an uploaded sequence in
HBM and builds network inputs and targets with ``pn_map_sample`` - both bit-identical to the reference's numpy expressions
(tests/test_map_dataset_gpu.py).  CUDA inputs never fall back to a host path: the library must be there.
"""
import os

import numpy as np

SAVE_STEPS = list(range(25, 525, 25))       # collect_maps.py:52
NUM_TARGET_CATEGORIES = 6                   # train_prediction_model.py:26
TIME_STEPS_PER_FILE = 10                    # train_prediction_model.py:136 ("first 10 timesteps as partial map inputs")


def quantize_full_map(full_map):
    """collect_maps.py:79-80: ``(full_map.cpu().numpy() * 255).astype(np.uint8)`` -> uint8 numpy array [C, W, H].
    ``full_map``: float32 numpy array or torch tensor (any device)."""
    if isinstance(full_map, np.ndarray):
        return (full_map * 255).astype(np.uint8)
    import torch
    if full_map.dtype != torch.float32:
        raise TypeError("full_map must be float32")
    if full_map.is_cuda:
        return quantize_full_map_device(full_map).cpu().numpy()
    return (full_map * 255).to(torch.uint8).numpy()    # fp32 multiply, truncation toward zero: numpy's astype on [0, 256)


def _ctx_for(tensor):
    from . import _lib
    key = tensor.device.index if tensor.device.index is not None else 0
    ctx = _CTX.get(key)
    if ctx is None:
        ctx = _CTX[key] = _lib.Context(key)
    return ctx


_CTX = {}


def quantize_full_map_device(full_map):
    """``pn_map_quantize``: CUDA float32 tensor (any shape) -> CUDA uint8 tensor of the same shape, on the current stream."""
    import torch
    from . import _lib
    if not (full_map.is_cuda and full_map.dtype == torch.float32):
        raise TypeError("quantize_full_map_device takes a CUDA float32 tensor")
    x = full_map.contiguous()
    out = torch.empty(x.shape, dtype=torch.uint8, device=x.device)
    ctx = _ctx_for(x)
    with torch.cuda.device(x.device):
        _lib.check(ctx.lib.pn_map_quantize(ctx.handle, x.data_ptr(), x.numel(), out.data_ptr(),
                                           torch.cuda.current_stream().cuda_stream))
    return out


class DeviceMapSequence:
    """One map-sequence file resident in HBM: ``LoadMapFromFile`` (train_prediction_model.py:47-91) per time step on the device.

    ``maps``: uint8 [T, C, W, H] (numpy array, e.g. ``np.load(path)["maps"]``, or a tensor).  ``sample(t_idx)`` returns the
    reference reader's two arrays as CUDA tensors - ``img`` float32 [W, H, C] and ``gt_semantic_seg`` int64 [W, H, 6] - plus
    ``input`` float32 [C, W, H], the layout ``PEANUT_Prediction_Model.get_prediction`` / ``pn_prednet_forward`` take."""

    def __init__(self, maps, device="cuda:0"):
        import torch
        t = torch.as_tensor(maps)
        if t.dtype != torch.uint8 or t.dim() != 4:
            raise TypeError("maps must be uint8 [T, C, W, H]")
        self.seq = t.to(device).contiguous()
        self.T, self.C, self.W, self.H = (int(v) for v in self.seq.shape)
        self.ctx = _ctx_for(self.seq)

    @classmethod
    def from_file(cls, path, device="cuda:0"):
        maps = np.load(path)
        if path[-1] == "z":
            maps = maps["maps"]
        return cls(maps, device)

    def sample(self, t_idx, hwc=True, chw=True, target=True):
        import torch
        from . import _lib
        dev = self.seq.device
        img = torch.empty((self.W, self.H, self.C), dtype=torch.float32, device=dev) if hwc else None
        inp = torch.empty((self.C, self.W, self.H), dtype=torch.float32, device=dev) if chw else None
        gt = torch.empty((self.W, self.H, NUM_TARGET_CATEGORIES), dtype=torch.int64, device=dev) if target else None
        ptr = lambda t: None if t is None else t.data_ptr()
        with torch.cuda.device(dev):
            _lib.check(self.ctx.lib.pn_map_sample(self.ctx.handle, self.seq.data_ptr(), self.T, self.C, self.W, self.H, int(t_idx), 4,
                                                  NUM_TARGET_CATEGORIES, ptr(img), ptr(inp), ptr(gt),
                                                  torch.cuda.current_stream().cuda_stream))
        return {"img": img, "input": inp, "gt_semantic_seg": gt}


class MapSequenceWriter:
    """One episode's ``full_map_seq`` (collect_maps.py:67-87)."""

    def __init__(self, num_channels, full_w, full_h, save_steps=None):
        self.save_steps = list(SAVE_STEPS if save_steps is None else save_steps)
        self.seq = np.zeros((len(self.save_steps), num_channels, full_w, full_h), dtype=np.uint8)
        self.seq_i = 0

    def record(self, step_i, full_map):
        """Call after every environment step with the 1-based step count; stores the map at the save steps."""
        if step_i in self.save_steps:
            self.seq[self.seq_i] = quantize_full_map(full_map)
            self.seq_i += 1
            return True
        return False

    def should_save(self):
        return bool(np.sum(self.seq[:, 4:]) > 0 and np.sum(self.seq[:, 1]) > 4000)

    def save(self, path):
        """Writes ``path`` (``.npz``, key ``maps``) if the episode qualifies; returns whether it did."""
        if not self.should_save():
            return False
        np.savez_compressed(path, maps=self.seq)
        return True


def load_map_sample(filename, t_idx, img_prefix=None, ori_filename=None):
    """``LoadMapFromFile.__call__`` (train_prediction_model.py:47-91): the fields it adds to the mmseg ``results`` dict."""
    path = os.path.join(img_prefix, filename) if img_prefix is not None else filename
    maps = np.load(path)
    if path[-1] == "z":
        maps = maps["maps"]
    img = maps[t_idx].transpose(1, 2, 0)
    img = img.astype(np.float32) / 255.
    num_channels = img.shape[0]             # sic: the reference sizes the (unused) normalisation vectors by the map height
    mask = (img[:, :, 1] > 0)
    goals = range(4, 4 + NUM_TARGET_CATEGORIES)
    return {
        "filename": path,
        "ori_filename": filename if ori_filename is None else ori_filename,
        "img": img,
        "img_shape": img.shape,
        "ori_shape": img.shape,
        "pad_shape": img.shape,
        "scale_factor": 1.0,
        "img_norm_cfg": dict(mean=np.zeros(num_channels, dtype=np.float32), std=np.ones(num_channels, dtype=np.float32),
                             to_rgb=False),
        "gt_semantic_seg": (maps[-1, goals] * (1 - mask)).transpose(1, 2, 0),
    }


def network_input(sample):
    """The sample's map as the [C, H, W] float32 array ``PEANUT_Prediction_Model.get_prediction`` takes."""
    return np.ascontiguousarray(sample["img"].transpose(2, 0, 1))


def list_samples(img_dir, img_suffix=".npz"):
    """``SemMapDataset.load_annotations`` (train_prediction_model.py:125-143): ten (file, t_idx) entries per file found
    recursively under ``img_dir``, ordered by file name."""
    infos = []
    for root, _, files in os.walk(img_dir):
        for f in files:
            if f.endswith(img_suffix):
                rel = os.path.relpath(os.path.join(root, f), img_dir)
                for t_idx in range(TIME_STEPS_PER_FILE):
                    infos.append(dict(filename=rel, t_idx=t_idx))
    return sorted(infos, key=lambda x: x["filename"])
