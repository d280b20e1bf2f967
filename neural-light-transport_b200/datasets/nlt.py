"""On-disk NLT scene -> the model's 11-tuple (reference: nlt/datasets/nlt.py:36-184).

Layout written by the reference's data generator / post-processor (data_gen/postproc.py:89-122):
`<data_root>.json` maps every id `{trainvali|test}_{index:09d}_{cam}_{light}` to its relative paths
(`diffuse`, `cvis`, `lvis`, `rgb`, `rgb_camspc`: uint8/uint16 PNG; `uv2cam`: float16 `.npy` of (imh, imw, 2)
texture coordinates in [0, 1]; `nn`: JSON naming the nearest training camera/light) plus a `complete` flag.

Per example (same arithmetic, in float64 like the reference, until the final float32 cast):
  image / dtype-max  ->  bilinear resize (cv2 default) of UV maps to height `uvh` (aspect kept) and of
  camera-space images to (imh, imw)  ->  float32;  the warp is NOT resized (resizing it smears the
  foreground/background edge -- always warp first, then resize);  `test` mode has no ground truth: zeros.
A neighbour that is missing from the status file yields black placeholders and an `incomplete-data_*` id.
"""
import json
import re
from itertools import product
from os.path import exists, join

import cv2
import numpy as np
from PIL import Image

from .base import Dataset as BaseDataset


def _read_uint_image(path):
    """PNG -> ndarray of its stored unsigned type (what `np.array(PIL.Image.open(...))` yields)."""
    with open(path, 'rb') as h:
        img = Image.open(h)
        img.load()
    return np.array(img)


def _unit_range(arr):
    """uint8 / uint16 -> float64 in [0, 1] with the dtype maximum mapped to 1 (xiuminglib img.normalize_uint)."""
    if arr.dtype not in (np.uint8, np.uint16):
        raise TypeError(arr.dtype)
    return arr.astype(float) / np.iinfo(arr.dtype).max


def _fit(arr, new_h=None, new_w=None):
    """cv2 bilinear resize; a missing side follows the aspect ratio, truncated like the reference
    (xiuminglib img.resize: `int(w / h * new_h)`)."""
    h, w = arr.shape[:2]
    if new_h is None and new_w is None:
        raise ValueError("At least one of new height or width must be given")
    if new_h is None:
        new_h = int(h / w * new_w)
    elif new_w is None:
        new_w = int(w / h * new_h)
    return cv2.resize(arr, (new_w, new_h))


class Dataset(BaseDataset):
    def __init__(self, config, mode, **kwargs):
        self.data_root = config.get('DEFAULT', 'data_root')
        status_path = self.data_root.rstrip('/') + '.json'
        if not exists(status_path):
            raise FileNotFoundError(
                "Data status JSON not found at \n\t%s\nRun $REPO/data_gen/postproc.py to generate it" % status_path)
        with open(status_path) as h:
            self.data_paths = json.load(h)
        for paths in self.data_paths.values():        # the JSON stores paths relative to the data root
            for k, v in paths.items():
                if k != 'complete':
                    paths[k] = join(self.data_root, v)
        # uint8_inputs (config key or keyword, default False = the reference's float32 tuple): an 8-bit image that needs
        # no resize is handed over as uint8 and normalised on the GPU -- same values, a quarter of the PCIe bytes
        self.uint8_inputs = bool(kwargs.pop('uint8_inputs', config.getboolean('DEFAULT', 'uint8_inputs', fallback=False)))
        super().__init__(config, mode, **kwargs)
        Image.init()    # PIL's lazy plugin registration is not thread-safe: do it before the worker threads start

    # ---- which ids belong to this mode (nlt/datasets/nlt.py:54-88) ----
    def _glob(self):
        prefix = 'test' if self.mode == 'test' else 'trainvali'
        ids = [i for i, p in self.data_paths.items() if i.startswith(prefix) and p['complete']]
        if self.mode == 'test':
            return ids
        cams = self.config.get('DEFAULT', 'holdout_cam').split(',')
        lights = self.config.get('DEFAULT', 'holdout_light').split(',')
        held_out = {'%s_%s' % cl for cl in product(cams, lights)}
        want_held_out = self.mode == 'vali'
        # id = {prefix}_{index:09d}_{cam}_{light}
        return [i for i in ids if ('_'.join(i.split('_')[-2:]) in held_out) == want_held_out]

    def _get_nn_id(self, nn):
        pattern = re.compile(r'trainvali_\d\d\d\d\d\d\d\d\d_{cam}_{light}'.format(**nn))
        hits = [i for i in self.data_paths if pattern.search(i) is not None]
        if not hits:
            return None
        if len(hits) > 1:
            raise ValueError("Found {n} matches:\n\t{matches}".format(n=len(hits), matches=hits))
        return hits[0]

    # ---- one example (nlt/datasets/nlt.py:120-184) ----
    def _uv_image(self, path, uvh, channels=3):
        img = _read_uint_image(path)
        if channels:
            img = img[:, :, :channels]        # drop alpha
        if self.uint8_inputs and img.dtype == np.uint8 and img.shape[0] == uvh and img.shape[0] == img.shape[1]:
            return np.ascontiguousarray(img)  # lossless: v / 255 happens on the device
        return _fit(_unit_range(img), new_h=uvh)

    def _cam_image(self, path, imh, imw):
        img = _read_uint_image(path)[:, :, :3]
        if self.uint8_inputs and img.dtype == np.uint8 and img.shape[:2] == (imh, imw):
            return np.ascontiguousarray(img)
        return _fit(_unit_range(img), new_h=imh, new_w=imw)

    def _process_example_precache(self, id_):
        if isinstance(id_, bytes):
            id_ = id_.decode()
        paths = self.data_paths[id_]
        cfg = self.config
        imh, imw, uvh = cfg.getint('DEFAULT', 'imh'), cfg.getint('DEFAULT', 'imw'), cfg.getint('DEFAULT', 'uvh')
        base = self._uv_image(paths['diffuse'], uvh)
        cvis = self._uv_image(paths['cvis'], uvh, channels=0)
        lvis = self._uv_image(paths['lvis'], uvh, channels=0)
        if cvis.ndim != 2 or lvis.ndim != 2:
            raise ValueError('visibility maps must be single-channel images: %s' % id_)
        warp = np.load(paths['uv2cam'])
        if self.mode == 'test':
            rgb = np.zeros_like(base)
            rgb_camspc = np.zeros((imh, imw, 3), dtype=base.dtype if base.dtype == np.uint8 else float)
        else:
            rgb = self._uv_image(paths['rgb'], uvh)
            rgb_camspc = self._cam_image(paths['rgb_camspc'], imh, imw)
        with open(paths['nn']) as h:
            nn = json.load(h)
        nn_id = self._get_nn_id(nn)
        if nn_id is None:
            nn_id = 'incomplete-data_{cam}_{light}'.format(**nn)
            nn_base, nn_rgb, nn_rgb_camspc = np.zeros_like(base), np.zeros_like(rgb), np.zeros_like(rgb_camspc)
        else:
            nn_paths = self.data_paths[nn_id]
            nn_base = self._uv_image(nn_paths['diffuse'], uvh)
            nn_rgb = self._uv_image(nn_paths['rgb'], uvh)
            nn_rgb_camspc = self._cam_image(nn_paths['rgb_camspc'], imh, imw)
        f32 = lambda a: a if np.asarray(a).dtype == np.uint8 else np.asarray(a, dtype=np.float32)
        return (id_.encode(), f32(base), f32(cvis)[:, :, None], f32(lvis)[:, :, None], f32(warp), f32(rgb),
                f32(rgb_camspc), nn_id.encode(), f32(nn_base), f32(nn_rgb), f32(nn_rgb_camspc))
