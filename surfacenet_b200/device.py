"""Device-resident scene constants: the packed image store and the camera matrices
(main_reconstruct.py:49-51 loads them once per scene; here they are uploaded once per scene)."""
import numpy as np
from . import _lib


class DeviceScene:
    """images: reference-style ``models_img`` (list of (H,W,3) uint8 arrays or None, indexed by view
    position, sizes may differ: utils/image.py:80-89); cameraPOs: (V,3,4) float64."""

    def __init__(self, cameraPOs, models_img, views=None):
        torch = _lib.require_cuda()
        cameraPOs = np.ascontiguousarray(cameraPOs, dtype=np.float64)
        if cameraPOs.ndim != 3 or cameraPOs.shape[1:] != (3, 4):
            raise ValueError("cameraPOs must have shape (N_views,3,4), got {}".format(cameraPOs.shape))
        self.n_views = cameraPOs.shape[0]
        if len(models_img) < self.n_views and views is None:
            pass                                   # views beyond the list are simply never referenced
        use = range(min(len(models_img), self.n_views)) if views is None else sorted(set(int(v) for v in views))
        hw = np.zeros((self.n_views, 2), dtype=np.int32)
        off = np.zeros(self.n_views, dtype=np.int64)
        chunks, cur = [], 0
        for v in use:
            img = models_img[v]
            if img is None:
                continue
            img = np.ascontiguousarray(img)
            if img.dtype != np.uint8 or img.ndim != 3 or img.shape[2] < 3:
                raise ValueError("image of view {} must be (H,W,3) uint8, got {} {}".format(v, img.shape, img.dtype))
            if img.shape[2] != 3:
                img = np.ascontiguousarray(img[:, :, :3])
            hw[v] = img.shape[:2]
            off[v] = cur
            chunks.append(img.reshape(-1))
            cur += img.size
        packed = np.concatenate(chunks) if chunks else np.zeros(4, np.uint8)
        self.images = torch.from_numpy(packed).cuda()
        self.img_hw = torch.from_numpy(hw).cuda()
        self.img_offset = torch.from_numpy(off).cuda()
        self.P = torch.from_numpy(cameraPOs).cuda()
        self.loaded_views = set(v for v in use if models_img[v] is not None)
        self.image_bytes = int(packed.size)

    def check_views(self, views):
        views = np.asarray(views)
        if views.size and (views.min() < 0 or views.max() >= self.n_views):
            raise ValueError("view index out of range [0,{})".format(self.n_views))
        missing = set(int(v) for v in np.unique(views)) - self.loaded_views
        if missing:
            raise ValueError("views {} are selected but have no image in models_img".format(sorted(missing)))
