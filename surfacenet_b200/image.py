"""Drop-ins for the patch helpers of utils/image.py used by early rejection: preprocess_patches (9-48), cropImgPatches (92-200, the
pyramidRate == 1 case earlyRejection.py:46 uses) and img_hw_cubesCorner_inScopeCheck (203-221).  "Next" row N3."""
import numpy as np
from . import _lib


def preprocess_patches(patches, mean_BGR):
    """utils/image.py:9-48: (...,h,w,c) RGB -> (...,c,h,w) BGR minus mean_BGR (host; shape bookkeeping only)."""
    patches = np.moveaxis(patches, -1, -3)
    patches = patches[..., ::-1, :, :]
    patches = patches - np.asarray(mean_BGR)[:, None, None]
    return patches


def img_hw_cubesCorner_inScopeCheck(hw_shape, img_h_cubesCorner, img_w_cubesCorner):
    """utils/image.py:203-221 -> (N_cubes,) bool: all 8 projected corners inside the image."""
    img_h, img_w = hw_shape
    return ((np.min(img_h_cubesCorner, axis=1) >= 0) & (np.max(img_h_cubesCorner, axis=1) <= img_h) &
            (np.min(img_w_cubesCorner, axis=1) >= 0) & (np.max(img_w_cubesCorner, axis=1) <= img_w))


def crop_preprocessed_patches_device(img_dev, center_h, center_w, patchSize, mean_BGR):
    """cropImgPatches(pyramidRate=1, cubeCenter_hw=(center_h, center_w)) + .astype(float32) + preprocess_patches in one kernel:
    img_dev cuda (H,W,3) uint8 -> cuda (N,3,patchSize,patchSize) float32 (BGR - mean)."""
    torch = _lib.require_cuda()
    n = int(len(center_h))
    ch = torch.from_numpy(np.ascontiguousarray(center_h, dtype=np.float64)).cuda()
    cw = torch.from_numpy(np.ascontiguousarray(center_w, dtype=np.float64)).cuda()
    mean = torch.from_numpy(np.ascontiguousarray(mean_BGR, dtype=np.float32)).cuda()
    out = torch.empty((n, 3, patchSize, patchSize), dtype=torch.float32, device="cuda")
    _lib.check(_lib.lib.sn_crop_patches(_lib.ptr(img_dev), int(img_dev.shape[0]), int(img_dev.shape[1]), _lib.ptr(ch), _lib.ptr(cw), n,
                                        int(patchSize), _lib.ptr(mean), _lib.ptr(out), _lib.stream_ptr()))
    return out


def cropImgPatches(img, range_h, range_w, patchSize=64, pyramidRate=1.2, interp_order=2, cubeCenter_hw=None):
    """utils/image.py:92-200 for pyramidRate == 1 (one pyramid level, resize rate 1: a clipped gather around the patch centres).
    -> (N_patches, patchSize, patchSize, 3) of img's dtype (uint8)."""
    if pyramidRate != 1:
        raise ValueError("only pyramidRate == 1 (the value earlyRejection.patch2embedding passes) is implemented")
    torch = _lib.require_cuda()
    img = np.ascontiguousarray(img)
    if img.dtype != np.uint8 or img.ndim != 3 or img.shape[2] != 3:
        raise ValueError("img must be (H,W,3) uint8, got {} {}".format(img.shape, img.dtype))
    if cubeCenter_hw is None:
        center_h, center_w = np.mean(range_h, axis=1), np.mean(range_w, axis=1)
    else:
        center_h, center_w = cubeCenter_hw
    out = crop_preprocessed_patches_device(torch.from_numpy(img).cuda(), center_h, center_w, patchSize, np.zeros(3, np.float32))
    return np.ascontiguousarray(out.cpu().numpy()[:, ::-1].transpose(0, 2, 3, 1)).astype(np.uint8)      # BGR (c,h,w) -> RGB (h,w,c)


def readImages(datasetFolder, imgNamePattern, viewList, return_list=True):
    """utils/image.py:50-89: the images of the listed views as (H,W,3) uint8 arrays ('#' -> zero-padded view index, '@' -> plain);
    decoded with PIL (the reference's scipy.misc.imread was a PIL wrapper)."""
    import os
    from PIL import Image
    imgs_list = []
    for viewIndx in viewList:
        imgPath = os.path.join(datasetFolder, imgNamePattern.replace('#', '{:03}'.format(viewIndx)).replace('@', '{}'.format(viewIndx)))
        with Image.open(imgPath) as im:
            imgs_list.append(np.array(im.convert("RGB") if im.mode not in ("RGB", "L") else im))       # a writable copy
        print('loaded img ' + imgPath)
    return imgs_list if return_list else np.stack(imgs_list)
