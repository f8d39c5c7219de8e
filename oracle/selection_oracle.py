"""Oracle restatement of the "next" row N3: utils/camera.py:275-309 (viewPairAngles_wrt_pts),
utils/viewPairSelection.py:8-82, utils/earlyRejection.py:6-93, utils/image.py:9-48,92-221 and nets/similarityNet.py:23-77
(the VGG-16 patch embedding and the embedding-pair similarity, torch-CPU fp32).  Test infrastructure only.

Pinned by (tests/test_oracle_golden.py): the reference's doctest known answers (camera.py:290-294, viewPairSelection.py:17-31,
image.py:24-34) and golden vectors produced by executing the reference's numpy functions (tests/golden/make_golden_select.py).
similarityNet itself (Theano/Lasagne) cannot run here: PARITY UNPINNED for the network arithmetic, restated from the Lasagne
layer semantics (Conv2DDNNLayer = cross-correlation + bias + ReLU, Pool2DLayer(2) = 2x2 max, DenseLayer = x.W + b).
Mechanical py3 fixes: `patchSize / 2` -> `//`, np.int/np.bool aliases."""
import itertools
import math
import numpy as np


def k_combination_np(iterable, k=2):
    return np.asarray(list(itertools.combinations(iterable, k)))                                   # utils.py:253-256


def viewPairAngles_wrt_pts(cameraTs, pts_xyz):
    """utils/camera.py:275-309."""
    unitize_array = lambda array, axis: array / np.linalg.norm(array, axis=axis, ord=2, keepdims=True)
    calc_arccos = lambda cos_values: np.arccos(np.clip(cos_values, -1.0, 1.0))
    N_views = cameraTs.shape[0]
    vector_pts2cameras = pts_xyz[:, None, :] - cameraTs[None, ...]
    unit_vector_pts2cameras = unitize_array(vector_pts2cameras, axis=-1)
    viewPairs = k_combination_np(range(N_views), k=2)
    viewPairCosine_wrt_pts = np.sum(np.multiply(unit_vector_pts2cameras[:, viewPairs[:, 0]], unit_vector_pts2cameras[:, viewPairs[:, 1]]), axis=-1)
    return calc_arccos(viewPairCosine_wrt_pts)


def argmaxN_viewPairs(viewPairs, w_viewPairs, N_argmax):
    """utils/viewPairSelection.py:8-41 (stable sort: numpy's default leaves the order of equal weights unspecified)."""
    N_validCubes, N_viewPairs = w_viewPairs.shape
    indice_cube, _ = np.indices((N_validCubes, N_argmax))
    indice_N_max = w_viewPairs.argsort(axis=1, kind="stable")[:, -1 * N_argmax:]
    argmaxN_viewPairs = np.repeat(viewPairs[None, ...], N_validCubes, axis=0)[indice_cube, indice_N_max]
    return argmaxN_viewPairs, w_viewPairs[indice_cube, indice_N_max]


def viewPairSelection(cameraTs_np, e_viewPairs, d_viewPairs, validCubes, cubeCenters_xyz, viewPair_relativeImpt_fn, batchSize,
                      N_viewPairs4inference, viewPairs):
    """utils/viewPairSelection.py:44-82."""
    N_cubes, N_viewPairs = d_viewPairs.shape[:2]
    N_validCubes = validCubes.sum()
    D_embedding = e_viewPairs.shape[-1]
    theta_viewPairs = viewPairAngles_wrt_pts(cameraTs=cameraTs_np, pts_xyz=cubeCenters_xyz[validCubes])[..., None]
    d_viewPairs = d_viewPairs[validCubes][..., None]
    w_viewPairs = np.empty((N_validCubes, N_viewPairs), dtype=np.float32)
    per = int(math.floor(float(batchSize) / N_viewPairs))
    for b0 in range(0, N_validCubes, per):
        sl = slice(b0, min(N_validCubes, b0 + per))
        N_batch = sl.stop - sl.start
        _e = e_viewPairs[validCubes][sl][:, viewPairs.flatten()].reshape((N_batch, N_viewPairs, 2 * D_embedding))
        feats = np.concatenate([_e, d_viewPairs[sl], theta_viewPairs[sl]], axis=-1).astype(np.float32).reshape((N_batch * N_viewPairs, 2 * D_embedding + 2))
        w_viewPairs[sl] = viewPair_relativeImpt_fn(feats, n_samples_perGroup=N_viewPairs)
    return argmaxN_viewPairs(viewPairs, w_viewPairs, N_viewPairs4inference)


def preprocess_patches(patches, mean_BGR):
    """utils/image.py:9-48."""
    patches = np.moveaxis(patches, -1, -3)
    patches = patches[..., ::-1, :, :]
    patches = patches - mean_BGR[:, None, None]
    return patches


def cropImgPatches_rate1(img, patchSize, cubeCenter_hw):
    """utils/image.py:92-200 with pyramidRate == 1 (resize rate 1; scipy's order-2 zoom by 1.0 returns the image itself)."""
    center_h, center_w = cubeCenter_hw
    patchSize_r = patchSize // 2
    H, W = img.shape[:2]
    h_min = (center_h * 1.0).astype(np.int64) - patchSize_r
    w_min = (center_w * 1.0).astype(np.int64) - patchSize_r
    rel = np.indices((patchSize, patchSize))
    ph = np.clip(h_min[:, None, None] + rel[0:1], 0, H - 1)
    pw = np.clip(w_min[:, None, None] + rel[1:2], 0, W - 1)
    return img[ph, pw, :]


def img_hw_cubesCorner_inScopeCheck(hw_shape, img_h_cubesCorner, img_w_cubesCorner):
    """utils/image.py:203-221."""
    img_h, img_w = hw_shape
    return ((np.min(img_h_cubesCorner, axis=1) >= 0) & (np.max(img_h_cubesCorner, axis=1) <= img_h) &
            (np.min(img_w_cubesCorner, axis=1) >= 0) & (np.max(img_w_cubesCorner, axis=1) <= img_w))


def patch2embedding(images_list, img_h_cubesCorner, img_w_cubesCorner, patch2embedding_fn, patches_mean_bgr, N_cubes, N_views, D_embedding,
                    patchSize, batchSize, cubeCenter_hw):
    """utils/earlyRejection.py:6-55."""
    inScope = np.zeros((N_cubes, N_views), dtype=bool)
    black = preprocess_patches(np.zeros((1, patchSize, patchSize, 3), dtype=np.float32), mean_BGR=patches_mean_bgr)
    emb = np.zeros((N_cubes, N_views, D_embedding), dtype=np.float32)
    emb[:, :] = patch2embedding_fn(black)[0]
    for _view, _image in enumerate(images_list):
        _in = img_hw_cubesCorner_inScopeCheck(_image.shape[:2], img_h_cubesCorner[_view], img_w_cubesCorner[_view])
        inScope[:, _view] = _in
        n = int(_in.sum())
        if not n:
            continue
        patches = cropImgPatches_rate1(_image, patchSize, cubeCenter_hw[:, _view, _in])
        pre = preprocess_patches(patches.astype(np.float32), mean_BGR=patches_mean_bgr)
        out = np.zeros((n, D_embedding), np.float32)
        for b0 in range(0, n, batchSize):
            out[b0:b0 + batchSize] = patch2embedding_fn(pre[b0:b0 + batchSize])
        emb[_in, _view] = out
    return emb, inScope


def embeddingPairs2simil(embeddings, N_views, embeddingPair2simil_fn, batchSize):
    """utils/earlyRejection.py:58-80 (the (cube, flattened pair) enumeration of utils.yield_batch_ij_npBool is row-major)."""
    viewPairs = k_combination_np(range(N_views), k=2)
    N_cubes = embeddings.shape[0]
    i, j = np.meshgrid(np.arange(N_cubes), viewPairs.flatten(), indexing="ij")
    i, j = i.ravel(), j.ravel()
    outs = []
    for b0 in range(0, i.size, int(batchSize * 2)):
        outs.append(embeddingPair2simil_fn(embeddings[i[b0:b0 + int(batchSize * 2)], j[b0:b0 + int(batchSize * 2)]]))
    return np.vstack(outs).reshape((N_cubes, viewPairs.shape[0]))


def selectFromSimilarity(dissimilarityProb, N_viewPairs4inference):
    """utils/earlyRejection.py:82-93."""
    similarityBool = (dissimilarityProb < 0.5) & (dissimilarityProb > 0.1)
    return (similarityBool.sum(axis=1) >= N_viewPairs4inference).astype(bool)


# ---- similarityNet (nets/similarityNet.py:23-77), torch-CPU fp32 ------------------------------------------------------------
POOL_AFTER = (1, 3, 6, 9, 12)


def patch2embedding_fn(patches, p):
    """similarityNet.py:23-56 on (N,3,64,64) float32 -> (N,128)."""
    import torch
    import torch.nn.functional as F
    with torch.no_grad():
        x = torch.from_numpy(np.ascontiguousarray(patches, dtype=np.float32))
        pools = []
        for l in range(13):
            x = F.relu(F.conv2d(x, torch.from_numpy(p[2 * l]), torch.from_numpy(p[2 * l + 1]), padding=1))
            if l in POOL_AFTER:
                x = F.max_pool2d(x, 2)
                pools.append(x)
        crops = []
        for t in pools[:4]:                                                                      # layers.py:71-76 (r = 1)
            c = t.shape[-1] // 2
            crops.append(t[:, :, c - 1:c + 1, c - 1:c + 1].flatten(1))
        v = torch.cat([pools[4].flatten(1)] + crops, dim=1)                                      # similarityNet.py:47-52
        v = v / (v ** 2).sum(dim=1).sqrt()[:, None]                                              # layers.py:34-38
        return (v @ torch.from_numpy(p[26]) + torch.from_numpy(p[27])).numpy()


def embeddingPair2simil_fn(pairs, p):
    """similarityNet.py:66-77: rows (2m, 2m+1) -> sigmoid(W * ||e1 - e2||_2 + b), (M,1)."""
    e = np.asarray(pairs, np.float32).reshape(-1, 2, pairs.shape[-1])
    d = np.sqrt((np.abs(e[:, 0] - e[:, 1]) ** 2).sum(axis=1, keepdims=True, dtype=np.float32))
    z = d * p[28][0, 0] + p[29][0]
    return (1.0 / (1.0 + np.exp(-z))).astype(np.float32)
