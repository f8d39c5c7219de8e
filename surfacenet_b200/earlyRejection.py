"""Drop-in for utils/earlyRejection.py: patch embeddings of every (cube, view), embedding-pair dissimilarity, cube rejection --
patch cropping, the similarityNet and the selection run on the GPU.  "Next" row N3."""
import numpy as np
from . import _lib, image
from .utils import k_combination_np


def patch2embedding(images_list, img_h_cubesCorner, img_w_cubesCorner, patch2embedding_fn, patches_mean_bgr, N_cubes, N_views, D_embedding,
                    patchSize, batchSize, cubeCenter_hw):
    """utils/earlyRejection.py:6-55 -> patches_embedding (N_cubes, N_views, D_embedding) f32, inScope_cubes_vs_views (N_cubes, N_views) bool.
    Out-of-scope (cube, view) entries hold the embedding of the all-black patch (:30-33)."""
    torch = _lib.require_cuda()
    inScope_cubes_vs_views = np.zeros((N_cubes, N_views), dtype=bool)
    patch_allBlack = image.preprocess_patches(np.zeros((1, patchSize, patchSize, 3), dtype=np.float32), mean_BGR=patches_mean_bgr)
    patches_embedding = np.zeros((N_cubes, N_views, D_embedding), dtype=np.float32)
    patches_embedding[:, :] = patch2embedding_fn(np.ascontiguousarray(patch_allBlack, dtype=np.float32))[0]
    batchSize = max(1, int(batchSize))
    for _view, _image in enumerate(images_list):
        _img_h, _img_w, _img_c = _image.shape
        _inScope = image.img_hw_cubesCorner_inScopeCheck((_img_h, _img_w), img_h_cubesCorner[_view], img_w_cubesCorner[_view])
        inScope_cubes_vs_views[:, _view] = _inScope
        N_in = int(_inScope.sum())
        if not N_in:
            continue
        img_dev = torch.from_numpy(np.ascontiguousarray(_image)).cuda()
        ch, cw = cubeCenter_hw[0][_view][_inScope], cubeCenter_hw[1][_view][_inScope]
        emb = torch.empty((N_in, D_embedding), dtype=torch.float32, device="cuda")
        for b0 in range(0, N_in, batchSize):
            sl = slice(b0, min(N_in, b0 + batchSize))
            patches = image.crop_preprocessed_patches_device(img_dev, ch[sl], cw[sl], patchSize, patches_mean_bgr)
            emb[sl] = patch2embedding_fn(patches)
        patches_embedding[_inScope, _view] = emb.cpu().numpy()
    return patches_embedding, inScope_cubes_vs_views


def embeddingPairs2simil(embeddings, N_views, inScope_cubes_vs_views, embeddingPair2simil_fn, batchSize, viewPairs):
    """utils/earlyRejection.py:58-80 -> dissimilarity (N_cubes, N_viewPairs) float32 (out-of-scope pairs are NOT masked, :77-79)."""
    torch = _lib.require_cuda()
    viewPairs = k_combination_np(range(N_views), k=2)
    N_viewPairs, N_cubes, E = viewPairs.shape[0], embeddings.shape[0], embeddings.shape[-1]
    emb = torch.from_numpy(np.ascontiguousarray(embeddings, dtype=np.float32)).cuda()
    flat = torch.from_numpy(viewPairs.reshape(-1).astype(np.int64)).cuda()
    out = np.empty((N_cubes, N_viewPairs), np.float32)
    cubes_per_batch = max(1, int(batchSize) // max(N_viewPairs, 1))
    for c0 in range(0, N_cubes, cubes_per_batch):
        c1 = min(N_cubes, c0 + cubes_per_batch)
        pairs = emb[c0:c1].index_select(1, flat).reshape(-1, E)                 # rows (2m, 2m+1) = the two views of pair m
        out[c0:c1] = embeddingPair2simil_fn(pairs).reshape(c1 - c0, N_viewPairs).cpu().numpy()
    return out


def selectFromSimilarity(dissimilarityProb, N_viewPairs4inference):
    """utils/earlyRejection.py:82-93 -> (N_cubes,) bool."""
    torch = _lib.require_cuda()
    d = np.ascontiguousarray(dissimilarityProb, dtype=np.float32)
    if d.ndim != 2:
        raise ValueError("dissimilarityProb must have shape (N_cubes, N_viewPairs), got {}".format(d.shape))
    dd = torch.from_numpy(d).cuda()
    out = torch.zeros(d.shape[0], dtype=torch.uint8, device="cuda")
    _lib.check(_lib.lib.sn_select_from_similarity(_lib.ptr(dd), d.shape[0], d.shape[1], int(N_viewPairs4inference), _lib.ptr(out),
                                                  _lib.stream_ptr()))
    return out.cpu().numpy().astype(bool)
