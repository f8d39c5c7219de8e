"""Drop-in for utils/sparseCubes.py:9-141 (dense2sparse, append_dense_2sparseList): ray pooling, centre crop,
thresholding and the ordered compaction run on the GPU (sn_raypool_votes, sn_dense2sparse).  "Next" row N1."""
import os
import numpy as np
from . import _lib, rayPooling


def dense2sparse_device(pred16, rgb, votes, D, Dcenter, min_prob, rayPool_thresh=0):
    """pred16 torch.cuda (B,D,D,D) f16; rgb (B,3,D,D,D) u8 | None; votes (B,D,D,D) u8 | None.
    -> dict(counts i32 (B,), offsets i32 (B+1,), ijk u8 (T,3), pred f16 (T,), rgb u8 (T,3)|None, votes u8 (T,)|None) on the
    host, T = total kept voxels; only the compacted entries cross PCIe."""
    torch = _lib.require_cuda()
    B = int(pred16.shape[0])
    cap = B * Dcenter ** 3
    counts = torch.zeros(B, dtype=torch.int32, device="cuda")
    offsets = torch.zeros(B + 1, dtype=torch.int32, device="cuda")
    ijk = torch.empty((cap, 3), dtype=torch.uint8, device="cuda")
    pred_o = torch.empty(cap, dtype=torch.float16, device="cuda")
    rgb_o = torch.empty((cap, 3), dtype=torch.uint8, device="cuda") if rgb is not None else None
    votes_o = torch.empty(cap, dtype=torch.uint8, device="cuda") if votes is not None else None
    need = _lib.lib.sn_dense2sparse_workspace_bytes(B, D, Dcenter)
    if need < 0:
        raise ValueError("dense2sparse: bad sizes (N_cubes={}, D={}, cube_Dcenter={})".format(B, D, Dcenter))
    ws = torch.empty(int(need), dtype=torch.uint8, device="cuda")
    mp = float(np.float16(min_prob))                 # float16 prediction compared with the python float in float16
    _lib.check(_lib.lib.sn_dense2sparse(_lib.ptr(pred16), _lib.ptr(rgb), _lib.ptr(votes), B, D, Dcenter, mp, int(rayPool_thresh),
                                        _lib.ptr(counts), _lib.ptr(offsets), _lib.ptr(ijk), _lib.ptr(pred_o), _lib.ptr(rgb_o),
                                        _lib.ptr(votes_o), cap, _lib.ptr(ws), ws.numel(), _lib.stream_ptr()))
    off = offsets.cpu().numpy()
    T = int(off[-1]) if B else 0
    return dict(counts=counts.cpu().numpy(), offsets=off, ijk=ijk[:T].cpu().numpy(), pred=pred_o[:T].cpu().numpy(),
                rgb=None if rgb_o is None else rgb_o[:T].cpu().numpy(), votes=None if votes_o is None else votes_o[:T].cpu().numpy())


def _split(flat, offsets, idx):
    return [flat[offsets[n]:offsets[n + 1]] for n in idx]


def dense2sparse(prediction, rgb, param, viewPair, min_prob=0.5, rayPool_thresh=0, enable_centerCrop=False, cube_Dcenter=None,
                 enable_rayPooling=False, cameraPOs=None, cameraTs=None):
    """utils/sparseCubes.py:9.
    prediction float16 (N_cubes,D,D,D); rgb uint8 (N_cubes,D,D,D,3); param structured 'ijk'/'xyz'/'resol'; viewPair (N_cubes,N_vp,2).
    -> nonempty_cube_indx, vxl_ijk_list, prediction_list, rgb_list, rayPooling_votes_list, param_new"""
    torch = _lib.require_cuda()
    prediction = np.asarray(prediction)
    if prediction.ndim != 4:
        raise ValueError("prediction must have shape (N_cubes, D, D, D), got {}".format(prediction.shape))
    N_cubes, D = prediction.shape[:2]
    Dc = int(cube_Dcenter) if enable_centerCrop else D
    param_new = np.copy(param)
    if enable_centerCrop:
        param_new['xyz'] += param_new['resol'][:, None] * ((D - Dc) // 2)              # sparseCubes.py:55
    if N_cubes == 0:
        return [], [], [], [], [], param_new
    pred16 = torch.from_numpy(np.ascontiguousarray(prediction.astype(np.float16))).cuda()
    rgb_d = torch.from_numpy(np.ascontiguousarray(np.transpose(np.asarray(rgb).astype(np.uint8), (0, 4, 1, 2, 3)))).cuda()
    votes_d = None
    if enable_rayPooling:
        P = torch.from_numpy(np.ascontiguousarray(cameraPOs, dtype=np.float64)).cuda()
        vp = torch.from_numpy(np.ascontiguousarray(np.asarray(viewPair).astype(np.int32))).cuda()
        xyz = torch.from_numpy(np.ascontiguousarray(param['xyz'], dtype=np.float32)).cuda()
        resol = torch.from_numpy(np.ascontiguousarray(param['resol'], dtype=np.float32)).cuda()
        votes_d = rayPooling.votes_device(pred16, vp, xyz, resol, P, P.shape[0], min_prob)   # sparseCubes.py:60-62
    out = dense2sparse_device(pred16, rgb_d, votes_d, D, Dc, min_prob, rayPool_thresh if enable_rayPooling else 0)
    idx = [n for n in range(N_cubes) if out["counts"][n] > 0]                               # sparseCubes.py:67-68
    off = out["offsets"]
    votes_l = _split(out["votes"], off, idx) if enable_rayPooling else []
    return idx, _split(out["ijk"], off, idx), _split(out["pred"], off, idx), _split(out["rgb"], off, idx), votes_l, param_new


def append_dense_2sparseList(prediction_sub, rgb_sub, param_sub, viewPair_sub, min_prob=0.5, rayPool_thresh=0,
                             enable_centerCrop=False, cube_Dcenter=None, enable_rayPooling=False, cameraPOs=None, cameraTs=None,
                             prediction_list=[], rgb_list=[], vxl_ijk_list=[], rayPooling_votes_list=[],
                             cube_ijk_np=None, param_np=None, viewPair_np=None):
    """Same contract as utils/sparseCubes.py:82-141 (including the mutable default lists; main_reconstruct.py always passes
    them): sparsify one dense batch on the GPU and append the non-empty cubes to the running lists / arrays."""
    pred = np.asarray(prediction_sub)
    if pred.ndim == 5:                                    # (N,1,D,D,D) float32 from the network -> (N,D,D,D) float16
        pred = pred.astype(np.float16)[:, 0]
    rgb_last = np.moveaxis(np.asarray(rgb_sub).astype(np.uint8), 1, -1)             # (N,3,D,D,D) -> (N,D,D,D,3)
    pairs16 = np.asarray(viewPair_sub).astype(np.uint16)
    kept, ijk_l, pred_l, rgb_l, votes_l, param_shifted = dense2sparse(
        pred, rgb_last, param_sub, pairs16, min_prob=min_prob, rayPool_thresh=rayPool_thresh, enable_centerCrop=enable_centerCrop,
        cube_Dcenter=cube_Dcenter, enable_rayPooling=enable_rayPooling, cameraPOs=cameraPOs, cameraTs=cameraTs)
    new_param, new_pairs, new_ijk = param_shifted[kept], pairs16[kept], param_sub['ijk'][kept]
    if len({len(pred_l), len(rgb_l), len(ijk_l), new_param.shape[0], new_pairs.shape[0], new_ijk.shape[0]}) != 1:
        raise Warning('load dense data, # of cubes is not consistent.')
    for dst, src in ((prediction_list, pred_l), (rgb_list, rgb_l), (vxl_ijk_list, ijk_l), (rayPooling_votes_list, votes_l)):
        dst.extend(src)
    stack = lambda old, new, fn: new if old is None else fn([old, new])
    return (prediction_list, rgb_list, vxl_ijk_list, rayPooling_votes_list,
            stack(cube_ijk_np, new_ijk, np.vstack), stack(param_np, new_param, lambda x: np.concatenate(x, axis=0)),
            stack(viewPair_np, new_pairs, np.vstack))


# ---- host-side consumers of the sparse lists: thresholding masks and the on-disk NPZ schema ---------------------------
NPZ_KEYS = ("cube_1st_vxlIndx_np", "prediction_np", "rgb_np", "vxl_ijk_np", "rayPooling_votes_np", "cube_ijk_np", "param_np", "viewPair_np")


def _and_into(masks, new_masks):
    """masks is extended when empty, otherwise and-ed element-wise in place (the reference's accumulate-or-create rule)."""
    if not masks:
        masks.extend(new_masks)
    else:
        for m, n in zip(masks, new_masks):
            m &= n
    return masks


def filter_voxels(vxl_mask_list=[], prediction_list=None, prob_thresh=None, rayPooling_votes_list=None, rayPool_thresh=None):
    """Same contract as utils/sparseCubes.py:205-243: per-cube boolean masks of `prediction >= prob_thresh` (scalar, or one
    threshold per cube when a list is given) and `votes >= rayPool_thresh`, combined with the masks passed in."""
    if prediction_list is not None:
        if prob_thresh is None:
            raise Warning('prob_thresh should not be None.')
        per_cube = prob_thresh if isinstance(prob_thresh, list) else [prob_thresh] * len(prediction_list)
        _and_into(vxl_mask_list, [p >= t for p, t in zip(prediction_list, per_cube)])
    if rayPooling_votes_list is not None:
        if rayPool_thresh is None:
            raise Warning('rayPool_thresh should not be None.')
        _and_into(vxl_mask_list, [v >= rayPool_thresh for v in rayPooling_votes_list])
    return vxl_mask_list


def save2ply(ply_filePath, xyz_np, rgb_np=None, normal_np=None):
    """utils/sparseCubes.py:246-283: binary little-endian PLY with vertex properties x,y,z[,nx,ny,nz][,red,green,blue] (what
    plyfile's PlyData([PlyElement.describe(saved_pts, 'vertex')]).write() produces for the same structured array)."""
    xyz_np = np.asarray(xyz_np)
    N_voxels = xyz_np.shape[0]
    atributes = [('x', '<f4'), ('y', '<f4'), ('z', '<f4')]
    if normal_np is not None:
        atributes += [('nx', '<f4'), ('ny', '<f4'), ('nz', '<f4')]
    if rgb_np is not None:
        atributes += [('red', 'u1'), ('green', 'u1'), ('blue', 'u1')]
    saved_pts = np.zeros(shape=(N_voxels,), dtype=np.dtype(atributes))
    if N_voxels:
        saved_pts['x'], saved_pts['y'], saved_pts['z'] = xyz_np[:, 0], xyz_np[:, 1], xyz_np[:, 2]
        if rgb_np is not None:
            saved_pts['red'], saved_pts['green'], saved_pts['blue'] = rgb_np[:, 0], rgb_np[:, 1], rgb_np[:, 2]
        if normal_np is not None:
            saved_pts['nx'], saved_pts['ny'], saved_pts['nz'] = normal_np[:, 0], normal_np[:, 1], normal_np[:, 2]
    outputFolder = os.path.dirname(ply_filePath)
    if outputFolder and not os.path.exists(outputFolder):
        os.makedirs(outputFolder)
    names = {'<f4': 'float', 'u1': 'uchar'}
    header = ["ply", "format binary_little_endian 1.0", "element vertex {}".format(N_voxels)]
    header += ["property {} {}".format(names[t], n) for n, t in atributes] + ["end_header"]
    with open(ply_filePath, 'wb') as f:
        f.write(("\n".join(header) + "\n").encode("ascii"))
        f.write(saved_pts.tobytes())
    return 1


_PLY_TYPES = {"char": "i1", "int8": "i1", "uchar": "u1", "uint8": "u1", "short": "i2", "int16": "i2", "ushort": "u2", "uint16": "u2",
              "int": "i4", "int32": "i4", "uint": "u4", "uint32": "u4", "float": "f4", "float32": "f4", "double": "f8", "float64": "f8"}


def readPointCloud_xyz(pointCloudFile):
    """utils/scene.py:111-114 (PlyData.read(file)['vertex'] x / y / z -> (N, 3)) without plyfile: the vertex element of an
    ascii / binary_little_endian / binary_big_endian PLY whose properties are scalars, as the initial point clouds of
    main_reconstruct.py:57 are.  The columns keep the file's dtype, as np.c_ of the plyfile columns does."""
    with open(pointCloudFile, "rb") as f:
        raw = f.read()
    end = raw.find(b"end_header")
    if not raw.startswith(b"ply") or end < 0:
        raise ValueError("{} is not a PLY file".format(pointCloudFile))
    end = raw.index(b"\n", end) + 1
    fmt, n_vertex, props, element, before = None, 0, [], None, 0
    for line in raw[:end].decode("ascii", "replace").splitlines():
        tok = line.split()
        if not tok:
            continue
        if tok[0] == "format":
            fmt = tok[1]
        elif tok[0] == "element":
            element = tok[1]
            if element == "vertex":
                n_vertex = int(tok[2])
            elif not props:
                before += int(tok[2])
        elif tok[0] == "property" and element == "vertex":
            if tok[1] == "list":
                raise ValueError("list properties in the vertex element are not supported ({})".format(pointCloudFile))
            props.append((tok[2], _PLY_TYPES[tok[1]]))
    if before:
        raise ValueError("elements before 'vertex' are not supported ({})".format(pointCloudFile))
    names = [n for n, _ in props]
    if not all(a in names for a in "xyz"):
        raise ValueError("the vertex element of {} has no x / y / z".format(pointCloudFile))
    if fmt == "ascii":
        rows = raw[end:].decode("ascii").split("\n")[:n_vertex]
        table = np.array([r.split()[:len(props)] for r in rows], dtype=np.float64).reshape(n_vertex, len(props))
        cols = [table[:, names.index(a)].astype(props[names.index(a)][1]) for a in "xyz"]
    else:
        order = {"binary_little_endian": "<", "binary_big_endian": ">"}.get(fmt)
        if order is None:
            raise ValueError("unknown PLY format {!r} ({})".format(fmt, pointCloudFile))
        v = np.frombuffer(raw[end:], dtype=np.dtype([(n, order + t) for n, t in props]), count=n_vertex)
        cols = [v[a].astype(v[a].dtype.newbyteorder("=")) for a in "xyz"]
    return np.c_[cols[0], cols[1], cols[2]]


def save_sparseCubes_2ply(vxl_mask_list, vxl_ijk_list, rgb_list, param, ply_filePath, normal_list=None):
    """utils/sparseCubes.py:287-327: voxel xyz = ijk * resol + cube xyz (float32) of the masked voxels -> PLY."""
    vxl_mask_np = np.concatenate([np.asarray(m).astype(bool) for m in vxl_mask_list], axis=0)
    vxl_ijk_np = np.vstack(vxl_ijk_list)
    rgb_np = np.vstack(rgb_list)
    if not vxl_mask_np.shape[0] == vxl_ijk_np.shape[0] == rgb_np.shape[0]:
        raise Warning('make sure # of voxels in each cube are consistent.')
    normal_np = None if normal_list is None else np.vstack(normal_list)[vxl_mask_np]
    xyz_list = []
    for _cube, _select in enumerate(vxl_mask_list):
        _select = np.asarray(_select).astype(bool)
        xyz_list.append(vxl_ijk_list[_cube][_select] * param[_cube]['resol'] + param[_cube]['xyz'][None, :])
    xyz_np = np.vstack(xyz_list) if xyz_list else np.zeros((0, 3), np.float32)
    return save2ply(ply_filePath, xyz_np, rgb_np[vxl_mask_np], normal_np)


def save_sparseCubes(filePath, prediction_list, rgb_list, vxl_ijk_list, rayPooling_votes_list, cube_ijk_np, param_np, viewPair_np):
    """Writes the NPZ schema of utils/sparseCubes.py:330-366 (keys NPZ_KEYS; cube_1st_vxlIndx_np = uint32 start offsets, N+1)."""
    starts = np.zeros(cube_ijk_np.shape[0] + 1, dtype=np.uint32)
    sizes = np.fromiter((p.size for p in prediction_list), dtype=np.uint32, count=len(prediction_list))
    starts[1:1 + sizes.size] = np.cumsum(sizes, dtype=np.uint32)
    flat = dict(prediction_np=np.concatenate(prediction_list, axis=0), rgb_np=np.vstack(rgb_list), vxl_ijk_np=np.vstack(vxl_ijk_list),
                rayPooling_votes_np=np.concatenate(rayPooling_votes_list, axis=0) if rayPooling_votes_list else np.empty((0,), np.uint8))
    if not starts[-1] == flat["prediction_np"].shape[0] == flat["rgb_np"].shape[0] == flat["vxl_ijk_np"].shape[0]:
        raise Warning("# of voxels is not consistent while saving sparseCubes.")
    with open(filePath, 'wb') as f:
        np.savez_compressed(f, cube_1st_vxlIndx_np=starts, cube_ijk_np=cube_ijk_np, param_np=param_np, viewPair_np=viewPair_np, **flat)


def load_sparseCubes(filePath):
    """Reads that schema back (utils/sparseCubes.py:369-405)
    -> prediction_list, rgb_list, vxl_ijk_list, rayPooling_votes_list, cube_ijk_np, param_np, viewPair_np"""
    with np.load(filePath) as npz:
        z = {k: npz[k] for k in NPZ_KEYS}
    starts, n_vox = z["cube_1st_vxlIndx_np"].astype(np.int64), int(z["cube_1st_vxlIndx_np"][-1])
    if not n_vox == z["prediction_np"].shape[0] == z["rgb_np"].shape[0] == z["vxl_ijk_np"].shape[0]:
        raise Warning("# of voxels is not consistent while saving sparseCubes.")
    if z["rayPooling_votes_np"].shape[0] not in (0, n_vox):
        raise Warning("rayPooling_votes_np.shape[0] != 0 / # of voxels.")
    n_cube = z["cube_ijk_np"].shape[0]
    cut = lambda a: [a[starts[i]:starts[i + 1]] for i in range(n_cube)]
    return (cut(z["prediction_np"]), cut(z["rgb_np"]), cut(z["vxl_ijk_np"]), cut(z["rayPooling_votes_np"]),
            z["cube_ijk_np"], z["param_np"], z["viewPair_np"])


def lists_from_flat(sparse, param_sub, viewPair_sub, cube_Dcenter, D):
    """Turn the flat output of HotPath.infer_batch_sparse into exactly what append_dense_2sparseList would have appended:
    (prediction_list, rgb_list, vxl_ijk_list, rayPooling_votes_list, cube_ijk_np, param_np, viewPair_np) of the non-empty cubes,
    with the centre-crop shift of the cube origin applied (sparseCubes.py:55)."""
    off = sparse["offsets"]
    idx = [n for n in range(len(sparse["counts"])) if sparse["counts"][n] > 0]
    param_new = np.copy(param_sub)
    param_new['xyz'] += param_new['resol'][:, None] * ((D - cube_Dcenter) // 2)
    return (_split(sparse["pred"], off, idx), _split(sparse["rgb"], off, idx), _split(sparse["ijk"], off, idx),
            _split(sparse["votes"], off, idx), param_sub['ijk'][idx], param_new[idx], np.asarray(viewPair_sub).astype(np.uint16)[idx])
