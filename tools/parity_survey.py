#!/usr/bin/env python
"""Parity as a distribution (GPU box): whole-network max-abs error of the exact mode over weight seeds x view-pair counts x cubes,
against (a) the torch-CPU fp32 oracle (the parity criterion) and (b) an fp64 evaluation of the same graph on the GPU (torch
float64), which separates OUR error from the oracle's own fp32 rounding.

    python tools/parity_survey.py --seeds 0 1 2 3 4 --nvp 1 5 8 --D 64 --cubes 2 [--tag name]

Environment switches of the library (SN_WG, SN_WG_RZSCALE, SN_TC_RZCOMP ...) are read once per process: run one process per
setting; the oracle / fp64 results are cached in /tmp between processes of the same gpurun call.  Prints one JSON line per case
and a summary line; `--out` appends them to a file."""
import argparse, json, os, sys, time
import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def truth_forward(X, p, dev, dtype):
    """oracle/surfacenet_oracle.one_viewpair_forward restated on (dev, dtype) -- the same graph, fp64 on the GPU."""
    import torch
    import torch.nn.functional as F
    from oracle.surfacenet_oracle import LAYOUT as L
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(device=dev, dtype=dtype)

    def cb(x, i, act, dil=False):
        W = t(p[i])
        if dil:
            W = W.permute(1, 0, 2, 3, 4).contiguous()
            y = F.conv3d(x, W, padding=2 * (W.shape[-1] // 2), dilation=2)
        else:
            y = F.conv3d(x, W, padding=W.shape[-1] // 2)
        beta, gamma, mean, inv_std = (t(p[i + k]) for k in (1, 2, 3, 4))
        sh = (1, -1, 1, 1, 1)
        y = (y - mean.view(sh)) * (gamma * inv_std).view(sh) + beta.view(sh)
        return torch.relu(y) if act == "relu" else torch.sigmoid(y)

    def up(x, W, f):
        n, c, d, h, w = x.shape
        z = torch.zeros((n, c, d * f, h * f, w * f), dtype=dtype, device=dev)
        z[:, :, ::f, ::f, ::f] = x
        W = t(W)
        return F.conv3d(z.reshape(n * c, 1, d * f, h * f, w * f), W, padding=W.shape[-1] // 2).reshape(n, c, d * f, h * f, w * f)

    x = t(X)
    c13 = cb(cb(cb(x, L["conv1_1"], "relu"), L["conv1_2"], "relu"), L["conv1_3"], "relu")
    s1 = cb(c13, L["side_op1"], "sigmoid")
    c23 = cb(cb(cb(F.max_pool3d(c13, 2, 2), L["conv2_1"], "relu"), L["conv2_2"], "relu"), L["conv2_3"], "relu")
    s2u = up(cb(c23, L["side_op2"], "sigmoid"), p[L["up2_W"]], 2)
    c33 = cb(cb(cb(F.max_pool3d(c23, 2, 2), L["conv3_1"], "relu"), L["conv3_2"], "relu"), L["conv3_3"], "relu")
    s3u = up(cb(c33, L["side_op3"], "sigmoid"), p[L["up3_W"]], 4)
    c43 = cb(cb(cb(c33, L["conv4_1"], "relu", True), L["conv4_2"], "relu", True), L["conv4_3"], "relu", True)
    s4u = up(cb(c43, L["side_op4"], "sigmoid", True), p[L["up4_W"]], 4)
    m2 = cb(cb(torch.cat([s1, s2u, s3u, s4u], dim=1), L["merge_conv"], "relu"), L["merge_conv2"], "relu")
    return cb(m2, L["merge_conv3"], "sigmoid")


def make_inputs(cams, D, n_cubes, n_vp, seed):
    from oracle import cvc_oracle
    from tests import util
    rs = np.random.RandomState(7000 + seed * 31 + n_vp)
    used = [8, 9, 22, 23, 30, 33, 40, 44]
    imgs = util.image_list(49, used)
    pairs = rs.choice(used, size=(n_cubes, n_vp, 2))
    xyz = (np.array([0.0, -40.0, 610.0]) + rs.rand(n_cubes, 3) * 40).astype(np.float32)
    resol = np.full(n_cubes, 0.4, np.float32)
    X = cvc_oracle.gen_coloredCubes(pairs, xyz, resol, cams, imgs, D)
    _, X = cvc_oracle.preprocess_augmentation(None, X, util.MEAN6[None, :, None, None, None], False, False)
    w = (0.1 + rs.rand(n_cubes, n_vp)).astype(np.float32)
    return X, w


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seeds", type=int, nargs="+", default=[0])
    ap.add_argument("--nvp", type=int, nargs="+", default=[2])
    ap.add_argument("--D", type=int, default=64)
    ap.add_argument("--cubes", type=int, default=2)
    ap.add_argument("--mode", default="exact")
    ap.add_argument("--tag", default="")
    ap.add_argument("--out", default="")
    ap.add_argument("--no-oracle", action="store_true", help="skip the torch-CPU fp32 oracle (fp64 truth only)")
    a = ap.parse_args()
    import torch
    from oracle import surfacenet_oracle as so
    from surfacenet_b200 import SurfaceNet, weights
    from tests import util
    cams = util.dtu_cameras()
    rows = []
    for seed in a.seeds:
        params = weights.synthetic_params(seed)
        net = SurfaceNet.Net(params)
        for n_vp in a.nvp:
            X, w = make_inputs(cams, a.D, a.cubes, n_vp, seed)
            cache = "/tmp/parity_ref_s%d_v%d_D%d_c%d.npz" % (seed, n_vp, a.D, a.cubes)
            if os.path.exists(cache):
                z = np.load(cache)
                truth, orc = z["truth"], (z["orc"] if "orc" in z else None)
            else:
                t0 = time.time()
                with torch.no_grad():
                    truth = torch.cat([truth_forward(X[i:i + 1], params, "cuda", torch.float64) for i in range(X.shape[0])], 0).cpu().numpy()
                orc = None
                if not a.no_oracle:
                    with torch.no_grad():
                        orc = torch.cat([so.one_viewpair_forward(X[i:i + 1], params) for i in range(X.shape[0])], 0).numpy()
                    np.savez(cache, truth=truth, orc=orc)
                else:
                    np.savez(cache, truth=truth)
                sys.stderr.write("refs for seed %d n_vp %d: %.1f s\n" % (seed, n_vp, time.time() - t0))
            fused, unf = net.forward(torch.from_numpy(X).cuda(), torch.from_numpy(w).cuda() if n_vp > 1 else None, n_vp, a.mode)
            torch.cuda.synchronize()
            unf = unf.cpu().numpy().reshape(-1, 1, a.D, a.D, a.D).astype(np.float64)
            row = dict(tag=a.tag, seed=seed, n_vp=n_vp, D=a.D, cubes=a.cubes, mode=a.mode,
                       ours_vs_truth=float(np.abs(unf - truth).max()))
            if orc is not None:
                cw = (w / w.sum(1, keepdims=True)).astype(np.float32)
                fused_o = (orc.reshape(a.cubes, n_vp, a.D, a.D, a.D) * cw[:, :, None, None, None]).sum(1, keepdims=True)
                row.update(ours_vs_oracle_unfused=float(np.abs(unf - orc).max()), oracle_vs_truth=float(np.abs(orc - truth).max()),
                           ours_vs_oracle_fused=float(np.abs(fused.cpu().numpy() - fused_o).max()) if n_vp > 1 else float(np.abs(unf - orc).max()))
            rows.append(row)
            print(json.dumps(row), flush=True)
        del net
    keys = [k for k in ("ours_vs_oracle_unfused", "ours_vs_oracle_fused", "ours_vs_truth", "oracle_vs_truth") if k in rows[0]]
    summ = dict(tag=a.tag, summary=True, n=len(rows), **{"max_" + k: max(r[k] for r in rows) for k in keys})
    print(json.dumps(summ), flush=True)
    if a.out:
        with open(a.out, "a") as f:
            for r in rows + [summ]:
                f.write(json.dumps(r) + "\n")


if __name__ == "__main__":
    main()
