#!/usr/bin/env python
"""bench.py -- throughput of the SurfaceNet per-cube inference hot path (CVC gather -> 3D SurfaceNet ->
view-pair fusion -> float16 -> ray-pool votes) on N B200s of one node.

    python bench.py [--gpus N --steps K --warmup W] [--workload c3|c2|c4] [--mode fp32|exact|fast]
    python bench.py --impl reference ...      the CPU restatement of the reference path on the host cores
    python bench.py --workload post ...       "next" row N4: filter + denoise + adapthresh on the scene's sparse cubes
    python bench.py --workload simnet ...     "next" row N3: patch crop + similarityNet embedding (early rejection)

A "step" = one batch of the hot loop of main_reconstruct.py:132-162 per GPU.  Workload (default c3,
BASELINE.json configs[2], the 64^3 configuration the headline target is quoted on; largest single-GPU
config): 16 cubes of 64^3 x 5 view pairs per GPU, weighted fusion + ray pooling, DTU cal18 cameras,
synthetic uint8 1200x1600 images, synthetic calibrated weights.  Weak scaling: every rank owns its own
16 cubes; the per-cube float16 prediction and the votes (u8) volumes -- what the consumer reads, utils/sparseCubes.py:115 --
are all-gathered ONCE per step, inside the timed region; the float16 gather runs under the ray-pool kernels.
At 8 GPUs (or with --c4-shard) the line also carries "c4_shard": one BASELINE configs[3] shard (64 cubes x 8 view pairs per GPU).

value  = fused surface-probability voxels / s, whole job, inputs resident in HBM (device timed, max over ranks); taken from a plain
         pass of K steps -- a second pass of K steps carries the per-launch CUDA events of the conv units for `roofline`
e2e    = same metric through HotPath.infer_batch_host: pinned HOST per-batch arguments in, fused f32 +
         float16 prediction + votes back to pinned host memory, copies inside the timed region.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

FLOP_PER_PAIR_VOXEL = 1358389.0        # 2 x 679,194.5 MAC (conv + 1x1x1 layers)
WORKLOADS = {                          # cubes per GPU, view pairs, cube side
    "c2": dict(cubes=64, n_vp=5, D=32, name="synthetic-image DTU-camera s=32 cubes, 64 cubes x 5 view-pairs per GPU (BASELINE configs[1])"),
    "c3": dict(cubes=16, n_vp=5, D=64, name="synthetic-image DTU-camera s=64 cubes, 16 cubes x 5 weighted view-pairs + rayPooling per GPU (BASELINE configs[2])"),
    "c4": dict(cubes=64, n_vp=8, D=64, name="synthetic 64^3 cubes, 64 cubes x 8 view-pairs per GPU (BASELINE configs[3] = 512 cubes on 8 GPUs)"),
    # whole-scene workload (own GPU arm, run_c5); the reference arm times one cube x 3 pairs of the same size per step
    "c5": dict(cubes=16, n_vp=3, D=64, name="Middlebury dinoSparseRing views 7-12, whole-scene reconstruction, s=64, N_vp=3, cube-sharded (BASELINE configs[4])"),
}


def unit_macs_per_voxel():
    """MAC per full-resolution input voxel for every conv unit."""
    from surfacenet_b200 import weights
    res = {"conv1": 1, "side_op1": 1, "merge": 1, "conv2": 8, "side_op2": 8, "conv3": 64, "side_op3": 64, "conv4": 64, "side_op4": 64}
    macs = []
    for name, kind, cin, cout, k in weights.UNITS:
        if kind == "up":
            macs.append(0.0)
            continue
        key = [p for p in res if name.startswith(p)][0]
        macs.append(cin * cout * k ** 3 / res[key])
    return macs


def peaks():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained") or d["bf16_tflops"], d["hbm_gbs"], "measured (MEASURED_PEAKS.json, bf16 sustained)"
    return 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def window(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        rows = [r for t, r in self.rows if t0 <= t <= t1 and len(r) >= 6] or [r for t, r in self.rows[-3:] if len(r) >= 6]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(float(r[0]) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][1]), "reasons": reasons, "samples": len(rows)}

    def stop(self):
        if self.proc is not None:
            self.proc.kill()


def make_inputs(wl, rank, n_views=49):
    """cube origins uniform in the scan9 BB shrunk by one cube, distinct views per pair, w = rand + 0.1."""
    import numpy as np
    from tests import util
    B, n_vp, D = wl["cubes"], wl["n_vp"], wl["D"]
    rs = np.random.RandomState(100 + rank)
    lo, hi = util.SCAN9_BB[0], util.SCAN9_BB[1] - D * 0.4
    xyz = (lo + rs.rand(B, 3) * (hi - lo)).astype(np.float32)
    resol = np.full(B, 0.4, np.float32)
    pairs = np.stack([np.stack([rs.choice(n_views, 2, replace=False) for _ in range(n_vp)]) for _ in range(B)]).astype(np.int32)
    w = (rs.rand(B, n_vp) + 0.1).astype(np.float32)
    return pairs, xyz, resol, w


def make_images(n_views=49, H=1200, W=1600):
    import numpy as np
    rs = np.random.RandomState(0)
    return [rs.randint(0, 256, size=(H, W, 3), dtype=np.uint8) for _ in range(n_views)]


# ------------------------------------------------------------------------------------------------------
def cpu_reference_step(wl, params, cams, imgs, sample_cubes, threads, return_outputs=False):
    """The reference path on the host: numpy CVC + mean (utils/CVC.py), torch-CPU fp32 network + fusion
    (nets/SurfaceNet.py restated), float16 cast + numpy ray pooling (utils/sparseCubes.py:115,57-62).
    Returns seconds for `sample_cubes` cubes x n_vp pairs."""
    import numpy as np
    import torch
    from oracle import cvc_oracle, raypool_oracle, surfacenet_oracle
    from tests import util
    torch.set_num_threads(threads)
    pairs, xyz, resol, w = make_inputs(dict(wl, cubes=sample_cubes), 0)
    t0 = time.perf_counter()
    X = cvc_oracle.gen_coloredCubes(pairs.astype(np.int64), xyz, resol, cams, imgs, wl["D"])
    _, X = cvc_oracle.preprocess_augmentation(None, X, util.MEAN6[None, :, None, None, None], False, False)
    fused, _ = surfacenet_oracle.nViewPair_SurfaceNet_fn(X, params, w, N_vp=wl["n_vp"], chunk=1)
    votes = raypool_oracle.votes_batch(fused, pairs, xyz, resol, cams, 0.46)
    dt = time.perf_counter() - t0
    if return_outputs:
        return dt, dict(pairs=pairs, xyz=xyz, resol=resol, w=w, fused=fused, votes=votes)
    return dt


def run_reference(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # torchrun pins OMP_NUM_THREADS=1 for its children; the CPU arm is meant to use every host core
    for k in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS"):
        os.environ[k] = str(os.cpu_count() or 1)
    import numpy as np  # noqa: F401
    from surfacenet_b200 import weights
    from tests import util
    cams = util.dtu_cameras()
    imgs = make_images()
    params = weights.synthetic_params(0)
    threads = os.cpu_count() or 1
    sample = 1
    V = wl["D"] ** 3
    for _ in range(args.warmup):
        cpu_reference_step(wl, params, cams, imgs, sample, threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_reference_step(wl, params, cams, imgs, sample, threads)
    dt = (time.perf_counter() - t0) / args.steps
    val = sample * V / dt
    desc = "%d cube x %d view-pairs of %d^3 per step (CVC numpy + torch-CPU fp32 net + numpy ray pooling)" % (sample, wl["n_vp"], wl["D"])
    print(json.dumps({
        "impl": "reference", "metric": "fused surface-probability voxels/sec (CVC + SurfaceNet fwd + fusion + rayPooling)", "value": val,
        "unit": "voxels/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["name"], "sample": desc},
        "cpu_baseline": {"value": val, "unit": "voxels/s", "cores": threads, "kind": "port", "sample": desc},
        "e2e": {"value": val, "unit": "voxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "pair_voxels_per_s": val * wl["n_vp"], "gpu_launches": 0}))


# ------------------------------------------------------------------------------------------------------
def run_gpu(args, wl):
    import ctypes as C
    import numpy as np
    import torch
    import torch.distributed as dist
    from surfacenet_b200 import SurfaceNet, _lib, pipeline, weights
    from surfacenet_b200.device import DeviceScene
    from tests import util

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    mode = args.mode or _lib.DEFAULT_MODE
    B, n_vp, D = wl["cubes"], wl["n_vp"], wl["D"]
    V = D ** 3

    cams = util.dtu_cameras()
    imgs = make_images()
    scene = DeviceScene(cams, imgs)
    params = weights.synthetic_params(0)
    net = SurfaceNet.Net(params)
    hp = pipeline.HotPath(net, scene, mode=mode)
    pairs, xyz, resol, w = make_inputs(wl, rank)
    d_pairs, d_xyz, d_resol, d_w = (torch.from_numpy(a).to(dev) for a in (pairs, xyz, resol, w))
    from surfacenet_b200 import rayPooling
    # multi-GPU: what the consumer of the per-cube volumes reads is the float16 prediction (utils/sparseCubes.py:115) and the votes: gather
    # THOSE (3 B / voxel instead of 5).  The float16 gather is issued right after the cast and runs on NCCL's stream under the ray-pool kernels.
    gathered_p = torch.empty((world * B, D, D, D), dtype=torch.float16, device=dev) if world > 1 else None
    gathered_v = torch.empty((world * B, D, D, D), dtype=torch.uint8, device=dev) if world > 1 else None
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)          # > 126 MB L2
    rp_ws = torch.empty(int(_lib.lib.sn_raypool_workspace_bytes(wl["cubes"], n_vp, D)), dtype=torch.uint8, device=dev)

    def device_step(d_pairs, d_xyz, d_resol, d_w, g_p, g_v):
        if world == 1:
            return hp.infer_batch(d_pairs, d_xyz, d_resol, d_w, D, want_unfused=False, ray_pool=True)
        out = hp.infer_batch(d_pairs, d_xyz, d_resol, d_w, D, want_unfused=False, ray_pool=False)
        h1 = dist.all_gather_into_tensor(g_p, out["pred16"], async_op=True)
        out["votes"] = rayPooling.votes_device(out["pred16"], d_pairs, d_xyz, d_resol, scene.P, scene.n_views, hp.min_prob, workspace=rp_ws)
        h2 = dist.all_gather_into_tensor(g_v, out["votes"], async_op=True)
        h1.wait(); h2.wait()
        return out

    def step_device():
        return device_step(d_pairs, d_xyz, d_resol, d_w, gathered_p, gathered_v)

    def step_host():
        # the host-buffer entry returns THIS rank's cubes in host memory; the scene-level exchange of the (sparse) results happens once per
        # scene in reconstruct.reconstruct_cubes(gather=True), not per batch -- no device re-upload of host results
        return hp.infer_batch_host(pairs, xyz, resol, w, D, want_fused=True, ray_pool=True)

    def timed(fn, steps, warmup, profile=False):
        for _ in range(warmup):
            fn(); flush.fill_(1)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        _lib.lib.sn_launch_count_reset()
        if profile:
            _lib.lib.sn_profile_enable(1)
            torch.cuda.cudart().cudaProfilerStart()      # ncu --profile-from-start off captures the timed region only
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        t0 = time.time()
        for a, b in evs:
            a.record(); fn(); b.record()
            flush.fill_(1)                                                   # L2 flush between timed steps, outside the events
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t1 = time.time()
        if profile:
            torch.cuda.cudart().cudaProfilerStop()
        ms = sum(a.elapsed_time(b) for a, b in evs)
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        launches = int(_lib.lib.sn_launch_count())
        return float(t.item()) / steps, launches // steps, (t0, t1)

    sampler = ClockSampler(local) if rank == 0 else None
    # `value` comes from a plain pass; a second pass of the same K steps carries the per-launch CUDA events of the conv units (36 event
    # records per step inside the library) for the roofline and the per-unit table
    ms_dev, launches, win = timed(step_device, args.steps, args.warmup)
    ms_prof, _, _ = timed(step_device, args.steps, 1, profile=True)
    n_units = len(weights.UNITS)
    ms_u = (C.c_double * n_units)(); cnt_u = (C.c_int64 * n_units)()
    _lib.check(_lib.lib.sn_profile_collect(ms_u, cnt_u, n_units))
    _lib.lib.sn_profile_enable(0)
    clocks = sampler.window(*win) if sampler else None
    ms_e2e, _, _ = timed(step_host, args.steps, max(args.warmup, 1))
    h2d_e2e, d2h_e2e = hp.h2d_bytes, hp.d2h_bytes
    # the same loop body including colour fusion + dense2sparse (main_reconstruct.py:150-162): only kept voxels cross PCIe
    Dc = {32: 26, 64: 52}.get(D, D)                                          # params.py:107 __cube_Dcenter
    def step_sparse():
        return hp.infer_batch_sparse(pairs, xyz, resol, w, D, Dc)
    ms_sparse, _, _ = timed(step_sparse, max(2, args.steps // 2), 1)
    d2h_sparse = hp.d2h_bytes
    if sampler:
        sampler.stop()

    # BASELINE configs[3] (512 cubes x 8 view pairs over 8 GPUs): one shard = 64 cubes x 8 pairs per rank, timed next to the c3 line on
    # the multi-GPU runs (and on request with --c4-shard); same step, same gather
    c4 = None
    if (world >= 8 or args.c4_shard) and args.workload == "c3":
        w4 = WORKLOADS["c4"]
        p4, x4, r4, ww4 = make_inputs(w4, rank)
        dd = [torch.from_numpy(a).to(dev) for a in (p4, x4, r4, ww4)]
        g4p = torch.empty((world * w4["cubes"], D, D, D), dtype=torch.float16, device=dev) if world > 1 else None
        g4v = torch.empty((world * w4["cubes"], D, D, D), dtype=torch.uint8, device=dev) if world > 1 else None
        rp_ws = torch.empty(int(_lib.lib.sn_raypool_workspace_bytes(w4["cubes"], w4["n_vp"], D)), dtype=torch.uint8, device=dev)
        ms4, l4, _ = timed(lambda: device_step(dd[0], dd[1], dd[2], dd[3], g4p, g4v), max(2, args.steps // 4), 1)
        c4 = {"workload": w4["name"], "cubes_per_gpu": w4["cubes"], "view_pairs": w4["n_vp"], "ms_per_step": ms4, "comm_nranks": world,
              "value": world * w4["cubes"] * V / (ms4 * 1e-3), "unit": "voxels/s",
              "pair_voxels_per_s": world * w4["cubes"] * w4["n_vp"] * V / (ms4 * 1e-3), "gpu_launches": l4}
    if rank == 0:
        tc_peak, hbm_peak, peak_src = peaks()
        macs = unit_macs_per_voxel()
        pair_vox_step = B * n_vp * V
        conv_ms = sum(ms_u[i] for i in range(n_units))
        conv_launches = sum(cnt_u[i] for i in range(n_units))
        conv_flop_step = 2.0 * sum(macs) * pair_vox_step
        conv_tflops = conv_flop_step * args.steps / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
        # dominant kernel = the merge_conv2 launch of conv_tc_kernel (3x3x3, 100->100, + fused merge_conv3 + sigmoid epilogue)
        names = [u[0] for u in weights.UNITS]
        dom = names.index("merge_conv2")
        dom_flop_launch = 2.0 * (macs[dom] + macs[names.index("merge_conv3")]) * pair_vox_step * args.steps / max(cnt_u[dom], 1)
        dom_ms_launch = ms_u[dom] / max(cnt_u[dom], 1)
        achieved = dom_flop_launch / (dom_ms_launch * 1e-3) / 1e12 if dom_ms_launch > 0 else 0.0
        traffic = None
        summ = os.path.join(REPO, "profiles", "ncu_summary.json")
        if os.path.exists(summ):
            try:
                traffic = json.load(open(summ)).get(mode, {}).get("merge_conv2", {}).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        per_unit = {weights.UNITS[i][0]: {"ms_per_step": ms_u[i] / args.steps, "tflops": (2.0 * macs[i] * pair_vox_step * args.steps / (ms_u[i] * 1e-3) / 1e12) if ms_u[i] > 0 else 0.0}
                    for i in range(n_units) if cnt_u[i]}
        fused_vox = world * B * V
        line = {
            "metric": "fused surface-probability voxels/sec (CVC + SurfaceNet fwd + fusion + rayPooling)",
            "value": fused_vox / (ms_dev * 1e-3), "unit": "voxels/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {"fp32": "f32", "exact": "f16x2-split operands, f32 accumulate", "tc_exact": "f16x2-split operands, f32 accumulate",
                      "fast": "f16", "tc_fast": "f16"}[mode],
            "data": "synthetic",
            "config": {"workload": wl["name"], "mode": mode, "cubes_per_gpu": B, "view_pairs": n_vp, "cube_D": D,
                       "l2": "256 MiB flush write between timed steps; per-step activations exceed L2",
                       "parallelism": "cube-sharded x%d, one all-gather each of the float16 prediction (overlapped with ray pooling) and the votes per step" % world},
            "pair_voxels_per_s": world * pair_vox_step / (ms_dev * 1e-3),
            "path_tflops": FLOP_PER_PAIR_VOXEL * world * pair_vox_step / (ms_dev * 1e-3) / 1e12,
            "e2e": {"value": fused_vox / (ms_e2e * 1e-3), "unit": "voxels/s", "h2d_bytes_per_step": h2d_e2e, "d2h_bytes_per_step": d2h_e2e,
                    "ms_per_step": ms_e2e},
            "e2e_sparse": {"value": fused_vox / (ms_sparse * 1e-3), "unit": "voxels/s", "ms_per_step": ms_sparse, "d2h_bytes_per_step": d2h_sparse,
                           "what": "numpy in -> per-cube sparse lists out (adds colour fusion + centre-crop/threshold compaction on the GPU; no all-gather)"},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {"bound": "tensor", "kernel": "conv_wg_kernel<1,112,FINAL>, merge_conv2 launch (3x3x3 100->100 as w-axis Winograd F(2,3) + fused merge_conv3/sigmoid), %d launches/step" % (cnt_u[dom] // args.steps),
                         "achieved": achieved, "peak": tc_peak, "unit": "TFLOP/s", "frac": achieved / tc_peak, "traffic": traffic,
                         "peak_source": peak_src, "flop_per_launch": dom_flop_launch, "ms_per_launch": dom_ms_launch,
                         "kernel_share_of_step": ms_u[dom] / args.steps / ms_prof, "ms_per_step_with_events": ms_prof,
                         "all_conv_units": {"launches_per_step": conv_launches // args.steps, "tflops": conv_tflops, "frac": conv_tflops / tc_peak,
                                            "share_of_step": conv_ms / args.steps / ms_prof},
                         # merge_conv2, exact + Winograd F(2,3) along w: algorithmic 27 x 100 x 100 MAC per voxel; executed per 256-voxel tile 4 frequencies
                         # x 60 (channel block, tap) stages x (208 + 112) columns x 128 rows x 16 K = 614,400 MAC per voxel -> 0.4395 of the fp16 pipe
                         "exact_mode_ceiling_frac": 270000.0 / 614400.0 if mode in ("exact", "tc_exact") else None,
                         "per_unit": per_unit},
        }
        if c4 is not None:
            line["c4_shard"] = c4
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            cpu_reference_step(dict(wl, D=16), params, cams, imgs, 1, threads)        # warm torch's thread pool
            dt, ref = cpu_reference_step(wl, params, cams, imgs, 1, threads, return_outputs=True)
            line["cpu_baseline"] = {"value": V / dt, "unit": "voxels/s", "cores": threads, "kind": "port",
                                    "sample": "1 cube x %d view-pairs of %d^3 (%.1f s): numpy CVC + torch-CPU fp32 net + numpy ray pooling" % (n_vp, D, dt)}
            # the CPU leg doubles as the checker: the same cube through the timed GPU entry, compared with what the port computed
            got = hp.infer_batch_host(ref["pairs"], ref["xyz"], ref["resol"], ref["w"], D)
            line["parity_check"] = {"case": "the cpu_baseline cube (1 cube x %d view-pairs of %d^3), mode %s" % (n_vp, D, mode),
                                    "max_abs_prob_vs_port": float(np.abs(got["fused"] - ref["fused"]).max()), "tolerance": 1e-4,
                                    "votes_differing_voxels": int((got["votes"] != ref["votes"]).sum()), "voxels": int(V),
                                    "votes_note": "votes are counted on the float16 cast of each side's OWN probabilities: a difference <= tolerance can flip a "
                                                  "float16 rounding next to min_prob or a ray's arg-max; the vote logic itself is bit-exact on equal inputs "
                                                  "(tests/test_gpu_parity.py::test_infer_batch_host_matches_oracle_pipeline, test_raypool_*)"}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------------
def run_c5(args, wl):
    """BASELINE configs[4]: Middlebury dinoSparseRing (real images + calibration, tests/golden/real), views 7-12, s=64 cubes, N_vp=3,
    params.py:174-182.  One step = the whole reconstruct.reconstruction call: cube grid (2,016 cubes) -> early rejection (similarityNet)
    -> view-pair selection -> SurfaceNet inference on this rank's share of the cube batches -> gather of the sparse lists -> fixed
    threshold + cross-cube denoising.  value = fused voxels of the cubes that survive early rejection / s (whole job)."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from surfacenet_b200 import _lib, camera, image, reconstruct, similarityNet, weights
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    real = os.path.join(REPO, "tests", "golden", "real")
    views = list(range(7, 13))
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        imgs = image.readImages(real, "dinoSparseRing/dinoSR0#.png", views)
    P = camera.readCameraPOs_as_np(real, "Middlebury", "dinoSparseRing/dinoSR_par.txt", "dinoSparseRing", views)
    BB = np.array([(-0.061897, 0.010897), (-0.018874, 0.068227), (-0.057845, 0.015495)], dtype=np.float32)
    params = weights.synthetic_params(1)
    sp = similarityNet.synthetic_params(0)
    sp[28] = np.array([[-0.02]], np.float32); sp[29] = np.array([-0.4], np.float32)
    D = wl["D"]

    def step():
        return reconstruct.reconstruction(imgs, P, BB, np.float32(0.00025), wl["n_vp"], params, sp, outputFolder=None, cube_D=D, batch_size=16,
                                          tau=0.7, gamma=0.8, model="dinoSparseRing", rank=rank, world_size=world)
    sampler = ClockSampler(local) if rank == 0 else None
    for _ in range(max(1, min(args.warmup, 1))):
        out = step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    _lib.lib.sn_launch_count_reset()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = max(1, min(args.steps, 3))
    t0 = time.time()
    e0.record()
    for _ in range(steps):
        out = step()
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t1 = time.time()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / steps
    if rank == 0:
        n_valid = int(out["validCubes"].sum())
        vox = n_valid * D ** 3
        line = {"metric": "fused surface-probability voxels/sec, whole-scene reconstruction (early rejection + selection + CVC + SurfaceNet + rayPooling + sparsify + denoise)",
                "value": vox / (ms * 1e-3), "unit": "voxels/s", "n_gpus": world, "steps": steps, "warmup": 1, "ms_per_step": ms, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f16x2-split operands, f32 accumulate", "data": "real images (reference fixtures), synthetic weights",
                "config": {"workload": wl["name"], "cubes_in_grid": int(out["validCubes"].size), "cubes_after_early_rejection": n_valid,
                           "sparse_voxels": int(sum(len(x) for x in out["result"][0])), "view_pairs": wl["n_vp"], "cube_D": D,
                           "parallelism": "cube batches dealt round-robin to %d ranks, all_gather_object of the sparse lists" % world},
                "gpu_launches": int(_lib.lib.sn_launch_count()) // steps, "clocks": sampler.window(t0, t1) if sampler else None,
                "e2e": {"value": vox / (ms * 1e-3), "unit": "voxels/s", "h2d_bytes_per_step": int(sum(i.nbytes for i in imgs)), "d2h_bytes_per_step": int(sum(len(x) for x in out["result"][0]) * 8),
                        "what": "the step IS the host-facing call: numpy images / cameras in, sparse numpy lists out"}}
        print(json.dumps(line))
    if sampler:
        sampler.stop()
    if world > 1:
        dist.barrier(); dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------------
# "next" row N4: the post-processing of the scene's sparse cubes (main_reconstruct.py:170-173 + utils/adapthresh.py:91-178)
POST_GRID, POST_D, POST_ITERS = (8, 8, 4), 52, 8          # 256 half-overlapping 52^3 centre cubes, params.py:107,113


def post_step_oracle(sc, iters):
    """The reference's CPU post-processing (py3 restatement): fixed-threshold filter + denoise, then `iters` refinement
    iterations each followed by a denoise."""
    from oracle import postprocess_oracle as post
    D = sc["D"]
    m = post.filter_voxels([], sc["pred_list"], 0.7, sc["votes_list"], 8)
    post.denoise_crossCubes(sc["cube_ijk"], sc["ijk_list"], m, D)
    post.adapthresh_core(sc["pred_list"], sc["ijk_list"], sc["votes_list"], sc["cube_ijk"], iters, D, 0.5, 0.5, 0.9, 8, 6)


def post_traffic(dom):
    try:
        d = json.load(open(os.path.join(REPO, "profiles", "ncu_summary.json")))["post"]
        return d["denoise_call"]["dram_bytes_per_call"] if dom == "denoise" else None
    except Exception:
        return None


def run_post(args):
    import numpy as np
    rank = int(os.environ.get("RANK", "0"))
    from tests import util
    if args.impl == "reference":
        if rank != 0:
            return
        sc = util.sparse_scene((2, 2, 2), POST_D, seed=50, floaters=40, thick=0.05)
        n = sum(x.shape[0] for x in sc["ijk_list"])
        for _ in range(min(args.warmup, 1)):
            post_step_oracle(sc, 1)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            post_step_oracle(sc, POST_ITERS)
        dt = (time.perf_counter() - t0) / args.steps
        desc = "8 cubes of %d^3 (%d voxels), %d refinement iterations per step: numpy/scipy restatement, 1 thread" % (POST_D, n, POST_ITERS)
        print(json.dumps({"impl": "reference", "metric": "sparse voxels/sec through post-processing (filter + denoise + adapthresh)",
                          "value": n / dt, "unit": "voxels/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                          "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/f16",
                          "data": "synthetic", "config": {"workload": "post", "sample": desc},
                          "cpu_baseline": {"value": n / dt, "unit": "voxels/s", "cores": 1, "kind": "port", "sample": desc},
                          "e2e": {"value": n / dt, "unit": "voxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}))
        return
    import torch
    import torch.distributed as dist
    from surfacenet_b200 import _lib, adapthresh
    from surfacenet_b200.sparse_device import DeviceSparseCubes
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    # every rank post-processes its own slab of the scene (replicas of the same synthetic generator with another seed)
    sc = util.sparse_scene(POST_GRID, POST_D, seed=60 + rank, floaters=40, thick=0.05)
    dsc = DeviceSparseCubes(sc["cube_ijk"], sc["ijk_list"], sc["pred_list"], sc["votes_list"])
    N, C = dsc.N, dsc.C
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ev = lambda: torch.cuda.Event(enable_timing=True)
    parts = {"filter": 0.0, "denoise": 0.0, "adapthresh_iter": 0.0}
    n_calls = {"filter": 0, "denoise": 0, "adapthresh_iter": 0}

    def timed_call(name, fn, record):
        if not record:
            return fn()
        a, b = ev(), ev()
        a.record(); r = fn(); b.record()
        pending.append((name, a, b))
        return r

    pending = []

    def step_device(record=False):
        m = timed_call("filter", lambda: dsc.filter_voxels(None, prob_thresh=0.7, rayPool_thresh=8), record)
        timed_call("denoise", lambda: dsc.denoise(m, POST_D), record)                       # main_reconstruct.py:170-173
        init = dsc.filter_voxels(None, prob_thresh=0.5, rayPool_thresh=8)                     # adapthresh.py:103
        timed_call("denoise", lambda: dsc.denoise(init, POST_D), record)
        mask = init.clone()
        thresh = torch.full((C,), 0.5, dtype=torch.float64, device=dev)
        for _ in range(POST_ITERS):
            timed_call("adapthresh_iter", lambda: dsc.adapthresh(init, mask, thresh, POST_D, 0.9, 6, n_iter=1), record)
            timed_call("denoise", lambda: dsc.denoise(mask, POST_D), record)
        return thresh

    def step_host():
        return adapthresh.adapthresh_lists(sc["pred_list"], sc["ijk_list"], sc["votes_list"], sc["cube_ijk"], POST_ITERS, POST_D, 0.5, 0.9, 8, 6)

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn(); flush.fill_(1)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        _lib.lib.sn_launch_count_reset()
        evs = [(ev(), ev()) for _ in range(steps)]
        t0 = time.time()
        for a, b in evs:
            a.record(); fn(); b.record(); flush.fill_(1)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t1 = time.time()
        t = torch.tensor([sum(a.elapsed_time(b) for a, b in evs)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) / steps, int(_lib.lib.sn_launch_count()) // steps, (t0, t1)

    sampler = ClockSampler(local) if rank == 0 else None
    ms_dev, launches, win = timed(lambda: step_device(False), args.steps, args.warmup)
    clocks = sampler.window(*win) if sampler else None
    step_device(True)                                         # per-call split, separate pass (events around each C-ABI call)
    torch.cuda.synchronize()
    for name, a, b in pending:
        parts[name] += a.elapsed_time(b); n_calls[name] += 1
    t0 = time.perf_counter()
    e2e_steps = max(1, args.steps // 2)
    for _ in range(e2e_steps):
        step_host()
    torch.cuda.synchronize()
    ms_e2e = (time.perf_counter() - t0) * 1e3 / e2e_steps
    if sampler:
        sampler.stop()
    if rank == 0:
        _, hbm_peak, peak_src = peaks()
        # algorithmic bytes per voxel and call: every kernel streams each per-voxel array it needs once
        #   denoise (11 launches): vox_cube 4 w | bitmap: ijk 3 + mask 1 + vox_cube 4 | link: 8 + parent 4 w | flatten: vox_cube 4 +
        #     parent 4 r + 4 w | reduce: 8 | root: vox_cube 4 + parent 4 + root 4 w | overlap: ijk 3 + mask 1 + vox_cube 4 + root 4 |
        #     keep: ijk 3 + mask 1 + vox_cube 4 + root 4 + ovl 1 + keep 1 w | ovl memset 1                = 83 B per voxel
        #   adapthresh iteration (9 launches): vox_cube 4 w + bitmap 8 | occ0: ijk 3 + pred 2 + mask 1 | count: same 6 |
        #     filter: pred 2 + mask 1 + vox_cube 4 + mask 1 w                                             = 32 B per voxel
        #   (the per-cube bitmaps / prefixes, 2 x 17.6 KB per cube, stay in L2 and are not counted)
        bytes_call = {"denoise": 83.0 * N, "adapthresh_iter": 32.0 * N, "filter": 12.0 * N}
        dom = max(("denoise", "adapthresh_iter"), key=lambda k: parts[k])
        dom_ms = parts[dom] / max(n_calls[dom], 1)
        achieved = bytes_call[dom] / (dom_ms * 1e-3) / 1e9
        line = {"metric": "sparse voxels/sec through post-processing (filter + denoise + adapthresh)", "value": world * N / (ms_dev * 1e-3),
                "unit": "voxels/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/f16", "data": "synthetic",
                "config": {"workload": "post: %dx%dx%d half-overlapping %d^3 centre cubes (%d cubes, %d sparse voxels) per GPU; step = filter(tau,gamma) + denoise + "
                                       "adapthresh init + %d x (refinement iteration + denoise)" % (POST_GRID + (POST_D, C, N, POST_ITERS)),
                           "l2": "256 MiB flush write between timed steps", "parallelism": "independent slabs x%d, no collective" % world},
                "e2e": {"value": world * N / (ms_e2e * 1e-3), "unit": "voxels/s", "ms_per_step": ms_e2e,
                        "h2d_bytes_per_step": int(N * (3 + 2 + 1) + C * 20), "d2h_bytes_per_step": int(N * (POST_ITERS * 2 + 2) + C * 8),
                        "what": "adapthresh_lists: per-cube numpy lists in, per-iteration masks + denoised masks out (wall clock)"},
                "gpu_launches": launches, "clocks": clocks,
                "roofline": {"bound": "hbm", "kernel": "the %s C-ABI call (%d launches)" % (dom, {"denoise": 11, "adapthresh_iter": 9}[dom]),
                             "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": post_traffic(dom),
                             "peak_source": peak_src, "bytes_per_call": bytes_call[dom], "ms_per_call": dom_ms,
                             "calls_ms_per_step": {k: parts[k] for k in parts}, "calls_per_step": n_calls}}
        if c4 is not None:
            line["c4_shard"] = c4
        if world == 1 and not args.no_cpu_baseline:
            small = util.sparse_scene((2, 2, 2), POST_D, seed=50, floaters=40, thick=0.05)
            n_small = sum(x.shape[0] for x in small["ijk_list"])
            t0 = time.perf_counter()
            post_step_oracle(small, 2)
            dt = (time.perf_counter() - t0) * (1 + 2 + 2 * POST_ITERS) / (1 + 2 + 2 * 2)          # scale the iteration count up to POST_ITERS
            line["cpu_baseline"] = {"value": n_small / dt, "unit": "voxels/s", "cores": 1, "kind": "port",
                                    "sample": "8 cubes of %d^3 (%d voxels), 2 of %d refinement iterations timed and scaled: numpy/scipy restatement of "
                                              "denoising.py + adapthresh.py" % (POST_D, n_small, POST_ITERS)}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------------
# "next" row N3: early rejection's patch embedding (utils/earlyRejection.py:6-55 on nets/similarityNet.py:23-56)
SIMNET_FLOP_PER_PATCH = 2.0 * (7.08e6 + 151.0e6 + 75.5e6 + 151.0e6 + 75.5e6 + 2 * 151.0e6 + 75.5e6 + 2 * 151.0e6 + 3 * 37.75e6 + 0.754e6)


def run_simnet(args):
    import numpy as np
    rank = int(os.environ.get("RANK", "0"))
    from tests import util
    n_patches = 2048
    if args.impl == "reference":
        if rank != 0:
            return
        for k in ("OMP_NUM_THREADS", "MKL_NUM_THREADS"):
            os.environ[k] = str(os.cpu_count() or 1)
        import torch
        from oracle import selection_oracle as so
        from surfacenet_b200 import similarityNet
        torch.set_num_threads(os.cpu_count() or 1)
        params = similarityNet.synthetic_params(0)
        patches = np.random.RandomState(0).randn(64, 3, 64, 64).astype(np.float32) * 50
        for _ in range(args.warmup):
            so.patch2embedding_fn(patches, params)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            so.patch2embedding_fn(patches, params)
        dt = (time.perf_counter() - t0) / args.steps
        desc = "64 patches of 3x64x64 per step, torch-CPU fp32 VGG-16 embedding"
        print(json.dumps({"impl": "reference", "metric": "patch embeddings/sec (similarityNet patch2embedding)", "value": 64 / dt, "unit": "patches/s",
                          "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
                          "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": {"workload": "simnet", "sample": desc},
                          "cpu_baseline": {"value": 64 / dt, "unit": "patches/s", "cores": os.cpu_count() or 1, "kind": "port", "sample": desc},
                          "e2e": {"value": 64 / dt, "unit": "patches/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}))
        return
    import torch
    from surfacenet_b200 import _lib, image, similarityNet
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    net = similarityNet.SimNet(similarityNet.synthetic_params(0))
    img = torch.from_numpy(util.synth_image(5, 1200, 1600)).cuda()
    rs = np.random.RandomState(3)
    ch, cw = rs.rand(n_patches) * 1200, rs.rand(n_patches) * 1600
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def step():
        return net.patch2embedding(image.crop_preprocessed_patches_device(img, ch, cw, 64, util.MEAN_BGR))

    for _ in range(args.warmup):
        step(); flush.fill_(1)
    torch.cuda.synchronize()
    _lib.lib.sn_launch_count_reset()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    sampler = ClockSampler(local)
    t0 = time.time()
    for a, b in evs:
        a.record(); out = step(); b.record(); flush.fill_(1)
    torch.cuda.synchronize()
    t1 = time.time()
    clocks = sampler.window(t0, t1); sampler.stop()
    ms = sum(a.elapsed_time(b) for a, b in evs) / args.steps
    launches = int(_lib.lib.sn_launch_count()) // args.steps
    t0 = time.perf_counter()
    for _ in range(args.steps):
        emb = step().cpu().numpy()                                  # e2e: centres up, embeddings back
    ms_e2e = (time.perf_counter() - t0) * 1e3 / args.steps
    tf = SIMNET_FLOP_PER_PATCH * n_patches / (ms * 1e-3) / 1e12
    fma_peak = 148 * 128 * 2 * (clocks["sm_mhz"] or 1965.0) * 1e6 / 1e12
    line = {"metric": "patch embeddings/sec (similarityNet patch2embedding)", "value": n_patches / (ms * 1e-3), "unit": "patches/s", "n_gpus": 1,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "simnet: %d patches of 3x64x64 cropped from a 1200x1600 image per step (crop + VGG-16 embedding)" % n_patches,
                       "l2": "256 MiB flush write between timed steps"},
            "e2e": {"value": n_patches / (ms_e2e * 1e-3), "unit": "patches/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": n_patches * 16,
                    "d2h_bytes_per_step": n_patches * 128 * 4},
            "gpu_launches": launches, "clocks": clocks,
            "roofline": {"bound": "cuda-core fp32 FMA (not a tensor-core kernel)", "kernel": "conv2d3x3_kernel x13 (whole embedding)", "achieved": tf,
                         "peak": fma_peak, "unit": "TFLOP/s", "frac": tf / fma_peak, "traffic": None,
                         "peak_source": "148 SMs x 128 FMA lanes x 2 x measured SM clock", "flop_per_patch": SIMNET_FLOP_PER_PATCH}}
    if not args.no_cpu_baseline:
        import torch as _t
        from oracle import selection_oracle as so
        _t.set_num_threads(os.cpu_count() or 1)
        params = similarityNet.synthetic_params(0)
        pp = np.random.RandomState(0).randn(64, 3, 64, 64).astype(np.float32) * 50
        so.patch2embedding_fn(pp[:8], params)
        t0 = time.perf_counter(); so.patch2embedding_fn(pp, params); dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": 64 / dt, "unit": "patches/s", "cores": os.cpu_count() or 1, "kind": "port",
                                "sample": "64 patches, torch-CPU fp32 VGG-16 embedding (%.2f s)" % dt}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS) + ["post", "simnet"])
    ap.add_argument("--mode", default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--c4-shard", action="store_true", help="also time one BASELINE configs[3] shard (64 cubes x 8 view pairs per GPU); always on at 8 GPUs")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.workload == "post":
        return run_post(args)
    if args.workload == "simnet":
        return run_simnet(args)
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl)
    elif args.workload == "c5":
        run_c5(args, wl)
    else:
        run_gpu(args, wl)


if __name__ == "__main__":
    main()
