"""surfacenet_b200 -- B200-native implementation of the SurfaceNet per-cube inference hot path
(CVC construction -> 3D SurfaceNet forward + view-pair fusion -> ray-pool votes) behind the
reference's Python call surface (mjiUST/SurfaceNet: utils/CVC.py, nets/SurfaceNet.py,
utils/rayPooling.py, utils/camera.py:perspectiveProj, main_reconstruct.py:132-162).

Sub-modules mirror the reference's module names:
    surfacenet_b200.CVC          gen_coloredCubes, preprocess_augmentation
    surfacenet_b200.SurfaceNet   SurfaceNet_inference -> (viewPair_relativeImpt_fn, nViewPair_SurfaceNet_fn)
    surfacenet_b200.rayPooling   rayPooling_1cube_numpy
    surfacenet_b200.camera       perspectiveProj
    surfacenet_b200.pipeline     HotPath.infer_batch  (the fused per-batch loop body)
    surfacenet_b200.weights      parameter-list layout, loader, synthetic generator (pure host logic)
All compute runs in libsurfacenet_b200.so (hand-written sm_100a CUDA); importing any compute
sub-module without the built library raises ImportError, running without a GPU raises RuntimeError.
"""
__version__ = "0.1.0"
