"""Drop-in for main_reconstruct.reconstruction (main_reconstruct.py:28-183) with the reference's argument list: reads the images and
cameras (utils/image.py:50-89, utils/camera.py:55-82), then runs surfacenet_b200.reconstruct.reconstruction (cube grid -> early
rejection -> view-pair selection -> SurfaceNet inference -> thresholding + denoising -> PLY + NPZ) on the GPU.

The constants the reference takes from params.py are keyword arguments with the values of params.py:65-119 as defaults; the two
model files default to the reference's paths below `input_data_rootFld` (they are not distributed with the reference, so callers
normally pass parameter lists or their own files)."""
import os
import numpy as np
from . import camera, image, reconstruct, similarityNet, sparseCubes, weights

PRETRAINED_SURFACENET = 'SurfaceNet_models/2D_2_3D-19-0.918_0.951.model'          # params.py:106
PRETRAINED_SIMILNET = 'SurfaceNet_models/epoch33_acc_tr0.707_val0.791.model'      # params.py:92


def reconstruction(datasetFolder, model, imgNamePattern, poseNamePattern, initialPtsNamePattern, outputFolder, N_viewPairs4inference,
                   resol, BB, viewList, datasetName="DTU", cube_D=64, input_data_rootFld=None, surfacenet_model=None, similnet_model=None,
                   mode="exact", weighted_fusion=True, batch_size=16, min_prob=0.46, tau=0.7, gamma=0.8, cube_overlapping_ratio=0.5,
                   rank=0, world_size=1):
    """-> path of the saved NPZ ('model{model}-{N_views}views.npz' in outputFolder, main_reconstruct.py:180-183) or "Empty!"."""
    initial_pts_xyz = None
    if initialPtsNamePattern is not None:                                                                                                # :57
        initial_pts_xyz = sparseCubes.readPointCloud_xyz(os.path.join(datasetFolder, initialPtsNamePattern))
    images_list = image.readImages(datasetFolder=datasetFolder, imgNamePattern=imgNamePattern, viewList=viewList, return_list=True)      # :48
    cameraPOs_np = camera.readCameraPOs_as_np(datasetFolder=datasetFolder, datasetName=datasetName, poseNamePattern=poseNamePattern,
                                              model=model, viewList=viewList)                                                         # :49
    root = input_data_rootFld if input_data_rootFld is not None else os.path.dirname(os.path.abspath(datasetFolder))
    sp = surfacenet_model if surfacenet_model is not None else weights.load_model_file(os.path.join(root, PRETRAINED_SURFACENET))
    if isinstance(sp, str):
        sp = weights.load_model_file(sp)
    mp = similnet_model if similnet_model is not None else os.path.join(root, PRETRAINED_SIMILNET)
    if isinstance(mp, str):
        mp = similarityNet.load_model_file(mp)
    out = reconstruct.reconstruction(images_list, cameraPOs_np, np.asarray(BB, dtype=np.float64), resol, N_viewPairs4inference, sp, mp,
                                     outputFolder=outputFolder, cube_D=cube_D, mode=mode, weighted_fusion=weighted_fusion,
                                     batch_size=batch_size, min_prob=min_prob, tau=tau, gamma=gamma,
                                     cube_overlapping_ratio=cube_overlapping_ratio, model=model, rank=rank, world_size=world_size,
                                     initial_pts_xyz=initial_pts_xyz)
    return out if out == "Empty!" else out["npz_path"]
