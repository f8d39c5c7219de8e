#!/usr/bin/env python
"""Golden fixture for the cube-grid builders of utils/scene.py (initializeCubes 7-61, quantizePts2Cubes 63-107), produced by executing
the reference functions.  Writes tests/golden/scene_golden.npz.  Shims, in memory only: the plyfile / mesh_util imports (unused by the
two functions) are dropped, `cubes_ijk.size / 3` is the python-2 integer division, the module-level doctest.testmod() is removed.
`resol` is passed as float(np.float32(0.4)): the reference passes np.float32(0.4) (params.py) under numpy 1.x, whose scalar promotion
makes `resol * cube_D` (float32 scalar * python int) a FLOAT64 product; numpy 2 (NEP 50) would keep float32 and move 432 of 486 cube
origins by up to 1.5e-5 mm.  A python float reproduces the numpy-1.x arithmetic under numpy 2."""
import os, sys, types
import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    src = open(os.path.join(REF, "utils", "scene.py")).read().replace("import doctest\ndoctest.testmod()", "")
    src = src.replace("from plyfile import PlyData, PlyElement\n", "").replace("import mesh_util\n", "")
    src = src.replace("N_cubes = cubes_ijk.size / 3", "N_cubes = cubes_ijk.size // 3")
    mod = types.ModuleType("ref_scene")
    exec(compile(src, os.path.join(REF, "utils", "scene.py"), "exec"), mod.__dict__)
    rs = np.random.RandomState(0)
    pts = rs.rand(200, 3) * np.array([30.0, 20.0, 10.0]) + np.array([5.0, -3.0, 600.0])
    cubes, D_mm = mod.quantizePts2Cubes(pts, resol=float(np.float32(0.4)), cube_D=32, cube_Dcenter=26, cube_overlapping_ratio=0.5,
                                        BB=np.array([[0., 40.], [-5., 20.], [598., 612.]]))
    c2, _ = mod.initializeCubes(resol=float(np.float32(0.4)), cube_D=64, cube_Dcenter=52, cube_overlapping_ratio=0.5,
                                BB=np.array([[-73., 129.], [-197., 183.], [472., 810.]]))            # DTU scan9: 24,420 cubes (q.log)
    np.savez_compressed(os.path.join(HERE, "scene_golden.npz"), q_pts=pts, q_xyz=cubes["xyz"], q_ijk=cubes["ijk"], q_resol=cubes["resol"],
                        q_D_mm=np.array([D_mm]), init_n=np.array([len(c2)]), init_xyz_sample=c2["xyz"][::997], init_ijk_sample=c2["ijk"][::997])
    print("wrote scene_golden.npz:", len(cubes), "quantised cubes,", len(c2), "grid cubes")


if __name__ == "__main__":
    main()
