"""CPU oracle for the SurfaceNet per-cube inference hot path.

TEST INFRASTRUCTURE ONLY.  This package restates, for python3 / numpy 2 / torch-CPU, the
arithmetic of the reference (mjiUST/SurfaceNet) files named in SURVEY.md section 8(c).  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it, and only as the checker / the CPU baseline.  Nothing under
``surfacenet_b200/`` imports it; the product path fails loudly when the CUDA library is missing.

Pinning status (see DESIGN.md "Oracle"):
  * camera.perspectiveProj  - pinned by the reference doctest known answers (utils/camera.py:144-160)
    and by golden vectors produced by importing the reference module itself
    (tests/golden/make_golden.py).
  * CVC.__colorize_cube__ / gen_coloredCubes, rayPooling_1cube_numpy - the reference has no tests
    for them; pinned instead against OUTPUTS OF THE REFERENCE CODE ITSELF executed in the build
    container with mechanical py3/numpy-2 shims (tests/golden/make_golden.py, fixtures committed).
  * SurfaceNet network forward (Theano/Lasagne + cuDNN only; cannot run anywhere here) -
    PARITY UNPINNED: restated from nets/SurfaceNet.py and the published Lasagne@7992faa layer
    semantics; only the fixed up-sampling kernels (nets/layers.py:363-374 __W_5D__) are pinned by
    executing the reference function.
"""
