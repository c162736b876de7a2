"""CPU oracle for the Thermal3D-Vision per-pixel hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and there only as the checker (or as
the thing timed on the host cores), never as the implementation that ships.
The product path (``thermal3d_vision_b200``) fails loudly when the CUDA
library is missing and never routes through this package.

What is here
------------
* ``ref_loss``        torch-CPU / numpy-fp64 restatement of ``utils/loss.py``
* ``ref_preprocess``  numpy restatement of ``utils/preprocessing.py`` + the
                      ``cv2.resize`` recipes used around it
* ``ref_metrics``     numpy restatement of ``utils/metrics.py`` and
                      ``utils/evaluate_depth_metrics.py:20-80``
* ``ref_depth``       pointmap->depth, intrinsics (``scripts/pseudo_gt.py``)
* ``ref_sobel``       torch restatement of ``ThermalDUSt3R.preprocess_thermal``
* ``reference_bridge`` imports the *unmodified* reference from ``/root/reference``
                      (only available in the build container, never on the GPU
                      box) to pin the restatements and to generate
                      ``tests/golden/*`` via ``gen_golden.py``.

Parity status: PINNED.  The reference has no tests or golden vectors of its
own, so the restatements are pinned against outputs of
the reference itself run in the build container (``gen_golden.py`` ->
``tests/golden/``) and, when ``/root/reference`` is mounted, against the live
reference functions on fresh random inputs (``tests/test_oracle_pin.py``).
"""
