"""CPU oracle for the RoDyGS dynamic-splatting hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``rodygs_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs do, and there only as the checker
or as the timed CPU baseline.

Parity status
-------------
* Rasterizer internals (preprocess / binning / blend): **parity unpinned**.
  The reference's rasterizer is the un-vendored submodule
  ``slothfulxtx/diff-gaussian-rasterization`` (branch ``pose``, no pinned SHA,
  ``/root/reference/.gitmodules:1-4``) and the reference ships no tests or
  golden vectors.  The restatement follows the published 3DGS algorithm
  (Kerbl et al. 2023) as written down in SURVEY.md Appendix A and is anchored
  on the reference's call sites (``src/trainer/renderer.py:50-101``).
* SH basis, L1 / SSIM / Pearson losses, time embedding + motion-basis MLP and
  the deformation rule: **pinned** against the reference's own Python
  (``src/utils/sh_utils.py``, ``src/utils/loss_utils.py``,
  ``src/model/rodygs_dynamic.py``) through the committed fixtures under
  ``tests/golden/`` (generator: ``tests/golden/make_golden.py``).
"""
