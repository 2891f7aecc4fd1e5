"""CPU oracle for the triplane volume-rendering hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in ``triplaneturbo_b200/`` may import this
package: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs use it, and only as the checker or
the timed CPU baseline.

What it is: a pure-PyTorch (CPU, fp32 or fp64) restatement of the reference's
algorithm for the path named by BASELINE.json, every function citing the
reference ``file:line`` it follows (paths relative to ``/root/reference``).

Parity pinning status
---------------------
* PINNED against the reference's own Python source executed in the build
  container: geometry (plane rotation, triplane sampling, decoder MLPs,
  shifted SDF, analytic normal), NeuS alpha, the volume renderer ``_forward``
  (image outputs and per-sample extras), ``ImportanceEstimator.sampling``
  control flow, ``NoMaterial``, ``scale_tensor``, ``chunk_batch``,
  ``PatchRenderer``.  ``tests/golden/make_golden.py`` imports those reference
  files (with the absent third-party packages stubbed) and commits the vectors
  in ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks the oracle
  against them.
* UNPINNED: the three nerfacc v0.5.2 entry points (``importance_sampling``,
  ``render_weight_from_alpha``/``render_transmittance_from_density``,
  ``accumulate_along_rays``).  nerfacc is an un-vendored dependency
  (``requirements.txt:5``, ``git+https://github.com/KAIR-BAIR/nerfacc.git@v0.5.2``)
  whose source is nowhere in this image and the reference has no tests or
  golden vectors at that boundary.  ``oracle/nerfacc_restated.py`` restates the
  documented semantics; the quantile placement of ``importance_sampling`` is
  the one unverifiable detail (see that file's header).
* The second derivative of bilinear sampling (reference:
  ``extern/grid_sample_gradfix/gridsample_cuda.cu:87-209``, CUDA only) is
  restated as a gather-based bilinear that autograd differentiates twice; it
  is pinned against ``F.grid_sample`` (value and first derivative) and against
  fp64 ``gradgradcheck``.  In addition ``oracle/build_ref.py`` compiles the
  reference's own extension (from the sources under ``/root/reference``, output
  only in ``oracle/_ref/``) and ``tests/test_sampler.py`` runs its ``grad2_2d``
  next to ``tt_sample_planes_bwdbwd`` on the GPU.
"""
