"""CPU oracle for the MR-Flow plane+parallax hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``mrflow_b200/`` may import this package; the only
permitted users are ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` (as the checker / the timed CPU arm, never as the
product).

What it is: a numpy + OpenCV restatement of the reference functions on the path named by
SURVEY.md section 8a, each function citing the reference file:line it follows.  The reference is
Python, so the restatement calls the very same third-party primitives the reference calls
(``cv2.remap``, ``cv2.perspectiveTransform``, ``cv2.cvtColor``, ``cv2.GaussianBlur``,
``cv2.erode``, ``scipy.special.ive``; OpenCV 4.13.0 / scipy 1.18.1 in this image -- the reference
pins neither, README.md:19 only says "OpenCV >= 3.0").  ``oracle.cvemul`` additionally restates
those OpenCV primitives in plain numpy, bit-exactly, so the CUDA kernels have an arithmetic
specification that does not depend on OpenCV's binary.

``oracle/build_ref.py`` additionally compiles the reference's own C++ MRF solver (extern/eff_trws) from
the sources under /root/reference into ``oracle/_ref/`` (git-ignored); ``oracle/trws.py`` wraps it as the
reference's ``infer_mrf`` so that ``compute_structure`` can be compared end to end.

Parity status
-------------
* a1-a8, a10 (flow<->structure maps, sampler, robust functions, rigidity unaries): PINNED.
  ``tests/golden/make_golden.py`` imports the real reference from ``/root/reference`` (with the
  import shims of SURVEY.md section 8c, none of which alters arithmetic) and stores its outputs on
  seeded inputs in ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks this oracle
  against them.
* a11 (``computeDataTerm``): PINNED the same way -- the energy, and the complete linear system
  (diag(A_data), b_data) of the reference function with its auto-sigma Lorentzian; a9
  (``compute_derivs_through_H``), ``computeOcclusionsFromConsistency`` and ``get_image_weights`` likewise.
* a5 (``correct_epipole``, pipeline/compute_structure.py:453-511): PINNED.  The generator runs the reference's
  own function -- objective, BFGS call, acceptance rule untouched -- with a small ``chumpy.ch`` stand-in whose
  derivative comes from torch autograd (tests/golden/make_golden.py::_Ch); the restatement here with the
  analytic gradient lands within 1.6e-9 px of it (bar 1e-6; one epipole of the 'inside' fixture moves 1.7 px,
  the other is kept).
* a4 end to end: the reference's ``compute_structure`` run with its own solver (oracle/_ref) is stored
  for all four fixtures (one with moving objects, 14 % non-rigid after the MRF; one with the epipole inside the
  frame, where the reference's own q_new is fed to both sides so that everything downstream is compared bit for
  bit); the oracle reproduces all seven outputs bit for bit.
* configs[0] at FULL size: example/frame1-3.png (1024x436), synthetic flows, uniform rigidity,
  override_optimization=1 through the reference's chain of mrflow.py:98-140; the oracle reproduces the sha256
  of every output (tests/golden/ref_example_full.npz, test_config1_full_size_chain_pinned).
* stage 5 (``oracle/variational.py``): the variational refinement ``optimizeStructureSingleScale``
  (optimize_structure_single_scale.py:310-652) and the reference's whole default stage 5
  (``optimize`` -> ``optimizeStructureMultiScale`` -> ...): PINNED -- bit-identical structure and energies
  on all four fixtures, plus a case with a carved non-rigid block (same scipy / OpenCV calls; the explicit stencil matrices are asserted
  equal to construct_matrices.py's when the goldens are generated).
* a12 (the per-label cost volume): the reference's arithmetic for this step lived in a compiled
  extension that is NOT shipped (``pipeline/build_cost_volume.py:26-27``).  The oracle composes
  it from shipped reference functions exactly as SURVEY.md section 8c defines; each label plane is
  pinned against the reference's own ``computeDataTerm`` geometry + ``interpolate_through_H`` +
  ``robust_functions`` run at that label (golden ``costvol_*.npz``), but the *choice* of
  composition (rho, sigma0, channels) is ours: for a12 as a whole, "parity unpinned" by the
  reference's own tests -- it has none.
"""
