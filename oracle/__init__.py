"""CPU oracle for the cGP per-cluster GP hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``cgp_b200/`` imports this package.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and there only as the checker.

The reference (gptune/cGP) is pure Python and delegates the arithmetic of this path to
scikit-learn / SciPy / LAPACK, which are third-party packages *not* vendored under
``/root/reference`` (``requirements.txt:1-6`` lists them unpinned).  This package restates
the published algorithm of the versions pinned by this image (scikit-learn 1.9.0,
SciPy 1.18.1) in plain NumPy + LAPACK, citing the file:line each function follows.

Parity pin status: the reference ships no tests.  The oracle is pinned against
  (1) the reference's only result-bearing fixture ``cGP/BUKIN.pkl`` (two fitted GPRs and a
      fitted k-NN classifier; arrays extracted into ``tests/golden/bukin.npz``), and
  (2) outputs of the reference's own code path run in the build container
      (scikit-learn objects configured exactly as ``cGP/cGP_constrained.py:169,315-316`` and
      ``cGP/acquisition.py:4-32`` imported from ``/root/reference``), committed under
      ``tests/golden/`` together with the generating script ``oracle/make_golden.py``.
"""
from .gp_oracle import (  # noqa: F401
    LOG_BOUND_LO,
    LOG_BOUND_HI,
    kernel_matrix,
    kernel_cross,
    kernel_matrix_grad,
    normalize_y,
    lml_and_grad,
    restart_thetas,
    fit_gpr,
    predict_mean_std,
    knn_labels,
    ei_scores,
    select_next,
    merge_small_clusters,
    recode,
)
