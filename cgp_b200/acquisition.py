"""Acquisition functions with the reference's plugin signature (REF/cGP/acquisition.py:4-32, selected by name
through ``-A``, REF/cGP/cGP_constrained.py:156-162) plus the batched form the GPU scan uses."""
from __future__ import annotations

import numpy as np


def expected_improvement(X, surrogate, X_sample, Y_sample, correct_label, classify, USE_SKLEARN=True, VERBOSE=True,
                         boundary_penalty=lambda a, b: 0):
    """Pointwise EI objective, same arguments / return convention as the reference: returns ``-EI/len(X_sample)``
    for a single point, ``-0`` when the k-NN router sends the point to another cluster."""
    if not USE_SKLEARN:
        raise NotImplementedError("the GPy branch of the reference is disabled (USE_SKLEARN=True, "
                                  "REF/cGP/cGP_constrained.py:132)")
    my_X = np.asarray(X, dtype=np.float64).reshape(1, -1)
    if classify.predict(my_X) != np.asarray(correct_label).astype(int):
        return -0
    mu, sigma = surrogate.predict(my_X, return_std=True, return_cov=False)
    mu, sigma = np.asarray(mu, dtype=np.float64).reshape(-1), np.asarray(sigma, dtype=np.float64).reshape(-1)
    from math import erfc, exp, pi, sqrt

    ei = 0.0
    if sigma[0] > 0.0:
        imp = mu[0] - np.max(Y_sample)
        Z = imp / sigma[0]
        ei = imp * 0.5 * erfc(-Z / sqrt(2.0)) + sigma[0] * exp(-0.5 * Z * Z) / sqrt(2.0 * pi)
    ei = float(ei + boundary_penalty(my_X, X_sample))
    if VERBOSE:
        print("EI=", ei, "\n")
    return -ei / X_sample.shape[0]


def expected_improvement_batch(model, candidates):
    """Scores of a whole candidate array through the fused GPU scan: ``model`` is a fitted CGPRegressor."""
    return model.score_candidates(candidates)["score"]


def mean_square_prediction_error(X, surrogate, X_sample, Y_sample, correct_label, classify, USE_SKLEARN=True,
                                 VERBOSE=True, boundary_penalty=lambda a, b: 0):
    """Second acquisition of the reference (REF/cGP/acquisition.py:34-75), same signature and return convention:
    joint posterior covariance of [x; X_sample] from ``surrogate.predict(return_cov=True)``, then
    ``(s_xx - s_xo S_oo s_xo^T + penalty) / len(X_sample)`` — to be minimised, one point per call."""
    if not USE_SKLEARN:
        raise NotImplementedError("the GPy branch of the reference is disabled (USE_SKLEARN=True)")
    my_X = np.asarray(X, dtype=np.float64).reshape(1, -1)
    if classify.predict(my_X) != np.asarray(correct_label).astype(int):
        return -0
    joint = np.vstack((my_X, np.asarray(X_sample, dtype=np.float64)))
    _, cov = surrogate.predict(joint, return_std=False, return_cov=True)
    s_xx = cov[0:1, 0:1]
    s_oo = cov[1:, 1:]
    s_xo = cov[0:1, 1:]
    mspe = float((s_xx - s_xo @ s_oo @ s_xo.T).ravel()[0]) + boundary_penalty(my_X, X_sample)
    if VERBOSE:
        print("MSPE=", mspe, "\n")
    return mspe / X_sample.shape[0]


def mean_square_prediction_error_batch(model, candidates):
    """``mspe / n_c`` of a whole candidate array through the dense GPU scan (``model`` is a fitted CGPRegressor); equals
    ``mean_square_prediction_error`` point by point for candidates routed to the cluster they are scored in."""
    return model.mspe_candidates(candidates)
