"""Example -f black-box plugin with the reference's contract (REF/cGP/f_truth.py:5-66, README.md:76-90):
EXAMPLE_NAME, f_truth(X), boundary_penalty(X, data_X), get_bounds(restrict), censor_function(Y).
Bukin N.6 on [-15,-5] x [-3,3]; an optional candidate_mask(Q) replaces the scipy constraint objects."""
import numpy as np

EXAMPLE_NAME = "BUKIN_N6"


def f_truth(X):
    X = np.asarray(X, dtype=np.float64).reshape(1, -1)
    return (100.0 * np.sqrt(np.abs(X[:, 1] - 0.01 * X[:, 0] ** 2)) + 0.01 * np.abs(X[:, 0] + 10.0))[0]


def boundary_penalty(X, data_X=None):
    return 0


def get_bounds(restrict):
    return np.array([[-15.0, -5.0], [-3.0, 3.0]])


def censor_function(Y):
    return Y
