"""Interop with model files written by the reference (REF/cGP/cGP_constrained.py:538-545: a dill pickle
``{'Cluster', 'Classify', 'GPR_list', ...}`` of scikit-learn objects; consumer REF/cGP/f_truth_dill.py:17-30).

Files written with old scikit-learn versions (the shipped ``BUKIN.pkl`` is 0.24.1) reference private modules that no
longer exist (``sklearn.neighbors._kd_tree`` / ``_dist_metrics``); they are loaded with a permissive unpickler and
converted to the B200 objects, which only need the stored arrays (training inputs, normalised targets, theta)."""
from __future__ import annotations

import warnings

import numpy as np


class _Opaque:
    """Stand-in for pickled classes that no longer exist (KD-tree internals): keeps the state, does nothing."""

    def __init__(self, *a, **k):
        pass

    def __setstate__(self, state):
        self.state = state


def _permissive_load(path):
    import dill

    gone = ("sklearn.neighbors._kd_tree", "sklearn.neighbors._ball_tree", "sklearn.neighbors._dist_metrics",
            "sklearn.neighbors.kd_tree", "sklearn.neighbors.dist_metrics")

    class Unpickler(dill.Unpickler):
        def find_class(self, module, name):
            if module.startswith(gone):
                if name.startswith("newObj"):
                    return lambda cls, *a: cls.__new__(cls)
                return _Opaque
            return super().find_class(module, name)

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        with open(path, "rb") as f:
            return Unpickler(f).load()


def extract_arrays(obj):
    """Plain arrays of a reference model dict: per GPR (X, y_norm, y_mean, y_std, theta, alpha, kernel) and the router's
    (X, labels, k).  No GPU needed."""
    gprs = []
    for g in obj["GPR_list"]:
        if g is None:
            gprs.append(None)
            continue
        d = g.__dict__
        gprs.append(dict(X=np.asarray(d["X_train_"], dtype=np.float64),
                         y_norm=np.asarray(d["y_train_"], dtype=np.float64).reshape(-1),
                         y_mean=float(np.asarray(d["_y_train_mean"]).reshape(-1)[0]),
                         y_std=float(np.asarray(d["_y_train_std"]).reshape(-1)[0]),
                         theta=np.asarray(d["kernel_"].theta, dtype=np.float64), alpha=float(d.get("alpha", 1e-10)),
                         kernel=d["kernel_"], lml=float(d.get("log_marginal_likelihood_value_", np.nan))))
    c = obj["Classify"].__dict__
    clf = dict(X=np.asarray(c["_fit_X"], dtype=np.float64), labels=np.asarray(c["classes_"])[np.asarray(c["_y"]).astype(int)],
               k=int(c["n_neighbors"]))
    return dict(gprs=gprs, classify=clf)


def model_from_arrays(arr, cluster=None):
    """B200-backed {'Cluster', 'Classify', 'GPR_list'} from the plain arrays of extract_arrays (needs a CUDA device).
    A GPR entry needs X, y_norm, y_mean, y_std, theta, alpha and either ``kernel`` (the fitted sklearn kernel object)
    or ``kind`` ("matern32" | "rbf": the reference's template REF/cGP/cGP_constrained.py:169 is rebuilt)."""
    from .gpr import GaussianProcessRegressor, KNeighborsClassifier

    out = {"Cluster": cluster, "GPR_list": []}
    out["Classify"] = KNeighborsClassifier(arr["classify"]["k"]).fit(arr["classify"]["X"], arr["classify"]["labels"])
    for a in arr["gprs"]:
        if a is None:
            out["GPR_list"].append(None)
            continue
        kern = a.get("kernel")
        if kern is None:
            from sklearn.gaussian_process.kernels import RBF, Matern, WhiteKernel

            d = a["X"].shape[1]
            base = Matern(length_scale=np.ones(d), length_scale_bounds=(1e-05, 100000.0), nu=1.5) \
                if a.get("kind", "matern32") == "matern32" else RBF(length_scale=np.ones(d), length_scale_bounds=(1e-05, 100000.0))
            kern = base + WhiteKernel(noise_level=1.0, noise_level_bounds=(1e-05, 100000.0))
        kern = kern.clone_with_theta(a["theta"])
        g = GaussianProcessRegressor(kernel=kern, alpha=a["alpha"], normalize_y=True, optimizer=None, random_state=0)
        g.fit(a["X"], (a["y_norm"] * a["y_std"] + a["y_mean"]).reshape(-1, 1))  # one factorisation at the stored theta
        out["GPR_list"].append(g)
    return out


def load_reference_model(path):
    """-> {'Cluster', 'Classify', 'GPR_list'} with B200-backed objects (needs a CUDA device for GPR_list / predict)."""
    obj = _permissive_load(path)
    return model_from_arrays(extract_arrays(obj), obj.get("Cluster"))


def f_truth(model, X):
    """REF/cGP/f_truth_dill.py:20-30: route one point with the k-NN classifier, return that cluster's posterior mean."""
    x = np.asarray(X, dtype=np.float64).reshape(1, -1)
    lab = int(model["Classify"].predict(x)[0])
    mu = model["GPR_list"][lab].predict(x, return_std=True, return_cov=False)[0]
    return float(np.asarray(mu).reshape(-1)[0])
