"""Sequential-sampling driver on the B200 backend.

Same surface as the reference drivers — the callable class ``cGP_constrained(options).run()`` with its ``*_CGP`` option
keys and overridable hooks (REF/cGP/cGP_constrained_module.py:61-251) and the ``cGP.py`` command-line flags
(REF/cGP/cGP.py:7-43) — with the per-cluster body of the loop (REF/cGP/cGP_constrained.py:305-379) replaced by ONE
batched GPU fit of all clusters x restarts and ONE dense EI scan of a candidate set (the reference runs one
trust-constr search per cluster, one point per call).  Clustering, the black box, merging / recoding of labels, the
exploration coin and all bookkeeping stay on the host exactly as in the reference.

    python -m cgp_b200.driver -p 20 -s 10 -c 3 -n 3 -f f_truth --candidates 65536
"""
from __future__ import annotations

import importlib
import os
import random
import string
import sys
from datetime import datetime

import numpy as np


class cGP_constrained(object):
    """Callable interface (option keys and hook names of REF/cGP/cGP_constrained_module.py:61-251)."""

    def __init__(self, options):
        g = options.get
        self.N_PILOT = g("N_PILOT_CGP", 20)
        self.N_SEQUENTIAL = g("N_SEQUENTIAL_CGP", 20)
        self.RND_SEED = g("RND_SEED_CGP", 1)
        self.EXAMPLE_NAME = g("EXAMPLE_NAME_CGP", "app_name")
        self.METHOD = g("METHOD_CGP", "FREQUENTIST")
        self.EXPLORATION_RATE = g("EXPLORATION_RATE_CGP", 1.0)
        self.NO_CLUSTER = g("NO_CLUSTER_CGP", False)
        self.N_NEIGHBORS = g("N_NEIGHBORS_CGP", 3)
        self.CLUSTER_METHOD = g("CLUSTER_METHOD_CGP", "BGM")
        self.N_COMPONENTS = g("N_COMPONENTS_CGP", 3)
        self.ACQUISITION = g("ACQUISITION_CGP", "EI")
        # additions of the GPU backend: size of the dense candidate scan and restart count (default 10*d, :412-413)
        self.N_CANDIDATES = g("N_CANDIDATES_CGP", 65536)
        self.N_RESTARTS = g("N_RESTARTS_CGP", None)
        # incremental refit: when the clustering keeps the labels of the old samples, append the
        # new sample to its cluster and re-optimise from the previous theta instead of refitting from scratch
        self.INCREMENTAL = g("INCREMENTAL_CGP", False)
        # exact bound-pruned EI scan (cgp_ei_argmax_pruned, default): same selected sample, several times faster acquisition;
        # False = score every candidate
        self.PRUNED_SCAN = g("PRUNED_SCAN_CGP", True)
        self.PILOT_SAMPLES = g("PILOT_SAMPLES_CGP", None)
        self.OUTPUT_NAME = g("OUTPUT_NAME_CGP", None)
        self.OBJ_NAME = g("OBJ_NAME_CGP", None)
        self.VERBOSE = g("VERBOSE_CGP", False)
        self.REPEAT_SAMPLE = False
        self.ALPHA_SKLEARN = 1e-5
        self.SKLEARN_normalizer = True
        if self.METHOD != "FREQUENTIST":
            raise NotImplementedError("only METHOD_CGP='FREQUENTIST' exists on the scikit-learn path of the reference "
                                      "(REF/cGP/cGP_constrained_module.py:423)")
        if self.ACQUISITION not in ("EI", "MSPE"):
            raise NotImplementedError("ACQUISITION_CGP must be 'EI' or 'MSPE' (REF/cGP/acquisition.py)")

    # ---- hooks (same names / meaning as the reference) ------------------------------------------------------------
    def f_Classify(self):
        from .gpr import KNeighborsClassifier

        return KNeighborsClassifier(self.N_NEIGHBORS)

    def f_Cluster(self, RND_SEED):
        from .plugins import _factory

        if self.CLUSTER_METHOD == "KMeans":
            return _factory.kmeans(self.N_COMPONENTS, RND_SEED)
        return _factory.dgm(self.N_COMPONENTS, RND_SEED)

    def boundary_penalty(self, X, data_X=None):
        return 0

    def get_bounds(self, restrict):
        return np.array([[-1, 1], [-1, 1]]).astype(float)

    def censor_function(self, Y):
        return Y

    def candidate_mask(self, Q):
        """Extra hard constraints as a candidate filter (a hook of this backend).  Return a bool array [M]; default keeps
        all.  The reference's own constraint objects are honoured through get_linear_constraint /
        get_nonlinear_constraint / get_bounds (see constraint_mask)."""
        return None

    def get_linear_constraint(self):
        """scipy.optimize.LinearConstraint or None (REF/cGP/cGP_constrained_module.py:220-222)."""
        return None

    def get_nonlinear_constraint(self):
        """scipy.optimize.NonlinearConstraint or None (REF/cGP/cGP_constrained_module.py:202-206)."""
        return None

    def constraint_mask(self, Q, bounds):
        """The hard constraints the reference hands to trust-constr (REF/cGP/cGP_constrained.py:362-366:
        ``constraints=[linear_constraint, nonlinear_constraint], bounds=bounds_constraint``) as a filter of the dense
        candidate set: keep q iff  lb <= A q <= ub,  lb <= fun(q) <= ub  and  bounds.lb <= q <= bounds.ub.
        ``fun`` is tried vectorised (columns = candidates, the SciPy convention) and verified on a few points against
        point-wise calls; a function that only works on one point is called per candidate."""
        keep = np.ones(Q.shape[0], dtype=bool)
        lin, nonlin = self.get_linear_constraint(), self.get_nonlinear_constraint()
        if lin is not None:
            A = np.atleast_2d(np.asarray(lin.A, dtype=np.float64))
            v = Q @ A.T
            keep &= np.all((v >= np.asarray(lin.lb, dtype=np.float64)) & (v <= np.asarray(lin.ub, dtype=np.float64)), axis=1)
        if nonlin is not None:
            fun = nonlin.fun
            lb, ub = np.atleast_1d(np.asarray(nonlin.lb, dtype=np.float64)), np.atleast_1d(np.asarray(nonlin.ub, dtype=np.float64))
            if not (np.all(np.isneginf(lb)) and np.all(np.isposinf(ub))):
                point = lambda q: np.atleast_1d(np.asarray(fun(q), dtype=np.float64))  # noqa: E731
                vals = None
                try:
                    v = np.asarray(fun(Q.T), dtype=np.float64)
                    v = v.reshape(1, -1) if v.ndim <= 1 else v
                    if v.shape[1] == Q.shape[0] and v.size and all(
                            np.allclose(v[:, j], point(Q[j]), rtol=1e-12, atol=0.0) for j in range(0, Q.shape[0], max(1, Q.shape[0] // 8))):
                        vals = v.T
                except Exception:  # noqa: BLE001
                    vals = None
                if vals is None:
                    vals = np.stack([point(q) for q in Q])
                keep &= np.all((vals >= lb) & (vals <= ub), axis=1)
        keep &= np.all((Q >= bounds[:, 0]) & (Q <= bounds[:, 1]), axis=1)
        return keep

    def recode(self, label):
        level = np.unique(np.array(label))
        out = np.array(label).copy()
        for ck, j in enumerate(level):
            out[np.array(label) == j] = ck
        return out

    def f_truth(self, X):
        X = X.reshape(1, -1)
        return 1 / (1 + (X[:, 0] - .25) ** 2 + (X[:, 1] - .25) ** 2)

    def get_kernel(self, d):
        """Matern(nu=3/2, ARD) + WhiteKernel, REF/cGP/cGP_constrained_module.py:289."""
        from sklearn.gaussian_process.kernels import Matern, WhiteKernel

        return Matern(length_scale=np.ones(d), length_scale_bounds=(1e-05, 100000.0), nu=3 / 2) + \
            WhiteKernel(noise_level=1.0, noise_level_bounds=(1e-05, 100000.0))

    def get_random_string(self, length):
        return "".join(random.choice(string.ascii_lowercase) for _ in range(length))

    # ---- one iteration pieces --------------------------------------------------------------------------------------
    def _labels(self, X, Y, dgm):
        d = X.shape[1]
        if self.NO_CLUSTER:
            return np.zeros(X.shape[0], dtype=np.int64)
        XY = np.concatenate((X, Y.reshape(-1, 1)), axis=1)
        lab = np.asarray(dgm.fit_predict(XY)).astype(np.int64)
        for c in np.unique(lab):  # REF/cGP/cGP_constrained.py:279-282
            if np.sum(lab == c) <= d + 2:
                lab[lab == c] = np.argmax(np.bincount(lab))
        return self.recode(lab)

    def _candidates(self, bounds):
        Q = np.random.uniform(bounds[:, 0], bounds[:, 1], size=(int(self.N_CANDIDATES), bounds.shape[0]))
        keep = self.constraint_mask(Q, bounds)
        mask = self.candidate_mask(Q)
        if mask is not None:
            keep &= np.asarray(mask, dtype=bool)
        if not np.all(keep):
            Q = Q[keep]
            if Q.shape[0] == 0:
                raise RuntimeError("the constraints (linear / nonlinear / bounds / candidate_mask) reject every candidate")
        return Q

    def run(self):
        from .cgp import CGPRegressor

        np.random.seed(self.RND_SEED)
        start = datetime.now()
        bounds = self.get_bounds(1)
        d = bounds.shape[0]
        if self.PILOT_SAMPLES is not None:  # -F: X-only whitespace file or array (REF/cGP/cGP_constrained.py:185-188)
            X = np.loadtxt(self.PILOT_SAMPLES) if isinstance(self.PILOT_SAMPLES, str) else np.asarray(self.PILOT_SAMPLES, float)
            X = X.reshape(-1, d)
        else:
            X = np.zeros((self.N_PILOT, d))
            for j in range(d):  # column by column like :191-193
                X[:, j] = np.random.uniform(bounds[j, 0], bounds[j, 1], size=(self.N_PILOT, 1)).ravel()
        Y = np.zeros((X.shape[0], 1))
        for k in range(X.shape[0]):
            Y[k, 0] = np.asarray(self.censor_function(self.f_truth(X[k, :].reshape(1, -1))), dtype=np.float64).reshape(-1)[0]
        dgm = self.f_Cluster(self.RND_SEED)
        model, labels = None, None
        fitted_n, fitted_labels = 0, None  # what `model` currently holds
        self.history = []
        for it in range(self.N_SEQUENTIAL):
            labels = self._labels(X, Y, dgm)
            coin = np.random.uniform(0, 1, 1)
            if coin < self.EXPLORATION_RATE:
                R = 10 * d if self.N_RESTARTS is None else int(self.N_RESTARTS)
                n = X.shape[0]
                if (self.INCREMENTAL and model is not None and fitted_n < n and int(labels.max()) + 1 == model.n_clusters_
                        and np.array_equal(labels[:fitted_n], fitted_labels)):
                    for k in range(fitted_n, n):  # samples drawn since the last fit (random-search steps included)
                        model.partial_fit(X[k], Y[k, 0], labels[k], reoptimize=(k == n - 1), n_restarts=0)
                    mode = "acquisition+incremental"
                else:
                    model = CGPRegressor(kernel=self.get_kernel(d), alpha=self.ALPHA_SKLEARN, n_restarts_optimizer=R,
                                         normalize_y=self.SKLEARN_normalizer, random_state=0,
                                         n_neighbors=min(self.N_NEIGHBORS, X.shape[0]))
                    model.fit(X, Y.reshape(-1), labels)
                    mode = "acquisition"
                fitted_n, fitted_labels = n, labels.copy()
                Q = self._candidates(bounds)
                if self.ACQUISITION == "MSPE":  # minimised, (mspe + boundary_penalty) / n_c, REF/cGP/acquisition.py:69-75
                    X_next, score, idx = model.select_next_mspe(Q, boundary_penalty=self.boundary_penalty)
                else:  # soft constraint inside the GPU score: (EI + boundary_penalty(x, X_c)) / n_c (REF/cGP/acquisition.py:29-32)
                    X_next, score, idx = model.select_next(Q, prune=bool(self.PRUNED_SCAN), boundary_penalty=self.boundary_penalty)
                self.history.append(dict(step=it, mode=mode, score=float(score), n_clusters=int(labels.max()) + 1))
            else:  # random search inside the bounds (:380-412)
                X_next = np.random.uniform(bounds[:, 0], bounds[:, 1])
                while not self.REPEAT_SAMPLE and np.any(np.all(X == X_next[None, :], axis=1)):
                    X_next = np.random.uniform(bounds[:, 0], bounds[:, 1])
                self.history.append(dict(step=it, mode="random", n_clusters=int(labels.max()) + 1))
            X_next = np.asarray(X_next, dtype=np.float64).reshape(1, -1)
            Y_next = np.asarray(self.f_truth(X_next)).reshape(1, -1)
            if self.VERBOSE:
                print(">>Next sample input is chosen to be: ", X_next, " response: ", Y_next.ravel())
            X = np.vstack((X, X_next))
            Y = np.vstack((Y, np.asarray(self.censor_function(Y_next)).reshape(1, -1)))
        self.model_, self.labels_, self.cluster_ = model, labels, dgm
        self._write_outputs(X, Y, dgm, start, datetime.now())
        return X, Y

    # ---- outputs: same table / log / pickle layout as REF/cGP/cGP_constrained.py:443-545 -----------------------------
    def _write_outputs(self, X, Y, dgm, start, end):
        if self.OUTPUT_NAME is None and self.OBJ_NAME is None:
            return
        final = np.zeros((X.shape[0], 1), dtype=np.int64) if self.NO_CLUSTER else \
            np.asarray(dgm.fit_predict(np.concatenate((X, Y), axis=1))).reshape(-1, 1)
        if self.OUTPUT_NAME is not None:
            names = ",".join(["Y"] + ["X%d" % (i + 1) for i in range(X.shape[1])] + ["label"])
            np.savetxt(self.OUTPUT_NAME + ".txt", np.hstack((Y, X, final)), delimiter=",", header=names, comments="",
                       fmt=",".join(["%1.12f"] * (1 + X.shape[1]) + ["%i"]))
            with open(self.OUTPUT_NAME + ".log", "w") as f:
                print("Example: ", self.EXAMPLE_NAME, file=f)
                print("Sample start date and time = ", start, file=f)
                print("Sample end date and time = ", end, file=f)
                print("Python version: ", sys.version, file=f)
                print("Random seed: ", self.RND_SEED, file=f)
                print("Backend: cgp_b200 (dense EI scan, %d candidates per step)" % self.N_CANDIDATES, file=f)
                print("Maximum in sample = ", np.round(Y[np.argmax(Y), :], 12), " attained at ", X[np.argmax(Y), :], file=f)
                print("Minimum in sample = ", np.round(Y[np.argmin(Y), :], 12), " attained at ", X[np.argmin(Y), :], file=f)
        if self.OBJ_NAME is not None and self.model_ is not None:
            import dill

            with open(self.OBJ_NAME + ".pkl", "wb+") as f:  # keys of REF/cGP/cGP_constrained.py:539
                dill.dump({"Cluster": dgm, "Classify": self.model_.classify_, "GPR_list": self.model_.GPR_list}, f)


def _plugin(name, candidates):
    """Import a plugin module by the name the reference's -C/-N/-f flags take (bare module name on sys.path, or one of
    the modules shipped in cgp_b200.plugins)."""
    for mod in candidates:
        try:
            return importlib.import_module(mod)
        except ModuleNotFoundError:
            continue
    raise ModuleNotFoundError(name)


def main(argv=None):
    from optparse import OptionParser

    p = OptionParser("Usage: python -m cgp_b200.driver -p <N_PILOT> -s <N_SEQUENTIAL> [options]  (flags of REF/cGP/cGP.py)")
    p.add_option("-p", "--N_PILOT", type="int", dest="N_PILOT")
    p.add_option("-s", "--N_SEQUENTIAL", type="int", dest="N_SEQUENTIAL")
    p.add_option("-c", "--N_COMPONENT", type="int", dest="N_COMPONENT", default=4)
    p.add_option("-C", "--CLUSTER", type="string", dest="CLUSTER", default=None)
    p.add_option("-n", "--N_NEIGHBORS", type="int", dest="N_NEIGHBORS", default=3)
    p.add_option("-N", "--CLASSIFY", type="string", dest="CLASSIFY", default=None)
    p.add_option("-e", "--EXPLORATION_RATE", type="float", dest="EXPLORATION_RATE", default=1.0)
    p.add_option("-g", "--NO_CLUSTER", type="int", dest="NO_CLUSTER", default=0)
    p.add_option("-f", "--f_truth", type="string", dest="f_truth_py", default="f_truth")
    p.add_option("-r", "--RND_SEED", type="int", dest="RND_SEED", default=123)
    p.add_option("-o", "--OUTPUT_NAME", type="string", dest="OUTPUT_NAME", default=None)
    p.add_option("-d", "--OBJ_NAME", type="string", dest="OBJ_NAME", default=None)
    p.add_option("-F", "--PILOT_FILE", type="string", dest="PILOT_FILE", default=None)
    p.add_option("-A", "--ACQUISITION", type="string", dest="ACQUISITION", default="expected_improvement")
    p.add_option("-X", "--QUIET", action="store_true", dest="QUIET")
    p.add_option("--candidates", type="int", dest="CANDIDATES", default=65536)
    p.add_option("--restarts", type="int", dest="RESTARTS", default=None)
    p.add_option("--incremental", action="store_true", dest="INCREMENTAL", default=False)
    p.add_option("--pruned-scan", action="store_true", dest="PRUNED", default=True)
    p.add_option("--exhaustive-scan", action="store_false", dest="PRUNED")
    o, _ = p.parse_args(argv)
    if o.N_SEQUENTIAL is None:
        raise Exception("ERROR: N_SEQEUNTIAL(int) is a compulsory parameter that must be supplied. \n e.g. -s 20")
    if o.PILOT_FILE is None and o.N_PILOT is None:
        raise Exception("ERROR: Either N_PILOT(int) or PILOT_FILE(string, with extension) must be supplied.")
    if o.ACQUISITION not in ("expected_improvement", "mean_square_prediction_error"):
        raise NotImplementedError("-A must be expected_improvement or mean_square_prediction_error (REF/cGP/acquisition.py)")
    sys.path.insert(0, os.getcwd())
    ft = importlib.import_module(o.f_truth_py)  # -f: module with f_truth / get_bounds / censor_function / ... (:139-145)
    stamp = "".join(random.choice(string.ascii_lowercase) for _ in range(8))
    name = getattr(ft, "EXAMPLE_NAME", "app_name")
    out = o.OUTPUT_NAME or (name + ("_local_GP(" if o.NO_CLUSTER else "_local_cGP_k=%d(" % o.N_COMPONENT) + stamp + ")")

    class Run(cGP_constrained):
        def f_truth(self, X):
            return ft.f_truth(X)

        def get_bounds(self, restrict):
            return ft.get_bounds(restrict)

        def censor_function(self, Y):
            return ft.censor_function(Y) if hasattr(ft, "censor_function") else Y

        def boundary_penalty(self, X, data_X=None):
            return ft.boundary_penalty(X, data_X) if hasattr(ft, "boundary_penalty") else 0

        def candidate_mask(self, Q):
            return ft.candidate_mask(Q) if hasattr(ft, "candidate_mask") else None

        # constraint objects a reference -f module exports at module level (REF/cGP/f_truth.py:44-59)
        def get_linear_constraint(self):
            return getattr(ft, "linear_constraint", None)

        def get_nonlinear_constraint(self):
            return getattr(ft, "nonlinear_constraint", None)

        def f_Cluster(self, seed):
            if o.CLUSTER is not None:  # -C <module>: f_Cluster(RND_SEED) (REF/cGP/cluster.py:3-15)
                return _plugin(o.CLUSTER, [o.CLUSTER, "cgp_b200.plugins." + o.CLUSTER]).f_Cluster(seed)
            from .plugins import _factory

            return _factory.dgm(o.N_COMPONENT, 0)  # built-in uses random_state=0 (REF/cGP/cGP_constrained.py:110-114)

        def f_Classify(self):
            if o.CLASSIFY is not None:  # -N <module>: f_Classify() (REF/cGP/classify.py:3-9)
                return _plugin(o.CLASSIFY, [o.CLASSIFY, "cgp_b200.plugins." + o.CLASSIFY]).f_Classify()
            return cGP_constrained.f_Classify(self)

    run = Run({"N_PILOT_CGP": o.N_PILOT, "N_SEQUENTIAL_CGP": o.N_SEQUENTIAL, "RND_SEED_CGP": o.RND_SEED,
               "EXAMPLE_NAME_CGP": name, "EXPLORATION_RATE_CGP": o.EXPLORATION_RATE, "NO_CLUSTER_CGP": bool(o.NO_CLUSTER),
               "N_NEIGHBORS_CGP": o.N_NEIGHBORS, "N_COMPONENTS_CGP": o.N_COMPONENT, "N_CANDIDATES_CGP": o.CANDIDATES,
               "N_RESTARTS_CGP": o.RESTARTS, "INCREMENTAL_CGP": o.INCREMENTAL, "PRUNED_SCAN_CGP": o.PRUNED,
               "ACQUISITION_CGP": "MSPE" if o.ACQUISITION == "mean_square_prediction_error" else "EI", "PILOT_SAMPLES_CGP": o.PILOT_FILE, "OUTPUT_NAME_CGP": out,
               "OBJ_NAME_CGP": o.OBJ_NAME, "VERBOSE_CGP": not o.QUIET})
    X, Y = run.run()
    if not o.QUIET:
        print("Maximum in sample:", float(Y.max()), "at", X[int(np.argmax(Y))])
        print("outputs:", out + ".txt", out + ".log", (o.OBJ_NAME + ".pkl") if o.OBJ_NAME else "")
    return X, Y


if __name__ == "__main__":
    main()
