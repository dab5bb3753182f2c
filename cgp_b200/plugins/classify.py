"""-N plugin: ``f_Classify()`` -> object with ``fit(X, labels)`` / ``predict(X)`` (REF/cGP/classify.py:3-9),
here the GPU k-NN router."""
from cgp_b200.gpr import KNeighborsClassifier


def f_Classify():
    return KNeighborsClassifier(n_neighbors=3)
