"""-C plugin: ``f_Cluster(RND_SEED)`` -> object with ``fit_predict(XY)`` (REF/cGP/cluster.py:3-15; same as DGM3)."""
from cgp_b200.plugins._factory import dgm


def f_Cluster(RND_SEED):
    return dgm(3, RND_SEED)
