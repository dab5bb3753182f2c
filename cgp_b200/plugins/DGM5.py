"""-C plugin: Dirichlet-process Gaussian mixture with at most 5 components (REF/cGP/DGM5.py)."""
from cgp_b200.plugins._factory import dgm


def f_Cluster(RND_SEED):
    return dgm(5, RND_SEED)
