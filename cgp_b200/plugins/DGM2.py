"""-C plugin: Dirichlet-process Gaussian mixture with at most 2 components (REF/cGP/DGM2.py)."""
from cgp_b200.plugins._factory import dgm


def f_Cluster(RND_SEED):
    return dgm(2, RND_SEED)
