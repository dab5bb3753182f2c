"""-C plugin: Dirichlet-process Gaussian mixture with at most 4 components (REF/cGP/DGM4.py)."""
from cgp_b200.plugins._factory import dgm


def f_Cluster(RND_SEED):
    return dgm(4, RND_SEED)
