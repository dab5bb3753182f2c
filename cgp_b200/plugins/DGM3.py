"""-C plugin: Dirichlet-process Gaussian mixture with at most 3 components (REF/cGP/DGM3.py)."""
from cgp_b200.plugins._factory import dgm


def f_Cluster(RND_SEED):
    return dgm(3, RND_SEED)
