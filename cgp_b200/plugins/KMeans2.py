"""-C plugin: KMeans with 2 clusters (REF/cGP/KMeans2.py)."""
from cgp_b200.plugins._factory import kmeans


def f_Cluster(RND_SEED):
    return kmeans(2, RND_SEED)
