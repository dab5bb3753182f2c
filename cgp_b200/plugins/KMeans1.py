"""-C plugin: KMeans with 1 clusters (REF/cGP/KMeans1.py)."""
from cgp_b200.plugins._factory import kmeans


def f_Cluster(RND_SEED):
    return kmeans(1, RND_SEED)
