"""-C plugin: KMeans with 3 clusters (REF/cGP/KMeans3.py)."""
from cgp_b200.plugins._factory import kmeans


def f_Cluster(RND_SEED):
    return kmeans(3, RND_SEED)
