"""-C plugin: KMeans with 4 clusters (REF/cGP/KMeans4.py)."""
from cgp_b200.plugins._factory import kmeans


def f_Cluster(RND_SEED):
    return kmeans(4, RND_SEED)
