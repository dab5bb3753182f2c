"""-C plugin: KMeans with 5 clusters (REF/cGP/KMeans5.py)."""
from cgp_b200.plugins._factory import kmeans


def f_Cluster(RND_SEED):
    return kmeans(5, RND_SEED)
