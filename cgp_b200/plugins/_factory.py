"""Shared constructors of the cluster plugins.  Clustering stays on the host (scikit-learn), as in the reference."""


def dgm(n_components, seed):
    from sklearn.mixture import BayesianGaussianMixture

    return BayesianGaussianMixture(weight_concentration_prior_type="dirichlet_process", n_components=n_components,
                                   random_state=seed)


def kmeans(n_clusters, seed):
    from sklearn.cluster import KMeans

    return KMeans(n_clusters=n_clusters, random_state=seed)
