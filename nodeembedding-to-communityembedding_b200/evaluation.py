"""Quality evaluators for the Hogwild acceptance tests: community NMI and node-classification
micro-F1 on the learned node table, and the o1 / o2 objectives.  Host-side (sklearn); not part of the SGD path."""
import numpy as np


def _np(x):
    return x.detach().cpu().numpy() if hasattr(x, "detach") else np.asarray(x)


def community_nmi(embedding, labels, k=None, method="gmm", seed=0):
    """NMI between ground-truth communities and a clustering of the node table.  method="gmm" mirrors the reference's
    community assignment (GaussianMixture, community_embeddings.py:16-37, diagonal covariance for d >> n robustness);
    "kmeans" is the cheaper proxy used on large graphs."""
    from sklearn.metrics import normalized_mutual_info_score
    x = _np(embedding).astype(np.float64)
    labels = np.asarray(labels)
    k = int(k or len(np.unique(labels)))
    if method == "kmeans":
        from sklearn.cluster import KMeans
        pred = KMeans(k, n_init=5, random_state=seed).fit_predict(x)
    else:
        from sklearn.mixture import GaussianMixture
        pred = GaussianMixture(n_components=k, covariance_type="diag", n_init=5, reg_covar=1e-5,
                               random_state=seed).fit_predict(x)
    return float(normalized_mutual_info_score(labels, pred))


def node_classification_micro_f1(embedding, labels, train_fraction=0.5, seed=0):
    """One-vs-rest logistic regression on a random train split, micro-F1 on the rest (the DeepWalk protocol)."""
    from sklearn.linear_model import LogisticRegression
    from sklearn.metrics import f1_score
    x = _np(embedding).astype(np.float64)
    y = np.asarray(labels)
    rs = np.random.RandomState(seed)
    perm = rs.permutation(len(y))
    cut = int(train_fraction * len(y))
    tr, te = perm[:cut], perm[cut:]
    clf = LogisticRegression(max_iter=500).fit(x[tr], y[tr])
    return float(f1_score(y[te], clf.predict(x[te]), average="micro"))


def o2_positive_loss(model, walks, walk_off, window):
    """Mean -log sigmoid(x_j . c_i) over window pairs (device kernel comemb_o2_pos_loss)."""
    from .utils import training_sdg_inner as K
    s, n = K.o2_pos_loss(model.node_embedding, model.context_embedding, walks, walk_off, window)
    return s / max(1, n)
