"""Quality evaluators for the Hogwild acceptance tests: community NMI and node-classification
micro-F1 on the learned node table, and the o1 / o2 objectives.  The objectives run on the device (Node2Vec.loss,
comemb_o2_pos_loss; pinned against reference-generated values in tests/golden/golden_losses.json); `nmi` and
`community_nmi(method="device")` keep the node table on the device as well (k-means assignment + contingency table in
torch); the micro-F1 protocol (logistic regression on a train split) runs through sklearn on the host or, with
method="device", as an L-BFGS fit in torch on the table's device.  Not part of the SGD path."""
import numpy as np


def _np(x):
    return x.detach().cpu().numpy() if hasattr(x, "detach") else np.asarray(x)


def nmi(labels_true, labels_pred):
    """Normalised mutual information (arithmetic-mean normalisation, sklearn's default) of two labelings given as integer
    tensors / arrays -- computed with torch on whatever device `labels_pred` lives on (contingency table by bincount)."""
    import torch
    b = labels_pred if isinstance(labels_pred, torch.Tensor) else torch.as_tensor(np.asarray(labels_pred))
    a = torch.as_tensor(np.asarray(_np(labels_true))).to(b.device)
    a = torch.unique(a.long(), return_inverse=True)[1]
    b = torch.unique(b.long(), return_inverse=True)[1]
    ka, kb, n = int(a.max()) + 1, int(b.max()) + 1, a.numel()
    cont = torch.bincount(a * kb + b, minlength=ka * kb).reshape(ka, kb).double()
    pa, pb, pab = cont.sum(1) / n, cont.sum(0) / n, cont / n
    nz = pab > 0
    mi = (pab[nz] * (pab[nz].log() - (pa[:, None] * pb[None, :])[nz].log())).sum()
    ha = -(pa[pa > 0] * pa[pa > 0].log()).sum()
    hb = -(pb[pb > 0] * pb[pb > 0].log()).sum()
    denom = 0.5 * (ha + hb)
    if float(denom) <= 0.0:
        return 1.0  # both labelings are constant
    return float((mi / denom).clamp(min=0.0))


def community_nmi(embedding, labels, k=None, method="gmm", seed=0):
    """NMI between ground-truth communities and a clustering of the node table.  method="gmm" mirrors the reference's
    community assignment (GaussianMixture, community_embeddings.py:16-37, diagonal covariance for d >> n robustness);
    "kmeans" is the cheaper proxy used on large graphs; "device" = k-means++/Lloyd assignment and the NMI itself on the
    device the table lives on (nothing is copied to the host)."""
    labels = np.asarray(_np(labels))
    k = int(k or len(np.unique(labels)))
    if method == "device":
        import torch
        from .ADSCModel.gmm_device import DeviceGaussianMixture
        x = embedding if isinstance(embedding, torch.Tensor) else torch.as_tensor(np.asarray(embedding))
        gen = torch.Generator(device=x.device)
        gen.manual_seed(int(seed))
        one = DeviceGaussianMixture(n_components=k, kmeans_iter=30)._kmeans_resp(x.float(), gen)
        return nmi(labels, one.argmax(1))
    from sklearn.metrics import normalized_mutual_info_score
    x = _np(embedding).astype(np.float64)
    if method == "kmeans":
        from sklearn.cluster import KMeans
        pred = KMeans(k, n_init=5, random_state=seed).fit_predict(x)
    else:
        from sklearn.mixture import GaussianMixture
        pred = GaussianMixture(n_components=k, covariance_type="diag", n_init=5, reg_covar=1e-5,
                               random_state=seed).fit_predict(x)
    return float(normalized_mutual_info_score(labels, pred))


def node_classification_micro_f1(embedding, labels, train_fraction=0.5, seed=0, method="sklearn", l2=1.0, max_iter=100):
    """Logistic regression on a random train split, micro-F1 on the rest (the DeepWalk protocol).  method="sklearn":
    host LogisticRegression; method="device": the same L2-regularised multinomial model (C = 1/l2, intercept not
    penalised) fitted by L-BFGS in torch on the device the table lives on -- the table is never copied to the host.
    With one label per node micro-F1 is the accuracy on the test split; both methods use the same split."""
    y = np.asarray(_np(labels))
    rs = np.random.RandomState(seed)
    perm = rs.permutation(len(y))
    cut = int(train_fraction * len(y))
    tr, te = perm[:cut], perm[cut:]
    if method == "device":
        import torch
        x = embedding if isinstance(embedding, torch.Tensor) else torch.as_tensor(np.asarray(embedding))
        x = x.detach().double()
        dev = x.device
        classes, yi = np.unique(y, return_inverse=True)
        yt = torch.from_numpy(yi).to(dev)
        tr_t, te_t = torch.from_numpy(tr).to(dev), torch.from_numpy(te).to(dev)
        xtr, ytr = x[tr_t], yt[tr_t]
        w = torch.zeros((x.shape[1], len(classes)), dtype=torch.float64, device=dev, requires_grad=True)
        b = torch.zeros(len(classes), dtype=torch.float64, device=dev, requires_grad=True)
        opt = torch.optim.LBFGS([w, b], lr=1.0, max_iter=max_iter, tolerance_grad=1e-6, tolerance_change=1e-10,
                                history_size=10, line_search_fn="strong_wolfe")

        def closure():  # sklearn's objective: C * sum_i loss_i + 0.5 ||w||^2 with C = 1 / l2
            opt.zero_grad()
            loss = torch.nn.functional.cross_entropy(xtr @ w + b, ytr, reduction="sum") / l2 + 0.5 * (w * w).sum()
            loss.backward()
            return loss

        opt.step(closure)
        with torch.no_grad():
            pred = (x[te_t] @ w + b).argmax(1)
            return float((pred == yt[te_t]).double().mean())
    from sklearn.linear_model import LogisticRegression
    from sklearn.metrics import f1_score
    x = _np(embedding).astype(np.float64)
    clf = LogisticRegression(max_iter=500).fit(x[tr], y[tr])
    return float(f1_score(y[te], clf.predict(x[te]), average="micro"))


def o2_positive_loss(model, walks, walk_off, window):
    """Mean -log sigmoid(x_j . c_i) over window pairs (device kernel comemb_o2_pos_loss)."""
    from .utils import training_sdg_inner as K
    s, n = K.o2_pos_loss(model.node_embedding, model.context_embedding, walks, walk_off, window)
    return s / max(1, n)


def sgns_objective(node, ctx, walks2d, window, table, negative, seed=0, max_centres=200000):
    """The skip-gram negative-sampling objective the o2 path minimises, per (centre, context) pair, with exact sigmoids:
        -log sigma(x_j . c_i) - sum_{t in `negative` draws from `table`} log sigma(-x_j . c_t)
    evaluated in torch on the device the tables live on, over the window pairs of `walks2d` ([n_walks, L] row tokens,
    TOKEN_NONE padded).  Untrained tables (context table zero) give (1 + negative) * ln 2 (less the skipped draws).  Returns (objective per pair,
    positive part per pair, number of pairs).  An evaluator, not part of the SGD path."""
    import torch
    none = 0xFFFFFFFF
    w = walks2d.long() & 0xFFFFFFFF
    n_walks, L = w.shape
    gen = torch.Generator(device=w.device)
    gen.manual_seed(int(seed))
    tot = pos_tot = 0.0
    cnt = 0
    tab = table.long() & 0xFFFFFFFF
    rows_per_chunk = max(1, int(max_centres // L))
    for r0 in range(0, n_walks, rows_per_chunk):
        wc = w[r0:r0 + rows_per_chunk]
        for off in range(1, int(window) + 1):
            for a, b in ((wc[:, off:], wc[:, :-off]), (wc[:, :-off], wc[:, off:])):  # (centre i, context j) both ways
                ok = (a != none) & (b != none)
                ci, xj = a[ok], b[ok]
                if ci.numel() == 0:
                    continue
                x = node[xj].double()
                f = (x * ctx[ci].double()).sum(1)
                p = torch.nn.functional.softplus(-f)
                loss = p.clone()
                t = tab[torch.randint(tab.numel(), (ci.numel(), int(negative)), generator=gen, device=w.device)]
                fn = (x[:, None, :] * ctx[t].double()).sum(2)
                live = t != ci[:, None]  # pyx:135-136: a draw equal to the centre is skipped
                loss += (torch.nn.functional.softplus(fn) * live).sum(1)
                tot += float(loss.sum())
                pos_tot += float(p.sum())
                cnt += int(ci.numel())
    return tot / max(cnt, 1), pos_tot / max(cnt, 1), cnt
