"""Community2Vec: GMM fit + community-gradient step (/root/reference/ADSCModel/community_embeddings.py:12-77).

`train(nodes, model, beta, chunksize, iter)` runs csrc/o3_community.cu.  `fit` stays sklearn's GaussianMixture on the
host exactly as the reference (:16-37) -- it is the producer of o3's inputs, not part of the SGD path.
"""
import logging as log

import numpy as np

from ..utils import training_sdg_inner as K


class Community2Vec(object):
    def __init__(self, model, lr, reg_covar=0, gmm_backend="sklearn"):
        """gmm_backend: "sklearn" = the reference's host fit (:18, bit-compatible inputs for o3); "device" = the same EM
        on the GPU (gmm_device.DeviceGaussianMixture: library GEMMs, tables never leave HBM)."""
        self.lr = lr
        self.reg_covar = reg_covar
        self.k = model.k
        self.g_mixture = None
        self.gmm_backend = gmm_backend

    def fit(self, model):
        import torch
        log.info("Fitting: {} communities".format(model.k))
        if self.gmm_backend == "device":
            from .gmm_device import DeviceGaussianMixture
            gm = DeviceGaussianMixture(n_components=model.k, reg_covar=self.reg_covar, n_init=10)
            x = model.node_embedding.detach()
            gm.fit(x)
            self.g_mixture = gm
            model.centroid = gm.means_.float().contiguous()
            model.covariance_mat = gm.covariances_.float().contiguous()
            model.inv_covariance_mat = torch.linalg.inv(model.covariance_mat).contiguous()
            model.pi = gm.predict_proba(x).float().contiguous()
            return
        import sklearn.mixture as mixture
        if self.g_mixture is None:
            self.g_mixture = mixture.GaussianMixture(n_components=model.k, reg_covar=self.reg_covar,
                                                     covariance_type='full', n_init=10)
        x = model.node_embedding.detach().cpu().numpy()
        self.g_mixture.fit(x)
        dev = model.node_embedding.device
        model.centroid = torch.from_numpy(self.g_mixture.means_.astype(np.float32)).to(dev)
        cov = self.g_mixture.covariances_.astype(np.float32)
        model.covariance_mat = torch.from_numpy(cov).to(dev)
        model.inv_covariance_mat = torch.from_numpy(np.linalg.inv(cov).astype(np.float32)).to(dev)
        model.pi = torch.from_numpy(self.g_mixture.predict_proba(x).astype(np.float32)).to(dev)

    def loss(self, nodes, model, beta, chunksize=150):
        """sum_i sum_k pi_ik * log N(x_i | mu_k, Sigma_k), scaled by beta/K (the intent of
        community_embeddings.py:40-59, whose own code raises: `model.vocab(x)`)."""
        import torch
        rows = torch.as_tensor([model.vocab[x].index for x in nodes], dtype=torch.int64,
                               device=model.node_embedding.device)
        x = model.node_embedding[rows].double()
        total = 0.0
        for com in range(model.k):
            mvn = torch.distributions.MultivariateNormal(model.centroid[com].double(),
                                                         covariance_matrix=model.covariance_mat[com].double())
            total += float((mvn.log_prob(x) * model.pi[rows, com].double()).sum().item())
        return abs(total) * (beta / model.k)

    def train(self, nodes, model, beta, chunksize=150, iter=1):
        """`nodes` may list a node more than once.  The reference adds the node's (frozen) gradient once per CHUNK the
        node occurs in -- `grad_input[node_index] += batch_grad_input` is a buffered fancy-index add, so repeats inside
        one chunk of `chunksize` collapse, repeats in different chunks accumulate (:64-73) -- and applies the sum once.
        The gradient is linear in pi, so the same result is the single-occurrence update with the node's
        responsibilities scaled by that multiplicity; every row is updated by exactly one warp/tile."""
        import torch
        dev = model.node_embedding.device
        rows = _rows_of(model, nodes)
        pi = model.pi.contiguous()
        if rows.size and np.unique(rows).size != rows.size:
            chunk = np.arange(rows.size, dtype=np.int64) // int(chunksize)
            pairs = np.unique(rows.astype(np.int64) * (int(chunk[-1]) + 1) + chunk)
            rows, mult = np.unique((pairs // (int(chunk[-1]) + 1)).astype(np.uint32), return_counts=True)
            if (mult != 1).any():
                pi = pi.clone()
                sel_m = torch.from_numpy(rows.astype(np.int64)).to(dev)
                pi[sel_m] *= torch.from_numpy(mult.astype(np.float32)).to(dev)[:, None]
        rows_d = torch.from_numpy(rows.view(np.int32)).to(dev)
        with torch.cuda.device(dev):
            inv_t = K.transpose_blocks(model.inv_covariance_mat.contiguous())
            # Rows whose pi has at most one non-zero entry (sklearn's predict_proba is one-hot in fp32 on separated
            # data) take the top-1 form: same arithmetic, but rows are grouped by community on the device and share
            # the inv_cov reads.  The others keep the dense form.
            single_all = (pi != 0).sum(1) <= 1
            single = single_all[rows_d.long()]
            mu = model.centroid.contiguous()
            if bool(single.any()):
                pi1 = pi if bool(single_all.all()) else pi * single_all[:, None].to(pi.dtype)
                comm, weight = K.pi_top1(pi1)
                sel = rows_d if bool(single.all()) else rows_d[single].contiguous()
                K.o3_batch_top1(model.node_embedding, sel, mu, inv_t, comm, weight, beta, self.lr, iters=iter)
            if not bool(single.all()):
                K.o3_batch(model.node_embedding, rows_d[~single].contiguous(), mu, inv_t, pi, beta, self.lr, iters=iter)


def _rows_of(model, nodes):
    """[model.vocab[x].index for x in nodes] (community_embeddings.py:63) without a Python loop for integer ids; an id
    that is not in the vocabulary raises KeyError like the reference's dict lookup."""
    if not isinstance(nodes, np.ndarray):
        nodes = list(nodes)  # may be a one-shot iterator
    try:
        a = np.asarray(nodes)
    except Exception:
        a = None
    if a is None or a.ndim != 1 or a.dtype.kind not in "iu" or a.size == 0 or not hasattr(model, "id_index"):
        return np.fromiter((model.vocab[x].index for x in nodes), dtype=np.uint32)
    ids, rows, _ = model.id_index()
    pos = np.searchsorted(ids, a)
    pos[pos >= ids.size] = 0
    bad = ids[pos] != a
    if bad.any():
        raise KeyError(a[bad][0].item())
    return rows[pos].astype(np.uint32)
