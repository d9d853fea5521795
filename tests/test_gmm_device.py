"""N1: the device GMM (ADSCModel/gmm_device.py) against sklearn's own EM, step by step from identical initial
responsibilities (sklearn's private `_initialize` / `_e_step` / `_m_step` driven by hand).  Runs on CPU tensors here
and on the GPU under -m gpu."""
import numpy as np
import pytest


def _blobs(n, d, k, seed):
    rs = np.random.RandomState(seed)
    centres = rs.normal(size=(k, d)) * 3
    lab = rs.randint(0, k, n)
    x = centres[lab] + rs.normal(size=(n, d)) * (0.5 + rs.rand(k)[lab, None])
    return x, lab


def _compare(device, dtype_np, tol):
    import torch
    from sklearn.mixture import GaussianMixture
    from comemb_b200.ADSCModel.gmm_device import DeviceGaussianMixture
    x, lab = _blobs(1500, 16, 4, 0)
    x = x.astype(dtype_np)
    rs = np.random.RandomState(1)
    resp = rs.rand(1500, 4).astype(dtype_np) ** 3
    resp /= resp.sum(1, keepdims=True)
    sk = GaussianMixture(n_components=4, covariance_type="full", reg_covar=1e-5, tol=0.0, max_iter=5)
    sk._initialize(x, resp)                       # sklearn's own initialisation from responsibilities
    lowers = []
    for _ in range(5):                            # fit_predict's loop body (sklearn/mixture/_base.py)
        log_prob_norm, log_resp = sk._e_step(x)
        sk._m_step(x, log_resp)
        lowers.append(float(log_prob_norm))
    gm = DeviceGaussianMixture(n_components=4, reg_covar=1e-5, tol=0.0, max_iter=5)
    gm.fit(torch.as_tensor(x, device=device), resp_init=torch.as_tensor(resp, device=device))
    assert gm.n_iter_ == 5
    assert abs(gm.lower_bound_ - lowers[-1]) <= tol * abs(lowers[-1])
    for got, want in ((gm.weights_, sk.weights_), (gm.means_, sk.means_), (gm.covariances_, sk.covariances_),
                      (gm.precisions_cholesky_, sk.precisions_cholesky_)):
        got = got.cpu().numpy()
        assert np.abs(got - want).max() <= tol * max(1.0, np.abs(want).max()), np.abs(got - want).max()
    p = gm.predict_proba(torch.as_tensor(x, device=device)).cpu().numpy()
    assert np.abs(p - sk.predict_proba(x)).max() <= 20 * tol


def test_device_gmm_em_steps_match_sklearn_cpu_float64():
    _compare("cpu", np.float64, 1e-9)


def test_device_gmm_em_steps_match_sklearn_cpu_float32():
    # float32 (what sklearn itself computes in for float32 embeddings): five EM iterations from a deliberately diffuse
    # start amplify fp32 summation-order differences between numpy's and torch's GEMMs to ~4e-4 relative
    _compare("cpu", np.float32, 2e-3)


def _sparse_m_step(device):
    """Well separated blobs: most responsibilities underflow to exactly 0, the M-step then gathers only the non-zero
    (point, component) pairs.  Must equal the dense batched form and sklearn's own M-step."""
    import torch
    from sklearn.mixture import GaussianMixture
    from comemb_b200.ADSCModel.gmm_device import DeviceGaussianMixture
    rs = np.random.RandomState(5)
    n, d, k = 6000, 12, 16
    centres = rs.normal(size=(k, d)) * 40
    lab = rs.randint(0, k, n)
    x = centres[lab] + rs.normal(size=(n, d))
    sk = GaussianMixture(n_components=k, covariance_type="full", reg_covar=1e-6, tol=0.0, max_iter=3)
    one_hot = np.eye(k)[lab]
    sk._initialize(x, one_hot)
    for _ in range(2):
        _, log_resp = sk._e_step(x)
        sk._m_step(x, log_resp)
    resp = np.exp(log_resp)
    assert (resp == 0).mean() > 0.9
    X, R = torch.as_tensor(x, device=device), torch.as_tensor(resp, device=device)
    out = {}
    for sparse in (True, False):
        gm = DeviceGaussianMixture(n_components=k, reg_covar=1e-6, sparse_m_step=sparse)
        if sparse:
            nk = R.sum(0) + 10 * torch.finfo(R.dtype).eps
            assert gm._covariances_sparse(X, R, nk, (R.T @ X) / nk[:, None]) is not None  # the sparse path is taken
        out[sparse] = [t.cpu().numpy() for t in gm._estimate_parameters(X, R)]
    for a, b in zip(out[True], out[False]):
        assert np.abs(a - b).max() <= 1e-10 * max(1.0, np.abs(b).max())
    assert np.abs(out[True][2] - sk.covariances_).max() <= 1e-9 * np.abs(sk.covariances_).max()
    # dense responsibilities: the sparse form declines
    gm = DeviceGaussianMixture(n_components=k)
    dense = torch.full_like(R, 1.0 / k)
    assert gm._covariances_sparse(X, dense, dense.sum(0), (dense.T @ X) / dense.sum(0)[:, None]) is None


def test_device_gmm_sparse_m_step_cpu():
    _sparse_m_step("cpu")


@pytest.mark.gpu
def test_device_gmm_sparse_m_step_gpu():
    _sparse_m_step("cuda")


def test_device_gmm_full_fit_recovers_blobs_cpu():
    import torch
    from sklearn.metrics import normalized_mutual_info_score as nmi
    from comemb_b200.ADSCModel.gmm_device import DeviceGaussianMixture
    x, lab = _blobs(2000, 16, 5, 3)
    gm = DeviceGaussianMixture(n_components=5, reg_covar=1e-6, n_init=3, random_state=0).fit(torch.as_tensor(x))
    assert gm.converged_ and nmi(lab, gm.predict(torch.as_tensor(x)).numpy()) > 0.9


@pytest.mark.gpu
def test_device_gmm_em_steps_match_sklearn_gpu():
    import torch
    torch.backends.cuda.matmul.allow_tf32 = False
    _compare("cuda", np.float64, 1e-9)
    _compare("cuda", np.float32, 2e-3)


@pytest.mark.gpu
def test_community2vec_fit_device_backend_feeds_o3(golden):
    """Community2Vec.fit(gmm_backend='device') -> centroid / inv_cov / pi on the GPU, consumed by the o3 kernel."""
    import torch
    from comemb_b200.ADSCModel.community_embeddings import Community2Vec

    class M(object):
        pass
    x, lab = _blobs(600, 128, 3, 5)
    m = M()
    m.k = 3
    m.node_embedding = torch.as_tensor(x.astype(np.float32), device="cuda")
    m.vocab = {i + 1: type("V", (), {"index": i})() for i in range(600)}
    learner = Community2Vec(m, lr=0.1, reg_covar=1e-3, gmm_backend="device")
    learner.fit(m)
    assert m.pi.shape == (600, 3) and m.inv_covariance_mat.shape == (3, 128, 128)
    assert torch.allclose(m.pi.sum(1), torch.ones(600, device="cuda"), atol=1e-4)
    before = m.node_embedding.clone()
    learner.train(list(range(1, 601)), m, beta=0.1, iter=1)
    assert torch.isfinite(m.node_embedding).all() and not torch.equal(before, m.node_embedding)
    # the step pulls nodes towards their component means
    mu = m.centroid[m.pi.argmax(1)]
    assert ((m.node_embedding - mu).norm(dim=1) < (before - mu).norm(dim=1) + 1e-6).float().mean() > 0.95


@pytest.mark.gpu
def test_estep_kernel_matches_library_form_and_sklearn_at_d128():
    """comemb_gmm_estep (tcgen05 3xTF32 tiles, squared norm reduced in the epilogue) at d = 128: the [N, K] matrix of
    squared Mahalanobis norms against a float64 evaluation of the same formula (<= 1e-5 relative to the row maximum: the
    expression x.P - mu.P cancels, any fp32 evaluation carries that), ragged N (not a multiple of the 64-row tile),
    K = 1 and K > number of SMs' worth of jobs; then five EM iterations from identical responsibilities against sklearn's
    own float32 EM like the library path."""
    import torch
    from sklearn.mixture import GaussianMixture
    from comemb_b200 import _lib
    from comemb_b200.ADSCModel.gmm_device import DeviceGaussianMixture
    rs = np.random.RandomState(4)
    for n, K in ((1, 1), (63, 2), (1000, 5), (4099, 7)):
        d = 128
        x = rs.normal(size=(n, d)).astype(np.float32)
        mu = rs.normal(size=(K, d)).astype(np.float32) * 0.5
        P = (np.triu(rs.normal(size=(K, d, d))) * 0.05 + np.eye(d)).astype(np.float32)  # upper triangular like sklearn's
        bias = np.einsum("kb,kbj->kj", mu.astype(np.float64), P.astype(np.float64)).astype(np.float32)
        want = ((np.einsum("nb,kbj->nkj", x.astype(np.float64), P.astype(np.float64)) - bias[None].astype(np.float64)) ** 2).sum(2)
        dx, dP, db = (torch.as_tensor(a, device="cuda") for a in (x, P, bias))
        out = torch.full((n, K), -1.0, device="cuda")
        _lib.check(_lib.load().comemb_gmm_estep(dx.data_ptr(), n, d, dP.data_ptr(), db.data_ptr(), K, out.data_ptr(), None))
        got = out.cpu().numpy()
        assert np.abs(got - want).max() <= 1e-5 * want.max(), (n, K, np.abs(got - want).max(), want.max())
    x, lab = _blobs(3000, 128, 4, 9)
    x = x.astype(np.float32)
    resp = rs.rand(3000, 4).astype(np.float32) ** 3
    resp /= resp.sum(1, keepdims=True)
    sk = GaussianMixture(n_components=4, covariance_type="full", reg_covar=1e-4, tol=0.0, max_iter=5)
    sk._initialize(x, resp)
    for _ in range(5):
        log_prob_norm, log_resp = sk._e_step(x)
        sk._m_step(x, log_resp)
    res = {}
    for kern in (True, False):
        gm = DeviceGaussianMixture(n_components=4, reg_covar=1e-4, tol=0.0, max_iter=5, estep_kernel=kern)
        gm.fit(torch.as_tensor(x, device="cuda"), resp_init=torch.as_tensor(resp, device="cuda"))
        res[kern] = gm
        assert abs(gm.lower_bound_ - float(log_prob_norm)) <= 2e-3 * abs(float(log_prob_norm))
        for got, want in ((gm.means_, sk.means_), (gm.covariances_, sk.covariances_)):
            assert np.abs(got.cpu().numpy() - want).max() <= 2e-3 * max(1.0, np.abs(want).max())
    assert abs(res[True].lower_bound_ - res[False].lower_bound_) <= 1e-4 * abs(res[False].lower_bound_)


@pytest.mark.gpu
def test_mstep_kernel_matches_float64_scatter_and_library_form_at_d128():
    """comemb_gmm_mstep (csrc/gmm_mstep.cu: tcgen05 3xTF32, four components per CTA in TMEM, centred / weighted tiles
    built in shared memory) against a float64 evaluation of  S_k = sum_i r_ik (x_i - mu_k)(x_i - mu_k)^T : <= 1e-5 of the
    largest entry (measured <= 2.9e-6: fp32 accumulation over up to 20 000 points; the 3xTF32 split itself drops 2^-22),
    for ragged n (not a multiple of the 32-point tile), K = 1, K not a
    multiple of 4, more component groups than fit one wave, dense and sparse responsibilities (all-zero tiles are
    skipped; a component without any weight gives exactly 0); then five EM iterations from identical responsibilities
    with and without the kernel: means and covariances within 5e-4 of the largest entry (measured 6e-5: both paths are
    fp32 and five EM iterations amplify their rounding differences), lower bound within 1e-5."""
    import torch
    from comemb_b200 import _lib
    from comemb_b200.ADSCModel.gmm_device import DeviceGaussianMixture
    rs = np.random.RandomState(11)
    d = 128
    for n, K, sparse in ((1, 1, False), (33, 2, False), (1000, 5, False), (4099, 7, True), (20000, 50, True), (257, 9, False)):
        x = (rs.normal(size=(n, d)) + rs.normal(size=(1, d))).astype(np.float32)
        mu = (x.mean(0)[None] + rs.normal(size=(K, d)) * 0.3).astype(np.float32)
        resp = rs.rand(n, K).astype(np.float32) ** 4
        if sparse:
            lab = rs.randint(0, K, size=n)
            keep = np.zeros((n, K), bool)
            keep[np.arange(n), lab] = True
            keep |= rs.rand(n, K) < 0.02
            resp = np.where(keep, resp + 1e-3, 0).astype(np.float32)
            if K > 3:
                resp[:, 3] = 0  # a component without any weight
        diff = x.astype(np.float64)[None] - mu.astype(np.float64)[:, None]          # [K, n, d]
        want = np.einsum("kna,knb->kab", diff * resp.T.astype(np.float64)[:, :, None], diff)
        dx, dr, dm = (torch.as_tensor(a, device="cuda") for a in (x, resp, mu))
        out = torch.full((K, d, d), -7.0, device="cuda")
        _lib.check(_lib.load().comemb_gmm_mstep(dx.data_ptr(), n, d, dr.data_ptr(), dm.data_ptr(), K, out.data_ptr(), None))
        got = out.cpu().numpy()
        assert np.abs(got - want).max() <= 1e-5 * max(np.abs(want).max(), 1e-30), (n, K, np.abs(got - want).max(), np.abs(want).max())
        if sparse and K > 3:
            assert not got[3].any()
    x, lab = _blobs(3000, 128, 4, 9)
    x = x.astype(np.float32)
    resp = rs.rand(3000, 4).astype(np.float32) ** 3
    resp /= resp.sum(1, keepdims=True)
    res = {}
    for kern in (True, False):
        gm = DeviceGaussianMixture(n_components=4, reg_covar=1e-4, tol=0.0, max_iter=5, mstep_kernel=kern, sparse_m_step=False)
        gm.fit(torch.as_tensor(x, device="cuda"), resp_init=torch.as_tensor(resp, device="cuda"))
        res[kern] = gm
    assert abs(res[True].lower_bound_ - res[False].lower_bound_) <= 1e-5 * abs(res[False].lower_bound_)
    for a, b in ((res[True].means_, res[False].means_), (res[True].covariances_, res[False].covariances_)):
        assert float((a - b).abs().max()) <= 5e-4 * max(1.0, float(b.abs().max()))
