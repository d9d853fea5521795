"""Full-covariance Gaussian mixture EM on the device (the producer of o3's inputs).

The reference fits `sklearn.mixture.GaussianMixture(k, 'full', n_init=10, reg_covar)` on the host
(ADSCModel/community_embeddings.py:16-37); at 100K x 128 points and K=50 that is minutes per fit and dominates an
outer iteration once o1/o2/o3 run at GPU speed.  This class restates sklearn's EM step by step
(sklearn/mixture/_gaussian_mixture.py: `_estimate_gaussian_parameters`, `_compute_precision_cholesky`,
`_estimate_log_gaussian_prob`, `_e_step`/`_m_step`, lower bound, `tol` on its change) on torch tensors.  The
E-step -- the dominant 2*N*K*d^2 contraction -- is the hand-written tcgen05 kernel comemb_gmm_estep at d == 128 in fp32
(csrc/o3_gemm.cu: P_k resident in shared memory as the 3xTF32 A operand, points streamed as 64-row tiles, squared norm
reduced in the epilogue, only [N, K] written); the fallback (other sizes, float64, CPU) is one library GEMM [N,d] x [d,K*d]
per row block.  The M-step's covariances are the hand-written tcgen05 kernel comemb_gmm_mstep
under the same condition (csrc/gmm_mstep.cu: four components per CTA accumulate in TMEM, centred / weighted point tiles built
in shared memory, no [K, N, d] temporary), otherwise batched library GEMMs ([kc,d,N] x [kc,N,d]; cuBLAS); the precision
factors are batched Cholesky / triangular solves (cuSOLVER / cuBLAS).

Parity: given the same initial responsibilities the iterations follow sklearn's to fp32 round-off
(tests/test_gmm_device.py, CPU and GPU).  The initialisation differs (own k-means++ / Lloyd instead of sklearn's KMeans
stream), so a whole `fit` is compared through its lower bound and the NMI of the hard assignments.
"""
import math

import numpy as np


class DeviceGaussianMixture(object):
    def __init__(self, n_components=1, reg_covar=1e-6, tol=1e-3, max_iter=100, n_init=1, random_state=None,
                 dtype=None, kmeans_iter=20, workspace_bytes=6 << 30, tf32=False, sparse_m_step=True, estep_kernel=True,
                 mstep_kernel=True):
        self.n_components = int(n_components)
        self.reg_covar = float(reg_covar)
        self.tol = float(tol)
        self.max_iter = int(max_iter)
        self.n_init = int(n_init)
        self.random_state = random_state
        self.dtype = dtype
        self.kmeans_iter = kmeans_iter
        self.workspace_bytes = int(workspace_bytes)  # bound on the temporaries of the batched E / M steps
        self.sparse_m_step = bool(sparse_m_step)
        self.estep_kernel = bool(estep_kernel)  # fp32, d == 128 on CUDA: the tcgen05 E-step kernel instead of the library GEMM
        self.mstep_kernel = bool(mstep_kernel)  # same condition: the tcgen05 covariance kernel instead of the batched GEMMs
        self.tf32 = bool(tf32)  # let cuBLAS use TF32 tensor cores for the fp32 GEMMs (about 1.7x per EM iteration;
        #                         the sklearn comparison of tests/test_gmm_device.py holds for tf32=False)
        self.converged_ = False

    # ---- sklearn: _estimate_gaussian_parameters + _estimate_gaussian_covariances_full ----------------------------------
    def _estimate_parameters(self, X, resp):
        """covariances[k] = (resp[:, k] * diff_k.T) @ diff_k / nk[k] + reg_covar * I with diff_k = X - means[k].
        Dense form: as many components at a time as fit the workspace, one batched GEMM [kc, d, N] x [kc, N, d] instead
        of kc launches of a 128 x 128-output GEMM.  Sparse form: responsibilities that are exactly 0 (exp() underflows
        for every far component once the clusters separate) contribute exactly 0, so when few are non-zero only those
        (point, component) pairs are gathered -- padded per component -- into one batched GEMM [K, d, M] x [K, M, d]."""
        import torch
        nk = resp.sum(0) + 10 * torch.finfo(resp.dtype).eps
        means = (resp.T @ X) / nk[:, None]
        K, d = means.shape
        n = X.shape[0]
        covs = None
        if self.mstep_kernel and X.is_cuda and X.dtype == torch.float32 and d == 128 and n > 0:
            # hand-written covariance kernel (csrc/gmm_mstep.cu, comemb_gmm_mstep): the centred / weighted tiles are
            # built in shared memory, four components per CTA accumulate in TMEM -- no [K, n, d] temporary, no host sync
            from .. import _lib
            covs = torch.empty((K, d, d), dtype=X.dtype, device=X.device)
            xc, rc, mc = X.contiguous(), resp.contiguous(), means.contiguous()
            with torch.cuda.device(X.device):
                _lib.check(_lib.load().comemb_gmm_mstep(_lib.ptr(xc), n, d, _lib.ptr(rc), _lib.ptr(mc), K, _lib.ptr(covs),
                                                        _lib.stream_ptr()))
            covs /= nk[:, None, None]
        if covs is None and self.sparse_m_step and n * K >= (1 << 16):
            covs = self._covariances_sparse(X, resp, nk, means)
        if covs is None:
            covs = torch.empty((K, d, d), dtype=X.dtype, device=X.device)
            kc = max(1, min(K, int(self.workspace_bytes // (2 * n * d * X.element_size()))))
            respT = resp.T.contiguous()
            for k0 in range(0, K, kc):
                diff = X[None, :, :] - means[k0:k0 + kc, None, :]
                wd = diff * respT[k0:k0 + kc, :, None]
                covs[k0:k0 + kc] = torch.bmm(wd.transpose(1, 2), diff) / nk[k0:k0 + kc, None, None]
                del diff, wd
        covs.diagonal(dim1=1, dim2=2).add_(self.reg_covar)
        return nk, means, covs

    def _covariances_sparse(self, X, resp, nk, means):
        """None if the responsibilities are not sparse enough to pay off (more than a quarter non-zero, or one component
        holding more than half of the points)."""
        import torch
        n, K = resp.shape
        nz = resp != 0
        per_k = nz.sum(0)
        m = int(per_k.max())  # one host sync per M-step
        if m == 0 or m * K > n * K // 4 or m > n // 2:
            return None
        kk, nn = nz.T.nonzero(as_tuple=True)  # sorted by component, then point
        start = torch.cumsum(per_k, 0) - per_k
        slot = torch.arange(kk.numel(), device=X.device) - start[kk]
        rows = torch.zeros((K, m), dtype=torch.int64, device=X.device)
        w = torch.zeros((K, m), dtype=X.dtype, device=X.device)
        rows[kk, slot] = nn
        w[kk, slot] = resp[nn, kk]
        diff = X[rows] - means[:, None, :]  # [K, m, d]; padding slots have weight 0
        return torch.bmm((diff * w[:, :, None]).transpose(1, 2), diff) / nk[:, None, None]

    # ---- sklearn: _compute_precision_cholesky ('full') ------------------------------------------------------------------
    @staticmethod
    def _precision_cholesky(covs):
        import torch
        L, info = torch.linalg.cholesky_ex(covs)
        if int(info.max()) != 0:
            raise ValueError("Fitting the mixture model failed because some components have ill-defined empirical "
                             "covariance. Try to decrease the number of components, or increase reg_covar.")
        eye = torch.eye(covs.shape[-1], dtype=covs.dtype, device=covs.device).expand_as(covs)
        return torch.linalg.solve_triangular(L, eye, upper=False).transpose(-1, -2).contiguous()

    # ---- sklearn: _estimate_log_gaussian_prob + weights, logsumexp (_estimate_log_prob_resp) -------------------------
    def _log_prob_resp(self, X):
        """log_prob[:, k] = sum((X @ P_k - mu_k @ P_k)**2, axis=1): all components in one GEMM X @ [P_0 | ... | P_K-1]
        per block of rows (the [rows, K*d] product is the workspace)."""
        import torch
        K, d = self.means_.shape
        n = X.shape[0]
        P = self.precisions_cholesky_
        log_det = torch.log(torch.diagonal(P, dim1=-2, dim2=-1)).sum(1)
        Pcat = P.permute(1, 0, 2).reshape(d, K * d)
        b_neg = -torch.bmm(self.means_[:, None, :], P).reshape(1, K * d)
        log_prob = torch.empty((n, K), dtype=X.dtype, device=X.device)
        if self.estep_kernel and X.is_cuda and X.dtype == torch.float32 and d == 128 and n > 0:
            # hand-written E-step (csrc/o3_gemm.cu, comemb_gmm_estep): tcgen05 3xTF32 tiles, P_k resident in shared
            # memory, squared norm reduced in the epilogue -- no [rows, K*d] product is ever written
            from .. import _lib
            xc, pc, bias = X.contiguous(), P.contiguous(), (-b_neg).reshape(K, d).contiguous()
            with torch.cuda.device(X.device):
                _lib.check(_lib.load().comemb_gmm_estep(_lib.ptr(xc), n, d, _lib.ptr(pc), _lib.ptr(bias), K,
                                                        _lib.ptr(log_prob), _lib.stream_ptr()))
        else:
            rows = max(1, int(self.workspace_bytes // (K * d * X.element_size())))
            for n0 in range(0, n, rows):
                y = torch.addmm(b_neg, X[n0:n0 + rows], Pcat)  # X @ Pcat - b, the bias added in the GEMM epilogue
                log_prob[n0:n0 + rows] = torch.linalg.vector_norm(y.view(-1, K, d), dim=2).square_()  # one pass over y
                del y
        weighted = -0.5 * (d * math.log(2 * math.pi) + log_prob) + log_det + torch.log(self.weights_)
        norm = torch.logsumexp(weighted, dim=1)
        return norm, weighted - norm[:, None]

    def _m_step(self, X, log_resp):
        nk, self.means_, self.covariances_ = self._estimate_parameters(X, torch_exp(log_resp))
        self.weights_ = nk / nk.sum()
        self.precisions_cholesky_ = self._precision_cholesky(self.covariances_)

    # ---- initial responsibilities: k-means++ seeding + Lloyd iterations (sklearn uses its own KMeans here) --------------
    def _kmeans_resp(self, X, gen):
        """No host synchronisation inside the seeding loop; Lloyd's convergence is checked every 5 iterations."""
        import torch
        n, K = X.shape[0], self.n_components
        first = torch.randint(n, (1,), generator=gen, device=X.device)
        centres = torch.empty((K, X.shape[1]), dtype=X.dtype, device=X.device)
        centres[0] = X[first[0]]
        d2 = ((X - centres[0]) ** 2).sum(1)
        for k in range(1, K):
            nxt = torch.multinomial(d2 / d2.sum(), 1, generator=gen)
            centres[k] = X[nxt[0]]
            d2 = torch.minimum(d2, ((X - centres[k]) ** 2).sum(1))
        xx = (X * X).sum(1, keepdim=True)
        one = None
        S = 64  # the [K, N] x [N, d] centre sums are split over the N dimension into S independent GEMMs (one
        #         [K, d]-output GEMM with a 100K-long reduction runs on a handful of SMs)
        n_main = (n // S) * S
        for it in range(self.kmeans_iter):
            dist = torch.addmm(xx, X, centres.T, alpha=-2.0)  # ||x||^2 - 2 x.c  (+ ||c||^2 below)
            dist.add_((centres * centres).sum(1)[None, :])
            lab = dist.argmin(1)
            one = torch.zeros((n, K), dtype=X.dtype, device=X.device)
            one[torch.arange(n, device=X.device), lab] = 1
            cnt = one.sum(0)
            if n_main >= S * 64:
                sums = torch.bmm(one[:n_main].view(S, n_main // S, K).transpose(1, 2),
                                 X[:n_main].view(S, n_main // S, -1)).sum(0)
                if n_main < n:
                    sums = sums + one[n_main:].T @ X[n_main:]
            else:
                sums = one.T @ X
            new = sums / cnt.clamp(min=1)[:, None]
            new = torch.where((cnt == 0)[:, None], centres, new)
            done = (it % 5 == 4) and bool(torch.allclose(new, centres))
            centres = new
            if done:
                break
        return one

    def _run_em(self, X, resp):
        nk, self.means_, self.covariances_ = self._estimate_parameters(X, resp)
        self.weights_ = nk / X.shape[0]  # sklearn _initialize: weights /= n_samples
        self.precisions_cholesky_ = self._precision_cholesky(self.covariances_)
        lower = -float("inf")
        converged = False
        n_iter = 0
        for n_iter in range(1, self.max_iter + 1):
            prev = lower
            norm, log_resp = self._log_prob_resp(X)
            self._m_step(X, log_resp)
            lower = float(norm.mean())
            if abs(lower - prev) < self.tol:
                converged = True
                break
        return lower, converged, n_iter

    def fit(self, X, resp_init=None):
        """X: [n, d] torch tensor (any device) or numpy array.  resp_init: optional [n, K] initial responsibilities
        (then n_init is ignored): used by the parity tests."""
        import torch
        if not isinstance(X, torch.Tensor):
            X = torch.as_tensor(np.asarray(X))
        if self.dtype is not None:
            X = X.to(self.dtype)
        prev_tf32 = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = self.tf32
        try:
            return self._fit(X, resp_init)
        finally:
            torch.backends.cuda.matmul.allow_tf32 = prev_tf32

    def _fit(self, X, resp_init):
        import torch
        gen = torch.Generator(device=X.device)
        gen.manual_seed(0 if self.random_state is None else int(self.random_state))
        best = None
        inits = [resp_init.to(X)] if resp_init is not None else [None] * self.n_init
        for init in inits:
            resp = init if init is not None else self._kmeans_resp(X, gen)
            lower, conv, n_iter = self._run_em(X, resp)
            if best is None or lower > best[0]:
                best = (lower, conv, n_iter, self.weights_, self.means_, self.covariances_, self.precisions_cholesky_)
        (self.lower_bound_, self.converged_, self.n_iter_, self.weights_, self.means_, self.covariances_,
         self.precisions_cholesky_) = best
        return self

    def predict_proba(self, X):
        import torch
        if not isinstance(X, torch.Tensor):
            X = torch.as_tensor(np.asarray(X)).to(self.means_)
        _, log_resp = self._log_prob_resp(X.to(self.means_))
        return torch.exp(log_resp)

    def predict(self, X):
        return self.predict_proba(X).argmax(1)


def torch_exp(x):
    import torch
    return torch.exp(x)
