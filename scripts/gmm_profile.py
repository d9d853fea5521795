"""Where the device GMM fit spends its time (N=100K, d=128, K=50): per-phase CUDA-event timings of the current
per-component formulation and of batched candidates."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from comemb_b200.ADSCModel.gmm_device import DeviceGaussianMixture  # noqa: E402


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


def main():
    N, d, K = 100000, 128, 50
    g = torch.Generator(device="cuda").manual_seed(0)
    centres = torch.randn((K, d), device="cuda", generator=g) * 0.5
    lab = torch.randint(0, K, (N,), device="cuda", generator=g)
    X = centres[lab] + 0.1 * torch.randn((N, d), device="cuda", generator=g)
    gm = DeviceGaussianMixture(n_components=K, reg_covar=1e-4, n_init=1, random_state=0)
    gen = torch.Generator(device="cuda").manual_seed(0)
    t, resp = timed(lambda: gm._kmeans_resp(X, gen), 1)
    print("kmeans init            %8.2f ms" % t)
    t, (nk, means, covs) = timed(lambda: gm._estimate_parameters(X, resp))
    print("M-step (params)        %8.2f ms" % t)
    gm.means_, gm.covariances_, gm.weights_ = means, covs, nk / N
    t, pc = timed(lambda: gm._precision_cholesky(covs))
    print("precision cholesky     %8.2f ms" % t)
    gm.precisions_cholesky_ = pc
    t, (norm, log_resp) = timed(lambda: gm._log_prob_resp(X))
    print("E-step                 %8.2f ms" % t)

    # ---- candidates ----
    def e_batched():
        Pcat = pc.permute(1, 0, 2).reshape(d, K * d)
        b = torch.bmm(means[:, None, :], pc).reshape(K * d)
        y = X @ Pcat
        y.sub_(b).square_()
        return y.view(N, K, d).sum(2)

    t, lp = timed(e_batched)
    ref = torch.stack([(((X @ pc[k]) - (means[k] @ pc[k])) ** 2).sum(1) for k in range(K)], 1)
    print("E-step batched GEMM    %8.2f ms  (max rel diff %.2e)" % (t, float(((lp - ref).abs() / ref.abs().clamp(min=1e-6)).max())))

    def m_batched(kc=10):
        covs2 = torch.empty((K, d, d), device="cuda")
        for k0 in range(0, K, kc):
            diff = X[None, :, :] - means[k0:k0 + kc, None, :]
            wd = diff * resp.T[k0:k0 + kc, :, None]
            covs2[k0:k0 + kc] = torch.bmm(wd.transpose(1, 2), diff) / nk[k0:k0 + kc, None, None]
        covs2.diagonal(dim1=1, dim2=2).add_(1e-4)
        return covs2

    for kc in (5, 10, 25, 50):
        t, c2 = timed(lambda: m_batched(kc))
        print("M-step covs bmm kc=%-3d %8.2f ms  (max abs diff %.2e)" % (kc, t, float((c2 - covs).abs().max())))

    def m_uncentred():
        # sum_n r_nk x x^T via one GEMM per chunk of components on sqrt-weighted copies is the same cost; the
        # algebraic shortcut: S_k = (X^T diag(r_k) X)/n_k - mu mu^T
        covs3 = torch.empty((K, d, d), device="cuda")
        for k in range(K):
            covs3[k] = (X.T * resp[:, k]) @ X / nk[k] - torch.outer(means[k], means[k])
        covs3.diagonal(dim1=1, dim2=2).add_(1e-4)
        return covs3

    t, c3 = timed(m_uncentred)
    print("M-step covs uncentred  %8.2f ms  (max abs diff %.2e)" % (t, float((c3 - covs).abs().max())))
    for tf32 in (False, True):
        torch.backends.cuda.matmul.allow_tf32 = tf32
        t, _ = timed(lambda: gm._estimate_parameters(X, resp))
        t2, _ = timed(lambda: gm._log_prob_resp(X))
        t3, _ = timed(e_batched)
        t4, _ = timed(lambda: m_batched(10))
        print("allow_tf32=%s: M %.2f ms, E %.2f ms, E batched %.2f ms, M bmm10 %.2f ms" % (tf32, t, t2, t3, t4))
    torch.backends.cuda.matmul.allow_tf32 = False
    t0 = time.perf_counter()
    DeviceGaussianMixture(n_components=K, reg_covar=1e-4, n_init=10, random_state=0).fit(X)
    torch.cuda.synchronize()
    print("full fit n_init=10: %.2f s" % (time.perf_counter() - t0))


if __name__ == "__main__":
    main()
