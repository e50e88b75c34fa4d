"""Device-level engine: thin Python over the C ABI (include/gp2d.h).

torch is used only for device memory, streams and host<->device copies; every number is
produced by the hand-written sm_100a kernels in csrc/.  There is no CPU path: calls raise
if CUDA or libgp2d.so is unavailable.
"""
from __future__ import annotations

import numpy as np
import torch

from ._lib import lib, check, Gp2dError, GP2D_OPT_PREDICT_I8

__all__ = ["set_predict_i8", "LinAlgError", "as_dev", "kernel_K", "kernel_Kdiag", "kernel_grad_sums", "potrf",
           "spd_inverse", "matmul", "HelmholtzGP", "HelmholtzBatch", "krig_snapshots", "fit_predict_host", "rbf_K",
           "rbf_grad_sums", "ScalarGP", "st_K", "st_grad_sums", "SpaceTimeGP"]


def set_predict_i8(mode=0):
    """Which kernel the predictive pass of the Helmholtz families uses (GP2D_OPT_PREDICT_I8, include/gp2d.h; per
    host thread; set it BEFORE the fit): 0 = automatic (int8-sliced tcgen05 kernel with 6 or 7 slices chosen from the
    conditioning bound, fp64 tensor-pipe kernel beyond it), 1 = fp64 kernel only, 6 / 7 = that slice count."""
    check(lib.gp2d_set_option(GP2D_OPT_PREDICT_I8, float(mode)), "gp2d_set_option")


class LinAlgError(np.linalg.LinAlgError):
    """Raised when the covariance is not positive definite (the reference surfaces this as
    numpy/GPy LinAlgError)."""


def _device(device=None) -> torch.device:
    if not torch.cuda.is_available():
        raise Gp2dError("gp2d needs a CUDA device: the hot path has no CPU fallback")
    if device is None:
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device(device)


def as_dev(a, device=None, shape=None) -> torch.Tensor:
    """float64, contiguous, on the GPU (numpy arrays go through pinned memory)."""
    dev = _device(device)
    if isinstance(a, torch.Tensor):
        t = a.to(device=dev, dtype=torch.float64, non_blocking=True)
    else:
        h = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64))
        t = h.pin_memory().to(dev, non_blocking=True) if h.numel() > 65536 else h.to(dev)
    t = t.contiguous()
    if shape is not None:
        t = t.reshape(shape)
    return t


def _ptr(t):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _points(X, device=None) -> torch.Tensor:
    t = as_dev(X, device)
    if t.dim() != 2 or t.shape[1] != 2:
        raise ValueError("points must be [N,2] (the kernels assume input_dim == 2, myKernel.py:15)")
    return t


# ------------------------------------------------------------------------------------------
def kernel_K(X, X2=None, l_df=1.0, l_cf=1.0, ratio=1.0, diag_add=0.0, out=None) -> torch.Tensor:
    """[2N,2M] Helmholtz covariance in the reference block layout (device tensor)."""
    Xd = _points(X)
    X2d = None if X2 is None else _points(X2, Xd.device)
    N, M = Xd.shape[0], (Xd.shape[0] if X2d is None else X2d.shape[0])
    if out is None:
        out = torch.empty((2 * N, 2 * M), dtype=torch.float64, device=Xd.device)
    if N and M:
        with torch.cuda.device(Xd.device):
            check(lib.gp2d_kernel_build(_ptr(Xd), N, _ptr(X2d), M, l_df, l_cf, ratio, diag_add,
                                        _ptr(out), out.stride(0), _stream()), "gp2d_kernel_build")
    return out


def kernel_Kdiag(M, l_df, l_cf, ratio, device=None) -> torch.Tensor:
    out = torch.empty(2 * int(M), dtype=torch.float64, device=_device(device))
    with torch.cuda.device(out.device):
        check(lib.gp2d_kdiag(int(M), l_df, l_cf, ratio, _ptr(out), _stream()), "gp2d_kdiag")
    return out


def kernel_grad_sums(dL_dK, X, X2, l_df, l_cf, ratio, reference_compat=False) -> torch.Tensor:
    """sum(dK/dtheta * dL_dK) for (l_df, l_cf, ratio) -> device tensor[3]."""
    Xd = _points(X)
    X2d = None if X2 is None else _points(X2, Xd.device)
    N, M = Xd.shape[0], (Xd.shape[0] if X2d is None else X2d.shape[0])
    W = as_dev(dL_dK, Xd.device)
    if tuple(W.shape) != (2 * N, 2 * M):
        raise ValueError("dL_dK must be [2N,2M]")
    nb = lib.gp2d_kernel_grad_workspace_bytes(N, M)
    ws = torch.empty(nb, dtype=torch.uint8, device=Xd.device)
    out = torch.empty(3, dtype=torch.float64, device=Xd.device)
    with torch.cuda.device(Xd.device):
        check(lib.gp2d_kernel_grad(_ptr(Xd), N, _ptr(X2d), M, l_df, l_cf, ratio, int(bool(reference_compat)),
                                   _ptr(W), W.stride(0), _ptr(ws), nb, _ptr(out), _stream()),
              "gp2d_kernel_grad")
    return out


def potrf(A, overwrite=False):
    """Lower Cholesky of a row-major SPD matrix on the device.  Returns (L, info)."""
    Ad = as_dev(A)
    if not (overwrite and isinstance(A, torch.Tensor) and Ad.data_ptr() == A.data_ptr()):
        Ad = Ad.clone()
    n = Ad.shape[0]
    nb = lib.gp2d_potrf_workspace_bytes(n)
    ws = torch.empty(nb, dtype=torch.uint8, device=Ad.device)
    info = torch.zeros(1, dtype=torch.int32, device=Ad.device)
    with torch.cuda.device(Ad.device):
        check(lib.gp2d_potrf(_ptr(Ad), n, Ad.stride(0), _ptr(ws), nb, _ptr(info), _stream()), "gp2d_potrf")
    return torch.tril(Ad), int(info.item())


def spd_inverse(A):
    """Inverse of a symmetric positive definite matrix (device tensor); raises LinAlgError."""
    Ad = as_dev(A).clone()
    n = Ad.shape[0]
    nb = lib.gp2d_spd_inverse_workspace_bytes(n)
    ws = torch.empty(nb, dtype=torch.uint8, device=Ad.device)
    info = torch.zeros(1, dtype=torch.int32, device=Ad.device)
    with torch.cuda.device(Ad.device):
        check(lib.gp2d_spd_inverse(_ptr(Ad), n, Ad.stride(0), _ptr(ws), nb, _ptr(info), _stream()),
              "gp2d_spd_inverse")
    i = int(info.item())
    if i > 0:
        raise LinAlgError("matrix not positive definite (pivot %d)" % i)
    return Ad


def _pad2(t, r, c):
    if t.shape[0] == r and t.shape[1] == c and t.is_contiguous():
        return t
    out = torch.zeros((r, c), dtype=torch.float64, device=t.device)
    out[:t.shape[0], :t.shape[1]] = t
    return out


def matmul(A, B) -> torch.Tensor:
    """A @ B in fp64 on the DMMA GEMM kernel (operands zero-padded to the tile sizes)."""
    Ad = as_dev(A)
    Bd = as_dev(B, Ad.device)
    vec = Bd.dim() == 1
    if vec:
        Bd = Bd.reshape(-1, 1)
    M, K = Ad.shape
    K2, N = Bd.shape
    if K != K2:
        raise ValueError("shape mismatch")
    Mp, Np, Kp = -(-M // 128) * 128, -(-N // 128) * 128, -(-K // 16) * 16
    Ap, Bp = _pad2(Ad, Mp, Kp), _pad2(Bd, Kp, Np)
    Cp = torch.empty((Mp, Np), dtype=torch.float64, device=Ad.device)
    with torch.cuda.device(Ad.device):
        check(lib.gp2d_dgemm(0, 0, Mp, Np, Kp, 1.0, _ptr(Ap), Ap.stride(0), _ptr(Bp), Bp.stride(0), 0.0,
                             _ptr(Cp), Cp.stride(0), _stream()), "gp2d_dgemm")
    out = Cp[:M, :N]
    return out.reshape(-1) if vec else out


# ------------------------------------------------------------------------------------------
# Iterated solves for ill-conditioned covariances (DESIGN.md §7 "Conditioning")
# ------------------------------------------------------------------------------------------
REFINE_CHUNK = 1 << 25      # doubles in one K* chunk (256 MB)
ROBUST_COND = 1e7           # same threshold as csrc/capi.cu refine_steps_for


def refined_predict(K_fn, X, y, Xs, ncomp, kdiag_fn, diag_add, var_add, steps=3, chunk_elems=REFINE_CHUNK, cache=None):
    """Prediction whose accuracy does not degrade with cond(K): the fused kernel applies the explicit
    inverse factor (error ~ cond(K) eps), here the refined inverse P is only a preconditioner and
    both solves are iterated against fp64 residuals,
        alpha += P (y - K alpha),   W += P (K*^T - K W),   var = k** - colsum(K*^T o W),
    ending at the accuracy of a backward-stable Cholesky solve (three steps at cond 1e13).  Built
    from this library's kernel builds, gp2d_spd_inverse and gp2d_dgemm; K* and W are materialised
    chunk by chunk over the grid points, ~8x the flops of the fused pass.
    K_fn(A, B, diag) -> covariance block (reference layout); ncomp = 2 for the stacked vector
    kernels; kdiag_fn(m) -> prior variances of an m-point chunk in the same stacking.
    ``cache``: a dict owned by the caller; the O(n^3) part (K, its inverse, the iterated alpha) is stored
    there under ``key`` and reused by later calls with the same key (krig.predict calls once per time
    slice): pass {"key": (hyper-parameters, diag_add, steps)} and invalidate by replacing the dict."""
    dev = X.device
    N = int(X.shape[0])
    n = ncomp * N
    key = None if cache is None else cache.get("key")
    if cache is not None and cache.get("have") == key and key is not None:
        Kh, P, a = cache["Kh"], cache["P"], cache["a"]
    else:
        Kh = K_fn(X, None, diag_add)
        P = spd_inverse(Kh)
        a = matmul(P, y)
        for _ in range(int(steps)):
            a = a + matmul(P, y - matmul(Kh, a))
        if cache is not None and key is not None:
            cache.update(Kh=Kh, P=P, a=a, have=key)
    M = int(Xs.shape[0])
    mean = torch.empty(ncomp * M, dtype=torch.float64, device=dev)
    var = torch.empty(ncomp * M, dtype=torch.float64, device=dev)
    step = max(1, int(chunk_elems) // max(n * ncomp, 1))
    for lo in range(0, M, step):
        hi = min(M, lo + step)
        Ks = K_fn(Xs[lo:hi], X, 0.0)                       # [ncomp (hi-lo), n]
        KsT = Ks.t().contiguous()
        W = matmul(P, KsT)
        for _ in range(int(steps)):
            dW = matmul(P, KsT - matmul(Kh, W))
            W = W + dW
            # converged to rounding: a further step would not change the variance (one host read per step)
            if float(dW.abs().max()) <= 1e-13 * float(W.abs().max()):
                break
        mc = matmul(Ks, a)
        vc = torch.clamp(kdiag_fn(hi - lo) - (KsT * W).sum(0), min=0.0) + var_add
        for c in range(ncomp):
            mean[c * M + lo:c * M + hi] = mc[c * (hi - lo):(c + 1) * (hi - lo)]
            var[c * M + lo:c * M + hi] = vc[c * (hi - lo):(c + 1) * (hi - lo)]
    return mean, var


def _refine_cache(gp, key):
    """Cache slot of the iterated-solve prediction on a GP object: keyed on the hyper-parameters, the
    EFFECTIVE diagonal term (noise and the jitter the fit settled on, models._jitchol raises gp.jitter) and
    the data pointers, so set_params / new data / a different jitter all miss and rebuild."""
    key = key + (gp.X.data_ptr(), gp.y.data_ptr(), int(gp.X.shape[0]))
    c = getattr(gp, "_refined", None)
    if c is None or c.get("key") != key:
        c = {"key": key}
        gp._refined = c
    return c


# ------------------------------------------------------------------------------------------
class HelmholtzGP:
    """Fit state of one snapshot on one GPU.

    fit(): K + (noise+jitter) I -> Cholesky / L^-1 -> alpha, LML  (gp2d_fit)
    predict(Xs): fused K* / mean / variance                       (gp2d_predict)
    lml_and_grad(): LML and d/d(l_df, l_cf, ratio, noise)          (gp2d_lml_grad)
    """

    def __init__(self, X, y, l_df, l_cf, ratio, noise, jitter=0.0, device=None):
        self.X = _points(X, device)
        self.N = int(self.X.shape[0])
        self.y = as_dev(y, self.X.device).reshape(-1)
        if self.y.numel() != 2 * self.N:
            raise ValueError("y must stack both components: length 2N (GP_laser.py:98,174)")
        self.set_params(l_df, l_cf, ratio, noise)
        self.jitter = float(jitter)
        self.ws_bytes = lib.gp2d_fit_workspace_bytes(self.N)
        self.ws = torch.empty(self.ws_bytes, dtype=torch.uint8, device=self.X.device)
        self._scal = torch.zeros(8, dtype=torch.float64, device=self.X.device)
        self._info = torch.zeros(1, dtype=torch.int32, device=self.X.device)
        self._pws = None                 # predict scratch (K* panels), grown on demand
        self.fitted = False
        self.lml = None

    @property
    def device(self):
        return self.X.device

    def set_params(self, l_df, l_cf, ratio, noise):
        self.l_df, self.l_cf, self.ratio, self.noise = float(l_df), float(l_cf), float(ratio), float(noise)
        self.fitted = False

    def fit_async(self, alpha_out=None):
        with torch.cuda.device(self.device):
            check(lib.gp2d_fit(_ptr(self.X), self.N, _ptr(self.y), self.l_df, self.l_cf, self.ratio,
                               self.noise, self.jitter, _ptr(self.ws), self.ws_bytes, _ptr(alpha_out),
                               _ptr(self._scal), _ptr(self._info), _stream()), "gp2d_fit")
        self.fitted = True

    def fit(self):
        """Returns the log marginal likelihood; raises LinAlgError when not PD."""
        self.fit_async()
        info = int(self._info.item())
        if info > 0:
            self.fitted = False
            raise LinAlgError("covariance not positive definite (pivot %d)" % info)
        self.lml = float(self._scal[0].item())
        return self.lml

    def predict_state(self) -> torch.Tensor:
        """Byte view of the part of the fit workspace that predict() reads (one contiguous range);
        dist.broadcast_fit ships it to the ranks that predict other grid shards."""
        import ctypes as C
        off, nb = C.c_size_t(), C.c_size_t()
        check(lib.gp2d_fit_predict_state(self.N, C.byref(off), C.byref(nb)), "gp2d_fit_predict_state")
        return self.ws[off.value:off.value + nb.value]

    def alpha(self) -> torch.Tensor:
        """K^-1 y in the caller's stacked order."""
        out = torch.empty(2 * self.N, dtype=torch.float64, device=self.device)
        self.fit_async(alpha_out=out)
        info = int(self._info.item())
        if info > 0:
            self.fitted = False
            raise LinAlgError("covariance not positive definite (pivot %d)" % info)
        return out

    def predict(self, Xs, include_noise=False, out_mean=None, out_var=None):
        """mean[2M], var[2M] (component-major) as device tensors."""
        if not self.fitted:
            self.fit()
        Xsd = _points(Xs, self.device)
        M = int(Xsd.shape[0])
        mean = out_mean if out_mean is not None else torch.empty(2 * M, dtype=torch.float64, device=self.device)
        var = out_var if out_var is not None else torch.empty(2 * M, dtype=torch.float64, device=self.device)
        if M:
            with torch.cuda.device(self.device):
                nb = lib.gp2d_predict_workspace_bytes(self.N, M)
                if self._pws is None or self._pws.numel() < nb:
                    self._pws = torch.empty(nb, dtype=torch.uint8, device=self.device)
                check(lib.gp2d_predict(_ptr(self.ws), self.N, self.l_df, self.l_cf, self.ratio, _ptr(Xsd), M, M,
                                       self.noise if include_noise else 0.0, _ptr(mean), _ptr(var),
                                       _ptr(self._pws), self._pws.numel(), _stream()), "gp2d_predict")
        return mean, var

    def kss(self):
        """Largest prior variance k(x, x) of a component."""
        return self.ratio / self.l_df ** 2 + (1.0 - self.ratio) / self.l_cf ** 2

    def cond_bound(self):
        """n k** / (noise + jitter) >= cond(K + (noise + jitter) I), from the arguments alone."""
        d = self.noise + self.jitter
        return float("inf") if d <= 0 else 2 * self.N * self.kss() / d

    def predict_refined(self, Xs, include_noise=False, steps=3, chunk_elems=REFINE_CHUNK):
        """Iterated-solve prediction for ill-conditioned covariances, see refined_predict()."""
        Xsd = _points(Xs, self.device)
        th = (self.l_df, self.l_cf, self.ratio)
        return refined_predict(lambda A, B, d=0.0: kernel_K(A, B, *th, diag_add=d), self.X, self.y, Xsd, 2,
                               lambda m: kernel_Kdiag(m, *th, device=self.device), self.noise + self.jitter,
                               self.noise if include_noise else 0.0, steps, chunk_elems,
                               cache=_refine_cache(self, (th, self.noise, self.jitter, int(steps))))

    def lml_and_grad(self, reference_compat=False):
        """(LML, grad[4]) with grad over (l_df, l_cf, ratio, noise) as host floats."""
        with torch.cuda.device(self.device):
            check(lib.gp2d_lml_grad(_ptr(self.X), self.N, _ptr(self.y), self.l_df, self.l_cf, self.ratio,
                                    self.noise, self.jitter, int(bool(reference_compat)), _ptr(self.ws),
                                    self.ws_bytes, _ptr(self._scal), _ptr(self._info), _stream()),
                  "gp2d_lml_grad")
        self.fitted = True
        host = self._scal[:5].cpu().numpy()
        info = int(self._info.item())
        if info > 0:
            self.fitted = False
            raise LinAlgError("covariance not positive definite (pivot %d)" % info)
        self.lml = float(host[0])
        return self.lml, host[1:5].copy()


# ------------------------------------------------------------------------------------------
class HelmholtzBatch:
    """B independent Helmholtz GPs of one size advanced together (gp2d_fit_batched /
    gp2d_lml_grad_batched): every kernel launch of the fit covers the whole batch.

    Two uses, both from the reference: the restarts of ``optimize_restarts`` (krig.py:450;
    GP_plots.py:765) -- one data set, B hyper-parameter points: ``X [N,2]``, ``y [2N]`` -- and the time
    slices / snapshots of the predict loop (krig.py:541-557) -- ``X [B,N,2]``, ``y [B,2N]``.
    Results are bit-identical to B separate HelmholtzGP calls."""

    def __init__(self, X, y, B=None, jitter=0.0, device=None):
        Xd = as_dev(X, device)
        if Xd.dim() == 2:
            if Xd.shape[1] != 2:
                raise ValueError("points must be [N,2] or [B,N,2]")
            if B is None:
                raise ValueError("B is required when the problems share one data set")
            self.shared = True
            self.N = int(Xd.shape[0])
            self.B = int(B)
        else:
            if Xd.dim() != 3 or Xd.shape[2] != 2:
                raise ValueError("points must be [N,2] or [B,N,2]")
            self.shared = False
            self.B, self.N = int(Xd.shape[0]), int(Xd.shape[1])
            if B is not None and int(B) != self.B:
                raise ValueError("B does not match X")
        self.X = Xd
        self.y = as_dev(y, Xd.device).reshape(-1) if self.shared else as_dev(y, Xd.device).reshape(self.B, -1)
        if self.y.shape[-1] != 2 * self.N:
            raise ValueError("y must stack both components: length 2N per problem")
        self.jitter = float(jitter)
        self.ws_stride = lib.gp2d_fit_workspace_bytes(self.N)
        if not self.ws_stride:
            raise ValueError("problem size out of range")
        self.ws = torch.empty(self.ws_stride * self.B, dtype=torch.uint8, device=Xd.device)
        self._out = torch.zeros((self.B, 5), dtype=torch.float64, device=Xd.device)
        self._info = torch.zeros(self.B, dtype=torch.int32, device=Xd.device)
        self._pws = None
        self.theta4 = None

    @property
    def device(self):
        return self.X.device

    def _theta(self, theta4, nb):
        t = np.ascontiguousarray(np.asarray(theta4, dtype=np.float64).reshape(-1, 4))
        if t.shape[0] != nb:
            raise ValueError("theta4 must be [%d,4] rows (l_df, l_cf, ratio, noise)" % nb)
        return t

    def _strides(self):
        return (0, 0) if self.shared else (2 * self.N, 2 * self.N)

    def fit_async(self, theta4, nb=None, alpha_out=None):
        """Fit the first ``nb`` problems (default all) at theta4[nb,4]; LML and info stay on the device
        (``lml_device()``, ``info_device()``)."""
        nb = self.B if nb is None else int(nb)
        t = self._theta(theta4, nb)
        xs, ys = self._strides()
        with torch.cuda.device(self.device):
            check(lib.gp2d_fit_batched(_ptr(self.X), xs, self.N, _ptr(self.y), ys, nb, t.ctypes.data, self.jitter,
                                       _ptr(self.ws), self.ws.numel(), _ptr(alpha_out), _ptr(self._out), _ptr(self._info),
                                       _stream()), "gp2d_fit_batched")
        self.theta4 = t
        self._nb = nb

    def fit(self, theta4, nb=None):
        """Returns (lml[nb], info[nb]) as numpy arrays (info > 0: not positive definite, LAPACK-style)."""
        self.fit_async(theta4, nb)
        lml = self._out.reshape(-1)[:self._nb].cpu().numpy()
        return lml, self._info[:self._nb].cpu().numpy()

    def lml_and_grad(self, theta4, nb=None, reference_compat=False):
        """(lml[nb], grad[nb,4] over (l_df, l_cf, ratio, noise), info[nb]) -- one batched objective evaluation."""
        nb = self.B if nb is None else int(nb)
        t = self._theta(theta4, nb)
        xs, ys = self._strides()
        with torch.cuda.device(self.device):
            check(lib.gp2d_lml_grad_batched(_ptr(self.X), xs, self.N, _ptr(self.y), ys, nb, t.ctypes.data, self.jitter,
                                            int(bool(reference_compat)), _ptr(self.ws), self.ws.numel(), _ptr(self._out),
                                            _ptr(self._info), _stream()), "gp2d_lml_grad_batched")
        self.theta4 = t
        self._nb = nb
        host = self._out[:nb].cpu().numpy()
        return host[:, 0].copy(), host[:, 1:5].copy(), self._info[:nb].cpu().numpy()

    def state(self, b):
        """Fit workspace of problem b (a view): what gp2d_predict reads."""
        return self.ws[b * self.ws_stride:(b + 1) * self.ws_stride]

    def predict(self, b, Xs, include_noise=False, out_mean=None, out_var=None):
        """Fused prediction from problem b's fit state: mean[2M], var[2M] device tensors."""
        if self.theta4 is None or b >= self._nb:
            raise Gp2dError("problem %d has not been fitted" % b)
        Xsd = _points(Xs, self.device)
        M = int(Xsd.shape[0])
        mean = out_mean if out_mean is not None else torch.empty(2 * M, dtype=torch.float64, device=self.device)
        var = out_var if out_var is not None else torch.empty(2 * M, dtype=torch.float64, device=self.device)
        l_df, l_cf, ratio, noise = (float(v) for v in self.theta4[b])
        if M:
            with torch.cuda.device(self.device):
                nbytes = lib.gp2d_predict_workspace_bytes(self.N, M)
                if self._pws is None or self._pws.numel() < nbytes:
                    self._pws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
                check(lib.gp2d_predict(_ptr(self.state(b)), self.N, l_df, l_cf, ratio, _ptr(Xsd), M, M,
                                       noise if include_noise else 0.0, _ptr(mean), _ptr(var), _ptr(self._pws),
                                       self._pws.numel(), _stream()), "gp2d_predict")
        return mean, var


def krig_snapshots(X, y, Xs, l_df, l_cf, ratio, noise, jitter=0.0, include_noise=False, batch=4, device=None):
    """Independent snapshots kriged with one hyper-parameter set: X [S,N,2], y [S,2N], Xs [M,2] (one grid for
    all) or [S,M,2].  The fits of ``batch`` snapshots share every kernel launch (gp2d_fit_batched); each
    snapshot is then predicted by the fused kernel.  Returns (mean [S,2M], var [S,2M], lml [S]) as numpy
    arrays; raises LinAlgError for a snapshot whose covariance is not positive definite.  This is the
    per-time-slice loop of krig.predict (krig.py:539-557) for the Helmholtz kernel."""
    X = np.ascontiguousarray(X, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64).reshape(X.shape[0], -1)
    Xs = np.ascontiguousarray(Xs, dtype=np.float64)
    S, N = X.shape[0], X.shape[1]
    per_grid = Xs.ndim == 3
    M = Xs.shape[-2]
    dev = _device(device)
    batch = max(1, min(int(batch), S))
    mean = np.empty((S, 2 * M))
    var = np.empty((S, 2 * M))
    lml = np.empty(S)
    theta = np.tile(np.array([l_df, l_cf, ratio, noise], dtype=np.float64), (batch, 1))
    hb = HelmholtzBatch(torch.zeros((batch, N, 2), dtype=torch.float64, device=dev),
                        torch.zeros((batch, 2 * N), dtype=torch.float64, device=dev), jitter=jitter)
    Xsd = None if per_grid else as_dev(Xs, dev)
    dm = torch.empty((batch, 2 * M), dtype=torch.float64, device=dev)
    dv = torch.empty((batch, 2 * M), dtype=torch.float64, device=dev)
    for s0 in range(0, S, batch):
        nb = min(batch, S - s0)
        hb.X[:nb].copy_(torch.from_numpy(X[s0:s0 + nb]))
        hb.y[:nb].copy_(torch.from_numpy(y[s0:s0 + nb]))
        l, info = hb.fit(theta[:nb], nb)
        if np.any(info > 0):
            bad = int(np.argmax(info > 0))
            raise LinAlgError("snapshot %d: covariance not positive definite (pivot %d)" % (s0 + bad, info[bad]))
        for b in range(nb):
            g = as_dev(Xs[s0 + b], dev) if per_grid else Xsd
            hb.predict(b, g, include_noise=include_noise, out_mean=dm[b], out_var=dv[b])
        mean[s0:s0 + nb] = dm[:nb].cpu().numpy()
        var[s0:s0 + nb] = dv[:nb].cpu().numpy()
        lml[s0:s0 + nb] = l
    return mean, var, lml


# ------------------------------------------------------------------------------------------
# space-time product kernel Kt(t) * Helmholtz(a, b)  (scratch.py:506-508; myKernel.py:337-363)
# ------------------------------------------------------------------------------------------
def _points3(X, device=None) -> torch.Tensor:
    t = as_dev(X, device)
    if t.dim() != 2 or t.shape[1] != 3:
        raise ValueError("space-time points must be [N,3] rows (t, a, b)")
    return t


def st_K(X3, X3b, l_df, l_cf, ratio, tvar, lt, diag_add=0.0) -> torch.Tensor:
    """[2N,2M] tvar exp(-dt^2/2lt^2) * Helmholtz block matrix (device tensor)."""
    Xd = _points3(X3)
    X2d = None if X3b is None else _points3(X3b, Xd.device)
    N, M = Xd.shape[0], (Xd.shape[0] if X2d is None else X2d.shape[0])
    out = torch.empty((2 * N, 2 * M), dtype=torch.float64, device=Xd.device)
    if N and M:
        with torch.cuda.device(Xd.device):
            check(lib.gp2d_st_kernel_build(_ptr(Xd), N, _ptr(X2d), M, l_df, l_cf, ratio, tvar, lt, diag_add, _ptr(out),
                                           out.stride(0), _stream()), "gp2d_st_kernel_build")
    return out


def st_grad_sums(dL_dK, X3, X3b, l_df, l_cf, ratio, tvar, lt) -> torch.Tensor:
    """sum(dK/dtheta * dL_dK) for theta = (l_df, l_cf, ratio, tvar, lt) -> device tensor[5]."""
    Xd = _points3(X3)
    X2d = None if X3b is None else _points3(X3b, Xd.device)
    N, M = Xd.shape[0], (Xd.shape[0] if X2d is None else X2d.shape[0])
    W = as_dev(dL_dK, Xd.device)
    if tuple(W.shape) != (2 * N, 2 * M):
        raise ValueError("dL_dK must be [2N,2M]")
    nb = lib.gp2d_kernel_grad_workspace_bytes(N, M)
    ws = torch.empty(nb, dtype=torch.uint8, device=Xd.device)
    out = torch.empty(5, dtype=torch.float64, device=Xd.device)
    with torch.cuda.device(Xd.device):
        check(lib.gp2d_st_kernel_grad(_ptr(Xd), N, _ptr(X2d), M, l_df, l_cf, ratio, tvar, lt, _ptr(W), W.stride(0),
                                      _ptr(ws), nb, _ptr(out), _stream()), "gp2d_st_kernel_grad")
    return out


class SpaceTimeGP:
    """HelmholtzGP with the time factor: points [N,3] rows (t, a, b), y the stacked components."""

    def __init__(self, X3, y, l_df, l_cf, ratio, tvar, lt, noise, jitter=0.0, device=None):
        self.X = _points3(X3, device)
        self.N = int(self.X.shape[0])
        self.y = as_dev(y, self.X.device).reshape(-1)
        if self.y.numel() != 2 * self.N:
            raise ValueError("y must stack both components: length 2N")
        self.set_params(l_df, l_cf, ratio, tvar, lt, noise)
        self.jitter = float(jitter)
        self.ws_bytes = lib.gp2d_st_fit_workspace_bytes(self.N)
        self.ws = torch.empty(self.ws_bytes, dtype=torch.uint8, device=self.X.device)
        self._scal = torch.zeros(8, dtype=torch.float64, device=self.X.device)
        self._info = torch.zeros(1, dtype=torch.int32, device=self.X.device)
        self._pws = None
        self.lml = None

    @property
    def device(self):
        return self.X.device

    def set_params(self, l_df, l_cf, ratio, tvar, lt, noise):
        self.theta = tuple(float(v) for v in (l_df, l_cf, ratio, tvar, lt))
        self.noise = float(noise)
        self.fitted = False

    def fit_async(self, alpha_out=None):
        with torch.cuda.device(self.device):
            check(lib.gp2d_st_fit(_ptr(self.X), self.N, _ptr(self.y), *self.theta, self.noise, self.jitter, _ptr(self.ws),
                                  self.ws_bytes, _ptr(alpha_out), _ptr(self._scal), _ptr(self._info), _stream()),
                  "gp2d_st_fit")
        self.fitted = True

    def fit(self):
        self.fit_async()
        info = int(self._info.item())
        if info > 0:
            self.fitted = False
            raise LinAlgError("covariance not positive definite (pivot %d)" % info)
        self.lml = float(self._scal[0].item())
        return self.lml

    def alpha(self) -> torch.Tensor:
        out = torch.empty(2 * self.N, dtype=torch.float64, device=self.device)
        self.fit_async(alpha_out=out)
        info = int(self._info.item())
        if info > 0:
            self.fitted = False
            raise LinAlgError("covariance not positive definite (pivot %d)" % info)
        return out

    def predict_state(self) -> torch.Tensor:
        import ctypes as C
        off, nb = C.c_size_t(), C.c_size_t()
        check(lib.gp2d_st_fit_predict_state(self.N, C.byref(off), C.byref(nb)), "gp2d_st_fit_predict_state")
        return self.ws[off.value:off.value + nb.value]

    def predict(self, Xs3, include_noise=False, out_mean=None, out_var=None):
        if not self.fitted:
            self.fit()
        Xsd = _points3(Xs3, self.device)
        M = int(Xsd.shape[0])
        mean = out_mean if out_mean is not None else torch.empty(2 * M, dtype=torch.float64, device=self.device)
        var = out_var if out_var is not None else torch.empty(2 * M, dtype=torch.float64, device=self.device)
        if M:
            with torch.cuda.device(self.device):
                nb = lib.gp2d_predict_workspace_bytes(self.N, M)
                if self._pws is None or self._pws.numel() < nb:
                    self._pws = torch.empty(nb, dtype=torch.uint8, device=self.device)
                check(lib.gp2d_st_predict(_ptr(self.ws), self.N, *self.theta, _ptr(Xsd), M, M,
                                          self.noise if include_noise else 0.0, _ptr(mean), _ptr(var), _ptr(self._pws),
                                          self._pws.numel(), _stream()), "gp2d_st_predict")
        return mean, var

    def kss(self):
        l_df, l_cf, ratio, tvar, _ = self.theta
        return tvar * (ratio / l_df ** 2 + (1.0 - ratio) / l_cf ** 2)

    def cond_bound(self):
        d = self.noise + self.jitter
        return float("inf") if d <= 0 else 2 * self.N * self.kss() / d

    def predict_refined(self, Xs3, include_noise=False, steps=3, chunk_elems=REFINE_CHUNK):
        """Iterated-solve prediction for ill-conditioned covariances, see refined_predict()."""
        Xsd = _points3(Xs3, self.device)
        l_df, l_cf, ratio, tvar, _ = self.theta
        return refined_predict(lambda A, B, d=0.0: st_K(A, B, *self.theta, diag_add=d), self.X, self.y, Xsd, 2,
                               lambda m: tvar * kernel_Kdiag(m, l_df, l_cf, ratio, device=self.device),
                               self.noise + self.jitter, self.noise if include_noise else 0.0, steps, chunk_elems,
                               cache=_refine_cache(self, (self.theta, self.noise, self.jitter, int(steps))))

    def lml_and_grad(self):
        """(LML, grad[6]) over (l_df, l_cf, ratio, tvar, lt, noise)."""
        with torch.cuda.device(self.device):
            check(lib.gp2d_st_lml_grad(_ptr(self.X), self.N, _ptr(self.y), *self.theta, self.noise, self.jitter,
                                       _ptr(self.ws), self.ws_bytes, _ptr(self._scal), _ptr(self._info), _stream()),
                  "gp2d_st_lml_grad")
        self.fitted = True
        host = self._scal[:7].cpu().numpy()
        info = int(self._info.item())
        if info > 0:
            self.fitted = False
            raise LinAlgError("covariance not positive definite (pivot %d)" % info)
        self.lml = float(host[0])
        return self.lml, host[1:7].copy()


# ------------------------------------------------------------------------------------------
# scalar ARD-RBF sum family (krig.py:174-181,388,405-407)
# ------------------------------------------------------------------------------------------
def _rbf_params(variances, lengthscales, D):
    """Host arrays var[Q], ls[Q*D] for the C ABI (kept alive by the caller during the call)."""
    var = np.ascontiguousarray(np.atleast_1d(np.asarray(variances, dtype=np.float64)))
    ls = np.asarray(lengthscales, dtype=np.float64)
    Q = var.size
    if ls.ndim == 0:
        ls = np.full((Q, D), float(ls))
    ls = np.atleast_2d(ls)
    if ls.shape == (Q, 1) and D > 1:
        ls = np.repeat(ls, D, axis=1)
    if ls.shape != (Q, D):
        raise ValueError("lengthscales must be [Q,D] = [%d,%d]" % (Q, D))
    return var, np.ascontiguousarray(ls), Q


def _coords(X, device=None) -> torch.Tensor:
    t = as_dev(X, device)
    if t.dim() != 2 or not (1 <= t.shape[1] <= 4):
        raise ValueError("inputs must be [N,D] with 1 <= D <= 4")
    return t


def rbf_K(X, X2, variances, lengthscales, diag_add=0.0, out=None) -> torch.Tensor:
    """[N,M] sum of ARD squared-exponential kernels (device tensor)."""
    Xd = _coords(X)
    X2d = None if X2 is None else _coords(X2, Xd.device)
    N, D = Xd.shape
    M = N if X2d is None else X2d.shape[0]
    var, ls, Q = _rbf_params(variances, lengthscales, D)
    if out is None:
        out = torch.empty((N, M), dtype=torch.float64, device=Xd.device)
    if N and M:
        with torch.cuda.device(Xd.device):
            check(lib.gp2d_rbf_kernel_build(_ptr(Xd), N, _ptr(X2d), M, D, Q, var.ctypes.data, ls.ctypes.data, diag_add,
                                            _ptr(out), out.stride(0), _stream()), "gp2d_rbf_kernel_build")
    return out


def rbf_grad_sums(dL_dK, X, X2, variances, lengthscales) -> torch.Tensor:
    """sum(dK/dtheta * dL_dK), theta ordered (variance_q, lengthscale_q[0..D-1]) per component."""
    Xd = _coords(X)
    X2d = None if X2 is None else _coords(X2, Xd.device)
    N, D = Xd.shape
    M = N if X2d is None else X2d.shape[0]
    var, ls, Q = _rbf_params(variances, lengthscales, D)
    W = as_dev(dL_dK, Xd.device)
    if tuple(W.shape) != (N, M):
        raise ValueError("dL_dK must be [N,M]")
    nb = lib.gp2d_rbf_kernel_grad_workspace_bytes(N, M)
    ws = torch.empty(nb, dtype=torch.uint8, device=Xd.device)
    out = torch.empty(Q * (1 + D), dtype=torch.float64, device=Xd.device)
    with torch.cuda.device(Xd.device):
        check(lib.gp2d_rbf_kernel_grad(_ptr(Xd), N, _ptr(X2d), M, D, Q, var.ctypes.data, ls.ctypes.data, _ptr(W),
                                       W.stride(0), _ptr(ws), nb, _ptr(out), _stream()), "gp2d_rbf_kernel_grad")
    return out


class ScalarGP:
    """Fit state of one scalar GP (sum of ARD-RBF kernels + white noise) on one GPU; same life
    cycle as HelmholtzGP: fit() / predict() / lml_and_grad() through gp2d_rbf_*."""

    def __init__(self, X, y, variances, lengthscales, noise, jitter=0.0, device=None):
        self.X = _coords(X, device)
        self.N, self.D = int(self.X.shape[0]), int(self.X.shape[1])
        self.y = as_dev(y, self.X.device).reshape(-1)
        if self.y.numel() != self.N:
            raise ValueError("y must hold one observation per row of X")
        self.set_params(variances, lengthscales, noise)
        self.jitter = float(jitter)
        self.ws_bytes = lib.gp2d_rbf_fit_workspace_bytes(self.N, self.D)
        self.ws = torch.empty(self.ws_bytes, dtype=torch.uint8, device=self.X.device)
        self._scal = torch.zeros(32, dtype=torch.float64, device=self.X.device)
        self._info = torch.zeros(1, dtype=torch.int32, device=self.X.device)
        self._pws = None
        self.lml = None

    @property
    def device(self):
        return self.X.device

    def set_params(self, variances, lengthscales, noise):
        self.var, self.ls, self.Q = _rbf_params(variances, lengthscales, self.D)
        self.noise = float(noise)
        self.fitted = False

    def fit_async(self, alpha_out=None):
        with torch.cuda.device(self.device):
            check(lib.gp2d_rbf_fit(_ptr(self.X), self.N, self.D, _ptr(self.y), self.Q, self.var.ctypes.data,
                                   self.ls.ctypes.data, self.noise, self.jitter, _ptr(self.ws), self.ws_bytes,
                                   _ptr(alpha_out), _ptr(self._scal), _ptr(self._info), _stream()), "gp2d_rbf_fit")
        self.fitted = True

    def fit(self):
        self.fit_async()
        info = int(self._info.item())
        if info > 0:
            self.fitted = False
            raise LinAlgError("covariance not positive definite (pivot %d)" % info)
        self.lml = float(self._scal[0].item())
        return self.lml

    def alpha(self) -> torch.Tensor:
        out = torch.empty(self.N, dtype=torch.float64, device=self.device)
        self.fit_async(alpha_out=out)
        info = int(self._info.item())
        if info > 0:
            self.fitted = False
            raise LinAlgError("covariance not positive definite (pivot %d)" % info)
        return out

    def predict_state(self) -> torch.Tensor:
        import ctypes as C
        off, nb = C.c_size_t(), C.c_size_t()
        check(lib.gp2d_rbf_fit_predict_state(self.N, self.D, C.byref(off), C.byref(nb)), "gp2d_rbf_fit_predict_state")
        return self.ws[off.value:off.value + nb.value]

    def predict(self, Xs, include_noise=False, out_mean=None, out_var=None):
        """mean[M], var[M] as device tensors; include_noise adds the white-noise variance (what
        sklearn's predict(return_std=True) with a WhiteKernel and GPy's predict both return)."""
        if not self.fitted:
            self.fit()
        Xsd = _coords(Xs, self.device)
        if Xsd.shape[1] != self.D:
            raise ValueError("prediction points must have %d columns" % self.D)
        M = int(Xsd.shape[0])
        mean = out_mean if out_mean is not None else torch.empty(M, dtype=torch.float64, device=self.device)
        var = out_var if out_var is not None else torch.empty(M, dtype=torch.float64, device=self.device)
        if M:
            with torch.cuda.device(self.device):
                nb = lib.gp2d_rbf_predict_workspace_bytes(self.N, M)
                if self._pws is None or self._pws.numel() < nb:
                    self._pws = torch.empty(nb, dtype=torch.uint8, device=self.device)
                check(lib.gp2d_rbf_predict(_ptr(self.ws), self.N, self.D, self.Q, self.var.ctypes.data,
                                           self.ls.ctypes.data, _ptr(Xsd), M, self.noise if include_noise else 0.0,
                                           _ptr(mean), _ptr(var), _ptr(self._pws), self._pws.numel(), _stream()),
                      "gp2d_rbf_predict")
        return mean, var

    def kss(self):
        return float(np.sum(self.var))

    def cond_bound(self):
        d = self.noise + self.jitter
        return float("inf") if d <= 0 else self.N * self.kss() / d

    def predict_refined(self, Xs, include_noise=False, steps=3, chunk_elems=REFINE_CHUNK):
        """Iterated-solve prediction for ill-conditioned covariances, see refined_predict()."""
        Xsd = _coords(Xs, self.device)
        if Xsd.shape[1] != self.D:
            raise ValueError("prediction points must have %d columns" % self.D)
        kss = float(np.sum(self.var))
        return refined_predict(lambda A, B, d=0.0: rbf_K(A, B, self.var, self.ls, diag_add=d), self.X, self.y, Xsd, 1,
                               lambda m: torch.full((m,), kss, dtype=torch.float64, device=self.device),
                               self.noise + self.jitter, self.noise if include_noise else 0.0, steps, chunk_elems,
                               cache=_refine_cache(self, (self.var.tobytes(), self.ls.tobytes(), self.noise, self.jitter, int(steps))))

    def lml_and_grad(self):
        """(LML, grad) with grad over (variance_q, lengthscale_q[..])_q then the noise variance."""
        ng = self.Q * (1 + self.D) + 1
        with torch.cuda.device(self.device):
            check(lib.gp2d_rbf_lml_grad(_ptr(self.X), self.N, self.D, _ptr(self.y), self.Q, self.var.ctypes.data,
                                        self.ls.ctypes.data, self.noise, self.jitter, _ptr(self.ws), self.ws_bytes,
                                        _ptr(self._scal), _ptr(self._info), _stream()), "gp2d_rbf_lml_grad")
        self.fitted = True
        host = self._scal[:1 + ng].cpu().numpy()
        info = int(self._info.item())
        if info > 0:
            self.fitted = False
            raise LinAlgError("covariance not positive definite (pivot %d)" % info)
        self.lml = float(host[0])
        return self.lml, host[1:1 + ng].copy()


# ------------------------------------------------------------------------------------------
# sum of space-time Helmholtz terms (krig.py:396-407: myKernel2.divFreeK / curlFreeK, their sum,
# nKernels copies)
# ------------------------------------------------------------------------------------------
HSUM_MAXQ = 8


def _hsum_params(types, params):
    """Host arrays type[Q] (0 divergence-free, 1 curl-free) and params[Q,4] = (var, lt, la, lb)."""
    ty = np.ascontiguousarray(np.atleast_1d(np.asarray(types)).astype(np.int32))
    pr = np.ascontiguousarray(np.atleast_2d(np.asarray(params, dtype=np.float64)))
    Q = int(ty.shape[0])
    if not (1 <= Q <= HSUM_MAXQ):
        raise ValueError("between 1 and %d terms" % HSUM_MAXQ)
    if pr.shape != (Q, 4):
        raise ValueError("params must be [Q,4] rows (var, lt, la, lb)")
    if np.any((ty != 0) & (ty != 1)):
        raise ValueError("term types are 0 (divergence-free) or 1 (curl-free)")
    return ty, pr, Q


def _points23(X, device=None) -> torch.Tensor:
    t = as_dev(X, device)
    if t.dim() != 2 or t.shape[1] not in (2, 3):
        raise ValueError("points must be [N,2] rows (a, b) or [N,3] rows (t, a, b)")
    return t


def hsum_K(X, X2, types, params, diag_add=0.0) -> torch.Tensor:
    """[2N,2M] block matrix of the term sum (device tensor)."""
    Xd = _points23(X)
    X2d = None if X2 is None else _points23(X2, Xd.device)
    N, M = Xd.shape[0], (Xd.shape[0] if X2d is None else X2d.shape[0])
    ldx = int(Xd.shape[1])
    if X2d is not None and X2d.shape[1] != ldx:
        raise ValueError("X and X2 must have the same number of columns")
    ty, pr, Q = _hsum_params(types, params)
    out = torch.empty((2 * N, 2 * M), dtype=torch.float64, device=Xd.device)
    if N and M:
        with torch.cuda.device(Xd.device):
            check(lib.gp2d_hsum_kernel_build(_ptr(Xd), N, _ptr(X2d), M, ldx, Q, ty.ctypes.data, pr.ctypes.data, diag_add,
                                             _ptr(out), out.stride(0), _stream()), "gp2d_hsum_kernel_build")
    return out


def hsum_Kdiag(M, ldx, types, params, device=None) -> torch.Tensor:
    ty, pr, Q = _hsum_params(types, params)
    dev = device or _device()
    out = torch.empty(2 * M, dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        check(lib.gp2d_hsum_kdiag(M, ldx, Q, ty.ctypes.data, pr.ctypes.data, _ptr(out), _stream()), "gp2d_hsum_kdiag")
    return out


def hsum_grad_sums(dL_dK, X, X2, types, params) -> torch.Tensor:
    """sum(dK/d(var, lt, la, lb)_q * dL_dK) -> device tensor [Q,4]."""
    Xd = _points23(X)
    X2d = None if X2 is None else _points23(X2, Xd.device)
    N, M = Xd.shape[0], (Xd.shape[0] if X2d is None else X2d.shape[0])
    ldx = int(Xd.shape[1])
    ty, pr, Q = _hsum_params(types, params)
    W = as_dev(dL_dK, Xd.device)
    if tuple(W.shape) != (2 * N, 2 * M):
        raise ValueError("dL_dK must be [2N,2M]")
    nb = lib.gp2d_hsum_kernel_grad_workspace_bytes(N, M, Q)
    ws = torch.empty(nb, dtype=torch.uint8, device=Xd.device)
    out = torch.empty((Q, 4), dtype=torch.float64, device=Xd.device)
    with torch.cuda.device(Xd.device):
        check(lib.gp2d_hsum_kernel_grad(_ptr(Xd), N, _ptr(X2d), M, ldx, Q, ty.ctypes.data, pr.ctypes.data, _ptr(W),
                                        W.stride(0), _ptr(ws), nb, _ptr(out), _stream()), "gp2d_hsum_kernel_grad")
    return out


class HelmholtzSumGP:
    """Fit state of a GP whose covariance is a sum of divergence-free / curl-free space-time terms;
    same life cycle as HelmholtzGP through gp2d_hsum_*.  Points [N,2] (a, b) or [N,3] (t, a, b)."""

    def __init__(self, X, y, types, params, noise, jitter=0.0, device=None):
        self.X = _points23(X, device)
        self.N, self.ldx = int(self.X.shape[0]), int(self.X.shape[1])
        self.y = as_dev(y, self.X.device).reshape(-1)
        if self.y.numel() != 2 * self.N:
            raise ValueError("y must stack both components: length 2N")
        self.set_params(types, params, noise)
        self.jitter = float(jitter)
        self.ws_bytes = lib.gp2d_hsum_fit_workspace_bytes(self.N, self.ldx, self.Q)
        if not self.ws_bytes:
            raise ValueError("problem size out of range")
        self.ws = torch.empty(self.ws_bytes, dtype=torch.uint8, device=self.X.device)
        self._scal = torch.zeros(2 + 4 * HSUM_MAXQ, dtype=torch.float64, device=self.X.device)
        self._info = torch.zeros(1, dtype=torch.int32, device=self.X.device)
        self._pws = None
        self.lml = None

    @property
    def device(self):
        return self.X.device

    def set_params(self, types, params, noise):
        ty, pr, Q = _hsum_params(types, params)
        if getattr(self, "Q", Q) != Q:
            raise ValueError("the number of terms is fixed at construction (workspace layout)")
        self.types, self.params, self.Q = ty, pr, Q
        self.noise = float(noise)
        self.fitted = False

    def fit_async(self, alpha_out=None):
        with torch.cuda.device(self.device):
            check(lib.gp2d_hsum_fit(_ptr(self.X), self.N, self.ldx, _ptr(self.y), self.Q, self.types.ctypes.data,
                                    self.params.ctypes.data, self.noise, self.jitter, _ptr(self.ws), self.ws_bytes,
                                    _ptr(alpha_out), _ptr(self._scal), _ptr(self._info), _stream()), "gp2d_hsum_fit")
        self.fitted = True

    def fit(self):
        self.fit_async()
        info = int(self._info.item())
        if info > 0:
            self.fitted = False
            raise LinAlgError("covariance not positive definite (pivot %d)" % info)
        self.lml = float(self._scal[0].item())
        return self.lml

    def alpha(self) -> torch.Tensor:
        out = torch.empty(2 * self.N, dtype=torch.float64, device=self.device)
        self.fit_async(alpha_out=out)
        info = int(self._info.item())
        if info > 0:
            self.fitted = False
            raise LinAlgError("covariance not positive definite (pivot %d)" % info)
        return out

    def predict_state(self) -> torch.Tensor:
        import ctypes as C
        off, nb = C.c_size_t(), C.c_size_t()
        check(lib.gp2d_hsum_fit_predict_state(self.N, self.ldx, self.Q, C.byref(off), C.byref(nb)),
              "gp2d_hsum_fit_predict_state")
        return self.ws[off.value:off.value + nb.value]

    def predict(self, Xs, include_noise=False, out_mean=None, out_var=None):
        if not self.fitted:
            self.fit()
        Xsd = _points23(Xs, self.device)
        if Xsd.shape[1] != self.ldx:
            raise ValueError("prediction points must have %d columns" % self.ldx)
        M = int(Xsd.shape[0])
        mean = out_mean if out_mean is not None else torch.empty(2 * M, dtype=torch.float64, device=self.device)
        var = out_var if out_var is not None else torch.empty(2 * M, dtype=torch.float64, device=self.device)
        if M:
            with torch.cuda.device(self.device):
                nb = lib.gp2d_predict_workspace_bytes(self.N, M)
                if self._pws is None or self._pws.numel() < nb:
                    self._pws = torch.empty(nb, dtype=torch.uint8, device=self.device)
                check(lib.gp2d_hsum_predict(_ptr(self.ws), self.N, self.ldx, self.Q, self.types.ctypes.data,
                                            self.params.ctypes.data, _ptr(Xsd), M, M,
                                            self.noise if include_noise else 0.0, _ptr(mean), _ptr(var),
                                            _ptr(self._pws), self._pws.numel(), _stream()), "gp2d_hsum_predict")
        return mean, var

    def kss(self):
        v0 = sum(p[0] / (p[2] ** 2 if t else p[3] ** 2) for t, p in zip(self.types, self.params))
        v1 = sum(p[0] / (p[3] ** 2 if t else p[2] ** 2) for t, p in zip(self.types, self.params))
        return float(max(v0, v1))

    def cond_bound(self):
        d = self.noise + self.jitter
        return float("inf") if d <= 0 else 2 * self.N * self.kss() / d

    def predict_refined(self, Xs, include_noise=False, steps=3, chunk_elems=REFINE_CHUNK):
        """Iterated-solve prediction for ill-conditioned covariances, see refined_predict()."""
        Xsd = _points23(Xs, self.device)
        if Xsd.shape[1] != self.ldx:
            raise ValueError("prediction points must have %d columns" % self.ldx)
        return refined_predict(lambda A, B, d=0.0: hsum_K(A, B, self.types, self.params, diag_add=d), self.X, self.y, Xsd, 2,
                               lambda m: hsum_Kdiag(m, self.ldx, self.types, self.params, device=self.device),
                               self.noise + self.jitter, self.noise if include_noise else 0.0, steps, chunk_elems,
                               cache=_refine_cache(self, (self.types.tobytes(), self.params.tobytes(), self.noise, self.jitter,
                                                          int(steps))))

    def lml_and_grad(self):
        """(LML, grad) with grad over (var, lt, la, lb)_q for every term, then the noise variance."""
        ng = 4 * self.Q + 1
        with torch.cuda.device(self.device):
            check(lib.gp2d_hsum_lml_grad(_ptr(self.X), self.N, self.ldx, _ptr(self.y), self.Q, self.types.ctypes.data,
                                         self.params.ctypes.data, self.noise, self.jitter, _ptr(self.ws), self.ws_bytes,
                                         _ptr(self._scal), _ptr(self._info), _stream()), "gp2d_hsum_lml_grad")
        self.fitted = True
        host = self._scal[:1 + ng].cpu().numpy()
        info = int(self._info.item())
        if info > 0:
            self.fitted = False
            raise LinAlgError("covariance not positive definite (pivot %d)" % info)
        self.lml = float(host[0])
        return self.lml, host[1:1 + ng].copy()


def fit_predict_host(X, y, l_df, l_cf, ratio, noise, Xs, jitter=0.0, include_noise=False):
    """Whole pipeline through the host-pointer C entry point (numpy in, numpy out)."""
    import ctypes as C
    X = np.ascontiguousarray(X, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64).reshape(-1)
    Xs = np.ascontiguousarray(Xs, dtype=np.float64)
    N, M = X.shape[0], Xs.shape[0]
    mean = np.empty(2 * M)
    var = np.empty(2 * M)
    lml = np.zeros(1)
    _device()
    rc = lib.gp2d_fit_predict_host(X.ctypes.data, N, y.ctypes.data, l_df, l_cf, ratio, noise, jitter,
                                   Xs.ctypes.data, M, int(include_noise), mean.ctypes.data, var.ctypes.data,
                                   lml.ctypes.data)
    check(rc, "gp2d_fit_predict_host")
    if rc > 0:
        raise LinAlgError("covariance not positive definite (pivot %d)" % rc)
    return mean, var, float(lml[0])
