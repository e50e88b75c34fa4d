"""GPRegression-like model over the CUDA engine: the subset of GPy's model API the reference
scripts use (GP_plots.py:760-770; krig.py:411-412,438-457,478-483,543-544).

    m = GPRegression(X, Y, kern)            X [N,2]; Y [2N,1] stacked components
    m.optimize(); m.optimize_restarts(num_restarts=..., messages=False)
    mean, var = m.predict(Xnew)             [2M,1] each, Gaussian noise included (GPy)
    m.param_array                           [length_df, length_cf, ratio, noise]  (GP_plots.py:810-813)
    m.log_likelihood(); m.pickle(path); load(path)

Every objective evaluation is one gp2d_lml_grad call (kernel build + Cholesky + K^-1 +
gradient reductions on the GPU); only the L-BFGS-B driver (scipy, like paramz) is host code.
Independent restarts shard across ranks (``optimize_restarts(..., rank, world)``).
"""
from __future__ import annotations

import pickle

import numpy as np
import scipy.optimize as sopt

from .engine import HelmholtzBatch, HelmholtzGP, HelmholtzSumGP, LinAlgError, ROBUST_COND, ScalarGP, SpaceTimeGP
from .kern import RBF, Add, Prod, _ScalarKern
from .myKernel import myKernel, nonDivK, nonRotK, _HelmholtzBase, Kt, SpaceTimeKern
from .myKernel2 import _HelmholtzSumKern, HelmholtzSum, divFreeK, curlFreeK
from .params import Param

GPY_JITTER = 1e-8       # GPy adds 1e-8 to the diagonal before factorising


class _Run:
    """One optimisation run (GPy keeps these in model.optimization_runs; krig.py:439-457)."""

    def __init__(self, x_opt, f_opt, status, nfev):
        self.x_opt, self.f_opt, self.status, self.funct_eval = x_opt, f_opt, status, nfev


def _grad_natural_for(k, grad4):
    """d/d(l_df, l_cf, ratio, noise) -> gradients of the kernel's own parameters, then the noise."""
    if isinstance(k, myKernel):
        g = [grad4[0], grad4[1], grad4[2]]
    elif isinstance(k, nonDivK):
        g = [grad4[0]]
    else:
        g = [grad4[1]]
    return g + [grad4[3]]


class GPRegression:
    def __init__(self, X, Y, kernel=None, noise_var=1.0, jitter=GPY_JITTER, reference_compat=None, device=None):
        X = np.asarray(X, dtype=np.float64)
        Y = np.asarray(Y, dtype=np.float64)
        if kernel is None:
            kernel = myKernel(2, [0, 1], 1.0, 1.0, 0.5)
        if not isinstance(kernel, (_HelmholtzBase, _ScalarKern, SpaceTimeKern, _HelmholtzSumKern)):
            raise TypeError("kernel must be myKernel / nonDivK / nonRotK, Kt * one of them, a sum of "
                            "myKernel2.divFreeK / curlFreeK terms, or an RBF (sum)")
        self.kern = kernel
        self.scalar = isinstance(kernel, _ScalarKern)
        self.spacetime = isinstance(kernel, SpaceTimeKern)
        self.hsum = isinstance(kernel, _HelmholtzSumKern)
        self.Y = Y.reshape(-1, 1)
        self.Gaussian_noise = Param("Gaussian_noise.variance", noise_var).constrain_positive()
        self.likelihood = self
        self.variance = self.Gaussian_noise
        self.jitter = float(jitter)
        self.optimization_runs = []
        if self.scalar:
            # scalar GP over (t, y, x): one model per velocity component (krig.py:409-412)
            self.X = self.kern._slice(X)
            if self.Y.shape[0] != self.X.shape[0]:
                raise ValueError("Y must hold one observation per row of X: shape [N,1]")
            self._gp = ScalarGP(self.X, self.Y.reshape(-1), *self.kern.rbf_params(), float(self.Gaussian_noise),
                                jitter=self.jitter, device=device)
        elif self.hsum:
            # sum of divFreeK / curlFreeK terms over (t, y, x), stacked [v; u] (krig.py:392-407)
            self.X = self.kern.points(X)
            if self.Y.shape[0] != 2 * self.X.shape[0]:
                raise ValueError("Y must stack both velocity components: shape [2N,1] (krig.py:392)")
            self._gp = HelmholtzSumGP(self.X, self.Y.reshape(-1), *self.kern.hsum_params(), float(self.Gaussian_noise),
                                      jitter=self.jitter, device=device)
        elif self.spacetime:
            # Kt(t) * Helmholtz(y, x): points (t, a, b), stacked components (scratch.py:495-510)
            self.X = self.kern.points3(X)
            if self.Y.shape[0] != 2 * self.X.shape[0]:
                raise ValueError("Y must stack both velocity components: shape [2N,1]")
            self._gp = SpaceTimeGP(self.X, self.Y.reshape(-1), *self.kern.theta5(), float(self.Gaussian_noise),
                                   jitter=self.jitter, device=device)
        else:
            if reference_compat is not None:
                self.kern.reference_compat = bool(reference_compat)
            self.X = X[:, self.kern.active_dims] if X.shape[1] != 2 else X
            if self.Y.shape[0] != 2 * self.X.shape[0]:
                raise ValueError("Y must stack both velocity components: shape [2N,1] (GP_plots.py:722)")
            self._gp = HelmholtzGP(self.X, self.Y.reshape(-1), *self.kern._theta(), float(self.Gaussian_noise),
                                   jitter=self.jitter, device=device)
        self._ll = None
        self.parameters_changed()          # GPy evaluates the likelihood in the constructor

    # ---- GPy attribute idioms ---------------------------------------------------------------
    def __setattr__(self, name, value):
        # ``model.Gaussian_noise = 1.7e-7`` sets the value of the parameter (laser_io_methods.py:497,503)
        if name == "Gaussian_noise" and "Gaussian_noise" in self.__dict__ and not isinstance(value, Param):
            self.__dict__["Gaussian_noise"].value = float(np.asarray(value).reshape(-1)[0])
            if "_gp" in self.__dict__:
                self.parameters_changed()
            return
        object.__setattr__(self, name, value)

    def __getattr__(self, name):
        # ``model.rbf.lengthscale[0]`` / ``model.myKern.ratio``: the kernel under its GPy name
        kern = self.__dict__.get("kern")
        if kern is not None and name == getattr(kern, "name", None):
            return kern
        raise AttributeError(name)

    # ---- parameters -----------------------------------------------------------------------
    @property
    def parameters(self):
        return list(self.kern.parameters) + [self.Gaussian_noise]

    @property
    def param_array(self):
        return np.array([float(p) for p in self.parameters])

    def parameter_names(self):
        if self.scalar or self.spacetime or self.hsum:
            return self.kern.parameter_names() + ["Gaussian_noise.variance"]
        return ["%s.%s" % (self.kern.name, p.name) for p in self.kern.parameters] + ["Gaussian_noise.variance"]

    def _free_params(self):
        return [p for p in self.parameters if p.constraint != "fixed"]

    def _sync(self):
        if self.scalar:
            self._gp.set_params(*self.kern.rbf_params(), float(self.Gaussian_noise))
        elif self.spacetime:
            self._gp.set_params(*self.kern.theta5(), float(self.Gaussian_noise))
        elif self.hsum:
            self._gp.set_params(*self.kern.hsum_params(), float(self.Gaussian_noise))
        else:
            self._gp.set_params(*self.kern._theta(), float(self.Gaussian_noise))

    # ---- likelihood ------------------------------------------------------------------------
    def _grad_natural(self, grad4):
        """Map d/d(l_df, l_cf, ratio, noise) onto this kernel's own parameters."""
        return _grad_natural_for(self.kern, grad4)

    def parameters_changed(self):
        self._sync()
        if self.scalar:
            try:
                self._ll, g = self._jitchol(self._gp.lml_and_grad)
                self.kern._scatter_gradient(g[:-1])
                self.Gaussian_noise.gradient = float(g[-1])
            except LinAlgError:
                self._ll = -np.inf
                for p in self.parameters:
                    p.gradient = 0.0
            return self._ll
        if self.hsum:
            try:
                self._ll, g = self._jitchol(self._gp.lml_and_grad)
                self.kern.scatter_gradient(g[:-1])
                self.Gaussian_noise.gradient = float(g[-1])
            except LinAlgError:
                self._ll = -np.inf
                for p in self.parameters:
                    p.gradient = 0.0
            return self._ll
        if self.spacetime:
            try:
                self._ll, g = self._jitchol(self._gp.lml_and_grad)
                self.kern.scatter_gradient(g[:5])
                self.Gaussian_noise.gradient = float(g[5])
            except LinAlgError:
                self._ll = -np.inf
                for p in self.parameters:
                    p.gradient = 0.0
            return self._ll
        try:
            self._ll, g = self._jitchol(lambda: self._gp.lml_and_grad(reference_compat=self.kern.reference_compat))
        except LinAlgError:
            self._ll, g = -np.inf, np.zeros(4)
        for p, gi in zip(self.parameters, self._grad_natural(g)):
            p.gradient = float(gi)
        return self._ll

    def _jitchol(self, fn):
        """GPy's jitchol (GPy/util/linalg.py): when the factorisation finds a non-positive pivot,
        retry with a diagonal jitter of 1e-6 mean(diag K) 10^k, k = 0..4, warn, and give up after
        that (the caller then reports -inf, as the optimiser expects)."""
        try:
            return fn()
        except LinAlgError:
            pass
        base = self._gp.jitter
        scale = 1e-6 * (self._gp.kss() + float(self.Gaussian_noise))
        try:
            for k in range(5):
                self._gp.jitter = base + scale * 10 ** k
                try:
                    out = fn()
                except LinAlgError:
                    continue
                import warnings
                warnings.warn("Added jitter of %.3e" % (scale * 10 ** k), RuntimeWarning)
                return out
            raise LinAlgError("not positive definite, even with jitter")
        finally:
            self._gp.jitter = base

    def log_likelihood(self):
        return float(self._ll)

    def objective_function(self):
        return -float(self._ll)

    # ---- optimisation (host driver; paramz uses scipy L-BFGS-B the same way) ---------------
    def _set_free(self, x):
        for p, xi in zip(self._free_params(), x):
            p.from_free(float(xi))

    def _objective(self, x):
        self._set_free(x)
        ll = self.parameters_changed()
        if not np.isfinite(ll):
            return 1e100, np.zeros(len(x))
        g = np.array([-p.gradient * p.dvalue_dfree(float(xi)) for p, xi in zip(self._free_params(), x)])
        return -ll, g

    def optimize(self, optimizer=None, max_iters=1000, messages=False, start=None, **kw):
        x0 = np.array([p.to_free() for p in self._free_params()]) if start is None else np.asarray(start, float)
        res = sopt.minimize(self._objective, x0, jac=True, method="L-BFGS-B", options={"maxiter": int(max_iters)})
        self._set_free(res.x)
        self.parameters_changed()
        run = _Run(res.x.copy(), float(res.fun), res.message, int(res.nfev))
        self.optimization_runs.append(run)
        if messages:
            print("optimize: f=%.6f nfev=%d %s" % (res.fun, res.nfev, res.message))
        return run

    def randomize(self, rng=None):
        """Draw the unconstrained parameters from N(0,1), as GPy's model.randomize()."""
        rng = np.random.default_rng() if rng is None else rng
        self._set_free(rng.standard_normal(len(self._free_params())))
        self._sync()

    def _clone(self):
        """Independent model on the same data and device (own fit workspace), for concurrent restarts."""
        k = self.kern.copy()
        if self.hsum:                       # self.X is already sliced to the kernel's columns
            for t in k.terms_list() + [k]:
                t.active_dims = list(range(t.input_dim))
        m = GPRegression(self.X, self.Y, k, noise_var=float(self.Gaussian_noise), jitter=self.jitter,
                         device=self._gp.device)
        for a, b in zip(m.parameters, self.parameters):
            a.constraint = b.constraint
        return m

    def _one_restart(self, r, init, child, max_iters):
        if r == 0:
            self._set_free(init)
        else:
            self.randomize(np.random.default_rng(child))
        return self.optimize(max_iters=max_iters, messages=False)

    # ---- restarts in lock step: one batched objective evaluation per round --------------------------
    def _lockstep_ok(self):
        return not (self.scalar or self.spacetime or self.hsum)

    def _restarts_lockstep(self, todo, first, init, children, max_iters, nbatch, robust, report):
        """Run the restarts ``todo`` with ``nbatch`` of them in flight.  Every restart is an ordinary
        scipy L-BFGS-B run in its own host thread; its objective hands theta to the serving loop below and
        sleeps.  When every live restart has asked, ONE gp2d_lml_grad_batched call evaluates them all (each
        kernel launch of the factorisation covers the whole batch) and the answers are handed back.  A
        restart sees exactly the values the single-problem path would give it (bit-identical entry points),
        so the set of runs is the same as with ``parallel=1``; only the wall time changes."""
        import copy
        import queue
        import threading
        hb = HelmholtzBatch(self._gp.X, self._gp.y, B=nbatch, jitter=self.jitter, device=self._gp.device)
        compat = bool(self.kern.reference_compat)
        cv = threading.Condition()
        pending, answers = {}, {}
        live = [nbatch]
        work = queue.SimpleQueue()
        for r in todo:
            work.put(r)
        results, errors = {}, []
        fallback_lock = threading.Lock()
        stats = {"rounds": 0, "evaluations": 0}

        def worker(slot):
            kern = copy.deepcopy(self.kern)
            noise = copy.deepcopy(self.Gaussian_noise)
            params = list(kern.parameters) + [noise]
            free = [p for p in params if p.constraint != "fixed"]

            def objective(x):
                for p, xi in zip(free, x):
                    p.from_free(float(xi))
                theta = tuple(kern._theta()) + (float(noise),)
                with cv:
                    pending[slot] = theta
                    cv.notify_all()
                    while slot not in answers:
                        cv.wait()
                    ll, g4, info = answers.pop(slot)
                if info > 0:
                    # not positive definite: GPy's jitchol retries on the single-problem path
                    with fallback_lock:
                        f, g = self._objective(x)
                    return f, g
                if not np.isfinite(ll):
                    return 1e100, np.zeros(len(x))
                for p, gi in zip(params, _grad_natural_for(kern, g4)):
                    p.gradient = float(gi)
                return -ll, np.array([-p.gradient * p.dvalue_dfree(float(xi)) for p, xi in zip(free, x)])

            try:
                while True:
                    try:
                        r = work.get_nowait()
                    except queue.Empty:
                        break
                    try:
                        rr = r if first else max(r, 1)       # only the very first restart starts from the current point
                        # the start point takes the same round trip through the constrained values as in
                        # _one_restart (set, then read back in optimize), so the runs are the same bit for bit
                        z = init if rr == 0 else np.random.default_rng(children[r]).standard_normal(len(free))
                        for p, zi in zip(free, z):
                            p.from_free(float(zi))
                        x0 = np.array([p.to_free() for p in free])
                        res = sopt.minimize(objective, x0, jac=True, method="L-BFGS-B", options={"maxiter": int(max_iters)})
                        results[r] = _Run(res.x.copy(), float(res.fun), res.message, int(res.nfev))
                        report(r, results[r])
                    except Exception as e:                  # noqa: BLE001
                        if not robust:
                            errors.append(e)
                            break
            finally:
                with cv:
                    live[0] -= 1
                    cv.notify_all()

        threads = [threading.Thread(target=worker, args=(s,), daemon=True) for s in range(nbatch)]
        for t in threads:
            t.start()
        while True:
            with cv:
                while live[0] > 0 and len(pending) < live[0]:
                    cv.wait()
                if live[0] == 0:
                    break
                batch = sorted(pending.items())
                pending.clear()
            try:
                ll, g, info = hb.lml_and_grad([th for _, th in batch], nb=len(batch), reference_compat=compat)
                out = {slot: (float(ll[i]), g[i], int(info[i])) for i, (slot, _) in enumerate(batch)}
            except Exception as e:                          # noqa: BLE001  hand the failure to every waiting restart
                errors.append(e)
                out = {slot: (-np.inf, np.zeros(4), 0) for slot, _ in batch}
            stats["rounds"] += 1
            stats["evaluations"] += len(batch)
            with cv:
                answers.update(out)
                cv.notify_all()
        for t in threads:
            t.join()
        self.lockstep_stats = stats
        if errors:
            raise errors[0]
        return results

    def optimize_restarts(self, num_restarts=10, robust=False, verbose=True, messages=False, max_iters=1000,
                          seed=None, rank=0, world=1, parallel=1, batched=None, **kw):
        """GPy semantics: optimise from the current point, then from random points; keep the
        best.  ``rank``/``world`` shard the restart indices (restart r runs on rank r % world);
        use gp2d_b200.dist.gather_best to pick the global winner.  ``parallel`` > 1 runs that many
        restarts of this rank concurrently, each on its own CUDA stream and workspace: one
        small-n factorisation is a latency-bound chain of kernels that leaves most SMs idle, so
        independent restarts overlap almost for free.  ``batched`` (default: on for the Helmholtz kernels
        when this rank has at least two restarts) advances up to ``batched`` restarts (True: all of this
        rank's, at most 64) in lock step through gp2d_lml_grad_batched instead -- one chain of launches per
        round for all of them, see _restarts_lockstep.  Restart r always starts from the point
        drawn from seed-sequence child r, so the set of runs does not depend on the sharding or the mode."""
        import torch
        base = np.random.SeedSequence(seed)
        children = base.spawn(int(num_restarts))
        init = np.array([p.to_free() for p in self._free_params()])
        first = not self.optimization_runs
        mine = [r for r in range(int(num_restarts)) if r % world == rank]
        results = {}

        def work(model, todo):
            for r in todo:
                try:
                    rr = r if first else max(r, 1)           # only the very first restart starts from the current point
                    results[r] = model._one_restart(rr, init, children[r], max_iters)
                    if messages or verbose:
                        print("Optimization restart %d/%d, f = %s" % (r + 1, num_restarts, results[r].f_opt))
                except Exception:
                    if not robust:
                        raise

        if batched is None:
            batched = self._lockstep_ok() and len(mine) >= 2 and parallel == 1
        if batched and not self._lockstep_ok():
            raise ValueError("batched restarts are implemented for the Helmholtz kernels (myKernel, nonDivK, nonRotK)")
        parallel = max(1, min(int(parallel), len(mine)))
        if batched and mine:
            nbatch = len(mine) if batched is True else int(batched)
            nbatch = max(1, min(nbatch, len(mine), 64))
            free_b, _ = torch.cuda.mem_get_info(self._gp.device)
            nbatch = max(1, min(nbatch, int(0.8 * free_b) // max(1, self._gp.ws_bytes)))

            def report(r, run):
                if messages or verbose:
                    print("Optimization restart %d/%d, f = %s" % (r + 1, num_restarts, run.f_opt))
            results.update(self._restarts_lockstep(mine, first, init, children, max_iters, nbatch, robust, report))
        elif parallel == 1:
            keep = len(self.optimization_runs)
            work(self, mine)
            del self.optimization_runs[keep:]
        else:
            import queue
            import threading
            workers = [self] + [self._clone() for _ in range(parallel - 1)]
            errors = []
            pending = queue.SimpleQueue()
            for r in mine:
                pending.put(r)

            def pull():                   # restarts differ a lot in length: hand them out one by one
                while True:
                    try:
                        yield pending.get_nowait()
                    except queue.Empty:
                        return

            def run(model, todo):
                try:
                    with torch.cuda.device(self._gp.device), torch.cuda.stream(torch.cuda.Stream(self._gp.device)):
                        keep = len(model.optimization_runs)
                        work(model, todo)
                        del model.optimization_runs[keep:]
                        torch.cuda.current_stream().synchronize()
                except Exception as e:          # re-raised in the caller's thread
                    errors.append(e)
            torch.cuda.current_stream().synchronize()
            threads = [threading.Thread(target=run, args=(w, pull())) for w in workers]
            for t in threads:
                t.start()
            for t in threads:
                t.join()
            if errors:
                raise errors[0]
        self.optimization_runs += [results[r] for r in sorted(results)]
        if self.optimization_runs:
            best = min(self.optimization_runs, key=lambda o: o.f_opt)
            self._set_free(best.x_opt)
            self.parameters_changed()
        return self.optimization_runs

    # ---- prediction ------------------------------------------------------------------------
    def predict(self, Xnew, full_cov=False, include_likelihood=True, refined=None):
        """GPy's model.predict: ([M or 2M, 1] mean, variance), Gaussian noise included.  ``refined``:
        None picks the iterated-solve path when the covariance is ill-conditioned
        (n k** / (noise + jitter) > 1e7, engine.refined_predict; DESIGN.md §7 "Conditioning") and the
        fused kernel otherwise; True / False force one of them."""
        if full_cov:
            raise NotImplementedError("only marginal variances (the reference never asks for full_cov)")
        Xnew = np.asarray(Xnew, dtype=np.float64)
        if self.scalar:
            Xnew = self.kern._slice(Xnew)
        elif self.spacetime:
            Xnew = self.kern.points3(Xnew)
        elif self.hsum:
            Xnew = self.kern.points(Xnew)
        elif Xnew.shape[1] != 2:
            Xnew = Xnew[:, self.kern.active_dims]
        if not self._gp.fitted:
            self._sync()
            self._gp.fit()
        if refined is None:
            refined = self._gp.cond_bound() > ROBUST_COND
        if refined:
            self._sync()
            mean, var = self._gp.predict_refined(Xnew, include_noise=include_likelihood)
        else:
            mean, var = self._gp.predict(Xnew, include_noise=include_likelihood)
        return mean.cpu().numpy()[:, None], var.cpu().numpy()[:, None]

    # ---- persistence (krig.py:412,438,452) -----------------------------------------------------
    def _state(self):
        k = self.kern
        st = {"kernel": type(k).__name__, "active_dims": getattr(k, "active_dims", None), "params": self.param_array,
              "constraints": [p.constraint for p in self.parameters],
              "X": self.X, "Y": self.Y, "jitter": self.jitter,
              "runs": [r.__dict__ for r in self.optimization_runs]}
        if self.scalar and isinstance(k, Prod):
            # X is stored already sliced to k.active_dims: factor columns become positions in that list
            st["factors"] = [{"input_dim": f.input_dim, "ARD": f.ARD, "name": f.name,
                              "dims": [k.active_dims.index(d) for d in f.active_dims]} for f in k.factors]
            st["active_dims"] = list(range(self.X.shape[1]))
        elif self.scalar:
            st["parts"] = [{"input_dim": p.input_dim, "ARD": p.ARD, "name": p.name} for p in k.parts_list()]
            st["active_dims"] = list(range(self.X.shape[1]))      # X is stored already sliced
        elif self.spacetime:
            st["space"] = type(k.kxy).__name__                    # X is stored as (t, a, b)
            st["active_dims"] = [0, 1, 2]
        elif self.hsum:
            st["terms"] = [t.TYPE for t in k.terms_list()]        # X is stored already sliced
            st["input_dim"] = k.terms_list()[0].input_dim
            st["active_dims"] = list(range(self.X.shape[1]))
        else:
            st["reference_compat"] = k.reference_compat
        return st

    def pickle(self, path):
        with open(path, "wb") as f:
            pickle.dump(self._state(), f, protocol=2)

    def __str__(self):
        rows = ["  %-28s %g" % (n, v) for n, v in zip(self.parameter_names(), self.param_array)]
        return "GPRegression  log-likelihood %s\n%s" % (self._ll, "\n".join(rows))


def _restore(m, st):
    """Constraints and optimisation history of a pickled model onto the rebuilt one."""
    for prm, c in zip(m.parameters, st["constraints"]):
        prm.constraint = c
    for d in st["runs"]:
        r = _Run(None, None, None, None)
        r.__dict__.update(d)
        m.optimization_runs.append(r)
    return m


class _StateUnpickler(pickle.Unpickler):
    """Model pickles written by GPRegression.pickle are dicts of numpy arrays, lists, strings and numbers:
    only numpy's array / dtype / scalar reconstructors are allowed as globals."""

    def find_class(self, module, name):
        if module.split(".")[0] == "numpy" and name in ("_reconstruct", "ndarray", "dtype", "scalar", "_frombuffer"):
            import importlib
            try:
                return getattr(importlib.import_module(module), name)
            except (ImportError, AttributeError):
                import numpy._core.multiarray as ma
                return getattr(ma, name) if hasattr(ma, name) else getattr(np, name)
        if (module, name) == ("_codecs", "encode"):          # how protocol 2 writes the bytes of an array under py3
            import codecs
            return codecs.encode
        raise pickle.UnpicklingError("refusing to load %s.%s from a model pickle" % (module, name))


def load(path, device=None):
    """Counterpart of GPy.load (krig.py:438,478-482)."""
    with open(path, "rb") as f:
        st = _StateUnpickler(f).load()
    p = st["params"]
    name = st["kernel"]
    if "terms" in st:
        D, np_term = st["input_dim"], (4 if st["input_dim"] == 3 else 3)
        terms = []
        for q, ty in enumerate(st["terms"]):
            v = p[q * np_term:(q + 1) * np_term]
            kw = dict(var=v[0], lt=v[1], ly=v[2], lx=v[3]) if D == 3 else dict(var=v[0], ly=v[1], lx=v[2])
            terms.append((curlFreeK if ty else divFreeK)(input_dim=D, **kw))
        k = terms[0] if name != "HelmholtzSum" else HelmholtzSum(terms)
        m = GPRegression(st["X"], st["Y"], k, noise_var=p[-1], jitter=st["jitter"], device=device)
        return _restore(m, st)
    if "space" in st:
        kxy = {"myKernel": lambda: myKernel(2, [1, 2], p[2], p[3], p[4]), "nonDivK": lambda: nonDivK(2, [1, 2], p[2]),
               "nonRotK": lambda: nonRotK(2, [1, 2], p[2])}[st["space"]]()
        k = Kt(1, [0], p[0], p[1]) * kxy
        m = GPRegression(st["X"], st["Y"], k, noise_var=p[-1], jitter=st["jitter"], device=device)
        return _restore(m, st)
    if "factors" in st:
        fs, o = [], 0
        for d in st["factors"]:
            nl = d["input_dim"] if d["ARD"] else 1
            fs.append(RBF(d["input_dim"], p[o], p[o + 1:o + 1 + nl], ARD=d["ARD"], active_dims=d["dims"], name=d["name"]))
            o += 1 + nl
        m = GPRegression(st["X"], st["Y"], Prod(fs), noise_var=p[-1], jitter=st["jitter"], device=device)
        return _restore(m, st)
    if "parts" in st:
        parts, o = [], 0
        for d in st["parts"]:
            nl = d["input_dim"] if d["ARD"] else 1
            parts.append(RBF(d["input_dim"], p[o], p[o + 1:o + 1 + nl], ARD=d["ARD"], active_dims=st["active_dims"],
                             name=d["name"]))
            o += 1 + nl
        k = parts[0] if len(parts) == 1 else Add(parts)
        m = GPRegression(st["X"], st["Y"], k, noise_var=p[-1], jitter=st["jitter"], device=device)
        return _restore(m, st)
    if name == "myKernel":
        k = myKernel(2, st["active_dims"], p[0], p[1], p[2])
    elif name == "nonDivK":
        k = nonDivK(2, st["active_dims"], p[0])
    else:
        k = nonRotK(2, st["active_dims"], p[0])
    k.reference_compat = st.get("reference_compat", False)
    m = GPRegression(st["X"], st["Y"], k, noise_var=p[-1], jitter=st["jitter"], device=device)
    return _restore(m, st)
