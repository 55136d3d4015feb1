"""Thompson-sampling Bayesian optimisation on top of the CUDA hot path (reference: BO.py:9-95).

Same class, constructor and methods as the reference's ``BO``.  What changes is how a sampled network is minimised
(``nn_opt``, BO.py:47-73): the reference hands platypus' NSGA-II a Python callback that evaluates ONE point per call through a
CPU ``nn.Sequential`` (5000 single-row forwards per sampled net); here every generation's whole population goes through the
fused forward kernel in one call, against the sampled network's row of the device-resident bank.  platypus is not needed: the
NSGA-II used by the reference with default operators (binary tournament, SBX eta = 15, polynomial mutation eta = 20 with
probability 1/dim, constrained non-dominated sorting + crowding distance, population 50, 5000 evaluations) is restated in
``nsga2`` below for a *batched* objective.  ``see()`` returns the curves it would plot and only draws when matplotlib is there.
"""
import numpy as np
import torch
import torch.nn as nn

from .BNN_SGDMC import BNN_SGDMC


# ---- NSGA-II for a batched objective ---------------------------------------------------------------------------------------
def _dominates(f1, c1, f2, c2):
    """Constrained domination (Deb 2002): feasible beats infeasible, less violation beats more, else Pareto on objectives."""
    if c1 != c2 and (c1 > 0 or c2 > 0):
        return c1 < c2
    return bool(np.all(f1 <= f2) and np.any(f1 < f2))


def _sort_fronts(F, cv):
    n = F.shape[0]
    S = [[] for _ in range(n)]
    cnt = np.zeros(n, dtype=np.int64)
    rank = np.zeros(n, dtype=np.int64)
    fronts = [[]]
    for p in range(n):
        for q in range(n):
            if p == q:
                continue
            if _dominates(F[p], cv[p], F[q], cv[q]):
                S[p].append(q)
            elif _dominates(F[q], cv[q], F[p], cv[p]):
                cnt[p] += 1
        if cnt[p] == 0:
            fronts[0].append(p)
    i = 0
    while fronts[i]:
        nxt = []
        for p in fronts[i]:
            for q in S[p]:
                cnt[q] -= 1
                if cnt[q] == 0:
                    rank[q] = i + 1
                    nxt.append(q)
        i += 1
        fronts.append(nxt)
    return fronts[:-1], rank


def _crowding(F, front):
    d = np.zeros(len(front))
    if len(front) <= 2:
        d[:] = np.inf
        return d
    Ff = F[front]
    for m in range(F.shape[1]):
        o = np.argsort(Ff[:, m], kind="stable")
        d[o[0]] = d[o[-1]] = np.inf
        span = Ff[o[-1], m] - Ff[o[0], m]
        if span > 0:
            d[o[1:-1]] += (Ff[o[2:], m] - Ff[o[:-2], m]) / span
    return d


def nsga2(evaluate, dim, lb, ub, nobj=1, ncons=0, population=50, max_evals=5000, rng=None):
    """Minimise a batched objective.  ``evaluate(X[P, dim]) -> Y[P, nobj + ncons]`` (numpy); constraints are "<= 0"
    (BO.py:63).  Returns (X, Y) of the final population's non-dominated feasible-first front."""
    rng = np.random.default_rng() if rng is None else rng
    X = rng.uniform(lb, ub, size=(population, dim))
    Y = np.asarray(evaluate(X), dtype=np.float64).reshape(population, nobj + ncons)
    evals = population

    def viol(Yv):
        return np.maximum(Yv[:, nobj:], 0.0).sum(axis=1) if ncons else np.zeros(Yv.shape[0])

    def rank_crowd(Yv):
        fronts, rank = _sort_fronts(Yv[:, :nobj], viol(Yv))
        crowd = np.zeros(Yv.shape[0])
        for fr in fronts:
            crowd[fr] = _crowding(Yv[:, :nobj], fr)
        return fronts, rank, crowd

    fronts, rank, crowd = rank_crowd(Y)
    while evals < max_evals:
        # binary tournament on (rank, crowding)
        a, b = rng.integers(population, size=population), rng.integers(population, size=population)
        win = np.where((rank[a] < rank[b]) | ((rank[a] == rank[b]) & (crowd[a] > crowd[b])), a, b)
        P1, P2 = X[win], X[np.roll(win, 1)]
        # simulated binary crossover (eta = 15, probability 1, per-variable 0.5)
        u = rng.random(P1.shape)
        beta = np.where(u <= 0.5, (2 * u) ** (1 / 16.0), (1 / (2 * (1 - u))) ** (1 / 16.0))
        cross = rng.random(P1.shape) < 0.5
        C = np.where(cross, 0.5 * ((1 + beta) * P1 + (1 - beta) * P2), P1)
        # polynomial mutation (eta = 20, probability 1 / dim)
        mut = rng.random(C.shape) < (1.0 / dim)
        r = rng.random(C.shape)
        delta = np.where(r < 0.5, (2 * r) ** (1 / 21.0) - 1, 1 - (2 * (1 - r)) ** (1 / 21.0))
        C = np.clip(np.where(mut, C + delta * (ub - lb), C), lb, ub)
        YC = np.asarray(evaluate(C), dtype=np.float64).reshape(population, nobj + ncons)
        evals += population
        # elitist survival: best `population` of parents + children by (front, crowding)
        XA, YA = np.concatenate([X, C]), np.concatenate([Y, YC])
        frontsA, rankA, crowdA = rank_crowd(YA)
        keep = []
        for fr in frontsA:
            if len(keep) + len(fr) <= population:
                keep += fr
            else:
                o = np.argsort(-crowdA[fr], kind="stable")
                keep += [fr[i] for i in o[: population - len(keep)]]
                break
        X, Y = XA[keep], YA[keep]
        fronts, rank, crowd = rank_crowd(Y)
    best = fronts[0]
    return X[best], Y[best]


class BO:
    def __init__(self, f, dim, nobj, ncons, max_eval=2, num_init=1, act=nn.Tanh(), num_hiddens=[50], conf={}):
        """f: objective function, with (-1, dim) tensor as input (BO.py:10-25)."""
        self.f = f
        self.dim, self.nobj, self.ncons, self.max_eval = dim, nobj, ncons, max_eval
        self.lb = -1 * torch.tensor(3.).sqrt()
        self.ub = torch.tensor(3.).sqrt()        # Uniform(-sqrt(3), sqrt(3)).variance = 1
        self.bnn = BNN_SGDMC(self.dim, act=act, num_hiddens=num_hiddens, nout=nobj + ncons, conf=conf)
        self.X = torch.rand(num_init, self.dim) * (self.ub - self.lb) + self.lb
        self.y = self.f(self.X)
        assert (self.y.dim() == 2)
        assert (self.y.shape[1] == self.nobj + self.ncons)
        self.population = conf.get('nsga_population', 50)
        self.nsga_evals = conf.get('nsga_evals', 5000)
        self.rng = np.random.default_rng(conf.get('seed', 0))

    def see(self, nn_idxs):
        """BO.py:27-38: all sampled curves (de-normalised) over [lb, ub]; returns (xs, pred[S, 100]) and plots if it can."""
        assert (self.dim == 1)
        assert (self.nobj + self.ncons == 1)
        xs = torch.linspace(float(self.lb), float(self.ub), 100).view(-1, 1)
        with torch.no_grad():
            pred = self.bnn.sample_predict(self.bnn.nns, xs) * self.y.std(dim=0) + self.y.mean(dim=0)
        try:
            import matplotlib.pyplot as plt
            plt.plot(xs.numpy(), pred.squeeze(-1).t().numpy(), 'g', alpha=0.1)
            plt.plot(xs.numpy(), pred[nn_idxs].squeeze(-1).t().numpy(), 'r')
            plt.plot(self.X.numpy(), self.y.numpy(), 'k+')
            plt.show()
        except ImportError:
            pass
        return xs, pred.squeeze(-1)

    def train(self):
        self.bnn.train(self.X, (self.y - self.y.mean(dim=0)) / self.y.std(dim=0))     # BO.py:41

    def _std(self):
        s = self.y.std(dim=0)
        return torch.where(torch.isfinite(s) & (s > 0), s, torch.ones_like(s))

    def nn_opt(self, nn_index):
        """BO.py:47-73: minimise ONE sampled network (objectives, then "<= 0" constraints) with NSGA-II; every generation is one
        batched forward of the population through that network's bank row.  Returns (suggested_x[1, dim], suggested_y[1, nout])."""
        one = self.bnn.nns[[int(nn_index)]]                    # SampleBank view holding that one network

        def evaluate(Xp):
            with torch.no_grad():
                out = self.bnn.sample_predict(one, torch.as_tensor(Xp, dtype=torch.float32))   # [1, P, nout]
            return out[0].double().numpy()

        Xb, Yb = nsga2(evaluate, self.dim, float(self.lb), float(self.ub), self.nobj, self.ncons, population=self.population,
                       max_evals=self.nsga_evals, rng=self.rng)
        k = int(self.rng.integers(len(Xb)))                     # np.random.randint(len(optimized)), BO.py:69
        sx = torch.as_tensor(Xb[k], dtype=torch.float32)
        sy = torch.as_tensor(Yb[k], dtype=torch.float32)
        return sx.view(-1, self.dim), sy.view(-1, self.nobj + self.ncons)

    def bo_iter(self, num_samples=1):
        """BO.py:75-88."""
        self.train()
        assert (num_samples <= len(self.bnn.nns))
        nn_idxs = torch.randperm(len(self.bnn.nns))[:num_samples]
        suggested = torch.zeros(num_samples, self.dim)
        if self.dim == 1 and self.nobj + self.ncons == 1:
            self.see(nn_idxs)
        for i in range(num_samples):
            sx, sy = self.nn_opt(i)                 # the reference optimises nns[i], not nns[nn_idxs[i]] (BO.py:82) -- kept
            suggested[i] = sx
        evaluated = self.f(suggested)
        self.X = torch.cat((self.X, suggested))
        self.y = torch.cat((self.y, evaluated))
        print(suggested)
        print(evaluated)

    def OSFTA(self):
        """One Sample to Find Them All (BO.py:90-95)."""
        for i in range(self.max_eval):
            self.bo_iter()
