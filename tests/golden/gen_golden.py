#!/usr/bin/env python
"""Generate golden vectors by EXECUTING THE REFERENCE'S OWN problem definition.

Runs here (authoring container) only: needs /root/reference.  The reference
defines the NMPC symbolically with CasADi and hands it to OpEn
(src/pkg_mpc_tracker/solver_build/mpc_builder.py:28-205).  Neither casadi nor
opengen is installed, so this script installs two tiny stand-ins into
``sys.modules`` BEFORE importing the reference:

* ``casadi.casadi``: an ``SX`` class that is a 2-D float64 torch tensor with
  CasADi's matrix semantics (column vectors from ``sym``, column-major
  ``reshape``, linear single-index, scalar and horizontal-repmat broadcasting,
  ``fmax/fmin`` with the 0.5/0.5 tie derivative, ``mmin`` as a fold of ``fmin``).
  The reference code then runs *numerically* at a concrete ``(u, p)`` and
  ``torch.autograd`` provides the gradient CasADi's AD would generate.
* ``opengen``: records what ``MpcModule.build`` passes to ``og.builder.Problem``
  / ``with_constraints`` / ``with_aug_lagrangian_constraints`` /
  ``with_penalty_constraints`` and forms psi exactly as opengen's builder does
  (psi = f + c/2 [dist2_C(F1 + y/max(c,1)) + F2.F2]).

The unmodified reference files executed are mpc_builder.py, mpc_cost.py,
mpc_helper.py, motion_model.py (unicycle_model) and configs.py.
Output: tests/golden/psi_cases.npz  (committed; the tests do not need the
reference).
Usage:  python tests/golden/gen_golden.py                re-run the reference on the COMMITTED inputs and
                                                         print the largest difference from the committed
                                                         outputs (0 expected; --write stores them)
        python tests/golden/gen_golden.py --new-inputs   draw new inputs with the current generator
"""
import os
import sys
import tempfile
import types

import numpy as np
import torch
import yaml

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

torch.set_default_dtype(torch.float64)


# --------------------------------------------------------------------------- casadi stand-in
def _t(x):
    if isinstance(x, SX):
        return x.t
    if isinstance(x, torch.Tensor):
        return x.reshape(1, 1) if x.dim() == 0 else x
    if isinstance(x, (int, float, np.floating, np.integer)):
        return torch.tensor([[float(x)]])
    if isinstance(x, (list, tuple, np.ndarray)):
        a = torch.as_tensor(np.asarray(x, dtype=np.float64))
        if a.dim() == 1:
            a = a.reshape(-1, 1)
        return a
    raise TypeError(type(x))


def _bcast(a, b):
    """CasADi binary-op shape rule: equal, scalar, or same rows with horizontal repmat."""
    if a.shape == b.shape:
        return a, b
    if a.numel() == 1:
        return a.expand(b.shape), b
    if b.numel() == 1:
        return a, b.expand(a.shape)
    if a.shape[0] == b.shape[0] and b.shape[1] % a.shape[1] == 0:
        return a.repeat(1, b.shape[1] // a.shape[1]), b
    if a.shape[0] == b.shape[0] and a.shape[1] % b.shape[1] == 0:
        return a, b.repeat(1, a.shape[1] // b.shape[1])
    raise ValueError(f"dimension mismatch {tuple(a.shape)} vs {tuple(b.shape)}")


class SX:
    def __init__(self, x=0.0):
        self.t = _t(x)

    # construction
    @staticmethod
    def sym(name, n, m=1):
        raise RuntimeError("symbols are injected by the harness (see SymbolFeed)")

    @staticmethod
    def ones(n, m=1):
        return SX(torch.ones(n, m))

    @property
    def shape(self):
        return tuple(self.t.shape)

    @property
    def T(self):
        return SX(self.t.transpose(0, 1))

    def _lin(self, idx):
        flat = self.t.transpose(0, 1).reshape(-1)       # column-major nonzeros
        if isinstance(idx, int):
            return SX(flat[idx].reshape(1, 1))
        sel = flat[idx]
        if self.t.shape[0] == 1:                        # row vector stays a row
            return SX(sel.reshape(1, -1))
        return SX(sel.reshape(-1, 1))

    def __getitem__(self, idx):
        if isinstance(idx, tuple):
            r, c = idx
            r = [r] if isinstance(r, int) else r
            c = [c] if isinstance(c, int) else c
            out = self.t[r, :] if not isinstance(r, list) else self.t[torch.as_tensor(r), :]
            out = out[:, c] if not isinstance(c, list) else out[:, torch.as_tensor(c)]
            return SX(out)
        if isinstance(idx, list):
            idx = torch.as_tensor(idx)
        return self._lin(idx)

    # arithmetic
    def _bin(self, other, fn, rev=False):
        a, b = _t(self), _t(other)
        if rev:
            a, b = b, a
        a, b = _bcast(a, b)
        return SX(fn(a, b))

    def __add__(self, o): return self._bin(o, torch.add)
    def __radd__(self, o): return self._bin(o, torch.add, True)
    def __sub__(self, o): return self._bin(o, torch.sub)
    def __rsub__(self, o): return self._bin(o, torch.sub, True)
    def __mul__(self, o): return self._bin(o, torch.mul)
    def __rmul__(self, o): return self._bin(o, torch.mul, True)
    def __truediv__(self, o): return self._bin(o, torch.div)
    def __rtruediv__(self, o): return self._bin(o, torch.div, True)
    def __neg__(self): return SX(-self.t)

    def __pow__(self, k):
        if k == 2:
            return SX(self.t * self.t)                  # CasADi: pow(x,2) -> sq(x)
        return SX(self.t ** k)


def _cat(args, dim):
    return SX(torch.cat([_t(a) for a in args], dim=dim))


def vertcat(*args): return _cat(args, 0)
def horzcat(*args): return _cat(args, 1)
def vcat(lst): return _cat(lst, 0)
def hcat(lst): return _cat(lst, 1)
def transpose(x): return SX(_t(x).transpose(0, 1))


def reshape(x, shp):
    t = _t(x)
    flat = t.transpose(0, 1).reshape(-1)                # column-major
    return SX(flat.reshape(shp[1], shp[0]).transpose(0, 1))


def fmax(a, b):
    a, b = _bcast(_t(a), _t(b))
    return SX(torch.maximum(a, b))                      # ties: gradient split 0.5/0.5


def fmin(a, b):
    a, b = _bcast(_t(a), _t(b))
    return SX(torch.minimum(a, b))


def cos(x): return SX(torch.cos(_t(x)))
def sin(x): return SX(torch.sin(_t(x)))
def sqrt(x): return SX(torch.sqrt(_t(x)))


def _seqsum(t, dim):
    out = None
    for sl in t.unbind(dim):
        out = sl if out is None else out + sl
    return out


def sum1(x): return SX(_seqsum(_t(x), 0).reshape(1, -1))
def sum2(x): return SX(_seqsum(_t(x), 1).reshape(-1, 1))


def dot(a, b):
    a, b = _t(a).reshape(-1), _t(b).reshape(-1)
    return SX(_seqsum(a * b, 0).reshape(1, 1))


def norm_2(x):
    v = _t(x).reshape(-1)
    return SX(torch.sqrt(_seqsum(v * v, 0)).reshape(1, 1))


def mtimes(a, b):
    a, b = _t(a), _t(b)
    prod = a[:, :, None] * b[None, :, :]                # [i,k,j], summed over k in order
    return SX(_seqsum(prod, 1))


def mmin(x):
    v = _t(x).transpose(0, 1).reshape(-1)
    r = torch.tensor(float("inf"))
    for e in v:
        r = torch.minimum(r, e)
    return SX(r.reshape(1, 1))


def install_standins(feed):
    cs = types.ModuleType("casadi.casadi")
    for k, v in dict(SX=SX, vertcat=vertcat, horzcat=horzcat, vcat=vcat, hcat=hcat,
                     transpose=transpose, reshape=reshape, fmax=fmax, fmin=fmin, cos=cos,
                     sin=sin, sqrt=sqrt, sum1=sum1, sum2=sum2, dot=dot, norm_2=norm_2,
                     mtimes=mtimes, mmin=mmin, pi=np.pi).items():
        setattr(cs, k, v)
    SX.sym = staticmethod(feed.sym)
    pkg = types.ModuleType("casadi")
    pkg.casadi = cs
    sys.modules["casadi"] = pkg
    sys.modules["casadi.casadi"] = cs

    # ---- opengen stand-in: just records the problem
    og = types.ModuleType("opengen.opengen")

    class Rectangle:
        def __init__(self, xmin, xmax):
            self.xmin, self.xmax = list(xmin), list(xmax)

        def distance_squared(self, z):
            # opengen.constraints.Rectangle.distance_squared (both bounds finite)
            d = 0.0
            for i in range(len(self.xmin)):
                d = d + fmax(0.0, fmax(z[i] - self.xmax[i], self.xmin[i] - z[i])) ** 2
            return d

    class Problem:
        last = None

        def __init__(self, u, p, cost):
            self.u, self.p, self.cost = u, p, cost
            Problem.last = self

        def with_constraints(self, c):
            self.constraints = c
            return self

        def with_aug_lagrangian_constraints(self, f1, set_c, set_y=None):
            self.f1, self.set_c = f1, set_c
            return self

        def with_penalty_constraints(self, f2):
            self.f2 = f2
            return self

    class _Chain:
        def __init__(self, *a, **k): pass
        def __getattr__(self, name):
            return lambda *a, **k: self

    og.constraints = types.SimpleNamespace(Rectangle=Rectangle)
    og.builder = types.SimpleNamespace(Problem=Problem, OpEnOptimizerBuilder=_Chain)
    og.config = types.SimpleNamespace(BuildConfiguration=_Chain, OptimizerMeta=_Chain,
                                      SolverConfiguration=_Chain)
    ogpkg = types.ModuleType("opengen")
    ogpkg.opengen = og
    sys.modules["opengen"] = ogpkg
    sys.modules["opengen.opengen"] = og
    return Problem


class SymbolFeed:
    """Hands concrete values to the builder's ``cs.SX.sym`` calls, in call order."""

    def __init__(self):
        self.values = {}
        self.leaves = {}

    def set(self, **kw):
        self.values = kw
        self.leaves = {}

    def sym(self, name, n, m=1):
        v = torch.as_tensor(np.asarray(self.values[name], dtype=np.float64)).reshape(n, m).clone()
        v.requires_grad_(name == "u")
        self.leaves[name] = v
        return SX(v)


def reference_eval(builder_mod, motion_model, cfg_path, dims, feed, Problem, p, u, y, c):
    """Run MpcModule.build(test=True) of the reference at (u, p); return goldens."""
    from configs import MpcConfiguration, CircularRobotSpecification
    lay = dims.layout()
    names = {"u_m1": "u_m1", "s_0": "s_0", "s_N": "s_N", "q": "q", "r_s": "r_s", "r_v": "r_v",
             "c_0": "c_0", "c": "c", "o_s": "os", "o_d": "od", "q_stc": "qstc", "q_dyn": "qdyn"}
    vals = {names[k]: p[o:o + ln] for k, (o, ln) in lay.items()}
    vals["u"] = u
    feed.set(**vals)
    mod = builder_mod.MpcModule(MpcConfiguration.from_yaml(cfg_path),
                                CircularRobotSpecification.from_yaml(cfg_path))
    assert mod.build(motion_model.unicycle_model, test=True) == 1
    prob = Problem.last
    f = prob.cost.t.reshape(())
    F1 = prob.f1.t.reshape(-1)
    F2 = prob.f2.t.reshape(-1)
    # opengen builder __construct_function_psi
    xi0 = torch.tensor(float(c))
    yv = torch.as_tensor(np.asarray(y, dtype=np.float64))
    z = SX((F1 + yv / torch.maximum(xi0, torch.tensor(1.0))).reshape(-1, 1))
    psi = f + xi0 * prob.set_c.distance_squared(z).t.reshape(()) / 2 + xi0 * (F2 * F2).sum() / 2
    (g,) = torch.autograd.grad(psi, feed.leaves["u"])
    return dict(f=f.item(), psi=psi.item(), grad=g.reshape(-1).numpy().copy(),
                F1=F1.detach().numpy().copy(), F2=F2.detach().numpy().copy(),
                umin=np.asarray(prob.constraints.xmin), umax=np.asarray(prob.constraints.xmax),
                cmin=np.asarray(prob.set_c.xmin), cmax=np.asarray(prob.set_c.xmax))


def reevaluate_committed(mpc_builder, motion_model, base_cfg, feed, Problem, write):
    """The committed INPUTS (p, u, y, c, dims of every case in psi_cases.npz) are the source of truth:
    run the reference on them again and compare with (or rewrite) the committed outputs.  This is
    what makes the fixture reproducible whatever happens to the synthetic instance generator."""
    from dyobav_mpcnwta_warehouse_b200 import Dims
    path = os.path.join(HERE, "psi_cases.npz")
    old = dict(np.load(path))
    worst = 0.0
    for key in [str(k) for k in old["cases"]]:
        dims = Dims(*[int(v) for v in old[f"{key}/dims"]])
        cfgd = dict(base_cfg)
        cfgd.update(N_hor=dims.N, Nother=dims.Nother, Nstcobs=dims.Nstc, nstcobs=3 * dims.nedge, Ndynobs=dims.Ndyn)
        with tempfile.NamedTemporaryFile("w", suffix=".yaml", delete=False) as fh:
            yaml.safe_dump(cfgd, fh)
            cfg_path = fh.name
        g = reference_eval(mpc_builder, motion_model, cfg_path, dims, feed, Problem, old[f"{key}/p"],
                           old[f"{key}/u"], old[f"{key}/y"], float(old[f"{key}/c"]))
        os.unlink(cfg_path)
        for k, v in g.items():
            d = float(np.max(np.abs(np.asarray(v) - old[f"{key}/{k}"]))) if np.size(v) else 0.0
            worst = max(worst, d)
            if write:
                old[f"{key}/{k}"] = np.asarray(v)
        print(key, "max |reference - committed| over its outputs so far: %.3g" % worst)
    if write:
        np.savez_compressed(path, **old)
        print("rewrote", path)
    print("MAXDIFF %.17g" % worst)
    return worst


def main():
    from dyobav_mpcnwta_warehouse_b200 import Dims, instances

    feed = SymbolFeed()
    Problem = install_standins(feed)
    sys.path.insert(0, os.path.join(REF, "src"))
    from pkg_mpc_tracker.solver_build import mpc_builder
    from basic_motion_model import motion_model

    with open(os.path.join(REF, "config", "mpc_fast.yaml")) as fh:
        base_cfg = yaml.safe_load(fh)

    if "--new-inputs" not in sys.argv:
        # default: the reference re-evaluated on the committed inputs (--write stores the outputs)
        reevaluate_committed(mpc_builder, motion_model, base_cfg, feed, Problem, "--write" in sys.argv)
        return
    # --new-inputs: draw fresh inputs with the CURRENT instance generator (changes the fixture)

    cases = []
    rng = np.random.default_rng(20231017)
    dim_sets = [
        ("default", Dims(), 4),
        ("small", Dims(N=5, Nother=2, Nstc=2, nedge=4, Ndyn=3), 3),
        ("ndyn40", Dims(Ndyn=40), 2),
        ("n40", Dims(N=40, Nother=3, Nstc=4, nedge=4, Ndyn=12), 2),
    ]
    out = {}
    for name, dims, count in dim_sets:
        cfgd = dict(base_cfg)
        cfgd.update(N_hor=dims.N, Nother=dims.Nother, Nstcobs=dims.Nstc,
                    nstcobs=3 * dims.nedge, Ndynobs=dims.Ndyn)
        # make the dead weights live so every term is exercised (rv, rw, qN, qthetaN)
        with tempfile.NamedTemporaryFile("w", suffix=".yaml", delete=False) as fh:
            yaml.safe_dump(cfgd, fh)
            cfg_path = fh.name
        P = instances.generate(dims, count, seed=100 + len(out), pedestrians=2,
                               modes=max(1, min(20, dims.Ndyn // 2)))
        lay = dims.layout()
        for i in range(count):
            p = P[i].copy()
            o, ln = lay["q"]
            p[o:o + ln] = [0.3, 10.0, 0.2, 0.7, 0.4, 1.5, 2.5, 100.0, 10.0, 20.0]
            # other robots: put a few near the robot so the fleet hinges are active
            s0 = p[lay["s_0"][0]:lay["s_0"][0] + 3]
            o, ln = lay["c_0"]
            p[o:o + ln] = np.tile(s0, dims.Nother) + rng.normal(0, 0.3, size=ln)
            o, ln = lay["c"]
            p[o:o + ln] = np.tile(s0, dims.N * dims.Nother) + rng.normal(0, 0.8, size=ln)
            # rotate / squash some ellipses and drop one right on the robot's way
            o, ln = lay["o_d"]
            od = p[o:o + ln].reshape(dims.Ndyn, dims.N + 1, 6)
            od[..., 4] = rng.uniform(-3.0, 3.0, size=od[..., 4].shape) * (od[..., 5] > 0)
            od[..., 3] *= rng.uniform(0.5, 2.0, size=od[..., 3].shape)
            if dims.Ndyn:
                od[0, :, 0] = s0[0] + 0.3 * np.cos(s0[2]) * np.arange(dims.N + 1) * 0.5
                od[0, :, 1] = s0[1] + 0.3 * np.sin(s0[2]) * np.arange(dims.N + 1) * 0.5
                od[0, :, 2:4] = [0.6, 0.4]
                od[0, :, 5] = 0.8
            p[o:o + ln] = od.reshape(-1)
            # one static polygon right around the robot so the polygon terms are active
            if dims.Nstc:
                o, ln = lay["o_s"]
                ne = dims.nedge
                cx, cy = s0[0] + 0.5, s0[1] + 0.2
                a0 = np.array([1.0, -1.0, 0.0, 0.0]) / 1.5
                a1 = np.array([0.0, 0.0, 1.0, -1.0]) / 1.2
                p[o:o + 3 * ne] = np.concatenate([a0 * cx + a1 * cy + 1.0, a0, a1])
            u = np.stack([rng.uniform(-0.5, 1.5, dims.N), rng.uniform(-0.5, 0.5, dims.N)], 1).reshape(-1)
            if i == 0:
                u[:] = 0.0                                   # the reference's own start point
            y = rng.normal(0, 2.0, size=dims.n1) if i % 2 else np.zeros(dims.n1)
            c = [10.0, 50.0, 1250.0, 1.0][i % 4]
            g = reference_eval(mpc_builder, motion_model, cfg_path, dims, feed, Problem, p, u, y, c)
            key = f"{name}_{i}"
            out[f"{key}/dims"] = np.array([dims.N, dims.Nother, dims.Nstc, dims.nedge, dims.Ndyn])
            out[f"{key}/p"], out[f"{key}/u"], out[f"{key}/y"], out[f"{key}/c"] = p, u, y, np.array(c)
            for k, v in g.items():
                out[f"{key}/{k}"] = np.asarray(v)
            cases.append(key)
            print(key, "f=%.6g psi=%.6g |grad|=%.4g |F2|=%.4g" % (
                g["f"], g["psi"], np.linalg.norm(g["grad"]), np.linalg.norm(g["F2"])))
        os.unlink(cfg_path)
    out["cases"] = np.array(cases)
    np.savez_compressed(os.path.join(HERE, "psi_cases.npz"), **out)
    print("wrote", os.path.join(HERE, "psi_cases.npz"), len(cases), "cases")


if __name__ == "__main__":
    main()
