"""CPU study (oracle only, no GPU): how well do quantities known at the equality-only optimum x0 forecast the active-set
iteration count of an env?    python tools/iter_forecast_study.py"""
import sys, os, numpy as np
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), 'tests'))
from common import setup
from tsid_control_b200 import synth
s = setup("v1")
orc = s["oracle"]
n = 600
q, v = synth.random_states(s["q0"], n, 0)
mask, refs = synth.walking_batch(s["refs"], n, 0, 0.3, 0.2, 0.2, 0.5, float(s["refs"]["com"][2]))
rows = []
for e in range(n):
    m = int(mask[e])
    r = orc.tick(q[e], v[e], m, {k: a[e] for k, a in refs.items()}, dump=True)
    d = r["dump"]
    H, g, CE, ce0, CI, ci0 = d["H"], d["g"], d["CE"], d["ce0"], d["CI"], d["ci0"]
    nn, me = H.shape[0], CE.shape[0]
    K = np.block([[H, CE.T], [CE, np.zeros((me, me))]])
    x0 = np.linalg.solve(K, np.r_[-g, -ce0])[:nn]
    nv = 26
    nc = (m & 1) + (m >> 1)
    f = x0[nv:].reshape(nc * 4, 3)
    fz_neg = int((f[:, 2] < 0).sum())
    # friction cone violation per corner: |fx|,|fy| > mu fz (mu = conf.mu)
    mu = s["conf"].mu if hasattr(s["conf"], "mu") else 0.5
    cone = int(((np.abs(f[:, 0]) > mu * f[:, 2]) | (np.abs(f[:, 1]) > mu * f[:, 2])).sum())
    sv = CI @ x0 + ci0
    viol = sv < -1e-9
    rows.append((m, r["iters"], fz_neg, cone, int(viol.sum()), float(-sv[viol].sum())))
rows = np.array(rows)
for m in (3, 1, 2):
    sel = rows[:, 0] == m
    it = rows[sel, 1]
    for name, col in (("fz_neg", 2), ("cone", 3), ("nviol", 4)):
        x = rows[sel, col]
        print("mask", m, name, "corr %.3f" % np.corrcoef(it, x)[0, 1], end=" | ")
    X = np.c_[np.ones(sel.sum()), rows[sel, 2], rows[sel, 3], rows[sel, 4]]
    beta, *_ = np.linalg.lstsq(X, it, rcond=None)
    pred = X @ beta
    print("LS fit R2 %.3f" % (1 - ((it - pred) ** 2).sum() / ((it - it.mean()) ** 2).sum()), "beta", np.round(beta, 2))
    k = max(1, int(0.1 * sel.sum()))
    top = set(np.argsort(-it, kind="stable")[:k])
    for name, p in (("cone", rows[sel, 3]), ("LS", pred)):
        pt = set(np.argsort(-p, kind="stable")[:2 * k])
        print("   top-10%% caught in top-20%% by %s: %.2f" % (name, len(top & pt) / k))
