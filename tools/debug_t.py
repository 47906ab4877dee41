import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as ol
rb = ol.rb
import test_gpu_parity2 as T
wl = rb.configs.dragon(640, 360, samples_per_pixel=1, max_bounces=2)
r = rb.Renderer(wl.width, wl.height, wl.tables, flags=rb.RB200_FLAG_NEE)
rng = np.random.RandomState(11)
n = 1_000_000
u = rng.normal(size=(n, 3)); u /= np.linalg.norm(u, axis=1, keepdims=True)
o = (np.array([0.0, 0.62, 0.0]) + 1.6 * u).astype(np.float32)
tgt = np.array([0.0, 0.62, 0.0]) + rng.uniform(-0.55, 0.55, (n, 3))
d = (tgt - o)
d = (d / np.linalg.norm(d, axis=1, keepdims=True) * rng.uniform(0.5, 2.0, (n, 1))).astype(np.float32)
hits = r.trace_rays(o, d, 1e4)
ok = hits["t"] > 0
tri = T.world_triangles(wl.tables, hits["instance"][ok], hits["primitive"][ok])
O, D = o[ok].astype(np.float64), d[ok].astype(np.float64)
e1, e2 = tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0]
p = np.cross(D, e2); det = (e1 * p).sum(1); s = O - tri[:, 0]
b1 = (s * p).sum(1) / det; q = np.cross(s, e1); b2 = (D * q).sum(1) / det; t = (e2 * q).sum(1) / det
rel = np.abs(hits["t"][ok] - t) / np.abs(t)
print("hits", ok.sum(), "rel quantiles 50/99/99.9/99.99/max", [float(np.percentile(rel, q)) for q in (50, 99, 99.9, 99.99, 100)])
print("count > 1e-5:", int((rel > 1e-5).sum()), " > 1e-6:", int((rel > 1e-6).sum()))
nrm = np.cross(e1, e2); nl = np.linalg.norm(nrm, axis=1)
cosang = np.abs((nrm * D).sum(1)) / (nl * np.linalg.norm(D, axis=1))
aspect = (np.maximum(np.linalg.norm(e1, axis=1), np.linalg.norm(e2, axis=1)) ** 2) / nl
w = np.argsort(-rel)[:12]
for i in w:
    print("rel %.3e t %.5f cos(ray,normal) %.2e aspect %.1f inst %d b1 %.4f b2 %.4f  du %.2e" % (rel[i], t[i], cosang[i], aspect[i], hits["instance"][ok][i], b1[i], b2[i], abs(hits["u"][ok][i]-b1[i])))
# error relative to the bound 2^-20 * t / cos
print("max rel*cos:", float((rel * cosang).max()), " 99.99%:", float(np.percentile(rel * cosang, 99.99)))
db = np.maximum(np.abs(hits["u"][ok] - b1), np.abs(hits["v"][ok] - b2))
print("db quantiles 99/99.9/99.99/max", [float(np.percentile(db, q)) for q in (99, 99.9, 99.99, 100)])
inside = (b1 > -1e-4) & (b2 > -1e-4) & ((b1 + b2) < 1 + 1e-4)
print("inside frac", inside.mean(), "min b1", b1.min(), "min b2", b2.min(), "max sum", (b1 + b2).max())
