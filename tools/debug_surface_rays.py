"""Rays that start ON the surface (as every bounce ray does): CUDA traversal vs the oracle's hierarchy vs brute force."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as ol
rb = ol.rb
wl = rb.configs.bunny(64, 48, levels=5, samples_per_pixel=1, max_bounces=2)     # 2 x 20,480 triangles + showroom
t = wl.tables
print("triangles", t.num_triangles())
inst = np.frombuffer(t.instances.tobytes(), dtype=np.uint32).reshape(t.numInstances, 20)
M = inst[:, :16].view(np.float32)
rng = np.random.RandomState(3)
n = 300000
# pick the two blob instances (largest triangle counts after the showroom)
cand = list(range(t.numInstances))
ii = rng.choice(cand, n)
prim = (rng.rand(n) * inst[ii, 19]).astype(np.int64)
ib = 3 * prim + inst[ii, 18].astype(np.int64)
V = []
for k in range(3):
    v = t.vertices[t.indices[ib + k]][:, :3].astype(np.float64)
    m = M[ii].astype(np.float64)
    w = np.stack([m[:, r] * v[:, 0] + m[:, 4 + r] * v[:, 1] + m[:, 8 + r] * v[:, 2] + m[:, 12 + r] for r in range(3)], 1)
    V.append(w)
b = rng.dirichlet([1, 1, 1], n)
P = (V[0] * b[:, :1] + V[1] * b[:, 1:2] + V[2] * b[:, 2:3])
nrm = np.cross(V[1] - V[0], V[2] - V[0]); nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
off = rng.choice([-1, 1], n)[:, None] * rng.choice([0.0, 1e-7, 1e-6, 1e-5, 1e-4], n)[:, None]
o = (P + nrm * off).astype(np.float32)
d = rng.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
d = (d * rng.uniform(0.3, 3.0, (n, 1))).astype(np.float32)
r = rb.Renderer(wl.width, wl.height, t)
sc = ol.OracleScene(t)
g = r.trace_rays(o, d, 1e4)
h = sc.trace_rays(o, d, 1e4)
bf = sc.trace_rays(o, d, 1e4, brute=True)
def diff(a, b):
    return (a["instance"] != b["instance"]) | (a["primitive"] != b["primitive"]) | (a["t"].view(np.uint32) != b["t"].view(np.uint32)) | \
           (a["u"].view(np.uint32) != b["u"].view(np.uint32)) | (a["v"].view(np.uint32) != b["v"].view(np.uint32))
print("cuda vs brute:", int(diff(g, bf).sum()), " oracle-bvh vs brute:", int(diff(h, bf).sum()), " cuda vs oracle-bvh:", int(diff(g, h).sum()))
for name, a in (("cuda", g), ("oracle-bvh", h)):
    bad = np.nonzero(diff(a, bf))[0][:6]
    for i in bad:
        print(name, i, "o", o[i], "d", d[i], "got", a[i], "brute", bf[i])
tm = rng.uniform(0.001, 2.5, n).astype(np.float32)
ga = r.trace_rays(o, d, tm, any_hit=True)["t"] >= 0
ha = sc.trace_rays(o, d, tm, any_hit=True)["t"] >= 0
ba = sc.trace_rays(o, d, tm, any_hit=True, brute=True)["t"] >= 0
print("any-hit: cuda vs brute", int((ga != ba).sum()), " oracle-bvh vs brute", int((ha != ba).sum()))
