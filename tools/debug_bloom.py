import os, sys
import numpy as np
ROOT = "/root/repo"
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as ol
rb = ol.rb
W, H = 1920, 1080
rng = np.random.RandomState(4)
img = np.zeros((H, W, 4), np.float32)
img[..., :3] = rng.uniform(0.0, 0.3, (H, W, 3)).astype(np.float32)
img[..., 3] = 1.0
for _ in range(40):
    y, x = rng.randint(0, H), rng.randint(0, W)
    img[y:y + rng.randint(1, 6), x:x + rng.randint(1, 6), :3] = rng.uniform(1.0, 30.0, 3).astype(np.float32)
cases = {"base": lambda a: None, "nan": lambda a: a.__setitem__((5, 7, slice(0, 3)), np.nan), "inf": lambda a: a.__setitem__((700, 1500, slice(0, 3)), np.inf),
         "negzero": lambda a: a.__setitem__((300, 300, slice(0, 3)), -0.0), "corner": lambda a: a.__setitem__((0, 0, slice(0, 3)), 50.0)}
wl = rb.configs.small_mixed(32, 24)
r = rb.Renderer(W, H, wl.tables, flags=0)
for name, f in cases.items():
    a = img.copy(); f(a)
    bloom = rb.BloomPushConsts(5.0, 1.0, 0.3)
    r.write_hdr(a); r.postprocess(bloom=bloom)
    got = r.read_ldr(); want = ol.postprocess(a, bloom=bloom)
    d = (got != want).any(axis=2)
    print(name, "differing pixels", int(d.sum()), (np.nonzero(d)[0][:3], np.nonzero(d)[1][:3]) if d.any() else "")
    if d.any():
        y, x = np.nonzero(d)[0][0], np.nonzero(d)[1][0]
        print("   ", got[y, x], want[y, x])
