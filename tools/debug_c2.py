import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as ol
rb = ol.rb
W, H = 960, 540
for nee in (1, 0):
    for mb in (1, 2, 3, 4, 6, 16):
        wl = rb.configs.bunny(W, H, levels=6, samples_per_pixel=1, max_bounces=mb)
        flags = rb.RB200_FLAG_NEE if nee else 0
        r = rb.Renderer(W, H, wl.tables, flags=flags)
        sc = ol.OracleScene(wl.tables)
        pc = wl.push_constants(0)
        r.render_batch(pc)
        o, cnt = sc.render_batch(W, H, flags, pc)
        g = r.read_hdr()
        last, _ = r.stats()
        diff = (g.view(np.uint32) != o.view(np.uint32)).any(axis=2)
        print("nee", nee, "bounces", mb, "differing pixels", int(diff.sum()), "counters", (last["extendRays"], last["shadowRays"]), (cnt["extendRays"], cnt["shadowRays"]), flush=True)
        if diff.any():
            ys, xs = np.nonzero(diff)
            for k in range(min(4, len(ys))):
                print("   px", xs[k], ys[k], g[ys[k], xs[k]], o[ys[k], xs[k]])
        r.close(); sc.close()
