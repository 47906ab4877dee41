"""Fold gpurun_out/configs/<c>.json (tools/run_configs.sh on the GPU box) into profiles/r02_configs.json."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = {}
for c in ("c1", "c2", "c3", "c4", "c5"):
    p = os.path.join(ROOT, "gpurun_out", "configs", c + ".json")
    try:
        d = json.loads(open(p).read().strip().splitlines()[-1])
    except Exception as e:
        out[c] = {"error": str(e)}
        continue
    r = d.get("roofline", {})
    k = r.get("kernel_ms_per_batch", {})
    out[c] = {"workload": d["config"]["workload"], "triangles": d["config"]["triangles"], "width": d["config"]["width"],
              "height": d["config"]["height"], "n_gpus": d["n_gpus"], "mrays_s": d["value"], "e2e_mrays_s": d["e2e"]["value"],
              "ms_per_step": d["ms_per_step"], "spp_per_s": d["spp_per_s"], "rays_per_step": d["rays_per_step"],
              "bvh_build_ms": d["bvh"]["buildMs"], "wide_nodes": d["bvh"]["numWideNodes"],
              "extend_rays_per_step": r.get("extend_rays_per_step"), "shadow_rays_per_step": r.get("shadow", {}).get("rays_per_step"),
              "extend_frac_of_hbm": r.get("frac"), "shadow_frac_of_hbm": r.get("shadow", {}).get("frac"),
              "nodes_per_extend_ray": r.get("nodes_per_ray"), "tris_per_extend_ray": r.get("tris_per_ray"),
              "kernel_ms_per_step": k, "post_ms": r.get("post_ms_per_call"),
              "as_shipped_mrays_s": d.get("as_shipped", {}).get("value"), "clocks": d.get("clocks"),
              "engines": d["config"].get("engines"), "lanes": d["config"].get("lanes_per_engine")}
json.dump(out, open(os.path.join(ROOT, "profiles", "r02_configs.json"), "w"), indent=1)
for c, v in out.items():
    print(c, {k: v.get(k) for k in ("mrays_s", "ms_per_step", "spp_per_s", "extend_frac_of_hbm", "as_shipped_mrays_s")})
