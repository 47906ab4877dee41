"""Golden vectors from the reference's compiled RAY-TRACING shaders (python tests/golden/make_spirv_rt_golden.py).

raytrace.rgen.spv + lambertian / metal / dielectric / disney .rchit.spv + raytrace.rmiss.spv, executed on the CPU by
tests/spirv_interp.py and wired into a pipeline by tests/spirv_rt.py (whose header lists exactly what is supplied from
outside the binaries: the intersection, the texture filter, the libm-like GLSL.std.450 functions). The shipped binaries
have next-event estimation compiled out, so the fixture is the AS-SHIPPED estimator: flags = 0.

spirv_rt.npz, per case: the RGBA32F image after every batch (running average, raytrace.rgen.glsl:277-284) and the number
of rays traced. Cases (sizes chosen so that the interpreter needs about a minute each):
  mixed     configs.small_mixed: every material, albedo / normal / alpha textures, cull and alpha skips, instancing
  cornell   configs.cornell (BASELINE C1 at reduced size), textured walls
  parallax  configs.parallax: texutils.h.glsl bumpMapping compiled into all four closest-hit shaders
  showroom  configs.showroom_mixed (BASELINE C5 at reduced size): open scene (sky misses after bounces), a Disney blob with
            subsurface and sheen, one sphere per material id, Disney with specular transmission
The workloads are rebuilt from the same configs by tests/test_spirv_golden.py; the oracle (CPU) and the CUDA path (GPU)
must reproduce every image bit for bit."""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
SHADERS = "/root/reference/shaders/raytrace"
CASES = {"mixed": dict(config="small_mixed", width=32, height=24, samples_per_pixel=2, max_bounces=6, batches=2),
         "cornell": dict(config="cornell", width=24, height=18, samples_per_pixel=2, max_bounces=8, batches=2),
         "parallax": dict(config="parallax", width=32, height=24, samples_per_pixel=2, max_bounces=6, batches=2),
         "showroom": dict(config="showroom_mixed", width=32, height=18, samples_per_pixel=2, max_bounces=8, batches=2, levels=2)}


def workload(rb, case):
    extra = {"levels": case["levels"]} if "levels" in case else {}
    return getattr(rb.configs, case["config"])(case["width"], case["height"], nee=False, samples_per_pixel=case["samples_per_pixel"],
                                               max_bounces=case["max_bounces"], **extra)


def main():
    import oracle_lib as ol
    import spirv_rt
    out = {}
    for name, case in CASES.items():
        wl = workload(ol.rb, case)
        pipe = spirv_rt.Pipeline(SHADERS, wl.tables, ol)
        img = np.zeros((case["height"], case["width"], 4), np.float32)
        t0 = time.time()
        for b in range(case["batches"]):
            before = pipe.rays
            pipe.render_batch(wl.push_constants(b), case["width"], case["height"], img)
            out["%s_hdr_%d" % (name, b)] = img.copy()
            out["%s_rays_%d" % (name, b)] = np.int64(pipe.rays - before)
        print(name, "%.0f s" % (time.time() - t0), "rays", pipe.rays, "mean", float(np.nanmean(img[..., :3])), "NaN pixels", int(np.isnan(img).any(axis=2).sum()))
    np.savez_compressed(os.path.join(HERE, "spirv_rt.npz"), **out)


if __name__ == "__main__":
    main()
