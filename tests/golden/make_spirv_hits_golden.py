"""Mints tests/golden/spirv_hits.npz: single closest-hit invocations of the reference's COMPILED shaders
(/root/reference/shaders/raytrace/{lambertian,metal,dielectric,disney}.rchit.spv, executed by tests/spirv_interp.py /
spirv_rt.py) on random hits of the procedural scenes of tests/spirv_hit_scenes.py — every payload field the shader
leaves. Run in the build container (the reference checkout is needed):

    python tests/golden/make_spirv_hits_golden.py [hits_per_scene]

The intersection (which triangle, barycentrics) is the oracle's brute-force closest hit: the traversal lives in the
Vulkan driver and is not part of the shader binaries.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle_lib as ol          # noqa: E402
import spirv_hit_scenes as hs    # noqa: E402
import spirv_rt                  # noqa: E402

RT_SHADERS = "/root/reference/shaders/raytrace"
F = np.float32


def main():
    per_scene = int(sys.argv[1]) if len(sys.argv) > 1 else 1625
    rb = ol.rb
    rows = {k: [] for k in ("scene", "o", "d", "state", "inside", "acc", "vec", "pdf", "acc_out", "rng_out", "skip", "inside_out", "material")}
    for seed in range(hs.N_SCENES):
        pipe = spirv_rt.Pipeline(RT_SHADERS, hs.hit_scene(rb, seed), ol)
        m = pipe.rgen.m
        ptype = [m.types[pt][2] for v, (st, pt, _) in m.globals.items() if st == spirv_rt.SC_RAY_PAYLOAD][0]
        o, d, state, inside, acc = hs.random_rays(seed, per_scene * 2)
        hits = pipe.scene.trace_rays(o, d, 1e4, brute=True)
        done = 0
        for i in range(len(o)):
            if done == per_scene:
                break
            hit = hits[i]
            if hit["t"] < 0:
                continue
            inst = int(hit["instance"])
            kk = min(int(pipe.inst_material[inst]), 3)
            payload = m.zero(ptype)
            payload[4], payload[12], payload[11] = int(state[i]), bool(inside[i]), F(acc[i])
            M = pipe.inst_transform[inst]
            builtins = {spirv_rt.BUILTIN_WORLD_RAY_ORIGIN: [F(x) for x in o[i]], spirv_rt.BUILTIN_WORLD_RAY_DIRECTION: [F(x) for x in d[i]],
                        spirv_rt.BUILTIN_OBJECT_TO_WORLD: [[F(M[4 * c + r]) for r in range(3)] for c in range(4)],
                        spirv_rt.BUILTIN_INSTANCE_CUSTOM_INDEX: int(pipe.inst_props[inst]), spirv_rt.BUILTIN_PRIMITIVE_ID: int(hit["primitive"])}
            b = dict(pipe.hit_bindings[kk])
            b[("storage", spirv_rt.SC_INCOMING_RAY_PAYLOAD)] = payload
            b[("storage", spirv_rt.SC_HIT_ATTRIBUTE)] = [F(hit["u"]), F(hit["v"])]
            pipe.rchit[kk].run(b, builtins=builtins)
            # color, albedo, origin, direction, emission, normal
            vec = np.array([payload[1], payload[0], payload[2], payload[3], payload[6], payload[7]], np.float32).reshape(18)
            rows["scene"].append(seed); rows["o"].append(o[i]); rows["d"].append(d[i]); rows["state"].append(state[i])
            rows["inside"].append(inside[i]); rows["acc"].append(acc[i]); rows["vec"].append(vec)
            rows["pdf"].append(F(payload[10])); rows["acc_out"].append(F(payload[11])); rows["rng_out"].append(int(payload[4]))
            rows["skip"].append(bool(payload[9])); rows["inside_out"].append(bool(payload[12])); rows["material"].append(kk)
            done += 1
        print("scene %d: %d hits" % (seed, done), flush=True)
        pipe.scene.close()
    out = os.path.join(HERE, "spirv_hits.npz")
    np.savez_compressed(out, scene=np.array(rows["scene"], np.uint8), o=np.array(rows["o"], np.float32), d=np.array(rows["d"], np.float32),
                        state=np.array(rows["state"], np.uint32), inside=np.array(rows["inside"], np.uint8), acc=np.array(rows["acc"], np.float32),
                        vec=np.array(rows["vec"], np.float32), pdf=np.array(rows["pdf"], np.float32), acc_out=np.array(rows["acc_out"], np.float32),
                        rng_out=np.array(rows["rng_out"], np.uint32), skip=np.array(rows["skip"], np.uint8),
                        inside_out=np.array(rows["inside_out"], np.uint8), material=np.array(rows["material"], np.uint8))
    print("wrote", out, len(rows["scene"]), "hits;", {k: int((np.array(rows["material"]) == k).sum()) for k in range(4)})


if __name__ == "__main__":
    main()
