"""Scenes and sampling of the random-hit campaign against the reference's compiled closest-hit shaders
(tests/golden/make_spirv_hits_golden.py mints the fixture from the shaders, tests/test_spirv_hits.py replays it on the
oracle and on the CUDA shading code). Everything here is procedural: nothing of /root/reference is needed to rebuild the
scenes, so the fixture can be replayed on the GPU box."""
import numpy as np

N_SCENES = 8
POS = [(-0.5, 0.5, -0.3), (0.0, 0.5, 0.3), (0.5, 0.5, -0.3), (0.0, 1.2, -0.2)]


def hit_scene(rb, seed):
    """Cornell box + light + four spheres, one per material kernel, with random material parameters (every Disney lobe
    weight, anisotropy, tints, ior, absorption, roughness; albedo / normal / height maps on or off; culling), random
    non-uniform scales. Deterministic in `seed`."""
    Cf = rb.configs
    R = np.random.RandomState(1000 + seed)
    u = lambda a=0.0, b=1.0: float(R.uniform(a, b))
    s = rb.Scene()
    tex, nmap = s.defineTexture(rb.meshes.cornell_texture(32, 48)), s.defineTexture(Cf._bumpy_normal_map(32))
    leaf, hmap = s.defineTexture(Cf._leaf_texture(32)), s.defineTexture(Cf.brick_height_map(32))
    pick = lambda t: t if R.rand() < 0.5 else -1
    s.addObject(rb.meshes.cornell_box(), Cf.IDENT, rb.Material(**Cf.CORNELL_WALL))
    s.addObject(rb.meshes.cornell_light(), Cf.IDENT, rb.Material(**Cf.LIGHT))
    sph = s.defineObject(rb.meshes.uv_sphere(12, 6, radius=0.22))
    mats = [rb.Material(materialIdx=0, albedo=(u(), u(), u()), interpNormals=True, textureID=pick(leaf), normalMapID=pick(nmap),
                        bumpMapID=pick(hmap), cullBackface=bool(R.rand() < 0.5)),
            rb.Material(materialIdx=1, albedo=(u(), u(), u()), roughness=u(0, 0.6), interpNormals=True, textureID=pick(tex),
                        normalMapID=pick(nmap), bumpMapID=pick(hmap)),
            rb.Material(materialIdx=2, albedo=(u(), u(), u()), roughness=u(0, 0.4), ior=u(1.1, 2.0), absorption=u(0, 3),
                        interpNormals=True, textureID=pick(tex), normalMapID=pick(nmap)),
            rb.Material(materialIdx=3, albedo=(u(), u(), u()), roughness=u(0.05, 1), ior=u(1.1, 2), interpNormals=True, metallic=u(),
                        clearcoat=u(), clearcoatGloss=u(), specularTransmission=u(), sheen=u(), subsurface=u(), anisotropic=u(),
                        sheenTint=(u(), u(), u()), specularTint=(u(), u(), u()), textureID=pick(tex), normalMapID=pick(nmap))]
    for k in range(4):
        s.addInstance(sph, Cf.compose(Cf.translate(POS[k]), Cf.scale((u(0.7, 1.3), u(0.7, 1.3), u(0.7, 1.3)))), mats[k])
    return s.build()


def random_rays(seed, n):
    """n candidate rays for scene `seed`: aimed at one of the spheres from a random point of the box or from inside the
    sphere, un-normalised, with a random RNG state and a random dielectric entry state."""
    R = np.random.RandomState(5000 + seed)
    o = np.empty((n, 3), np.float32)
    d = np.empty((n, 3), np.float32)
    state = R.randint(0, 2 ** 31, n).astype(np.uint32)
    inside = (R.rand(n) < 0.3).astype(np.uint32)
    acc = np.where(inside != 0, R.uniform(0, 2, n), 0.0).astype(np.float32)
    for i in range(n):
        while True:
            k = int(R.randint(0, 4))
            c = np.array(POS[k], np.float32)
            oo = (c + R.normal(size=3) * 0.03).astype(np.float32) if R.rand() < 0.25 else \
                np.array([R.uniform(-0.9, 0.9), R.uniform(0.1, 1.9), R.uniform(-0.9, 0.9)], np.float32)
            dd = (c + R.uniform(-0.2, 0.2, 3)).astype(np.float32) - oo
            if np.linalg.norm(dd) >= 1e-3:
                break
        o[i] = oo
        d[i] = (dd / np.linalg.norm(dd) * R.uniform(0.5, 2.0)).astype(np.float32)
    return o, d, state, inside, acc
