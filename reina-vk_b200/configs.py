"""The five workloads named in BASELINE.json `configs`, built from procedural geometry.

The reference hard-codes its scene in Reina::Reina() (src/Reina.cpp:91-116, camera :138-140); the Cornell materials
below are the worked example in its comments (`cornellWall`, `lightMaterial`, `subjectMaterial`, `glass`,
src/Reina.cpp:108-111). stanford_dragon / stanford_bunny / max_planck are absent from the reference snapshot
(.MISSING_LARGE_BLOBS), and nothing under /root/reference exists on the GPU box, so every mesh here is generated:
the stand-ins are labelled as such in bench output (`data`: "synthetic").
"""
import numpy as np

from . import meshes
from .camera import compose, push_constants, rotate, scale, translate
from .scene import Material, Scene

IDENT = np.eye(4, dtype=np.float32)

# src/Reina.cpp:108-111
CORNELL_WALL = dict(materialIdx=0, albedo=(0.9, 0.9, 0.9), emission=(0, 0, 0), cullBackface=True,
                    sheenTint=(0, 0, 0), specularTint=(1, 1, 1))
LIGHT = dict(materialIdx=0, albedo=(0.9, 0.9, 0.9), emission=(16.0, 16.0, 16.0), cullBackface=True,
             sheenTint=(0, 0, 0), specularTint=(1, 1, 1))
GLASS = dict(materialIdx=2, albedo=(0.2, 0.9, 0.4), roughness=0.3, ior=1.5, interpNormals=True, absorption=0.7,
             sheenTint=(0, 0, 0), specularTint=(1, 1, 1))


class Workload:
    def __init__(self, name, tables, width, height, pc_kwargs, nee, spp_total, note=""):
        self.name, self.tables, self.width, self.height = name, tables, width, height
        self.pc_kwargs, self.nee, self.spp_total, self.note = pc_kwargs, nee, spp_total, note

    def push_constants(self, sample_batch=0, **over):
        kw = dict(self.pc_kwargs)
        kw.update(over)
        return push_constants(self.width, self.height, total_emissive_weight=self.tables.totalEmissiveWeight,
                              sample_batch=sample_batch, **kw)


def cornell(width=800, height=600, with_sphere=False, nee=True, samples_per_pixel=1, max_bounces=8, textured=True):
    """C1: Cornell box + light, Lambertian, 800x600, 64 spp, 8 bounces, NEE+MIS. Camera outside the +z wall looking
    -z: that wall is a back face, so every primary ray spends one segment on a cull-skip."""
    s = Scene()
    tex = s.defineTexture(meshes.cornell_texture(256, 384)) if textured else -1
    s.addObject(meshes.cornell_box(), IDENT, Material(textureID=tex, **CORNELL_WALL))
    s.addObject(meshes.cornell_light(), IDENT, Material(**LIGHT))
    if with_sphere:
        t = compose(scale(0.25), translate((0, 2, 0)))      # glm::translate(glm::scale(I, .25), (0,2,0)), Reina.cpp:106
        s.addObject(meshes.uv_sphere(48, 24), t, Material(materialIdx=3, albedo=(1, 1, 1), roughness=0.5, ior=1.5,
                                                          interpNormals=True, sheenTint=(1, 1, 1), specularTint=(1, 1, 1)))
    tables = s.build(require_emitter=nee)
    pc = dict(pos=(0.0, 1.0, 3.9), look=(0.0, 1.0, 0.0), fovy_deg=40.0, samples_per_pixel=samples_per_pixel,
              max_bounces=max_bounces)
    return Workload("cornell", tables, width, height, pc, nee, 64)


def _showroom_scene(s):
    floor = Material(materialIdx=0, albedo=(0.8, 0.8, 0.8), interpNormals=True, cullBackface=False)
    s.addObject(meshes.showroom(), IDENT, floor)
    s.addObject(meshes.cornell_light(), IDENT, Material(**LIGHT))


def bunny(width=1920, height=1080, levels=6, nee=True, samples_per_pixel=1, max_bounces=16):
    """C2: two noise-displaced blobs (bunny stand-in, 20*4^levels triangles each) — metal and glass with Beer's-law
    absorption — in the showroom."""
    s = Scene()
    _showroom_scene(s)
    blob = s.defineObject(meshes.subdivided_blob(levels=levels, seed=0xB0771, displacement=0.35, radius=0.28))
    s.addInstance(blob, translate((-0.35, 0.3, 0.1)), Material(materialIdx=1, albedo=(0.95, 0.8, 0.6), roughness=0.1,
                                                               interpNormals=True))
    s.addInstance(blob, translate((0.35, 0.3, -0.1)), Material(**GLASS))
    tables = s.build(require_emitter=nee)
    pc = dict(pos=(-1.6899, 0.317017 + 0.4, -1.6386), look=(0.0, 0.35, 0.0), fovy_deg=25.0,
              samples_per_pixel=samples_per_pixel, max_bounces=max_bounces)
    return Workload("bunny", tables, width, height, pc, nee, 256, "stand-in geometry")


def dragon(width=1920, height=1080, n_along=10627, n_ring=41, nee=True, samples_per_pixel=8, max_bounces=16):
    """C3 (headline): Stanford-dragon stand-in — torus-knot tube with noise displacement, 871,414 triangles by
    default — Disney BSDF (albedo (0.6,0.3,0.8), roughness 0.3, metallic 0.8, clearcoat 1, ior 1.5) in the
    showroom, the default camera of src/Reina.cpp:138-140 mirrored to the open (-z) side of the cyclorama and raised,
    1080p, 16 bounces, NEE on."""
    s = Scene()
    _showroom_scene(s)
    knot = meshes.torus_knot(n_along=n_along, n_ring=n_ring, fit=((0.0, 0.62, 0.0), 1.15))
    s.addObject(knot, IDENT, Material(materialIdx=3, albedo=(0.6, 0.3, 0.8), roughness=0.3, ior=1.5, interpNormals=True,
                                      metallic=0.8, clearcoat=1.0, clearcoatGloss=0.5, sheenTint=(1, 1, 1),
                                      specularTint=(1, 1, 1)))
    tables = s.build(require_emitter=nee)
    pc = dict(pos=(-1.6899, 0.317017 + 0.5, -1.6386), look=(0.0, 0.6, 0.0), fovy_deg=30.0,
              samples_per_pixel=samples_per_pixel, max_bounces=max_bounces)
    return Workload("dragon", tables, width, height, pc, nee, 1024, "stand-in geometry")


def _leaf_texture(size=256, seed=7):
    """RGBA leaf-card atlas with a binary alpha channel (the reference's leaf textures are JPG without alpha; the
    alpha-tested path of lambertian.rchit.glsl:48-52 is exercised with a synthetic cut-out)."""
    y, x = np.mgrid[0:size, 0:size].astype(np.float32) / size
    r = np.sqrt(((x - 0.5) / 0.45) ** 2 + ((y - 0.5) / 0.3) ** 2)
    a = (r < 1.0)
    img = np.zeros((size, size, 4), np.uint8)
    img[..., 0] = 40 + 30 * np.sin(x * 40) ** 2
    img[..., 1] = 120 + 80 * np.cos(y * 25) ** 2
    img[..., 2] = 30
    img[..., 3] = np.where(a, 255, 0)
    return img


def _bumpy_normal_map(size=256):
    y, x = np.mgrid[0:size, 0:size].astype(np.float32) / size
    nx = 0.35 * np.sin(x * 2 * np.pi * 8)
    ny = 0.35 * np.cos(y * 2 * np.pi * 8)
    nz = np.sqrt(np.maximum(0.0, 1 - nx * nx - ny * ny))
    img = np.zeros((size, size, 4), np.uint8)
    img[..., 0] = np.clip((nx * 0.5 + 0.5) * 255, 0, 255)
    img[..., 1] = np.clip((ny * 0.5 + 0.5) * 255, 0, 255)
    img[..., 2] = np.clip((nz * 0.5 + 0.5) * 255, 0, 255)
    img[..., 3] = 255
    return img


def brick_height_map(size=128, rows=4, cols=2, mortar=0.08):
    """Height map in the spirit of the reference's textures/bricks2/parallax.png (which cannot travel to the GPU box):
    0 (black) = surface level, 255 = 0.2 uv-units deep. Bricks are shallow, mortar grooves deep, with a smooth ramp
    between them and a faint ripple so that neighbouring texels differ."""
    y, x = np.mgrid[0:size, 0:size].astype(np.float32) / size
    row = np.floor(y * rows)
    xs = (x + 0.5 * (row % 2) / cols) % 1.0
    fx = np.abs(((xs * cols) % 1.0) - 0.5) * 2          # 0 at the brick centre, 1 at the joint
    fy = np.abs(((y * rows) % 1.0) - 0.5) * 2
    edge = np.maximum(fx - (1 - mortar * cols), (fy - (1 - mortar * rows))) / (mortar * rows)
    groove = np.clip(edge * 2.0, 0.0, 1.0)
    ripple = 0.06 * (np.sin(x * 2 * np.pi * 9) * np.cos(y * 2 * np.pi * 7) + 1)
    hgt = np.clip(0.1 + ripple + 0.8 * groove, 0, 1)
    img = np.zeros((size, size, 4), np.uint8)
    img[..., 0] = img[..., 1] = img[..., 2] = np.round(hgt * 255)
    img[..., 3] = 255
    return img


def parallax(width=160, height=120, nee=True, samples_per_pixel=2, max_bounces=6):
    """every material with a parallax height map (texutils.h.glsl:4-41) — a Lambertian floor whose
    shifted UVs leave [0, 1] near the borders (range skip), a metal sphere (no range check: REPEAT addressing), a
    dielectric sphere and a Disney sphere that combines the height map with a normal map and an albedo texture."""
    s = Scene()
    hmap = s.defineTexture(brick_height_map(128))
    tex = s.defineTexture(meshes.cornell_texture(64, 96))
    nmap = s.defineTexture(_bumpy_normal_map(64))
    s.addObject(meshes.cornell_box(), IDENT, Material(**CORNELL_WALL))
    s.addObject(meshes.cornell_light(), IDENT, Material(**LIGHT))
    s.addObject(meshes.quad((-0.9, 0.02, 0.9), (0.9, 0.02, 0.9), (0.9, 0.02, -0.9), (-0.9, 0.02, -0.9)), IDENT,
                Material(materialIdx=0, albedo=(0.8, 0.5, 0.4), textureID=tex, bumpMapID=hmap))
    sph = s.defineObject(meshes.uv_sphere(24, 12, radius=0.3))
    s.addInstance(sph, translate((-0.55, 0.32, -0.3)), Material(materialIdx=1, albedo=(0.9, 0.8, 0.5), roughness=0.1,
                                                                interpNormals=True, textureID=tex, bumpMapID=hmap))
    glass = dict(GLASS); glass.update(bumpMapID=hmap, textureID=tex)
    s.addInstance(sph, translate((0.0, 0.32, 0.35)), Material(**glass))
    s.addInstance(sph, compose(translate((0.55, 0.37, -0.3)), scale((1.0, 1.15, 0.9))),
                  Material(materialIdx=3, albedo=(0.6, 0.3, 0.8), roughness=0.3, ior=1.5, interpNormals=True,
                           metallic=0.4, clearcoat=0.5, sheenTint=(1, 1, 1), specularTint=(1, 1, 1),
                           textureID=tex, normalMapID=nmap, bumpMapID=hmap))
    tables = s.build(require_emitter=nee)
    pc = dict(pos=(0.0, 1.3, 3.9), look=(0.0, 0.6, 0.0), fovy_deg=40.0, samples_per_pixel=samples_per_pixel,
              max_bounces=max_bounces)
    return Workload("parallax", tables, width, height, pc, nee, 2)


def plant(width=1920, height=1080, n_leaves=16000, nee=True, samples_per_pixel=1, max_bounces=16, seed=3):
    """C4: plant-like scene — a normal-mapped Disney pot, a soil disc and ~n_leaves alpha-tested two-triangle leaf cards
    (the reference's plant_* OBJs cannot travel to the GPU box; same feature set: albedo + normal textures,
    stochastic alpha skip)."""
    rng = np.random.RandomState(seed)
    s = Scene()
    _showroom_scene(s)
    leaf_tex = s.defineTexture(_leaf_texture())
    nmap = s.defineTexture(_bumpy_normal_map())
    pot = meshes.uv_sphere(64, 32, radius=0.22)
    s.addObject(pot, translate((0, 0.2, 0)), Material(materialIdx=3, albedo=(0.72, 0.45, 0.2), roughness=0.4, ior=1.5,
                                                      interpNormals=True, metallic=0.6, normalMapID=nmap,
                                                      sheenTint=(1, 1, 1), specularTint=(1, 1, 1)))
    # leaf cards: random quads around a stem
    pos, uv, nrm, tris = [], [], [], []
    for i in range(n_leaves):
        c = np.array([rng.normal(0, 0.16), 0.45 + abs(rng.normal(0, 0.22)), rng.normal(0, 0.16)])
        u = rng.normal(size=3); u /= np.linalg.norm(u)
        w = np.cross(u, rng.normal(size=3)); w /= np.linalg.norm(w)
        n = np.cross(u, w)
        a, b = 0.05 * u, 0.03 * w
        base = len(pos)
        pos += [c - a - b, c + a - b, c + a + b, c - a + b]
        uv += [(0, 0), (1, 0), (1, 1), (0, 1)]
        nrm += [n, n, n, n]
        tris += [(base, base + 1, base + 2), (base, base + 2, base + 3)]
    leaves = meshes.make_model(np.array(pos, np.float32), np.array(uv, np.float32), np.array(nrm, np.float32),
                               np.array(tris, np.uint32))
    s.addObject(leaves, IDENT, Material(materialIdx=0, albedo=(0.9, 0.9, 0.9), textureID=leaf_tex, interpNormals=False))
    tables = s.build(require_emitter=nee)
    pc = dict(pos=(-1.6899, 0.317017 + 0.5, -1.6386), look=(0.0, 0.55, 0.0), fovy_deg=30.0,
              samples_per_pixel=samples_per_pixel, max_bounces=max_bounces)
    return Workload("plant", tables, width, height, pc, nee, 256, "stand-in geometry")


def showroom_mixed(width=3840, height=2160, levels=6, nee=True, samples_per_pixel=1, max_bounces=16):
    """C5: showroom + Max-Planck stand-in (displaced icosphere) + one sphere per material id, bloom + tonemap with the
    config defaults, 3840x2160."""
    s = Scene()
    _showroom_scene(s)
    s.addObject(meshes.subdivided_blob(levels=levels, seed=0x9A8, displacement=0.3, radius=0.33), translate((0, 0.4, 0)),
                Material(materialIdx=3, albedo=(0.8, 0.75, 0.7), roughness=0.5, ior=1.5, interpNormals=True,
                         subsurface=0.3, sheen=0.5, sheenTint=(1, 1, 1), specularTint=(1, 1, 1)))
    sph = s.defineObject(meshes.uv_sphere(64, 32, radius=0.12))
    s.addInstance(sph, translate((-0.6, 0.12, 0.5)), Material(materialIdx=0, albedo=(0.8, 0.2, 0.2), interpNormals=True))
    s.addInstance(sph, translate((-0.2, 0.12, 0.65)), Material(materialIdx=1, albedo=(0.9, 0.9, 0.9), roughness=0.05,
                                                               interpNormals=True))
    s.addInstance(sph, translate((0.2, 0.12, 0.65)), Material(**GLASS))
    s.addInstance(sph, translate((0.6, 0.12, 0.5)), Material(materialIdx=3, albedo=(0.2, 0.4, 0.9), roughness=0.2, ior=1.5,
                                                             interpNormals=True, specularTransmission=0.9,
                                                             sheenTint=(1, 1, 1), specularTint=(1, 1, 1)))
    tables = s.build(require_emitter=nee)
    pc = dict(pos=(-1.6899, 0.317017 + 0.5, -1.6386), look=(0.0, 0.4, 0.0), fovy_deg=30.0,
              samples_per_pixel=samples_per_pixel, max_bounces=max_bounces)
    return Workload("showroom_mixed", tables, width, height, pc, nee, 4096, "stand-in geometry")


def two_lights(width=96, height=72, emitters_first=True, nee=True, samples_per_pixel=2, max_bounces=6):
    """Two emitters (NEE with more than one light, nee.h.glsl:52-124). nee.h.glsl:97-105 addresses an emitter's triangles
    by their position in the CONCATENATED triangle CDF (indices[3 * cdfIndex + indexOffset]), so the second emitter (CDF
    slots 2..49) reads 2 triangles past its own mesh — with the emitters first those slots still lie inside the index
    buffer (the box's triangles, moved by the emitter's transform: upstream's behaviour, restated literally); with the
    emitters last they lie past its end and the scene is refused with NEE on."""
    s = Scene()
    if not emitters_first:
        s.addObject(meshes.cornell_box(), IDENT, Material(**CORNELL_WALL))
    s.addObject(meshes.cornell_light(), IDENT, Material(**LIGHT))
    # a second, differently sized, double-sided and differently coloured emitter with a non-trivial transform
    M = compose(translate((-0.5, 0.9, -0.2)), scale((0.5, 1.0, 0.7)))
    s.addObject(meshes.uv_sphere(8, 4, radius=0.2), M, Material(materialIdx=0, albedo=(1, 1, 1), emission=(2.0, 7.0, 4.0), cullBackface=False))
    if emitters_first:
        s.addObject(meshes.cornell_box(), IDENT, Material(**CORNELL_WALL))
    tables = s.build(require_emitter=True)
    pc = dict(pos=(0.0, 1.0, 3.9), look=(0.0, 1.0, 0.0), fovy_deg=40.0, samples_per_pixel=samples_per_pixel,
              max_bounces=max_bounces)
    return Workload("two_lights", tables, width, height, pc, nee, 2)


def small_mixed(width=160, height=120, nee=True, samples_per_pixel=2, max_bounces=8, textured=True):
    """A small closed scene touching every material, texture and skip path; sized so the CPU oracle renders it in
    seconds. Used by the parity tests and by __graft_entry__.smoke()."""
    s = Scene()
    tex = s.defineTexture(meshes.cornell_texture(64, 96)) if textured else -1
    nmap = s.defineTexture(_bumpy_normal_map(64)) if textured else -1
    leaf = s.defineTexture(_leaf_texture(64)) if textured else -1
    s.addObject(meshes.cornell_box(), IDENT, Material(textureID=tex, **CORNELL_WALL))
    s.addObject(meshes.cornell_light(), IDENT, Material(**LIGHT))
    sph = s.defineObject(meshes.uv_sphere(24, 12, radius=0.3))
    s.addInstance(sph, translate((-0.55, 0.3, -0.3)), Material(materialIdx=1, albedo=(0.9, 0.8, 0.5), roughness=0.15,
                                                               interpNormals=True, normalMapID=nmap))
    s.addInstance(sph, translate((0.0, 0.3, 0.35)), Material(**GLASS))
    s.addInstance(sph, compose(translate((0.55, 0.35, -0.3)), scale((1.0, 1.15, 0.9))),
                  Material(materialIdx=3, albedo=(0.6, 0.3, 0.8), roughness=0.3, ior=1.5, interpNormals=True,
                           metallic=0.5, clearcoat=1.0, clearcoatGloss=0.6, specularTransmission=0.3, sheen=0.4,
                           subsurface=0.2, anisotropic=0.3, sheenTint=(1, 1, 1), specularTint=(1, 1, 1)))
    s.addObject(meshes.quad((-0.5, 0.9, -0.6), (0.5, 0.9, -0.6), (0.5, 1.6, -0.6), (-0.5, 1.6, -0.6)), IDENT,
                Material(materialIdx=0, albedo=(0.9, 0.9, 0.9), textureID=leaf))
    tables = s.build(require_emitter=nee)
    pc = dict(pos=(0.0, 1.0, 3.9), look=(0.0, 1.0, 0.0), fovy_deg=40.0, samples_per_pixel=samples_per_pixel,
              max_bounces=max_bounces)
    return Workload("small_mixed", tables, width, height, pc, nee, 2)


def instanced(width=160, height=120, grid=4, segments=24, rings=12, nee=True, samples_per_pixel=2, max_bounces=8, seed=11):
    """grid x grid instances of ONE sphere model and as many of ONE box-ish blob, each with its own rotation, non-uniform
    scale, translation and material, over the Cornell box and its light: what the reference's Scene keeps as one BLAS per
    object and a TLAS entry per instance (src/scene/Scene.cpp:93-111). Workload of the two-level tests."""
    rng = np.random.RandomState(seed)
    s = Scene()
    s.addObject(meshes.cornell_box(), IDENT, Material(**CORNELL_WALL))
    s.addObject(meshes.cornell_light(), IDENT, Material(**LIGHT))
    sph = s.defineObject(meshes.uv_sphere(segments, rings, radius=1.0))
    blob = s.defineObject(meshes.subdivided_blob(levels=2, seed=0x51, displacement=0.3, radius=1.0))
    mats = [dict(materialIdx=0, albedo=(0.8, 0.5, 0.3)),
            dict(materialIdx=1, albedo=(0.9, 0.8, 0.5), roughness=0.2, interpNormals=True),
            dict(GLASS),
            dict(materialIdx=3, albedo=(0.5, 0.4, 0.8), roughness=0.35, ior=1.5, interpNormals=True, metallic=0.4,
                 clearcoat=0.5, clearcoatGloss=0.6, sheenTint=(1, 1, 1), specularTint=(1, 1, 1))]
    k = 0
    for iy in range(grid):
        for ix in range(grid):
            for obj, lift in ((sph, 0.0), (blob, 0.45)):
                r = 0.32 / grid * (1.0 + 0.5 * rng.rand())
                pos = (-0.75 + 1.5 * (ix + 0.5) / grid + 0.02 * rng.randn(), 0.18 + lift + 0.9 * rng.rand() * (lift + 0.2),
                       -0.75 + 1.5 * (iy + 0.5) / grid + 0.02 * rng.randn())
                M = compose(translate(pos), rotate(rng.randn(3), rng.rand() * 6.28),
                            scale((r, r * (0.6 + 0.8 * rng.rand()), r * (0.6 + 0.8 * rng.rand()))))
                s.addInstance(obj, M, Material(**mats[k % 4]))
                k += 1
    tables = s.build(require_emitter=nee)
    pc = dict(pos=(0.0, 1.0, 3.9), look=(0.0, 0.8, 0.0), fovy_deg=40.0, samples_per_pixel=samples_per_pixel,
              max_bounces=max_bounces)
    return Workload("instanced", tables, width, height, pc, nee, 2)
