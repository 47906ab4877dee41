"""The reference's compiled RAY-TRACING pipeline on the CPU (test infrastructure): raytrace.rgen.spv, the four
*.rchit.spv closest-hit shaders and raytrace.rmiss.spv executed by tests/spirv_interp.py, wired together the way
vkCmdTraceRaysKHR wires them — OpTraceRayKHR looks the ray up, then runs the hit group the instance's shader-binding-
table offset selects (its material id, src/tools/vktools.cpp:479-486) or the miss shader, on the caller's payload.

What the modules leave to the Vulkan implementation is supplied from outside and stated here:
  * the intersection itself (acceleration structure + ray/triangle test live in the driver: SURVEY.md 8c) — the
    oracle's brute-force closest hit, i.e. the closest-hit RULE of this repository (smallest t, ties to the smallest
    primitive id); the shaders receive what the driver would hand them: instance custom index, primitive id,
    barycentrics, object-to-world matrix, the ray;
  * texture sampling (RGBA8 UNORM, bilinear, REPEAT, LOD 0: src/tools/vktools.cpp:765-788) — the filter arithmetic of
    oracle/pathtrace.cpp, restated in fp32 below;
  * GLSL.std.450 functions without a prescribed precision (sin, cos, exp, log, pow, acos, normalize, length,
    inversesqrt, inverse) — the shared elementary layer (csrc/rb_math.h, rb_vec.h); pow(x, 2) and pow(x, 5), which the
    reference also calls with negative x (undefined in GLSL), are products — the documented deviation of SURVEY.md A2.
Everything else — the RNG, the camera, the bounce loop, every material, the accumulation — is the reference's compiled
code. The shipped binaries were compiled with NEE switched off (raytrace.rgen.spv holds one OpTraceRayKHR and no
shadow ray: SURVEY.md F1), so this pipeline is compared with the oracle's as-shipped estimator (flags = 0).
"""
import os
import struct

import numpy as np

import spirv_interp as sp

F32 = np.float32
MASK = 0xFFFFFFFF
BUILTIN_LAUNCH_ID, BUILTIN_LAUNCH_SIZE = 5319, 5320
BUILTIN_WORLD_RAY_ORIGIN, BUILTIN_WORLD_RAY_DIRECTION = 5321, 5322
BUILTIN_INSTANCE_CUSTOM_INDEX, BUILTIN_OBJECT_TO_WORLD, BUILTIN_PRIMITIVE_ID = 5327, 5330, 7
SC_RAY_PAYLOAD, SC_HIT_ATTRIBUTE, SC_INCOMING_RAY_PAYLOAD = 5338, 5339, 5342
OP_TRACE_RAY = 4445


def decode(m, tid, buf, off=0, mstride=None):
    """bytes -> nested python value for SPIR-V type `tid` laid out by its Offset / ArrayStride / MatrixStride decorations."""
    t = m.types[tid]
    if t[0] == "float":
        return F32(struct.unpack_from("<f", buf, off)[0])
    if t[0] == "int":
        return struct.unpack_from("<I", buf, off)[0]
    if t[0] == "vector":
        return [decode(m, t[1], buf, off + 4 * k) for k in range(t[2])]
    if t[0] == "matrix":
        stride = mstride or 4 * m.types[t[1]][2]
        return [decode(m, t[1], buf, off + c * stride) for c in range(t[2])]
    if t[0] == "struct":
        out = []
        for k, mt in enumerate(t[1]):
            d = m.mdecor.get((tid, k), {})
            out.append(decode(m, mt, buf, off + d.get(35, [0])[0], d.get(7, [None])[0]))
        return out
    if t[0] in ("runtimearray", "array"):
        stride = m.decor[tid][6][0]
        n = (len(buf) - off) // stride if t[0] == "runtimearray" else m.consts[t[2]]
        return [decode(m, t[1], buf, off + i * stride, mstride) for i in range(n)]
    raise NotImplementedError(t)


class Texture:
    """RGBA8 UNORM, bilinear, REPEAT, LOD 0 — the arithmetic of oracle/pathtrace.cpp sample_texture, in fp32."""

    def __init__(self, rgba8):
        self.a = np.ascontiguousarray(rgba8, np.uint8)

    def sample(self, uv):
        h, w = self.a.shape[:2]
        x = F32(F32(uv[0] * F32(w)) - F32(0.5))
        y = F32(F32(uv[1] * F32(h)) - F32(0.5))
        fx, fy = F32(np.floor(x)), F32(np.floor(y))
        ax, ay = F32(x - fx), F32(y - fy)
        if not abs(fx) < 1073741824.0:
            fx, ax = F32(0), F32(0)
        if not abs(fy) < 1073741824.0:
            fy, ay = F32(0), F32(0)
        x0, y0 = int(fx) % w, int(fy) % h
        x1, y1 = (x0 + 1) % w, (y0 + 1) % h
        out = []
        one = F32(1)
        for c in range(4):
            c00, c10 = F32(F32(self.a[y0, x0, c]) / F32(255)), F32(F32(self.a[y0, x1, c]) / F32(255))
            c01, c11 = F32(F32(self.a[y1, x0, c]) / F32(255)), F32(F32(self.a[y1, x1, c]) / F32(255))
            top = F32(F32(c00 * F32(one - ax)) + F32(c10 * ax))
            bot = F32(F32(c01 * F32(one - ax)) + F32(c11 * ax))
            out.append(F32(F32(top * F32(one - ay)) + F32(bot * ay)))
        return out


def elementary_layer(ol):
    """GLSL.std.450 functions with implementation-defined precision -> csrc/rb_math.h / rb_vec.h definitions."""
    one = lambda fn: (lambda v: F32(ol.rb_math(fn, np.array([v], np.float32))[0]))
    rb_log, rb_exp = one(2), one(3)

    def dot(a, b):
        s = F32(a[0] * b[0])
        for x, y in zip(a[1:], b[1:]):
            s = F32(s + F32(x * y))
        return s

    def normalize(v):                       # rb_normalize: a * (1 / sqrt(dot(a, a)))
        inv = F32(F32(1) / F32(np.sqrt(dot(v, v))))
        return [F32(c * inv) for c in v]

    def cross(a, b):
        return [F32(F32(a[1] * b[2]) - F32(a[2] * b[1])), F32(F32(a[2] * b[0]) - F32(a[0] * b[2])), F32(F32(a[0] * b[1]) - F32(a[1] * b[0]))]

    def inverse3(M):                        # rb_m3_inverse_transpose by cofactors, transposed back
        c0, c1, c2 = M
        r = [cross(c1, c2), cross(c2, c0), cross(c0, c1)]
        inv = F32(F32(1) / dot(c0, r[0]))
        it = [[F32(x * inv) for x in rj] for rj in r]       # columns of the inverse transpose
        return [[it[r_][c_] for r_ in range(3)] for c_ in range(3)]

    def power(x, y):
        # GLSL leaves pow(x, y) undefined for x < 0, and the reference calls pow(x, 2) / pow(x, 5) on values that do go
        # negative (a real GPU returns NaN there and the sample is dropped). Documented deviation of this repository
        # (SURVEY.md A2): integer powers 2 and 5 are products — rb_sq, rb_pow5 — and everything else exp(y * log(x)).
        if y == 2.0:
            return F32(x * x)
        if y == 5.0:
            x2 = F32(x * x)
            return F32(F32(x2 * x2) * x)
        return rb_exp(F32(y * rb_log(x)))
    return {"sin": one(0), "cos": one(1), "log": rb_log, "exp": rb_exp, "exp2": one(4), "acos": one(5),
            "normalize": normalize, "inverse": inverse3, "pow": power,
            "inversesqrt": lambda v: F32(F32(1) / F32(np.sqrt(v)))}


class Pipeline:
    def __init__(self, shader_dir, tables, ol, brute=True):
        self.ol, self.tables = ol, tables
        self.scene = ol.OracleScene(tables)
        self.brute = brute
        ext = elementary_layer(ol)
        ext["ops"] = {OP_TRACE_RAY: self._trace_ray}
        load = lambda f: sp.Interpreter(sp.Module(os.path.join(shader_dir, f)), ext, max_steps=50_000_000)
        self.rgen = load("raytrace.rgen.spv")
        self.rmiss = load("raytrace.rmiss.spv")
        self.rchit = [load(f) for f in ("lambertian.rchit.spv", "metal.rchit.spv", "dielectric.rchit.spv", "disney.rchit.spv")]
        t = tables
        raw = {2: t.vertices.tobytes(), 3: t.indices.tobytes(), 4: t.instanceProperties.tobytes()[: t.numInstanceProperties * 120],
               5: t.tbns.tobytes(), 6: t.tbnIndices.tobytes(), 10: t.texCoords.tobytes(), 11: t.texIndices.tobytes()}
        self.textures = [Texture(x) for x in t.textures]
        self.hit_bindings = []
        for it in self.rchit:
            m = it.m
            b = {}
            for vid, (storage, ptype, _) in m.globals.items():
                dec = m.decor.get(vid, {})
                if storage == 12 and 33 in dec:
                    b[(dec[34][0], dec[33][0])] = decode(m, m.types[ptype][2], raw[dec[33][0]])
                elif storage == 0 and 33 in dec and dec[33][0] == 13:
                    b[(dec[34][0], 13)] = self.textures
            self.hit_bindings.append(b)
        inst = np.frombuffer(t.instances.tobytes()[: t.numInstances * 80], np.uint32).reshape(-1, 20)
        self.inst_transform = np.frombuffer(t.instances.tobytes()[: t.numInstances * 80], np.float32).reshape(-1, 20)[:, :16]
        self.inst_props, self.inst_material = inst[:, 16], inst[:, 17]
        self.rays = 0

    # OpTraceRayKHR accel flags cullMask sbtOffset sbtStride missIndex origin tmin direction tmax payload
    def _trace_ray(self, interp, env, a):
        V = lambda i: interp.val(env, i)
        origin, direction, tmax = V(a[6]), V(a[8]), V(a[9])
        payload_ptr = V(a[10])
        o = np.array([origin], np.float32)
        d = np.array([direction], np.float32)
        hit = self.scene.trace_rays(o, d, float(tmax), brute=self.brute, threads=1)[0]
        self.rays += 1
        payload = payload_ptr.load()          # the callee works on the caller's payload object (by reference)
        common = {BUILTIN_WORLD_RAY_ORIGIN: [F32(x) for x in origin], BUILTIN_WORLD_RAY_DIRECTION: [F32(x) for x in direction]}
        if hit["t"] < 0:
            self.rmiss.run({("storage", SC_INCOMING_RAY_PAYLOAD): payload}, builtins=common)
            return
        i = int(hit["instance"])
        M = self.inst_transform[i]
        common[BUILTIN_OBJECT_TO_WORLD] = [[F32(M[4 * c + r]) for r in range(3)] for c in range(4)]
        common[BUILTIN_INSTANCE_CUSTOM_INDEX] = int(self.inst_props[i])
        common[BUILTIN_PRIMITIVE_ID] = int(hit["primitive"])
        k = min(int(self.inst_material[i]), 3)
        b = dict(self.hit_bindings[k])
        b[("storage", SC_INCOMING_RAY_PAYLOAD)] = payload
        b[("storage", SC_HIT_ATTRIBUTE)] = [F32(hit["u"]), F32(hit["v"])]
        self.rchit[k].run(b, builtins=common)

    def render_batch(self, pc, width, height, image):
        """One vkCmdTraceRaysKHR(W, H, 1) with the push-constant block `pc` over the RGBA32F `image` (in place)."""
        m = self.rgen.m
        pc_var = [v for v, (st, _, _) in m.globals.items() if st == 9][0]
        pc_val = decode(m, m.types[m.globals[pc_var][1]][2], bytes(pc))
        img = sp.Image(image)
        payload_type = [m.types[pt][2] for v, (st, pt, _) in m.globals.items() if st == SC_RAY_PAYLOAD][0]
        for y in range(height):
            for x in range(width):
                self.rgen.run({"push_constant": pc_val, (0, 0): img, ("storage", SC_RAY_PAYLOAD): m.zero(payload_type)},
                              builtins={BUILTIN_LAUNCH_ID: [x, y, 0], BUILTIN_LAUNCH_SIZE: [width, height, 1]})
        return image
