"""A small SPIR-V interpreter (test infrastructure): executes the reference's own COMPILED shaders
(/root/reference/shaders/**/*.spv, committed upstream next to their GLSL sources) one invocation at a time on the CPU, so
that golden vectors for the oracle can be minted from the reference's artefacts themselves instead of from a reading of
its GLSL (the reference has no tests, and no Vulkan driver / glslang / spirv-cross exists in this image).

Scope: the logical-addressing subset glslang + spirv-opt emit for these shaders — scalar / vector / matrix fp32 and
32-bit integer arithmetic, composites, structured control flow with OpPhi, function calls, Function / Private / Input /
PushConstant / UniformConstant / StorageBuffer variables, storage images, GLSL.std.450. Every fp32 operation rounds to
fp32 (numpy float32 scalars); nothing is contracted or reassociated. The GLSL.std.450 functions whose precision GLSL
leaves to the implementation (sin, cos, exp, exp2, log, pow, acos ...) are pluggable (`ext_math`): the golden scripts
plug in the repository's elementary layer (csrc/rb_math.h through the oracle library), because that is the one place
where a choice has to be made — everything else (constants, operation order, control flow) is the reference's.
"""
import struct

import numpy as np

F32 = np.float32
MASK = 0xFFFFFFFF


def _s32(x):
    x &= MASK
    return x - (1 << 32) if x & 0x80000000 else x


class Module:
    def __init__(self, path):
        raw = open(path, "rb").read()
        w = struct.unpack("<%dI" % (len(raw) // 4), raw)
        assert w[0] == 0x07230203, "not a SPIR-V module"
        self.bound = w[3]
        self.types, self.consts, self.names, self.decor, self.mdecor = {}, {}, {}, {}, {}
        self.globals, self.functions, self.ext_sets, self.entry = {}, {}, {}, None
        self.insts = []
        i = 5
        while i < len(w):
            op, n = w[i] & 0xFFFF, w[i] >> 16
            self.insts.append((op, list(w[i + 1: i + n])))
            i += n
        self._index()

    @staticmethod
    def _string(words):
        b = b"".join(struct.pack("<I", x) for x in words)
        return b.split(b"\0", 1)[0].decode()

    def _index(self):
        T, C = self.types, self.consts
        fn = None
        block = None
        for op, a in self.insts:
            if op == 5:
                self.names[a[0]] = self._string(a[1:])
            elif op == 11:
                self.ext_sets[a[0]] = self._string(a[1:])
            elif op == 15:
                self.entry = a[1]
            elif op == 71:
                self.decor.setdefault(a[0], {})[a[1]] = a[2:]
            elif op == 72:
                self.mdecor.setdefault((a[0], a[1]), {})[a[2]] = a[3:]
            elif op == 19:
                T[a[0]] = ("void",)
            elif op == 20:
                T[a[0]] = ("bool",)
            elif op == 21:
                T[a[0]] = ("int", a[1], a[2])
            elif op == 22:
                T[a[0]] = ("float", a[1])
            elif op == 23:
                T[a[0]] = ("vector", a[1], a[2])
            elif op == 24:
                T[a[0]] = ("matrix", a[1], a[2])
            elif op == 25:
                T[a[0]] = ("image",) + tuple(a[1:])
            elif op == 26:
                T[a[0]] = ("sampler",)
            elif op == 27:
                T[a[0]] = ("sampledimage", a[1])
            elif op == 28:
                T[a[0]] = ("array", a[1], a[2])
            elif op == 29:
                T[a[0]] = ("runtimearray", a[1])
            elif op == 30:
                T[a[0]] = ("struct", a[1:])
            elif op == 32:
                T[a[0]] = ("pointer", a[1], a[2])
            elif op == 33:
                T[a[0]] = ("function", a[1], a[2:])
            elif op == 5341:      # OpTypeAccelerationStructureKHR
                T[a[0]] = ("accel",)
            elif op == 41:
                C[a[1]] = True
            elif op == 42:
                C[a[1]] = False
            elif op == 43:
                t = T[a[0]]
                C[a[1]] = F32(struct.unpack("<f", struct.pack("<I", a[2]))[0]) if t[0] == "float" else a[2] & MASK
            elif op == 44:
                C[a[1]] = [C[x] for x in a[2:]]
            elif op == 46:        # OpConstantNull
                C[a[1]] = self.zero(a[0])
            elif op == 59 and fn is None:
                self.globals[a[1]] = (a[2], a[0], a[3] if len(a) > 3 else None)
            elif op == 54:
                fn = {"id": a[1], "type": a[0], "params": [], "blocks": {}, "first": None}
                self.functions[a[1]] = fn
            elif op == 55:
                fn["params"].append(a[1])
            elif op == 56:
                fn = None
            elif op == 248:
                block = []
                fn["blocks"][a[0]] = block
                if fn["first"] is None:
                    fn["first"] = a[0]
            elif fn is not None and op not in (8,):      # OpLine
                block.append((op, a))

    def zero(self, tid):
        t = self.types[tid]
        if t[0] == "float":
            return F32(0)
        if t[0] == "int":
            return 0
        if t[0] == "bool":
            return False
        if t[0] in ("vector", "matrix"):
            return [self.zero(t[1]) for _ in range(t[2])]
        if t[0] == "array":
            return [self.zero(t[1]) for _ in range(self.consts[t[2]])]
        if t[0] == "struct":
            return [self.zero(m) for m in t[1]]
        if t[0] == "runtimearray":
            return []
        return None


class Pointer:
    __slots__ = ("root", "path")

    def __init__(self, root, path=()):
        self.root, self.path = root, path

    def load(self):
        v = self.root
        for p in self.path:
            v = v[p]
        return v

    def store(self, val):
        v = self.root
        for p in self.path[:-1]:
            v = v[p]
        v[self.path[-1]] = val


def _copy(v):
    return [_copy(x) for x in v] if isinstance(v, list) else v


class Interpreter:
    """One invocation at a time. `bindings`: {variable name or (set, binding): python object}; scalars/composites are
    plain values (they are wrapped into cells), images are objects with read(coord) / write(coord, texel) / size()."""

    def __init__(self, module, ext_math=None, max_steps=5_000_000):
        self.m = module
        self.ext = ext_math or {}
        self.max_steps = max_steps

    # -- helpers -------------------------------------------------------------------------------------------------
    def _is_float(self, tid):
        t = self.m.types[tid]
        return t[0] == "float" or (t[0] in ("vector", "matrix") and self._is_float(t[1]))

    def _signed(self, tid):
        t = self.m.types[tid]
        return t[2] == 1 if t[0] == "int" else self._signed(t[1])

    @staticmethod
    def _map(f, *vs):
        if isinstance(vs[0], list):
            return [Interpreter._map(f, *[v[i] if isinstance(v, list) else v for v in vs]) for i in range(len(vs[0]))]
        return f(*vs)

    def _math(self, name, default):
        return self.ext.get(name, default)

    # -- execution -----------------------------------------------------------------------------------------------
    def run(self, bindings, builtins=None, entry=None):
        m = self.m
        self.cells = {}
        for vid, (storage, ptype, init) in m.globals.items():
            name = m.names.get(vid, "")
            dec = m.decor.get(vid, {})
            key = (dec[34][0], dec[33][0]) if 34 in dec and 33 in dec else None
            val = None
            if 11 in dec and builtins is not None and dec[11][0] in builtins:      # BuiltIn
                val = builtins[dec[11][0]]
            elif storage == 9 and "push_constant" in bindings:
                val = _copy(bindings["push_constant"])
            elif ("storage", storage) in bindings:           # e.g. ray payload (5338 / 5342), hit attributes (5339)
                val = bindings[("storage", storage)]
            elif name in bindings:
                val = bindings[name]
            elif key in bindings:
                val = bindings[key]
            elif init is not None:
                val = _copy(m.consts[init])
            else:
                val = m.zero(m.types[ptype][2])
            self.cells[vid] = [val]
        self.steps = 0
        return self.call(entry or m.entry, [])

    def call(self, fid, args):
        m = self.m
        fn = m.functions[fid]
        env = dict(zip(fn["params"], args))
        label, prev = fn["first"], None
        while True:
            block = fn["blocks"][label]
            nxt = None
            # phis read the values of the predecessor: evaluate them together first
            phis = {}
            for op, a in block:
                if op != 245:
                    break
                for k in range(2, len(a), 2):
                    if a[k + 1] == prev:
                        phis[a[1]] = self.val(env, a[k])
            env.update(phis)
            for op, a in block:
                self.steps += 1
                if self.steps > self.max_steps:
                    raise RuntimeError("step limit exceeded")
                if op == 245 or op in (246, 247):
                    continue
                if op == 249:
                    nxt = a[0]
                    break
                if op == 250:
                    nxt = a[1] if self.val(env, a[0]) else a[2]
                    break
                if op == 251:          # OpSwitch
                    sel = self.val(env, a[0])
                    nxt = a[1]
                    for k in range(2, len(a), 2):
                        if a[k] == sel:
                            nxt = a[k + 1]
                    break
                if op == 253:
                    return None
                if op == 254:
                    return self.val(env, a[0])
                if op in (252, 255):   # OpKill, OpUnreachable
                    return None
                self.step(env, op, a)
            prev, label = label, nxt

    def val(self, env, i):
        if i in env:
            return env[i]
        if i in self.m.consts:
            return self.m.consts[i]
        if i in self.cells:
            return Pointer(self.cells[i], (0,))
        raise KeyError("undefined id %d" % i)

    def step(self, env, op, a):
        m, V = self.m, (lambda i: self.val(env, i))
        mp = self._map
        if op == 59:                                   # OpVariable (function storage)
            t = m.types[a[0]][2]
            env[a[1]] = Pointer([_copy(m.consts[a[3]]) if len(a) > 3 else m.zero(t)], (0,))
        elif op == 61:
            env[a[1]] = _copy(V(a[2]).load())
        elif op == 62:
            V(a[0]).store(_copy(V(a[1])))
        elif op in (65, 66):                           # OpAccessChain / InBounds
            p = V(a[2])
            idx = []
            for x in a[3:]:
                v = V(x)
                idx.append(_s32(v) if not isinstance(v, bool) else int(v))
            env[a[1]] = Pointer(p.root, p.path + tuple(idx))
        elif op == 57:
            env[a[1]] = self.call(a[2], [V(x) for x in a[3:]])
        elif op == 12:
            env[a[1]] = self.ext_inst(a[0], a[3], [V(x) for x in a[4:]])
        elif op == 79:
            src = list(V(a[2])) + list(V(a[3]))
            env[a[1]] = [src[k] for k in a[4:]]
        elif op == 80:
            out = []
            for x in a[2:]:
                v = V(x)
                if isinstance(v, list) and m.types[a[0]][0] == "vector":
                    out.extend(v)
                else:
                    out.append(_copy(v))
            env[a[1]] = out
        elif op == 81:
            v = V(a[2])
            for k in a[3:]:
                v = v[k]
            env[a[1]] = _copy(v)
        elif op == 82:
            obj = _copy(V(a[3]))
            t = obj
            for k in a[4:-1]:
                t = t[k]
            t[a[-1]] = _copy(V(a[2]))
            env[a[1]] = obj
        elif op == 83:
            env[a[1]] = _copy(V(a[2]))
        elif op == 104:
            env[a[1]] = list(V(a[2]).size())
        elif op == 98:
            c = V(a[3])
            env[a[1]] = V(a[2]).read([_s32(x) for x in c])
        elif op == 99:
            c = V(a[1])
            V(a[0]).write([_s32(x) for x in c], V(a[2]))
        elif op == 100:
            env[a[1]] = V(a[2])
        elif op == 88:                                 # OpImageSampleExplicitLod (Lod 0): the sampler object decides
            env[a[1]] = V(a[2]).sample(V(a[3]))
        elif op in (109, 110):                         # ConvertFToU / FToS: truncation toward zero
            env[a[1]] = mp(lambda x: int(np.trunc(np.float64(x))) & MASK if np.isfinite(x) else 0, V(a[2]))
        elif op == 111:
            env[a[1]] = mp(lambda x: F32(_s32(x)), V(a[2]))
        elif op == 112:
            env[a[1]] = mp(lambda x: F32(x & MASK), V(a[2]))
        elif op == 124:                                # OpBitcast
            src, dstf = V(a[2]), self._is_float(a[0])
            def bc(x):
                if isinstance(x, (F32, float)):
                    return x if dstf else struct.unpack("<I", struct.pack("<f", x))[0]
                return F32(struct.unpack("<f", struct.pack("<I", x & MASK))[0]) if dstf else x & MASK
            env[a[1]] = mp(bc, src)
        elif op == 126:
            env[a[1]] = mp(lambda x: (-_s32(x)) & MASK, V(a[2]))
        elif op == 127:
            env[a[1]] = mp(lambda x: F32(-x), V(a[2]))
        elif op in (128, 130, 132):
            f = {128: lambda x, y: (x + y) & MASK, 130: lambda x, y: (x - y) & MASK, 132: lambda x, y: (x * y) & MASK}[op]
            env[a[1]] = mp(f, V(a[2]), V(a[3]))
        elif op in (129, 131, 133, 136):
            f = {129: lambda x, y: F32(x + y), 131: lambda x, y: F32(x - y), 133: lambda x, y: F32(x * y),
                 136: lambda x, y: F32(np.divide(x, y))}[op]
            with np.errstate(all="ignore"):
                env[a[1]] = mp(f, V(a[2]), V(a[3]))
        elif op == 134:
            env[a[1]] = mp(lambda x, y: (x // y) & MASK if y else 0, V(a[2]), V(a[3]))
        elif op == 135:
            env[a[1]] = mp(lambda x, y: int(_s32(x) / _s32(y)) & MASK if y else 0, V(a[2]), V(a[3]))
        elif op == 137:
            env[a[1]] = mp(lambda x, y: (x % y) & MASK if y else 0, V(a[2]), V(a[3]))
        elif op in (138, 139):                         # SRem / SMod
            def srem(x, y):
                x, y = _s32(x), _s32(y)
                if y == 0:
                    return 0
                r = abs(x) % abs(y)
                r = -r if x < 0 else r
                if op == 139 and r != 0 and (r < 0) != (y < 0):
                    r += y
                return r & MASK
            env[a[1]] = mp(srem, V(a[2]), V(a[3]))
        elif op == 141:                                # OpFMod: x - y * floor(x / y)
            with np.errstate(all="ignore"):
                env[a[1]] = mp(lambda x, y: F32(x - F32(y * F32(np.floor(F32(x / y))))), V(a[2]), V(a[3]))
        elif op == 142:
            s = V(a[3])
            env[a[1]] = [F32(x * s) for x in V(a[2])]
        elif op == 143:
            s = V(a[3])
            env[a[1]] = [[F32(x * s) for x in col] for col in V(a[2])]
        elif op == 144:                                # vector * matrix: result[c] = dot(v, M[c])
            v, M = V(a[2]), V(a[3])
            env[a[1]] = [self._dot(v, col) for col in M]
        elif op == 145:                                # matrix * vector: sum_c M[c] * v[c], accumulated in column order
            M, v = V(a[2]), V(a[3])
            out = [F32(M[0][r] * v[0]) for r in range(len(M[0]))]
            for c in range(1, len(M)):
                out = [F32(out[r] + F32(M[c][r] * v[c])) for r in range(len(out))]
            env[a[1]] = out
        elif op == 146:
            A, B = V(a[2]), V(a[3])
            cols = []
            for bc in B:
                out = [F32(A[0][r] * bc[0]) for r in range(len(A[0]))]
                for c in range(1, len(A)):
                    out = [F32(out[r] + F32(A[c][r] * bc[c])) for r in range(len(out))]
                cols.append(out)
            env[a[1]] = cols
        elif op == 84:                                 # OpTranspose
            M = V(a[2])
            env[a[1]] = [[M[c][r] for c in range(len(M))] for r in range(len(M[0]))]
        elif op == 148:
            env[a[1]] = self._dot(V(a[2]), V(a[3]))
        elif op in (154, 155):                         # OpAny / OpAll
            env[a[1]] = (any if op == 154 else all)(V(a[2]))
        elif op == 156:
            env[a[1]] = mp(lambda x: bool(np.isnan(x)), V(a[2]))
        elif op == 157:
            env[a[1]] = mp(lambda x: bool(np.isinf(x)), V(a[2]))
        elif op in (164, 165, 166):                    # LogicalEqual / LogicalNotEqual / LogicalOr
            f = {164: lambda x, y: x == y, 165: lambda x, y: x != y, 166: lambda x, y: x or y}[op]
            env[a[1]] = mp(f, V(a[2]), V(a[3]))
        elif op == 167:
            env[a[1]] = mp(lambda x, y: x and y, V(a[2]), V(a[3]))
        elif op == 168:
            env[a[1]] = mp(lambda x: not x, V(a[2]))
        elif op == 169:
            c, x, y = V(a[2]), V(a[3]), V(a[4])
            env[a[1]] = [(xi if ci else yi) for ci, xi, yi in zip(c, x, y)] if isinstance(c, list) else (_copy(x) if c else _copy(y))
        elif op in (170, 171):
            env[a[1]] = mp((lambda x, y: x == y) if op == 170 else (lambda x, y: x != y), V(a[2]), V(a[3]))
        elif op in (172, 174, 176, 178):               # unsigned >, >=, <, <=
            f = {172: lambda x, y: x > y, 174: lambda x, y: x >= y, 176: lambda x, y: x < y, 178: lambda x, y: x <= y}[op]
            env[a[1]] = mp(lambda x, y: f(x & MASK, y & MASK), V(a[2]), V(a[3]))
        elif op in (173, 175, 177, 179):               # signed
            f = {173: lambda x, y: x > y, 175: lambda x, y: x >= y, 177: lambda x, y: x < y, 179: lambda x, y: x <= y}[op]
            env[a[1]] = mp(lambda x, y: f(_s32(x), _s32(y)), V(a[2]), V(a[3]))
        elif 180 <= op <= 191:                         # float comparisons, ordered (even) / unordered (odd)
            base = {180: "eq", 182: "ne", 184: "lt", 186: "gt", 188: "le", 190: "ge"}[op - (op & 1)]
            unordered = bool(op & 1)
            def cmp(x, y):
                if np.isnan(x) or np.isnan(y):
                    return unordered
                return {"eq": x == y, "ne": x != y, "lt": x < y, "gt": x > y, "le": x <= y, "ge": x >= y}[base]
            env[a[1]] = mp(lambda x, y: bool(cmp(x, y)), V(a[2]), V(a[3]))
        elif op == 194:
            env[a[1]] = mp(lambda x, y: ((x & MASK) >> (y & 31)) & MASK, V(a[2]), V(a[3]))
        elif op == 195:
            env[a[1]] = mp(lambda x, y: (_s32(x) >> (y & 31)) & MASK, V(a[2]), V(a[3]))
        elif op == 196:
            env[a[1]] = mp(lambda x, y: (x << (y & 31)) & MASK, V(a[2]), V(a[3]))
        elif op in (197, 198, 199):
            f = {197: lambda x, y: x | y, 198: lambda x, y: x ^ y, 199: lambda x, y: x & y}[op]
            env[a[1]] = mp(lambda x, y: f(x, y) & MASK, V(a[2]), V(a[3]))
        elif op == 200:
            env[a[1]] = mp(lambda x: (~x) & MASK, V(a[2]))
        elif op in self.ext.get("ops", {}):
            self.ext["ops"][op](self, env, a)
        else:
            raise NotImplementedError("SPIR-V opcode %d" % op)

    @staticmethod
    def _dot(a, b):
        s = F32(a[0] * b[0])
        for x, y in zip(a[1:], b[1:]):
            s = F32(s + F32(x * y))
        return s

    def ext_inst(self, rtype, inst, x):
        mp, M = self._map, self._math
        with np.errstate(all="ignore"):
            if inst == 4:
                return mp(lambda v: F32(abs(v)), x[0])
            if inst == 5:
                return mp(lambda v: abs(_s32(v)) & MASK, x[0])
            if inst == 6:
                return mp(lambda v: F32(np.sign(v)), x[0])
            if inst == 8:
                return mp(lambda v: F32(np.floor(v)), x[0])
            if inst == 9:
                return mp(lambda v: F32(np.ceil(v)), x[0])
            if inst == 10:
                return mp(lambda v: F32(v - F32(np.floor(v))), x[0])
            if inst == 3:
                return mp(lambda v: F32(np.trunc(v)), x[0])
            if inst in (1, 2):
                return mp(lambda v: F32(np.rint(v)), x[0])
            if inst == 11:
                return mp(lambda v: F32(v * F32(0.017453292519943295)), x[0])
            if inst == 13:
                return mp(M("sin", lambda v: F32(np.sin(np.float64(v)))), x[0])
            if inst == 14:
                return mp(M("cos", lambda v: F32(np.cos(np.float64(v)))), x[0])
            if inst == 17:
                return mp(M("acos", lambda v: F32(np.arccos(np.float64(v)))), x[0])
            if inst == 26:
                return mp(M("pow", lambda v, w: F32(np.power(np.float64(v), np.float64(w)))), x[0], x[1])
            if inst == 27:
                return mp(M("exp", lambda v: F32(np.exp(np.float64(v)))), x[0])
            if inst == 28:
                return mp(M("log", lambda v: F32(np.log(np.float64(v)))), x[0])
            if inst == 29:
                return mp(M("exp2", lambda v: F32(np.exp2(np.float64(v)))), x[0])
            if inst == 31:
                return mp(lambda v: F32(np.sqrt(v)), x[0])
            if inst == 32:
                return mp(M("inversesqrt", lambda v: F32(F32(1) / F32(np.sqrt(v)))), x[0])
            if inst in (37, 40):
                f = np.minimum if inst == 37 else np.maximum
                return mp(lambda v, w: F32(f(v, w)), x[0], x[1])
            if inst in (38, 41):
                return mp((lambda v, w: min(v & MASK, w & MASK)) if inst == 38 else (lambda v, w: max(v & MASK, w & MASK)), x[0], x[1])
            if inst in (39, 42):
                f = min if inst == 39 else max
                return mp(lambda v, w: f(_s32(v), _s32(w)) & MASK, x[0], x[1])
            if inst == 43:
                return mp(lambda v, lo, hi: F32(np.minimum(np.maximum(v, lo), hi)), x[0], x[1], x[2])
            if inst == 44:
                return mp(lambda v, lo, hi: min(max(v & MASK, lo & MASK), hi & MASK), x[0], x[1], x[2])
            if inst == 45:
                return mp(lambda v, lo, hi: min(max(_s32(v), _s32(lo)), _s32(hi)) & MASK, x[0], x[1], x[2])
            if inst == 46:       # mix(a, b, t) = a * (1 - t) + b * t
                return mp(lambda p, q, t: F32(F32(p * F32(F32(1) - t)) + F32(q * t)), x[0], x[1], x[2])
            if inst == 48:
                return mp(lambda edge, v: F32(0.0) if v < edge else F32(1.0), x[0], x[1])
            if inst == 50:
                return mp(lambda p, q, r: F32(np.float64(p) * np.float64(q) + np.float64(r)), x[0], x[1], x[2])
            if inst == 66:
                v = x[0] if isinstance(x[0], list) else [x[0]]
                return M("length", lambda vv: F32(np.sqrt(self._dot(vv, vv))))(v)
            if inst == 67:
                d = [F32(p - q) for p, q in zip(x[0], x[1])]
                return M("length", lambda vv: F32(np.sqrt(self._dot(vv, vv))))(d)
            if inst == 68:
                p, q = x
                return [F32(F32(p[1] * q[2]) - F32(p[2] * q[1])), F32(F32(p[2] * q[0]) - F32(p[0] * q[2])), F32(F32(p[0] * q[1]) - F32(p[1] * q[0]))]
            if inst == 69:
                def normalize(vv):
                    inv = F32(F32(1) / F32(np.sqrt(self._dot(vv, vv))))
                    return [F32(c * inv) for c in vv]
                return M("normalize", normalize)(x[0])
            if inst == 70:       # faceforward(N, I, Nref)
                return list(x[0]) if self._dot(x[2], x[1]) < 0 else [F32(-c) for c in x[0]]
            if inst == 71:       # reflect(I, N) = I - 2 dot(N, I) N
                d2 = F32(F32(2) * self._dot(x[1], x[0]))
                return [F32(i - F32(n * d2)) for i, n in zip(x[0], x[1])]
            if inst == 72:
                I, N, eta = x
                ni = self._dot(N, I)
                k = F32(F32(1) - F32(F32(eta * eta) * F32(F32(1) - F32(ni * ni))))
                if k < 0:
                    return [F32(0)] * len(I)
                f = F32(F32(eta * ni) + F32(np.sqrt(k)))
                return [F32(F32(i * eta) - F32(n * f)) for i, n in zip(I, N)]
            if inst == 34:       # MatrixInverse (3x3 by cofactors, columns)
                return M("inverse", None)(x[0])
        raise NotImplementedError("GLSL.std.450 instruction %d" % inst)


class Image:
    """Storage image over a (H, W, 4) float32 array; reads outside return zeros, writes outside are dropped."""

    def __init__(self, array):
        self.a = array

    def size(self):
        return [self.a.shape[1], self.a.shape[0]]

    def read(self, c):
        x, y = c[0], c[1]
        if 0 <= x < self.a.shape[1] and 0 <= y < self.a.shape[0]:
            return [F32(v) for v in self.a[y, x]]
        return [F32(0)] * 4

    def write(self, c, texel):
        x, y = c[0], c[1]
        if 0 <= x < self.a.shape[1] and 0 <= y < self.a.shape[0]:
            self.a[y, x] = [np.float32(v) for v in texel]
