"""Golden vectors minted from the reference's own COMPILED shaders (python tests/golden/make_spirv_golden.py).

/root/reference ships the SPIR-V binaries of every shader next to its GLSL source. No Vulkan driver, glslang or
spirv-cross exists in this image, so tests/spirv_interp.py interprets those binaries on the CPU, one invocation at a
time, with every fp32 operation rounded to fp32 and nothing contracted or reassociated. The only thing the modules
leave open is the precision of GLSL.std.450 Exp / Exp2 (implementation-defined in GLSL); for those the repository's
elementary layer (csrc/rb_math.h, through the oracle library) is plugged in. Everything else — constants, operation
order, loop bounds, the threshold / weight-sum rules, the ACES matrices as the compiler laid them out — is the
reference's compiled code, not a reading of its source.

spirv_post.npz: a 48x36 HDR frame with bright spots through
    shaders/postprocessing/bloom/blurX.comp.spv -> blurY.comp.spv -> combine.comp.spv -> tonemap/tonemapping.comp.spv
for three parameter sets; stored: the input, the push constants, and every intermediate image (fp32) plus the final
colour before and after the RGBA8 UNORM store (round to nearest, ties to even — the one step the image store does, not
the shader).
The fixture travels to the GPU box (the reference does not): tests/test_spirv_golden.py checks the oracle against it
on the CPU and the CUDA path against it on the GPU."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
REF = "/root/reference/shaders/postprocessing/"
CASES = [dict(radius=5.0, threshold=1.0, intensity=0.05, exposure=1.0),       # config/config.toml defaults
         dict(radius=3.0, threshold=0.5, intensity=0.3, exposure=0.0),
         dict(radius=8.0, threshold=2.5, intensity=1.0, exposure=-1.25)]
W, H = 48, 36


def ext_math():
    import oracle_lib as ol
    F = np.float32
    one = lambda fn: (lambda v: F(ol.rb_math(fn, np.array([v], np.float32))[0]))
    return {"sin": one(0), "cos": one(1), "log": one(2), "exp": one(3), "exp2": one(4), "acos": one(5)}


def input_frame():
    rng = np.random.RandomState(3)
    hdr = np.zeros((H, W, 4), np.float32)
    hdr[..., :3] = (rng.uniform(0, 0.9, (H, W, 3)) ** 3).astype(np.float32)
    hdr[..., 3] = 1
    for _ in range(12):
        y, x = rng.randint(0, H), rng.randint(0, W)
        hdr[y, x, :3] = rng.uniform(2, 40, 3)
    return hdr


def run_chain(hdr, case, shader_dir=REF):
    """The reference's four compute shaders over the whole frame, in the order Reina::applyBloom / applyTonemapping
    dispatch them (src/Reina.cpp:472-577)."""
    import spirv_interp as sp
    F = np.float32
    ext = ext_math()
    load = lambda f: sp.Interpreter(sp.Module(os.path.join(shader_dir, f)), ext)
    bx, by, cb, tm = (load(f) for f in ("bloom/blurX.comp.spv", "bloom/blurY.comp.spv", "bloom/combine.comp.spv",
                                        "tonemap/tonemapping.comp.spv"))
    h, w = hdr.shape[:2]
    pc = [[F(case["radius"]), F(case["threshold"]), F(case["intensity"])]]       # BloomPushConsts, polyglot/bloom.h
    ping, pong, comb, ldrf = (np.zeros_like(hdr) for _ in range(4))

    def dispatch(it, bind):
        for y in range(h):
            for x in range(w):
                it.run(bind, builtins={28: [x, y, 0]})                            # gl_GlobalInvocationID
    dispatch(bx, {(0, 0): sp.Image(hdr), (0, 1): sp.Image(ping), "push_constant": pc})
    dispatch(by, {(0, 0): sp.Image(ping), (0, 1): sp.Image(pong), "push_constant": pc})
    dispatch(cb, {(0, 0): sp.Image(hdr), (0, 1): sp.Image(pong), (0, 2): sp.Image(comb), "push_constant": pc})
    dispatch(tm, {(0, 0): sp.Image(comb), (0, 1): sp.Image(ldrf), "push_constant": [[F(case["exposure"])]]})
    ldr = np.rint(np.clip(ldrf, 0, 1).astype(np.float64) * 255).astype(np.uint8)  # VK_FORMAT_R8G8B8A8_UNORM image store
    return dict(blur_x=ping, blur_y=pong, combined=comb, ldr_float=ldrf, ldr=ldr)


def main():
    hdr = input_frame()
    out = {"hdr": hdr, "params": np.array([[c["radius"], c["threshold"], c["intensity"], c["exposure"]] for c in CASES], np.float32)}
    for i, c in enumerate(CASES):
        r = run_chain(hdr, c)
        for k, v in r.items():
            out["%s_%d" % (k, i)] = v
        print("case", i, c, "bloom energy", float(np.abs(r["combined"][..., :3] - hdr[..., :3]).max()), "ldr mean", float(r["ldr"][..., :3].mean()))
    np.savez_compressed(os.path.join(HERE, "spirv_post.npz"), **out)


if __name__ == "__main__":
    main()
