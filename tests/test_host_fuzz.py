"""Robustness of the C++ host's file parsers (PNG inflate / defilter, baseline and progressive JPEG, JSON + glTF / GLB, OBJ, the TOML configuration): mutated inputs
must be either decoded or refused with an exception — never a crash, an out-of-bounds access or undefined behaviour.
The harness (tools/fuzz_host.cpp) is built with AddressSanitizer + UndefinedBehaviorSanitizer and fed ~2000 mutations
(byte flips, truncations, zeroed and inserted runs, 0xFFFFFFFF words) of valid files."""
import os
import pathlib
import shutil
import subprocess

import numpy as np
import pytest

import gltf_fixture as gf

ROOT = pathlib.Path(__file__).resolve().parent.parent
HOST = ROOT / "reina-vk_b200" / "host"


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    out = tmp_path_factory.mktemp("fuzz") / "fuzz_host"
    src = [str(ROOT / "tools" / "fuzz_host.cpp")] + [str(HOST / f) for f in ("model.cpp", "scene.cpp", "texture.cpp", "jpeg.cpp", "gltf.cpp", "png.cpp", "config.cpp")]
    err = ""
    for cxx in ("/usr/bin/g++", shutil.which("g++"), os.environ.get("CXX"), shutil.which("clang++")):     # the first with sanitizer runtimes
        if not cxx or not os.path.exists(cxx):
            continue
        r = subprocess.run([cxx, "-O1", "-g", "-std=c++17", "-fsanitize=address,undefined", "-fno-omit-frame-pointer", "-I", str(HOST), "-o", str(out)] + src + ["-lz"],
                           capture_output=True, text=True)
        if r.returncode == 0:
            return str(out)
        err = r.stderr[-300:]
    pytest.skip("sanitizer build not available: " + err)


def _mutations(raw, rng, count):
    for it in range(count):
        b = bytearray(raw)
        mode = it % 5
        if mode == 0:
            for _ in range(int(rng.integers(1, 6))):
                b[int(rng.integers(0, len(b)))] = int(rng.integers(0, 256))
        elif mode == 1:
            b = b[: int(rng.integers(0, len(b)))]
        elif mode == 2:
            i = int(rng.integers(0, len(b)))
            b[i: i + int(rng.integers(1, 32))] = bytes(int(rng.integers(1, 32)))
        elif mode == 3:
            i = int(rng.integers(0, len(b)))
            b = b[:i] + bytes(rng.integers(0, 256, int(rng.integers(1, 64)), dtype=np.uint8)) + b[i:]
        else:
            i = int(rng.integers(0, max(1, len(b) - 4)))
            b[i: i + 4] = b"\xff\xff\xff\xff"
        yield bytes(b)


def test_mutated_files_never_crash_the_parsers(harness, tmp_path):
    from PIL import Image
    rng = np.random.default_rng(7)
    pic = rng.integers(0, 256, (40, 56, 3), dtype=np.uint8)
    d = tmp_path
    Image.fromarray(pic).save(d / "a.jpg", quality=80, subsampling=2)
    Image.fromarray(pic).save(d / "b.jpg", quality=60, subsampling=1, restart_marker_blocks=2)
    Image.fromarray(pic).save(d / "p.jpg", quality=70, subsampling=2, progressive=True)
    Image.fromarray(pic).save(d / "q.jpg", quality=90, subsampling=0, progressive=True, restart_marker_blocks=3)
    Image.fromarray(pic).save(d / "a.png")
    Image.fromarray(pic[..., 0]).save(d / "c.png")
    Image.fromarray(pic).convert("P").save(d / "d.png")
    glb, _ = gf.build(d, sparse=True)
    (d / "ext").mkdir()
    gltf, _ = gf.build(d / "ext", external=True, jpeg=True)
    from test_cpp_host import OBJ_FULL, REFERENCE_SCHEMA, _write_png
    _write_png(d / "i.png", pic[:19, :21].astype(np.uint32), 2, 8, 1, rng)                        # Adam7
    _write_png(d / "j.png", (pic[:9, :13, :1] >> 6).astype(np.uint32), 0, 2, 1, rng)             # Adam7, 2-bit grey
    (d / "a.obj").write_text(OBJ_FULL)
    (d / "a.toml").write_text(REFERENCE_SCHEMA + "\n[render]\nwidth = 80\nheight = 60\ncamera_pos = [0.4, 0.9, 2.6]\nscene = \"cornell\"\n")
    seeds = [d / "a.jpg", d / "b.jpg", d / "p.jpg", d / "q.jpg", d / "a.png", d / "c.png", d / "d.png", d / "i.png", d / "j.png", pathlib.Path(glb), pathlib.Path(gltf), d / "a.obj", d / "a.toml"]
    files = [str(s) for s in seeds]                       # the valid files themselves must decode
    for s in seeds:
        raw = s.read_bytes()
        for k, m in enumerate(_mutations(raw, rng, 220)):
            out = s.parent / ("m%03d_%s" % (k, s.name))
            out.write_bytes(m)
            files.append(str(out))
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=0:abort_on_error=0", UBSAN_OPTIONS="print_stacktrace=1")
    r = subprocess.run([harness] + files, capture_output=True, text=True, env=env)
    noise = [l for l in r.stderr.splitlines() if "falling back to UV" not in l and l.strip()]
    assert r.returncode == 0, r.stderr[-2000:]
    assert not [l for l in noise if "runtime error" in l or "AddressSanitizer" in l], "\n".join(noise[:20])
    ok, refused = (int(x) for x in r.stdout.split()[1::2])
    assert ok >= len(seeds) and ok + refused == len(files)
