"""The kernels' pure shading functions (reina-vk_b200/csrc/shade.cuh: disney_sample, disney_eval, fuzzy_reflection,
random_emissive_point) compiled for the HOST by g++ next to the oracle's independent restatement (oracle/disney.h,
oracle/pathtrace.cpp) and compared bit for bit on 200,000 random inputs and 20,000 light samples of a three-emitter
scene (tools/host_shade.cpp). The GPU parity tests compare images; this one compares the two
source texts directly, without a GPU, so an edit to one side that is not mirrored on the other fails here first."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CUDA_INC = "/usr/local/cuda/include"


@pytest.mark.skipif(not os.path.exists(os.path.join(CUDA_INC, "cuda_runtime.h")), reason="CUDA headers not present")
def test_shade_cuh_matches_the_oracle_on_the_host(tmp_path):
    exe = str(tmp_path / "host_shade")
    err = ""
    for cxx in ("/usr/bin/g++", shutil.which("g++"), os.environ.get("CXX")):
        if not cxx or not os.path.exists(cxx):
            continue
        r = subprocess.run([cxx, "-O1", "-std=c++17", "-ffp-contract=off", "-mfma", "-w", "-I", CUDA_INC,
                            "-I", os.path.join(ROOT, "reina-vk_b200", "csrc"), "-I", os.path.join(ROOT, "oracle"),
                            "-o", exe, os.path.join(ROOT, "tools", "host_shade.cpp"), os.path.join(ROOT, "oracle", "pathtrace.cpp"),
                            os.path.join(ROOT, "oracle", "intersect.cpp"), "-pthread"], capture_output=True, text=True)
        if r.returncode == 0:
            break
        err = r.stderr[-500:]
    else:
        pytest.skip("host build of shade.cuh not possible: " + err)
    run = subprocess.run([exe], capture_output=True, text=True)
    assert run.returncode == 0, run.stdout + run.stderr
    assert "disney_sample mismatches 0  disney_eval mismatches 0  fuzzy_reflection state not restored 0" in run.stdout
    assert "random_emissive_point mismatches 0" in run.stdout
