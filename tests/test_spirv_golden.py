"""Parity pinned to the reference's COMPILED shaders: tests/golden/spirv_post.npz holds what the reference's own
SPIR-V binaries (blurX / blurY / combine / tonemapping .comp.spv) compute for a 48x36 HDR frame, executed by
tests/spirv_interp.py (see tests/golden/make_spirv_golden.py). SURVEY.md 8a rows a17, a18, a19.
  * CPU: the oracle's post-processing equals the fixture bit for bit (combined HDR image) and byte for byte (LDR);
  * CPU, build container only: the reference's .spv files are re-executed on a smaller frame and compared with the oracle;
  * GPU: the CUDA post-processing kernels, through the C ABI, produce the fixture's LDR frame."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "spirv_post.npz")
REF = "/root/reference/shaders/postprocessing/"


def _cases():
    g = np.load(GOLD)
    for i, (radius, threshold, intensity, exposure) in enumerate(g["params"]):
        yield i, g, float(radius), float(threshold), float(intensity), float(exposure)


def test_oracle_postprocessing_equals_the_reference_spirv(ol, rb):
    n = 0
    for i, g, radius, threshold, intensity, exposure in _cases():
        ldr, comb = ol.postprocess(g["hdr"], bloom=rb.abi.BloomPushConsts(radius, threshold, intensity),
                                   tonemap=rb.abi.TonemappingPushConsts(exposure), want_combined=True)
        want = g["combined_%d" % i]
        assert (comb[..., :3].view(np.uint32) == want[..., :3].view(np.uint32)).all(), i      # bloom + combine: same bits
        assert (ldr[..., :3] == g["ldr_%d" % i][..., :3]).all(), i                               # tonemap + UNORM8 store
        assert (ldr[..., 3] == 255).all() and (g["ldr_%d" % i][..., 3] == 255).all()
        n += 1
    assert n == 3


def test_fixture_is_not_trivial():
    for i, g, radius, threshold, intensity, exposure in _cases():
        assert np.abs(g["combined_%d" % i][..., :3] - g["hdr"][..., :3]).max() > 0.03         # bloom moved energy
        assert np.abs(g["blur_y_%d" % i][..., :3]).max() > 0 and g["ldr_%d" % i][..., :3].std() > 10
        # below-threshold pixels never enter the blur: the darkest region of the blurred image is exactly zero
        assert (g["blur_x_%d" % i][..., :3] == 0).any()


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present")
def test_reexecuting_the_reference_spirv_matches_the_oracle(ol, rb):
    """Run the reference's binaries again, on a different (smaller) frame than the fixture's."""
    import golden.make_spirv_golden as mk
    rng = np.random.RandomState(11)
    hdr = np.zeros((14, 20, 4), np.float32)
    hdr[..., :3] = (rng.uniform(0, 1.2, (14, 20, 3)) ** 2).astype(np.float32)
    hdr[3, 5, :3] = (30, 2, 9)
    hdr[..., 3] = 1
    case = dict(radius=6.0, threshold=0.8, intensity=0.4, exposure=0.5)
    r = mk.run_chain(hdr, case)
    ldr, comb = ol.postprocess(hdr, bloom=rb.abi.BloomPushConsts(6.0, 0.8, 0.4), tonemap=rb.abi.TonemappingPushConsts(0.5),
                               want_combined=True)
    assert (comb[..., :3].view(np.uint32) == r["combined"][..., :3].view(np.uint32)).all()
    assert (ldr[..., :3] == r["ldr"][..., :3]).all()


@pytest.mark.gpu
def test_cuda_postprocessing_equals_the_reference_spirv(rb):
    g = np.load(GOLD)
    hdr = np.ascontiguousarray(g["hdr"], np.float32)
    h, w = hdr.shape[:2]
    wl = rb.configs.cornell(w, h)
    r = rb.Renderer(w, h, wl.tables)
    for i, (radius, threshold, intensity, exposure) in enumerate(g["params"]):
        r.write_hdr(hdr)
        r.postprocess(bloom=rb.abi.BloomPushConsts(float(radius), float(threshold), float(intensity)),
                      tonemap=rb.abi.TonemappingPushConsts(float(exposure)))
        got = r.read_ldr()
        assert (got[..., :3] == g["ldr_%d" % i][..., :3]).all(), i
    r.close()
