"""Several GPUs behind the C ABI (rb200_group_*): NCCL is called from inside librb200.so. A group of
one device runs everywhere; the two-device tests need a box with at least two GPUs (gpurun --gpus 2) and are skipped
otherwise."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def device_count():
    import torch
    return torch.cuda.device_count()


def sum_image(rb, wl, batches, device=0):
    """SUM image (rb200 sum mode) of the given batches rendered by one context, in the given order."""
    r = rb.Renderer(wl.width, wl.height, wl.tables, flags=rb.RB200_FLAG_NEE | rb.RB200_FLAG_ACCUM_SUM, device=device)
    for b in batches:
        r.render_batch(wl.push_constants(b))
    img = r.read_hdr()
    r.close()
    return img


def peer_access(a, b):
    import torch
    return torch.cuda.can_device_access_peer(b, a)


def test_group_of_one_device_is_the_sum_mode_renderer(rb):
    wl = rb.configs.small_mixed(96, 72, nee=True, samples_per_pixel=2, max_bounces=6)
    g = rb.Group(wl.width, wl.height, wl.tables, [0], flags=rb.RB200_FLAG_NEE)
    assert g.size() == 1
    g.render_batches(wl.push_constants(0), 0, 3)
    g.render_batches(wl.push_constants(0), 3, 2)
    got = g.read_hdr()
    g.present()
    ldr = g.read_ldr()
    st = g.stats()
    g.close()
    want = sum_image(rb, wl, range(5))
    inv = np.float32(1.0) / np.float32(5)
    mean = want * inv
    mean[..., 3] = 1.0
    assert (bits(got) == bits(mean)).all()
    r = rb.Renderer(wl.width, wl.height, wl.tables, flags=rb.RB200_FLAG_NEE | rb.RB200_FLAG_ACCUM_SUM)
    for b in range(5):
        r.render_batch(wl.push_constants(b))
    r.present_sum(5)
    assert (r.read_ldr() == ldr).all()
    _, cum = r.stats()
    assert (st["extendRays"], st["shadowRays"], st["paths"]) == (cum["extendRays"], cum["shadowRays"], cum["paths"])
    r.close()


def test_group_rejects_bad_device_lists(rb):
    wl = rb.configs.small_mixed(32, 24, nee=True)
    with pytest.raises(rb.RB200Error, match="listed twice"):
        rb.Group(wl.width, wl.height, wl.tables, [0, 0], flags=rb.RB200_FLAG_NEE)
    with pytest.raises(rb.RB200Error):
        rb.Group(wl.width, wl.height, wl.tables, [device_count() + 3], flags=rb.RB200_FLAG_NEE)


@pytest.mark.skipif("device_count() < 2")
def test_two_devices_sample_split_reduces_exactly_the_two_partial_sums(ol, rb):
    """Device 0 renders batches 0, 2, 4, device 1 batches 1, 3, 5; the library's ncclReduce adds the two SUM images.
    The mean image it returns is (sum0 + sum1) / 6 bit for bit, with sum0 / sum1 rendered by single contexts — and it
    agrees with the oracle's running average of the six batches up to the fp32 summation order."""
    wl = rb.configs.small_mixed(96, 72, nee=True, samples_per_pixel=2, max_bounces=6)
    g = rb.Group(wl.width, wl.height, wl.tables, [0, 1], flags=rb.RB200_FLAG_NEE)
    assert g.bvh_info(0)["hash"] == g.bvh_info(1)["hash"]
    g.render_batches(wl.push_constants(0), 0, 2)
    g.render_batches(wl.push_constants(0), 4, 1)
    got = g.read_hdr()
    g.present()
    ldr = g.read_ldr()
    g.close()
    s0, s1 = sum_image(rb, wl, [0, 2, 4], device=0), sum_image(rb, wl, [1, 3, 5], device=1)
    want = (s0 + s1) * (np.float32(1.0) / np.float32(6))
    want[..., 3] = 1.0
    assert (bits(got) == bits(want)).all()
    sc = ol.OracleScene(wl.tables)
    hdr_o = np.zeros((wl.height, wl.width, 4), np.float32)
    for b in range(6):
        hdr_o, _ = sc.render_batch(wl.width, wl.height, rb.RB200_FLAG_NEE, wl.push_constants(b), hdr_o)
    sc.close()
    rel = np.abs(got[..., :3] - hdr_o[..., :3]) / np.maximum(np.abs(hdr_o[..., :3]), 1e-3)
    assert rel.max() < 1e-5      # tolerance: six fp32 additions in another order
    assert (np.abs(ldr.astype(int) - ol.postprocess(hdr_o).astype(int)) <= 1).all()


@pytest.mark.skipif("device_count() < 2")
@pytest.mark.parametrize("reduce", [False, True])
def test_two_devices_interleaved_tiles_give_the_single_gpu_image_bit_for_bit(ol, rb, reduce, monkeypatch):
    """Latency mode: every device traces its 16 x 16 tiles of every batch (running average). With peer access the devices
    store their pixels straight into device 0's image from k_accumulate (no collective); with RB200_GROUP_TILES_REDUCE=1
    the images — zero outside a device's tiles — are added by one ncclReduce per frame. Either way every presented frame
    (one per batch: the host-side hand-shake of the peer mode runs four times) is the single-GPU frame, HDR bit for bit
    and LDR byte for byte."""
    if reduce:
        monkeypatch.setenv("RB200_GROUP_TILES_REDUCE", "1")
    else:
        monkeypatch.delenv("RB200_GROUP_TILES_REDUCE", raising=False)
    wl = rb.configs.small_mixed(96, 72, nee=True, samples_per_pixel=2, max_bounces=6)
    g = rb.Group(wl.width, wl.height, wl.tables, [0, 1], flags=rb.RB200_FLAG_NEE, tiles=True)
    assert g.uses_peer_stores() == (False if reduce else peer_access(0, 1))
    rb.abi.check(g.lib, g.lib.rb200_group_set_tile_size(g._g, 16))
    r = rb.Renderer(wl.width, wl.height, wl.tables, flags=rb.RB200_FLAG_NEE)
    for b in range(4):
        g.render_batches(wl.push_constants(0), b, 1)
        g.present()
        ldr = g.read_ldr()
        r.render_batch(wl.push_constants(b))
        r.postprocess()
        assert (ldr == r.read_ldr()).all(), "frame %d" % b
    got = g.read_hdr()
    g.render_batches(wl.push_constants(0), 4, 2)        # two batches behind one present
    g.present()
    ldr = g.read_ldr()
    got2 = g.read_hdr()
    g.close()
    want = r.read_hdr()
    for b in (4, 5):
        r.render_batch(wl.push_constants(b))
    r.postprocess()
    want_ldr, want2 = r.read_ldr(), r.read_hdr()
    r.close()
    assert (bits(got[..., :3]) == bits(want[..., :3])).all()
    assert (bits(got2[..., :3]) == bits(want2[..., :3])).all()
    assert (ldr == want_ldr).all()
