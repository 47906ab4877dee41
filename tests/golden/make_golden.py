"""Writes the golden fixtures in this directory from the CPU oracle (python tests/golden/make_golden.py).

The reference ships no tests, golden vectors or fixtures and cannot be run here, so these files
do not come from the reference: they freeze the oracle's output for four small cases (every material + textures +
alpha skip + NEE; the as-shipped NEE-off estimator; the Cornell config C1 at reduced size; every material with a
parallax height map) so that any later change to
the oracle or to the shared elementary layer shows up as a diff, and so that the GPU parity test has a fixed target
that does not depend on executing the oracle.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))


def _small_nee(rb):
    return rb.configs.small_mixed(64, 48, nee=True, samples_per_pixel=2, max_bounces=6), rb.RB200_FLAG_NEE, 2


def _small_shipped(rb):
    return rb.configs.small_mixed(64, 48, nee=False, samples_per_pixel=2, max_bounces=6), 0, 2


def _cornell(rb):
    return rb.configs.cornell(80, 60, nee=True, samples_per_pixel=4, max_bounces=8), rb.RB200_FLAG_NEE, 1


def _parallax(rb):
    return rb.configs.parallax(64, 48, nee=True, samples_per_pixel=2, max_bounces=6), rb.RB200_FLAG_NEE, 2


CASES = {"small_mixed_nee": _small_nee, "small_mixed_shipped": _small_shipped, "cornell_nee": _cornell,
         "parallax_nee": _parallax}


def main():
    import oracle_lib as ol
    rb = ol.rb
    only = sys.argv[1:]            # python make_golden.py [case ...]: regenerate just these
    for name, make in CASES.items():
        if only and name not in only:
            continue
        wl, flags, batches = make(rb)
        sc = ol.OracleScene(wl.tables)
        hdr = np.zeros((wl.height, wl.width, 4), np.float32)
        for b in range(batches):
            hdr, cnt = sc.render_batch(wl.width, wl.height, flags, wl.push_constants(b), hdr)
        hits = sc.trace_primary(wl.width, wl.height, wl.push_constants(0))
        np.savez_compressed(os.path.join(HERE, name + ".npz"), hdr=hdr, ldr=ol.postprocess(hdr), prim=hits["primitive"],
                            inst=hits["instance"], t=hits["t"], extend=cnt["extendRays"], shadow=cnt["shadowRays"])
        print(name, hdr.shape, "mean", float(hdr[..., :3].mean()), cnt)


if __name__ == "__main__":
    main()
