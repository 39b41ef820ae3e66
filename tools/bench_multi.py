#!/usr/bin/env python
"""One host process driving every GPU of the box through sage_b200_score_batch_multi (what the Rust shim calls from `search_processed_spectra`):
one index replica + one scorer per device, the batch cut into contiguous blocks, one NUMA-bound host thread per device inside the library.

    python tools/bench_multi.py [--gpus N] [--spectra-per-gpu 50000] [--steps 10]

Prints one JSON line: e2e spectra/s (host buffers in, Feature rows out, wall clock around the C-ABI call) for N = 1, 2, 4, ... visible GPUs, each
with the rows of GPU 0's block checked against the single-GPU result of the same spectra (bitwise).
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=0)
    ap.add_argument("--spectra-per-gpu", type=int, default=50_000)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--peptides", type=int, default=2_000_000)
    ap.add_argument("--only", type=int, default=0, help="measure this GPU count only (plus 1 as the base)")
    ap.add_argument("--numa-blocks", type=int, default=1, help="1: batch buffers from sage_b200_host_alloc_blocks (block g next to GPU g); 0: one pinned buffer")
    args = ap.parse_args()
    from sage_b200 import IndexedDatabase, Scorer, SpectraBatch, Tolerance, api, synth
    ndev = api.device_count()
    gmax = args.gpus or ndev
    assert 1 <= gmax <= ndev
    pep = synth.make_peptides(args.peptides)
    spectra = synth.make_spectra(pep, args.spectra_per_gpu * gmax, seed=0xB202)
    kw = dict(precursor_tol=Tolerance.ppm(-20, 20), fragment_tol=Tolerance.ppm(-20, 20))
    dbs = [IndexedDatabase.build_from_peptides(pep, device=g) for g in range(gmax)]
    scorers = [Scorer(dbs[g], **kw) for g in range(gmax)]

    def pin(a):
        p = api.pinned_empty(a.shape, a.dtype)
        p[...] = a
        return p
    ref_f, ref_c = scorers[0].score_batch(spectra.slice(0, args.spectra_per_gpu))
    out = {"metric": "spectra/sec", "mode": "one host process, sage_b200_score_batch_multi", "spectra_per_gpu": args.spectra_per_gpu, "steps": args.steps, "runs": []}
    n = 1
    while n <= gmax:
        if args.only and n not in (1, args.only):
            n *= 2
            continue
        sub = spectra.slice(0, args.spectra_per_gpu * n)
        devs = list(range(n))
        if args.numa_blocks:   # block g of every batch array on the NUMA node next to GPU g
            def pin(a):
                p = api.pinned_empty_blocks(a.shape, a.dtype, devs)
                p[...] = a
                return p
            f = api.pinned_empty_blocks((len(sub),), api.FEATURE_DTYPE, devs)
            c = api.pinned_empty_blocks((len(sub),), np.uint32, devs)
        else:
            f = api.pinned_empty((len(sub),), api.FEATURE_DTYPE)
            c = api.pinned_empty((len(sub),), np.uint32)
        hs = SpectraBatch(**{**sub.__dict__, "masses": pin(sub.masses), "intensities": pin(sub.intensities)})
        for _ in range(3):
            api.score_batch_multi(scorers[:n], hs, f, c)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            api.score_batch_multi(scorers[:n], hs, f, c)
        dt = (time.perf_counter() - t0) / args.steps
        m = args.spectra_per_gpu
        same = bool(np.array_equal(np.array(c[:m]), ref_c) and np.array(f[:m]).tobytes() == ref_f.tobytes())
        out["runs"].append({"n_gpus": n, "numa_blocks": bool(args.numa_blocks), "e2e_spectra_per_s": len(sub) / dt, "ms_per_call": dt * 1e3, "first_block_equals_single_gpu": same})
        f, c = np.array(f), np.array(c)
        api.pinned_free(hs.masses); api.pinned_free(hs.intensities)
        n *= 2
    base = out["runs"][0]["e2e_spectra_per_s"]
    for r in out["runs"]:
        r["efficiency_vs_1gpu"] = r["e2e_spectra_per_s"] / (base * r["n_gpus"])
    print(json.dumps(out))


if __name__ == "__main__":
    main()
