#!/usr/bin/env python3
"""NTT / LDE / leaf-hash / Merkle micro-benchmarks (BASELINE.json config 5) on device-resident data.

Times each building block with CUDA events on the library's stream (dg_dev_* entry points), L2 flushed between
iterations, and reports algorithmic GB/s (SURVEY.md 8d byte counts) against the measured HBM peak.
    python tools/microbench.py [--max-log 24] [--iters 5] [--out gpurun_out/microbench.json]
"""
import argparse
import ctypes
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from distaff_b200 import backend, felt  # noqa: E402


def peak_gbs():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"], "measured"
    except Exception:
        return 6650.0, "fallback"


def timed(fn, iters):
    L = backend.lib()
    ms = ctypes.c_float(0)
    best = []
    for i in range(iters + 2):
        backend.check(L.dg_dev_flush_l2())
        fn(ctypes.byref(ms))
        if i >= 2:
            best.append(ms.value)
    return float(np.median(best)), float(min(best))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--min-log", type=int, default=14)
    ap.add_argument("--max-log", type=int, default=24)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    L = backend.lib()
    info = backend.device_info()
    peak, peak_kind = peak_gbs()
    results = {"device": info, "peak_gbs": peak, "peak_kind": peak_kind, "ntt": [], "lde": [], "hash_rows": [], "merkle": []}

    for log_n in range(args.min_log, args.max_log + 1, 2):
        n = 1 << log_n
        buf = backend.DeviceBuffer(n * 16).upload(felt.random_elements(n, log_n))
        med, best = timed(lambda ms: backend.check(L.dg_dev_ntt(buf.ptr, log_n, 1, 0, ms)), args.iters)
        gbs = 32.0 * n / (med * 1e-3) / 1e9
        results["ntt"].append({"log_n": log_n, "ms": med, "ms_best": best, "alg_gbs": gbs, "frac": gbs / peak})
        print(f"ntt      2^{log_n}: {med:8.3f} ms  {gbs:8.1f} GB/s  ({100 * gbs / peak:.1f}% of {peak_kind} HBM peak)", flush=True)
        buf.free()

    # batched NTT (16 columns) and LDE x32 of 16 columns
    for log_n in (12, 14, 16, 18, 20):
        w = 16
        n = 1 << log_n
        if n * 32 * w * 16 > 40 << 30:
            continue
        polys = backend.DeviceBuffer(w * n * 16).upload(felt.random_elements(w * n, 7))
        ext = backend.DeviceBuffer(w * n * 32 * 16)
        med, best = timed(lambda ms: backend.check(L.dg_dev_lde(polys.ptr, ext.ptr, log_n, 5, w, ms)), args.iters)
        alg = w * (16.0 * n + 16.0 * n * 32)
        gbs = alg / (med * 1e-3) / 1e9
        results["lde"].append({"log_n": log_n, "w": w, "blowup": 32, "ms": med, "ms_best": best, "alg_gbs": gbs, "frac": gbs / peak})
        print(f"lde x32  2^{log_n} x{w}: {med:8.3f} ms  {gbs:8.1f} GB/s  ({100 * gbs / peak:.1f}%)", flush=True)
        leaves = backend.DeviceBuffer(n * 32 * 32)
        med, best = timed(lambda ms: backend.check(L.dg_dev_hash_rows(ext.ptr, w, log_n, 5, leaves.ptr, ms)), args.iters)
        alg = 16.0 * w * n * 32 + 32.0 * n * 32
        gbs = alg / (med * 1e-3) / 1e9
        results["hash_rows"].append({"log_n": log_n, "w": w, "rows": n * 32, "ms": med, "alg_gbs": gbs, "frac": gbs / peak})
        print(f"leafhash 2^{log_n + 5} rows x{w}: {med:8.3f} ms  {gbs:8.1f} GB/s  ({100 * gbs / peak:.1f}%)", flush=True)
        nodes = backend.DeviceBuffer(n * 32 * 32)
        med, best = timed(lambda ms: backend.check(L.dg_dev_merkle_build(leaves.ptr, n * 32, nodes.ptr, ms)), args.iters)
        gbs = 64.0 * n * 32 / (med * 1e-3) / 1e9
        results["merkle"].append({"log_leaves": log_n + 5, "ms": med, "alg_gbs": gbs, "frac": gbs / peak})
        print(f"merkle   2^{log_n + 5} leaves: {med:8.3f} ms  {gbs:8.1f} GB/s  ({100 * gbs / peak:.1f}%)", flush=True)
        for b in (polys, ext, leaves, nodes):
            b.free()

    # Merkle trees with the algebraic hashes (benches/hash.rs + merkle.rs with rescue / poseidon); leaves = pairs of field elements
    results["alg_merkle"] = []
    for name, hid in (("rescue", 1), ("poseidon", 2)):
        for log_l in (14, 16, 18, 20):
            L_ = 1 << log_l
            leaves = backend.DeviceBuffer(L_ * 32).upload(felt.random_elements(2 * L_, 7 + log_l))
            nodes = backend.DeviceBuffer(L_ * 32)
            med, best = timed(lambda ms: backend.check(L.dg_dev_merkle_build_with(hid, leaves.ptr, L_, nodes.ptr, ms)), max(3, args.iters // 2))
            results["alg_merkle"].append({"hash": name, "log_leaves": log_l, "ms": med, "hashes_per_s": (L_ - 1) / (med * 1e-3)})
            print(f"{name:8s} merkle 2^{log_l} leaves: {med:8.3f} ms  {(L_ - 1) / (med * 1e-3) / 1e6:8.2f} M hashes/s", flush=True)
            leaves.free(); nodes.free()

    if args.out:
        os.makedirs(os.path.dirname(args.out), exist_ok=True)
        json.dump(results, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
