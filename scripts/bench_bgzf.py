#!/usr/bin/env python3
"""The device BGZF compressor on SAM text: ratio and kernel throughput against zlib level 1 / 6 on one host core.

    python scripts/bench_bgzf.py [records]
"""
import gzip
import json
import os
import sys
import time
import zlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from xenomapper_b200 import _lib, synth
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
    p, _ = synth.generate(n, seed=5, style=synth.STYLE_PE_BOWTIE2)
    data = bytes(p)
    ctx = _lib.Context(0)
    ctx.bgzf_deflate_host(data[:1 << 20])
    best = None
    for k in range(3):
        ctx.bgzf_stats(reset=True)
        t0 = time.perf_counter()
        z = ctx.bgzf_deflate_host(data)
        wall = time.perf_counter() - t0
        st = ctx.bgzf_stats()
        row = dict(bytes_in=len(data), bytes_out=len(z), ratio=len(data) / len(z), kernel_ms=st.kernel_ms,
                   deflate_gb_per_s=len(data) / (st.kernel_ms / 1e3) / 1e9, wall_s_host_to_host=wall, members=int(st.members))
        if best is None or row["kernel_ms"] < best["kernel_ms"]:
            best = row
    sample = data[:64 << 20]
    t0 = time.perf_counter()
    ok = gzip.decompress(z + bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")) == data
    best["inflates_to_input"] = ok
    for level in (1, 6):
        t0 = time.perf_counter()
        zz = sum(len(zlib.compress(sample[o:o + 0xff00], level)) for o in range(0, len(sample), 0xff00))
        dt = time.perf_counter() - t0
        best["zlib_level_%d" % level] = dict(ratio=len(sample) / zz, gb_per_s_one_core=len(sample) / dt / 1e9)
    best["workload"] = "%d synthetic 2x150 bp Bowtie2 SAM records (primary stream)" % n
    print(json.dumps(best))


if __name__ == "__main__":
    main()
