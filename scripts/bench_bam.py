#!/usr/bin/env python3
"""BAM leg (BASELINE configs[4] shape): --cigar_scores --paired on a synthetic interlaced 2x150 bp BAM pair.

    python scripts/bench_bam.py --make N      write tmp_bam/{p,s}.bam with N records per file (slow Python writer: do it off the GPU box)
    python scripts/bench_bam.py               time xm_classify_bam_host on them; one JSON line

Reports host inflate GB/s (BGZF scan + zlib on host threads + record chain) separately from the GPU rendering GB/s
(three BAM kernels, bytes of SAM text produced) and from the walk itself.
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
DIR = os.path.join(ROOT, "tmp_bam")


def make(n):
    from tests import _bamwriter
    from tests.test_bam import FULL_HEADER
    from xenomapper_b200 import synth
    os.makedirs(DIR, exist_ok=True)
    p, s = synth.generate(n, seed=7, style=synth.STYLE_PE_BOWTIE2)
    hdr2 = FULL_HEADER.replace("SN:chr", "SN:").replace("SN:M\t", "SN:MT\t")
    open(os.path.join(DIR, "p.bam"), "wb").write(_bamwriter.sam_to_bam(FULL_HEADER, bytes(p), level=6))
    open(os.path.join(DIR, "s.bam"), "wb").write(_bamwriter.sam_to_bam(hdr2, bytes(s), level=6))


def main():
    if "--make" in sys.argv:
        make(int(sys.argv[sys.argv.index("--make") + 1]))
        return
    from xenomapper_b200 import _lib
    bp, bs = open(os.path.join(DIR, "p.bam"), "rb").read(), open(os.path.join(DIR, "s.bam"), "rb").read()
    ctx = _lib.Context(0)
    opts = ctx.opts(_lib.MODE_PE_LIBERAL, _lib.SCORE_CIGAR_NM, False, -40.0)
    best = None
    for k in range(5):
        ctx.bam_stats(reset=True)
        t0 = time.perf_counter()
        rc, res, outs = ctx.classify_bam_host(bp, bs, opts, want_outputs=False)
        wall = time.perf_counter() - t0
        st = ctx.bam_stats()
        assert rc == 0, ctx.error()
        row = dict(wall_s=wall, inflate_s=st.inflate_s, render_ms=st.render_ms, walk_ms=res.ms_total, records=int(st.records) // 2,
                   bam_bytes=int(st.bam_bytes), inflated_bytes=int(st.inflated_bytes), text_bytes=int(st.text_bytes))
        if k and (best is None or wall < best["wall_s"]):
            best = row
    # parity of the outputs with the walk on the SAM twins
    rc, res, outs = ctx.classify_bam_host(bp, bs, opts)
    from xenomapper_b200 import synth
    p, s = synth.generate(best["records"], seed=7, style=synth.STYLE_PE_BOWTIE2)        # the SAM text the BAMs were made from
    rc2, res2, outs2 = ctx.classify_host(p, s, opts)
    best["outputs_equal_sam_walk"] = outs == outs2 and list(res.counts) == list(res2.counts)
    best.update(host_inflate_gb_per_s=best["inflated_bytes"] / best["inflate_s"] / 1e9,
                gpu_render_gb_per_s=best["text_bytes"] / (best["render_ms"] / 1e3) / 1e9,
                reads_per_s_end_to_end=2 * best["records"] / best["wall_s"], host_threads=min(os.cpu_count() or 1, 32),
                workload="synthetic interlaced 2x150bp BAM pair, --paired --cigar_scores --min_score -40 (BASELINE configs[4] shape)")
    print(json.dumps(best))


if __name__ == "__main__":
    main()
