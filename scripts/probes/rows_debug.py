import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import ctypes as C
from xenomapper_b200 import _lib, synth
from oracle import oracle
ctx = _lib.Context(0)
for style, mode, skip in ((synth.STYLE_SE_BOWTIE2, 0, True), (synth.STYLE_PE_BOWTIE2, 1, False)):
    p, s = synth.generate(int(sys.argv[1]) if len(sys.argv) > 1 else 3000, seed=3, style=style)
    opts = ctx.opts(mode=mode, skip_repeated=skip)
    ref = oracle.classify(p, s, mode=mode, skip_repeated=skip)
    dp, ds = ctx.dev_alloc(p.nbytes), ctx.dev_alloc(s.nbytes)
    ctx.h2d(dp, p); ctx.h2d(ds, s)
    res = _lib.Result()
    zero_out = (C.c_void_p * 6)(*([None] * 6)); zero_cap = (C.c_uint64 * 6)(*([0] * 6))
    rc = ctx.lib.xm_classify_device(ctx.h, dp, p.nbytes, ds, s.nbytes, C.byref(opts), zero_out, zero_cap, C.byref(res))
    print("mode", mode, "rc", rc, ctx.error(), "n", res.n_records, "kernels", ctx.walk_kernels())
    print(" out_len", list(res.out_len), "ref", [len(o) for o in ref["outputs"]])
    print(" counts", [c for c in res.counts if c], "ref", [c for c in ref["counts"] if c])
    rc, r2, outs = ctx.classify_host(p, s, opts)
    print(" host rc", rc, "equal outputs", outs == ref["outputs"], "counts eq", list(r2.counts) == ref["counts"], ctx.walk_kernels())
    for b in range(6):
        if outs[b] != ref["outputs"][b]:
            a, r = outs[b], ref["outputs"][b]
            k = next((i for i in range(min(len(a), len(r))) if a[i] != r[i]), min(len(a), len(r)))
            print("  bin", b, "len", len(a), len(r), "first diff at", k, a[max(0,k-30):k+30], r[max(0,k-30):k+30])
