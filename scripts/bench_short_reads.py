#!/usr/bin/env python3
"""Short-read lines: the span kernels with spans of half the size against the exact kernels, resident, per line width."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from tests.test_emu_tiles import _fixed_width_pair
    from xenomapper_b200 import _lib
    rows = []
    for width in (104, 128, 160, 190, 256):
        p, s = _fixed_width_pair(3_000_000, width)
        opts = _lib.Context.opts(0, 0, True)
        row = dict(line_bytes=width, records=3_000_000)
        for name, dbg in (("span_kernels", 0), ("exact_kernels", _lib.DEBUG_FORCE_GENERIC)):
            c = _lib.Context(0)
            c.set_debug(dbg)
            best = 1e9
            for k in range(6):
                rc, res, _ = c.classify_host(p, s, opts, want_outputs=False)
                assert rc == 0
                if k >= 2:
                    best = min(best, res.ms_scan + res.ms_classify)
            row[name] = dict(kernels=c.walk_kernels(), ms=round(best, 3), m_reads_per_s=round(3e6 / best / 1e3, 1))
            c.close()
        rows.append(row)
    print(json.dumps(rows))


if __name__ == "__main__":
    main()
