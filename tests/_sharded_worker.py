"""One rank of a sharded walk under gloo, with the CPU emulation of the kernels as the engine.
TEST SCAFFOLDING: exercises xenomapper_b200/sharded.py (byte ranges, count exchange, partition points,
context records, output placement) without a GPU.

    python -m tests._sharded_worker RANK WORLD PORT CASE.json
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


class EmuEngine:
    def __init__(self, debug=0):
        from tests import _emu
        self.emu = _emu
        self.debug = debug

    def index(self, buf, queries, skip):
        info, off = self.emu.index(buf, list(queries), skip_repeated=skip, debug=self.debug)
        return int(info.n_records), info.stop_at != (1 << 64) - 1, int(info.end_off), off

    def walk(self, prim, sec, mode, score_src, skip, min_score, enabled_bins, first_is_context):
        r = self.emu.classify(prim, sec, mode=mode, score_src=score_src, skip_repeated=skip, min_score=min_score,
                              enabled_bins=enabled_bins, debug=self.debug, first_is_context=first_is_context)
        return dict(status=r["status"], counts=r["counts"], outputs=r["outputs"], n_records=int(r["n_records"]),
                    err_record=int(r["err_record"]), message=r["message"])


def main():
    rank, world, port, casefile = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
    case = json.load(open(casefile))
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    from xenomapper_b200 import sharded
    prim = sharded.FileSource(case["prim"], case.get("prim_off", 0))
    sec = sharded.FileSource(case["sec"], case.get("sec_off", 0))
    if case.get("engine") == "gpu":
        from xenomapper_b200 import _lib
        ctx = _lib.Context(int(case.get("device", 0)))
        if case.get("debug", 0):
            ctx.set_debug(case["debug"])
        engine = sharded.GpuEngine(ctx)
    else:
        engine = EmuEngine(case.get("debug", 0))
    res = sharded.sharded_walk(engine, prim, sec, mode=case["mode"], score_src=case["score_src"],
                               skip=case["skip"], min_score=case["min_score"], enabled_bins=case["enabled_bins"])
    fds = [os.open(p, os.O_WRONLY) for p in case["outs"]]
    sharded.write_outputs(res, fds, [0] * 6)
    [os.close(f) for f in fds]
    if world > 1:
        dist.barrier()
    if rank == 0:
        json.dump({k: res[k] for k in ("status", "message", "err_record", "counts", "n_records", "out_total")},
                  open(case["result"], "w"))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
