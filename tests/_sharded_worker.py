"""One rank of a sharded walk.  TEST SCAFFOLDING.

    python -m tests._sharded_worker RANK WORLD PORT CASE.json

engine "emu": the walk across ranks of csrc/xm_shard.h on the CPU (tests/emu), its all-gather and send/receive
served by torch.distributed over gloo -- byte shards, line heads, context lines, filler, counts, slivers and the
placement in the bins are exercised without a GPU.  engine "gpu": the same through libxenomapper_b200.so, one
device per rank, NCCL inside the library (needs WORLD devices).
"""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def gloo_collectives(dist, torch, rank, world):
    def view(addr, n):
        return torch.frombuffer((C.c_uint8 * n).from_address(addr), dtype=torch.uint8) if n else torch.empty(0, dtype=torch.uint8)

    def all_gather(send, recv, nbytes):
        try:
            mine = view(send, nbytes).clone()
            parts = [torch.empty(nbytes, dtype=torch.uint8) for _ in range(world)]
            dist.all_gather(parts, mine)
            view(recv, nbytes * world).copy_(torch.cat(parts))
            return 0
        except Exception as e:                                    # noqa: BLE001 -- reported as a failed collective
            print("all_gather callback failed:", e, file=sys.stderr)
            return 1

    def exchange(sends, ns, recvs, nr):
        try:
            # messages between one pair of ranks are matched in the order both sides list them (NCCL's rule): tag = position
            ops, keep, count = [], [], {}
            for k in range(nr):
                x = recvs[k]
                t = view(x.ptr, x.bytes)
                tag = count.get(("r", x.peer), 0); count[("r", x.peer)] = tag + 1
                ops.append(dist.irecv(t, src=x.peer, tag=tag))
                keep.append(t)
            for k in range(ns):
                x = sends[k]
                t = view(x.ptr, x.bytes).clone()
                tag = count.get(("s", x.peer), 0); count[("s", x.peer)] = tag + 1
                ops.append(dist.isend(t, dst=x.peer, tag=tag))
                keep.append(t)
            for op in ops:
                op.wait()
            return 0
        except Exception as e:                                    # noqa: BLE001
            print("exchange callback failed:", e, file=sys.stderr)
            return 1

    return all_gather, exchange


def main():
    rank, world, port, casefile = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
    case = json.load(open(casefile))
    from xenomapper_b200 import sharded
    prim_size = os.path.getsize(case["prim"]) - case.get("prim_off", 0)
    sec_size = os.path.getsize(case["sec"]) - case.get("sec_off", 0)
    fp, fs = os.open(case["prim"], os.O_RDONLY), os.open(case["sec"], os.O_RDONLY)
    kw = dict(mode=case["mode"], score_src=case["score_src"], min_score=case["min_score"], enabled_bins=case["enabled_bins"])
    if case.get("engine") == "gpu":
        from xenomapper_b200 import _lib
        ctx = _lib.Context(rank if case.get("one_device_per_rank", True) else 0)
        rv = sharded.Rendezvous(rank, world, directory=case["rendezvous"])
        sharded.init_comm(ctx, rank, world, rv)
        res = sharded.sharded_walk(ctx, rank, world, fp, case.get("prim_off", 0), fs, case.get("sec_off", 0), skip=case["skip"], **kw)
        ctx.comm_barrier()
        done = ctx.close
    else:
        import torch
        import torch.distributed as dist
        from tests import _emu
        if world > 1:
            dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
        ag, ex = gloo_collectives(dist, torch, rank, world)
        lo, hi = sharded.byte_range(prim_size, rank, world)
        p = sharded.read_range(fp, case.get("prim_off", 0) + lo, case.get("prim_off", 0) + hi)
        lo, hi = sharded.byte_range(sec_size, rank, world)
        s = sharded.read_range(fs, case.get("sec_off", 0) + lo, case.get("sec_off", 0) + hi)
        res = _emu.classify_sharded(p, s, rank, world, ag, ex, skip_repeated=case["skip"], room=case.get("room", 1 << 16), **kw)
        if res["status"] == -2:
            print("rank %d: sharded walk declined: %s" % (rank, res["message"]), file=sys.stderr)
            # every rank declined together (a blank line, a dirty line ...): the library's host entry point then gathers
            # the shards on rank 0 for the exact walk; here rank 0 reads the files and the others contribute nothing
            if rank == 0:
                r = _emu.classify(open(case["prim"], "rb").read()[case.get("prim_off", 0):], open(case["sec"], "rb").read()[case.get("sec_off", 0):],
                                  skip_repeated=case["skip"], **kw)
                res = dict(r, out_offset=[0] * 6, out_total=[len(x) for x in r["outputs"]], declined=True)
            else:
                res = dict(status=0, message="", err_record=0, counts=[0] * 36, n_records=0, outputs=[b""] * 6, out_offset=[0] * 6,
                           out_total=[0] * 6, declined=True)

        def done():
            if world > 1:
                dist.barrier()
                dist.destroy_process_group()
    fds = [os.open(p_, os.O_WRONLY) for p_ in case["outs"]]
    sharded.write_outputs(res, fds, [0] * 6)
    [os.close(f) for f in fds]
    json.dump({k: res.get(k) for k in ("status", "message", "err_record", "counts", "n_records", "out_total", "records", "sliver_bytes", "declined")},
              open("%s.%d" % (case["result"], rank), "w"))
    done()


if __name__ == "__main__":
    main()
