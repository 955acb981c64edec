"""NVLink exchange probe (2+ ranks under torchrun): times b200_exchange_copy on plain contiguous
blocks -- local -> local, remote -> local (pull), local -> remote (push) -- next to torch's own
peer copy (cudaMemcpyPeerAsync) and the thread-level gather kernel, so that a slow exchange can be
attributed to the link, the copy engine or the kernel.  Prints one JSON line per rank 0."""
import ctypes as C
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from strawberryfields_b200 import lib as L
    from strawberryfields_b200.sharding import ShardedCircuit

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    world = dist.get_world_size()
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 9
    circ = ShardedCircuit(n, 10, exchange="p2p")       # owns two peer-mapped buffers per rank
    size = circ._size()
    nbytes = 16 * (size // 2)                            # copy half a shard, like a 2-rank exchange
    peer = (rank + 1) % world
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    out = {"world": world, "shard_GB": 16 * size / 1e9, "copy_GB": nbytes / 1e9}

    def timed(fn, reps=3):
        fn()
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        dist.barrier()
        return nbytes / (e0.elapsed_time(e1) / reps * 1e-3) / 1e9

    def xcopy(src_t, dst_t, run, ctas=0, local=False):
        d = L.XchgDesc()
        d.n_src, d.first_src = 1, 0
        total = nbytes // 16
        d.run = run
        d.n_axes = 1 if total // run > 1 else 0
        d.ext[0], d.ss[0], d.ds[0] = total // run, run, run
        d.src[0], d.dst[0] = src_t.data_ptr(), dst_t.data_ptr()
        d.src_base[0] = d.dst_base[0] = 0
        return lambda: L.call("b200_exchange_copy", C.byref(d), 0 if local else -1, ctas, stream)

    mine0, mine1 = circ._bufs[0], circ._bufs[1]
    theirs0, theirs1 = circ._peers[0][peer], circ._peers[1][peer]
    for run in (size // 2, 10 ** 5, 1000, 100):
        out["bulk_local_run%d" % run] = timed(xcopy(mine0, mine1, run))
        out["threads_local_run%d" % run] = timed(xcopy(mine0, mine1, run, local=True))
        out["bulk_pull_run%d" % run] = timed(xcopy(theirs0, mine1, run))
        out["bulk_push_run%d" % run] = timed(xcopy(mine0, theirs1, run))
    for ctas in (16, 32, 64):
        out["bulk_pull_ctas%d" % ctas] = timed(xcopy(theirs0, mine1, 10 ** 5, ctas))
        out["bulk_push_ctas%d" % ctas] = timed(xcopy(mine0, theirs1, 10 ** 5, ctas))
    half = size // 2
    out["torch_copy_local"] = timed(lambda: mine1[:half].copy_(mine0[:half]))
    out["torch_copy_pull"] = timed(lambda: mine1[:half].copy_(theirs0[:half]))
    out["torch_copy_push"] = timed(lambda: theirs1[:half].copy_(mine0[:half]))
    oa = [(half, 1, 0, 1)]
    out["gather_local"] = timed(lambda: circ._gather(mine0, None, mine1, oa))
    out["gather_pull"] = timed(lambda: circ._gather(theirs0, None, mine1, oa))
    out["gather_push"] = timed(lambda: circ._gather(mine0, None, theirs1, oa))
    if rank == 0:
        print(json.dumps({k: (round(v, 1) if isinstance(v, float) else v) for k, v in out.items()}))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
