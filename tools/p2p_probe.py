"""Diagnostic (run under torchrun on >= 2 GPUs): can a kernel on this rank read a peer's buffer that
was shared through torch's CUDA-IPC storage handles?  Tries opening the handle in the exporter's
and in the importer's device context."""
import os
import sys
import ctypes as C

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from strawberryfields_b200 import lib as L  # noqa: E402

rank = int(os.environ["RANK"])
world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl")
dev = torch.cuda.current_device()
n = 1 << 20
mine = (torch.arange(n, dtype=torch.float64, device="cuda") + 1000 * rank).to(torch.complex128)
shared = (mine.untyped_storage()._share_cuda_(), mine.storage_offset())
everyone = [None] * world
dist.all_gather_object(everyone, (dev, shared))
peer = (rank + 1) % world
pdev, (handle, off) = everyone[peer]
L.load()
for where in ("importer", "exporter"):
    try:
        if where == "importer":
            L.call("b200_enable_peer_access", int(pdev))
            h = (dev,) + tuple(handle[1:])
        else:
            h = tuple(handle)
        st = torch.UntypedStorage._new_shared_cuda(*h)
        typed = torch.storage.TypedStorage(wrap_storage=st, dtype=torch.complex128, _internal=True)
        t = torch._utils._rebuild_tensor(typed, off, (n,), (1,))
        out = torch.zeros(n, dtype=torch.complex128, device="cuda")
        d = L.GatherDesc()
        d.n_out_axes, d.n_red_axes = 1, 0
        d.out_ext[0], d.out_sa[0], d.out_sb[0], d.out_sc[0] = n, 1, 0, 1
        L.call("b200_gather_reduce", C.byref(d), C.c_void_p(t.data_ptr()), None, C.c_void_p(out.data_ptr()), 0, None, None)
        torch.cuda.synchronize()
        ok = bool((out.real[:4].cpu() == torch.arange(4, dtype=torch.float64) + 1000 * peer).all())
        print("rank", rank, "open in", where, "context: kernel read of peer memory", "OK" if ok else "WRONG", flush=True)
        break
    except Exception as exc:
        print("rank", rank, "open in", where, "context FAILED:", repr(exc)[:200], flush=True)
        break
dist.barrier()
dist.destroy_process_group()
