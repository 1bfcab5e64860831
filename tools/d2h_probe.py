#!/usr/bin/env python
"""Platform probe for the end-to-end line at N GPUs: what device->host (and host->device) copy rate does each GPU get
when all N copy at once, and on which NUMA node do the pinned pages of a rank land?  Run under torchrun like bench.py:
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/d2h_probe.py [--gb 4] [--out file]
Development aid (not part of the product path): plain cudaMemcpyAsync through torch, the library is used only for the
binding call and its pinned allocator, so the probe measures the same allocations bench.py makes."""
import argparse, ctypes as C, json, os, sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import jpegdec_b200 as J


def page_nodes(ptr, nbytes, samples=256):
    """NUMA node of `samples` pages spread over [ptr, ptr+nbytes) via move_pages(2) in query mode."""
    libc = C.CDLL(None, use_errno=True)
    step = max(nbytes // samples, 4096) & ~4095
    pages = [(ptr & ~4095) + i * step for i in range(samples) if i * step < nbytes]
    arr = (C.c_void_p * len(pages))(*pages)
    status = (C.c_int * len(pages))()
    rc = libc.syscall(279, 0, C.c_ulong(len(pages)), arr, None, status, 0)      # __NR_move_pages on x86-64
    if rc != 0:
        return {"error": C.get_errno()}
    hist = {}
    for s in status:
        hist[int(s)] = hist.get(int(s), 0) + 1
    return hist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gb", type=float, default=4.0)
    ap.add_argument("--chunk-mb", type=int, default=512)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = J.Context(local, J.JPEG_ARITH_SSE2)
    bound = 0 if os.environ.get("JPEGDEC_B200_NO_BIND") else ctx.bind_host_to_device()
    nbytes = int(args.gb * (1 << 30)) & ~((1 << 20) - 1)
    chunk = args.chunk_mb << 20
    L = J.lib()
    hp = L.JPEGB200_hostAlloc(nbytes)                      # the library's pinned allocator (what bench.py's buffers use)
    host_np = np.ctypeslib.as_array(C.cast(hp, C.POINTER(C.c_ubyte)), shape=(nbytes,))
    host_np[::4096] = 1                                    # touch
    host = torch.from_numpy(host_np)
    tp = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)    # torch's pinned allocator, for comparison
    dev = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    dev2 = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    res = {"rank": rank, "gpu_node": ctx.numa_node(), "bound_cpus": bound, "cpus": len(os.sched_getaffinity(0)),
           "pages_lib": page_nodes(hp, nbytes), "pages_torch": page_nodes(tp.data_ptr(), nbytes)}
    # cudaMemcpyAsync needs torch to know the lib buffer is pinned: register is not needed for speed measurements of
    # the DMA path if the memory is already page-locked by cudaHostAlloc (the runtime recognises the range).

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def run(kind, hbuf):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s1)
        for o in range(0, nbytes, chunk):
            if kind in ("d2h", "both"):
                with torch.cuda.stream(s1):
                    hbuf[o:o + chunk].copy_(dev[o:o + chunk], non_blocking=True)
            if kind in ("h2d", "both"):
                with torch.cuda.stream(s2 if kind == "both" else s1):
                    dev2[o:o + chunk].copy_(hbuf[o:o + chunk] if kind == "h2d" else tp[o:o + chunk], non_blocking=True)
        s1.wait_stream(s2)
        e1.record(s1)
        torch.cuda.synchronize()
        return nbytes / (e0.elapsed_time(e1) / 1e3) / 1e9

    for name, hbuf in (("lib", host), ("torch", tp)):
        for kind in ("d2h", "h2d", "both"):
            # all ranks at once
            run(kind, hbuf); sync_all()
            v = []
            for _ in range(3):
                sync_all(); v.append(run(kind, hbuf))
            res["%s_%s_all" % (name, kind)] = round(max(v), 2)
    # one rank at a time (the others idle)
    for r in range(world):
        sync_all()
        if r == rank:
            run("d2h", host)
            res["lib_d2h_alone"] = round(max(run("d2h", host) for _ in range(3)), 2)
    sync_all()
    line = json.dumps(res)
    if args.out:
        with open("%s.rank%d" % (args.out, rank), "w") as f:
            f.write(line + "\n")
    print(line, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
