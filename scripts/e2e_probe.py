"""Scratch: host-ingest diagnostics on the GPU box: NUMA layout, syl_pack2 scaling, per-call e2e timings
for several packer thread counts (run each thread count in its own process: the pool size is fixed per ctx)."""
import ctypes as C
import os
import subprocess
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import sylph_b200
    from sylph_b200 import _lib, synth
    mode = sys.argv[1] if len(sys.argv) > 1 else "all"
    if mode == "all":
        print(subprocess.run("lscpu | egrep 'Model name|Socket|NUMA|^CPU\\(s\\)|Thread'; nvidia-smi topo -m | head -12; free -g | head -2",
                             shell=True, stdout=subprocess.PIPE, text=True).stdout)
        for thr in (8, 16, 32, 64, 128):
            env = dict(os.environ, SYL_PACK_THREADS=str(thr))
            subprocess.run([sys.executable, __file__, "one"], env=env)
        return
    n_reads = 6_666_667
    b, o = synth.reads(n_reads, device="cuda")
    hb = torch.empty(b.numel(), dtype=torch.uint8, pin_memory=True)
    ho = torch.empty(n_reads + 1, dtype=torch.int64, pin_memory=True)
    hb.copy_(b)
    ho.copy_(o)
    torch.cuda.synchronize()
    hbn, hon = hb.numpy(), ho.numpy().view(np.uint64)
    L = _lib.lib()
    thr = L.syl_pack_threads()
    words = np.zeros((hbn.size + 15) // 16, np.uint32)
    ts = []
    for _ in range(4):
        t = time.perf_counter()
        L.syl_pack2(hbn.ctypes.data_as(C.c_void_p), hbn.size, words.ctypes.data_as(C.c_void_p), thr)
        ts.append((time.perf_counter() - t) * 1e3)
    ctx = sylph_b200.Context(0, stream=torch.cuda.current_stream().cuda_stream)
    oh = torch.empty(3_000_000, dtype=torch.int64, pin_memory=True).numpy().view(np.uint64)
    oc = torch.empty(3_000_000, dtype=torch.int32, pin_memory=True).numpy().view(np.uint32)
    es = []
    for _ in range(12):
        torch.cuda.synchronize()
        t = time.perf_counter()
        s = ctx.sketch_sequences(hbn, hon)
        s.download(oh, oc)
        s.free()
        torch.cuda.synchronize()
        es.append((time.perf_counter() - t) * 1e3)
    print("threads %3d: syl_pack2 1 GB ms %s | e2e sketch ms/call %s" % (thr, " ".join("%.1f" % x for x in ts), " ".join("%.1f" % x for x in es)), flush=True)


if __name__ == "__main__":
    main()
