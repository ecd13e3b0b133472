"""Two batches of PacBio reads left on the device, for ncu:
    ncu --set full --import-source on -k regex:k_pb_warp -s 2 -c 2 -o gpurun_out/prof_pb python tools/pacbio_ncu_probe.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import jackalope_b200 as J

ctx = J.Context(0)
g = J.random_genome(10, 10_000_000, seed=3)
for i in range(2):
    st = J.pacbio(g, "", 16384, seed=5 + i, ctx=ctx, sink="device", n_threads=16)
    print("kernels %.3f ms for %d reads, %.0f Mbases" % (st["reads_ms"], st["pairs"], st["bytes_out"][0] / 2e6))
