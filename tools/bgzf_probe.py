"""A few batches of PE150 HS25 FASTQ compressed on the device (k_bgzf), for ncu:
    ncu --set full --import-source on -k regex:k_bgzf -c 1 -o gpurun_out/bgzf python tools/bgzf_probe.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import jackalope_b200 as J

ctx = J.Context(0)
g = J.create_genome(8, 25_000_000, seed=5, ctx=ctx)
B = 1 << 20
level = int(sys.argv[2]) if len(sys.argv) > 2 else 6
for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    st = J.illumina(g, "", 2 * 2 * B, 150, True, seed=7 + i, ctx=ctx, sink="device", batch_pairs=B, seq_sys="HS25",
                    compress=level, comp_engine="device")
    print("level %d:" % level, "bgzf %.3f ms/batch, reads %.3f ms/batch, ratio %.4f" % (st["bgzf_ms"] / st["batches"], st["reads_ms"] / st["batches"],
                                                                  sum(st["z_bytes"]) / sum(st["bytes_out"])))
