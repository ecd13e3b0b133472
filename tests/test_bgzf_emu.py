"""The device BGZF coder's kernels on host threads (tests/emu): jackalope_b200/csrc/jlp_bgzf.cu is compiled unchanged as
C++ over a small CUDA execution model (one OS thread per CUDA thread, barriers and warp collectives on pthread barriers)
and run the way jlp_bgzf_device runs it on the GPU.  What comes out must inflate (zlib, CRC-32 and ISIZE of every member
checked) to what went in.  This is the CPU suite's view of the kernels' logic -- code sharing, the own-code fallback, the
bit-level join of the segment images; tests/test_gpu_bgzf.py runs the same inputs through the GPU."""
import gzip
import importlib.util
import os
import subprocess
import sys
import types

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _gpu_cases():
    """CASES / helpers of tests/test_gpu_bgzf.py without importing the product (there is no GPU here)."""
    saved = sys.modules.get("jackalope_b200")
    sys.modules["jackalope_b200"] = types.ModuleType("jackalope_b200")
    try:
        spec = importlib.util.spec_from_file_location("_gpu_bgzf_cases", os.path.join(HERE, "test_gpu_bgzf.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        if saved is None:
            del sys.modules["jackalope_b200"]
        else:
            sys.modules["jackalope_b200"] = saved
    return mod


G = _gpu_cases()
# a subset that reaches every path and keeps the CPU suite short (a block takes about half a second)
NAMES = ["empty", "one_byte", "short_text", "block_plus_one", "three_blocks_ragged", "random_bytes_stored", "fibonacci_depth",
         "repeated_lines", "short_lines", "long_codes_before_matches", "late_binary_bytes", "nonstationary", "lines_of_128",
         "lines_of_129", "identical_300", "long_lines", "chunk_tail_129"]


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("emu") / "bgzf_emu")
    cmd = ["g++", "-O1", "-std=c++17", "-pthread", "-Wno-unknown-pragmas", "-I", os.path.join(HERE, "emu"), "-x", "c++",
           os.path.join(HERE, "emu", "bgzf_emu.cpp"), "-o", exe]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    return exe


@pytest.mark.parametrize("level", [1, 6])
@pytest.mark.parametrize("name", NAMES)
def test_emulated_kernels_round_trip(emu, tmp_path, name, level):
    data = G.CASES[name]
    src, dst = tmp_path / "in.bin", tmp_path / "out.gz"
    src.write_bytes(data)
    assert subprocess.run([emu, str(src), str(level), str(dst)], timeout=600).returncode == 0
    z = dst.read_bytes()
    blocks = G.bgzf_blocks(z)
    assert blocks[-1] == (28, 0)
    assert [i for _, i in blocks[:-1]] == [min(0xff00, len(data) - o) for o in range(0, len(data), 0xff00)]
    assert gzip.decompress(z) == data
    if len(data) >= 1000 and name != "random_bytes_stored":
        assert len(z) <= 1.01 * G.huffman_only_size(data) + 150 * len(blocks)
