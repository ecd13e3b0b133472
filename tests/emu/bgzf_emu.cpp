// The device BGZF coder's kernels (jackalope_b200/csrc/jlp_bgzf.cu, compiled as C++ over tests/emu/cuda_emu.h)
// run on host threads the way jlp_bgzf_device runs them on the GPU:  bgzf_emu <in> <level> <out>
// Test infrastructure (tests/test_bgzf_emu.py); the product never builds or loads this.
#define JLP_CPU_EMU 1
#include "../../jackalope_b200/csrc/jlp_bgzf.cu"

#include <cstdio>

namespace jlp_emu {
thread_local Block* tl_block;
thread_local uint3 tl_tid, tl_bid, tl_gdim;
}

int main(int argc, char** argv) {
    if (argc != 4) { std::fprintf(stderr, "usage: bgzf_emu <in> <level> <out>\n"); return 2; }
    std::FILE* f = std::fopen(argv[1], "rb");
    if (!f) return 2;
    std::vector<uint8_t> in;
    uint8_t tmp[65536];
    for (size_t n; (n = std::fread(tmp, 1, sizeof tmp, f)) > 0;) in.insert(in.end(), tmp, tmp + n);
    std::fclose(f);
    const int level = std::atoi(argv[2]);
    const uint64_t n = in.size();
    const uint32_t nblk = (uint32_t)((n + jlp::kBgzfIn - 1) / jlp::kBgzfIn);
    in.resize(n + 64, 0xee);                                  // the slack the device buffers have
    std::vector<uint8_t> slots((size_t)nblk * jlp::kBgzfSlot + 16, 0xcd), out((size_t)nblk * jlp::kBgzfSlot + 16, 0xcd);
    std::vector<uint8_t> codes(jlp::kBgzfCodeBytes, 0xcd);
    std::vector<uint32_t> zlen(nblk + 1, 0xcdcdcdcdu);
    std::vector<uint64_t> zoff(nblk + 1);
    uint64_t tot[4] = {n, 0, 0, 0};
    if (jlp::bgzf_init() != cudaSuccess) return 3;
    if (jlp::launch_bgzf(in.data(), in.data(), tot, nblk, level >= 4, slots.data(), slots.data(), zlen.data(), zlen.data(), zoff.data(),
                         zoff.data(), out.data(), out.data(), codes.data(), nullptr) != cudaSuccess)
        return 3;
    static const uint8_t eof[28] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0, 0x1b, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    f = std::fopen(argv[3], "wb");
    if (!f) return 2;
    std::fwrite(out.data(), 1, tot[2], f);
    std::fwrite(eof, 1, sizeof eof, f);
    std::fclose(f);
    return 0;
}
