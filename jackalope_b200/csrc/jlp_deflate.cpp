#include "jlp_deflate.h"

#include <zlib.h>

#include <algorithm>
#include <cstring>

namespace jlp {

const uint8_t kBgzfEof[28] = {0x1f, 0x8b, 0x08, 0x04, 0, 0, 0, 0, 0, 0xff, 0x06, 0, 0x42, 0x43, 0x02, 0,
                              0x1b, 0, 0x03, 0, 0, 0, 0, 0, 0, 0, 0, 0};

namespace {

void put32(std::vector<uint8_t>& v, uint32_t x) {
    for (int i = 0; i < 4; i++) v.push_back(static_cast<uint8_t>(x >> (8 * i)));
}

// one member: header, raw deflate stream, CRC32, ISIZE
std::string one_member(int method, int level, const uint8_t* p, size_t n, std::vector<uint8_t>& out) {
    const size_t start = out.size();
    static const uint8_t gz_hdr[10] = {0x1f, 0x8b, 0x08, 0x00, 0, 0, 0, 0, 0, 0xff};
    static const uint8_t bgzf_hdr[18] = {0x1f, 0x8b, 0x08, 0x04, 0, 0, 0, 0, 0, 0xff, 0x06, 0, 0x42, 0x43, 0x02, 0, 0, 0};
    if (method == DEFLATE_BGZF) out.insert(out.end(), bgzf_hdr, bgzf_hdr + 18);
    else out.insert(out.end(), gz_hdr, gz_hdr + 10);
    z_stream zs;
    std::memset(&zs, 0, sizeof zs);
    if (deflateInit2(&zs, level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) return "deflateInit2 failed";
    const size_t bound = deflateBound(&zs, static_cast<uLong>(n));
    const size_t body = out.size();
    out.resize(body + bound);
    zs.next_in = const_cast<Bytef*>(p);
    zs.avail_in = static_cast<uInt>(n);
    zs.next_out = out.data() + body;
    zs.avail_out = static_cast<uInt>(bound);
    int rc = deflate(&zs, Z_FINISH);
    const size_t produced = bound - zs.avail_out;
    deflateEnd(&zs);
    if (rc != Z_STREAM_END) return "deflate failed";
    out.resize(body + produced);
    put32(out, static_cast<uint32_t>(crc32(crc32(0L, Z_NULL, 0), p, static_cast<uInt>(n))));
    put32(out, static_cast<uint32_t>(n));
    if (method == DEFLATE_BGZF) {
        const size_t bsize = out.size() - start;          // whole block
        if (bsize > 65536) return "BGZF block does not fit 64 KiB";
        out[start + 16] = static_cast<uint8_t>((bsize - 1) & 0xff);
        out[start + 17] = static_cast<uint8_t>((bsize - 1) >> 8);
    }
    return std::string();
}

}  // namespace

std::string deflate_members(int method, int level, const uint8_t* p, size_t n, std::vector<uint8_t>& out) {
    if (level < 0 || level > 9) return "\nInvalid bgzip compress level. It must be in range [0,9].";   // src/io.h:113-117
    const size_t chunk = method == DEFLATE_BGZF ? 0xff00 : (1u << 20);
    for (size_t o = 0; o < n; o += chunk) {
        std::string e = one_member(method, level, p + o, std::min(chunk, n - o), out);
        if (!e.empty()) return e;
    }
    return std::string();
}

}  // namespace jlp
