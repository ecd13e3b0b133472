// Compressed FASTQ output on the host (zlib): independent deflate members, so that the
// writer threads compress the slices of a batch in parallel.
//   bgzip: BGZF blocks of at most 0xff00 input bytes, each a gzip member carrying the BC
//          extra field with its size, plus the 28-byte EOF block at the end of the file --
//          what htslib's bgzf_write / bgzf_close produce for FileBGZF
//          (/root/reference/src/io.h:58-135) and for bgzip_file (src/hts.h:140-180);
//   gzip:  plain gzip members of 1 MiB of input; concatenated members are one valid .gz
//          file with the same contents FileGZ's gzwrite stream has (src/io.h:140-236).
#ifndef JLP_DEFLATE_H
#define JLP_DEFLATE_H

#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

namespace jlp {

enum DeflateMethod { DEFLATE_GZIP = 0, DEFLATE_BGZF = 1 };

// Append the compressed form of [p, p + n) to `out`.  Returns an error text, empty on success.
std::string deflate_members(int method, int level, const uint8_t* p, size_t n, std::vector<uint8_t>& out);

// the BGZF end-of-file marker block
extern const uint8_t kBgzfEof[28];

}  // namespace jlp
#endif
