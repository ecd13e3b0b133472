// micro-benchmark: how fast fresh tmpfs pages can be produced for two output files.
//   mode 0  posix_fallocate of both files, one after the other, on one thread
//   mode 1  posix_fallocate of the two files on two threads
//   mode 2  posix_fallocate of 64 MiB pieces of both files from N threads
//   mode 3  ftruncate, then N threads touch one byte per page through mappings (page faults only)
//   mode 4  ftruncate, then N threads memcpy 8 MiB slices through mappings (faults + copy)
//   mode 5  mode 0, then mode 4's copy (pages exist already)
//   mode 6  N threads pwrite 8 MiB slices
//   mode 7  one thread per file runs posix_fallocate in 352 MiB pieces while N threads copy the slices whose pages exist
//   mode 8  posix_fallocate of the two files on two threads, then N threads copy
//   mode 9  per 352 MiB batch: two threads fallocate batch k + 1 of the two files while N threads copy batch k
// usage: fill_probe <threads> <mode> <dir> <MiB per file>
#include <fcntl.h>
#include <sys/mman.h>
#include <unistd.h>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main(int argc, char** argv) {
    if (argc < 5) return 1;
    const int nthr = atoi(argv[1]), mode = atoi(argv[2]);
    const char* dir = argv[3];
    const size_t per_file = (size_t)atol(argv[4]) << 20, slice = 8u << 20, piece = 64u << 20;
    std::vector<uint8_t> src(slice * 4, 'A');
    int fd[2];
    char nm[2][256];
    for (int e = 0; e < 2; e++) { snprintf(nm[e], 256, "%s/fp_%d", dir, e); fd[e] = open(nm[e], O_RDWR | O_CREAT | O_TRUNC, 0644); if (fd[e] < 0) return 2; }
    auto par = [&](int n, size_t items, auto f) {
        std::atomic<size_t> next{0};
        std::vector<std::thread> th;
        for (int t = 0; t < n; t++) th.emplace_back([&]() { for (size_t k = next.fetch_add(1); k < items; k = next.fetch_add(1)) f(k); });
        for (auto& t : th) t.join();
    };
    const double t0 = now();
    double t_alloc = 0;
    if (mode == 0 || mode == 5) { for (int e = 0; e < 2; e++) if (posix_fallocate(fd[e], 0, per_file)) return 3; t_alloc = now() - t0; }
    if (mode == 1) par(2, 2, [&](size_t e) { if (posix_fallocate(fd[e], 0, per_file)) exit(3); });
    if (mode == 2) { const size_t np = (per_file + piece - 1) / piece; par(nthr, 2 * np, [&](size_t k) { const size_t o = (k >> 1) * piece; if (posix_fallocate(fd[k & 1], o, std::min(piece, per_file - o))) exit(3); }); }
    if (mode == 3 || mode == 4) for (int e = 0; e < 2; e++) if (ftruncate(fd[e], per_file)) return 4;
    if (mode >= 3 && mode <= 6) {
        const size_t ns = (per_file + slice - 1) / slice;
        par(nthr, 2 * ns, [&](size_t k) {
            const int e = k & 1; const size_t o = (k >> 1) * slice, len = std::min(slice, per_file - o);
            if (mode == 6) { size_t d = 0; while (d < len) { ssize_t w = pwrite(fd[e], src.data() + d, len - d, o + d); if (w <= 0) exit(5); d += w; } return; }
            uint8_t* p = (uint8_t*)mmap(nullptr, len, PROT_READ | PROT_WRITE, MAP_SHARED, fd[e], o);
            if (p == MAP_FAILED) exit(6);
            if (mode == 3) for (size_t i = 0; i < len; i += 4096) p[i] = 1; else memcpy(p, src.data() + (o % (slice * 3)), len);
            munmap(p, len);
        });
    }
    if (mode == 7) {
        std::atomic<size_t> ready[2];
        ready[0] = 0; ready[1] = 0;
        const size_t batch = 352u << 20;
        std::thread fa[2];
        for (int e = 0; e < 2; e++) fa[e] = std::thread([&, e]() {
            for (size_t o = 0; o < per_file; o += batch) { if (posix_fallocate(fd[e], o, std::min(batch, per_file - o))) exit(3); ready[e].store(std::min(per_file, o + batch)); }
        });
        const size_t ns = (per_file + slice - 1) / slice;
        par(nthr, 2 * ns, [&](size_t k) {
            const int e = k & 1; const size_t o = (k >> 1) * slice, len = std::min(slice, per_file - o);
            while (ready[e].load() < o + len) std::this_thread::yield();
            uint8_t* p = (uint8_t*)mmap(nullptr, len, PROT_READ | PROT_WRITE, MAP_SHARED, fd[e], o);
            if (p == MAP_FAILED) exit(6);
            memcpy(p, src.data() + (o % (slice * 3)), len);
            munmap(p, len);
        });
        for (auto& t : fa) t.join();
    }
    if (mode == 8) {
        par(2, 2, [&](size_t e) { if (posix_fallocate(fd[e], 0, per_file)) exit(3); });
        t_alloc = now() - t0;
        const size_t ns = (per_file + slice - 1) / slice;
        par(nthr, 2 * ns, [&](size_t k) {
            const int e = k & 1; const size_t o = (k >> 1) * slice, len = std::min(slice, per_file - o);
            uint8_t* p = (uint8_t*)mmap(nullptr, len, PROT_READ | PROT_WRITE, MAP_SHARED, fd[e], o);
            if (p == MAP_FAILED) exit(6);
            memcpy(p, src.data() + (o % (slice * 3)), len);
            munmap(p, len);
        });
    }
    if (mode == 9) {
        const size_t batch = 352u << 20, nb = (per_file + batch - 1) / batch;
        auto alloc = [&](size_t b) { par(2, 2, [&](size_t e) { const size_t o = b * batch; if (posix_fallocate(fd[e], o, std::min(batch, per_file - o))) exit(3); }); };
        alloc(0);
        for (size_t b = 0; b < nb; b++) {
            std::thread next;
            if (b + 1 < nb) next = std::thread(alloc, b + 1);
            const size_t b0 = b * batch, bl = std::min(batch, per_file - b0), ns = (bl + slice - 1) / slice;
            // one mapping per file and batch (as the library does), slices copied by the workers
            uint8_t* m[2];
            for (int e = 0; e < 2; e++) { m[e] = (uint8_t*)mmap(nullptr, bl, PROT_READ | PROT_WRITE, MAP_SHARED, fd[e], b0); if (m[e] == MAP_FAILED) exit(6); }
            par(nthr, 2 * ns, [&](size_t k) {
                const int e = k & 1; const size_t o = (k >> 1) * slice, len = std::min(slice, bl - o);
                memcpy(m[e] + o, src.data() + (o % (slice * 3)), len);
            });
            for (int e = 0; e < 2; e++) munmap(m[e], bl);
            if (next.joinable()) next.join();
        }
    }
    const double s = now() - t0;
    printf("threads %2d mode %d: %6.2f GB/s (%.3f s; fallocate part %.3f s)\n", nthr, mode, 2.0 * per_file / s / 1e9, s, t_alloc);
    for (int e = 0; e < 2; e++) { close(fd[e]); unlink(nm[e]); }
}
