// micro-benchmark: N threads filling 2 files on tmpfs, pwrite vs mmap+memcpy
#include <fcntl.h>
#include <sys/mman.h>
#include <unistd.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
#include <atomic>
int main(int argc, char** argv) {
    int nthr = atoi(argv[1]); int mode = atoi(argv[2]); const char* dir = argv[3];
    size_t per_file = (size_t)atol(argv[4]) << 20;
    size_t slice = 8u << 20;
    std::vector<uint8_t> src(per_file, 'A');
    for (size_t i = 0; i < per_file; i += 4096) src[i] = (uint8_t)i;
    int fd[2];
    for (int e = 0; e < 2; e++) { char nm[256]; snprintf(nm, 256, "%s/wb_%d", dir, e); fd[e] = open(nm, O_RDWR | O_CREAT | O_TRUNC, 0644); }
    auto t0 = std::chrono::steady_clock::now();
    if (mode >= 1) for (int e = 0; e < 2; e++) if (ftruncate(fd[e], per_file)) return 1;
    std::atomic<size_t> next{0};
    size_t ns = (per_file + slice - 1) / slice;
    std::vector<std::thread> th;
    for (int t = 0; t < nthr; t++) th.emplace_back([&]() {
        for (;;) {
            size_t k = next.fetch_add(1); if (k >= 2 * ns) break;
            int e = k & 1; size_t o = (k >> 1) * slice; size_t len = std::min(slice, per_file - o);
            if (mode == 0 || mode == 1) { size_t d = 0; while (d < len) { ssize_t w = pwrite(fd[e], src.data() + o + d, len - d, o + d); if (w <= 0) exit(2); d += w; } }
            else { void* p = mmap(nullptr, len, PROT_READ | PROT_WRITE, MAP_SHARED | (mode == 3 ? MAP_POPULATE : 0), fd[e], o); if (p == MAP_FAILED) exit(3); memcpy(p, src.data() + o, len); munmap(p, len); }
        }
    });
    for (auto& t : th) t.join();
    double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    printf("threads %d mode %d: %.2f GB/s\n", nthr, mode, 2.0 * per_file / s / 1e9);
    for (int e = 0; e < 2; e++) { close(fd[e]); char nm[256]; snprintf(nm, 256, "%s/wb_%d", dir, e); unlink(nm); }
}
