// Stress test of csrc/lsr_copy_pool.h (the host threads behind the staged pageable path of lwe_commit_batch):
// three caller threads, sizes around every slicing boundary.  Test infrastructure.
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>

#include "lsr_copy_pool.h"

int main() {
    auto work = [](unsigned seed) {
        std::vector<char> a(24 << 20), b(24 << 20);
        for (size_t i = 0; i < a.size(); i++) a[i] = (char)(i * 31 + seed);
        const size_t sizes[] = {0, 1, 4095, 1 << 20, (1 << 20) + 1, 5000000, (8 << 20) + 7, (24 << 20) - 3, 24 << 20};
        for (int rep = 0; rep < 10; rep++)
            for (size_t s : sizes) {
                std::fill(b.begin(), b.begin() + (s < 64 ? 64 : s), 0);
                lsr::CopyPool::get().copy(b.data(), a.data(), s);
                if (memcmp(a.data(), b.data(), s)) { fprintf(stderr, "MISMATCH %zu\n", s); _Exit(1); }
            }
    };
    std::thread t1(work, 1), t2(work, 2), t3(work, 3);
    t1.join(); t2.join(); t3.join();
    fprintf(stderr, "copy pool ok\n");
    return 0;
}
