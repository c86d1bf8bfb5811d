// Developer tool: what the GPU box's host memory gives the upload path.  Reads a large buffer with T threads (AVX2 xor/or like the
// uniform-chunk scan of hcm_upload_chunks) and prints GB/s per thread count -- the ceiling of the scan, and together with the DMA
// traffic of a step the ceiling of `e2e`.
//   g++ -O3 -march=x86-64-v3 -pthread -o tools/bin/host_bw tools/host_bw.cpp && tools/bin/host_bw [GiB]
#include <immintrin.h>

#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

int main(int argc, char **argv) {
	const size_t bytes = (size_t)(argc > 1 ? atof(argv[1]) : 4.0) * (1ull << 30);
	uint8_t *buf = (uint8_t *)aligned_alloc(4096, bytes);
	memset(buf, 1, bytes);
	const unsigned hw = std::thread::hardware_concurrency();
	printf("host cores %u, buffer %.1f GiB\n", hw, bytes / double(1ull << 30));
	for (unsigned T : {1u, 2u, 4u, 8u, 12u, 15u, 16u, 24u, 32u}) {
		if (T > hw)
			break;
		double best = 0;
		for (int rep = 0; rep < 3; ++rep) {
			std::atomic<uint64_t> sink{0};
			auto t0 = std::chrono::steady_clock::now();
			std::vector<std::thread> ts;
			for (unsigned t = 0; t < T; ++t)
				ts.emplace_back([&, t] {
					const size_t lo = bytes / T * t / 128 * 128, hi = bytes / T * (t + 1) / 128 * 128; // aligned loads
					__m256i acc = _mm256_setzero_si256();
					for (size_t i = lo; i + 128 <= hi; i += 128) {
						_mm_prefetch((const char *)buf + i + 4096, _MM_HINT_T0);
						_mm_prefetch((const char *)buf + i + 4096 + 64, _MM_HINT_T0);
						acc = _mm256_or_si256(acc, _mm256_load_si256((const __m256i *)(buf + i)));
						acc = _mm256_or_si256(acc, _mm256_load_si256((const __m256i *)(buf + i + 32)));
						acc = _mm256_or_si256(acc, _mm256_load_si256((const __m256i *)(buf + i + 64)));
						acc = _mm256_or_si256(acc, _mm256_load_si256((const __m256i *)(buf + i + 96)));
					}
					sink += (uint64_t)_mm256_extract_epi64(acc, 0);
				});
			for (auto &t : ts)
				t.join();
			const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
			best = std::max(best, bytes / s / 1e9);
		}
		printf("threads %2u: %7.1f GB/s\n", T, best);
	}
	free(buf);
	return 0;
}
