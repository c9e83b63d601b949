// Per-call latency of mc2_get_close as a C++ caller sees it (the relinked meshclust2's accumulate stage issues one call per
// query, each depending on the previous result).  Development aid; build and run on a GPU box from the repo root:
//   g++ -O2 -std=c++17 -Iinclude tools/call_latency.cpp -o tools/call_latency -Lmeshclust2_b200/lib -lmeshclust2_b200 \
//       -Wl,-rpath,$PWD/meshclust2_b200/lib && tools/call_latency tests/golden/weights_cfg1_id90.txt
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "meshclust2_b200.h"

#define OK(x)                                                                  \
	do {                                                                   \
		if ((x) != MC2_OK) {                                           \
			std::fprintf(stderr, "%s: %s\n", #x, mc2_last_error()); \
			return 1;                                              \
		}                                                              \
	} while (0)

int main(int argc, char **argv)
{
	const char *weights = argc > 1 ? argv[1] : "tests/golden/weights_cfg1_id90.txt";
	const uint64_t n = 20000, N = 1024;
	std::mt19937_64 rng(1);
	std::vector<uint8_t> H(n * N);
	for (auto &v : H) v = 1 + rng() % 6;
	std::vector<uint64_t> len(n);
	for (auto &v : len) v = 950 + rng() % 100;
	mc2_ctx *ctx = nullptr;
	OK(mc2_ctx_create(0, &ctx));
	mc2_hset *hs = nullptr;
	OK(mc2_hset_from_host(ctx, H.data(), n, 5, 1, nullptr, len.data(), &hs));
	mc2_model_desc desc;
	int k, eb, mode;
	double id;
	OK(mc2_model_desc_from_file(weights, 0, &desc, &k, &id, &eb, &mode));
	mc2_model *model = nullptr;
	OK(mc2_model_create(ctx, &desc, &model));
	for (uint64_t m : {1, 8, 32, 64, 128, 192, 512}) {
		std::vector<uint64_t> cand(m);
		std::vector<uint8_t> marks(m);
		int64_t best = 0;
		double bd = 0;
		int32_t is_min = 0;
		const int reps = 20000;
		uint64_t q = 7;
		for (int warm = 0; warm < 2; warm++) {
			const auto t0 = std::chrono::steady_clock::now();
			for (int r = 0; r < reps; r++) {
				for (auto &c : cand) c = (q * 2654435761u + (&c - cand.data()) * 97) % n; // depends on the previous answer
				OK(mc2_get_close(ctx, model, hs, q, hs, cand.data(), 0, m, 0.9, &best, &bd, &is_min, marks.data()));
				q = (uint64_t)(best >= 0 ? cand[best] : (q + 1) % n);
			}
			const double us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count() / reps;
			if (warm) std::printf("get_close, %4lu candidates: %6.2f us per call\n", (unsigned long)m, us);
		}
	}
	mc2_model_free(model);
	mc2_hset_free(hs);
	mc2_ctx_destroy(ctx);
	return 0;
}
