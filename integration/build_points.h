// build_points.h -- the host objects of the device reader (integration/GetPoints_b200.cpp), header-only so that the CPU test
// (tests/cpp/test_build_points.cpp) can hold them against Loader<T>::get_point.  Given what the device returned for the
// records of one file, builds the DivergencePoint<T> objects exactly as Loader<T>::get_point leaves them
// (src/clutil/Loader.cpp:151-175): values + size, 1-mers, header, effective length, the encoded sequence string, k, stddev.
// The ids are assigned by the caller.
#ifndef MC2_BUILD_POINTS_H
#define MC2_BUILD_POINTS_H

#include <algorithm>
#include <atomic>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "clutil/DivergencePoint.h"
#include "clutil/Point.h"

#include "fasta_records.h"

namespace mc2i {

// The caller sits inside get_points' `omp parallel for` over files, where a nested OpenMP region would get one thread; plain
// std::thread workers over blocks of 256 records do not care.  Returns "" or the reference's message for a record whose
// letters ChromosomeOneDigit::encode would have rejected (the points are then deleted again).
template <class T>
std::string build_points(const FileRecords &rec, int k, const T *bins, const uint64_t *len, const uint64_t *mers1,
			 const double *stddev, const int32_t *segs, const uint64_t *seg_off, unsigned n_threads,
			 std::vector<Point<T> *> &made)
{
	const uint64_t n = rec.headers.size();
	const size_t N = (size_t)1 << (2 * k);
	made.assign(n, nullptr);
	std::string bad;
	std::mutex bad_mu;
	std::atomic<uint64_t> next(0);
	const uint64_t block = 256;
	auto work = [&]() {
		for (;;) {
			const uint64_t b0 = next.fetch_add(block);
			if (b0 >= n) {
				return;
			}
			const uint64_t b1 = std::min(n, b0 + block);
			for (uint64_t i = b0; i < b1; i++) {
				std::string data(rec.text, rec.seq_off[i], rec.seq_off[i + 1] - rec.seq_off[i]);
				const char bad_letter = encode_data_string(data, segs + 2 * seg_off[i], seg_off[i + 1] - seg_off[i]);
				if (bad_letter) {
					std::lock_guard<std::mutex> lock(bad_mu);
					bad = std::string("ChromosomeOneDigit::encode() found invalid letter: ") + bad_letter;
				}
				std::vector<T> values(bins + i * N, bins + (i + 1) * N);
				DivergencePoint<T> *p = new DivergencePoint<T>(values, data.size());
				p->set_1mers(std::vector<uint64_t>(mers1 + 4 * i, mers1 + 4 * i + 4));
				p->set_header(rec.headers[i]);
				p->set_length(len[i]);
				p->set_data_str(data);
				p->setK(k);
				p->set_stddev(stddev[i]);
				made[i] = p;
			}
		}
	};
	const unsigned workers = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(n_threads ? n_threads : 1, (n + block - 1) / block));
	std::vector<std::thread> pool;
	for (unsigned t = 1; t < workers; t++) {
		pool.emplace_back(work);
	}
	work();
	for (std::thread &t : pool) {
		t.join();
	}
	if (!bad.empty()) {
		for (Point<T> *p : made) {
			delete p;
		}
		made.clear();
	}
	return bad;
}

} // namespace mc2i

#endif
