// CPU check of the host objects of the device reader (integration/build_points.h) against Loader<T>::get_point, compiled
// against the reference sources and linked with oracle/_ref/libmc2ref.so (TEST INFRASTRUCTURE).
//   test_build_points <file.fa> <k> <threads>
// The reference reads the file and builds one point per record; the same numbers a device batch would return (bins, length,
// 1-mers, stddev, segment lists -- here taken from the reference's own objects) go through build_points, and every field of
// every point is compared.
#include <fstream>
#include <iostream>
#include <sstream>

#include "clutil/DivergencePoint.h"
#include "clutil/Loader.h"
#include "nonltr/ChromListMaker.h"
#include "nonltr/ChromosomeOneDigitDna.h"

#include "build_points.h"

using namespace nonltr;

template <class T>
int run(const char *path, int k, unsigned threads)
{
	std::ifstream in(path, std::ios::binary);
	std::stringstream ss;
	ss << in.rdbuf();
	mc2i::FileRecords rec;
	if (!mc2i::split_fasta(ss.str(), rec)) {
		std::cout << "DECLINED" << std::endl;
		return 0;
	}
	const uint64_t n = rec.headers.size();
	const size_t N = (size_t)1 << (2 * k);
	ChromListMaker maker(path, false);
	const std::vector<Chromosome *> *list = maker.makeChromOneDigitDnaList();
	if (list->size() != n) {
		std::cout << "FAIL record count" << std::endl;
		return 1;
	}
	std::vector<Point<T> *> want(n);
	std::vector<T> bins(n * N);
	std::vector<uint64_t> len(n), mers1(4 * n), seg_off(n + 1, 0);
	std::vector<double> stddev(n);
	std::vector<int32_t> segs;
	uintmax_t id = 0;
	for (uint64_t i = 0; i < n; i++) {
		ChromosomeOneDigit *c = dynamic_cast<ChromosomeOneDigit *>(list->at(i));
		want[i] = Loader<T>::get_point(c, id, k);
		const DivergencePoint<T> *d = dynamic_cast<const DivergencePoint<T> *>(want[i]);
		std::copy(d->points.begin(), d->points.end(), bins.begin() + i * N);
		len[i] = want[i]->get_length();
		const std::vector<uint64_t> m = want[i]->get_1mers();
		std::copy(m.begin(), m.end(), mers1.begin() + 4 * i);
		stddev[i] = d->get_stddev();
		for (const std::vector<int> *s : *c->getSegment()) {
			segs.push_back(s->at(0));
			segs.push_back(s->at(1));
		}
		seg_off[i + 1] = segs.size() / 2;
	}
	std::vector<Point<T> *> made;
	const std::string bad = mc2i::build_points<T>(rec, k, bins.data(), len.data(), mers1.data(), stddev.data(), segs.data(),
						      seg_off.data(), threads, made);
	if (!bad.empty() || made.size() != n) {
		std::cout << "FAIL build_points: " << bad << std::endl;
		return 1;
	}
	for (uint64_t i = 0; i < n; i++) {
		const DivergencePoint<T> *a = dynamic_cast<const DivergencePoint<T> *>(want[i]);
		const DivergencePoint<T> *b = dynamic_cast<const DivergencePoint<T> *>(made[i]);
		const bool same = b && a->points == b->points && a->getPseudoMagnitude() == b->getPseudoMagnitude() &&
				  a->getRealMagnitude() == b->getRealMagnitude() && want[i]->get_length() == made[i]->get_length() &&
				  want[i]->get_header() == made[i]->get_header() && want[i]->get_data_str() == made[i]->get_data_str() &&
				  want[i]->get_1mers() == made[i]->get_1mers() && want[i]->getK() == made[i]->getK() &&
				  a->get_stddev() == b->get_stddev() && want[i]->size() == made[i]->size();
		if (!same) {
			std::cout << "FAIL point " << i << " (" << rec.headers[i] << ")" << std::endl;
			return 1;
		}
	}
	std::cout << "OK " << n << " points" << std::endl;
	return 0;
}

int main(int argc, char **argv)
{
	if (argc < 4) {
		return 2;
	}
	const int k = atoi(argv[2]);
	const unsigned threads = (unsigned)atoi(argv[3]);
	const int rc8 = run<uint8_t>(argv[1], k, threads);
	return rc8 ? rc8 : run<uint16_t>(argv[1], k, threads);
}
