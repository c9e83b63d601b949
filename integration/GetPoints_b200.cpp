// GetPoints_b200.cpp -- K1 behind the reference's histogram producer: Runner::get_points (src/cluster/CRunner.cpp:504-544)
// hands the ChromosomeOneDigitDna objects of one FASTA file to mc2_batched_get_points, which builds for all of them at once
// what Loader<T>::get_point (src/clutil/Loader.cpp:138-179) builds one by one:
//   k-mer table (T, init 1) + 1-mer table (u64, init 1) over the segments, DivergencePoint<T>(values, size), header, effective
//   length, sequence string, k, stddev, id.
// The ChromosomeOneDigit objects already hold exactly the input contract of mc2_seqs_upload (codes 0..3 inside the segments,
// inclusive [start, end] segment lists), so nothing is re-encoded.  Histograms, 1-mers, magnitudes and lengths
// are bit-identical to the reference's (tests/test_gpu_count.py against the oracle and the golden vectors); stddev is
// computed on the device from the exact integer identity sqrt(N * sum p^2 - (sum p)^2) / N, which agrees with the reference's
// floating-point loop to 1e-12 relative (it feeds only `extraslow` singles, which are out of scope).
//
// Loader.cpp keeps its "histogram type too small" counter in a file-static that nothing outside can reach; sequences whose
// device histogram reports an overflowing segment are therefore ALSO passed through the reference's Loader<T>::fill_table,
// whose values are discarded -- its only retained effect is that counter, so get_warning() prints what it always printed.
#include "get_points_b200.h"

#include <cctype>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <sys/stat.h>

#include "clutil/DivergencePoint.h"
#include "clutil/Loader.h"
#include "exception/InvalidInputException.h"
#include "nonltr/ChromosomeOneDigit.h"
#include "nonltr/ChromosomeOneDigitDna.h"
#include "nonltr/KmerHashTable.h"

#include <omp.h>

#include "build_points.h"
#include "device_b200.h"
#include "fasta_records.h"

namespace {

// owning handles: the device objects are released on every path out of a scope, exceptions included
struct SeqsHandle {
	mc2_seqs *p = nullptr;
	~SeqsHandle()
	{
		if (p) {
			mc2_seqs_free(p);
		}
	}
};
struct HsetHandle {
	mc2_hset *p = nullptr;
	~HsetHandle()
	{
		if (p) {
			mc2_hset_free(p);
		}
	}
};


template <class T>
struct DeviceWidth {
	static const int bytes = 0; // int / double histograms: not a device width
};
template <>
struct DeviceWidth<uint8_t> {
	static const int bytes = 1;
};
template <>
struct DeviceWidth<uint16_t> {
	static const int bytes = 2;
};
template <>
struct DeviceWidth<uint32_t> {
	static const int bytes = 4;
};
template <>
struct DeviceWidth<uint64_t> {
	static const int bytes = 8;
};

// MC2_TIMING=1: phases of the batched readers on stderr (development aid)
struct Phase {
	const bool on = std::getenv("MC2_TIMING") != nullptr;
	std::chrono::steady_clock::time_point t = std::chrono::steady_clock::now();
	void mark(const char *what)
	{
		if (on) {
			const auto now = std::chrono::steady_clock::now();
			std::fprintf(stderr, "meshclust2_b200 timing: %-28s %.3f s\n", what, std::chrono::duration<double>(now - t).count());
			t = now;
		}
	}
};

uint64_t min_bases()
{
	static const char *min_env = std::getenv("MC2_K1_MIN_BASES");
	return min_env ? std::strtoull(min_env, nullptr, 10) : (32ull << 20);
}

} // namespace

template <class T>
bool mc2_batched_get_points(const std::vector<Chromosome *> &chroms, uintmax_t &id, int k, std::vector<Point<T> *> &points)
{
	const int eb = DeviceWidth<T>::bytes;
	if (!mc2i::batching_enabled() || eb == 0) {
		return false;
	}
	const uint64_t n = chroms.size();
	if (n == 0) {
		return true;
	}
	const size_t N = (size_t)1 << (2 * k);
	// the input contract, gathered from the objects the reader produced
	std::vector<ChromosomeOneDigit *> cs(n);
	std::vector<uint64_t> seq_off(n + 1, 0), seg_off(n + 1, 0);
	std::vector<int32_t> segs;
	uint64_t total = 0;
	for (uint64_t i = 0; i < n; i++) {
		cs[i] = dynamic_cast<ChromosomeOneDigit *>(chroms[i]);
		if (!cs[i]) {
			throw std::runtime_error("meshclust2_b200 integration: get_points expects ChromosomeOneDigit objects");
		}
		total += cs[i]->getBase()->size();
		seq_off[i + 1] = total;
		for (const std::vector<int> *s : *cs[i]->getSegment()) {
			segs.push_back(s->at(0));
			segs.push_back(s->at(1));
		}
		seg_off[i + 1] = segs.size() / 2;
	}
	// Below a few tens of megabases the reference's reader + host loop (~90 ns per base over its three passes) finish before
	// the CUDA context is even up (one to two seconds, otherwise hidden behind the FASTA parse and the GLM fit): decline, and
	// let the device start where it pays.  MC2_K1_MIN_BASES overrides the threshold (default 32 Mi; 0: always on the device
	// -- what tests/test_integrated_cluster.py runs).
	if (total < min_bases()) {
		return false;
	}
	std::string codes;
	codes.reserve(total);
	for (uint64_t i = 0; i < n; i++) {
		codes += *cs[i]->getBase();
	}
	std::vector<T> bins(n * N);
	std::vector<uint64_t> mag(n), len(n), mers1(4 * n);
	std::vector<double> stddev(n);
	std::vector<int32_t> novf(n);
	std::vector<uint32_t> maxc(n);
	{
		std::lock_guard<std::mutex> lock(mc2i::device_mutex());
		mc2_ctx *ctx = mc2i::shared_ctx();
		mc2_seqs *sq = nullptr;
		mc2_hset *hs = nullptr;
		mc2i::ok(mc2_seqs_upload(ctx, codes.data(), seq_off.data(), n, segs.data(), seg_off.data(), &sq));
		int rc = mc2_count_kmers(ctx, sq, k, eb, &hs);
		if (rc == MC2_OK) {
			rc = mc2_hset_download(ctx, hs, 0, n, bins.data(), mag.data(), len.data(), mers1.data(), stddev.data(), novf.data(),
					       maxc.data());
		}
		mc2_hset_free(hs);
		mc2_seqs_free(sq);
		mc2i::ok(rc);
	}
	// the host objects, as Loader<T>::get_point leaves them (Loader.cpp:151-175)
	std::vector<Point<T> *> made(n);
	for (uint64_t i = 0; i < n; i++) {
		std::vector<T> values(bins.begin() + i * N, bins.begin() + (i + 1) * N);
		Point<T> *p = new DivergencePoint<T>(values, cs[i]->size());
		p->set_1mers(std::vector<uint64_t>(mers1.begin() + 4 * i, mers1.begin() + 4 * i + 4));
		p->set_header(cs[i]->getHeader());
		p->set_length(cs[i]->getEffectiveSize());
		p->set_data_str(*cs[i]->getBase());
		p->setK(k);
		dynamic_cast<DivergencePoint<T> *>(p)->set_stddev(stddev[i]);
		if ((uint64_t)cs[i]->getEffectiveSize() != len[i]) {
			throw std::runtime_error("meshclust2_b200 integration: effective length differs from the segment list");
		}
		if (novf[i] > 0) {
			KmerHashTable<unsigned long, T> table(k, 1);
			std::vector<T> discarded;
			Loader<T>::fill_table(table, cs[i], discarded); // bumps Loader.cpp's private overflow counter, nothing else kept
		}
		made[i] = p;
	}
#pragma omp critical
	{
		for (uint64_t i = 0; i < n; i++) {
			made[i]->set_id(id);
			id++;
			points.push_back(made[i]);
		}
	}
	return true;
}

template bool mc2_batched_get_points<uint8_t>(const std::vector<Chromosome *> &, uintmax_t &, int, std::vector<Point<uint8_t> *> &);
template bool mc2_batched_get_points<uint16_t>(const std::vector<Chromosome *> &, uintmax_t &, int, std::vector<Point<uint16_t> *> &);
template bool mc2_batched_get_points<uint32_t>(const std::vector<Chromosome *> &, uintmax_t &, int, std::vector<Point<uint32_t> *> &);
template bool mc2_batched_get_points<uint64_t>(const std::vector<Chromosome *> &, uintmax_t &, int, std::vector<Point<uint64_t> *> &);

// ---- FASTA file -> records (host) -> the input contract on the device ------------------------------------------------------
namespace {

// The reference reads each input file three times (find_k, the width detection of Runner::run, get_points); the records are
// split once (integration/fasta_records.h) and kept until get_points has consumed them.
using mc2i::FileRecords;

// per file: split once, by whichever caller comes first; other files load concurrently (get_points runs an `omp parallel
// for` over files)
struct FileSlot {
	std::once_flag once;
	std::shared_ptr<FileRecords> rec; // nullptr = declined (remembered), or consumed by get_points
};
std::mutex g_files_mu;
std::map<std::string, std::shared_ptr<FileSlot>> g_files;

// nullptr = declined: MC2_NO_BATCH=1, MC2_NO_DEVICE_READER=1, --single-file, a file below MC2_K1_MIN_BASES bytes, or a file
// whose shape trips the reader's own corner cases (bases before any header, a record without bases)
std::shared_ptr<FileRecords> load_records(const std::string &fasta, bool is_single_file)
{
	if (!mc2i::batching_enabled() || is_single_file || std::getenv("MC2_NO_DEVICE_READER")) {
		return nullptr;
	}
	std::shared_ptr<FileSlot> slot;
	{
		std::lock_guard<std::mutex> lock(g_files_mu);
		std::shared_ptr<FileSlot> &s = g_files[fasta];
		if (!s) {
			s = std::make_shared<FileSlot>();
		}
		slot = s;
	}
	std::call_once(slot->once, [&]() {
		struct stat st;
		if (stat(fasta.c_str(), &st) != 0 || (uint64_t)st.st_size < min_bases()) {
			return;
		}
		std::string raw((size_t)st.st_size, '\0');
		std::ifstream in(fasta.c_str(), std::ios::binary);
		if (!in.read(&raw[0], st.st_size)) {
			return;
		}
		std::shared_ptr<FileRecords> rec = std::make_shared<FileRecords>();
		if (mc2i::split_fasta(raw, *rec)) {
			std::lock_guard<std::mutex> lock(g_files_mu);
			slot->rec = rec;
		}
	});
	std::lock_guard<std::mutex> lock(g_files_mu);
	return slot->rec;
}

void drop_records(const std::string &fasta)
{
	std::lock_guard<std::mutex> lock(g_files_mu);
	auto it = g_files.find(fasta);
	if (it != g_files.end() && it->second) {
		it->second->rec.reset(); // consumed: a later reader of the same file takes the reference's path
	}
}

} // namespace

// Runner::find_k's per-file body (src/cluster/CRunner.cpp:484-493): the sum of getEffectiveSize() over makeChromList()'s
// Chromosome objects and their number.  makeChromList pre-fills each Chromosome with `size` blanks and then APPENDS the
// sequence (ChromListMaker.cpp:72,87 vs Chromosome.cpp:18-25, 88-97), so the segmentation runs over
// blanks + sequence.  Only N versus not-N matters to it, so the device segments `A...A + sequence` instead.
bool mc2_batched_effective_length(const std::string &fasta, bool is_single_file, unsigned long long &sum_effective,
				  unsigned long long &n_records)
{
	std::shared_ptr<FileRecords> rec = load_records(fasta, is_single_file);
	if (!rec) {
		return false;
	}
	const uint64_t n = rec->headers.size();
	std::string doubled;
	std::vector<uint64_t> off;
	mc2i::doubled_for_find_k(*rec, doubled, off);
	std::vector<int32_t> segs;
	std::vector<uint64_t> seg_off(n + 1);
	{
		std::lock_guard<std::mutex> lock(mc2i::device_mutex());
		mc2_ctx *ctx = mc2i::shared_ctx();
		SeqsHandle sq;
		mc2i::ok(mc2_seqs_from_text(ctx, doubled.data(), off.data(), n, &sq.p));
		segs.resize(2 * mc2_seqs_total_segments(sq.p));
		mc2i::ok(mc2_seqs_download_segments(ctx, sq.p, segs.data(), seg_off.data(), nullptr));
	}
	sum_effective = 0;
	for (size_t s2 = 0; s2 + 1 < segs.size(); s2 += 2) {
		sum_effective += (unsigned long long)(segs[s2 + 1] - segs[s2] + 1);
	}
	n_records = n;
	return true;
}

// The per-file body of Runner::run's width detection (src/cluster/CRunner.cpp:61-81): the largest entry of the u64 k-mer
// tables (initial value 1) over the file's sequences.  mc2_count_kmers_auto counts once in 32-bit shared bins and returns it;
// a segment shorter than k makes it decline (the reference's pass has no length guard there).
bool mc2_batched_largest_count(const std::string &fasta, bool is_single_file, int k, uint64_t &largest)
{
	std::shared_ptr<FileRecords> rec = load_records(fasta, is_single_file);
	if (!rec) {
		return false;
	}
	std::lock_guard<std::mutex> lock(mc2i::device_mutex());
	mc2_ctx *ctx = mc2i::shared_ctx();
	SeqsHandle sq;
	HsetHandle hs;
	int rc = mc2_seqs_from_text(ctx, rec->text.data(), rec->seq_off.data(), rec->headers.size(), &sq.p);
	if (rc == MC2_ERR_INPUT) {
		throw InvalidInputException(std::string("Invalid nucleotide: ") + mc2_last_error());
	}
	mc2i::ok(rc);
	// (when the largest count exceeds 255, mc2_count_kmers_auto counts a second time at the wider width; that set is not
	// needed here, but the call is the one that declines on segments shorter than k)
	int eb = 0;
	rc = mc2_count_kmers_auto(ctx, sq.p, k, &largest, &eb, &hs.p);
	if (rc == MC2_ERR_INPUT) {
		return false;
	}
	mc2i::ok(rc);
	return true;
}

// The whole per-file body of get_points (see get_points_b200.h)
template <class T>
bool mc2_batched_read_points(const std::string &fasta, bool is_single_file, uintmax_t &id, int k, std::vector<Point<T> *> &points)
{
	const int eb = DeviceWidth<T>::bytes;
	if (eb == 0) {
		return false;
	}
	Phase ph;
	std::shared_ptr<FileRecords> rec = load_records(fasta, is_single_file);
	if (!rec) {
		return false;
	}
	const std::vector<std::string> &headers = rec->headers;
	const std::vector<uint64_t> &seq_off = rec->seq_off;
	const std::string &text = rec->text;
	const uint64_t n = headers.size();
	ph.mark("read + split records");
	const size_t N = (size_t)1 << (2 * k);
	std::vector<T> bins(n * N);
	std::vector<uint64_t> mag(n), len(n), mers1(4 * n), seg_off(n + 1);
	std::vector<double> stddev(n);
	std::vector<int32_t> novf(n), segs;
	std::vector<uint32_t> maxc(n);
	{
		std::lock_guard<std::mutex> lock(mc2i::device_mutex());
		mc2_ctx *ctx = mc2i::shared_ctx();
		ph.mark("wait for the CUDA context");
		SeqsHandle sqh;
		HsetHandle hsh;
		mc2_seqs *&sq = sqh.p;
		mc2_hset *&hs = hsh.p;
		int rc = mc2_seqs_from_text(ctx, text.data(), seq_off.data(), n, &sq);
		ph.mark("mc2_seqs_from_text");
		if (rc == MC2_ERR_INPUT) {
			throw InvalidInputException(std::string("Invalid nucleotide: ") + mc2_last_error());
		}
		mc2i::ok(rc);
		segs.resize(2 * mc2_seqs_total_segments(sq));
		rc = mc2_seqs_download_segments(ctx, sq, segs.data(), seg_off.data(), nullptr);
		ph.mark("segments back");
		if (rc == MC2_OK) {
			rc = mc2_count_kmers(ctx, sq, k, eb, &hs);
		}
		ph.mark("mc2_count_kmers");
		if (rc == MC2_OK) {
			rc = mc2_hset_download(ctx, hs, 0, n, bins.data(), mag.data(), len.data(), mers1.data(), stddev.data(), novf.data(),
					       maxc.data());
		}
		ph.mark("histograms back");
		mc2i::ok(rc);
	}
	// the host objects (integration/build_points.h)
	std::vector<Point<T> *> made;
	const std::string bad = mc2i::build_points<T>(*rec, k, bins.data(), len.data(), mers1.data(), stddev.data(), segs.data(),
						      seg_off.data(), (unsigned)omp_get_max_threads(), made);
	ph.mark("host point objects");
	if (!bad.empty()) {
		throw InvalidInputException(bad);
	}
	for (uint64_t i = 0; i < n; i++) {
		if (novf[i] > 0) {
			// rare: rebuild the reader's object for this one record so that Loader.cpp's private overflow counter moves
			std::string line(text, seq_off[i], seq_off[i + 1] - seq_off[i]);
			ChromosomeOneDigitDna chrom((uint64_t)line.size());
			std::string header = headers[i];
			chrom.setHeader(header);
			chrom.insert(line);
			chrom.finalize();
			KmerHashTable<unsigned long, T> table(k, 1);
			std::vector<T> discarded;
			Loader<T>::fill_table(table, &chrom, discarded);
		}
	}
#pragma omp critical
	{
		for (uint64_t i = 0; i < n; i++) {
			made[i]->set_id(id);
			id++;
			points.push_back(made[i]);
		}
	}
	rec.reset();
	drop_records(fasta);
	return true;
}

template bool mc2_batched_read_points<uint8_t>(const std::string &, bool, uintmax_t &, int, std::vector<Point<uint8_t> *> &);
template bool mc2_batched_read_points<uint16_t>(const std::string &, bool, uintmax_t &, int, std::vector<Point<uint16_t> *> &);
template bool mc2_batched_read_points<uint32_t>(const std::string &, bool, uintmax_t &, int, std::vector<Point<uint32_t> *> &);
template bool mc2_batched_read_points<uint64_t>(const std::string &, bool, uintmax_t &, int, std::vector<Point<uint64_t> *> &);
