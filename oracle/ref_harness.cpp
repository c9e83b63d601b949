// ref_harness.cpp — TEST INFRASTRUCTURE ONLY (parity oracle, "reference" arm of the CPU baseline).
//
// A thin extern "C" shim over the UNMODIFIED MeShClust2 reference classes, compiled together
// with the reference's own sources (from /root/reference, see oracle/Makefile) into
// oracle/_ref/libmc2ref.so.  Nothing in the product (libmeshclust2_b200.so, meshclust2_b200/)
// links, imports or calls this file.  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may load it.
//
// It calls, never re-implements, the reference:
//   Loader<T>::get_point(ChromosomeOneDigit*, id, k)         src/clutil/Loader.cpp:138-179
//   ChromosomeOneDigitDna (encode / segments)                src/nonltr/Chromosome.cpp:130-154
//   Feature<T>::{manhattan,...} static raw singles           src/predict/Feature.cpp
//   Feature<T>::compute / operator()                         src/predict/Feature.h:197-239
//   Predictor<T>(file), close(), similarity(), classify_sum  src/predict/Predictor.cpp:47-79,231-333
//   Trainer<T>::get_close / merge / filter / closest         src/cluster/Trainer.cpp:23-157
//   DivergencePoint<T>::distance / distance_d                src/clutil/DivergencePoint.cpp:55-82
//
// `private` is opened up for Trainer/Predictor only so the harness can install a pinned
// Feature + weight matrix without going through the (nondeterministic, slow) training path
// and without the --recover path that crashes.
#include <cstdint>
#include <cstring>
#include <limits>
#include <string>
#include <vector>
#include <tuple>
#include <fstream>
#include <chrono>
#include <omp.h>

#include <sstream>
#include <iostream>
#include <map>
#include <set>
#include <functional>
#include <algorithm>
#include <random>
#include <numeric>
#include <iterator>
#include <cmath>
#include <cstdlib>
#include <cstdio>
#include <unordered_map>
#include <memory>
#include <thread>
#include <mutex>
#include <iomanip>
#include <list>
#include <queue>
#include <stack>
#include <utility>
#include <exception>
#include <stdexcept>

#define private public
#include "predict/Predictor.h"
#include "cluster/Trainer.h"
#undef private
#include "cluster/ClusterFactory.h"
#include "Loader.h"
#include "DivergencePoint.h"
#include "../predict/Feature.h"
#include "../predict/GLM.h"
#include "../predict/Matrix.h"
#include "ChromosomeOneDigitDna.h"
#include "Center.h"
#include "bvec.h"

namespace {

template <class T>
DivergencePoint<T>* make_point(const T* h, uint64_t N, uint64_t mag_override, uint64_t length, uint64_t id)
{
	std::vector<T> v(h, h + N);
	DivergencePoint<T>* p = nullptr;
	if (mag_override == 0) {
		p = new DivergencePoint<T>(v, length);
	} else {
		// Reproduce the reference's stale-mag state: build a point whose
		// constructor-computed `mag` is the wanted one, then set() the real bins onto it
		// (DivergencePoint::set copies points/length/header/id but not mag, DivergencePoint.cpp:182-190).
		std::vector<T> dummy(N, 0);
		uint64_t left = mag_override;
		const uint64_t cap = (uint64_t)std::numeric_limits<T>::max();
		for (uint64_t i = 0; i < N && left > 0; i++) {
			uint64_t t = left < cap ? left : cap;
			dummy[i] = (T)t;
			left -= t;
		}
		if (left != 0) { return nullptr; }
		p = new DivergencePoint<T>(dummy, length);
		DivergencePoint<T> real(v, length);
		p->set(real);
	}
	p->set_length(length);
	p->set_id(id);
	return p;
}

template <class T>
double raw_single(uint64_t flag, Feature<T>& f, Point<T>& a, Point<T>& b)
{
	switch (flag) {
	case FEAT_MANHATTAN: return Feature<T>::manhattan(a, b);
	case FEAT_EUCLIDEAN: return Feature<T>::euclidean(a, b);
	case FEAT_NORMALIZED_VECTORS: return Feature<T>::normalized_vectors(a, b);
	case FEAT_JEFFEREY_DIV: return Feature<T>::jefferey_divergence(a, b);
	case FEAT_PEARSON_COEFF: return Feature<T>::pearson(a, b);
	case FEAT_INTERSECTION: return Feature<T>::intersection(a, b);
	case FEAT_EMD: return Feature<T>::emd(a, b);
	case FEAT_LENGTHD: return Feature<T>::length_difference(a, b);
	case FEAT_KULCZYNSKI2: return Feature<T>::kulczynski2(a, b);
	case FEAT_SIMRATIO: return Feature<T>::simratio(a, b);
	case FEAT_JENSEN_SHANNON: return f.jensen_shannon(a, b);
	default: return std::numeric_limits<double>::quiet_NaN();
	}
}

struct ModelBase {
	virtual ~ModelBase() {}
	int elem_bytes;
};

template <class T>
struct Model : ModelBase {
	Predictor<T>* pred;   // heap, never destroyed (the file ctor leaves members uninitialised)
	Trainer<T>* trainer;  // carries a copy of the classifier's Feature + weights
};

template <class T>
Model<T>* load_model(const char* weights_file, double cutoff)
{
	Model<T>* m = new Model<T>();
	m->elem_bytes = sizeof(T);
	m->pred = new Predictor<T>(std::string(weights_file));
	// Feature(int k) leaves do_save uninitialised and read_from() builds raw_funcs while it is
	// garbage: when it happens to be non-zero the singles memoise into an unlocked std::map and concurrent
	// close() calls crash.  Pin it to the state every trained Feature handed to clustering has (copy-ctor: false).
	if (m->pred->get_mode() & PRED_MODE_CLASS) {
		m->pred->feat_c->set_save(false);
		m->pred->feat_c->reset_funcs();
	}
	if (m->pred->get_mode() & PRED_MODE_REGR) {
		m->pred->feat_r->set_save(false);
		m->pred->feat_r->reset_funcs();
	}
	std::vector<Point<T>*> none;
	m->trainer = new Trainer<T>(none, 0, 0, cutoff, 0, m->pred->get_k());
	if (m->pred->get_mode() & PRED_MODE_CLASS) {
		auto pr = m->pred->get_class();
		delete m->trainer->feat;
		m->trainer->feat = pr.first;
		m->trainer->feat->set_save(false);
		m->trainer->weights = pr.second.get_weights();
	}
	return m;
}

template <class T>
int get_point_impl(const char* seq, long len, int k, void* hist_out, uint64_t* mers1, uint64_t* mag,
		   uint64_t* length, double* stddev, long* n_overflow_dummy)
{
	ChromosomeOneDigitDna chrom;
	std::string header(">s");
	std::string s(seq, (size_t)len);
	chrom.setHeader(header);
	chrom.appendToSequence(s);
	chrom.finalize();
	uintmax_t id = 0;
	Point<T>* p = Loader<T>::get_point(&chrom, id, k, false);
	DivergencePoint<T>* d = dynamic_cast<DivergencePoint<T>*>(p);
	std::memcpy(hist_out, d->points.data(), d->points.size() * sizeof(T));
	auto om = p->get_1mers();
	for (int i = 0; i < 4; i++) { mers1[i] = om[i]; }
	*mag = d->getPseudoMagnitude();
	*length = d->get_length();
	*stddev = d->get_stddev();
	delete p;
	return 0;
}


template <class T>
int kmer_table_impl(const char* codes, int first, int last, int k, uint64_t init, void* values_out, int* ret)
{
	KmerHashTable<unsigned long, T> table(k, (T)init);
	*ret = table.wholesaleIncrementNoOverflow(codes, first, last);
	std::memcpy(values_out, table.getValues(), table.getMaxTableSize() * sizeof(T));
	return 0;
}

template <class T>
int raw_single_impl(uint64_t flag, int k, uint64_t N, const void* P, const void* Q, uint64_t mop, uint64_t moq,
		    uint64_t len_p, uint64_t len_q, double* out)
{
	Feature<T> f(k);
	DivergencePoint<T>* a = make_point<T>((const T*)P, N, mop, len_p, 1);
	DivergencePoint<T>* b = make_point<T>((const T*)Q, N, moq, len_q, 2);
	if (!a || !b) { return -4; }
	int rc = 0;
	try {
		*out = raw_single<T>(flag, f, *a, *b);
	} catch (...) {
		rc = -1;
	}
	delete a;
	delete b;
	return rc;
}

template <class T>
int distance_impl(uint64_t N, const void* P, const void* Q, uint64_t mop, uint64_t moq, uint64_t* out)
{
	DivergencePoint<T>* a = make_point<T>((const T*)P, N, mop, 1, 1);
	DivergencePoint<T>* b = make_point<T>((const T*)Q, N, moq, 1, 2);
	if (!a || !b) { return -4; }
	*out = a->distance(*b);
	delete a;
	delete b;
	return 0;
}

template <class T>
int distance_d_impl(uint64_t N, const void* P, const double* C, double* out)
{
	DivergencePoint<T>* a = make_point<T>((const T*)P, N, 0, 1, 1);
	std::vector<double> cv(C, C + N);
	DivergencePoint<double> c(cv, 1);
	*out = a->distance_d(c);
	delete a;
	return 0;
}

template <class T>
int score_pairs_impl(ModelBase* mb, int mode, uint64_t N, uint64_t n, const void* H, const uint64_t* mag,
		     const uint64_t* len, uint64_t m, const uint64_t* ia, const uint64_t* ib, double* out_score,
		     double* out_dist, uint8_t* out_close, double* out_cache, int threads, double* seconds)
{
	Model<T>* mm = (Model<T>*)mb;
	std::vector<DivergencePoint<T>*> pts(n);
	const T* h = (const T*)H;
	for (uint64_t i = 0; i < n; i++) {
		pts[i] = make_point<T>(h + i * N, N, mag ? mag[i] : 0, len[i], i);
		if (!pts[i]) { return -4; }
	}
	Feature<T>* feat = mm->trainer->feat;
	const size_t S = feat->get_lookup().size();
	int bad = 0;
	auto t0 = std::chrono::steady_clock::now();
#pragma omp parallel for num_threads(threads) schedule(static)
	for (uint64_t j = 0; j < m; j++) {
		try {
			Point<T>* a = pts[ia[j]];
			Point<T>* b = pts[ib[j]];
			if (mode == 0) {
				auto cache = feat->compute(*a, *b);
				double dist = (*feat)(0, cache);
				double s = mm->trainer->classify(a, b);
				if (out_score) out_score[j] = s;
				if (out_dist) out_dist[j] = dist;
				if (out_close) out_close[j] = round(s) > 0;
				if (out_cache) for (size_t c = 0; c < S; c++) out_cache[j * S + c] = cache[c];
			} else if (mode == 1) {
				out_close[j] = mm->pred->close(a, b);
			} else {
				out_score[j] = mm->pred->similarity(a, b);
			}
		} catch (...) {
#pragma omp atomic write
			bad = 1;
		}
	}
	auto t1 = std::chrono::steady_clock::now();
	if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
	for (auto p : pts) delete p;
	return bad ? -1 : 0;
}

template <class T>
int get_close_impl(ModelBase* mb, uint64_t N, const void* H, const uint64_t* mag, const uint64_t* len,
		   uint64_t q, uint64_t m, const uint64_t* cand, int64_t* best, double* best_dist, int* is_min,
		   uint8_t* marks, int threads)
{
	Model<T>* mm = (Model<T>*)mb;
	const T* h = (const T*)H;
	std::vector<std::vector<std::pair<Point<T>*, bool> > > col(1);
	for (uint64_t j = 0; j < m; j++) {
		uint64_t i = cand[j];
		col[0].push_back(std::make_pair((Point<T>*)make_point<T>(h + i * N, N, mag ? mag[i] : 0, len[i], i), false));
	}
	DivergencePoint<T>* qp = make_point<T>(h + q * N, N, mag ? mag[q] : 0, len[q], q);
	bvec_iterator<T> b(0, 0, &col), e(0, m, &col);
	bool ismin = true;
	omp_set_num_threads(threads);
	auto r = mm->trainer->get_close(qp, b, e, ismin);
	*best = std::get<0>(r) ? (int64_t)std::get<3>(r) : -1;
	*best_dist = std::get<1>(r);
	*is_min = ismin;
	for (uint64_t j = 0; j < m; j++) {
		marks[j] = col[0][j].second;
		delete col[0][j].first;
	}
	delete qp;
	return 0;
}

template <class T>
int filter_impl(ModelBase* mb, uint64_t N, const void* H, const uint64_t* mag, const uint64_t* len, uint64_t c,
		uint64_t m, const uint64_t* members, uint8_t* keep)
{
	Model<T>* mm = (Model<T>*)mb;
	const T* h = (const T*)H;
	std::vector<std::pair<Point<T>*, bool> > vec;
	std::vector<Point<T>*> all;
	for (uint64_t j = 0; j < m; j++) {
		uint64_t i = members[j];
		Point<T>* p = make_point<T>(h + i * N, N, mag ? mag[i] : 0, len[i], j);
		vec.push_back(std::make_pair(p, false));
		all.push_back(p);
		keep[j] = 0;
	}
	DivergencePoint<T>* cp = make_point<T>(h + c * N, N, mag ? mag[c] : 0, len[c], c);
	mm->trainer->filter(cp, vec);
	for (auto& kv : vec) keep[kv.first->get_id()] = 1;
	for (auto p : all) delete p;
	delete cp;
	return 0;
}

template <class T>
int merge_impl(ModelBase* mb, uint64_t N, const void* H, const uint64_t* mag, const uint64_t* len, uint64_t ncen,
	       const uint64_t* rows, long cur, long begin, long last, long* out, int threads)
{
	Model<T>* mm = (Model<T>*)mb;
	const T* h = (const T*)H;
	std::vector<Center<T> > centers;
	std::vector<Point<T>*> nopts;
	centers.reserve(ncen);
	for (uint64_t j = 0; j < ncen; j++) {
		uint64_t i = rows[j];
		DivergencePoint<T>* p = make_point<T>(h + i * N, N, 0, len[i], i);
		centers.emplace_back(p, nopts); // Center clones (recomputing mag)
		if (mag && mag[i]) {
			// install the stale-mag variant as the center itself
			delete centers.back().center;
			centers.back().center = make_point<T>(h + i * N, N, mag[i], len[i], i);
		}
		delete p;
	}
	omp_set_num_threads(threads);
	*out = mm->trainer->merge(centers, cur, begin, last);
	for (auto& c : centers) { delete c.center; c.center = nullptr; }
	return 0;
}

// get_mean's own sequence of calls (src/cluster/ClusterFactory.cpp:338-380) on the reference's objects; the arg-min is
// the sequential (--threads 1) order of its reduction, identical to Trainer::closest (Trainer.cpp:144-157)
template <class T>
int mean_closest_impl(uint64_t N, const void *H, const uint64_t *members, uint64_t n, int64_t *best, double *best_dist,
		      double *mean_out, double *dist_out)
{
	const T *h = (const T *)H;
	std::vector<DivergencePoint<T> *> pts(n);
	for (uint64_t j = 0; j < n; j++) {
		pts[j] = make_point<T>(h + members[j] * N, N, 0, 1, j);
	}
	Point<double> *top = pts[0]->create_double();
	top->zero();
	Point<double> *temp = top->clone();
	for (uint64_t j = 0; j < n; j++) {
		pts[j]->set_arg_to_this_d(*temp);
		*top += *temp;
	}
	*top /= (double)n;
	int64_t b = -1;
	double bd = std::numeric_limits<double>::max();
	for (uint64_t j = 0; j < n; j++) {
		double d = pts[j]->distance_d(*top);
		if (dist_out) dist_out[j] = d;
		if (d < bd) {
			bd = d;
			b = (int64_t)j;
		}
	}
	if (mean_out) {
		const std::vector<double> &m = top->get_data();
		for (uint64_t i = 0; i < N; i++) mean_out[i] = m[i];
	}
	*best = b;
	*best_dist = bd;
	delete top;
	delete temp;
	for (auto p : pts) delete p;
	return 0;
}

template <class T>
int count_batch_impl(const char* text, const uint64_t* off, uint64_t n, int k, void* hist_out, int threads,
		     double* seconds)
{
	const uint64_t N = 1ULL << (2 * k);
	int bad = 0;
	auto t0 = std::chrono::steady_clock::now();
#pragma omp parallel for num_threads(threads) schedule(dynamic, 16)
	for (uint64_t i = 0; i < n; i++) {
		try {
			ChromosomeOneDigitDna chrom;
			std::string header(">s");
			std::string s(text + off[i], (size_t)(off[i + 1] - off[i]));
			chrom.setHeader(header);
			chrom.appendToSequence(s);
			chrom.finalize();
			uintmax_t id = i;
			Point<T>* p = Loader<T>::get_point(&chrom, id, k, false);
			if (hist_out) {
				DivergencePoint<T>* d = dynamic_cast<DivergencePoint<T>*>(p);
				std::memcpy((T*)hist_out + i * N, d->points.data(), N * sizeof(T));
			}
			delete p;
		} catch (...) {
#pragma omp atomic write
			bad = 1;
		}
	}
	auto t1 = std::chrono::steady_clock::now();
	if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
	return bad ? -1 : 0;
}

} // namespace

#define DISPATCH(eb, FN, ...)                          \
	switch (eb) {                                  \
	case 1: return FN<uint8_t>(__VA_ARGS__);       \
	case 2: return FN<uint16_t>(__VA_ARGS__);      \
	case 4: return FN<uint32_t>(__VA_ARGS__);      \
	case 8: return FN<uint64_t>(__VA_ARGS__);      \
	default: return -2;                            \
	}

extern "C" {

// Raw text (any case, N runs, IUPAC) -> reference codes + inclusive segments + effective size.
// base_out must hold len bytes; segs_out 2*max_segs ints. Returns 0, -1 on exception, -3 if too many segments.
int ref_encode(const char* seq, long len, char* base_out, int* segs_out, int max_segs, int* nseg, long* eff_size)
{
	try {
		ChromosomeOneDigitDna chrom;
		std::string header(">s");
		std::string s(seq, (size_t)len);
		chrom.setHeader(header);
		chrom.appendToSequence(s);
		chrom.finalize();
		const std::string* b = chrom.getBase();
		std::memcpy(base_out, b->data(), b->size());
		auto segs = chrom.getSegment();
		if ((int)segs->size() > max_segs) { return -3; }
		*nseg = (int)segs->size();
		for (size_t i = 0; i < segs->size(); i++) {
			segs_out[2 * i] = segs->at(i)->at(0);
			segs_out[2 * i + 1] = segs->at(i)->at(1);
		}
		*eff_size = chrom.getEffectiveSize();
		return 0;
	} catch (...) {
		return -1;
	}
}

// Loader<T>::get_point on raw text. hist_out: 4^k elements of elem_bytes.
int ref_get_point(const char* seq, long len, int k, int elem_bytes, void* hist_out, uint64_t* mers1,
		  uint64_t* mag, uint64_t* length, double* stddev)
{
	try {
		long dummy = 0;
		DISPATCH(elem_bytes, get_point_impl, seq, len, k, hist_out, mers1, mag, length, stddev, &dummy);
	} catch (...) {
		return -1;
	}
}

// KmerHashTable<unsigned long,V>(k, init).wholesaleIncrementNoOverflow on pre-encoded codes.
int ref_kmer_table(const char* codes, int first, int last, int k, int elem_bytes, uint64_t init, void* values_out, int* ret)
{
	try {
		DISPATCH(elem_bytes, kmer_table_impl, codes, first, last, k, init, values_out, ret);
	} catch (...) {
		return -1;
	}
}

// One raw single (static Feature<T>::xxx) on two histograms with explicit side-band.
// mag_override_{p,q}: 0 = let the constructor sum the bins; else a stale pseudo-magnitude (Q4).
int ref_raw_single(uint64_t flag, int elem_bytes, int k, uint64_t N, const void* P, const void* Q,
		   uint64_t mag_override_p, uint64_t mag_override_q, uint64_t len_p, uint64_t len_q, double* out)
{
	try {
		DISPATCH(elem_bytes, raw_single_impl, flag, k, N, P, Q, mag_override_p, mag_override_q, len_p, len_q, out);
	} catch (...) {
		return -1;
	}
}

// DivergencePoint<T>::distance (u64-truncated) on two histograms.
int ref_distance(int elem_bytes, uint64_t N, const void* P, const void* Q, uint64_t mag_override_p,
		 uint64_t mag_override_q, uint64_t* out)
{
	try {
		DISPATCH(elem_bytes, distance_impl, N, P, Q, mag_override_p, mag_override_q, out);
	} catch (...) {
		return -1;
	}
}

// DivergencePoint<T>::distance_d(Point<double>&) : histogram vs double-valued mean.
int ref_distance_d(int elem_bytes, uint64_t N, const void* P, const double* C, double* out)
{
	try {
		DISPATCH(elem_bytes, distance_d_impl, N, P, C, out);
	} catch (...) {
		return -1;
	}
}

// ---- pinned model (weights.txt written by Predictor::save) ----
void* ref_model_load(const char* weights_file, int elem_bytes, double cutoff)
{
	try {
		switch (elem_bytes) {
		case 1: return load_model<uint8_t>(weights_file, cutoff);
		case 2: return load_model<uint16_t>(weights_file, cutoff);
		case 4: return load_model<uint32_t>(weights_file, cutoff);
		case 8: return load_model<uint64_t>(weights_file, cutoff);
		}
	} catch (...) {
	}
	return nullptr;
}

void ref_set_bias(double b) { Predictor<uint8_t>::set_bias(b); }

// Score m pairs (ia[j], ib[j]) given as row indices into a histogram matrix H[n x N].
//   mode 0: Trainer-style classify: cache=compute(a,b); out_score=classify_sum(w0+sum w_c*combo_c),
//           out_dist = combo_0, out_close = round(score)>0, out_cache (m x S, may be NULL)
//   mode 1: Predictor::close(a,b)  -> out_close only
//   mode 2: Predictor::similarity(a,b) -> out_score only (regression model)
// mag[] : pseudo-magnitudes (NULL or 0 entries = from bins); len[] : lengths.  threads: omp threads.
int ref_score_pairs(void* model, int mode, uint64_t N, uint64_t n, const void* H, const uint64_t* mag,
		    const uint64_t* len, uint64_t m, const uint64_t* ia, const uint64_t* ib, double* out_score,
		    double* out_dist, uint8_t* out_close, double* out_cache, int threads, double* seconds)
{
	ModelBase* mb = (ModelBase*)model;
	try {
		DISPATCH(mb->elem_bytes, score_pairs_impl, mb, mode, N, n, H, mag, len, m, ia, ib, out_score, out_dist,
			 out_close, out_cache, threads, seconds);
	} catch (...) {
		return -1;
	}
}

// Trainer<T>::get_close: query row q vs candidate rows cand[0..m) of H (in that bvec order).
// Outputs: best index into cand (-1 if none), best dist, is_min, marks[m] (second==true).
int ref_get_close(void* model, uint64_t N, const void* H, const uint64_t* mag, const uint64_t* len, uint64_t q,
		  uint64_t m, const uint64_t* cand, int64_t* best, double* best_dist, int* is_min, uint8_t* marks,
		  int threads)
{
	ModelBase* mb = (ModelBase*)model;
	try {
		DISPATCH(mb->elem_bytes, get_close_impl, mb, N, H, mag, len, q, m, cand, best, best_dist, is_min, marks, threads);
	} catch (...) {
		return -1;
	}
}

// Trainer<T>::filter: center row c vs member rows; keep[j]=1 iff the member survives.
int ref_filter(void* model, uint64_t N, const void* H, const uint64_t* mag, const uint64_t* len, uint64_t c,
	       uint64_t m, const uint64_t* members, uint8_t* keep)
{
	ModelBase* mb = (ModelBase*)model;
	try {
		DISPATCH(mb->elem_bytes, filter_impl, mb, N, H, mag, len, c, m, members, keep);
	} catch (...) {
		return -1;
	}
}

// Trainer<T>::merge: center `cur` vs centers begin..last (indices into rows[]); returns chosen index (0 = none).
int ref_merge(void* model, uint64_t N, const void* H, const uint64_t* mag, const uint64_t* len, uint64_t ncen,
	      const uint64_t* rows, long cur, long begin, long last, long* out, int threads)
{
	ModelBase* mb = (ModelBase*)model;
	try {
		DISPATCH(mb->elem_bytes, merge_impl, mb, N, H, mag, len, ncen, rows, cur, begin, last, out, threads);
	} catch (...) {
		return -1;
	}
}

// Timing helper for the CPU baseline: Loader<T>::get_point over n raw sequences (concatenated text + offsets),
// omp over sequences. Optionally copies the histograms out.
int ref_count_batch(const char* text, const uint64_t* off, uint64_t n, int k, int elem_bytes, void* hist_out,
		    int threads, double* seconds)
{
	try {
		DISPATCH(elem_bytes, count_batch_impl, text, off, n, k, hist_out, threads, seconds);
	} catch (...) {
		return -1;
	}
}

// get_mean / mean_shift_update mean + closest member over rows members[0..n) of H
int ref_mean_closest(int elem_bytes, uint64_t N, const void *H, const uint64_t *members, uint64_t n, int64_t *best,
		     double *best_dist, double *mean_out, double *dist_out)
{
	try {
		DISPATCH(elem_bytes, mean_closest_impl, N, H, members, n, best, best_dist, mean_out, dist_out);
	} catch (...) {
		return -1;
	}
}

// Runner::run's histogram-width detection (src/cluster/CRunner.cpp:57-93): per sequence a u64 table (init 1) filled by
// the free fill_table<V> of src/cluster/ClusterFactory.h:40-54 (wholesaleIncrement, no length guard), max over everything.
int ref_largest_count(const char* text, const uint64_t* off, uint64_t n, int k, uint64_t* largest)
{
	try {
		uint64_t best = 0;
		for (uint64_t i = 0; i < n; i++) {
			ChromosomeOneDigitDna chrom;
			std::string header(">s");
			std::string s(text + off[i], (size_t)(off[i + 1] - off[i]));
			chrom.setHeader(header);
			chrom.appendToSequence(s);
			chrom.finalize();
			std::vector<uint64_t> values;
			KmerHashTable<unsigned long, uint64_t> table(k, 1);
			fill_table<uint64_t>(table, &chrom, values);
			uint64_t l_count = *std::max_element(values.begin(), values.end());
			if (l_count > best) { best = l_count; }
		}
		*largest = best;
		return 0;
	} catch (...) {
		return -1;
	}
}

int ref_max_threads(void) { return omp_get_max_threads(); }

} // extern "C"
