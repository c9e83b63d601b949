// mc2_shim.hpp — C++ host classes that keep the reference's class names and call signatures for the hot path, over
// the C ABI of include/meshclust2_b200.h.  Header-only; link with -lmeshclust2_b200.
//
// Mirrors (paths relative to the MeShClust2 reference root):
//   nonltr::KmerHashTable<I,V>      src/nonltr/KmerHashTable.h:20-79      (k-mer table: ctor(k, init), wholesaleIncrementNoOverflow,
//                                                                          getValues, getMaxTableSize, getK)
//   DivergencePoint<T> / Point<T>   src/clutil/DivergencePoint.h:13-88     (points, mag, length, id, header, 1-mers, clone(), set())
//   Loader<T>::get_point            src/clutil/Loader.h:70-74              (+ a batched get_points that is the intended fast entry)
//   Combo, Feature<T>               src/predict/Feature.h:66-71, 107-384   (add_feature, set_normal/get_normal, normalize, finalize,
//                                                                          compute, operator(), size, get_combos/lookup/mins/maxs)
//   Predictor<T>(file)              src/predict/Predictor.h:27-90          (close, similarity, get_k/get_id/get_mode/get_datatype,
//                                                                          classify_sum, set_bias)
//   Trainer<T> glue                 src/cluster/Trainer.h:30-36            (get_close / filter / merge on index lists)
// Error behaviour follows the reference: where it throws, these throw (std::runtime_error carrying mc2_last_error()).
// Single-pair calls (Feature::compute(p,q), Predictor::close(a,b)) work but move two rows per call; the batched members
// (PointSet + *_batch) are the ones the clustering driver should use (batch the candidates).
#pragma once
#include <cmath>
#include <cstdint>
#include <limits>
#include <memory>
#include <stdexcept>
#include <string>
#include <tuple>
#include <utility>
#include <vector>

#include "../../include/meshclust2_b200.h"

namespace mc2shim {

inline void check(int rc)
{
	if (rc != MC2_OK) {
		throw std::runtime_error(std::string("meshclust2_b200: ") + mc2_last_error());
	}
}

// one GPU context shared by the shim objects of a thread (the reference is single-process; fastcar's OpenMP workers would
// each own one)
class Context {
public:
	explicit Context(int device = 0) { check(mc2_ctx_create(device, &ctx_)); }
	~Context() { mc2_ctx_destroy(ctx_); }
	Context(const Context &) = delete;
	Context &operator=(const Context &) = delete;
	mc2_ctx *get() const { return ctx_; }
	static Context &instance()
	{
		static thread_local Context c(0);
		return c;
	}

private:
	mc2_ctx *ctx_ = nullptr;
};

// ---------------------------------------------------------------------------------------------------------------
// nonltr::KmerHashTable<I,V>
// ---------------------------------------------------------------------------------------------------------------
template <class I, class V>
class KmerHashTable {
public:
	KmerHashTable(int keyLength, V initValue = 0) : k_(keyLength), init_(initValue)
	{
		size_ = (I)1 << (2 * k_);
		values_.assign((size_t)size_, initValue);
	}
	// counts every k-mer starting in [firstKmerStart, lastKmerStart] on the device and folds the counts into the table
	// with the reference's saturation rule; returns -1 if an increment met a bin already at max(V), else 0
	int wholesaleIncrementNoOverflow(const char *sequence, int firstKmerStart, int lastKmerStart)
	{
		std::vector<uint64_t> counts((size_t)size_);
		int32_t dummy = 0;
		check(mc2_kmer_table_increment(Context::instance().get(), sequence, firstKmerStart, lastKmerStart, k_, 8, 0,
					       counts.data(), &dummy));
		const uint64_t vmax = (uint64_t)std::numeric_limits<V>::max();
		int ret = 0;
		for (size_t h = 0; h < counts.size(); h++) {
			if (counts[h] == 0) {
				continue;
			}
			uint64_t v = (uint64_t)values_[h] + counts[h];
			if (v > vmax) {
				v = vmax;
				ret = -1;
			}
			values_[h] = (V)v;
		}
		return ret;
	}
	const V *getValues() const { return values_.data(); }
	I getMaxTableSize() const { return size_; }
	int getK() const { return k_; }

private:
	int k_;
	V init_;
	I size_;
	std::vector<V> values_;
};

// ---------------------------------------------------------------------------------------------------------------
// DivergencePoint<T>
// ---------------------------------------------------------------------------------------------------------------
template <class T>
class DivergencePoint {
public:
	DivergencePoint(const std::vector<T> &pts, uint64_t len) : points(pts), nucl_length(len)
	{
		mag = 0;
		for (T v : points) {
			mag += v; // src/clutil/DivergencePoint.cpp:99-110
		}
	}
	DivergencePoint *clone() const
	{ // DivergencePoint.h:35-43: recomputes mag; copies header/id/length/stddev; NOT one_mers nor k
		auto *d = new DivergencePoint(points, nucl_length);
		d->header = header;
		d->id = id;
		d->s_dev = s_dev;
		return d;
	}
	void set(const DivergencePoint &p)
	{ // DivergencePoint.cpp:182-190: points, length, header, id — NOT mag (stays stale), NOT stddev
		points = p.points;
		nucl_length = p.nucl_length;
		header = p.header;
		id = p.id;
	}
	uint64_t getPseudoMagnitude() const { return mag; }
	uint64_t getRealMagnitude() const { return mag - points.size(); }
	unsigned long get_length() const { return nucl_length; }
	void set_length(unsigned long l) { nucl_length = l; }
	uintmax_t get_id() const { return id; }
	void set_id(uintmax_t i) { id = i; }
	const std::string &get_header() const { return header; }
	void set_header(const std::string &h) { header = h; }
	double get_stddev() const { return s_dev; }
	void set_stddev(double s) { s_dev = s; }
	std::vector<uint64_t> get_1mers() const { return one_mers; }
	void set_1mers(const std::vector<uint64_t> &v) { one_mers = v; }
	int getK() const { return k; }
	void setK(int kk) { k = kk; }
	unsigned long size() const { return points.size(); }
	const std::vector<T> &get_data() const { return points; }
	std::vector<T> points;

private:
	uintmax_t mag = 0;
	uint64_t id = 0;
	uint64_t nucl_length = 0;
	double s_dev = 0;
	std::string header;
	std::vector<uint64_t> one_mers;
	int k = 0;
};
template <class T>
using Point = DivergencePoint<T>;

// a device-resident copy of a set of points (rows in the order given); side-band = what the host objects report
template <class T>
class PointSet {
public:
	PointSet(const std::vector<DivergencePoint<T> *> &pts, int k) : n_(pts.size()), k_(k)
	{
		const size_t N = (size_t)1 << (2 * k);
		std::vector<T> bins(n_ * N);
		std::vector<uint64_t> mag(n_), len(n_);
		for (size_t i = 0; i < n_; i++) {
			if (pts[i]->points.size() != N) {
				throw std::runtime_error("PointSet: histogram size does not match k");
			}
			std::copy(pts[i]->points.begin(), pts[i]->points.end(), bins.begin() + i * N);
			mag[i] = pts[i]->getPseudoMagnitude();
			len[i] = pts[i]->get_length();
		}
		check(mc2_hset_from_host(Context::instance().get(), bins.data(), n_, k, (int)sizeof(T), mag.data(), len.data(), &h_));
	}
	explicit PointSet(mc2_hset *adopt) : h_(adopt), n_(mc2_hset_count(adopt)), k_(mc2_hset_k(adopt)) {}
	~PointSet() { mc2_hset_free(h_); }
	PointSet(const PointSet &) = delete;
	PointSet &operator=(const PointSet &) = delete;
	mc2_hset *get() const { return h_; }
	size_t size() const { return n_; }
	int k() const { return k_; }

private:
	mc2_hset *h_ = nullptr;
	size_t n_;
	int k_;
};

// ---------------------------------------------------------------------------------------------------------------
// Loader<T>
// ---------------------------------------------------------------------------------------------------------------
template <class T>
class Loader {
public:
	// Loader<T>::get_point(header, base, id, k): raw text; as in the reference everything but A,C,G,T is stripped first
	// (src/clutil/Loader.cpp:112-134)
	static DivergencePoint<T> *get_point(std::string header, const std::string &base, uintmax_t &id, int k)
	{
		std::string clean;
		clean.reserve(base.size());
		for (char c : base) {
			if (c == 'A' || c == 'C' || c == 'G' || c == 'T') {
				clean.push_back(c);
			}
		}
		std::vector<std::string> hs{header}, ss{clean};
		auto v = get_points(hs, ss, id, k);
		return v[0];
	}
	// batched form of Loader<T>::get_point(ChromosomeOneDigit*, id, k) (src/clutil/Loader.cpp:138-179): raw sequence
	// text per record (N runs, IUPAC, lower case handled as Chromosome::help does); ids are assigned in order
	static std::vector<DivergencePoint<T> *> get_points(const std::vector<std::string> &headers,
							    const std::vector<std::string> &seqs, uintmax_t &id, int k,
							    std::unique_ptr<PointSet<T>> *device_set = nullptr, int threads = 1)
	{
		const size_t n = seqs.size();
		std::vector<uint64_t> off(n + 1, 0);
		std::string blob;
		for (size_t i = 0; i < n; i++) {
			off[i + 1] = off[i] + seqs[i].size();
		}
		blob.reserve(off[n]);
		for (auto &s : seqs) {
			blob += s;
		}
		// segmentation, letter coding and packing happen on the device (mc2_seqs_from_text == Chromosome::help + encode)
		(void)threads;
		mc2_ctx *ctx = Context::instance().get();
		mc2_seqs *sq = nullptr;
		check(mc2_seqs_from_text(ctx, blob.data(), off.data(), n, &sq));
		mc2_hset *h = nullptr;
		int rc = mc2_count_kmers(ctx, sq, k, (int)sizeof(T), &h);
		mc2_seqs_free(sq);
		check(rc);
		const size_t N = (size_t)1 << (2 * k);
		std::vector<T> bins(n * N);
		std::vector<uint64_t> mag(n), len(n), mers(4 * n);
		std::vector<double> sd(n);
		std::vector<int32_t> novf(n);
		rc = mc2_hset_download(ctx, h, 0, n, bins.data(), mag.data(), len.data(), mers.data(), sd.data(), novf.data(), nullptr);
		if (rc != MC2_OK) {
			mc2_hset_free(h);
			check(rc);
		}
		std::vector<DivergencePoint<T> *> out(n);
		for (size_t i = 0; i < n; i++) {
			std::vector<T> v(bins.begin() + i * N, bins.begin() + (i + 1) * N);
			auto *p = new DivergencePoint<T>(v, seqs[i].size()); // ctor sums the bins -> mag
			p->set_1mers(std::vector<uint64_t>(mers.begin() + 4 * i, mers.begin() + 4 * i + 4));
			p->set_header(headers[i]);
			p->set_length(len[i]); // chrom->getEffectiveSize()
			p->setK(k);
			p->set_stddev(sd[i]);
			p->set_id(id++);
			num_overflow() += novf[i];
			out[i] = p;
		}
		if (device_set) {
			device_set->reset(new PointSet<T>(h));
		} else {
			mc2_hset_free(h);
		}
		return out;
	}
	static uint64_t &num_overflow()
	{
		static uint64_t n = 0;
		return n;
	}
	// Runner::run's histogram-width detection (src/cluster/CRunner.cpp:57-127): "Largest count" over raw sequences and the
	// number of bytes per bin the reference would pick for it (1, 2, 4 or 8 = uint8_t ... uint64_t)
	static uint64_t largest_count(const std::vector<std::string> &seqs, int k, int *elem_bytes = nullptr)
	{
		std::vector<uint64_t> off(seqs.size() + 1, 0);
		std::string blob;
		for (size_t i = 0; i < seqs.size(); i++) {
			off[i + 1] = off[i] + seqs[i].size();
			blob += seqs[i];
		}
		mc2_ctx *ctx = Context::instance().get();
		mc2_seqs *sq = nullptr;
		check(mc2_seqs_from_text(ctx, blob.data(), off.data(), seqs.size(), &sq));
		mc2_hset *h = nullptr;
		uint64_t largest = 0;
		int eb = 1;
		int rc = mc2_count_kmers_auto(ctx, sq, k, &largest, &eb, &h);
		mc2_seqs_free(sq);
		check(rc);
		mc2_hset_free(h);
		if (elem_bytes) {
			*elem_bytes = eb;
		}
		return largest;
	}
};

// ---------------------------------------------------------------------------------------------------------------
// Feature<T>
// ---------------------------------------------------------------------------------------------------------------
enum class Combo { xy, x2y2, xy2, x2y }; // src/predict/Feature.h:66-71

template <class T>
struct pra {
	DivergencePoint<T> *first;
	DivergencePoint<T> *second;
	double val;
};

template <class T>
class Feature {
public:
	explicit Feature(int k_) : k(k_) {}
	void add_feature(uint64_t f_flags, Combo combo = Combo::xy)
	{ // src/predict/Feature.cpp:102-127
		std::vector<int> indices;
		for (uint64_t f = 1; f <= f_flags && f != 0; f <<= 1) {
			if ((f_flags & f) != 0) {
				if ((flags & f) == 0) {
					lookup.push_back(f);
					mins.push_back(std::numeric_limits<double>::max());
					maxs.push_back(std::numeric_limits<double>::min());
					is_finalized.push_back(false);
					flags |= f;
				}
				indices.push_back(index_of(f));
			}
		}
		combos.push_back(std::make_pair(combo, indices));
		model_.reset();
	}
	void set_normal(uint64_t single_flag, double mn, double mx)
	{
		int idx = index_of(single_flag);
		mins.at(idx) = mn;
		maxs.at(idx) = mx;
		is_finalized.at(idx) = true;
		model_.reset();
	}
	std::pair<double, double> get_normal(uint64_t single_flag) const
	{
		int idx = index_of(single_flag);
		return std::make_pair(mins.at(idx), maxs.at(idx));
	}
	void finalize()
	{
		for (size_t i = 0; i < is_finalized.size(); i++) {
			is_finalized[i] = true;
		}
	}
	// min/max of every not-yet-finalised raw single over the given pairs (src/predict/Feature.cpp:216-268); the raw
	// singles of the whole batch come from one device launch
	void normalize(const std::vector<pra<T>> &pairs)
	{
		std::vector<double> raw = raw_batch(pairs);
		const size_t S = lookup.size();
		for (size_t i = 0; i < S; i++) {
			if (is_finalized[i]) {
				continue;
			}
			double small = mins[i], big = maxs[i];
			for (size_t j = 0; j < pairs.size(); j++) {
				double v = raw[j * S + i];
				small = v < small ? v : small;
				big = v > big ? v : big;
			}
			mins[i] = small;
			maxs[i] = big;
			if (std::fabs(maxs[i] - mins[i]) <= 0.000000001 || std::isinf(maxs[i]) || std::isinf(mins[i])) {
				throw std::runtime_error("Feature::normalize: degenerate range");
			}
		}
		model_.reset();
	}
	// normalised singles of one pair (src/predict/Feature.h:197-201)
	std::vector<double> compute(DivergencePoint<T> &p, DivergencePoint<T> &q)
	{
		std::vector<DivergencePoint<T> *> two{&p, &q};
		PointSet<T> set(two, k);
		std::vector<double> cache(lookup.size());
		mc2_pairs pr = pair_desc(set.get(), set.get(), 1);
		uint64_t ia = 0, ib = 1;
		pr.ia = &ia;
		pr.ib = &ib;
		check(mc2_score_pairs(Context::instance().get(), model(), &pr, nullptr, nullptr, nullptr, cache.data(), nullptr, nullptr));
		return cache;
	}
	// combination value (src/predict/Feature.h:205-239)
	double operator()(int col, const std::vector<double> &cache) const
	{
		auto pr = combos.at(col);
		auto &indices = pr.second;
		if (pr.first == Combo::xy) {
			double prod = 1;
			for (auto idx : indices) {
				prod *= cache[idx];
			}
			return prod;
		} else if (pr.first == Combo::x2y2) {
			double prod = 1;
			for (auto idx : indices) {
				prod *= cache[idx] * cache[idx];
			}
			return prod;
		} else if (pr.first == Combo::xy2) {
			if (indices.size() != 2) {
				throw "invalid";
			}
			return cache[indices[0]] * cache[indices[1]] * cache[indices[1]];
		}
		if (indices.size() != 2) {
			throw "invalid";
		}
		return cache[indices[0]] * cache[indices[0]] * cache[indices[1]];
	}
	size_t size() const { return combos.size(); }
	std::vector<std::pair<Combo, std::vector<int>>> get_combos() const { return combos; }
	std::vector<double> get_mins() const { return mins; }
	std::vector<double> get_maxs() const { return maxs; }
	std::vector<uint64_t> get_lookup() const { return lookup; }
	int get_k() const { return k; }

	// ---- batched members (what Trainer / selectors should call) ----
	std::vector<double> raw_batch(const std::vector<pra<T>> &pairs)
	{
		std::vector<DivergencePoint<T> *> pts;
		pts.reserve(2 * pairs.size());
		for (auto &p : pairs) {
			pts.push_back(p.first);
			pts.push_back(p.second);
		}
		PointSet<T> set(pts, k);
		std::vector<uint64_t> ia(pairs.size()), ib(pairs.size());
		for (size_t j = 0; j < pairs.size(); j++) {
			ia[j] = 2 * j;
			ib[j] = 2 * j + 1;
		}
		std::vector<double> raw(pairs.size() * lookup.size());
		mc2_pairs pr = pair_desc(set.get(), set.get(), pairs.size());
		pr.ia = ia.data();
		pr.ib = ib.data();
		// normalisation constants are irrelevant for the raw output; use a unit range so nothing is NaN
		mc2_model_desc d = desc(true);
		mc2_model *m = nullptr;
		check(mc2_model_create(Context::instance().get(), &d, &m));
		int rc = mc2_score_pairs(Context::instance().get(), m, &pr, nullptr, nullptr, nullptr, nullptr, raw.data(), nullptr);
		mc2_model_free(m);
		check(rc);
		return raw;
	}
	// the device model for this feature set with the given GLM weights (weights.size() == size()+1)
	mc2_model_desc desc(bool unit_range = false, const std::vector<double> *weights = nullptr, double bias = 0,
			    bool regression = false) const
	{
		mc2_model_desc d = mc2_model_desc();
		if (lookup.size() > MC2_MAX_SINGLES || combos.size() > MC2_MAX_COMBOS || combos.empty()) {
			throw std::runtime_error("Feature: too many (or no) singles / combos for the device model");
		}
		d.n_singles = (int32_t)lookup.size();
		for (size_t i = 0; i < lookup.size(); i++) {
			d.single_flag[i] = lookup[i];
			d.single_min[i] = unit_range ? 0.0 : mins[i];
			d.single_max[i] = unit_range ? 1.0 : maxs[i];
		}
		d.n_combos = (int32_t)combos.size();
		for (size_t c = 0; c < combos.size(); c++) {
			switch (combos[c].first) {
			case Combo::xy: d.combo_kind[c] = MC2_COMBO_XY; break;
			case Combo::xy2: d.combo_kind[c] = MC2_COMBO_XY2; break;
			case Combo::x2y: d.combo_kind[c] = MC2_COMBO_X2Y; break;
			case Combo::x2y2: d.combo_kind[c] = MC2_COMBO_X2Y2; break;
			}
			d.combo_nidx[c] = (int32_t)combos[c].second.size();
			for (size_t t = 0; t < combos[c].second.size() && t < MC2_MAX_COMBO_IDX; t++) {
				d.combo_idx[c][t] = combos[c].second[t];
			}
		}
		for (size_t c = 0; c <= combos.size(); c++) {
			d.weight[c] = weights ? (*weights)[c] : 0.0;
		}
		d.bias = bias;
		d.regression = regression ? 1 : 0;
		return d;
	}

private:
	int index_of(uint64_t single_flag) const
	{
		for (size_t i = 0; i < lookup.size(); i++) {
			if (lookup[i] == single_flag) {
				return (int)i;
			}
		}
		return -1;
	}
	static mc2_pairs pair_desc(const mc2_hset *a, const mc2_hset *b, uint64_t n)
	{
		mc2_pairs p = mc2_pairs();
		p.set_a = a;
		p.set_b = b;
		p.n_pairs = n;
		return p;
	}
	struct ModelDel {
		void operator()(mc2_model *m) const { mc2_model_free(m); }
	};
	mc2_model *model()
	{
		if (!model_) {
			mc2_model_desc d = desc();
			mc2_model *m = nullptr;
			check(mc2_model_create(Context::instance().get(), &d, &m));
			model_.reset(m, ModelDel());
		}
		return model_.get();
	}
	int k;
	uint64_t flags = 0;
	std::vector<std::pair<Combo, std::vector<int>>> combos;
	std::vector<double> mins, maxs;
	std::vector<bool> is_finalized;
	std::vector<uint64_t> lookup;
	std::shared_ptr<mc2_model> model_;
};

// ---------------------------------------------------------------------------------------------------------------
// Predictor<T> (file constructor: a trained model) and the Trainer<T> glue over index lists
// ---------------------------------------------------------------------------------------------------------------
#define MC2_PRED_MODE_CLASS 1
#define MC2_PRED_MODE_REGR 2

template <class T>
class Predictor {
public:
	explicit Predictor(const std::string &filename)
	{ // src/predict/Predictor.cpp:47-79
		mc2_model_desc d;
		int eb = 0, md = 0;
		check(mc2_model_desc_from_file(filename.c_str(), 0, &d, &k, &id, &eb, &md));
		mode = (uint8_t)md;
		datatype = eb == 1 ? "uint8_t" : eb == 2 ? "uint16_t" : eb == 4 ? "uint32_t" : "uint64_t";
		d.bias = bias();
		check(mc2_model_create(Context::instance().get(), &d, &c_model));
		c_desc = d;
		if (mode & MC2_PRED_MODE_REGR) {
			mc2_model_desc r;
			check(mc2_model_desc_from_file(filename.c_str(), 1, &r, nullptr, nullptr, nullptr, nullptr));
			check(mc2_model_create(Context::instance().get(), &r, &r_model));
		}
	}
	~Predictor()
	{
		mc2_model_free(c_model);
		mc2_model_free(r_model);
	}
	Predictor(const Predictor &) = delete;
	Predictor &operator=(const Predictor &) = delete;
	static double classify_sum(double sum) { return 1.0 / (1 + std::exp(-sum)) + bias(); } // Predictor.cpp:316-320
	static void set_bias(double b) { bias() = b; }
	bool close(DivergencePoint<T> *a, DivergencePoint<T> *b)
	{
		std::vector<uint8_t> c = close_batch({a}, {b});
		return c[0] != 0;
	}
	double similarity(DivergencePoint<T> *a, DivergencePoint<T> *b)
	{
		if (!r_model) {
			throw "Bad"; // Predictor.cpp:233-236
		}
		std::vector<DivergencePoint<T> *> two{a, b};
		PointSet<T> set(two, k);
		mc2_pairs pr = mc2_pairs();
		pr.set_a = pr.set_b = set.get();
		pr.n_pairs = 1;
		uint64_t ia = 0, ib = 1;
		pr.ia = &ia;
		pr.ib = &ib;
		double s = 0;
		check(mc2_score_pairs(Context::instance().get(), r_model, &pr, &s, nullptr, nullptr, nullptr, nullptr, nullptr));
		return s;
	}
	// close(a[j], b[j]) for a whole list in one launch
	std::vector<uint8_t> close_batch(const std::vector<DivergencePoint<T> *> &a, const std::vector<DivergencePoint<T> *> &b)
	{
		std::vector<DivergencePoint<T> *> pts(a);
		pts.insert(pts.end(), b.begin(), b.end());
		PointSet<T> set(pts, k);
		const size_t m = a.size();
		std::vector<uint64_t> ia(m), ib(m);
		for (size_t j = 0; j < m; j++) {
			ia[j] = j;
			ib[j] = m + j;
		}
		mc2_pairs pr = mc2_pairs();
		pr.set_a = pr.set_b = set.get();
		pr.n_pairs = m;
		pr.ia = ia.data();
		pr.ib = ib.data();
		std::vector<uint8_t> c(m);
		check(mc2_score_pairs(Context::instance().get(), c_model, &pr, nullptr, nullptr, c.data(), nullptr, nullptr, nullptr));
		return c;
	}
	// Trainer<T>::get_close (src/cluster/Trainer.cpp:23-71) on a device-resident set: candidates by row index
	std::tuple<int64_t, double, bool> get_close(const PointSet<T> &set, uint64_t query, const std::vector<uint64_t> &cand,
						     std::vector<uint8_t> &marks)
	{
		int64_t best = -1;
		double bd = -1;
		int32_t is_min = 1;
		marks.assign(cand.size(), 0);
		check(mc2_get_close(Context::instance().get(), c_model, set.get(), query, set.get(), cand.data(), 0, cand.size(), id, &best,
				    &bd, &is_min, marks.data()));
		return std::make_tuple(best, bd, is_min != 0);
	}
	// Trainer<T>::filter (Trainer.cpp:123-141)
	std::vector<uint8_t> filter(const PointSet<T> &centers, uint64_t center, const PointSet<T> &members,
				    const std::vector<uint64_t> &rows)
	{
		std::vector<uint8_t> keep(rows.size());
		check(mc2_filter(Context::instance().get(), c_model, centers.get(), center, members.get(), rows.data(), rows.size(), id,
				 keep.data()));
		return keep;
	}
	// Trainer<T>::merge (Trainer.cpp:74-109)
	long merge(const PointSet<T> &centers, const std::vector<uint64_t> &rows, long cur, long begin, long last)
	{
		int64_t out = 0;
		check(mc2_merge(Context::instance().get(), c_model, centers.get(), rows.data(), cur, begin, last, id, &out));
		return (long)out;
	}
	uint8_t get_mode() const { return mode; }
	int get_k() const { return k; }
	double get_id() const { return id; }
	std::string get_datatype() const { return datatype; }
	const mc2_model_desc &get_class_desc() const { return c_desc; }

private:
	static double &bias()
	{
		static double b = 0;
		return b;
	}
	int k = 0;
	double id = 0;
	uint8_t mode = 0;
	std::string datatype;
	mc2_model *c_model = nullptr, *r_model = nullptr;
	mc2_model_desc c_desc;
};

} // namespace mc2shim
