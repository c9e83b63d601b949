// FastcarWork_b200.cpp -- fastcar's work() (src/fastcar/FC_Runner.cpp:427-470) as one device batch per block.
//
// The relinked fastcar (oracle/Makefile: fastcar_b200) is the reference's fastcar with one call added to work()
// (integration/patch_fc_runner.py).  A block = the query chunk x one database chunk, both already loaded by the
// reference's Loader.  Here:
//   * both chunks' histograms go to the device once (the query chunk is kept while it is reused against later chunks),
//   * mc2_all_pairs runs the classifier over every (query, database point) pair inside the length window and returns the
//     close pairs -- Predictor<T>::close (src/predict/Predictor.cpp:255-281, :323-333),
//   * mc2_score_pairs with the regression model gives Predictor<T>::similarity for the survivors (:231-252, :284-300),
//   * the lines are written exactly as the reference writes them, query-major, database points in chunk order.
// work()'s start index comes from the reference's bin_search (FC_Runner.cpp:389-407), which returns 0 whenever the search
// runs off the right end of a sub-range: the loop then also visits points SHORTER than the window's lower bound.  The same
// function is restated below and the extra pairs are scored as a list, so the output is the reference's, quirk included.
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <limits>
#include <map>
#include <mutex>
#include <numeric>
#include <stdexcept>
#include <string>
#include <vector>

// The classifier is reachable through Predictor::get_class(); the regression model has no accessor.  Everything Predictor.h
// includes is included first, so the access override touches the Predictor class alone.
#include <omp.h>
#include <random>
#include <set>
#include <sstream>
#include "predict/GLM.h"
#include "clutil/Point.h"
#include "predict/Feature.h"
#include "clutil/Progress.h"
#include "clutil/Random.h"
#define private public
#include "predict/Predictor.h"
#undef private
#include "clutil/DivergencePoint.h"

#include "meshclust2_b200.h"
#include "device_b200.h"
#include "fastcar_b200.h"

namespace {

using mc2i::ok;

template <class T> struct DeviceWidth { static const int bytes = 0; };
template <> struct DeviceWidth<uint8_t> { static const int bytes = 1; };
template <> struct DeviceWidth<uint16_t> { static const int bytes = 2; };
template <> struct DeviceWidth<uint32_t> { static const int bytes = 4; };
template <> struct DeviceWidth<uint64_t> { static const int bytes = 8; };

template <class T>
mc2_model_desc describe(const Feature<T> &feat, const matrix::Matrix &weights, bool regression)
{
	mc2_model_desc d = mc2_model_desc();
	auto lookup = feat.get_lookup();
	auto mins = feat.get_mins();
	auto maxs = feat.get_maxs();
	auto combos = feat.get_combos();
	if (lookup.size() > MC2_MAX_SINGLES || combos.size() > MC2_MAX_COMBOS) {
		throw std::runtime_error("model too large for the device descriptor");
	}
	d.n_singles = (int32_t)lookup.size();
	for (size_t i = 0; i < lookup.size(); i++) {
		d.single_flag[i] = lookup[i];
		d.single_min[i] = mins[i];
		d.single_max[i] = maxs[i];
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
		for (size_t t = 0; t < combos[c].second.size(); t++) {
			d.combo_idx[c][t] = combos[c].second[t];
		}
	}
	for (int r = 0; r < weights.getNumRow(); r++) {
		d.weight[r] = weights.get(r, 0);
	}
	d.bias = regression ? 0.0 : Predictor<T>::classify_sum(0) - 0.5; // classify_sum(0) = logistic(0) + _bias
	d.regression = regression ? 1 : 0;
	return d;
}

struct Models {
	mc2_model *cls = nullptr, *reg = nullptr;
};
std::map<const void *, Models> g_models;

// the query chunk stays on the device while work() is called for it against one database chunk after the other
struct QueryCache {
	const void *first = nullptr;
	size_t n = 0;
	mc2_hset *set = nullptr;
};
std::map<const void *, QueryCache> g_queries; // keyed by the address of the caller's vector

template <class T>
mc2_hset *upload(mc2_ctx *ctx, const std::vector<Point<T> *> &v, int k)
{
	const size_t n = v.size(), N = (size_t)1 << (2 * k);
	std::vector<T> bins(n * N);
	std::vector<uint64_t> mag(n), len(n);
	for (size_t i = 0; i < n; i++) {
		const DivergencePoint<T> &q = dynamic_cast<const DivergencePoint<T> &>(*v[i]);
		std::copy(q.points.begin(), q.points.end(), bins.begin() + i * N);
		mag[i] = q.getPseudoMagnitude();
		len[i] = q.get_length();
	}
	mc2_hset *h = nullptr;
	ok(mc2_hset_from_host(ctx, bins.data(), n, k, (int)sizeof(T), mag.data(), len.data(), &h));
	return h;
}

// FC_Runner.cpp:389-407, restated (the relinked file's own copy is not visible from here)
template <class T>
long ref_bin_search(const std::vector<Point<T> *> &points, size_t begin, size_t last, size_t length)
{
	if (last < begin) {
		return 0;
	}
	size_t idx = begin + (last - begin) / 2;
	if (points.at(idx)->get_length() == length) {
		while (idx > 0 && points[idx - 1]->get_length() == length) {
			idx--;
		}
		return idx;
	} else if (points.at(idx)->get_length() > length) {
		if (begin == idx) {
			return idx;
		}
		return ref_bin_search(points, begin, idx - 1, length);
	} else {
		return ref_bin_search(points, idx + 1, last, length);
	}
}

struct Hit {
	uint64_t q, d;
	double sim;
};

} // namespace

template <class T>
bool mc2_batched_work(const std::vector<Point<T> *> &queries, const std::vector<Point<T> *> &pts, double similarity, Predictor<T> *pred,
		      const std::string &delim, std::ofstream &out, uintmax_t &num_pred_pos, bool format,
		      std::string (*format_header)(std::string))
{
	const int eb = DeviceWidth<T>::bytes;
	const uint8_t mode = pred->get_mode();
	if (eb == 0 || !(mode & PRED_MODE_CLASS) || std::getenv("MC2_NO_BATCH") || queries.empty()) {
		return false; // without a classifier every in-window pair is an output line: the reference's loop handles that
	}
	// the database chunk must be sorted by length, as work() assumes for its binary search
	for (size_t i = 1; i < pts.size(); i++) {
		if (pts[i - 1]->get_length() > pts[i]->get_length()) {
			return false;
		}
	}
	std::vector<Hit> hits;
	{
		std::lock_guard<std::mutex> lock(mc2i::device_mutex());
		mc2_ctx *ctx = mc2i::shared_ctx();
		const int k = pred->get_k();
		Models &m = g_models[pred];
		if (!m.cls) {
			mc2_model_desc dc = describe<T>(*pred->feat_c, pred->c_glm.get_weights(), false);
			ok(mc2_model_create(ctx, &dc, &m.cls));
			if (mode & PRED_MODE_REGR) {
				mc2_model_desc dr = describe<T>(*pred->feat_r, pred->r_glm.get_weights(), true);
				ok(mc2_model_create(ctx, &dr, &m.reg));
			}
		}
		QueryCache &qc = g_queries[&queries];
		if (qc.set == nullptr || qc.first != (const void *)queries[0] || qc.n != queries.size()) {
			if (qc.set) {
				mc2_hset_free(qc.set);
			}
			qc.set = upload<T>(ctx, queries, k);
			qc.first = (const void *)queries[0];
			qc.n = queries.size();
		}
		mc2_hset *dset = upload<T>(ctx, pts, k);
		// close pairs inside the window [(size_t)(len_q * id), (size_t)(len_q / id)] (FC_Runner.cpp:435-444)
		uint64_t cap = 1 << 20, n_out = 0, n_scored = 0;
		std::vector<uint64_t> oq, od;
		std::vector<double> os;
		for (;;) {
			oq.resize(cap);
			od.resize(cap);
			os.resize(cap);
			ok(mc2_all_pairs(ctx, m.cls, qc.set, 0, queries.size(), dset, 0, pts.size(), 0, similarity, cap, oq.data(), od.data(),
					 os.data(), &n_out, &n_scored));
			if (n_out <= cap) {
				break;
			}
			cap = n_out;
		}
		for (uint64_t i = 0; i < n_out; i++) {
			hits.push_back(Hit{oq[i], od[i], 1.0});
		}
		// the reference's start index may lie below the window (bin_search returns 0 off the right end of a sub-range)
		std::vector<uint64_t> xa, xb;
		for (size_t qi = 0; qi < queries.size(); qi++) {
			const size_t q_len = queries[qi]->get_length();
			const size_t begin_length = q_len * similarity, end_length = q_len / similarity;
			const size_t start = (size_t)ref_bin_search(pts, 0, pts.size() - 1, begin_length);
			for (size_t i = start; i < pts.size() && pts[i]->get_length() <= end_length && pts[i]->get_length() < begin_length; i++) {
				xa.push_back(i);
				xb.push_back(qi);
			}
		}
		if (!xa.empty()) {
			mc2_pairs pr = mc2_pairs();
			pr.set_a = dset; // close(pts[i], query): the database point first
			pr.set_b = qc.set;
			pr.n_pairs = xa.size();
			pr.ia = xa.data();
			pr.ib = xb.data();
			std::vector<uint8_t> cl(xa.size());
			ok(mc2_score_pairs(ctx, m.cls, &pr, nullptr, nullptr, cl.data(), nullptr, nullptr, nullptr));
			for (size_t j = 0; j < xa.size(); j++) {
				if (cl[j]) {
					hits.push_back(Hit{xb[j], xa[j], 1.0});
				}
			}
		}
		if ((mode & PRED_MODE_REGR) && !hits.empty()) {
			std::vector<uint64_t> ia(hits.size()), ib(hits.size());
			for (size_t j = 0; j < hits.size(); j++) {
				ia[j] = hits[j].d;
				ib[j] = hits[j].q;
			}
			mc2_pairs pr = mc2_pairs();
			pr.set_a = dset;
			pr.set_b = qc.set;
			pr.n_pairs = hits.size();
			pr.ia = ia.data();
			pr.ib = ib.data();
			std::vector<double> sim(hits.size());
			ok(mc2_score_pairs(ctx, m.reg, &pr, sim.data(), nullptr, nullptr, nullptr, nullptr, nullptr));
			for (size_t j = 0; j < hits.size(); j++) {
				hits[j].sim = sim[j];
			}
		}
		mc2_hset_free(dset);
	}
	std::sort(hits.begin(), hits.end(), [](const Hit &a, const Hit &b) { return a.q != b.q ? a.q < b.q : a.d < b.d; });
	for (const Hit &h : hits) {
		num_pred_pos++;
		if (h.sim > 0) {
			Point<T> *query = queries[h.q], *p = pts[h.d];
			if (format) {
				out << format_header(query->get_header()) << delim << format_header(p->get_header()) << delim << 100 * h.sim << endl;
			} else {
				out << query->get_header() << delim << p->get_header() << delim << 100 * h.sim << endl;
			}
		}
	}
	return true;
}

#define MC2_INSTANTIATE_WORK(T)                                                                                               \
	template bool mc2_batched_work<T>(const std::vector<Point<T> *> &, const std::vector<Point<T> *> &, double, Predictor<T> *, \
					  const std::string &, std::ofstream &, uintmax_t &, bool, std::string (*)(std::string));
MC2_INSTANTIATE_WORK(uint8_t)
MC2_INSTANTIATE_WORK(uint16_t)
MC2_INSTANTIATE_WORK(uint32_t)
MC2_INSTANTIATE_WORK(uint64_t)
MC2_INSTANTIATE_WORK(int)
MC2_INSTANTIATE_WORK(double)
