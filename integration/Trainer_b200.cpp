// Trainer_b200.cpp — the one reference file that is REPLACED for the drop-in (ClusterFactory.cpp and CRunner.cpp only gain the
// few-line offers of integration/patch_*.py): a replacement for
// src/cluster/Trainer.cpp that keeps Trainer<T>'s interface (src/cluster/Trainer.h:20-45, compiled from the reference
// header, unmodified) and sends the candidate batches of get_close / filter / merge to libmeshclust2_b200 through the C
// ABI (include/meshclust2_b200.h).  Everything else of MeShClust2 — CLI, FASTA I/O, mutation generator, GLM fit and
// feature selection, the mean-shift driver (ClusterFactory), bvec, CLSTR output — is the reference's own object code.
//
// It is built by oracle/Makefile (`make integrated`) against the reference sources where they lie, into
// oracle/_ref/meshclust2_b200, and exists to prove the boundary end to end: tests/test_integrated_cluster.py checks that
// this binary and the unmodified reference binary produce the same clusters from the same weights.
//
// Batching follows SURVEY.md section 3.1: one device call per get_close / filter / merge (the candidates of one query);
// the histograms of all points are uploaded once, centers are addressed by the id of the point whose bins they carry plus
// the (possibly stale, quirk Q4) pseudo-magnitude and length the host object reports.
//
// The update stage goes further (north_star: the mean-shift driver is "unchanged apart from batching candidate pairs to the
// device"): mc2_batched_update / mc2_batched_merge below serve every center of one pass of ClusterFactory<T>::MS's two loops
// with one device call each (mc2_update_centers, mc2_merge_centers).  They are reached through the two-hunk edit of
// src/cluster/ClusterFactory.cpp that integration/patch_cluster_factory.py applies at build time; with
// MC2_NO_BATCH=1 in the environment they decline and the reference's per-center loops run instead.
#include "cluster/Trainer.h"
#include "clutil/Datatype.h"
#include "clutil/DivergencePoint.h"
#include "predict/Predictor.h"

#include <atomic>
#include <cfloat>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <limits>
#include <map>
#include <mutex>
#include <stdexcept>
#include <thread>
#include <unordered_map>

#include "meshclust2_b200.h"
#include "device_b200.h"
#include "update_batch_b200.h"

namespace {

std::mutex &g_mu = mc2i::device_mutex(); // the update stage calls filter() from an OpenMP loop; one GPU context, one caller at a time
using mc2i::ok;

struct Device {
	mc2_ctx *ctx = nullptr;
	mc2_model *model = nullptr;
	mc2_hset *points = nullptr;   // row = Point::get_id(); rows are filled as their points are first seen (see rows_of)
	uint64_t cap = 0;             // rows allocated in `points`
	std::vector<uint8_t> present; // row id holds the histogram of the point that currently carries that id
	std::vector<const void *> owner; // the point object a row was filled from (NULL: filled from a center clone)
	std::vector<uint64_t> hmag, hlen; // host copies of the rows' side-band (needed when the set is re-allocated)
	mc2_hset *scratch = nullptr;  // assembled center rows (filter: 1, merge: <= 1 + delta)
	uint64_t scratch_rows = 0;
	mc2_hset *centers = nullptr;  // every center of an update / merge pass (batched stage)
	uint64_t center_rows = 0;
	uint64_t n = 0;
	int k = 0;
};

std::map<const void *, Device> g_dev;

// MC2_TIMING=1: wall-clock spent behind the boundary, printed to stderr at exit (development aid)
struct Timing {
	double init = 0, get_close = 0, filter = 0, closest = 0, merge = 0, update_batch = 0, merge_batch = 0;
	unsigned long n_get_close = 0, n_filter = 0, n_closest = 0, n_merge = 0, n_update_batch = 0, n_merge_batch = 0;
	bool on = std::getenv("MC2_TIMING") != nullptr;
	~Timing()
	{
		if (on) {
			std::fprintf(stderr,
				     "meshclust2_b200 timing: init %.3f s | get_close %lu calls %.3f s | filter %lu %.3f s | closest %lu %.3f s | "
				     "merge %lu %.3f s | update batches %lu %.3f s | merge batches %lu %.3f s\n",
				     init, n_get_close, get_close, n_filter, filter, n_closest, closest, n_merge, merge, n_update_batch,
				     update_batch, n_merge_batch, merge_batch);
		}
	}
} g_time;

struct Stopwatch {
	double &acc;
	unsigned long &n;
	std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
	Stopwatch(double &a, unsigned long &c) : acc(a), n(c) {}
	~Stopwatch()
	{
		acc += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
		n++;
	}
};

// Start-up off the critical path: the CUDA context is created by a thread started before main() (integration/device_b200.cpp);
// the upload of the point histograms does not depend on the model, so Trainer::train starts it on a second thread while the
// host fits the GLM.  device_for() joins.  Rows are uploaded in the order of the Trainer's point vector, which is the order
// CRunner.cpp:587-592 assigns the final ids in; device_for() checks that before use.  MC2_NO_PREWARM=1 turns both off.
struct Prewarm {
	std::thread th;
	mc2_ctx *ctx = nullptr;
	mc2_hset *points = nullptr;
	std::string err;
	bool started = false;
	std::atomic<bool> cancel{false};
	~Prewarm()
	{
		if (th.joinable()) {
			th.join();
		}
	}
};
std::map<const void *, Prewarm> g_pre;

template <class T>
const DivergencePoint<T> &dp(const Point<T> *p)
{
	return dynamic_cast<const DivergencePoint<T> &>(*p);
}

template <class T>
mc2_model_desc describe(const Feature<T> &feat, const matrix::Matrix &weights)
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
	d.bias = Predictor<T>::classify_sum(0) - 0.5; // classify_sum(0) = logistic(0) + _bias = 0.5 + _bias
	return d;
}

// The device mirror of the points is keyed by Point::get_id() and filled lazily.  CRunner assigns the final ids after the
// Trainer was built when --no-train-list files are given (src/cluster/CRunner.cpp:576-592: the extra points are appended, the
// whole vector re-sorted and re-numbered), so a row may hold the histogram of a point that no longer carries that id.  Every
// real point is therefore checked against the object its row was filled from; the first mismatch drops the whole mirror and
// the rows are refilled from the objects actually seen.  Center clones carry the id of the point whose bins they copied
// (DivergencePoint::set copies points + id, src/clutil/DivergencePoint.cpp:182-190), so a clone can fill a missing row too.
template <class T>
void grow_points(Device &d, uint64_t need)
{
	if (need <= d.cap) {
		return;
	}
	uint64_t cap = std::max<uint64_t>(need, d.cap + d.cap / 2 + 1024);
	const size_t N = (size_t)1 << (2 * d.k);
	mc2_hset *fresh = nullptr;
	{
		std::vector<T> ones((size_t)cap * N, 1);
		std::vector<uint64_t> len1(cap, 1);
		ok(mc2_hset_from_host(d.ctx, ones.data(), cap, d.k, (int)sizeof(T), nullptr, len1.data(), &fresh));
	}
	if (d.points) {
		std::vector<uint64_t> idx, mag, len;
		for (uint64_t r = 0; r < d.present.size(); r++) {
			if (d.present[r]) {
				idx.push_back(r);
				mag.push_back(d.hmag[r]);
				len.push_back(d.hlen[r]);
			}
		}
		if (!idx.empty()) {
			ok(mc2_hset_assign_rows(d.ctx, fresh, idx.size(), idx.data(), d.points, idx.data(), mag.data(), len.data()));
		}
		mc2_hset_free(d.points);
	}
	d.points = fresh;
	d.cap = cap;
	d.present.resize(cap, 0);
	d.owner.resize(cap, nullptr);
	d.hmag.resize(cap, 0);
	d.hlen.resize(cap, 0);
}

// rows of `objs` in the device mirror; real = the objects are the points themselves (not center clones)
template <class T>
void rows_of(Device &d, const std::vector<Point<T> *> &objs, bool real, std::vector<uint64_t> &rows)
{
	rows.resize(objs.size());
	for (int attempt = 0; attempt < 2; attempt++) {
		std::vector<size_t> missing;
		bool stale = false;
		uint64_t max_id = 0;
		for (size_t i = 0; i < objs.size(); i++) {
			const uint64_t id = objs[i]->get_id();
			rows[i] = id;
			max_id = std::max(max_id, id);
			if (id >= d.cap || !d.present[id]) {
				missing.push_back(i);
			} else if (real && d.owner[id] != nullptr && d.owner[id] != (const void *)objs[i]) {
				stale = true;
				break;
			}
		}
		if (stale) {
			// the ids were re-assigned after the rows were filled: forget everything, refill from what is seen
			std::fill(d.present.begin(), d.present.end(), 0);
			std::fill(d.owner.begin(), d.owner.end(), nullptr);
			continue;
		}
		if (missing.empty()) {
			return;
		}
		grow_points<T>(d, max_id + 1);
		// one upload for all missing rows of this call (a row may be asked for twice: keep the first)
		const size_t N = (size_t)1 << (2 * d.k);
		std::vector<size_t> uniq;
		for (size_t i : missing) {
			const uint64_t id = objs[i]->get_id();
			if (!d.present[id]) {
				d.present[id] = 1;
				d.owner[id] = real ? (const void *)objs[i] : nullptr;
				uniq.push_back(i);
			}
		}
		std::vector<T> bins(uniq.size() * N);
		std::vector<uint64_t> mag(uniq.size()), len(uniq.size()), dst(uniq.size()), src(uniq.size());
		for (size_t u = 0; u < uniq.size(); u++) {
			const DivergencePoint<T> &q = dp<T>(objs[uniq[u]]);
			const uint64_t id = objs[uniq[u]]->get_id();
			std::copy(q.points.begin(), q.points.end(), bins.begin() + u * N);
			// the row's own magnitude is the true bin sum: queries and centers pass theirs explicitly (quirk Q4)
			uint64_t sum = 0;
			for (const T &v : q.points) {
				sum += (uint64_t)v;
			}
			mag[u] = d.hmag[id] = real ? q.getPseudoMagnitude() : sum;
			len[u] = d.hlen[id] = q.get_length();
			dst[u] = id;
			src[u] = u;
		}
		mc2_hset *tmp = nullptr;
		ok(mc2_hset_from_host(d.ctx, bins.data(), uniq.size(), d.k, (int)sizeof(T), mag.data(), len.data(), &tmp));
		const int rc = mc2_hset_assign_rows(d.ctx, d.points, uniq.size(), dst.data(), tmp, src.data(), mag.data(), len.data());
		mc2_hset_free(tmp);
		ok(rc);
		return;
	}
	throw std::runtime_error("meshclust2_b200 integration: point ids changed while a batch was being staged");
}

template <class T>
uint64_t row_of(Device &d, Point<T> *obj, bool real)
{
	std::vector<Point<T> *> one{obj};
	std::vector<uint64_t> rows;
	rows_of<T>(d, one, real, rows);
	return rows[0];
}

// lazily mirror the Trainer's state on the device: the model once, the points as they are seen
template <class T>
Device &device_for(const void *key, const Feature<T> &feat, const matrix::Matrix &weights, const std::vector<Point<T> *> &points,
		   int k)
{
	Device &d = g_dev[key];
	if (d.ctx) {
		return d;
	}
	unsigned long once = 0;
	Stopwatch sw(g_time.init, once);
	d.k = k;
	d.n = points.size();
	auto pre = g_pre.find(key);
	if (pre != g_pre.end() && pre->second.started) {
		Prewarm &w = pre->second;
		w.cancel = true; // a thread still waiting for the device mutex (held by our caller) gives up; a finished one is joined
		if (w.th.joinable()) {
			w.th.join();
		}
		if (w.err.empty() && w.ctx && w.points) {
			// rows were uploaded in the order of the Trainer's vector; rows_of() verifies every row against its object
			d.ctx = w.ctx;
			d.points = w.points;
			d.cap = points.size();
			d.present.assign(d.cap, 1);
			d.owner.assign(d.cap, nullptr);
			d.hmag.resize(d.cap);
			d.hlen.resize(d.cap);
			for (size_t i = 0; i < points.size(); i++) {
				d.owner[i] = (const void *)points[i];
				d.hmag[i] = dp<T>(points[i]).getPseudoMagnitude();
				d.hlen[i] = points[i]->get_length();
				if (points[i]->get_id() != i) {
					d.present[i] = 0; // not where its id says: will be refilled on first use
					d.owner[i] = nullptr;
				}
			}
		} else if (w.points) {
			mc2_hset_free(w.points);
		}
		w.ctx = nullptr;
		w.points = nullptr;
		w.started = false;
	}
	if (!d.ctx) {
		d.ctx = mc2i::shared_ctx();
	}
	if (!d.points) {
		grow_points<T>(d, std::max<uint64_t>(points.size(), 1));
		std::vector<uint64_t> rows;
		rows_of<T>(d, points, true, rows); // the training-time points in one upload
	}
	mc2_model_desc desc = describe<T>(feat, weights);
	ok(mc2_model_create(d.ctx, &desc, &d.model));
	d.scratch_rows = 64;
	const size_t N = (size_t)1 << (2 * k);
	std::vector<T> zero(d.scratch_rows * N, 1);
	std::vector<uint64_t> ones(d.scratch_rows, 1);
	ok(mc2_hset_from_host(d.ctx, zero.data(), d.scratch_rows, k, (int)sizeof(T), nullptr, ones.data(), &d.scratch));
	return d;
}

template <class T>
void start_prewarm(const void *key, const std::vector<Point<T> *> &points, int k)
{
	if (std::getenv("MC2_NO_PREWARM")) {
		return;
	}
	std::lock_guard<std::mutex> lock(g_mu);
	Prewarm &w = g_pre[key];
	if (w.started || g_dev.count(key)) {
		return;
	}
	w.started = true;
	Prewarm *wp = &w; // std::map nodes are stable
	const std::vector<Point<T> *> *pts = &points;
	w.th = std::thread([wp, pts, k]() {
		try {
			const size_t n = pts->size(), N = (size_t)1 << (2 * k);
			std::vector<T> bins(n * N);
			std::vector<uint64_t> mag(n), len(n);
			for (size_t i = 0; i < n; i++) {
				const DivergencePoint<T> &q = dp<T>((*pts)[i]);
				std::copy(q.points.begin(), q.points.end(), bins.begin() + i * N);
				mag[i] = q.getPseudoMagnitude();
				len[i] = q.get_length();
			}
			// one caller per context (include/meshclust2_b200.h): the reader of --no-train-list files may be using the
			// shared context right now.  device_for() runs with the mutex held and sets `cancel` before it joins.
			std::unique_lock<std::mutex> lk(mc2i::device_mutex(), std::defer_lock);
			while (!lk.try_lock()) {
				if (wp->cancel) {
					return;
				}
				std::this_thread::sleep_for(std::chrono::microseconds(100));
			}
			wp->ctx = mc2i::shared_ctx();
			if (mc2_hset_from_host(wp->ctx, bins.data(), n, k, (int)sizeof(T), mag.data(), len.data(), &wp->points) != MC2_OK) {
				throw std::runtime_error(mc2_last_error());
			}
		} catch (const std::exception &e) {
			wp->err = e.what();
		}
	});
}

// put host center objects into rows 0..m-1 of `into`: bins of the point whose id they carry + their own mag / length
template <class T>
void stage_into(Device &d, mc2_hset *into, const std::vector<Point<T> *> &cs)
{
	const uint64_t m = cs.size();
	std::vector<uint64_t> dst(m), src, mag(m), len(m);
	rows_of<T>(d, cs, false, src);
	for (uint64_t i = 0; i < m; i++) {
		dst[i] = i;
		mag[i] = dp<T>(cs[i]).getPseudoMagnitude();
		len[i] = cs[i]->get_length();
	}
	ok(mc2_hset_assign_rows(d.ctx, into, m, dst.data(), d.points, src.data(), mag.data(), len.data()));
}

template <class T>
void stage_centers(Device &d, const std::vector<Point<T> *> &cs)
{
	if (cs.size() > d.scratch_rows) {
		throw std::runtime_error("meshclust2_b200 integration: --delta too large for the scratch set");
	}
	stage_into<T>(d, d.scratch, cs);
}

// all centers of a pass, staged into a set that is sized on first use (the number of centers only shrinks afterwards)
template <class T>
void stage_all_centers(Device &d, std::vector<Center<T>> &part)
{
	const uint64_t n = part.size();
	if (n > d.center_rows) {
		if (d.centers) {
			mc2_hset_free(d.centers);
			d.centers = nullptr;
		}
		const size_t N = (size_t)1 << (2 * d.k);
		std::vector<T> ones(n * N, 1);
		std::vector<uint64_t> len1(n, 1);
		ok(mc2_hset_from_host(d.ctx, ones.data(), n, d.k, (int)sizeof(T), nullptr, len1.data(), &d.centers));
		d.center_rows = n;
	}
	std::vector<Point<T> *> cs(n);
	for (uint64_t j = 0; j < n; j++) {
		cs[j] = part[j].getCenter();
	}
	stage_into<T>(d, d.centers, cs);
}

using mc2i::batching_enabled;

} // namespace

template <class T>
std::tuple<Point<T> *, double, size_t, size_t> Trainer<T>::get_close(Point<T> *p, bvec_iterator<T> istart, bvec_iterator<T> iend,
									bool &is_min_r) const
{
	std::lock_guard<std::mutex> lock(g_mu);
	Device &d = device_for<T>(this, *feat, weights, points, k);
	Stopwatch sw(g_time.get_close, g_time.n_get_close);
	std::vector<uint64_t> cand;
	std::vector<Point<T> *> cand_pts;
	std::vector<bvec_iterator<T>> where;
	// same trip count as the reference's `omp parallel for` over the iterator range: iend - istart (bvec_iterator::operator-),
	// which is 0 when only empty bins lie between the two positions
	const int64_t n_iter = iend - istart;
	bvec_iterator<T> i = istart;
	for (int64_t t = 0; t < n_iter; t++) {
		cand_pts.push_back((*i).first);
		where.push_back(i);
		if (t + 1 < n_iter) {
			++i;
		}
	}
	std::tuple<Point<T> *, double, size_t, size_t> result(NULL, -1, 0, 0);
	is_min_r = true;
	if (cand_pts.empty()) {
		return result;
	}
	rows_of<T>(d, cand_pts, true, cand);
	const uint64_t qrow = row_of<T>(d, p, false); // the query may be a clone: addressed by the id it carries
	int64_t best = -1;
	double best_dist = -1;
	int32_t is_min = 1;
	std::vector<uint8_t> marks(cand.size());
	// the query is the row of the point whose bins it carries, with its own (possibly stale) magnitude and length
	ok(mc2_get_close_as(d.ctx, d.model, d.points, qrow, dp<T>(p).getPseudoMagnitude(), p->get_length(), d.points,
			    cand.data(), 0, cand.size(), cutoff, &best, &best_dist, &is_min, marks.data()));
	for (size_t j = 0; j < cand.size(); j++) {
		if (marks[j]) {
			bvec_iterator<T> it = where[j];
			*it = std::make_pair((*it).first, true);
		}
	}
	if (best >= 0) {
		bvec_iterator<T> it = where[(size_t)best];
		result = std::make_tuple((*it).first, best_dist, it.r, it.c);
	}
	is_min_r = is_min != 0;
	return result;
}

template <class T>
long Trainer<T>::merge(vector<Center<T>> &centers, long current, long begin, long last) const
{
	if (last < begin) {
		return 0;
	}
	std::lock_guard<std::mutex> lock(g_mu);
	Device &d = device_for<T>(this, *feat, weights, points, k);
	Stopwatch sw(g_time.merge, g_time.n_merge);
	std::vector<Point<T> *> cs;
	cs.push_back(centers[current].getCenter());
	for (long i = begin; i <= last; i++) {
		cs.push_back(centers[i].getCenter());
	}
	stage_centers<T>(d, cs);
	std::vector<uint64_t> rows(cs.size());
	for (size_t i = 0; i < rows.size(); i++) {
		rows[i] = i;
	}
	int64_t out = 0;
	const int rc = mc2_merge(d.ctx, d.model, d.scratch, rows.data(), 0, 1, (int64_t)cs.size() - 1, get_id(), &out);
	if (rc == MC2_ERR_UNSUPPORTED) {
		// --bias outside [-0.5, 0.5): the device's close flag (round(score) > 0) is not the reference's rule
		// (round(classify_sum(sum)) == 1, src/cluster/Trainer.cpp:100-103).  Scores and first-combo values still come from
		// the device; the rule is applied here exactly as the reference's loop applies it (later tie wins, :104).
		const size_t m = cs.size() - 1;
		mc2_pairs pr = mc2_pairs();
		pr.set_a = pr.set_b = d.scratch;
		pr.n_pairs = m;
		pr.ia = rows.data() + 1;
		pr.b_begin = 0;
		pr.b_broadcast = 1;
		pr.len_filter = 1;
		pr.anchor_is_b = 1;
		pr.cutoff = get_id();
		std::vector<double> score(m), dist(m);
		std::vector<uint8_t> skipped(m);
		ok(mc2_score_pairs(d.ctx, d.model, &pr, score.data(), dist.data(), nullptr, nullptr, nullptr, skipped.data()));
		std::pair<long, double> best = std::make_pair(0L, std::numeric_limits<double>::min());
		for (size_t j = 0; j < m; j++) {
			if (!skipped[j] && round(score[j]) == 1) {
				best = best.second > dist[j] ? best : std::make_pair(begin + (long)j, dist[j]);
			}
		}
		return best.first;
	}
	ok(rc);
	return out == 0 ? 0 : begin + (out - 1);
}

template <class T>
double Trainer<T>::classify(Point<T> *a, Point<T> *b) const
{
	std::lock_guard<std::mutex> lock(g_mu);
	Device &d = device_for<T>(this, *feat, weights, points, k);
	std::vector<Point<T> *> two{a, b};
	stage_centers<T>(d, two);
	mc2_pairs pr = mc2_pairs();
	pr.set_a = pr.set_b = d.scratch;
	pr.n_pairs = 1;
	uint64_t ia = 0, ib = 1;
	pr.ia = &ia;
	pr.ib = &ib;
	double score = 0;
	ok(mc2_score_pairs(d.ctx, d.model, &pr, &score, nullptr, nullptr, nullptr, nullptr, nullptr));
	return score;
}

template <class T>
void Trainer<T>::filter(Point<T> *p, vector<pair<Point<T> *, bool>> &vec) const
{
	if (vec.empty()) {
		return;
	}
	std::vector<uint8_t> keep(vec.size());
	{
		std::lock_guard<std::mutex> lock(g_mu);
		Device &d = device_for<T>(this, *feat, weights, points, k);
		Stopwatch sw(g_time.filter, g_time.n_filter);
		std::vector<uint64_t> rows;
		std::vector<Point<T> *> pts(vec.size());
		for (size_t j = 0; j < vec.size(); j++) {
			pts[j] = vec[j].first;
		}
		rows_of<T>(d, pts, true, rows);
		const uint64_t crow = row_of<T>(d, p, false);
		ok(mc2_filter_as(d.ctx, d.model, d.points, crow, dp<T>(p).getPseudoMagnitude(), p->get_length(), d.points,
				 rows.data(), rows.size(), get_id(), keep.data()));
	}
	size_t w = 0;
	for (size_t j = 0; j < vec.size(); j++) {
		if (keep[j]) {
			vec[w] = vec[j];
			vec[w].second = false;
			w++;
		}
	}
	vec.resize(w);
}

// Trainer<T>::closest: the mean was built on the host by mean_shift_update; the distance_d arg-min runs on the device (K3)
template <class T>
Point<T> *Trainer<T>::closest(Point<double> *p, vector<pair<Point<T> *, bool>> &vec) const
{
	if (vec.empty()) {
		return NULL;
	}
	std::lock_guard<std::mutex> lock(g_mu);
	Device &d = device_for<T>(this, *feat, weights, points, k);
	Stopwatch sw(g_time.closest, g_time.n_closest);
	std::vector<uint64_t> rows;
	std::vector<Point<T> *> pts(vec.size());
	for (size_t j = 0; j < vec.size(); j++) {
		pts[j] = vec[j].first;
	}
	rows_of<T>(d, pts, true, rows);
	const std::vector<double> &mean = p->get_data();
	int64_t best = -1;
	double bd = 0;
	ok(mc2_closest(d.ctx, d.points, rows.data(), rows.size(), mean.data(), &best, &bd, nullptr));
	return best < 0 ? NULL : vec[(size_t)best].first;
}

// training stays on the host exactly as in the reference: Predictor builds the model, Trainer keeps a copy
template <class T>
void Trainer<T>::train(std::string dump_str)
{
	start_prewarm<T>(this, points, k);
	Predictor<T> *pred = new Predictor<T>(dump_str); // never destroyed: the reference's file ctor leaves members unset
	auto pr = pred->get_class();
	delete feat;
	feat = pr.first;
	feat->set_save(false);
	weights = pr.second.get_weights();
}

template <class T>
void Trainer<T>::train(int min_n_feat, int max_n_feat, uint64_t feat_type, int mut_type, double min_id, std::string dump_str,
		       double acc_cutoff)
{
	(void)acc_cutoff;
	if (dump_str == "") {
		start_prewarm<T>(this, points, k); // not for --dump runs, which exit right after training
	}
	std::cout << "Splitting data" << endl;
	uintmax_t next_id = points.size();
	Predictor<T> pred(k, cutoff, PRED_MODE_CLASS, feat_type, mut_type, min_n_feat, max_n_feat, min_id);
	pred.train(points, next_id, n_samples, n_templates);
	auto pr = pred.get_class();
	delete feat;
	feat = pr.first;
	weights = pr.second.get_weights();
	const bool dump_only = dump_str != "";
	pred.save(dump_only ? dump_str : std::string("weights.txt"), Datatype::get());
	if (dump_only) {
		exit(0);
	}
}

// ---- batched update stage -------------------------------------------------------------------------------------------
// One pass of `for j: mean_shift_update(part, j, trn, delta)` (src/cluster/ClusterFactory.cpp:639-642 and :648-651).  The
// iterations are independent: each reads the point lists of clusters j-delta..j+delta and its own center, and writes only
// its own center.  Member lists are built as mean_shift_update builds `good` (:292-307), the device filters them, averages
// the survivors and picks the closest one; the host applies center->set(*next) (:328) or, when nothing survives and
// delta == 0, center->set(*first) (:329-332).
template <class T>
bool mc2_batched_update(std::vector<Center<T>> &part, const Trainer<T> &trn, int delta)
{
	if (!batching_enabled() || part.empty()) {
		return false;
	}
	std::lock_guard<std::mutex> lock(g_mu);
	auto it = g_dev.find(&trn);
	if (it == g_dev.end() || !it->second.ctx) {
		return false; // no device mirror yet: the per-center path creates it
	}
	Device &d = it->second;
	Stopwatch sw(g_time.update_batch, g_time.n_update_batch);
	const long n = (long)part.size();
	std::vector<uint64_t> off((size_t)n + 1, 0), members;
	std::vector<Point<T> *> who;
	for (long j = 0; j < n; j++) {
		const long i_begin = std::max(0L, j - delta), i_end = std::min(j + (long)delta, n - 1);
		for (long i = i_begin; i <= i_end; i++) {
			for (Point<T> *p : part[i].getPoints()) {
				who.push_back(p);
			}
		}
		off[(size_t)j + 1] = who.size();
	}
	rows_of<T>(d, who, true, members);
	stage_all_centers<T>(d, part);
	std::vector<int64_t> next((size_t)n);
	const int rc = mc2_update_centers(d.ctx, d.model, d.centers, (uint64_t)n, d.points, off.data(), members.data(), trn.get_id(),
					  next.data(), nullptr);
	if (rc == MC2_ERR_UNSUPPORTED) {
		return false; // e.g. --bias below -0.5: the one-center calls handle it
	}
	ok(rc);
	for (long j = 0; j < n; j++) {
		Point<T> *center = part[j].getCenter();
		if (next[(size_t)j] >= 0) {
			center->set(*who[off[(size_t)j] + (uint64_t)next[(size_t)j]]);
		} else if (delta == 0) {
			center->set(*part[j].getPoints()[0]);
		}
	}
	return true;
}

// One call of merge() (src/cluster/ClusterFactory.cpp:382-401).  Trainer::merge reads only the center points, which the
// loop never changes, so every center's choice comes from one device call; the list splicing and the erase stay as they are.
template <class T>
bool mc2_batched_merge(std::vector<Center<T>> &centers, const Trainer<T> &trn, int delta)
{
	if (!batching_enabled() || centers.empty()) {
		return false;
	}
	std::vector<int64_t> ret(centers.size());
	{
		std::lock_guard<std::mutex> lock(g_mu);
		auto it = g_dev.find(&trn);
		if (it == g_dev.end() || !it->second.ctx) {
			return false;
		}
		Device &d = it->second;
		Stopwatch sw(g_time.merge_batch, g_time.n_merge_batch);
		stage_all_centers<T>(d, centers);
		const int rc = mc2_merge_centers(d.ctx, d.model, d.centers, centers.size(), delta, trn.get_id(), ret.data());
		if (rc == MC2_ERR_UNSUPPORTED) {
			return false; // --bias outside [-0.5, 0.5): the per-center calls apply the reference's rule on the host
		}
		ok(rc);
	}
	for (size_t i = 0; i < centers.size(); i++) {
		if (ret[i] > (int64_t)i) {
			auto &to_add = centers[(size_t)ret[i]].getPoints();
			auto &to_del = centers[i].getPoints();
			to_add.insert(std::end(to_add), std::begin(to_del), std::end(to_del));
			centers[i].lazy_remove();
		}
	}
	centers.erase(std::remove_if(centers.begin(), centers.end(), [](const Center<T> &c) { return c.is_delete(); }), centers.end());
	return true;
}

#define MC2_INSTANTIATE_BATCH(T)                                                                         \
	template bool mc2_batched_update<T>(std::vector<Center<T>> &, const Trainer<T> &, int);          \
	template bool mc2_batched_merge<T>(std::vector<Center<T>> &, const Trainer<T> &, int);
MC2_INSTANTIATE_BATCH(uint8_t)
MC2_INSTANTIATE_BATCH(uint16_t)
MC2_INSTANTIATE_BATCH(uint32_t)
MC2_INSTANTIATE_BATCH(uint64_t)
MC2_INSTANTIATE_BATCH(int)
MC2_INSTANTIATE_BATCH(double)

template class Trainer<uint8_t>;
template class Trainer<uint16_t>;
template class Trainer<uint32_t>;
template class Trainer<uint64_t>;
template class Trainer<int>;
template class Trainer<double>;
