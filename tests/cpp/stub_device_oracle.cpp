// stub_device_oracle.cpp -- TEST INFRASTRUCTURE ONLY.  A stand-in for libmeshclust2_b200.so that serves the subset of the C ABI
// the relinked meshclust2 uses (integration/*.cpp) from the CPU oracle (oracle/libmc2oracle.so), so that the HOST logic of the
// integration -- the build-time hunks, the FASTA record splitting, the batched update / merge passes, the point objects --
// can be run end to end without a GPU (tests/test_integration_cpu.py).  It is linked only into oracle/_ref/meshclust2_stub
// (`make -C oracle integrated_stub`); the product library never sees it, and nothing measured ever runs through it.
// Same role as the OracleEngine stand-in of tests/test_dist_gloo.py.
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "meshclust2_b200.h"
#include "mc2_oracle.h"

struct mc2_ctx {
	int unused;
};

struct mc2_seqs {
	uint64_t n = 0;
	std::string codes;              // 1 byte per base, as ChromosomeOneDigit leaves it
	std::vector<uint64_t> seq_off;  // [n + 1]
	std::vector<int32_t> segs;      // inclusive pairs
	std::vector<uint64_t> seg_off;  // [n + 1]
};

struct mc2_hset {
	uint64_t n = 0, N = 0;
	int k = 0, eb = 0;
	std::vector<unsigned char> bins; // (n + 1) x N x eb: one spare row for a center addressed "as" a row with its own side-band
	std::vector<uint64_t> mag, len;  // [n + 1]
	std::vector<uint64_t> mers1;     // [4 n]
	std::vector<double> stddev;
	std::vector<int32_t> novf;
	std::vector<uint32_t> maxc;
	void *row(uint64_t r) { return bins.data() + r * N * (uint64_t)eb; }
	const void *row(uint64_t r) const { return bins.data() + r * N * (uint64_t)eb; }
};

struct mc2_model {
	mc2o_model m;
	int regression;
};

namespace {

thread_local std::string g_err;

int fail(int rc, const char *msg)
{
	g_err = msg;
	return rc;
}

mc2_hset *new_hset(uint64_t n, int k, int eb)
{
	mc2_hset *h = new mc2_hset();
	h->n = n;
	h->k = k;
	h->eb = eb;
	h->N = (uint64_t)1 << (2 * k);
	h->bins.assign((n + 1) * h->N * (uint64_t)eb, 0);
	h->mag.assign(n + 1, 0);
	h->len.assign(n + 1, 0);
	h->mers1.assign(4 * n, 0);
	h->stddev.assign(n, 0);
	h->novf.assign(n, 0);
	h->maxc.assign(n, 0);
	return h;
}

// put (bins of src row, mag, len) into the spare row of `into`; returns its index
uint64_t stage_spare(mc2_hset *into, const mc2_hset *src, uint64_t src_row, uint64_t mag, uint64_t len)
{
	std::memcpy(into->row(into->n), src->row(src_row), into->N * (uint64_t)into->eb);
	into->mag[into->n] = mag;
	into->len[into->n] = len;
	return into->n;
}

mc2o_point point_of(const mc2_hset *h, uint64_t r)
{
	mc2o_point p;
	p.bins = h->row(r);
	p.mag = h->mag[r];
	p.len = h->len[r];
	return p;
}

int count_one(const mc2_seqs *s, uint64_t i, int k, int eb, mc2_hset *h)
{
	const int nseg = (int)(s->seg_off[i + 1] - s->seg_off[i]);
	const int *segs = s->segs.data() + 2 * s->seg_off[i];
	int novf = 0;
	if (mc2o_count(s->codes.data() + s->seq_off[i], segs, nseg, k, eb, h->row(i), h->mers1.data() + 4 * i, &novf) != 0) {
		return fail(MC2_ERR_INPUT, "stub: invalid code inside a segment");
	}
	h->novf[i] = novf;
	uint64_t mag = 0;
	double sd = 0;
	mc2o_point_stats(h->row(i), h->N, eb, &mag, &sd);
	h->mag[i] = mag;
	h->stddev[i] = sd;
	uint64_t len = 0;
	for (int g = 0; g < nseg; g++) {
		len += (uint64_t)(segs[2 * g + 1] - segs[2 * g] + 1);
	}
	h->len[i] = len;
	return MC2_OK;
}

} // namespace

extern "C" {

const char *mc2_last_error(void)
{
	return g_err.c_str();
}

int mc2_ctx_create(int, mc2_ctx **out)
{
	*out = new mc2_ctx();
	return MC2_OK;
}

void mc2_ctx_destroy(mc2_ctx *ctx)
{
	delete ctx;
}

/* ---- sequences ---- */
int mc2_seqs_upload(mc2_ctx *, const char *codes, const uint64_t *seq_off, uint64_t n, const int32_t *segs, const uint64_t *seg_off,
		    mc2_seqs **out)
{
	mc2_seqs *s = new mc2_seqs();
	s->n = n;
	s->seq_off.assign(seq_off, seq_off + n + 1);
	s->seg_off.assign(seg_off, seg_off + n + 1);
	s->codes.assign(codes, codes + seq_off[n]);
	s->segs.assign(segs, segs + 2 * seg_off[n]);
	*out = s;
	return MC2_OK;
}

int mc2_seqs_from_text(mc2_ctx *, const char *text, const uint64_t *seq_off, uint64_t n, mc2_seqs **out)
{
	mc2_seqs *s = new mc2_seqs();
	s->n = n;
	s->seq_off.assign(seq_off, seq_off + n + 1);
	s->seg_off.assign(n + 1, 0);
	s->codes.assign(seq_off[n], '\0');
	for (uint64_t i = 0; i < n; i++) {
		const long len = (long)(seq_off[i + 1] - seq_off[i]);
		std::vector<int> segs((size_t)(len / 10 + 4) * 2);
		int nseg = 0;
		long eff = 0;
		const int rc = mc2o_encode(text + seq_off[i], len, &s->codes[seq_off[i]], segs.data(), (int)(segs.size() / 2), &nseg, &eff);
		if (rc != 0) {
			delete s;
			return fail(rc == -1 ? MC2_ERR_INPUT : MC2_ERR_ARG, "stub: mc2o_encode rejected a sequence");
		}
		s->segs.insert(s->segs.end(), segs.begin(), segs.begin() + 2 * nseg);
		s->seg_off[i + 1] = s->segs.size() / 2;
	}
	*out = s;
	return MC2_OK;
}

uint64_t mc2_seqs_total_segments(const mc2_seqs *s)
{
	return s ? s->segs.size() / 2 : 0;
}

int mc2_seqs_download_segments(mc2_ctx *, const mc2_seqs *s, int32_t *segs_out, uint64_t *seg_off_out, uint64_t *lengths_out)
{
	if (segs_out) std::memcpy(segs_out, s->segs.data(), s->segs.size() * sizeof(int32_t));
	if (seg_off_out) std::memcpy(seg_off_out, s->seg_off.data(), s->seg_off.size() * sizeof(uint64_t));
	if (lengths_out) {
		for (uint64_t i = 0; i < s->n; i++) lengths_out[i] = s->seq_off[i + 1] - s->seq_off[i];
	}
	return MC2_OK;
}

void mc2_seqs_free(mc2_seqs *s)
{
	delete s;
}

/* ---- histograms ---- */
int mc2_count_kmers(mc2_ctx *, const mc2_seqs *s, int k, int eb, mc2_hset **out)
{
	mc2_hset *h = new_hset(s->n, k, eb);
	for (uint64_t i = 0; i < s->n; i++) {
		const int rc = count_one(s, i, k, eb, h);
		if (rc != MC2_OK) {
			delete h;
			return rc;
		}
	}
	*out = h;
	return MC2_OK;
}

int mc2_count_kmers_auto(mc2_ctx *ctx, const mc2_seqs *s, int k, uint64_t *largest_count, int *elem_bytes, mc2_hset **out)
{
	uint64_t largest = 0;
	for (uint64_t i = 0; i < s->n; i++) {
		uint64_t l = 0;
		const int nseg = (int)(s->seg_off[i + 1] - s->seg_off[i]);
		if (mc2o_largest_count(s->codes.data() + s->seq_off[i], s->segs.data() + 2 * s->seg_off[i], nseg, k, &l) != 0) {
			return fail(MC2_ERR_INPUT, "stub: a segment is shorter than k");
		}
		largest = l > largest ? l : largest;
	}
	*largest_count = largest;
	*elem_bytes = mc2o_width_for(largest);
	return mc2_count_kmers(ctx, s, k, *elem_bytes, out);
}

int mc2_hset_from_host(mc2_ctx *, const void *bins, uint64_t n, int k, int eb, const uint64_t *mag, const uint64_t *len, mc2_hset **out)
{
	mc2_hset *h = new_hset(n, k, eb);
	std::memcpy(h->bins.data(), bins, n * h->N * (uint64_t)eb);
	for (uint64_t i = 0; i < n; i++) {
		if (mag) {
			h->mag[i] = mag[i];
		} else {
			double sd;
			mc2o_point_stats(h->row(i), h->N, eb, &h->mag[i], &sd);
		}
		h->len[i] = len ? len[i] : 0;
	}
	*out = h;
	return MC2_OK;
}

void mc2_hset_free(mc2_hset *h)
{
	delete h;
}

int mc2_hset_download(mc2_ctx *, const mc2_hset *h, uint64_t first, uint64_t count, void *bins, uint64_t *mag, uint64_t *len,
		      uint64_t *mers1, double *stddev, int32_t *n_overflow, uint32_t *max_count)
{
	if (first + count > h->n) return fail(MC2_ERR_ARG, "stub: download range");
	if (bins) std::memcpy(bins, h->row(first), count * h->N * (uint64_t)h->eb);
	for (uint64_t i = 0; i < count; i++) {
		if (mag) mag[i] = h->mag[first + i];
		if (len) len[i] = h->len[first + i];
		if (stddev) stddev[i] = h->stddev[first + i];
		if (n_overflow) n_overflow[i] = h->novf[first + i];
		if (max_count) max_count[i] = h->maxc[first + i];
		if (mers1) std::memcpy(mers1 + 4 * i, h->mers1.data() + 4 * (first + i), 4 * sizeof(uint64_t));
	}
	return MC2_OK;
}

int mc2_hset_assign_rows(mc2_ctx *, mc2_hset *dst, uint64_t n, const uint64_t *dst_rows, const mc2_hset *src, const uint64_t *src_rows,
			 const uint64_t *mag, const uint64_t *len)
{
	for (uint64_t i = 0; i < n; i++) {
		if (dst_rows[i] >= dst->n || src_rows[i] >= src->n) return fail(MC2_ERR_ARG, "stub: assign_rows range");
		std::memcpy(dst->row(dst_rows[i]), src->row(src_rows[i]), dst->N * (uint64_t)dst->eb);
		dst->mag[dst_rows[i]] = mag ? mag[i] : src->mag[src_rows[i]];
		dst->len[dst_rows[i]] = len ? len[i] : src->len[src_rows[i]];
	}
	return MC2_OK;
}

/* ---- model ---- */
int mc2_model_create(mc2_ctx *, const mc2_model_desc *d, mc2_model **out)
{
	mc2_model *m = new mc2_model();
	std::memset(&m->m, 0, sizeof m->m);
	m->m.n_singles = d->n_singles;
	for (int i = 0; i < d->n_singles; i++) {
		m->m.single_flag[i] = d->single_flag[i];
		m->m.single_min[i] = d->single_min[i];
		m->m.single_max[i] = d->single_max[i];
	}
	m->m.n_combos = d->n_combos;
	for (int c = 0; c < d->n_combos; c++) {
		m->m.combo_kind[c] = d->combo_kind[c];
		m->m.combo_nidx[c] = d->combo_nidx[c];
		for (int t = 0; t < d->combo_nidx[c]; t++) m->m.combo_idx[c][t] = d->combo_idx[c][t];
	}
	for (int c = 0; c <= d->n_combos; c++) m->m.weight[c] = d->weight[c];
	m->m.bias = d->bias;
	m->regression = d->regression;
	*out = m;
	return MC2_OK;
}

/* ---- scoring ---- */
int mc2_score_pairs(mc2_ctx *, const mc2_model *model, const mc2_pairs *p, double *score, double *dist, uint8_t *close, double *cache,
		    double *raw, uint8_t *skipped)
{
	if (cache || raw) return fail(MC2_ERR_UNSUPPORTED, "stub: cache / raw outputs are not served");
	for (uint64_t j = 0; j < p->n_pairs; j++) {
		const uint64_t ra = p->ia ? p->ia[j] : p->a_begin + (p->a_broadcast ? 0 : j);
		const uint64_t rb = p->ib ? p->ib[j] : p->b_begin + (p->b_broadcast ? 0 : j);
		mc2o_point a = point_of(p->set_a, ra), b = point_of(p->set_b, rb);
		if (p->len_filter) {
			const uint64_t anchor = p->anchor_is_b ? b.len : a.len, other = p->anchor_is_b ? a.len : b.len;
			if (other < (uint64_t)((double)anchor * p->cutoff) || other > (uint64_t)((double)anchor / p->cutoff)) {
				if (score) score[j] = NAN;
				if (dist) dist[j] = NAN;
				if (close) close[j] = 0;
				if (skipped) skipped[j] = 1;
				continue;
			}
		}
		double cch[MC2O_MAX_SINGLES], d0 = 0, sum = 0, sc = 0;
		int cl = 0;
		if (mc2o_score_pair(&model->m, p->set_a->eb, p->set_a->N, &a, &b, cch, &d0, &sum, &sc, &cl) != 0) {
			return fail(MC2_ERR_FEATURE, "stub: a single feature failed");
		}
		if (model->regression) { // Predictor::p_predict (src/predict/Predictor.cpp:284-300): the clamped sum, no logistic
			sc = sum < 0 ? 0 : (sum > 1 ? 1 : sum);
			cl = 0;
		}
		if (score) score[j] = sc;
		if (dist) dist[j] = d0;
		if (close) close[j] = (uint8_t)cl;
		if (skipped) skipped[j] = 0;
	}
	return MC2_OK;
}

int mc2_all_pairs(mc2_ctx *, const mc2_model *model, const mc2_hset *set_q, uint64_t q_begin, uint64_t q_end, const mc2_hset *set_d,
		  uint64_t d_begin, uint64_t d_end, int32_t upper_only, double cutoff, uint64_t max_out, uint64_t *out_q, uint64_t *out_d,
		  double *out_score, uint64_t *n_out, uint64_t *n_scored)
{
	uint64_t no = 0, ns = 0;
	for (uint64_t q = q_begin; q < q_end; q++) {
		mc2o_point b = point_of(set_q, q);
		const uint64_t lo = (uint64_t)((double)b.len * cutoff), hi = (uint64_t)((double)b.len / cutoff);
		for (uint64_t d = d_begin; d < d_end; d++) {
			if (upper_only && d <= q) continue;
			mc2o_point a = point_of(set_d, d);
			if (a.len < lo || a.len > hi) continue;
			double cch[MC2O_MAX_SINGLES], d0 = 0, sum = 0, sc = 0;
			int cl = 0;
			if (mc2o_score_pair(&model->m, set_d->eb, set_d->N, &a, &b, cch, &d0, &sum, &sc, &cl) != 0) {
				return fail(MC2_ERR_FEATURE, "stub: a single feature failed");
			}
			ns++;
			if (cl) {
				if (no < max_out) {
					if (out_q) out_q[no] = q;
					if (out_d) out_d[no] = d;
					if (out_score) out_score[no] = sc;
				}
				no++;
			}
		}
	}
	*n_out = no;
	if (n_scored) *n_scored = ns;
	return MC2_OK;
}

int mc2_get_close_as(mc2_ctx *, const mc2_model *model, const mc2_hset *set_q, uint64_t q, uint64_t q_mag, uint64_t q_len,
		     const mc2_hset *set_c, const uint64_t *cand, uint64_t cand_begin, uint64_t n_cand, double cutoff, int64_t *best,
		     double *best_dist, int32_t *is_min, uint8_t *marks)
{
	mc2_hset *c = const_cast<mc2_hset *>(set_c);
	const uint64_t qi = stage_spare(c, set_q, q, q_mag, q_len);
	std::vector<uint64_t> rows;
	if (!cand) {
		for (uint64_t j = 0; j < n_cand; j++) rows.push_back(cand_begin + j);
		cand = rows.data();
	}
	int ismin = 0;
	if (mc2o_get_close(&model->m, c->eb, c->N, c->bins.data(), c->mag.data(), c->len.data(), qi, n_cand, cand, cutoff, best, best_dist,
			   &ismin, marks) != 0) {
		return fail(MC2_ERR_FEATURE, "stub: mc2o_get_close failed");
	}
	*is_min = ismin;
	return MC2_OK;
}

int mc2_filter_as(mc2_ctx *, const mc2_model *model, const mc2_hset *set_c, uint64_t center, uint64_t c_mag, uint64_t c_len,
		  const mc2_hset *set_m, const uint64_t *members, uint64_t n_members, double id, uint8_t *keep)
{
	mc2_hset *m = const_cast<mc2_hset *>(set_m);
	const uint64_t ci = stage_spare(m, set_c, center, c_mag, c_len);
	if (n_members == 0) return MC2_OK;
	if (mc2o_filter(&model->m, m->eb, m->N, m->bins.data(), m->mag.data(), m->len.data(), ci, n_members, members, id, keep) != 0) {
		return fail(MC2_ERR_FEATURE, "stub: mc2o_filter failed");
	}
	return MC2_OK;
}

int mc2_merge(mc2_ctx *, const mc2_model *model, const mc2_hset *centers, const uint64_t *rows, int64_t cur, int64_t begin, int64_t last,
	      double id, int64_t *out)
{
	*out = 0;
	if (last < begin) return MC2_OK;
	long o = 0;
	if (mc2o_merge(&model->m, centers->eb, centers->N, centers->bins.data(), centers->mag.data(), centers->len.data(), rows, (long)cur,
		       (long)begin, (long)last, id, &o) != 0) {
		return fail(MC2_ERR_FEATURE, "stub: mc2o_merge failed");
	}
	*out = o;
	return MC2_OK;
}

int mc2_closest(mc2_ctx *, const mc2_hset *set, const uint64_t *members, uint64_t n, const double *mean, int64_t *best, double *best_dist,
		double *dist_out)
{
	if (n == 0) return fail(MC2_ERR_ARG, "stub: empty member list");
	*best = -1;
	for (uint64_t j = 0; j < n; j++) {
		const double d = mc2o_distance_d(set->eb, set->N, set->row(members[j]), mean);
		if (dist_out) dist_out[j] = d;
		if (*best < 0 || d < *best_dist) {
			*best = (int64_t)j;
			*best_dist = d;
		}
	}
	return MC2_OK;
}

/* ---- batched update / merge stage ---- */
int mc2_update_centers(mc2_ctx *ctx, const mc2_model *model, const mc2_hset *centers, uint64_t n_centers, const mc2_hset *set_m,
		       const uint64_t *member_off, const uint64_t *members, double id, int64_t *next, uint64_t *n_good)
{
	if (model->m.bias < -0.5 || model->regression) return fail(MC2_ERR_UNSUPPORTED, "stub: needs a classifier with bias >= -0.5");
	for (uint64_t c = 0; c < n_centers; c++) {
		const uint64_t m0 = member_off[c], nm = member_off[c + 1] - m0;
		next[c] = -1;
		if (n_good) n_good[c] = 0;
		if (nm == 0) continue;
		std::vector<uint8_t> keep(nm);
		int rc = mc2_filter_as(ctx, model, centers, c, centers->mag[c], centers->len[c], set_m, members + m0, nm, id, keep.data());
		if (rc != MC2_OK) return rc;
		std::vector<uint64_t> good, pos;
		for (uint64_t j = 0; j < nm; j++) {
			if (keep[j]) {
				good.push_back(members[m0 + j]);
				pos.push_back(j);
			}
		}
		if (n_good) n_good[c] = good.size();
		if (good.empty()) continue;
		int64_t best = -1;
		double bd = 0;
		if (mc2o_mean_closest(set_m->eb, set_m->N, set_m->bins.data(), good.data(), good.size(), &best, &bd, nullptr, nullptr) != 0) {
			return fail(MC2_ERR_FEATURE, "stub: mc2o_mean_closest failed");
		}
		next[c] = (int64_t)pos[(size_t)best];
	}
	return MC2_OK;
}

int mc2_merge_centers(mc2_ctx *ctx, const mc2_model *model, const mc2_hset *centers, uint64_t n_centers, int64_t delta, double id,
		      int64_t *out)
{
	std::vector<uint64_t> rows(n_centers);
	for (uint64_t c = 0; c < n_centers; c++) rows[c] = c;
	for (uint64_t c = 0; c < n_centers; c++) {
		const int64_t last = (int64_t)((uint64_t)delta < n_centers - 1 - c ? c + (uint64_t)delta : n_centers - 1);
		const int rc = mc2_merge(ctx, model, centers, rows.data(), (int64_t)c, (int64_t)c + 1, last, id, &out[c]);
		if (rc != MC2_OK) return rc;
	}
	return MC2_OK;
}

} // extern "C"
