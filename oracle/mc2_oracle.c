/* mc2_oracle.c — CPU restatement of MeShClust2's hot path (k-mer histograms + pair features + GLM).
 *
 * TEST INFRASTRUCTURE ONLY — see mc2_oracle.h.  Parity status: PINNED against the reference
 * compiled here (oracle/_ref) and the committed fixtures under tests/golden/.
 *
 * Written from the reference's behaviour, function by function; each block cites the
 * reference file:line (relative to /root/reference) it restates.  The type-dependent integer
 * promotion of the reference's templates (SURVEY.md notes E1-E4) is reproduced by instantiating
 * the same expressions for uint8_t/uint16_t/uint32_t/uint64_t through macros: C and C++ share
 * the usual arithmetic conversions, so `p[i] - q[i]` has the same type and wrap-around here.
 * Built with -ffp-contract=off: the arithmetic below is the reference's expression order with
 * no fused multiply-adds.
 */
#include "mc2_oracle.h"
#include <ctype.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------------------------
 * a1  Input contract: Chromosome::help (src/nonltr/Chromosome.cpp:130-154)
 * ------------------------------------------------------------------------------------------ */

/* code map, src/nonltr/ChromosomeOneDigitDna.cpp:48-68 */
static int dna_code(char c)
{
	switch (c) {
	case 'A': return 0;
	case 'C': return 1;
	case 'G': return 2;
	case 'T': return 3;
	case 'R': return 2;
	case 'Y': return 1;
	case 'M': return 0;
	case 'K': return 3;
	case 'S': return 2;
	case 'W': return 3;
	case 'H': return 1;
	case 'B': return 3;
	case 'V': return 0;
	case 'D': return 3;
	case 'N': return 1;
	case 'X': return 2;
	default: return -1;
	}
}

typedef struct {
	int *v; /* pairs */
	int n, cap;
} seglist;

static void seg_push(seglist *l, int s, int e)
{
	if (l->n == l->cap) {
		l->cap = l->cap ? 2 * l->cap : 16;
		l->v = (int *)realloc(l->v, sizeof(int) * 2 * (size_t)l->cap);
	}
	l->v[2 * l->n] = s;
	l->v[2 * l->n + 1] = e;
	l->n++;
}

int mc2o_encode(const char *text, long len, char *base, int *segs_out, int max_segs, int *nseg, long *eff_size)
{
	seglist raw = {0, 0, 0}, merged = {0, 0, 0}, fin = {0, 0, 0};
	long i;
	int rc = 0;
	/* toUpperCase, Chromosome.cpp:254-258 */
	for (i = 0; i < len; i++) {
		base[i] = (char)toupper((unsigned char)text[i]);
	}
	/* removeAmbiguous, Chromosome.cpp:263-291 — note the else-if order: a non-N base that opens a run at the
	 * very last index never closes it (quirk Q7). */
	{
		int start = -1;
		for (i = 0; i < len; i++) {
			if (base[i] != 'N' && start == -1) {
				start = (int)i;
			} else if (base[i] == 'N' && start != -1) {
				seg_push(&raw, start, (int)i - 1);
				start = -1;
			} else if (i == len - 1 && base[i] != 'N' && start != -1) {
				seg_push(&raw, start, (int)i);
				start = -1;
			}
		}
	}
	/* mergeSegments, Chromosome.cpp:298-353: only when base.size() > 20 (Chromosome.cpp:147) */
	if (len > 20) {
		if (raw.n > 0) {
			int s = raw.v[0], e = raw.v[1], j;
			for (j = 1; j < raw.n; j++) {
				int s1 = raw.v[2 * j], e1 = raw.v[2 * j + 1];
				if (s1 - e < 10) {
					e = e1;
				} else {
					if (e - s + 1 >= 20) {
						seg_push(&merged, s, e);
					}
					s = s1;
					e = e1;
				}
			}
			if (e - s + 1 >= 20) {
				seg_push(&merged, s, e);
			}
		}
	} else {
		int j;
		for (j = 0; j < raw.n; j++) {
			seg_push(&merged, raw.v[2 * j], raw.v[2 * j + 1]);
		}
	}
	/* makeSegmentList, Chromosome.cpp:355-385, segLength = 1000000 */
	{
		const int segLength = 1000000;
		int j;
		for (j = 0; j < merged.n; j++) {
			int s = merged.v[2 * j], e = merged.v[2 * j + 1];
			if (e - s + 1 > segLength) {
				int fragNum = (e - s + 1) / segLength, h;
				for (h = 0; h < fragNum; h++) {
					int fragStart = s + h * segLength;
					int fragEnd = (h == fragNum - 1) ? e : fragStart + segLength - 1;
					seg_push(&fin, fragStart, fragEnd);
				}
			} else {
				seg_push(&fin, s, e);
			}
		}
	}
	/* ChromosomeOneDigit::encode, src/nonltr/ChromosomeOneDigit.cpp:79-133 */
	{
		int j;
		for (j = 0; j < fin.n && rc == 0; j++) {
			for (i = fin.v[2 * j]; i <= fin.v[2 * j + 1]; i++) {
				int c = dna_code(base[i]);
				if (c < 0) {
					rc = -1;
					break;
				}
				base[i] = (char)c;
			}
		}
		if (fin.n > 0 && rc == 0) {
			long gs = 0, ge = fin.v[0] - 1;
			for (j = 0; j <= fin.n && rc == 0; j++) {
				for (i = gs; i <= ge; i++) {
					char c = base[i];
					if (c != 'N') {
						int d = dna_code(c);
						if (d < 0) {
							rc = -1;
							break;
						}
						base[i] = (char)d;
					}
				}
				if (j < fin.n - 1) {
					gs = fin.v[2 * j + 1] + 1;
					ge = fin.v[2 * (j + 1)] - 1;
				} else if (j == fin.n - 1) {
					gs = fin.v[2 * j + 1] + 1;
					ge = len - 1;
				}
			}
		}
	}
	if (rc == 0) {
		int j;
		long eff = 0;
		if (fin.n > max_segs) {
			rc = -3;
		} else {
			for (j = 0; j < fin.n; j++) {
				segs_out[2 * j] = fin.v[2 * j];
				segs_out[2 * j + 1] = fin.v[2 * j + 1];
				eff += fin.v[2 * j + 1] - fin.v[2 * j] + 1; /* calculateEffectiveSize, Chromosome.cpp:420-427 */
			}
			*nseg = fin.n;
			*eff_size = eff;
		}
	}
	free(raw.v);
	free(merged.v);
	free(fin.v);
	return rc;
}

/* Loader<T>::get_point(header, string), src/clutil/Loader.cpp:114-120 */
long mc2o_strip_acgt(const char *text, long len, char *out)
{
	long i, n = 0;
	for (i = 0; i < len; i++) {
		char c = text[i];
		if (c == 'A' || c == 'C' || c == 'G' || c == 'T') {
			out[n++] = c;
		}
	}
	return n;
}

/* ------------------------------------------------------------------------------------------
 * a2/a3  KmerHashTable::wholesaleIncrementNoOverflow (src/nonltr/KmerHashTable.cpp:236-256) driven by
 *        Loader::fill_table (src/clutil/Loader.cpp:42-86)
 * ------------------------------------------------------------------------------------------ */
#define DEF_COUNT(T, SUF, TMAX)                                                                           \
	static int count_##SUF(const char *codes, const int *segs, int nseg, int k, T *values, int *novf)       \
	{                                                                                                       \
		const uint64_t N = 1ULL << (2 * k);                                                             \
		uint64_t i;                                                                                     \
		int s;                                                                                          \
		for (i = 0; i < N; i++) {                                                                       \
			values[i] = 1; /* KmerHashTable(k, 1), Loader.cpp:141 */                                  \
		}                                                                                               \
		*novf = 0;                                                                                      \
		for (s = 0; s < nseg; s++) {                                                                    \
			int start = segs[2 * s], end = segs[2 * s + 1];                                           \
			if (end - start + 1 >= k) { /* Loader.cpp:53 */                                           \
				int first = start, last = end - k + 1, ret = 0, p;                                 \
				/* hash(), KmerHashTable.cpp:108-131: first base most significant */               \
				uint64_t h = 0;                                                                    \
				for (p = first; p <= last; p++) {                                                  \
					if (!(codes[p] >= 0 && codes[p] <= 3)) {                                   \
						return -1; /* InvalidInputException, KmerHashTable.cpp:138-149 */    \
					}                                                                          \
				}                                                                                  \
				for (p = 0; p < k; p++) {                                                          \
					if (!(codes[first + p] >= 0 && codes[first + p] <= 3)) {                   \
						return -1;                                                           \
					}                                                                          \
					h = h * 4 + (uint64_t)codes[first + p];                                    \
				}                                                                                  \
				for (p = first;; p++) {                                                            \
					if (values[h] < (T)(TMAX)) {                                               \
						values[h]++;                                                         \
					} else {                                                                   \
						ret = -1;                                                            \
					}                                                                          \
					if (p == last) {                                                           \
						break;                                                               \
					}                                                                          \
					/* rolling update, KmerHashTable.cpp:154-158 (no range check on the incoming base; \
					 * an out-of-range code would index outside the table in the reference) */ \
					h = 4 * (h - (uint64_t)codes[p] * (N >> 2)) + (uint64_t)(int)codes[p + k]; \
					if (h >= N) {                                                              \
						return -1; /* "array out of bounds" throw, KmerHashTable.cpp:245-248 */ \
					}                                                                          \
				}                                                                                  \
				if (ret == -1) {                                                                   \
					(*novf)++; /* num_overflow++, Loader.cpp:55-56 */                          \
				}                                                                                  \
			}                                                                                         \
		}                                                                                               \
		return 0;                                                                                       \
	}

DEF_COUNT(uint8_t, u8, UINT8_MAX)
DEF_COUNT(uint16_t, u16, UINT16_MAX)
DEF_COUNT(uint32_t, u32, UINT32_MAX)
DEF_COUNT(uint64_t, u64, UINT64_MAX)

int mc2o_count(const char *codes, const int *segs, int nseg, int k, int elem_bytes, void *hist, uint64_t *mers1,
	       int *n_overflow_segs)
{
	int rc, dummy = 0;
	switch (elem_bytes) {
	case 1: rc = count_u8(codes, segs, nseg, k, (uint8_t *)hist, n_overflow_segs); break;
	case 2: rc = count_u16(codes, segs, nseg, k, (uint16_t *)hist, n_overflow_segs); break;
	case 4: rc = count_u32(codes, segs, nseg, k, (uint32_t *)hist, n_overflow_segs); break;
	case 8: rc = count_u64(codes, segs, nseg, k, (uint64_t *)hist, n_overflow_segs); break;
	default: return -2;
	}
	if (rc != 0) {
		return rc;
	}
	/* 1-mer table: KmerHashTable<unsigned long,uint64_t>(1,1) filled the same way, Loader.cpp:144,150 */
	if (mers1) {
		rc = count_u64(codes, segs, nseg, 1, mers1, &dummy);
	}
	return rc;
}

/* Histogram-width detection, Runner::run (src/cluster/CRunner.cpp:57-93): a u64 table (init 1) per sequence filled by
 * the free fill_table<V> (src/cluster/ClusterFactory.h:40-54) -> wholesaleIncrement(start, end-k+1) for EVERY segment
 * (quirk Q6: no `length >= k` guard), then max_element.  u64 never saturates here, so the result is
 * 1 + the largest k-mer multiplicity of the sequence.  A segment shorter than k makes the reference hash k characters
 * starting at `start`, i.e. read past the segment (and possibly the string); that is not restated: returns -1. */
int mc2o_largest_count(const char *codes, const int *segs, int nseg, int k, uint64_t *largest)
{
	const uint64_t N = 1ULL << (2 * k);
	uint64_t *values = (uint64_t *)malloc(N * sizeof(uint64_t));
	uint64_t i, best = 0;
	int s, novf = 0, rc;
	if (!values) {
		return -2;
	}
	for (s = 0; s < nseg; s++) {
		if (segs[2 * s + 1] - segs[2 * s] + 1 < k) {
			free(values);
			return -1;
		}
	}
	rc = count_u64(codes, segs, nseg, k, values, &novf); /* same increments: no guard needed, nothing saturates */
	if (rc == 0) {
		for (i = 0; i < N; i++) { /* std::max_element, CRunner.cpp:74 */
			if (values[i] > best) {
				best = values[i];
			}
		}
		*largest = best;
	}
	free(values);
	return rc;
}

/* Width choice from the largest count, CRunner.cpp:108-126: smallest of 8/16/32/64 bits whose max holds it. */
int mc2o_width_for(uint64_t largest_count)
{
	if (largest_count <= UINT8_MAX) {
		return 1;
	}
	if (largest_count <= UINT16_MAX) {
		return 2;
	}
	if (largest_count <= UINT32_MAX) {
		return 4;
	}
	return 8;
}

/* mag: DivergencePoint ctor, src/clutil/DivergencePoint.cpp:99-110; stddev: Loader.cpp:162-171 */
void mc2o_point_stats(const void *hist, uint64_t N, int elem_bytes, uint64_t *mag, double *stddev)
{
	uint64_t i, m = 0;
	double aq, sq = 0;
#define GET(i)                                                                                       \
	(elem_bytes == 1 ? (double)((const uint8_t *)hist)[i]                                        \
			 : elem_bytes == 2 ? (double)((const uint16_t *)hist)[i]                     \
					   : elem_bytes == 4 ? (double)((const uint32_t *)hist)[i]   \
							     : (double)((const uint64_t *)hist)[i])
	for (i = 0; i < N; i++) {
		switch (elem_bytes) {
		case 1: m += ((const uint8_t *)hist)[i]; break;
		case 2: m += ((const uint16_t *)hist)[i]; break;
		case 4: m += ((const uint32_t *)hist)[i]; break;
		default: m += ((const uint64_t *)hist)[i]; break;
		}
	}
	aq = (double)m / N;
	for (i = 0; i < N; i++) {
		double qdiff = GET(i) - aq;
		sq += qdiff * qdiff;
	}
#undef GET
	*mag = m;
	*stddev = sqrt(sq / N);
}

/* ------------------------------------------------------------------------------------------
 * a7/a8  raw singles, src/predict/Feature.cpp
 * ------------------------------------------------------------------------------------------ */
#define MINV(a, b) ((a) < (b) ? (a) : (b))

#define DEF_SINGLES(T, SUF)                                                                                  \
	/* manhattan, Feature.cpp:858-871: accumulates into `int` (E3) */                                      \
	static double manhattan_##SUF(const T *p, const T *q, uint64_t N)                                      \
	{                                                                                                      \
		int sum = 0;                                                                                   \
		uint64_t i;                                                                                    \
		for (i = 0; i < N; i++) {                                                                      \
			sum += p[i] > q[i] ? p[i] - q[i] : q[i] - p[i];                                          \
		}                                                                                              \
		return sum;                                                                                    \
	}                                                                                                      \
	/* euclidean, Feature.cpp:1112-1124: `auto diff = p-q` keeps T's promoted type (E1/E2) */              \
	static double euclidean_##SUF(const T *p, const T *q, uint64_t N)                                      \
	{                                                                                                      \
		uintmax_t sum = 0;                                                                             \
		uint64_t i;                                                                                    \
		for (i = 0; i < N; i++) {                                                                      \
			sum += (p[i] - q[i]) * (p[i] - q[i]);                                                    \
		}                                                                                              \
		return sqrt((double)sum);                                                                      \
	}                                                                                                      \
	/* normalized_vectors, Feature.cpp:1170-1184: u64 product d1*d2 before the sqrt (E4) */                \
	static double normalized_vectors_##SUF(const T *p, const T *q, uint64_t N)                             \
	{                                                                                                      \
		uintmax_t sum = 0, d1 = 0, d2 = 0;                                                             \
		uint64_t i;                                                                                    \
		for (i = 0; i < N; i++) {                                                                      \
			sum += p[i] * q[i];                                                                      \
			d1 += p[i] * p[i];                                                                       \
			d2 += q[i] * q[i];                                                                       \
		}                                                                                              \
		return (double)sum / sqrt((double)(d1 * d2));                                                  \
	}                                                                                                      \
	/* jefferey_divergence, Feature.cpp:1230-1263 */                                                       \
	static double jefferey_##SUF(const T *p, const T *q, uint64_t N, uint64_t mp, uint64_t mq)             \
	{                                                                                                      \
		double sum = 0;                                                                                \
		uint64_t i;                                                                                    \
		for (i = 0; i < N; i++) {                                                                      \
			double pp = (double)p[i] / mp;                                                           \
			double pq = (double)q[i] / mq;                                                           \
			double diff = pp - pq;                                                                   \
			sum += diff * log(pp / pq);                                                              \
		}                                                                                              \
		return sum;                                                                                    \
	}                                                                                                      \
	/* pearson, Feature.cpp:794-811 */                                                                     \
	static double pearson_##SUF(const T *p, const T *q, uint64_t N, uint64_t mp, uint64_t mq)              \
	{                                                                                                      \
		double dap = (double)mp / N;                                                                   \
		double daq = (double)mq / N;                                                                   \
		double dot = 0, np = 0, nq = 0;                                                                \
		uint64_t i;                                                                                    \
		for (i = 0; i < N; i++) {                                                                      \
			double dp = p[i] - dap;                                                                  \
			double dq = q[i] - daq;                                                                  \
			np += dp * dp;                                                                           \
			nq += dq * dq;                                                                           \
			dot += dp * dq;                                                                          \
		}                                                                                              \
		return dot / sqrt(np * nq);                                                                    \
	}                                                                                                      \
	/* intersection, Feature.cpp:763-777 */                                                                \
	static double intersection_##SUF(const T *p, const T *q, uint64_t N, uint64_t mp, uint64_t mq)         \
	{                                                                                                      \
		uintmax_t dist = 0;                                                                            \
		uintmax_t mag = mp + mq;                                                                       \
		uint64_t i;                                                                                    \
		for (i = 0; i < N; i++) {                                                                      \
			dist += 2 * MINV(p[i], q[i]);                                                            \
		}                                                                                              \
		return (double)dist / (double)mag;                                                             \
	}                                                                                                      \
	/* emd, Feature.cpp:1504-1518 */                                                                       \
	static double emd_##SUF(const T *p, const T *q, uint64_t N)                                            \
	{                                                                                                      \
		uintmax_t cp = 0, cq = 0, dist = 0;                                                            \
		uint64_t i;                                                                                    \
		for (i = 0; i < N; i++) {                                                                      \
			cp += p[i];                                                                              \
			cq += q[i];                                                                              \
			dist += cp > cq ? cp - cq : cq - cp;                                                     \
		}                                                                                              \
		return (double)dist;                                                                           \
	}                                                                                                      \
	/* kulczynski2, Feature.cpp:681-695 */                                                                 \
	static double kulczynski2_##SUF(const T *p, const T *q, uint64_t N, uint64_t mp, uint64_t mq)          \
	{                                                                                                      \
		uint64_t min_sum = 0, i;                                                                       \
		double ap = (double)mp / N;                                                                    \
		double aq = (double)mq / N;                                                                    \
		double coeff;                                                                                  \
		for (i = 0; i < N; i++) {                                                                      \
			min_sum += MINV(p[i], q[i]);                                                             \
		}                                                                                              \
		coeff = N * (ap + aq) / (2 * ap * aq);                                                         \
		return coeff * min_sum;                                                                        \
	}                                                                                                      \
	/* simratio, Feature.cpp:828-841: diff converted to intmax_t AFTER the T-typed subtraction (E2) */     \
	static double simratio_##SUF(const T *p, const T *q, uint64_t N)                                       \
	{                                                                                                      \
		uintmax_t dot = 0, norm2 = 0;                                                                  \
		uint64_t i;                                                                                    \
		for (i = 0; i < N; i++) {                                                                      \
			intmax_t diff = p[i] - q[i];                                                             \
			dot += p[i] * q[i];                                                                      \
			norm2 += diff * diff;                                                                    \
		}                                                                                              \
		return dot / (dot + sqrt((double)norm2));                                                      \
	}                                                                                                      \
	/* jensen_shannon, Feature.cpp:983-1009 (USETBL undefined) */                                          \
	static double jensen_shannon_##SUF(const T *p, const T *q, uint64_t N, uint64_t mp, uint64_t mq)       \
	{                                                                                                      \
		double sum = 0;                                                                                \
		uint64_t i;                                                                                    \
		for (i = 0; i < N; i++) {                                                                      \
			double pp = (double)p[i] / mp;                                                           \
			double pq = (double)q[i] / mq;                                                           \
			double avg = 0.5 * (pp + pq);                                                            \
			double lp = log(pp / avg);                                                               \
			double lq = log(pq / avg);                                                               \
			sum += pp * lp + pq * lq;                                                                \
		}                                                                                              \
		return sum / 2;                                                                                \
	}                                                                                                      \
	/* DivergencePoint::distance, src/clutil/DivergencePoint.cpp:70-82 */                                  \
	static uint64_t distance_##SUF(const T *p, const T *q, uint64_t N, uint64_t mp, uint64_t mq)           \
	{                                                                                                      \
		uint64_t dist = 0, i;                                                                          \
		const uint64_t mag = mp + mq;                                                                  \
		double frac;                                                                                   \
		for (i = 0; i < N; i++) {                                                                      \
			dist += MINV(p[i], q[i]);                                                                \
		}                                                                                              \
		dist *= 2;                                                                                     \
		frac = (double)dist / mag;                                                                     \
		return (uint64_t)(10000.0 * (1.0 - frac * frac));                                              \
	}                                                                                                      \
	/* DivergencePoint::distance_d, src/clutil/DivergencePoint.cpp:55-66:                                  \
	 * `mag += points[i] + c.points[i]` is u64 = (double)(u64 + (T + double)) truncated each step */       \
	static double distance_d_##SUF(const T *p, const double *c, uint64_t N)                                \
	{                                                                                                      \
		uint64_t dist = 0, mag = 0, i;                                                                 \
		double frac;                                                                                   \
		for (i = 0; i < N; i++) {                                                                      \
			T r = (T)round(c[i]);                                                                    \
			dist += 2 * MINV(p[i], r);                                                               \
			mag += p[i] + c[i];                                                                      \
		}                                                                                              \
		frac = (double)dist / mag;                                                                     \
		return 10000.0 * (1.0 - frac * frac);                                                          \
	}

DEF_SINGLES(uint8_t, u8)
DEF_SINGLES(uint16_t, u16)
DEF_SINGLES(uint32_t, u32)
DEF_SINGLES(uint64_t, u64)

int mc2o_raw_single(uint64_t flag, int eb, uint64_t N, const mc2o_point *p, const mc2o_point *q, double *out)
{
	const void *P = p->bins, *Q = q->bins;
	switch (flag) {
	case MC2O_FEAT_MANHATTAN:
		*out = eb == 1 ? manhattan_u8(P, Q, N) : eb == 2 ? manhattan_u16(P, Q, N) : eb == 4 ? manhattan_u32(P, Q, N) : manhattan_u64(P, Q, N);
		return 0;
	case MC2O_FEAT_EUCLIDEAN:
		*out = eb == 1 ? euclidean_u8(P, Q, N) : eb == 2 ? euclidean_u16(P, Q, N) : eb == 4 ? euclidean_u32(P, Q, N) : euclidean_u64(P, Q, N);
		return 0;
	case MC2O_FEAT_NORMALIZED_VECTORS:
		*out = eb == 1 ? normalized_vectors_u8(P, Q, N) : eb == 2 ? normalized_vectors_u16(P, Q, N) : eb == 4 ? normalized_vectors_u32(P, Q, N) : normalized_vectors_u64(P, Q, N);
		return 0;
	case MC2O_FEAT_JEFFEREY_DIV:
		*out = eb == 1 ? jefferey_u8(P, Q, N, p->mag, q->mag) : eb == 2 ? jefferey_u16(P, Q, N, p->mag, q->mag) : eb == 4 ? jefferey_u32(P, Q, N, p->mag, q->mag) : jefferey_u64(P, Q, N, p->mag, q->mag);
		return 0;
	case MC2O_FEAT_PEARSON_COEFF:
		*out = eb == 1 ? pearson_u8(P, Q, N, p->mag, q->mag) : eb == 2 ? pearson_u16(P, Q, N, p->mag, q->mag) : eb == 4 ? pearson_u32(P, Q, N, p->mag, q->mag) : pearson_u64(P, Q, N, p->mag, q->mag);
		return 0;
	case MC2O_FEAT_INTERSECTION:
		*out = eb == 1 ? intersection_u8(P, Q, N, p->mag, q->mag) : eb == 2 ? intersection_u16(P, Q, N, p->mag, q->mag) : eb == 4 ? intersection_u32(P, Q, N, p->mag, q->mag) : intersection_u64(P, Q, N, p->mag, q->mag);
		return 0;
	case MC2O_FEAT_EMD:
		*out = eb == 1 ? emd_u8(P, Q, N) : eb == 2 ? emd_u16(P, Q, N) : eb == 4 ? emd_u32(P, Q, N) : emd_u64(P, Q, N);
		return 0;
	case MC2O_FEAT_LENGTHD: {
		/* length_difference, Feature.cpp:873-887: throws 123 when a length is 0 */
		unsigned long lp = p->len, lq = q->len;
		if (lp == 0 || lq == 0) {
			return -1;
		}
		*out = (double)((lp > lq) ? (lp - lq) : (lq - lp));
		return 0;
	}
	case MC2O_FEAT_KULCZYNSKI2:
		*out = eb == 1 ? kulczynski2_u8(P, Q, N, p->mag, q->mag) : eb == 2 ? kulczynski2_u16(P, Q, N, p->mag, q->mag) : eb == 4 ? kulczynski2_u32(P, Q, N, p->mag, q->mag) : kulczynski2_u64(P, Q, N, p->mag, q->mag);
		return 0;
	case MC2O_FEAT_SIMRATIO:
		*out = eb == 1 ? simratio_u8(P, Q, N) : eb == 2 ? simratio_u16(P, Q, N) : eb == 4 ? simratio_u32(P, Q, N) : simratio_u64(P, Q, N);
		return 0;
	case MC2O_FEAT_JENSEN_SHANNON:
		*out = eb == 1 ? jensen_shannon_u8(P, Q, N, p->mag, q->mag) : eb == 2 ? jensen_shannon_u16(P, Q, N, p->mag, q->mag) : eb == 4 ? jensen_shannon_u32(P, Q, N, p->mag, q->mag) : jensen_shannon_u64(P, Q, N, p->mag, q->mag);
		return 0;
	default:
		return -2;
	}
}

uint64_t mc2o_distance(int eb, uint64_t N, const mc2o_point *p, const mc2o_point *q)
{
	const void *P = p->bins, *Q = q->bins;
	return eb == 1 ? distance_u8(P, Q, N, p->mag, q->mag) : eb == 2 ? distance_u16(P, Q, N, p->mag, q->mag) : eb == 4 ? distance_u32(P, Q, N, p->mag, q->mag) : distance_u64(P, Q, N, p->mag, q->mag);
}

double mc2o_distance_d(int eb, uint64_t N, const void *bins, const double *center)
{
	return eb == 1 ? distance_d_u8(bins, center, N) : eb == 2 ? distance_d_u16(bins, center, N) : eb == 4 ? distance_d_u32(bins, center, N) : distance_d_u64(bins, center, N);
}

/* ------------------------------------------------------------------------------------------
 * a6/a10/a11  Feature::compute, operator(), Trainer::classify, Predictor::classify_sum
 * ------------------------------------------------------------------------------------------ */

/* feat_is_sim, Feature.cpp:549-663 (in-scope singles) */
static int feat_is_sim(uint64_t flag)
{
	switch (flag) {
	case MC2O_FEAT_NORMALIZED_VECTORS:
	case MC2O_FEAT_PEARSON_COEFF:
	case MC2O_FEAT_INTERSECTION:
	case MC2O_FEAT_KULCZYNSKI2:
	case MC2O_FEAT_SIMRATIO:
		return 1;
	default:
		return 0;
	}
}

/* compute_all_raw + normalize_cache, Feature.cpp:136-171 */
static int compute_cache(const mc2o_model *m, int eb, uint64_t N, const mc2o_point *a, const mc2o_point *b, double *cache)
{
	int i;
	for (i = 0; i < m->n_singles; i++) {
		if (mc2o_raw_single(m->single_flag[i], eb, N, a, b, &cache[i]) != 0) {
			return -1;
		}
	}
	for (i = 0; i < m->n_singles; i++) {
		double val = (cache[i] - m->single_min[i]) / (m->single_max[i] - m->single_min[i]);
		if (isnan(val)) {
			return -1; /* throw std::exception(), Feature.cpp:143-146 */
		}
		cache[i] = feat_is_sim(m->single_flag[i]) ? val : 1 - val;
	}
	return 0;
}

/* Feature::operator()(col, cache), Feature.h:205-239 */
static int combo_value(const mc2o_model *m, int col, const double *cache, double *out)
{
	int j, n = m->combo_nidx[col];
	const int *idx = m->combo_idx[col];
	double prod = 1;
	switch (m->combo_kind[col]) {
	case MC2O_COMBO_XY:
		for (j = 0; j < n; j++) {
			prod *= cache[idx[j]];
		}
		*out = prod;
		return 0;
	case MC2O_COMBO_X2Y2:
		for (j = 0; j < n; j++) {
			prod *= cache[idx[j]] * cache[idx[j]];
		}
		*out = prod;
		return 0;
	case MC2O_COMBO_XY2:
		if (n != 2) {
			return -1;
		}
		*out = cache[idx[0]] * cache[idx[1]] * cache[idx[1]];
		return 0;
	case MC2O_COMBO_X2Y:
		if (n != 2) {
			return -1;
		}
		*out = cache[idx[0]] * cache[idx[0]] * cache[idx[1]];
		return 0;
	default:
		return -1;
	}
}

/* GLM::logistic, src/predict/GLM.cpp:26-29 */
static double logistic(double x)
{
	return 1.0 / (1 + exp(-x));
}

int mc2o_score_pair(const mc2o_model *m, int eb, uint64_t N, const mc2o_point *a, const mc2o_point *b, double *cache_out,
		    double *dist, double *sum_out, double *score, int *close)
{
	double cache[MC2O_MAX_SINGLES];
	double sum = m->weight[0], d0 = 0;
	int col;
	if (compute_cache(m, eb, N, a, b, cache) != 0) {
		return -1;
	}
	/* Trainer::classify, src/cluster/Trainer.cpp:112-120 */
	for (col = 0; col < m->n_combos; col++) {
		double d;
		if (combo_value(m, col, cache, &d) != 0) {
			return -1;
		}
		if (col == 0) {
			d0 = d;
		}
		sum += m->weight[col + 1] * d;
	}
	if (cache_out) {
		memcpy(cache_out, cache, sizeof(double) * (size_t)m->n_singles);
	}
	if (dist) {
		*dist = d0;
	}
	if (sum_out) {
		*sum_out = sum;
	}
	{
		/* Predictor::classify_sum, src/predict/Predictor.cpp:316-320 ; close <=> round(score) > 0 (:332) */
		double s = logistic(sum) + m->bias;
		if (score) {
			*score = s;
		}
		if (close) {
			*close = round(s) > 0;
		}
	}
	return 0;
}

/* Predictor::p_predict, src/predict/Predictor.cpp:284-300 */
int mc2o_predict_pair(const mc2o_model *m, int eb, uint64_t N, const mc2o_point *a, const mc2o_point *b, double *sim)
{
	double cache[MC2O_MAX_SINGLES];
	double sum = m->weight[0];
	int col;
	if (compute_cache(m, eb, N, a, b, cache) != 0) {
		return -1;
	}
	for (col = 0; col < m->n_combos; col++) {
		double d;
		if (combo_value(m, col, cache, &d) != 0) {
			return -1;
		}
		sum += m->weight[col + 1] * d;
	}
	if (sum < 0) {
		sum = 0;
	} else if (sum > 1) {
		sum = 1;
	}
	*sim = sum;
	return 0;
}

/* ------------------------------------------------------------------------------------------
 * a12  batched callers
 * ------------------------------------------------------------------------------------------ */
static mc2o_point row_point(int eb, uint64_t N, const void *H, const uint64_t *mag, const uint64_t *len, uint64_t i)
{
	mc2o_point p;
	p.bins = (const char *)H + i * N * (uint64_t)eb;
	p.mag = mag[i];
	p.len = len[i];
	return p;
}

/* Trainer::get_close, src/cluster/Trainer.cpp:23-71 (sequential order = --threads 1) */
int mc2o_get_close(const mc2o_model *m, int eb, uint64_t N, const void *H, const uint64_t *mag, const uint64_t *len,
		   uint64_t q, uint64_t ncand, const uint64_t *cand, double cutoff, int64_t *best, double *best_dist,
		   int *is_min_r, uint8_t *marks)
{
	mc2o_point p = row_point(eb, N, H, mag, len, q);
	int64_t b = -1;
	double bd = -1;
	int is_min = 1;
	uint64_t min_len = (uint64_t)(p.len * cutoff);
	uint64_t max_len = (uint64_t)(p.len / cutoff);
	uint64_t j;
	for (j = 0; j < ncand; j++) {
		mc2o_point pt = row_point(eb, N, H, mag, len, cand[j]);
		double dist, score;
		int close;
		marks[j] = 0;
		if (pt.len < min_len || pt.len > max_len) {
			continue;
		}
		if (mc2o_score_pair(m, eb, N, &pt, &p, NULL, &dist, NULL, &score, &close) != 0) { /* candidate first */
			return -1;
		}
		if (dist > bd) {
			bd = dist;
			b = (int64_t)j;
		}
		is_min = is_min && !close;
		if (close) {
			marks[j] = 1;
		}
	}
	*best = b;
	*best_dist = bd;
	*is_min_r = is_min;
	return 0;
}

/* Trainer::filter, src/cluster/Trainer.cpp:123-141 */
int mc2o_filter(const mc2o_model *m, int eb, uint64_t N, const void *H, const uint64_t *mag, const uint64_t *len, uint64_t c,
		uint64_t nmem, const uint64_t *members, double id, uint8_t *keep)
{
	mc2o_point p = row_point(eb, N, H, mag, len, c);
	uint64_t min_length = (uint64_t)(p.len * id);
	uint64_t max_length = (uint64_t)(p.len / id);
	uint64_t j;
	for (j = 0; j < nmem; j++) {
		mc2o_point pt = row_point(eb, N, H, mag, len, members[j]);
		int length_pass = pt.len >= min_length && pt.len <= max_length;
		keep[j] = 0; /* pt.second = true -> erased */
		if (length_pass) {
			double score;
			if (mc2o_score_pair(m, eb, N, &p, &pt, NULL, NULL, NULL, &score, NULL) != 0) { /* center first */
				return -1;
			}
			keep[j] = (round(score) != 0);
		}
	}
	return 0;
}

/* Trainer::merge, src/cluster/Trainer.cpp:74-109 (sequential order) */
int mc2o_merge(const mc2o_model *m, int eb, uint64_t N, const void *H, const uint64_t *mag, const uint64_t *len,
	       const uint64_t *rows, long cur, long begin, long last, double id, long *out)
{
	mc2o_point p = row_point(eb, N, H, mag, len, rows[cur]);
	long best = 0, i;
	double best_d = 2.2250738585072014e-308; /* std::numeric_limits<double>::min() */
	uint64_t min_length = (uint64_t)(p.len * id);
	uint64_t max_length = (uint64_t)(p.len / id);
	for (i = begin; i <= last; i++) {
		mc2o_point cen = row_point(eb, N, H, mag, len, rows[i]);
		int length_pass = cen.len >= min_length && cen.len <= max_length;
		if (length_pass) {
			double dist, score;
			if (mc2o_score_pair(m, eb, N, &cen, &p, NULL, &dist, NULL, &score, NULL) != 0) {
				return -1;
			}
			if (round(score) == 1) {
				if (!(best_d > dist)) {
					best = i;
					best_d = dist;
				}
			}
		}
	}
	*out = best;
	return 0;
}

/* ------------------------------------------------------------------------------------------
 * K3: cluster mean + closest member.  get_mean (src/cluster/ClusterFactory.cpp:338-380) and the mean of
 * mean_shift_update (:288-335) followed by Trainer::closest (src/cluster/Trainer.cpp:144-157):
 *   top = sum over members of the bins as doubles (operator+=), top /= count (operator/=, per-bin division),
 *   best = first member minimising distance_d(member, top) (strict <).
 * ------------------------------------------------------------------------------------------ */
int mc2o_mean_closest(int eb, uint64_t N, const void *H, const uint64_t *members, uint64_t n, int64_t *best, double *best_dist,
		      double *mean_out, double *dist_out)
{
	double *top = (double *)calloc(N, sizeof(double));
	uint64_t i, j;
	int64_t b = -1;
	double bd = 0;
	if (!top || n == 0) {
		free(top);
		return -1;
	}
	for (j = 0; j < n; j++) {
		const char *row = (const char *)H + members[j] * N * (uint64_t)eb;
		for (i = 0; i < N; i++) {
			double v = eb == 1 ? (double)((const uint8_t *)row)[i]
				 : eb == 2 ? (double)((const uint16_t *)row)[i]
				 : eb == 4 ? (double)((const uint32_t *)row)[i] : (double)((const uint64_t *)row)[i];
			top[i] += v;
		}
	}
	for (i = 0; i < N; i++) {
		top[i] /= (double)n;
	}
	for (j = 0; j < n; j++) {
		const char *row = (const char *)H + members[j] * N * (uint64_t)eb;
		double d = mc2o_distance_d(eb, N, row, top);
		if (dist_out) {
			dist_out[j] = d;
		}
		if (b < 0 || d < bd) {
			bd = d;
			b = (int64_t)j;
		}
	}
	if (mean_out) {
		memcpy(mean_out, top, N * sizeof(double));
	}
	*best = b;
	*best_dist = bd;
	free(top);
	return 0;
}

/* ------------------------------------------------------------------------------------------
 * timing helpers (cpu_baseline "port")
 * ------------------------------------------------------------------------------------------ */
static double now_s(void)
{
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

int mc2o_count_batch(const char *codes, const uint64_t *seq_off, const int *segs, const uint64_t *seg_off, uint64_t n, int k,
		     int eb, void *hist, int threads, double *seconds)
{
	const uint64_t N = 1ULL << (2 * k);
	int bad = 0;
	int64_t i;
	double t0 = now_s();
	(void)threads;
#pragma omp parallel for num_threads(threads) schedule(dynamic, 16)
	for (i = 0; i < (int64_t)n; i++) {
		uint64_t m1[4];
		int novf;
		int rc = mc2o_count(codes + seq_off[i], segs + 2 * seg_off[i], (int)(seg_off[i + 1] - seg_off[i]), k, eb,
				    (char *)hist + (uint64_t)i * N * (uint64_t)eb, m1, &novf);
		if (rc != 0) {
#pragma omp atomic write
			bad = 1;
		}
	}
	if (seconds) {
		*seconds = now_s() - t0;
	}
	return bad ? -1 : 0;
}

int mc2o_score_pairs(const mc2o_model *m, int eb, uint64_t N, const void *H, const uint64_t *mag, const uint64_t *len,
		     uint64_t npairs, const uint64_t *ia, const uint64_t *ib, double *score, double *dist, uint8_t *close,
		     double *cache, int threads, double *seconds)
{
	int bad = 0;
	int64_t j;
	double t0 = now_s();
	(void)threads;
#pragma omp parallel for num_threads(threads) schedule(static)
	for (j = 0; j < (int64_t)npairs; j++) {
		mc2o_point a = row_point(eb, N, H, mag, len, ia[j]);
		mc2o_point b = row_point(eb, N, H, mag, len, ib[j]);
		double s = 0, d = 0;
		int c = 0;
		int rc = mc2o_score_pair(m, eb, N, &a, &b, cache ? cache + (uint64_t)j * (uint64_t)m->n_singles : NULL, &d, NULL, &s, &c);
		if (rc != 0) {
#pragma omp atomic write
			bad = 1;
		}
		if (score) {
			score[j] = s;
		}
		if (dist) {
			dist[j] = d;
		}
		if (close) {
			close[j] = (uint8_t)c;
		}
	}
	if (seconds) {
		*seconds = now_s() - t0;
	}
	return bad ? -1 : 0;
}
