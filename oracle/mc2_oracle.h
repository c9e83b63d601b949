/* mc2_oracle.h — CPU restatement of MeShClust2's hot path, in plain C.
 *
 * TEST INFRASTRUCTURE ONLY.  This is the parity checker for the CUDA path; it is never linked
 * into, imported by or called from the product (libmeshclust2_b200.so / meshclust2_b200/).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it.
 *
 * Parity status: PINNED.  The reference ships no tests or golden vectors, so this
 * restatement is pinned against the reference itself compiled here (oracle/_ref/libmc2ref.so, built by
 * oracle/Makefile from the sources under /root/reference) — tests/test_oracle_vs_ref.py — and against
 * the committed fixtures in tests/golden/ that were generated from that build
 * (tests/golden/make_golden.py).
 *
 * Every function cites the reference file:line it follows (paths relative to /root/reference).
 */
#ifndef MC2_ORACLE_H
#define MC2_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* single-feature flags, src/predict/Feature.h:31-64 (only the in-scope ones) */
#define MC2O_FEAT_MANHATTAN           (1ULL << 2)
#define MC2O_FEAT_EUCLIDEAN           (1ULL << 3)
#define MC2O_FEAT_NORMALIZED_VECTORS  (1ULL << 5)
#define MC2O_FEAT_JEFFEREY_DIV        (1ULL << 7)
#define MC2O_FEAT_PEARSON_COEFF       (1ULL << 9)
#define MC2O_FEAT_INTERSECTION        (1ULL << 13)
#define MC2O_FEAT_EMD                 (1ULL << 18)
#define MC2O_FEAT_LENGTHD             (1ULL << 21)
#define MC2O_FEAT_KULCZYNSKI2         (1ULL << 27)
#define MC2O_FEAT_SIMRATIO            (1ULL << 28)
#define MC2O_FEAT_JENSEN_SHANNON      (1ULL << 29)

/* combo codes as written in weights.txt, src/predict/Predictor.cpp:96-110 */
#define MC2O_COMBO_XY   0
#define MC2O_COMBO_XY2  1
#define MC2O_COMBO_X2Y  2
#define MC2O_COMBO_X2Y2 3

#define MC2O_MAX_SINGLES 16
#define MC2O_MAX_COMBOS  16
#define MC2O_MAX_COMBO_IDX 4

typedef struct {
	int n_singles;
	uint64_t single_flag[MC2O_MAX_SINGLES];
	double single_min[MC2O_MAX_SINGLES];
	double single_max[MC2O_MAX_SINGLES];
	int n_combos;
	int combo_kind[MC2O_MAX_COMBOS];                 /* MC2O_COMBO_* */
	int combo_nidx[MC2O_MAX_COMBOS];
	int combo_idx[MC2O_MAX_COMBOS][MC2O_MAX_COMBO_IDX]; /* indices into singles, ascending flag bit */
	double weight[MC2O_MAX_COMBOS + 1];              /* weight[0] = intercept */
	double bias;                                     /* Predictor::set_bias, default 0 */
} mc2o_model;

/* one histogram + side-band, the DivergencePoint<T> state the features read */
typedef struct {
	const void *bins;   /* N elements of elem_bytes */
	uint64_t mag;       /* getPseudoMagnitude(): host-supplied, may be stale (quirk Q4) */
	uint64_t len;       /* get_length() */
} mc2o_point;

/* a1: Chromosome::help + ChromosomeOneDigit::encode. Returns 0; -1 invalid letter; -3 too many segments. */
int mc2o_encode(const char *text, long len, char *base_out, int *segs_out, int max_segs, int *nseg, long *eff_size);

/* a2/a3/a4: KmerHashTable + Loader::fill_table over (codes, segments); init 1, saturating.
 * hist: 4^k elements of elem_bytes; mers1: 4 x u64 (k=1 table, init 1);
 * n_overflow_segs: number of segments whose wholesaleIncrementNoOverflow returned -1. */
int mc2o_count(const char *codes, const int *segs, int nseg, int k, int elem_bytes, void *hist, uint64_t *mers1,
	       int *n_overflow_segs);
/* f2: Runner::run's width detection (CRunner.cpp:57-93, ClusterFactory.h:40-54): 1 + largest k-mer multiplicity;
 * -1 where the reference would read past a segment shorter than k (quirk Q6). mc2o_width_for: CRunner.cpp:108-126. */
int mc2o_largest_count(const char *codes, const int *segs, int nseg, int k, uint64_t *largest);
int mc2o_width_for(uint64_t largest_count);
/* a4: mag = sum(bins); stddev as in Loader::get_point */
void mc2o_point_stats(const void *hist, uint64_t N, int elem_bytes, uint64_t *mag, double *stddev);
/* Loader<T>::get_point(header, ACGT-string) front half: strip everything but A,C,G,T (Loader.cpp:112-134) */
long mc2o_strip_acgt(const char *text, long len, char *out);

/* a7/a8: one raw single. Returns 0, or -1 when the reference would throw (length 0). */
int mc2o_raw_single(uint64_t flag, int elem_bytes, uint64_t N, const mc2o_point *p, const mc2o_point *q, double *out);

/* a6/a10/a11: compute() -> normalised cache[S]; combos; sum; score = logistic(sum)+bias; close = round(score)>0.
 * Returns 0, -1 if a raw single throws or a normalised value is NaN (reference throws). */
int mc2o_score_pair(const mc2o_model *m, int elem_bytes, uint64_t N, const mc2o_point *a, const mc2o_point *b,
		    double *cache, double *dist, double *sum, double *score, int *close);
/* regression form, Predictor::p_predict: sum clamped to [0,1] */
int mc2o_predict_pair(const mc2o_model *m, int elem_bytes, uint64_t N, const mc2o_point *a, const mc2o_point *b,
		      double *sim);

/* a12: batched callers over rows of a histogram matrix H[n x N] with side-band mag[n], len[n]. */
int mc2o_get_close(const mc2o_model *m, int elem_bytes, uint64_t N, const void *H, const uint64_t *mag,
		   const uint64_t *len, uint64_t q, uint64_t ncand, const uint64_t *cand, double cutoff,
		   int64_t *best, double *best_dist, int *is_min, uint8_t *marks);
int mc2o_filter(const mc2o_model *m, int elem_bytes, uint64_t N, const void *H, const uint64_t *mag,
		const uint64_t *len, uint64_t c, uint64_t nmem, const uint64_t *members, double id, uint8_t *keep);
int mc2o_merge(const mc2o_model *m, int elem_bytes, uint64_t N, const void *H, const uint64_t *mag,
	       const uint64_t *len, const uint64_t *rows, long cur, long begin, long last, double id, long *out);

/* a13: DivergencePoint::distance / distance_d */
uint64_t mc2o_distance(int elem_bytes, uint64_t N, const mc2o_point *p, const mc2o_point *q);
double mc2o_distance_d(int elem_bytes, uint64_t N, const void *bins, const double *center);

/* K3: mean of member histograms (double) + first member minimising distance_d to it
 * (get_mean, src/cluster/ClusterFactory.cpp:338-380; Trainer::closest, src/cluster/Trainer.cpp:144-157) */
int mc2o_mean_closest(int elem_bytes, uint64_t N, const void *H, const uint64_t *members, uint64_t n, int64_t *best,
		      double *best_dist, double *mean_out, double *dist_out);

/* timing helpers for bench.py's cpu_baseline "port" leg (OpenMP over units) */
int mc2o_count_batch(const char *codes, const uint64_t *seq_off, const int *segs, const uint64_t *seg_off,
		     uint64_t n, int k, int elem_bytes, void *hist, int threads, double *seconds);
int mc2o_score_pairs(const mc2o_model *m, int elem_bytes, uint64_t N, const void *H, const uint64_t *mag,
		     const uint64_t *len, uint64_t npairs, const uint64_t *ia, const uint64_t *ib, double *score,
		     double *dist, uint8_t *close, double *cache, int threads, double *seconds);

#ifdef __cplusplus
}
#endif
#endif
