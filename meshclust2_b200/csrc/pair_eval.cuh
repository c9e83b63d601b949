// pair_eval.cuh — the per-pair fp64 epilogue shared by the pair kernels (pair_score.cu) and the tile sweep (tile_sweep.cu):
// side-band record, the exact integer reductions, and eval_pair_fast (singles -> normalise -> combos -> GLM -> logistic -> cutoff).
// Reference: Feature.cpp:136-171, Feature.h:205-239, Trainer.cpp:112-120, Predictor.cpp:284-333 (cited per block below).
#pragma once
#include "mc2_internal.cuh"
#include <math_constants.h>

namespace mc2 {

struct Side {
	u64 mag, sum, sumsq, len;
};
__device__ __forceinline__ Side load_side(const Sideband &sb, u64 row)
{
	Side s;
	s.mag = sb.mag[row];
	s.sum = sb.sum[row];
	s.sumsq = sb.sumsq[row];
	s.len = sb.len[row];
	return s;
}

// reductions for 8/16-bit histograms (exact integers) + optional log-feature sums
struct RedN {
	u64 smin, dot, emd;
	double jeff, js;
};

// ------------------------------------------------------------------------------------------------
// epilogue: raw singles (Feature.cpp), first = Feature::compute's first argument
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double i128_to_double(__int128 v)
{
	bool neg = v < 0;
	unsigned __int128 u = neg ? (unsigned __int128)(-v) : (unsigned __int128)v;
	double d = (double)(u64)(u >> 64) * 18446744073709551616.0 + (double)(u64)u;
	return neg ? -d : d;
}

// Feature<T>::pearson (Feature.cpp:794-811) through the exact integer identity
//   N*sum((P-a)(Q-b)) = N*S_pq - magP*sumQ - magQ*sumP + magP*magQ,  a = magP/N, b = magQ/N
__device__ __forceinline__ double pearson_exact(u64 N, u64 dot, const Side &p, const Side &q)
{
	__int128 n = (__int128)N;
	__int128 ndot = n * (__int128)dot - (__int128)p.mag * (__int128)q.sum - (__int128)q.mag * (__int128)p.sum +
			(__int128)p.mag * (__int128)q.mag;
	__int128 nnp = n * (__int128)p.sumsq - 2 * (__int128)p.mag * (__int128)p.sum + (__int128)p.mag * (__int128)p.mag;
	__int128 nnq = n * (__int128)q.sumsq - 2 * (__int128)q.mag * (__int128)q.sum + (__int128)q.mag * (__int128)q.mag;
	return i128_to_double(ndot) / sqrt(i128_to_double(nnp) * i128_to_double(nnq));
}

// a / b for a model constant b whose correctly rounded reciprocal y = RN(1/b) was computed on the host (Markstein):
// q0 = RN(a*y) is within one ulp of a/b, r = a - b*q0 is exact in one fused multiply-add, RN(q0 + r*y) is the correctly
// rounded quotient, i.e. bit-identical to a / b.  Anything outside the comfortable range (infinities, NaN, values near
// under/overflow, unusable reciprocal) takes the real division.
__device__ __forceinline__ double div_const(double a, double b, double y, int ok)
{
	if (ok) {
		const double q0 = __dmul_rn(a, y);
		const double r = __fma_rn(-b, q0, a);
		const double q = __fma_rn(r, y, q0);
		if ((fabs(q0) < 1e200 && fabs(a) > 1e-200) || a == 0.0) {
			return q;
		}
	}
	return a / b;
}

// pearson_exact in 64-bit integers when every term provably fits (always, for the 1 KiB / 128 KiB histograms of real
// data); identical value: the same integer converted to double once
__device__ __forceinline__ double pearson_fast(u64 N, u64 dot, const Side &p, const Side &q)
{
	if ((p.mag | q.mag | p.sum | q.sum) < (1ULL << 30) && N <= (1ULL << 20) && dot < (1ULL << 40) &&
	    (p.sumsq | q.sumsq) < (1ULL << 40)) {
		const long long n = (long long)N, pm = (long long)p.mag, qm = (long long)q.mag;
		const long long ndot = n * (long long)dot - pm * (long long)q.sum - qm * (long long)p.sum + pm * qm;
		const long long nnp = n * (long long)p.sumsq - 2 * pm * (long long)p.sum + pm * pm;
		const long long nnq = n * (long long)q.sumsq - 2 * qm * (long long)q.sum + qm * qm;
		return (double)ndot / sqrt((double)nnp * (double)nnq);
	}
	return pearson_exact(N, dot, p, q);
}

// eval_pair for 8/16-bit histograms without the interpreter: one guarded straight-line block per single code (uniform
// branches on the model), normalisation through div_const, the logistic only where the decision or the score needs it.
// Same operations in the same order as eval_pair, so the values are the same bits.
// lazy: the caller keeps `score` only for close pairs; with bias 0, sum < -1e-6 means logistic(sum) < 0.5 - 2e-7, the
// pair is not close whatever the last-bit rounding of exp, and exp + division are skipped.
__device__ __forceinline__ int eval_pair_fast(const DevModel &dm, u64 N, const RedN &r, const Side &p, const Side &q, bool lazy,
					      double &score, double &d0, int &close)
{
	double cache[MC2_MAX_SINGLES];
	int bad = 0;
#define MC2_PUT(CODE, RAW)                                                                                         \
	{                                                                                                          \
		const double nv_ = div_const((RAW) - dm.cmin[CODE], dm.crange[CODE], dm.crcp[CODE], dm.crcp_ok[CODE]); \
		if (isnan(nv_)) {                                                                                  \
			bad |= 1;                                                                                  \
		}                                                                                                  \
		cache[dm.slot[CODE]] = dm.csim[CODE] ? nv_ : 1 - nv_;                                              \
	}
	if (dm.slot[SC_MANHATTAN] >= 0) { // Feature.cpp:858-871 (int accumulator)
		MC2_PUT(SC_MANHATTAN, (double)(int)(p.sum + q.sum - 2 * r.smin));
	}
	if (dm.slot[SC_EUCLIDEAN] >= 0 || dm.slot[SC_SIMRATIO] >= 0) {
		const double rn2 = sqrt((double)(p.sumsq + q.sumsq - 2 * r.dot));
		if (dm.slot[SC_EUCLIDEAN] >= 0) { // Feature.cpp:1112-1124
			MC2_PUT(SC_EUCLIDEAN, rn2);
		}
		if (dm.slot[SC_SIMRATIO] >= 0) { // Feature.cpp:828-841
			const double dot = (double)r.dot;
			MC2_PUT(SC_SIMRATIO, dot / (dot + rn2));
		}
	}
	if (dm.slot[SC_NORMALIZED_VECTORS] >= 0) { // Feature.cpp:1170-1184 (u64 product, wraps like the reference)
		MC2_PUT(SC_NORMALIZED_VECTORS, (double)r.dot / sqrt((double)(p.sumsq * q.sumsq)));
	}
	if (dm.slot[SC_PEARSON] >= 0) {
		MC2_PUT(SC_PEARSON, pearson_fast(N, r.dot, p, q));
	}
	if (dm.slot[SC_INTERSECTION] >= 0) { // Feature.cpp:763-777
		MC2_PUT(SC_INTERSECTION, (double)(2 * r.smin) / (double)(p.mag + q.mag));
	}
	if (dm.slot[SC_EMD] >= 0) { // Feature.cpp:1504-1518
		MC2_PUT(SC_EMD, (double)r.emd);
	}
	if (dm.slot[SC_LENGTHD] >= 0) { // Feature.cpp:873-887 (throws 123 on a zero length)
		if (p.len == 0 || q.len == 0) {
			bad |= 1;
		}
		MC2_PUT(SC_LENGTHD, (double)(p.len > q.len ? p.len - q.len : q.len - p.len));
	}
	if (dm.slot[SC_KULCZYNSKI2] >= 0) { // Feature.cpp:681-695
		const double ap = (double)p.mag / (double)N;
		const double aq = (double)q.mag / (double)N;
		const double coeff = (double)N * (ap + aq) / (2 * ap * aq);
		MC2_PUT(SC_KULCZYNSKI2, coeff * (double)r.smin);
	}
	if (dm.slot[SC_JEFFEREY] >= 0) { // Feature.cpp:1230-1263
		MC2_PUT(SC_JEFFEREY, r.jeff);
	}
	if (dm.slot[SC_JENSEN_SHANNON] >= 0) { // Feature.cpp:983-1009
		MC2_PUT(SC_JENSEN_SHANNON, r.js / 2);
	}
#undef MC2_PUT
	double sum = dm.weight[0];
	d0 = 0;
#pragma unroll 1
	for (int c = 0; c < dm.n_combos; c++) {
		const int *ix = dm.idx[c];
		const int kind = dm.kind[c];
		double d;
		if (kind == MC2_COMBO_XY || kind == MC2_COMBO_X2Y2) {
			double prod = 1;
			for (int t = 0; t < dm.nidx[c]; t++) {
				const double x = cache[ix[t]];
				prod *= kind == MC2_COMBO_XY ? x : x * x;
			}
			d = prod;
		} else if (kind == MC2_COMBO_XY2) {
			d = cache[ix[0]] * cache[ix[1]] * cache[ix[1]];
		} else {
			d = cache[ix[0]] * cache[ix[0]] * cache[ix[1]];
		}
		if (c == 0) {
			d0 = d;
		}
		sum += dm.weight[c + 1] * d;
	}
	if (dm.regression) { // Predictor::p_predict, Predictor.cpp:284-300
		score = sum < 0 ? 0 : (sum > 1 ? 1 : sum);
		close = 0;
	} else if (lazy && dm.bias == 0.0 && sum < -1e-6) {
		score = 0;
		close = 0;
	} else {
		score = 1.0 / (1 + exp(-sum)) + dm.bias;
		close = round(score) > 0;
	}
	return bad;
}

} // namespace mc2
