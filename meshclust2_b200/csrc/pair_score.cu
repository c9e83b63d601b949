// pair_score.cu — K2: pair features + GLM score + cutoff, gather / one-vs-many form (sm_100a).
//
// Replaces, for whole batches of pairs, the per-pair chain of the reference
//   Feature<T>::compute  (src/predict/Feature.h:197-201, Feature.cpp:136-171)
//   Feature<T>::operator() combos (Feature.h:205-239)
//   Trainer<T>::classify / Predictor<T>::p_close / p_predict (Trainer.cpp:112-120, Predictor.cpp:284-333)
// One warp walks one pair's two histogram rows with coalesced 16-byte loads and produces the three exact
// integer reductions every "fast" single derives from:
//     S_sad = sum |p-q|   (VABSDIFF4.U8.ACC, 4 bins / instruction)      -> manhattan, intersection, kulczynski2
//     S_dot = sum p*q     (IDP.4A, 4 bins / instruction)                -> euclidean, simratio, normalized_vectors, pearson
//     S_emd = sum |cumP-cumQ| (per-lane prefix via IDP.4A with +1/-1 byte masks, warp scan, VABSDIFF accumulate)
// The per-pair fp64 epilogue (singles -> normalise -> combos -> w.x -> logistic -> cutoff) runs one pair per
// lane after every 32 pairs, so it costs 1/32 of a warp-serial epilogue.
// HBM-bound by design: algorithmic bytes per pair = N*w (streamed row) + side-band, see DESIGN.md.
#include "mc2_internal.cuh"
#include <math_constants.h>
#include <cstring>

namespace mc2 {

// ------------------------------------------------------------------------------------------------
// small device helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ u32 vabsdiff4_acc(u32 a, u32 b, u32 c)
{
	u32 d;
	asm("vabsdiff4.u32.u32.u32.add %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
	return d;
}
__device__ __forceinline__ int dp4a_us(u32 a, u32 b_s8, int c)
{
	int d;
	asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b_s8), "r"(c));
	return d;
}
__device__ __forceinline__ u32 dp2a_lo_uu(u32 a, u32 b, u32 c)
{
	u32 d;
	asm("dp2a.lo.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
	return d;
}
__device__ __forceinline__ u32 dp2a_hi_uu(u32 a, u32 b, u32 c)
{
	u32 d;
	asm("dp2a.hi.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
	return d;
}
__device__ __forceinline__ int dp2a_lo_us(u32 a, u32 b_s8, int c)
{
	int d;
	asm("dp2a.lo.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b_s8), "r"(c));
	return d;
}
__device__ __forceinline__ uint4 ldg_stream(const void *p)
{
	// streamed once: do not keep in L1
	uint4 r;
	asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
		     : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
		     : "l"(p));
	return r;
}
__device__ __forceinline__ uint4 ldg_keep(const void *p)
{
	return __ldg(reinterpret_cast<const uint4 *>(p));
}

// exact warp sum of a u64 per lane through three 24-bit limbs and the REDUX unit
__device__ __forceinline__ u64 warp_sum_u64(u64 v)
{
	u32 lo = (u32)(v & 0xFFFFFFu);
	u32 mid = (u32)((v >> 24) & 0xFFFFFFu);
	u32 hi = (u32)(v >> 48);
	u64 s = __reduce_add_sync(0xffffffffu, lo);
	// most totals fit the low limbs; the ballot keeps the extra REDUX off the common path
	if (__any_sync(0xffffffffu, (mid | hi) != 0)) {
		s += (u64)__reduce_add_sync(0xffffffffu, mid) << 24;
		s += (u64)__reduce_add_sync(0xffffffffu, hi) << 48;
	}
	return s;
}
__device__ __forceinline__ double warp_sum_f64(double v)
{
#pragma unroll
	for (int d = 16; d > 0; d >>= 1) {
		v += __shfl_xor_sync(0xffffffffu, v, d);
	}
	return v;
}
__device__ __forceinline__ s64 shfl_up_s64(s64 v, int d)
{
	return __shfl_up_sync(0xffffffffu, v, d);
}

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
// accumulators reproducing the reference's type-dependent arithmetic for 32/64-bit histograms
struct RedW {
	int man;
	u64 euc, nv_sum, nv_d1, nv_d2, smin, inter2, emd, norm2;
	double dpq, dpp, dqq, jeff, js;
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

__device__ int raw_single_narrow(int code, u64 N, const RedN &r, const Side &p, const Side &q, double *out)
{
	switch (code) {
	case SC_MANHATTAN: { // Feature.cpp:858-871 (int accumulator)
		u64 sad = p.sum + q.sum - 2 * r.smin;
		*out = (double)(int)sad;
		return 0;
	}
	case SC_EUCLIDEAN: // Feature.cpp:1112-1124
		*out = sqrt((double)(p.sumsq + q.sumsq - 2 * r.dot));
		return 0;
	case SC_NORMALIZED_VECTORS: // Feature.cpp:1170-1184 (u64 product, wraps like the reference)
		*out = (double)r.dot / sqrt((double)(p.sumsq * q.sumsq));
		return 0;
	case SC_PEARSON:
		*out = pearson_exact(N, r.dot, p, q);
		return 0;
	case SC_INTERSECTION: // Feature.cpp:763-777
		*out = (double)(2 * r.smin) / (double)(p.mag + q.mag);
		return 0;
	case SC_EMD: // Feature.cpp:1504-1518
		*out = (double)r.emd;
		return 0;
	case SC_LENGTHD: // Feature.cpp:873-887 (throws 123 on a zero length)
		if (p.len == 0 || q.len == 0) {
			return 1;
		}
		*out = (double)(p.len > q.len ? p.len - q.len : q.len - p.len);
		return 0;
	case SC_KULCZYNSKI2: { // Feature.cpp:681-695
		double ap = (double)p.mag / (double)N;
		double aq = (double)q.mag / (double)N;
		double coeff = (double)N * (ap + aq) / (2 * ap * aq);
		*out = coeff * (double)r.smin;
		return 0;
	}
	case SC_SIMRATIO: { // Feature.cpp:828-841
		double dot = (double)r.dot;
		u64 norm2 = p.sumsq + q.sumsq - 2 * r.dot;
		*out = dot / (dot + sqrt((double)norm2));
		return 0;
	}
	case SC_JEFFEREY: // Feature.cpp:1230-1263
		*out = r.jeff;
		return 0;
	case SC_JENSEN_SHANNON: // Feature.cpp:983-1009
		*out = r.js / 2;
		return 0;
	}
	return 2;
}

__device__ int raw_single_wide(int code, u64 N, const RedW &r, const Side &p, const Side &q, double *out)
{
	switch (code) {
	case SC_MANHATTAN:
		*out = (double)r.man;
		return 0;
	case SC_EUCLIDEAN:
		*out = sqrt((double)r.euc);
		return 0;
	case SC_NORMALIZED_VECTORS:
		*out = (double)r.nv_sum / sqrt((double)(r.nv_d1 * r.nv_d2));
		return 0;
	case SC_PEARSON: { // fp64 sums of products; the reference's own loop is fp64 too
		double dap = (double)p.mag / (double)N;
		double daq = (double)q.mag / (double)N;
		double n = (double)N;
		double dot = r.dpq - dap * (double)q.sum - daq * (double)p.sum + n * dap * daq;
		double np = r.dpp - 2 * dap * (double)p.sum + n * dap * dap;
		double nq = r.dqq - 2 * daq * (double)q.sum + n * daq * daq;
		*out = dot / sqrt(np * nq);
		return 0;
	}
	case SC_INTERSECTION:
		*out = (double)r.inter2 / (double)(p.mag + q.mag);
		return 0;
	case SC_EMD:
		*out = (double)r.emd;
		return 0;
	case SC_LENGTHD:
		if (p.len == 0 || q.len == 0) {
			return 1;
		}
		*out = (double)(p.len > q.len ? p.len - q.len : q.len - p.len);
		return 0;
	case SC_KULCZYNSKI2: {
		double ap = (double)p.mag / (double)N;
		double aq = (double)q.mag / (double)N;
		double coeff = (double)N * (ap + aq) / (2 * ap * aq);
		*out = coeff * (double)r.smin;
		return 0;
	}
	case SC_SIMRATIO: {
		double dot = (double)r.nv_sum;
		*out = dot / (dot + sqrt((double)r.norm2));
		return 0;
	}
	case SC_JEFFEREY:
		*out = r.jeff;
		return 0;
	case SC_JENSEN_SHANNON:
		*out = r.js / 2;
		return 0;
	}
	return 2;
}

// normalise (Feature.cpp:136-154), combos (Feature.h:205-239), GLM sum, logistic + bias (Predictor.cpp:316-320)
// returns a bit mask: 1 = the reference would throw, 2 = internal (unknown single code)
template <typename RED, bool WIDE>
__device__ int eval_pair(const DevModel &dm, u64 N, const RED &r, const Side &first, const Side &second, double *raw_out,
			 double *cache_out, double &score, double &d0, int &close)
{
	double cache[MC2_MAX_SINGLES];
	int bad = 0;
#pragma unroll 1
	for (int i = 0; i < dm.n_singles; i++) {
		double v = 0;
		int rc;
		if constexpr (WIDE) {
			rc = raw_single_wide(dm.code[i], N, r, first, second, &v);
		} else {
			rc = raw_single_narrow(dm.code[i], N, r, first, second, &v);
		}
		bad |= rc;
		if (raw_out) {
			raw_out[i] = v;
		}
		double nv = (v - dm.smin[i]) / (dm.smax[i] - dm.smin[i]);
		if (isnan(nv)) {
			bad |= 1;
		}
		cache[i] = dm.is_sim[i] ? nv : 1 - nv;
		if (cache_out) {
			cache_out[i] = cache[i];
		}
	}
	double sum = dm.weight[0];
	d0 = 0;
#pragma unroll 1
	for (int c = 0; c < dm.n_combos; c++) {
		double d;
		const int *ix = dm.idx[c];
		switch (dm.kind[c]) {
		case MC2_COMBO_XY: {
			double prod = 1;
			for (int t = 0; t < dm.nidx[c]; t++) {
				prod *= cache[ix[t]];
			}
			d = prod;
			break;
		}
		case MC2_COMBO_X2Y2: {
			double prod = 1;
			for (int t = 0; t < dm.nidx[c]; t++) {
				prod *= cache[ix[t]] * cache[ix[t]];
			}
			d = prod;
			break;
		}
		case MC2_COMBO_XY2:
			d = cache[ix[0]] * cache[ix[1]] * cache[ix[1]];
			break;
		default: // MC2_COMBO_X2Y
			d = cache[ix[0]] * cache[ix[0]] * cache[ix[1]];
			break;
		}
		if (c == 0) {
			d0 = d;
		}
		sum += dm.weight[c + 1] * d;
	}
	if (dm.regression) { // Predictor::p_predict, Predictor.cpp:284-300
		score = sum < 0 ? 0 : (sum > 1 ? 1 : sum);
		close = 0;
	} else {
		score = 1.0 / (1 + exp(-sum)) + dm.bias;
		close = round(score) > 0;
	}
	return bad;
}

template <typename RED, bool WIDE>
__device__ void finish_pair(const DevModel &dm, const PairArgs &a, u64 j, u64 N, const RED &r, const Side &first,
			    const Side &second)
{
	double score, d0;
	int close;
	const u64 S = (u64)dm.n_singles;
	int bad = eval_pair<RED, WIDE>(dm, N, r, first, second, a.raw ? a.raw + j * S : nullptr,
				       a.cache ? a.cache + j * S : nullptr, score, d0, close);
	if (bad) {
		atomicOr(a.err, bad & 1 ? 1 : 2);
	}
	if (a.score) {
		a.score[j] = score;
	}
	if (a.dist) {
		a.dist[j] = d0;
	}
	if (a.close) {
		a.close[j] = (uint8_t)close;
	}
	if (a.skipped) {
		a.skipped[j] = 0;
	}
	if (a.n_close && close) {
		atomicAdd(a.n_close, 1ULL);
	}
}

__device__ __forceinline__ void write_skipped(const PairArgs &a, u64 j)
{
	if (a.score) {
		a.score[j] = CUDART_NAN;
	}
	if (a.dist) {
		a.dist[j] = CUDART_NAN;
	}
	if (a.close) {
		a.close[j] = 0;
	}
	if (a.skipped) {
		a.skipped[j] = 1;
	}
}

// rows + length prefilter for pair j (Trainer.cpp:39-48 / 82-91 / 126-130: u64 truncation of len*cutoff, len/cutoff)
__device__ __forceinline__ bool resolve_pair(const PairArgs &a, u64 j, u64 &ra, u64 &rb)
{
	ra = a.ia ? a.ia[j] : a.a_begin + (a.a_bc ? 0 : j);
	rb = a.ib ? a.ib[j] : a.b_begin + (a.b_bc ? 0 : j);
	if (a.len_filter) {
		u64 la = a.sbA.len[ra], lb = a.sbB.len[rb];
		u64 anchor = a.anchor_is_b ? lb : la;
		u64 other = a.anchor_is_b ? la : lb;
		u64 min_len = (u64)((double)anchor * a.cutoff);
		u64 max_len = (u64)((double)anchor / a.cutoff);
		if (other < min_len || other > max_len) {
			return false;
		}
	}
	return true;
}

// ------------------------------------------------------------------------------------------------
// fast path: 8/16-bit bins, rows a multiple of 512 bytes, bin sums < 2^27
// ------------------------------------------------------------------------------------------------
template <int NEED>
__device__ __forceinline__ void slab_u8(const uint4 &pv, const uint4 &qv, u32 &sad, u32 &dot, int (&l)[16], int &tot)
{
	const u32 pw[4] = {pv.x, pv.y, pv.z, pv.w};
	const u32 qw[4] = {qv.x, qv.y, qv.z, qv.w};
	int c = 0;
#pragma unroll
	for (int w = 0; w < 4; w++) {
		u32 p = pw[w], q = qw[w];
		if (NEED & NEED_MIN) {
			sad = vabsdiff4_acc(p, q, sad);
		}
		if (NEED & NEED_DOT) {
			dot = __dp4a(p, q, dot);
		}
		if (NEED & NEED_EMD) {
			u32 w01 = __byte_perm(p, q, 0x5140); // p0 q0 p1 q1
			u32 w23 = __byte_perm(p, q, 0x7362); // p2 q2 p3 q3
			l[4 * w + 0] = dp4a_us(w01, 0x0000FF01u, c);
			l[4 * w + 1] = dp4a_us(w01, 0xFF01FF01u, c);
			c = l[4 * w + 1];
			l[4 * w + 2] = dp4a_us(w23, 0x0000FF01u, c);
			l[4 * w + 3] = dp4a_us(w23, 0xFF01FF01u, c);
			c = l[4 * w + 3];
		}
	}
	tot = c;
}

template <int NEED>
__device__ __forceinline__ void slab_u16(const uint4 &pv, const uint4 &qv, u32 &smin, u32 &dlo, u32 &dhi, int (&l)[8],
					 int &tot)
{
	const u32 pw[4] = {pv.x, pv.y, pv.z, pv.w};
	const u32 qw[4] = {qv.x, qv.y, qv.z, qv.w};
	int c = 0;
#pragma unroll
	for (int w = 0; w < 4; w++) {
		u32 p = pw[w], q = qw[w];
		if (NEED & NEED_MIN) {
			u32 mn = __vminu2(p, q);
			smin = dp2a_lo_uu(mn, 0x0101u, smin);
		}
		if (NEED & NEED_DOT) {
			u32 qp = __byte_perm(q, q, 0x3120); // q0.lo q1.lo q0.hi q1.hi
			dlo = dp2a_lo_uu(p, qp, dlo);
			dhi = dp2a_hi_uu(p, qp, dhi);
		}
		if (NEED & NEED_EMD) {
			l[2 * w + 0] = dp2a_lo_us(q, 0x00FFu, (int)dp2a_lo_uu(p, 0x0001u, (u32)c));
			l[2 * w + 1] = dp2a_lo_us(q, 0xFFFFu, (int)dp2a_lo_uu(p, 0x0101u, (u32)c));
			c = l[2 * w + 1];
		}
	}
	tot = c;
}

// exclusive prefix of `tot` across lanes, plus the warp total
__device__ __forceinline__ int warp_excl_scan(int tot, int lane, int &warp_total)
{
	int x = tot;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		int y = __shfl_up_sync(0xffffffffu, x, d);
		if (lane >= d) {
			x += y;
		}
	}
	warp_total = __shfl_sync(0xffffffffu, x, 31);
	return x - tot;
}

template <typename T, int NEED>
__device__ __forceinline__ RedN reduce_rows_fast(const T *__restrict__ P, const T *__restrict__ Q, u32 slabs, int lane,
						 bool q_hot)
{
	u64 a_min = 0, a_dot = 0, a_emd = 0;
	int carry = 0;
	const char *pp = reinterpret_cast<const char *>(P) + lane * 16;
	const char *qq = reinterpret_cast<const char *>(Q) + lane * 16;
	uint4 pv = ldg_stream(pp);
	uint4 qv = q_hot ? ldg_keep(qq) : ldg_stream(qq);
#pragma unroll 1
	for (u32 s = 0; s < slabs; s++) {
		uint4 pn = pv, qn = qv;
		if (s + 1 < slabs) { // prefetch the next slab before the ALU work of this one
			pn = ldg_stream(pp + (size_t)(s + 1) * 512);
			qn = q_hot ? ldg_keep(qq + (size_t)(s + 1) * 512) : ldg_stream(qq + (size_t)(s + 1) * 512);
		}
		int tot = 0;
		if constexpr (sizeof(T) == 1) {
			u32 sad = 0, dot = 0;
			int l[16];
			slab_u8<NEED>(pv, qv, sad, dot, l, tot);
			a_min += sad; // holds sum|p-q| for u8; converted after the loop
			a_dot += dot;
			if (NEED & NEED_EMD) {
				int wt;
				int off = warp_excl_scan(tot, lane, wt) + carry;
				carry += wt;
				u32 e = 0;
#pragma unroll
				for (int i = 0; i < 16; i++) {
					e = __sad(l[i], -off, e);
				}
				a_emd += e;
			}
		} else {
			u32 smin = 0, dlo = 0, dhi = 0;
			int l[8];
			slab_u16<NEED>(pv, qv, smin, dlo, dhi, l, tot);
			a_min += smin;
			a_dot += (u64)dlo + ((u64)dhi << 8);
			if (NEED & NEED_EMD) {
				int wt;
				int off = warp_excl_scan(tot, lane, wt) + carry;
				carry += wt;
				u32 e = 0;
#pragma unroll
				for (int i = 0; i < 8; i++) {
					e = __sad(l[i], -off, e);
				}
				a_emd += e;
			}
		}
		pv = pn;
		qv = qn;
	}
	RedN r;
	r.smin = (NEED & NEED_MIN) ? warp_sum_u64(a_min) : 0;
	r.dot = (NEED & NEED_DOT) ? warp_sum_u64(a_dot) : 0;
	r.emd = (NEED & NEED_EMD) ? warp_sum_u64(a_emd) : 0;
	r.jeff = r.js = 0;
	return r;
}

template <typename T, int NEED>
__global__ void __launch_bounds__(256) pair_fast_kernel(const __grid_constant__ DevModel dm, const __grid_constant__ PairArgs a)
{
	const int lane = threadIdx.x & 31;
	const u64 warps_total = (u64)gridDim.x * (blockDim.x >> 5);
	const u64 warp_id = (u64)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
	const u32 slabs = (u32)(a.N * sizeof(T) / 512);
	const T *A = reinterpret_cast<const T *>(a.binsA);
	const T *B = reinterpret_cast<const T *>(a.binsB);
	const u64 groups = (a.n_pairs + 31) / 32;
	// the streamed side is the one that is not broadcast; the broadcast row stays hot in L1
	const bool a_hot = a.a_bc && !a.ia;
	const bool b_hot = a.b_bc && !a.ib;
	for (u64 g = warp_id; g < groups; g += warps_total) {
		const u64 j = g * 32 + lane;
		const bool valid = j < a.n_pairs;
		u64 ra = 0, rb = 0;
		bool go = valid && resolve_pair(a, j, ra, rb);
		RedN mine;
		mine.smin = mine.dot = mine.emd = 0;
		mine.jeff = mine.js = 0;
		unsigned active = __ballot_sync(0xffffffffu, go);
		while (active) {
			int pi = __ffs(active) - 1;
			active &= active - 1;
			u64 xa = __shfl_sync(0xffffffffu, ra, pi);
			u64 xb = __shfl_sync(0xffffffffu, rb, pi);
			RedN r;
			if (a_hot) { // stream B, keep A
				r = reduce_rows_fast<T, NEED>(B + xb * a.N, A + xa * a.N, slabs, lane, true);
			} else {
				r = reduce_rows_fast<T, NEED>(A + xa * a.N, B + xb * a.N, slabs, lane, b_hot);
			}
			if (lane == pi) {
				mine = r;
			}
		}
		if (go) {
			Side sa = load_side(a.sbA, ra), sb = load_side(a.sbB, rb);
			if constexpr (sizeof(T) == 1) {
				// u8 path accumulated sum|p-q|; S_min = (sumP + sumQ - sad) / 2
				if (NEED & NEED_MIN) {
					mine.smin = (sa.sum + sb.sum - mine.smin) >> 1;
				}
			}
			finish_pair<RedN, false>(dm, a, j, a.N, mine, sa, sb);
		} else if (valid) {
			write_skipped(a, j);
		}
	}
}

// ------------------------------------------------------------------------------------------------
// generic path: any width, any N, log features; lane-interleaved bins, 64-bit accumulators.
// Reproduces the reference's type-dependent integer arithmetic for 32/64-bit bins.
// ------------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ void reduce_rows_generic(const T *__restrict__ P, const T *__restrict__ Q, u64 N, int lane,
						    int need, u64 magP, u64 magQ, RedN &rn, RedW &rw)
{
	constexpr bool WIDE = sizeof(T) > 2;
	u64 smin = 0, dot = 0, emd = 0;
	int man = 0;
	u64 euc = 0, d1 = 0, d2 = 0, inter2 = 0, norm2 = 0;
	double dpq = 0, dpp = 0, dqq = 0, jeff = 0, js = 0;
	s64 carry = 0;
	for (u64 base = 0; base < N; base += 32) {
		u64 i = base + lane;
		bool in = i < N;
		T p = in ? P[i] : (T)0;
		T q = in ? Q[i] : (T)0;
		smin += p < q ? p : q;
		if constexpr (WIDE) {
			// same expressions as Feature.cpp so the usual arithmetic conversions (and wrap-around) match
			man += p > q ? p - q : q - p;                       // :864-867
			euc += (p - q) * (p - q);                           // :1120-1121
			dot += p * q;                                       // :1179 / :837
			d1 += p * p;                                        // :1180
			d2 += q * q;                                        // :1181
			inter2 += 2 * (p < q ? p : q);                      // :773
			long long diff = p - q;                             // :836 (intmax_t from the T-typed difference)
			norm2 += diff * diff;                               // :838
			dpq += (double)p * (double)q;
			dpp += (double)p * (double)p;
			dqq += (double)q * (double)q;
		} else {
			dot += (u32)p * (u32)q;
		}
		if (need & NEED_EMD) { // Feature.cpp:1504-1518
			s64 d = (s64)((u64)p - (u64)q);
			s64 x = d;
#pragma unroll
			for (int s = 1; s < 32; s <<= 1) {
				s64 y = shfl_up_s64(x, s);
				if (lane >= s) {
					x += y;
				}
			}
			s64 c = x + carry;
			if (in) {
				emd += (u64)(c < 0 ? -c : c);
			}
			carry += __shfl_sync(0xffffffffu, x, 31);
		}
		if ((need & NEED_LOG) && in) {
			double pp = (double)p / (double)magP;
			double pq = (double)q / (double)magQ;
			double diff = pp - pq;
			jeff += diff * log(pp / pq);                        // :1240-1260
			double avg = 0.5 * (pp + pq);
			js += pp * log(pp / avg) + pq * log(pq / avg);      // :994-1006
		}
	}
	if constexpr (WIDE) {
		rw.man = __reduce_add_sync(0xffffffffu, man);
		rw.euc = warp_sum_u64(euc);
		rw.nv_sum = warp_sum_u64(dot);
		rw.nv_d1 = warp_sum_u64(d1);
		rw.nv_d2 = warp_sum_u64(d2);
		rw.smin = warp_sum_u64(smin);
		rw.inter2 = warp_sum_u64(inter2);
		rw.emd = warp_sum_u64(emd);
		rw.norm2 = warp_sum_u64(norm2);
		rw.dpq = warp_sum_f64(dpq);
		rw.dpp = warp_sum_f64(dpp);
		rw.dqq = warp_sum_f64(dqq);
		rw.jeff = (need & NEED_LOG) ? warp_sum_f64(jeff) : 0;
		rw.js = (need & NEED_LOG) ? warp_sum_f64(js) : 0;
	} else {
		rn.smin = warp_sum_u64(smin);
		rn.dot = warp_sum_u64(dot);
		rn.emd = warp_sum_u64(emd);
		rn.jeff = (need & NEED_LOG) ? warp_sum_f64(jeff) : 0;
		rn.js = (need & NEED_LOG) ? warp_sum_f64(js) : 0;
	}
}

template <typename T>
__global__ void __launch_bounds__(256) pair_generic_kernel(const __grid_constant__ DevModel dm, const __grid_constant__ PairArgs a)
{
	constexpr bool WIDE = sizeof(T) > 2;
	const int lane = threadIdx.x & 31;
	const u64 warps_total = (u64)gridDim.x * (blockDim.x >> 5);
	const u64 warp_id = (u64)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
	const T *A = reinterpret_cast<const T *>(a.binsA);
	const T *B = reinterpret_cast<const T *>(a.binsB);
	const u64 groups = (a.n_pairs + 31) / 32;
	for (u64 g = warp_id; g < groups; g += warps_total) {
		const u64 j = g * 32 + lane;
		const bool valid = j < a.n_pairs;
		u64 ra = 0, rb = 0;
		bool go = valid && resolve_pair(a, j, ra, rb);
		RedN mn;
		RedW mw;
		mn.smin = mn.dot = mn.emd = 0;
		mn.jeff = mn.js = 0;
		mw = RedW();
		unsigned active = __ballot_sync(0xffffffffu, go);
		while (active) {
			int pi = __ffs(active) - 1;
			active &= active - 1;
			u64 xa = __shfl_sync(0xffffffffu, ra, pi);
			u64 xb = __shfl_sync(0xffffffffu, rb, pi);
			u64 magA = a.sbA.mag[xa], magB = a.sbB.mag[xb];
			RedN rn;
			RedW rw;
			reduce_rows_generic<T>(A + xa * a.N, B + xb * a.N, a.N, lane, dm.need, magA, magB, rn, rw);
			if (lane == pi) {
				if constexpr (WIDE) {
					mw = rw;
				} else {
					mn = rn;
				}
			}
		}
		if (go) {
			Side sa = load_side(a.sbA, ra), sb = load_side(a.sbB, rb);
			if constexpr (WIDE) {
				finish_pair<RedW, true>(dm, a, j, a.N, mw, sa, sb);
			} else {
				finish_pair<RedN, false>(dm, a, j, a.N, mn, sa, sb);
			}
		} else if (valid) {
			write_skipped(a, j);
		}
	}
}

// ------------------------------------------------------------------------------------------------
// DivergencePoint<T>::distance (DivergencePoint.cpp:70-82): thin kernel, warp per pair
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) distance_kernel(const __grid_constant__ PairArgs a, u64 *out)
{
	const int lane = threadIdx.x & 31;
	const u64 warps_total = (u64)gridDim.x * (blockDim.x >> 5);
	const u64 warp_id = (u64)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
	const T *A = reinterpret_cast<const T *>(a.binsA);
	const T *B = reinterpret_cast<const T *>(a.binsB);
	for (u64 j = warp_id; j < a.n_pairs; j += warps_total) {
		u64 ra = a.ia ? a.ia[j] : a.a_begin + (a.a_bc ? 0 : j);
		u64 rb = a.ib ? a.ib[j] : a.b_begin + (a.b_bc ? 0 : j);
		const T *P = A + ra * a.N, *Q = B + rb * a.N;
		u64 s = 0;
		for (u64 i = lane; i < a.N; i += 32) {
			T p = P[i], q = Q[i];
			s += p < q ? p : q;
		}
		s = warp_sum_u64(s);
		if (lane == 0) {
			u64 dist = s * 2;
			u64 mag = a.sbA.mag[ra] + a.sbB.mag[rb];
			double frac = (double)dist / (double)mag;
			out[j] = (u64)(10000.0 * (1.0 - frac * frac));
		}
	}
}

// ------------------------------------------------------------------------------------------------
// argmax / any-close reductions for get_close and merge (single CTA, inputs are 9 bytes per candidate)
// mode 0 (get_close): best = first position with the maximum dist among non-skipped; init (-1, -1); is_min = !any(close)
// mode 1 (merge):     best = last position p (sequential `best.second > dist ? best : (i,dist)`) among close ones, init (0, DBL_MIN)
// ------------------------------------------------------------------------------------------------
struct ArgOut {
	long long best;
	double best_dist;
	int is_min;
	int has; // mode 1: a close candidate was found
};

__global__ void __launch_bounds__(1024) argmax_kernel(const double *dist, const uint8_t *skipped, const uint8_t *close, u64 n,
						      int mode, ArgOut *out)
{
	__shared__ double s_d[32];
	__shared__ long long s_i[32];
	__shared__ int s_any[32];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	double bd = mode == 0 ? -1.0 : 2.2250738585072014e-308;
	long long bi = mode == 0 ? -1 : 0;
	bool has = false; // mode 1: whether bi was set by a candidate
	int any = 0;
	for (u64 i = threadIdx.x; i < n; i += blockDim.x) {
		if (skipped[i]) {
			continue;
		}
		double d = dist[i];
		any |= close[i];
		if (mode == 0) {
			if (d > bd) { // strict: the first maximum in sequence order wins
				bd = d;
				bi = (long long)i;
			}
		} else if (close[i]) {
			// sequential semantics: replace unless best.second > dist  => later ties win
			if (!(bd > d)) {
				bd = d;
				bi = (long long)i;
				has = true;
			}
		}
	}
	// combine: mode 0 prefers larger dist then smaller index; mode 1 prefers larger dist then larger index
	auto better = [&](double d2, long long i2, bool h2) {
		if (mode == 0) {
			if (i2 < 0) {
				return false;
			}
			return bi < 0 || d2 > bd || (d2 == bd && i2 < bi);
		}
		if (!h2) {
			return false;
		}
		return !has || d2 > bd || (d2 == bd && i2 > bi);
	};
	for (int s = 16; s > 0; s >>= 1) {
		double d2 = __shfl_xor_sync(0xffffffffu, bd, s);
		long long i2 = __shfl_xor_sync(0xffffffffu, bi, s);
		int h2 = __shfl_xor_sync(0xffffffffu, (int)has, s);
		any |= __shfl_xor_sync(0xffffffffu, any, s);
		if (better(d2, i2, h2 != 0)) {
			bd = d2;
			bi = i2;
			has = h2 != 0;
		}
	}
	if (lane == 0) {
		s_d[warp] = bd;
		s_i[warp] = has || mode == 0 ? bi : -2;
		s_any[warp] = any;
	}
	__syncthreads();
	if (warp == 0) {
		int nw = blockDim.x >> 5;
		bd = lane < nw ? s_d[lane] : (mode == 0 ? -1.0 : 2.2250738585072014e-308);
		long long raw = lane < nw ? s_i[lane] : (mode == 0 ? -1 : -2);
		has = mode == 1 && raw != -2;
		bi = mode == 1 && raw == -2 ? 0 : raw;
		any = lane < nw ? s_any[lane] : 0;
		for (int s = 16; s > 0; s >>= 1) {
			double d2 = __shfl_xor_sync(0xffffffffu, bd, s);
			long long i2 = __shfl_xor_sync(0xffffffffu, bi, s);
			int h2 = __shfl_xor_sync(0xffffffffu, (int)has, s);
			any |= __shfl_xor_sync(0xffffffffu, any, s);
			if (better(d2, i2, h2 != 0)) {
				bd = d2;
				bi = i2;
				has = h2 != 0;
			}
		}
		if (lane == 0) {
			out->best = bi;
			out->best_dist = bd;
			out->is_min = !any;
			out->has = mode == 0 ? (bi >= 0) : (int)has;
		}
	}
}

// ------------------------------------------------------------------------------------------------
// query-vs-database sweep with fused length prefilter, cutoff and survivor compaction
// (fastcar work(), src/fastcar/FC_Runner.cpp:427-470).  A warp takes one query row r and 32 consecutive
// database rows; the query row stays hot in L1, database rows stream (from L2 when the set fits its 126 MB).
// ------------------------------------------------------------------------------------------------
struct SweepArgs {
	u64 q0, q1, d0, d1;
	int upper_only;
	double cutoff;
	u64 max_out;
	u64 *out_q, *out_d;
	double *out_score;
	u64 *counters; // [0] survivors, [1] scored pairs
};

template <typename T, int NEED, bool FAST>
__global__ void __launch_bounds__(256) sweep_kernel(const __grid_constant__ DevModel dm, const __grid_constant__ PairArgs a,
						    const __grid_constant__ SweepArgs g)
{
	constexpr bool WIDE = sizeof(T) > 2;
	const int lane = threadIdx.x & 31;
	const u64 warps_total = (u64)gridDim.x * (blockDim.x >> 5);
	const u64 warp_id = (u64)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
	const u32 slabs = (u32)(a.N * sizeof(T) / 512);
	const T *Dm = reinterpret_cast<const T *>(a.binsA); // database = first argument of close(pts[i], query)
	const T *Qm = reinterpret_cast<const T *>(a.binsB);
	const u64 cblocks = (g.d1 - g.d0 + 31) / 32;
	const u64 groups = (g.q1 - g.q0) * cblocks;
	for (u64 grp = warp_id; grp < groups; grp += warps_total) {
		const u64 r = g.q0 + grp / cblocks;
		const u64 c = g.d0 + (grp % cblocks) * 32 + lane;
		bool go = c < g.d1 && (!g.upper_only || c > r);
		if (go) { // FC_Runner.cpp:435-444: size_t truncation, window on the database length
			u64 lq = a.sbB.len[r], lc = a.sbA.len[c];
			u64 begin_length = (u64)((double)lq * g.cutoff);
			u64 end_length = (u64)((double)lq / g.cutoff);
			go = lc >= begin_length && lc <= end_length;
		}
		RedN mn;
		RedW mw;
		mn.smin = mn.dot = mn.emd = 0;
		mn.jeff = mn.js = 0;
		mw = RedW();
		unsigned active = __ballot_sync(0xffffffffu, go);
		const unsigned scored = active;
		while (active) {
			int pi = __ffs(active) - 1;
			active &= active - 1;
			u64 xc = __shfl_sync(0xffffffffu, c, pi);
			if constexpr (FAST) {
				RedN rr = reduce_rows_fast<T, NEED>(Dm + xc * a.N, Qm + r * a.N, slabs, lane, true);
				if (lane == pi) {
					mn = rr;
				}
			} else {
				RedN rn;
				RedW rw;
				reduce_rows_generic<T>(Dm + xc * a.N, Qm + r * a.N, a.N, lane, dm.need, a.sbA.mag[xc], a.sbB.mag[r], rn, rw);
				if (lane == pi) {
					mn = rn;
					mw = rw;
				}
			}
		}
		int close = 0;
		double score = 0, d0;
		if (go) {
			Side sd = load_side(a.sbA, c), sq = load_side(a.sbB, r);
			int bad;
			if constexpr (WIDE) {
				bad = eval_pair<RedW, true>(dm, a.N, mw, sd, sq, nullptr, nullptr, score, d0, close);
			} else {
				if (FAST && sizeof(T) == 1 && (NEED & NEED_MIN)) {
					mn.smin = (sd.sum + sq.sum - mn.smin) >> 1;
				}
				bad = eval_pair<RedN, false>(dm, a.N, mn, sd, sq, nullptr, nullptr, score, d0, close);
			}
			if (bad) {
				atomicOr(a.err, bad & 1 ? 1 : 2);
			}
		}
		unsigned cm = __ballot_sync(0xffffffffu, close);
		u64 base = 0;
		if (lane == 0) {
			if (cm) {
				base = atomicAdd(g.counters, (u64)__popc(cm));
			}
			if (scored) {
				atomicAdd(g.counters + 1, (u64)__popc(scored));
			}
		}
		base = __shfl_sync(0xffffffffu, base, 0);
		if (close) {
			u64 idx = base + __popc(cm & ((1u << lane) - 1));
			if (idx < g.max_out) {
				g.out_q[idx] = r;
				g.out_d[idx] = c;
				g.out_score[idx] = score;
			}
		}
	}
}

// ------------------------------------------------------------------------------------------------
// host launchers
// ------------------------------------------------------------------------------------------------
static int grid_for(mc2_ctx *ctx, u64 n_pairs, int warps_per_cta, int ctas_per_sm)
{
	u64 groups = (n_pairs + 31) / 32;
	u64 want = (groups + warps_per_cta - 1) / warps_per_cta;
	u64 cap = (u64)ctx->sm_count * ctas_per_sm;
	u64 g = want < cap ? want : cap;
	return (int)(g ? g : 1);
}

template <typename T>
static void launch_fast_need(int need, int grid, cudaStream_t st, const DevModel &dm, const PairArgs &a)
{
	switch (need & 7) {
#define CASE(n)                                                    \
	case n:                                                    \
		pair_fast_kernel<T, n><<<grid, 256, 0, st>>>(dm, a); \
		break;
		CASE(0)
		CASE(1)
		CASE(2)
		CASE(3)
		CASE(4)
		CASE(5)
		CASE(6)
		CASE(7)
#undef CASE
	}
}

int launch_pair_score(mc2_ctx *ctx, const DevModel &dm, const PairArgs &a)
{
	if (a.n_pairs == 0) {
		return MC2_OK;
	}
	const u64 row_bytes = a.N * (u64)a.eb;
	const bool fast = a.eb <= 2 && row_bytes % 512 == 0 && !(dm.need & NEED_LOG) && a.max_sum < (1ULL << 27);
	prof_begin(ctx, 2);
	if (fast) {
		int grid = grid_for(ctx, a.n_pairs, 8, 8);
		if (a.eb == 1) {
			launch_fast_need<uint8_t>(dm.need, grid, ctx->stream, dm, a);
		} else {
			launch_fast_need<uint16_t>(dm.need, grid, ctx->stream, dm, a);
		}
	} else {
		int grid = grid_for(ctx, a.n_pairs, 8, 8);
		switch (a.eb) {
		case 1: pair_generic_kernel<uint8_t><<<grid, 256, 0, ctx->stream>>>(dm, a); break;
		case 2: pair_generic_kernel<uint16_t><<<grid, 256, 0, ctx->stream>>>(dm, a); break;
		case 4: pair_generic_kernel<uint32_t><<<grid, 256, 0, ctx->stream>>>(dm, a); break;
		case 8: pair_generic_kernel<unsigned long long><<<grid, 256, 0, ctx->stream>>>(dm, a); break;
		default: set_error("elem_bytes must be 1, 2, 4 or 8"); return MC2_ERR_ARG;
		}
	}
	prof_end(ctx);
	ctx->launches++;
	MC2_CUDA(cudaGetLastError());
	return MC2_OK;
}

int launch_argmax(mc2_ctx *ctx, const double *dist, const uint8_t *skipped, const uint8_t *close, u64 n, int mode, void *d_out)
{
	prof_begin(ctx, 4);
	argmax_kernel<<<1, 1024, 0, ctx->stream>>>(dist, skipped, close, n, mode, reinterpret_cast<ArgOut *>(d_out));
	prof_end(ctx);
	ctx->launches++;
	MC2_CUDA(cudaGetLastError());
	return MC2_OK;
}

int launch_distance(mc2_ctx *ctx, const PairArgs &a, u64 *d_out)
{
	if (a.n_pairs == 0) {
		return MC2_OK;
	}
	u64 want = (a.n_pairs + 7) / 8;
	u64 cap = (u64)ctx->sm_count * 8;
	int grid = (int)(want < cap ? want : cap);
	switch (a.eb) {
	case 1: distance_kernel<uint8_t><<<grid, 256, 0, ctx->stream>>>(a, d_out); break;
	case 2: distance_kernel<uint16_t><<<grid, 256, 0, ctx->stream>>>(a, d_out); break;
	case 4: distance_kernel<uint32_t><<<grid, 256, 0, ctx->stream>>>(a, d_out); break;
	case 8: distance_kernel<unsigned long long><<<grid, 256, 0, ctx->stream>>>(a, d_out); break;
	default: set_error("elem_bytes must be 1, 2, 4 or 8"); return MC2_ERR_ARG;
	}
	ctx->launches++;
	MC2_CUDA(cudaGetLastError());
	return MC2_OK;
}

template <typename T>
static void launch_sweep_fast(int need, int grid, cudaStream_t st, const DevModel &dm, const PairArgs &a, const SweepArgs &g)
{
	switch (need & 7) {
#define CASE(n)                                                          \
	case n:                                                          \
		sweep_kernel<T, n, true><<<grid, 256, 0, st>>>(dm, a, g); \
		break;
		CASE(0)
		CASE(1)
		CASE(2)
		CASE(3)
		CASE(4)
		CASE(5)
		CASE(6)
		CASE(7)
#undef CASE
	}
}

int launch_all_pairs(mc2_ctx *ctx, const DevModel &dm, const mc2_hset *q, u64 q0, u64 q1, const mc2_hset *d, u64 d0, u64 d1,
		     int upper_only, double cutoff, u64 max_out, u64 *d_out_q, u64 *d_out_d, double *d_out_score,
		     u64 *d_counters)
{
	PairArgs a;
	memset(&a, 0, sizeof a);
	a.binsA = d->bins;
	a.binsB = q->bins;
	a.sbA = Sideband{d->mag, d->sum, d->sumsq, d->len};
	a.sbB = Sideband{q->mag, q->sum, q->sumsq, q->len};
	a.N = q->N;
	a.eb = q->eb;
	a.err = ctx->d_err;
	a.max_sum = q->max_sum > d->max_sum ? q->max_sum : d->max_sum;
	SweepArgs g;
	g.q0 = q0;
	g.q1 = q1;
	g.d0 = d0;
	g.d1 = d1;
	g.upper_only = upper_only;
	g.cutoff = cutoff;
	g.max_out = max_out;
	g.out_q = d_out_q;
	g.out_d = d_out_d;
	g.out_score = d_out_score;
	g.counters = d_counters;
	const u64 groups = (q1 - q0) * ((d1 - d0 + 31) / 32);
	u64 want = (groups + 7) / 8, cap = (u64)ctx->sm_count * 8;
	int grid = (int)(want < cap ? want : cap);
	const u64 row_bytes = a.N * (u64)a.eb;
	const bool fast = a.eb <= 2 && row_bytes % 512 == 0 && !(dm.need & NEED_LOG) && a.max_sum < (1ULL << 27);
	prof_begin(ctx, 3);
	if (fast) {
		if (a.eb == 1) {
			launch_sweep_fast<uint8_t>(dm.need, grid, ctx->stream, dm, a, g);
		} else {
			launch_sweep_fast<uint16_t>(dm.need, grid, ctx->stream, dm, a, g);
		}
	} else {
		switch (a.eb) {
		case 1: sweep_kernel<uint8_t, 0, false><<<grid, 256, 0, ctx->stream>>>(dm, a, g); break;
		case 2: sweep_kernel<uint16_t, 0, false><<<grid, 256, 0, ctx->stream>>>(dm, a, g); break;
		case 4: sweep_kernel<uint32_t, 0, false><<<grid, 256, 0, ctx->stream>>>(dm, a, g); break;
		case 8: sweep_kernel<unsigned long long, 0, false><<<grid, 256, 0, ctx->stream>>>(dm, a, g); break;
		default: set_error("elem_bytes must be 1, 2, 4 or 8"); return MC2_ERR_ARG;
		}
	}
	prof_end(ctx);
	ctx->launches++;
	MC2_CUDA(cudaGetLastError());
	return MC2_OK;
}

} // namespace mc2
