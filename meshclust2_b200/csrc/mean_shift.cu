// mean_shift.cu — K3: cluster mean + closest member (sm_100a).
//
// Replaces get_mean (src/cluster/ClusterFactory.cpp:338-380) and the mean of mean_shift_update (:288-335) followed by
// Trainer<T>::closest (src/cluster/Trainer.cpp:144-157):
//   top[i]   = (sum over members of bins[i] as double) / count          (operator+= then operator/=)
//   dist_j   = DivergencePoint<T>::distance_d(member_j, top)            (src/clutil/DivergencePoint.cpp:55-66)
//   best     = first j with the minimum dist_j                          (strict <, sequential order)
// The column sums are exact integers (order independent); distance_d's running magnitude `mag += p_i + c_i` truncates a
// double to u64 at EVERY bin, so each member's bins are walked sequentially by one thread, with the same IEEE operations.
#include "mc2_internal.cuh"

namespace mc2 {

// column sums over the member rows: thread t of a CTA owns 4 consecutive bins, CTAs split the member list
template <typename T>
__global__ void __launch_bounds__(256) colsum_kernel(const T *__restrict__ bins, u64 N, const u64 *__restrict__ members, u64 n,
						      u64 *__restrict__ sums)
{
	const u64 per = (n + gridDim.y - 1) / gridDim.y;
	const u64 j0 = (u64)blockIdx.y * per, j1 = min(n, j0 + per);
	const u64 b = ((u64)blockIdx.x * blockDim.x + threadIdx.x) * 4;
	if (b >= N) {
		return;
	}
	u64 s0 = 0, s1 = 0, s2 = 0, s3 = 0;
	for (u64 j = j0; j < j1; j++) {
		const T *row = bins + members[j] * N + b;
		s0 += row[0];
		s1 += row[1];
		s2 += row[2];
		s3 += row[3];
	}
	atomicAdd(sums + b, s0);
	atomicAdd(sums + b + 1, s1);
	atomicAdd(sums + b + 2, s2);
	atomicAdd(sums + b + 3, s3);
}

__global__ void __launch_bounds__(256) mean_kernel(const u64 *__restrict__ sums, u64 N, u64 n, double *__restrict__ mean)
{
	u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < N) {
		mean[i] = (double)sums[i] / (double)n; // sum of integer-valued doubles is exact below 2^53
	}
}

// one thread per member, the mean staged through shared memory in chunks.
// The reference's `u64 mag += p_i + c_i` truncates at every step (a sequential chain of 4^k steps per member, not
// associative).  For 8/16-bit histograms every partial sum is below 2^53, so the chain can stay in the double domain:
// (u64)((double)mag + x) == floor((double)mag + x) exactly, and DADD + FRND is half the latency of DADD + F2I + I2F.  The
// rounded mean (T)round(c_i) depends only on the bin and is computed once per chunk by the staging threads.
template <typename T>
__global__ void __launch_bounds__(128) distance_d_kernel(const T *__restrict__ bins, u64 N, const u64 *__restrict__ members, u64 n,
							   const double *__restrict__ mean, double *__restrict__ dist)
{
	__shared__ double c[512];
	__shared__ T rc[512];
	const u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x;
	const bool on = j < n;
	const T *row = on ? bins + members[j] * N : bins;
	constexpr bool NARROW = sizeof(T) <= 2;
	u64 d = 0, mag = 0;
	double magd = 0;
	for (u64 base = 0; base < N; base += 512) {
		const u64 cnt = min((u64)512, N - base);
		__syncthreads();
		for (u64 t = threadIdx.x; t < cnt; t += blockDim.x) {
			const double ci = mean[base + t];
			c[t] = ci;
			rc[t] = (T)round(ci);
		}
		__syncthreads();
		if (on) {
			if (NARROW && cnt % 16 == 0) {
				// 16 bytes of the member's row per load
				constexpr int PER = 16 / (int)sizeof(T);
				for (u64 t = 0; t < cnt; t += PER) {
					const uint4 v = *reinterpret_cast<const uint4 *>(row + base + t);
					const T *pv = reinterpret_cast<const T *>(&v);
#pragma unroll
					for (int u = 0; u < PER; u++) {
						const T p = pv[u];
						const T r = rc[t + u];
						d += 2 * (u64)(p < r ? p : r);
						magd = floor(magd + ((double)p + c[t + u]));
					}
				}
			} else {
				for (u64 t = 0; t < cnt; t++) {
					const T p = row[base + t];
					const double ci = c[t];
					const T r = rc[t];
					d += 2 * (p < r ? p : r);
					if (NARROW) {
						magd = floor(magd + ((double)p + ci));
					} else {
						mag = (u64)((double)mag + ((double)p + ci)); // u64 += T + double, truncating each step
					}
				}
			}
		}
	}
	if (on) {
		if (NARROW) {
			mag = (u64)magd;
		}
		double frac = (double)d / (double)mag;
		dist[j] = 10000.0 * (1.0 - frac * frac);
	}
}

struct MinOut {
	long long best;
	double best_dist;
};

__global__ void __launch_bounds__(1024) argmin_kernel(const double *__restrict__ dist, u64 n, MinOut *out)
{
	__shared__ double s_d[32];
	__shared__ long long s_i[32];
	double bd = 0;
	long long bi = -1;
	for (u64 i = threadIdx.x; i < n; i += blockDim.x) {
		double d = dist[i];
		if (bi < 0 || d < bd) {
			bd = d;
			bi = (long long)i;
		}
	}
	auto better = [&](double d2, long long i2) { return i2 >= 0 && (bi < 0 || d2 < bd || (d2 == bd && i2 < bi)); };
	for (int s = 16; s > 0; s >>= 1) {
		double d2 = __shfl_xor_sync(0xffffffffu, bd, s);
		long long i2 = __shfl_xor_sync(0xffffffffu, bi, s);
		if (better(d2, i2)) {
			bd = d2;
			bi = i2;
		}
	}
	if ((threadIdx.x & 31) == 0) {
		s_d[threadIdx.x >> 5] = bd;
		s_i[threadIdx.x >> 5] = bi;
	}
	__syncthreads();
	if (threadIdx.x < 32) {
		bd = s_d[threadIdx.x];
		bi = s_i[threadIdx.x];
		for (int s = 16; s > 0; s >>= 1) {
			double d2 = __shfl_xor_sync(0xffffffffu, bd, s);
			long long i2 = __shfl_xor_sync(0xffffffffu, bi, s);
			if (better(d2, i2)) {
				bd = d2;
				bi = i2;
			}
		}
		if (threadIdx.x == 0) {
			out->best = bi;
			out->best_dist = bd;
		}
	}
}

template <typename T>
static int run_mean_closest(mc2_ctx *ctx, const mc2_hset *h, const u64 *d_members, u64 n, u64 *d_sums, double *d_mean, double *d_dist,
			    void *d_out, bool have_mean)
{
	const u64 N = h->N;
	if (!have_mean) {
		MC2_CUDA(cudaMemsetAsync(d_sums, 0, N * 8, ctx->stream));
		dim3 grid((unsigned)((N / 4 + 255) / 256), (unsigned)min((u64)ctx->sm_count * 4, (n + 63) / 64));
		colsum_kernel<T><<<grid, 256, 0, ctx->stream>>>((const T *)h->bins, N, d_members, n, d_sums);
		mean_kernel<<<(unsigned)((N + 255) / 256), 256, 0, ctx->stream>>>(d_sums, N, n, d_mean);
		ctx->launches += 2;
	}
	distance_d_kernel<T><<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>((const T *)h->bins, N, d_members, n, d_mean, d_dist);
	argmin_kernel<<<1, 1024, 0, ctx->stream>>>(d_dist, n, reinterpret_cast<MinOut *>(d_out));
	ctx->launches += 2;
	MC2_CUDA(cudaGetLastError());
	return MC2_OK;
}

int launch_mean_closest(mc2_ctx *ctx, const mc2_hset *h, const u64 *d_members, u64 n, u64 *d_sums, double *d_mean, double *d_dist,
			void *d_out, bool have_mean)
{
	switch (h->eb) {
	case 1: return run_mean_closest<uint8_t>(ctx, h, d_members, n, d_sums, d_mean, d_dist, d_out, have_mean);
	case 2: return run_mean_closest<uint16_t>(ctx, h, d_members, n, d_sums, d_mean, d_dist, d_out, have_mean);
	case 4: return run_mean_closest<uint32_t>(ctx, h, d_members, n, d_sums, d_mean, d_dist, d_out, have_mean);
	case 8: return run_mean_closest<u64>(ctx, h, d_members, n, d_sums, d_mean, d_dist, d_out, have_mean);
	}
	set_error("elem_bytes must be 1, 2, 4 or 8");
	return MC2_ERR_ARG;
}

// ------------------------------------------------------------------------------------------------
// Batched update stage.  mean_shift_update (src/cluster/ClusterFactory.cpp:288-335) runs once per center inside an
// `omp parallel for` (:639-642, :650-653); the iterations are independent (each reads the point lists and writes only its own
// center), so one launch serves every center of an iteration: CTA c walks the members of center c that survived
// Trainer::filter (flags left on the device by the pair kernel), builds their mean and returns the first member minimising
// distance_d to it -- the same arithmetic as colsum/mean/distance_d/argmin above, per segment.
// ------------------------------------------------------------------------------------------------
struct UpdateArgs {
	const void *bins;
	u64 N;
	const u64 *member_off;  // [n_centers + 1] positions into members / close / skipped
	const u64 *members;     // rows of the point set
	const uint8_t *close;   // round(score) > 0, per position
	const uint8_t *skipped; // outside the length window, per position
	u64 c_begin;            // first center of this launch
	double *mean;           // [centers of this launch x N] scratch
	long long *next;        // [n_centers] position inside the center's member list, -1 when no member survives
	u64 *n_good;            // [n_centers] survivors
};

template <typename T>
__global__ void __launch_bounds__(256) update_batch_kernel(const UpdateArgs a)
{
	__shared__ double c[512];
	__shared__ T rc[512];
	__shared__ u32 s_cnt[8];
	__shared__ double s_d[8];
	__shared__ long long s_i[8];
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const u64 cidx = a.c_begin + blockIdx.x;
	const u64 p0 = a.member_off[cidx], p1 = a.member_off[cidx + 1];
	const T *bins = reinterpret_cast<const T *>(a.bins);
	const u64 N = a.N;
	double *mean = a.mean + (u64)blockIdx.x * N;
	auto kept = [&](u64 p) { return a.close[p] && !a.skipped[p]; };
	// survivors
	u32 cnt = 0;
	for (u64 p = p0 + tid; p < p1; p += blockDim.x) {
		cnt += kept(p) ? 1u : 0u;
	}
	cnt = __reduce_add_sync(0xffffffffu, cnt);
	if (lane == 0) {
		s_cnt[warp] = cnt;
	}
	__syncthreads();
	u64 n_good = 0;
	for (int w = 0; w < 8; w++) {
		n_good += s_cnt[w];
	}
	if (n_good == 0) {
		if (tid == 0) {
			a.next[cidx] = -1;
			a.n_good[cidx] = 0;
		}
		return;
	}
	// mean of the survivors: exact integer column sums, then one division per bin (operator+= / operator/= on doubles)
	for (u64 b = (u64)tid * 4; b < N; b += (u64)blockDim.x * 4) {
		u64 s0 = 0, s1 = 0, s2 = 0, s3 = 0;
		for (u64 p = p0; p < p1; p++) {
			if (!kept(p)) {
				continue;
			}
			const T *row = bins + a.members[p] * N + b;
			s0 += row[0];
			s1 += row[1];
			s2 += row[2];
			s3 += row[3];
		}
		mean[b] = (double)s0 / (double)n_good;
		mean[b + 1] = (double)s1 / (double)n_good;
		mean[b + 2] = (double)s2 / (double)n_good;
		mean[b + 3] = (double)s3 / (double)n_good;
	}
	// distance_d of every survivor to the mean, one thread per member (see distance_d_kernel), first minimum
	constexpr bool NARROW = sizeof(T) <= 2;
	double bd = 0;
	long long bi = -1;
	for (u64 q0 = p0; q0 < p1; q0 += blockDim.x) {
		const u64 p = q0 + tid;
		const bool on = p < p1 && kept(p);
		const T *row = on ? bins + a.members[p] * N : bins;
		u64 d = 0, mag = 0;
		double magd = 0;
		for (u64 base = 0; base < N; base += 512) {
			const u64 cn = min((u64)512, N - base);
			__syncthreads(); // also orders the mean written above before its first read
			for (u64 t = tid; t < cn; t += blockDim.x) {
				const double ci = mean[base + t];
				c[t] = ci;
				rc[t] = (T)round(ci);
			}
			__syncthreads();
			if (on) {
				if (NARROW && cn % 16 == 0) {
					constexpr int PER = 16 / (int)sizeof(T);
					for (u64 t = 0; t < cn; t += PER) {
						const uint4 v = *reinterpret_cast<const uint4 *>(row + base + t);
						const T *pv = reinterpret_cast<const T *>(&v);
#pragma unroll
						for (int u = 0; u < PER; u++) {
							const T pp = pv[u];
							const T r = rc[t + u];
							d += 2 * (u64)(pp < r ? pp : r);
							magd = floor(magd + ((double)pp + c[t + u]));
						}
					}
				} else {
					for (u64 t = 0; t < cn; t++) {
						const T pp = row[base + t];
						const double ci = c[t];
						const T r = rc[t];
						d += 2 * (pp < r ? pp : r);
						if (NARROW) {
							magd = floor(magd + ((double)pp + ci));
						} else {
							mag = (u64)((double)mag + ((double)pp + ci));
						}
					}
				}
			}
		}
		if (on) {
			if (NARROW) {
				mag = (u64)magd;
			}
			const double frac = (double)d / (double)mag;
			const double dist = 10000.0 * (1.0 - frac * frac);
			if (bi < 0 || dist < bd) {
				bd = dist;
				bi = (long long)(p - p0);
			}
		}
	}
	auto better = [&](double d2, long long i2) { return i2 >= 0 && (bi < 0 || d2 < bd || (d2 == bd && i2 < bi)); };
	for (int s = 16; s > 0; s >>= 1) {
		const double d2 = __shfl_xor_sync(0xffffffffu, bd, s);
		const long long i2 = __shfl_xor_sync(0xffffffffu, bi, s);
		if (better(d2, i2)) {
			bd = d2;
			bi = i2;
		}
	}
	if (lane == 0) {
		s_d[warp] = bd;
		s_i[warp] = bi;
	}
	__syncthreads();
	if (tid == 0) {
		for (int w = 1; w < 8; w++) {
			if (better(s_d[w], s_i[w])) {
				bd = s_d[w];
				bi = s_i[w];
			}
		}
		a.next[cidx] = bi;
		a.n_good[cidx] = n_good;
	}
}

int launch_update_batch(mc2_ctx *ctx, const mc2_hset *h, const u64 *d_member_off, const u64 *d_members, const uint8_t *d_close,
			const uint8_t *d_skipped, u64 c_begin, u64 count, double *d_mean, long long *d_next, u64 *d_n_good)
{
	if (count == 0) {
		return MC2_OK;
	}
	UpdateArgs a{h->bins, h->N, d_member_off, d_members, d_close, d_skipped, c_begin, d_mean, d_next, d_n_good};
	prof_begin(ctx, 6);
	switch (h->eb) {
	case 1: update_batch_kernel<uint8_t><<<(unsigned)count, 256, 0, ctx->stream>>>(a); break;
	case 2: update_batch_kernel<uint16_t><<<(unsigned)count, 256, 0, ctx->stream>>>(a); break;
	case 4: update_batch_kernel<uint32_t><<<(unsigned)count, 256, 0, ctx->stream>>>(a); break;
	case 8: update_batch_kernel<u64><<<(unsigned)count, 256, 0, ctx->stream>>>(a); break;
	default: set_error("elem_bytes must be 1, 2, 4 or 8"); return MC2_ERR_ARG;
	}
	prof_end(ctx);
	ctx->launches++;
	MC2_CUDA(cudaGetLastError());
	return MC2_OK;
}

// Batched merge stage: Trainer<T>::merge (src/cluster/Trainer.cpp:74-109) for every center of one pass of
// ClusterFactory.cpp:382-401.  Segment c holds the scored pairs (center c+1.., center c); out[c] = position of the chosen
// candidate (sequential semantics: a later tie wins, candidates whose first-combo value is not above DBL_MIN never win) or -1.
__global__ void __launch_bounds__(128) merge_batch_kernel(const double *__restrict__ dist, const uint8_t *__restrict__ skipped,
							     const uint8_t *__restrict__ close, const u64 *__restrict__ off, u64 n_centers,
							     long long *__restrict__ out)
{
	const u64 cix = (u64)blockIdx.x * blockDim.x + threadIdx.x;
	if (cix >= n_centers) {
		return;
	}
	double bd = 2.2250738585072014e-308;
	long long bi = -1;
	for (u64 p = off[cix]; p < off[cix + 1]; p++) {
		if (skipped[p] || !close[p]) {
			continue;
		}
		const double d = dist[p];
		if (!(bd > d)) {
			bd = d;
			bi = (long long)(p - off[cix]);
		}
	}
	out[cix] = bi;
}

int launch_merge_batch(mc2_ctx *ctx, const double *d_dist, const uint8_t *d_skipped, const uint8_t *d_close, const u64 *d_off,
		       u64 n_centers, long long *d_out)
{
	if (n_centers == 0) {
		return MC2_OK;
	}
	prof_begin(ctx, 6);
	merge_batch_kernel<<<(unsigned)((n_centers + 127) / 128), 128, 0, ctx->stream>>>(d_dist, d_skipped, d_close, d_off, n_centers, d_out);
	prof_end(ctx);
	ctx->launches++;
	MC2_CUDA(cudaGetLastError());
	return MC2_OK;
}

} // namespace mc2
