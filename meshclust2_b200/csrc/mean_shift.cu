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

} // namespace mc2
