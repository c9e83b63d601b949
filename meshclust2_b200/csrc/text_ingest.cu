// text_ingest.cu — raw nucleotide text -> (segments, 2-bit packed bases) on the device.
//
// The reference does this on the host, three passes over every sequence with a std::map lookup per base:
//   Chromosome::help            src/nonltr/Chromosome.cpp:130-154   upper-case, then
//   removeAmbiguous             :263-291   maximal runs of non-N (a run opening on the very last base is never closed)
//   mergeSegments               :298-353   only if the sequence is longer than 20: bridge gaps (next.s - cur.e < 10), then
//                                          drop merged segments shorter than 20
//   makeSegmentList             :355-385   cut segments longer than 1 Mbp into 1 Mbp pieces, the last takes the remainder
//   ChromosomeOneDigit::encode  src/nonltr/ChromosomeOneDigit.cpp:79-133 with the DNA code map
//                               (src/nonltr/ChromosomeOneDigitDna.cpp:48-68): IUPAC letters fold onto A,C,G,T (N -> C inside
//                               a bridged gap); any other byte throws when the sequence has at least one segment.
// Here: one warp per sequence.  The warp ballots "is N" over 32 bases at a time and every lane replays the same tiny state
// machine over the run boundaries (the merge / drop / split rules are sequential, but they act on runs, not on bases);
// a counting pass, an exclusive scan, a writing pass; then one thread per 16-base word maps letters to codes through a
// shared-memory table and packs them.  K1 consumes exactly this (codes, segments) contract.
#include "mc2_internal.cuh"

namespace mc2 {

__device__ __forceinline__ unsigned char up_char(unsigned char c)
{
	return (c >= 'a' && c <= 'z') ? (unsigned char)(c - 32) : c;
}

// code of an (upper-cased) letter, -1 if it is not a nucleotide letter
__device__ __forceinline__ int code_of(unsigned char u)
{
	switch (u) {
	case 'A': case 'M': case 'V': return 0;
	case 'C': case 'Y': case 'H': case 'N': return 1;
	case 'G': case 'R': case 'S': case 'X': return 2;
	case 'T': case 'K': case 'W': case 'B': case 'D': return 3;
	}
	return -1;
}

// WRITE = false: count[seq] = number of final segments; WRITE = true: fill segs[2 * (seg_off[seq] + j)] (inclusive pairs)
template <bool WRITE>
__global__ void __launch_bounds__(256) segment_kernel(const char *__restrict__ text, const u64 *__restrict__ seq_off, u64 n,
						       u32 *__restrict__ count, const u64 *__restrict__ seg_off, int *__restrict__ segs,
						       unsigned long long *min_seg)
{
	const int lane = threadIdx.x & 31;
	const u64 warps_total = (u64)gridDim.x * (blockDim.x >> 5);
	for (u64 s = (u64)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); s < n; s += warps_total) {
		const unsigned char *src = reinterpret_cast<const unsigned char *>(text) + seq_off[s];
		const int L = (int)(seq_off[s + 1] - seq_off[s]);
		u32 nseg = 0;
		u64 out0 = WRITE ? seg_off[s] : 0;
		int shortest = 0x7fffffff;
		auto emit = [&](int a, int b) {
			if (WRITE && lane == 0) {
				segs[2 * (out0 + nseg)] = a;
				segs[2 * (out0 + nseg) + 1] = b;
			}
			shortest = min(shortest, b - a + 1);
			nseg++;
		};
		auto flush = [&](int a, int b) { // makeSegmentList
			const int len = b - a + 1;
			if (len > 1000000) {
				const int nf = len / 1000000;
				for (int h = 0; h < nf; h++) {
					const int fs = a + h * 1000000;
					emit(fs, h == nf - 1 ? b : fs + 999999);
				}
			} else {
				emit(a, b);
			}
		};
		bool in_run = false, have_cur = false;
		int run_s = 0, cur_s = 0, cur_e = 0;
		auto on_run = [&](int a, int b) {
			if (a == L - 1) {
				return; // removeAmbiguous never closes a run that opens on the last base
			}
			if (L > 20) { // mergeSegments
				if (have_cur && a - cur_e < 10) {
					cur_e = b;
				} else {
					if (have_cur && cur_e - cur_s + 1 >= 20) {
						flush(cur_s, cur_e);
					}
					cur_s = a;
					cur_e = b;
					have_cur = true;
				}
			} else {
				flush(a, b);
			}
		};
		for (int base = 0; base < L; base += 32) {
			const int j = base + lane;
			const bool is_n = j >= L || up_char(src[j]) == 'N'; // past the end counts as N: it closes an open run
			const unsigned non = ~__ballot_sync(0xffffffffu, is_n);
			int pos = 0;
			while (pos < 32) {
				const unsigned from = 0xffffffffu << pos;
				if (!in_run) {
					const unsigned rem = non & from;
					if (!rem) {
						break;
					}
					const int b = __ffs(rem) - 1;
					in_run = true;
					run_s = base + b;
					pos = b + 1;
				} else {
					const unsigned rem = ~non & from;
					if (!rem) {
						break;
					}
					const int b = __ffs(rem) - 1;
					in_run = false;
					on_run(run_s, base + b - 1);
					pos = b + 1;
				}
			}
		}
		if (in_run) { // L is a multiple of 32 and the last base is not N
			on_run(run_s, L - 1);
		}
		if (L > 20 && have_cur && cur_e - cur_s + 1 >= 20) {
			flush(cur_s, cur_e);
		}
		if (lane == 0) {
			if (!WRITE) {
				count[s] = nseg;
			} else if (nseg) {
				atomicMin(min_seg, (unsigned long long)shortest);
			}
		}
	}
}

// exclusive scan of count[0..n) into off[0..n] (off[n] = total); one CTA, contiguous chunk per thread
__global__ void __launch_bounds__(1024) seg_scan_kernel(const u32 *__restrict__ count, u64 n, u64 *__restrict__ off)
{
	__shared__ u64 part[1024];
	const u64 chunk = (n + blockDim.x - 1) / blockDim.x;
	const u64 b = (u64)threadIdx.x * chunk, e = b + chunk < n ? b + chunk : n;
	u64 sum = 0;
	for (u64 i = b; i < e; i++) {
		sum += count[i];
	}
	part[threadIdx.x] = sum;
	__syncthreads();
	if (threadIdx.x == 0) {
		u64 run = 0;
		for (int t = 0; t < (int)blockDim.x; t++) {
			u64 v = part[t];
			part[t] = run;
			run += v;
		}
		off[n] = run;
	}
	__syncthreads();
	u64 run = part[threadIdx.x];
	for (u64 i = b; i < e; i++) {
		off[i] = run;
		run += count[i];
	}
}

// letters -> codes -> 2 bits/base (same layout as pack_kernel); a byte that is no nucleotide letter is an error when the
// sequence has at least one segment (ChromosomeOneDigit::encode walks the whole sequence then), ignored otherwise
__global__ void __launch_bounds__(128) pack_text_kernel(const char *__restrict__ text, const u64 *__restrict__ seq_off,
							 const u64 *__restrict__ word_off, const u64 *__restrict__ seg_off, u64 n,
							 u32 *__restrict__ packed, int *err)
{
	__shared__ signed char lut[256];
	for (int c = threadIdx.x; c < 256; c += blockDim.x) {
		lut[c] = (signed char)code_of(up_char((unsigned char)c));
	}
	__syncthreads();
	for (u64 s = blockIdx.x; s < n; s += gridDim.x) {
		const u64 b0 = seq_off[s];
		const u64 len = seq_off[s + 1] - b0;
		const u64 w0 = word_off[s];
		const u64 nw = word_off[s + 1] - w0;
		const bool have = seg_off[s + 1] > seg_off[s];
		const unsigned char *src = reinterpret_cast<const unsigned char *>(text) + b0;
		for (u64 w = threadIdx.x; w < nw; w += blockDim.x) {
			u32 word = 0;
			int bad = 0;
			const u64 j0 = w * 16;
#pragma unroll
			for (int t = 0; t < 16; t++) {
				const u64 j = j0 + t;
				const int c = j < len ? (int)lut[src[j]] : 0;
				bad |= c < 0;
				word |= ((u32)c & 3u) << (30 - 2 * t);
			}
			packed[w0 + w] = word;
			if (bad && have) {
				atomicOr(err, 4);
			}
		}
	}
}

int launch_segment(mc2_ctx *ctx, bool write, const char *d_text, const u64 *d_seq_off, u64 n, u32 *d_count, const u64 *d_seg_off,
		   int *d_segs, unsigned long long *d_min_seg)
{
	if (n == 0) {
		return MC2_OK;
	}
	u64 want = (n + 7) / 8, cap = (u64)ctx->sm_count * 8;
	int grid = (int)(want < cap ? want : cap);
	prof_begin(ctx, 0);
	if (write) {
		segment_kernel<true><<<grid, 256, 0, ctx->stream>>>(d_text, d_seq_off, n, d_count, d_seg_off, d_segs, d_min_seg);
	} else {
		segment_kernel<false><<<grid, 256, 0, ctx->stream>>>(d_text, d_seq_off, n, d_count, d_seg_off, d_segs, d_min_seg);
	}
	prof_end(ctx);
	ctx->launches++;
	MC2_CUDA(cudaGetLastError());
	return MC2_OK;
}

int launch_seg_scan(mc2_ctx *ctx, const u32 *d_count, u64 n, u64 *d_seg_off)
{
	seg_scan_kernel<<<1, 1024, 0, ctx->stream>>>(d_count, n, d_seg_off);
	ctx->launches++;
	MC2_CUDA(cudaGetLastError());
	return MC2_OK;
}

int launch_pack_text(mc2_ctx *ctx, const char *d_text, const u64 *d_seq_off, mc2_seqs *s)
{
	if (s->n == 0) {
		return MC2_OK;
	}
	u64 cap = (u64)ctx->sm_count * 16;
	int grid = (int)(s->n < cap ? s->n : cap);
	prof_begin(ctx, 0);
	pack_text_kernel<<<grid, 128, 0, ctx->stream>>>(d_text, d_seq_off, s->word_off, s->seg_off, s->n, s->packed, ctx->d_err);
	prof_end(ctx);
	ctx->launches++;
	MC2_CUDA(cudaGetLastError());
	return MC2_OK;
}

} // namespace mc2
