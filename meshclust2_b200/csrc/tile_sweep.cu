// tile_sweep.cu — K2, all-pairs tile form: query x database sweep over 1 KiB uint8 histograms (k = 5) on sm_100a.
//
// fastcar's work() for a whole block (src/fastcar/FC_Runner.cpp:427-470) / the all-pairs sweep of BASELINE configs[2]:
// every (query, database) pair inside the length window gets the model's reductions, the GLM score and the cutoff
// (Feature.cpp:136-171, Trainer.cpp:112-120, Predictor.cpp:323-333), survivors are appended to a list.
//
// One persistent CTA per SM walks 64 (query) x 128 (database) pair tiles.  Per tile the 1024 bins are streamed through a
// shared-memory ring in 16 chunks of 64 bins by TMA (cp.async.bulk.tensor, hardware swizzle), and three engines consume
// each chunk:
//   S_pq  = sum p*q              tcgen05.mma kind::i8 (u8 x u8 -> s32, exact), accumulator 128 x 64 in TMEM, issued by one
//                                thread; the accumulator is double buffered so the next tile's MMAs overlap the epilogue.
//   S_emd = sum |cumP - cumQ|    CUDA cores on precomputed u16 cumulative rows (mc2_hset::cum16), through the identity
//                                sum|a-b| = sum a + sum b - 2 sum min(a,b): VIMNMX.U16x2 (2 bins / instruction), packed
//                                16-bit partial sums added three at a time (IADD3) and flushed into a 32-bit accumulator
//                                with IDP.2A before a half can overflow.  Thread tile 8 query x 4 database rows: the
//                                query rows are warp-uniform shared-memory broadcasts, the database rows conflict-free
//                                128-bit loads from the swizzled tile.
//   S_sad = sum |p - q|          VABSDIFF4.U8.ACC on the u8 tile (4 bins / instruction)  -> S_min = (sumP+sumQ-S_sad)/2
// Epilogue per tile: reductions -> shared memory; one pair per thread at a time: length window (FC_Runner.cpp:435-444),
// an fp32 evaluation of the GLM sum with a running error bound that can only REJECT (sum + bound < -1e-6 => not close
// whatever the rounding); everything else goes through the exact fp64 epilogue shared with the other pair kernels
// (eval_pair_fast), so scores and decisions are the same bits as theirs.
#include "mc2_internal.cuh"
#include "pair_eval.cuh"
#include <cuda.h>
#include <cstring>
#include <cstdlib>

namespace mc2 {

namespace ts {

constexpr int TQ = 64;          // query rows per tile (UMMA N, TMEM columns)
constexpr int TD = 128;         // database rows per tile (UMMA M, TMEM lanes)
constexpr int KC = 64;          // bins per pipeline stage
constexpr int NBINS = 1024;
constexpr int NCHUNK = NBINS / KC;
constexpr int NCW = 8;          // compute warps
constexpr int THREADS = (NCW + 2) * 32;
constexpr int CUMD_BYTES = TD * KC * 2;  // 16 KB, 128-byte rows, SWIZZLE_128B
constexpr int CUMQ_BYTES = TQ * KC * 2;  //  8 KB
constexpr int U8D_BYTES = TD * KC;       //  8 KB, 64-byte rows, SWIZZLE_64B
constexpr int U8Q_BYTES = TQ * KC;       //  4 KB
constexpr int STAGE_BYTES = CUMD_BYTES + CUMQ_BYTES + U8D_BYTES + U8Q_BYTES; // 36 KB (a stage keeps this layout whatever NEED is)
constexpr int RED_BYTES = TQ * TD * 4;   // one staged reduction (32 KB)
constexpr int MAX_SUPER = 1024;          // entries of the tile schedule's prefix array

__host__ __device__ constexpr int n_red(int need) { return ((need & NEED_DOT) ? 1 : 0) + ((need & NEED_EMD) ? 1 : 0) + ((need & NEED_MIN) ? 1 : 0); }
__host__ __device__ constexpr int n_stages(int need)
{
	// 227 KB per CTA: staged reductions + candidate list (16 KB) + row info / schedule / barriers (16 KB) + ring
	const int fixed = n_red(need) * RED_BYTES + 16384 + 16384 + 1024;
	const int s = (227 * 1024 - fixed) / STAGE_BYTES;
	return s > 4 ? 4 : s;
}

struct Params {
	u64 q0, q1, d0, d1;     // row ranges (query set / database set)
	int upper_only;
	double cutoff;
	u64 max_out;
	u64 *out_q, *out_d;
	double *out_score;
	u64 *counters;          // [0] survivors, [1] scored pairs
	int *err;
	Sideband sbQ, sbD;
	const u32 *csQ, *csD;   // per-row sum of the cumulative row (EMD identity)
	u32 nqt, ndt;           // tiles along each side
	u32 group;              // query tiles per super-row of the schedule
	u32 n_super;
	const u32 *sched;       // [n_super + 1] exclusive prefix of items per super-row (device)
	int no_screen;          // experiments: skip the screen and the exact path (main-loop cost only)
	int flush_mode;         // 0: packed sums flushed every 16 words (row sums <= 4095), 1: every 4 (<= 16383), 2: IDP.2A per word
	// raw mode (tests): dense (q1-q0) x (d1-d0) matrices of the reductions instead of scoring
	u32 *raw_dot, *raw_emd, *raw_sad;
};

// ---------------------------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ u32 smem_u32(const void *p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(u32 bar, u32 count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(u32 bar, u32 bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(u32 bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ bool mbar_try(u32 bar, u32 parity)
{
	u32 ok;
	asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
		     : "=r"(ok)
		     : "r"(bar), "r"(parity)
		     : "memory");
	return ok != 0;
}
// bounded wait: a pipeline bug must end in an error, never in a hung GPU
__device__ __forceinline__ void mbar_wait(u32 bar, u32 parity, int *err)
{
	u32 spins = 0;
	while (!mbar_try(bar, parity)) {
		if (++spins > (1u << 24)) {
			atomicOr(err, 4);
			__trap();
		}
	}
}
__device__ __forceinline__ void tma_load_2d(u32 dst, const CUtensorMap *map, int x, int y, u32 bar)
{
	asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
		     "l"(map), "r"(x), "r"(y), "r"(bar)
		     : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(u32 bar)
{
	asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_i8(u32 tmem_d, u64 desc_a, u64 desc_b, u32 idesc, u32 accumulate)
{
	asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
		     "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
		     : "memory");
}
__device__ __forceinline__ void named_sync(int id, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }
__device__ __forceinline__ uint4 lds128(u32 addr)
{
	uint4 v;
	asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
	return v;
}
__device__ __forceinline__ u32 vmin2(u32 a, u32 b)
{
	u32 d;
	asm("min.u16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
	return d;
}
__device__ __forceinline__ u32 dp2a_sum(u32 packed, u32 acc)
{
	u32 d;
	asm("dp2a.lo.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(packed), "r"(0x0101u), "r"(acc));
	return d;
}
__device__ __forceinline__ u32 sad4(u32 a, u32 b, u32 c)
{
	u32 d;
	asm("vabsdiff4.u32.u32.u32.add %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
	return d;
}

// K-major shared-memory operand descriptor for 64-byte rows under SWIZZLE_64B (cute::UMMA::SmemDescriptor: start >> 4 at
// [0,14), leading byte offset >> 4 at [16,30) (1 for swizzled K-major), stride byte offset >> 4 at [32,46) = 8 rows x 64 B,
// version 1 at [46,48), layout type at [61,64): 4 = SWIZZLE_64B)
__device__ __forceinline__ u64 umma_desc_sw64(u32 saddr)
{
	return (u64)((saddr & 0x3FFFFu) >> 4) | ((u64)1 << 16) | ((u64)(512 >> 4) << 32) | ((u64)1 << 46) | ((u64)4 << 61);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): c_format S32 = 2 at [4,6), a/b format U8 = 0, K-major both,
// N >> 3 at [17,23), M >> 4 at [24,29)
constexpr u32 IDESC_U8 = (2u << 4) | ((u32)(TQ >> 3) << 17) | ((u32)(TD >> 4) << 24);

// ---------------------------------------------------------------------------------------------------------------
// tile schedule: super-rows of `group` query tiles; inside a super-row the items run database-tile major, query-tile
// minor, so the CTAs working at the same time share one database tile and a handful of query tiles (L2 reuse).
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ u32 first_dt(const Params &p, u32 qt)
{
	if (!p.upper_only) {
		return 0;
	}
	// smallest database offset that can pair with the tile's first query row: d > q
	const s64 x = (s64)(p.q0 + (u64)qt * TQ) + 1 - (s64)p.d0;
	return x <= 0 ? 0u : (u32)(x / TD);
}

__global__ void __launch_bounds__(1024) sched_kernel(Params p, u32 *sched)
{
	__shared__ u32 cnt[MAX_SUPER];
	const u32 s = threadIdx.x;
	u32 c = 0;
	if (s < p.n_super) {
		const u32 f = first_dt(p, s * p.group);
		c = f < p.ndt ? (p.ndt - f) * p.group : 0;
	}
	cnt[s] = c;
	__syncthreads();
	if (s == 0) {
		u32 run = 0;
		for (u32 i = 0; i < p.n_super; i++) {
			sched[i] = run;
			run += cnt[i];
		}
		sched[p.n_super] = run;
	}
}

struct Tile {
	u32 qt, dt;
	bool valid;
};
__device__ __forceinline__ Tile decode_item(const Params &p, const u32 *s_sched, u32 item)
{
	// largest s with sched[s] <= item
	u32 lo = 0, hi = p.n_super;
	while (hi - lo > 1) {
		const u32 mid = (lo + hi) >> 1;
		if (s_sched[mid] <= item) {
			lo = mid;
		} else {
			hi = mid;
		}
	}
	const u32 r = item - s_sched[lo];
	Tile t;
	t.qt = lo * p.group + r % p.group;
	t.dt = first_dt(p, lo * p.group) + r / p.group;
	t.valid = t.qt < p.nqt && t.dt < p.ndt && t.dt >= first_dt(p, t.qt);
	return t;
}

// ---------------------------------------------------------------------------------------------------------------
// per-chunk CUDA-core work of one compute warp: rows cw*8 .. cw*8+7 of the query tile x rows lane + 32 j of the database tile
// ---------------------------------------------------------------------------------------------------------------
template <int NEED, int FLUSH>
__device__ __forceinline__ void chunk_compute(u32 stage, int cw, int lane, u32 (&emd)[8][4], u32 (&sad)[8][4])
{
	if constexpr ((NEED & NEED_EMD) != 0) {
		// 128-byte rows, SWIZZLE_128B: 16-byte chunk c of row r sits at r*128 + ((c ^ (r & 7)) << 4)
		const u32 dbase = stage + (u32)lane * 128;
		const u32 dsw = ((u32)lane & 7) << 4;
		const u32 qbase = stage + CUMD_BYTES + (u32)cw * 1024;
#pragma unroll
		for (int g = 0; g < 2; g++) {
			u32 acc[8][4];
#pragma unroll
			for (int ks = 0; ks < 4; ks++) {
				const int c = g * 4 + ks;
				uint4 D[4];
#pragma unroll
				for (int j = 0; j < 4; j++) {
					D[j] = lds128(dbase + (u32)j * 4096 + (((u32)c << 4) ^ dsw));
				}
#pragma unroll
				for (int i = 0; i < 8; i++) {
					const uint4 Q = lds128(qbase + (u32)i * 128 + (u32)((c ^ i) << 4));
#pragma unroll
					for (int j = 0; j < 4; j++) {
						const u32 m0 = vmin2(Q.x, D[j].x), m1 = vmin2(Q.y, D[j].y);
						const u32 m2 = vmin2(Q.z, D[j].z), m3 = vmin2(Q.w, D[j].w);
						if constexpr (FLUSH == 2) {
							emd[i][j] = dp2a_sum(m0, emd[i][j]);
							emd[i][j] = dp2a_sum(m1, emd[i][j]);
							emd[i][j] = dp2a_sum(m2, emd[i][j]);
							emd[i][j] = dp2a_sum(m3, emd[i][j]);
						} else if constexpr (FLUSH == 1) {
							emd[i][j] = dp2a_sum(m0 + m1 + m2 + m3, emd[i][j]);
						} else {
							if (ks == 0) {
								acc[i][j] = m0 + m1 + m2 + m3;
							} else {
								acc[i][j] += m0 + m1;
								acc[i][j] += m2 + m3;
							}
						}
					}
				}
			}
			if constexpr (FLUSH == 0) {
#pragma unroll
				for (int i = 0; i < 8; i++) {
#pragma unroll
					for (int j = 0; j < 4; j++) {
						emd[i][j] = dp2a_sum(acc[i][j], emd[i][j]);
					}
				}
			}
		}
	}
	if constexpr ((NEED & NEED_MIN) != 0) {
		// 64-byte rows, SWIZZLE_64B: 16-byte chunk c of row r sits at r*64 + ((c ^ ((r >> 1) & 3)) << 4)
		const u32 dbase = stage + CUMD_BYTES + CUMQ_BYTES + (u32)lane * 64;
		const u32 dsw = (((u32)lane >> 1) & 3) << 4;
		const u32 qbase = stage + CUMD_BYTES + CUMQ_BYTES + U8D_BYTES + (u32)cw * 512;
#pragma unroll
		for (int c = 0; c < 4; c++) {
			uint4 D[4];
#pragma unroll
			for (int j = 0; j < 4; j++) {
				D[j] = lds128(dbase + (u32)j * 2048 + (((u32)c << 4) ^ dsw));
			}
#pragma unroll
			for (int i = 0; i < 8; i++) {
				const uint4 Q = lds128(qbase + (u32)i * 64 + (u32)((c ^ ((i >> 1) & 3)) << 4));
#pragma unroll
				for (int j = 0; j < 4; j++) {
					u32 s = sad[i][j];
					s = sad4(Q.x, D[j].x, s);
					s = sad4(Q.y, D[j].y, s);
					s = sad4(Q.z, D[j].z, s);
					s = sad4(Q.w, D[j].w, s);
					sad[i][j] = s;
				}
			}
		}
	}
}

// ---------------------------------------------------------------------------------------------------------------
// fp32 screen: the GLM sum with a running bound on |computed - exact|.  Returns true when the pair may be close (or the
// screen cannot tell): only `false` is a decision.  eps is 2^-23 (twice the unit round-off); every raw single below is a
// cancellation-free expression of exact integers, at most 8 operations of <= 2 ulp each -> relative bound 16 eps.
// ---------------------------------------------------------------------------------------------------------------
struct RowF {          // per-row values staged in shared memory once per tile
	u32 mag, sum;  // low words (valid when !big)
	float magf, sumsqf, lenf, nnf; // nnf = N*sumsq - 2*mag*sum + mag^2 (pearson's centred norm, exact integer -> float)
	u32 sumsq;
	u32 big;       // mag or len outside the screen's comfortable range -> exact path
};

__device__ __forceinline__ bool screen_pair(const DevModel &dm, u32 dot, u32 emd, u32 sad, const RowF &P, const RowF &Q)
{
	const float eps = 1.1920929e-07f;
	float x[MC2_MAX_SINGLES], ex[MC2_MAX_SINGLES];
#define MC2_SCR(CODE, RAW, DELTA)                                                                  \
	{                                                                                          \
		const float raw_ = (RAW);                                                          \
		const float cmin_ = (float)dm.cmin[CODE], rcp_ = (float)dm.crcp[CODE];             \
		const float t_ = raw_ - cmin_;                                                     \
		const float v_ = t_ * rcp_;                                                        \
		const float xv_ = dm.csim[CODE] ? v_ : 1.0f - v_;                                  \
		x[dm.slot[CODE]] = xv_;                                                            \
		ex[dm.slot[CODE]] = eps * (((DELTA)*fabsf(raw_) + fabsf(cmin_) + fabsf(t_)) * fabsf(rcp_) + 2.0f * fabsf(v_) + fabsf(xv_)); \
	}
	const u32 n2 = P.sumsq + Q.sumsq - 2u * dot; // exact: sum (p-q)^2 < 2^27 for 1024 uint8 bins
	if (dm.slot[SC_MANHATTAN] >= 0) {
		MC2_SCR(SC_MANHATTAN, (float)sad, 1.0f);
	}
	if (dm.slot[SC_EUCLIDEAN] >= 0) {
		MC2_SCR(SC_EUCLIDEAN, sqrtf((float)n2), 16.0f);
	}
	if (dm.slot[SC_SIMRATIO] >= 0) {
		const float d = (float)dot;
		MC2_SCR(SC_SIMRATIO, __fdividef(d, d + sqrtf((float)n2)), 16.0f);
	}
	if (dm.slot[SC_NORMALIZED_VECTORS] >= 0) {
		MC2_SCR(SC_NORMALIZED_VECTORS, (float)dot * rsqrtf(P.sumsqf * Q.sumsqf), 16.0f);
	}
	if (dm.slot[SC_PEARSON] >= 0) {
		const long long ndot = 1024ll * (long long)dot - (long long)P.mag * (long long)Q.sum - (long long)Q.mag * (long long)P.sum +
				       (long long)P.mag * (long long)Q.mag;
		MC2_SCR(SC_PEARSON, (float)ndot * rsqrtf(P.nnf * Q.nnf), 16.0f);
	}
	if (dm.slot[SC_INTERSECTION] >= 0) {
		MC2_SCR(SC_INTERSECTION, __fdividef((float)(P.sum + Q.sum - sad), P.magf + Q.magf), 16.0f);
	}
	if (dm.slot[SC_EMD] >= 0) {
		MC2_SCR(SC_EMD, (float)emd, 1.0f);
	}
	if (dm.slot[SC_LENGTHD] >= 0) {
		MC2_SCR(SC_LENGTHD, fabsf(P.lenf - Q.lenf), 4.0f);
	}
	if (dm.slot[SC_KULCZYNSKI2] >= 0) {
		const float ap = P.magf * (1.0f / 1024.0f), aq = Q.magf * (1.0f / 1024.0f);
		const float smin = 0.5f * (float)(P.sum + Q.sum - sad);
		MC2_SCR(SC_KULCZYNSKI2, __fdividef(1024.0f * (ap + aq), 2.0f * ap * aq) * smin, 16.0f);
	}
#undef MC2_SCR
	float s = (float)dm.weight[0];
	float M = fabsf(s), E = 0.0f;
#pragma unroll 1
	for (int c = 0; c < dm.n_combos; c++) {
		const int *ix = dm.idx[c];
		const int kind = dm.kind[c];
		float v = 1.0f, hi = 1.0f, lo = 1.0f; // product, product of (|f| + e), product of |f|
		const int n = dm.nidx[c];
		for (int t = 0; t < n; t++) {
			const float f = x[ix[t]], a = fabsf(f), b = a + ex[ix[t]];
			int pw = 1;
			if (kind == MC2_COMBO_X2Y2 || (kind == MC2_COMBO_XY2 && t == 1) || (kind == MC2_COMBO_X2Y && t == 0)) {
				pw = 2;
			}
			v *= f;
			hi *= b;
			lo *= a;
			if (pw == 2) {
				v *= f;
				hi *= b;
				lo *= a;
			}
		}
		const float w = (float)dm.weight[c + 1], aw = fabsf(w);
		s += w * v;
		E += aw * ((hi - lo) + 8.0f * eps * hi);
		M += aw * hi;
	}
	E += 16.0f * eps * M;
	if (!(M < 1.0e6f)) {
		return true;
	}
	return !(s + E < -1.0e-6f);
}

// ---------------------------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------------------------
template <int NEED> struct Smem {
	static constexpr int STAGES = n_stages(NEED);
	static constexpr int RING = 0;
	static constexpr int RED = STAGES * STAGE_BYTES;                 // staged reductions: dot, emd, sad (those needed)
	static constexpr int CAND = RED + n_red(NEED) * RED_BYTES;       // u16 pair indices, 8192 entries
	static constexpr int ROWD = CAND + 16384;                        // RowF[TD]
	static constexpr int ROWQ = ROWD + TD * (int)sizeof(RowF);       // RowF[TQ]
	static constexpr int WINQ = ROWQ + TQ * (int)sizeof(RowF);       // u64 window [TQ][2]
	static constexpr int LEND = WINQ + TQ * 16;                      // u64 len [TD]
	static constexpr int SCHED = LEND + TD * 8;                      // u32 [MAX_SUPER + 1]
	static constexpr int BARS = SCHED + (MAX_SUPER + 1) * 4 + 4;     // mbarriers
	static constexpr int MISC = BARS + 16 * 8;                       // tmem base, candidate count
	static constexpr int TOTAL = MISC + 64 + 1024; // + alignment slack
};

template <int NEED, int FLUSH, bool RAW>
__global__ void __launch_bounds__(THREADS, 1)
tile_sweep_kernel(const __grid_constant__ DevModel dm, const __grid_constant__ Params p, const __grid_constant__ CUtensorMap mapCumD,
		  const __grid_constant__ CUtensorMap mapCumQ, const __grid_constant__ CUtensorMap mapU8D,
		  const __grid_constant__ CUtensorMap mapU8Q)
{
	using L = Smem<NEED>;
	constexpr int STAGES = L::STAGES;
	constexpr bool DOT = (NEED & NEED_DOT) != 0, EMD = (NEED & NEED_EMD) != 0, MIN = (NEED & NEED_MIN) != 0;
	constexpr bool CUDA_STAGE = EMD || MIN;           // compute warps read the ring
	constexpr bool U8_STAGE = DOT || MIN;             // the ring carries the u8 tiles
	constexpr u32 TX_BYTES = (EMD ? CUMD_BYTES + CUMQ_BYTES : 0) + (U8_STAGE ? U8D_BYTES + U8Q_BYTES : 0);
	extern __shared__ unsigned char smem_raw[];
	// SWIZZLE_128B tiles and the UMMA descriptors want 1024-byte alignment: align by hand (the launch adds the slack)
	unsigned char *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
	const u32 sbase = smem_u32(smem);
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	u32 *s_sched = reinterpret_cast<u32 *>(smem + L::SCHED);
	const u32 bar_full = sbase + L::BARS, bar_empty = bar_full + 4 * 8, bar_tfull = bar_empty + 4 * 8, bar_tempty = bar_tfull + 2 * 8;
	u32 *s_tmem = reinterpret_cast<u32 *>(smem + L::MISC);
	u32 *s_ncand = s_tmem + 1;

	for (u32 i = threadIdx.x; i <= p.n_super; i += blockDim.x) {
		s_sched[i] = p.sched[i];
	}
	if (threadIdx.x == 0) {
		for (int s = 0; s < STAGES; s++) {
			mbar_init(bar_full + s * 8, 1);
			mbar_init(bar_empty + s * 8, (CUDA_STAGE ? NCW : 0) + (DOT ? 1 : 0));
		}
		for (int b = 0; b < 2; b++) {
			mbar_init(bar_tfull + b * 8, 1);
			mbar_init(bar_tempty + b * 8, NCW);
		}
		*s_ncand = 0;
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	if (DOT && warp == 1) { // TMEM: two 64-column accumulators
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(s_tmem)) : "memory");
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
	}
	tc_fence_before();
	__syncthreads();
	tc_fence_after();
	const u32 tmem_base = DOT ? *s_tmem : 0;
	const u32 n_items = s_sched[p.n_super];

	if (warp == 0) {
		// ===== TMA producer =====
		if (lane == 0) {
			u32 s = 0, ph = 0;
			for (u32 item = blockIdx.x; item < n_items; item += gridDim.x) {
				const Tile t = decode_item(p, s_sched, item);
				if (!t.valid) {
					continue;
				}
				const int qrow = (int)(p.q0 + (u64)t.qt * TQ), drow = (int)(p.d0 + (u64)t.dt * TD);
				for (int c = 0; c < NCHUNK; c++) {
					mbar_wait(bar_empty + s * 8, ph ^ 1, p.err);
					const u32 st = sbase + s * STAGE_BYTES, fb = bar_full + s * 8;
					mbar_expect_tx(fb, TX_BYTES);
					if (EMD) {
						tma_load_2d(st, &mapCumD, c * KC, drow, fb);
						tma_load_2d(st + CUMD_BYTES, &mapCumQ, c * KC, qrow, fb);
					}
					if (U8_STAGE) {
						tma_load_2d(st + CUMD_BYTES + CUMQ_BYTES, &mapU8D, c * KC, drow, fb);
						tma_load_2d(st + CUMD_BYTES + CUMQ_BYTES + U8D_BYTES, &mapU8Q, c * KC, qrow, fb);
					}
					if (++s == (u32)STAGES) {
						s = 0;
						ph ^= 1;
					}
				}
			}
		}
	} else if (warp == 1) {
		// ===== MMA issuer =====
		if (DOT && lane == 0) {
			u32 s = 0, ph = 0, it = 0;
			for (u32 item = blockIdx.x; item < n_items; item += gridDim.x) {
				const Tile t = decode_item(p, s_sched, item);
				if (!t.valid) {
					continue;
				}
				const u32 buf = it & 1;
				mbar_wait(bar_tempty + buf * 8, ((it >> 1) & 1) ^ 1, p.err);
				tc_fence_after();
				const u32 tacc = tmem_base + buf * TQ;
				for (int c = 0; c < NCHUNK; c++) {
					mbar_wait(bar_full + s * 8, ph, p.err);
					tc_fence_after();
					const u32 a = sbase + s * STAGE_BYTES + CUMD_BYTES + CUMQ_BYTES, b = a + U8D_BYTES;
#pragma unroll
					for (int k = 0; k < KC / 32; k++) {
						tc_mma_i8(tacc, umma_desc_sw64(a + k * 32), umma_desc_sw64(b + k * 32), IDESC_U8, (c | k) != 0);
					}
					tc_commit(bar_empty + s * 8);
					if (++s == (u32)STAGES) {
						s = 0;
						ph ^= 1;
					}
				}
				tc_commit(bar_tfull + buf * 8);
				it++;
			}
		}
	} else {
		// ===== compute warps =====
		const int cw = warp - 2;
		const int ctid = threadIdx.x - 64; // 0..255
		u32 s = 0, ph = 0, it = 0;
		RowF *s_rowD = reinterpret_cast<RowF *>(smem + L::ROWD);
		RowF *s_rowQ = reinterpret_cast<RowF *>(smem + L::ROWQ);
		u64 *s_winQ = reinterpret_cast<u64 *>(smem + L::WINQ);
		u64 *s_lenD = reinterpret_cast<u64 *>(smem + L::LEND);
		unsigned short *s_cand = reinterpret_cast<unsigned short *>(smem + L::CAND);
		u32 *s_dot = reinterpret_cast<u32 *>(smem + L::RED);
		u32 *s_emd = s_dot + (DOT ? TQ * TD : 0);
		u32 *s_sad = s_emd + (EMD ? TQ * TD : 0);
		for (u32 item = blockIdx.x; item < n_items; item += gridDim.x) {
			const Tile t = decode_item(p, s_sched, item);
			if (!t.valid) {
				continue;
			}
			const u64 qrow0 = p.q0 + (u64)t.qt * TQ, drow0 = p.d0 + (u64)t.dt * TD;
			// row info for the epilogue: issued now, consumed after the main loop
			if (!RAW && ctid < TQ + TD) {
				const bool isq = ctid >= TD;
				const u64 row = isq ? qrow0 + (ctid - TD) : drow0 + ctid;
				const bool ok = isq ? row < p.q1 : row < p.d1;
				const Sideband &sb = isq ? p.sbQ : p.sbD;
				RowF r;
				u64 len = 0;
				memset(&r, 0, sizeof r);
				r.big = 1;
				if (ok) {
					const u64 mag = sb.mag[row], sum = sb.sum[row], sumsq = sb.sumsq[row];
					len = sb.len[row];
					r.mag = (u32)mag;
					r.sum = (u32)sum;
					r.sumsq = (u32)sumsq;
					r.magf = (float)mag;
					r.sumsqf = (float)sumsq;
					r.lenf = (float)len;
					// the screen wants exact float lengths (their difference cancels) and 64-bit pearson terms
					r.big = (mag >= (1ull << 26) || len >= (1ull << 24) || len == 0 || mag == 0) ? 1u : 0u;
					const long long nn = 1024ll * (long long)sumsq - 2ll * (long long)(mag & 0x3FFFFFFull) * (long long)sum +
							     (long long)(mag & 0x3FFFFFFull) * (long long)(mag & 0x3FFFFFFull);
					r.nnf = (float)nn;
				}
				if (isq) {
					s_rowQ[ctid - TD] = r;
					// FC_Runner.cpp:435-444: size_t truncation of len * id and len / id; an empty window marks an unused row
					s_winQ[2 * (ctid - TD)] = ok ? (u64)((double)len * p.cutoff) : 1;
					s_winQ[2 * (ctid - TD) + 1] = ok ? (u64)((double)len / p.cutoff) : 0;
				} else {
					s_rowD[ctid] = r;
					s_lenD[ctid] = ok ? len : ~0ull;
				}
			}
			u32 emd[8][4], sad[8][4];
#pragma unroll
			for (int i = 0; i < 8; i++) {
#pragma unroll
				for (int j = 0; j < 4; j++) {
					emd[i][j] = 0;
					sad[i][j] = 0;
				}
			}
			if (CUDA_STAGE) {
#pragma unroll 1
				for (int c = 0; c < NCHUNK; c++) {
					mbar_wait(bar_full + s * 8, ph, p.err);
					chunk_compute<NEED, FLUSH>(sbase + s * STAGE_BYTES, cw, lane, emd, sad);
					__syncwarp();
					if (lane == 0) {
						mbar_arrive(bar_empty + s * 8);
					}
					if (++s == (u32)STAGES) {
						s = 0;
						ph ^= 1;
					}
				}
			}
			// ---- epilogue: reductions -> shared memory ----
			if (EMD || MIN) {
#pragma unroll
				for (int i = 0; i < 8; i++) {
#pragma unroll
					for (int j = 0; j < 4; j++) {
						const int idx = (cw * 8 + i) * TD + lane + 32 * j;
						if (EMD) s_emd[idx] = emd[i][j];
						if (MIN) s_sad[idx] = sad[i][j];
					}
				}
			}
			if (DOT) {
				const u32 buf = it & 1;
				mbar_wait(bar_tfull + buf * 8, (it >> 1) & 1, p.err);
				tc_fence_after();
				// warp w may read TMEM lanes 32 (w % 4) ..; compute warps 0-3 take accumulator columns 0-31, 4-7 columns 32-63
				const u32 quarter = (u32)warp & 3, chalf = (u32)cw >> 2;
				u32 v[32];
				const u32 taddr = tmem_base + buf * TQ + chalf * 32 + ((quarter * 32) << 16);
				asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
					     "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
					     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
					       "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
					       "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
					       "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
					     : "r"(taddr)
					     : "memory");
				asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
				for (int c = 0; c < 32; c++) {
					s_dot[(chalf * 32 + c) * TD + quarter * 32 + lane] = v[c];
				}
				tc_fence_before();
				__syncwarp();
				if (lane == 0) {
					mbar_arrive(bar_tempty + buf * 8);
				}
			}
			named_sync(1, NCW * 32);
			// ---- epilogue: one pair per thread at a time ----
			u32 scored = 0;
			for (int pi = ctid; pi < TQ * TD; pi += NCW * 32) {
				const int ql = pi >> 7, dl = pi & (TD - 1);
				const u64 q = qrow0 + ql, d = drow0 + dl;
				if constexpr (RAW) {
					if (q < p.q1 && d < p.d1) {
						const u64 o = (q - p.q0) * (p.d1 - p.d0) + (d - p.d0);
						if (DOT) p.raw_dot[o] = s_dot[pi];
						if (EMD) p.raw_emd[o] = p.csQ[q] + p.csD[d] - 2u * s_emd[pi];
						if (MIN) p.raw_sad[o] = s_sad[pi];
					}
				} else {
					const u64 lc = s_lenD[dl];
					bool go = lc >= s_winQ[2 * ql] && lc <= s_winQ[2 * ql + 1] && (!p.upper_only || d > q);
					if (go) {
						scored++;
						const RowF &P = s_rowD[dl], &Q = s_rowQ[ql];
						bool cand = !p.no_screen;
						if (cand && !(P.big | Q.big)) {
							const u32 dotv = DOT ? s_dot[pi] : 0;
							const u32 emdv = EMD ? p.csQ[q] + p.csD[d] - 2u * s_emd[pi] : 0;
							cand = screen_pair(dm, dotv, emdv, MIN ? s_sad[pi] : 0, P, Q);
						}
						if (cand) {
							const u32 slot = atomicAdd(s_ncand, 1u);
							s_cand[slot] = (unsigned short)pi;
						}
					}
				}
			}
			if constexpr (!RAW) {
				named_sync(1, NCW * 32);
				const u32 ncand = *s_ncand;
				for (u32 base = 0; base < ncand; base += NCW * 32) {
					const u32 e = base + (u32)ctid;
					int close = 0;
					double score = 0, d0v;
					u64 q = 0, d = 0;
					if (e < ncand) {
						const int pi = s_cand[e];
						const int ql = pi >> 7, dl = pi & (TD - 1);
						q = qrow0 + ql;
						d = drow0 + dl;
						const Side sd = load_side(p.sbD, d), sq = load_side(p.sbQ, q);
						RedN r;
						r.jeff = r.js = 0;
						r.dot = DOT ? s_dot[pi] : 0;
						r.emd = EMD ? (u64)(p.csQ[q] + p.csD[d] - 2u * s_emd[pi]) : 0;
						r.smin = MIN ? (sd.sum + sq.sum - (u64)s_sad[pi]) >> 1 : 0;
						const int bad = eval_pair_fast(dm, NBINS, r, sd, sq, true, score, d0v, close);
						if (bad) {
							atomicOr(p.err, bad & 1 ? 1 : 2);
						}
					}
					const unsigned cm = __ballot_sync(0xffffffffu, close);
					u64 obase = 0;
					if (lane == 0 && cm) {
						obase = atomicAdd(p.counters, (u64)__popc(cm));
					}
					obase = __shfl_sync(0xffffffffu, obase, 0);
					if (close) {
						const u64 idx = obase + __popc(cm & ((1u << lane) - 1));
						if (idx < p.max_out) {
							p.out_q[idx] = q;
							p.out_d[idx] = d;
							p.out_score[idx] = score;
						}
					}
				}
				scored = __reduce_add_sync(0xffffffffu, scored);
				if (lane == 0 && scored) {
					atomicAdd(p.counters + 1, (u64)scored);
				}
			}
			named_sync(1, NCW * 32); // staged reductions, candidate list and row info are free again
			if (ctid == 0) {
				*s_ncand = 0;
			}
			it++;
		}
	}
	tc_fence_before();
	__syncthreads();
	if (DOT && warp == 1) {
		asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem_base) : "memory");
	}
}

// cumulative rows: warp per row, lane l owns bins [32 l, 32 l + 32); inclusive prefix, u16 (row sums < 65536), plus the
// row's sum of prefixes for the EMD identity
__global__ void __launch_bounds__(256) cum16_kernel(const unsigned char *__restrict__ bins, u64 n, unsigned short *__restrict__ cum,
						    u32 *__restrict__ cumsum)
{
	const int lane = threadIdx.x & 31;
	const u64 warps_total = (u64)gridDim.x * (blockDim.x >> 5);
	for (u64 r = (u64)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < n; r += warps_total) {
		const uint4 *src = reinterpret_cast<const uint4 *>(bins + r * 1024 + lane * 32);
		const uint4 a = __ldg(src), b = __ldg(src + 1);
		const u32 w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
		u32 tot = 0;
#pragma unroll
		for (int i = 0; i < 8; i++) {
			tot = __dp4a(w[i], 0x01010101u, tot);
		}
		u32 x = tot;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			const u32 y = __shfl_up_sync(0xffffffffu, x, d);
			if (lane >= d) {
				x += y;
			}
		}
		u32 run = x - tot, cs = 0;
		u32 out[16];
#pragma unroll
		for (int i = 0; i < 8; i++) {
			u32 c0 = run + (w[i] & 0xFF);
			u32 c1 = c0 + ((w[i] >> 8) & 0xFF);
			u32 c2 = c1 + ((w[i] >> 16) & 0xFF);
			u32 c3 = c2 + (w[i] >> 24);
			run = c3;
			cs += c0 + c1 + c2 + c3;
			out[2 * i] = c0 | (c1 << 16);
			out[2 * i + 1] = c2 | (c3 << 16);
		}
		uint4 *dst = reinterpret_cast<uint4 *>(cum + r * 1024 + lane * 32);
		dst[0] = make_uint4(out[0], out[1], out[2], out[3]);
		dst[1] = make_uint4(out[4], out[5], out[6], out[7]);
		dst[2] = make_uint4(out[8], out[9], out[10], out[11]);
		dst[3] = make_uint4(out[12], out[13], out[14], out[15]);
		cs = __reduce_add_sync(0xffffffffu, cs);
		if (lane == 0) {
			cumsum[r] = cs;
		}
	}
}

} // namespace ts

int ensure_cum16(mc2_ctx *ctx, const mc2_hset *hc)
{
	mc2_hset *h = const_cast<mc2_hset *>(hc);
	if (h->eb != 1 || h->N != 1024 || h->max_sum >= 65536 || h->n == 0) {
		h->cum16_valid = 0;
		return MC2_OK;
	}
	if (h->cum16_valid) {
		return MC2_OK;
	}
	if (!h->cum16) {
		MC2_CUDA(cudaMalloc((void **)&h->cum16, h->n * 2048));
		MC2_CUDA(cudaMalloc((void **)&h->cumsum, h->n * 4));
	}
	u64 want = (h->n + 7) / 8, cap = (u64)ctx->sm_count * 8;
	int grid = (int)(want < cap ? want : cap);
	prof_begin(ctx, 5);
	ts::cum16_kernel<<<grid, 256, 0, ctx->stream>>>((const unsigned char *)h->bins, h->n, h->cum16, h->cumsum);
	prof_end(ctx);
	ctx->launches++;
	MC2_CUDA(cudaGetLastError());
	h->cum16_valid = 1;
	return MC2_OK;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
				  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
				  CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn()
{
	static EncodeTiledFn fn = nullptr;
	static bool tried = false;
	if (!tried) {
		tried = true;
		void *p = nullptr;
		cudaDriverEntryPointQueryResult qr;
		if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) == cudaSuccess && qr == cudaDriverEntryPointSuccess) {
			fn = reinterpret_cast<EncodeTiledFn>(p);
		}
	}
	return fn;
}

// rows of 1024 elements (u8 or u16), box = 64 elements x `box_rows` rows
static int make_map(CUtensorMap *m, const void *base, u64 n_rows, int elem_bytes, int box_rows)
{
	EncodeTiledFn fn = encode_fn();
	if (!fn) {
		set_error("tile sweep: cuTensorMapEncodeTiled is not available from this driver");
		return MC2_ERR_CUDA;
	}
	const cuuint64_t dims[2] = {1024, n_rows};
	const cuuint64_t strides[1] = {(cuuint64_t)1024 * elem_bytes};
	const cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
	const cuuint32_t es[2] = {1, 1};
	const CUresult r = fn(m, elem_bytes == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, const_cast<void *>(base), dims,
			      strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, elem_bytes == 1 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
			      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
	if (r != CUDA_SUCCESS) {
		set_error("tile sweep: cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
		return MC2_ERR_CUDA;
	}
	return MC2_OK;
}

bool tile_sweep_supported(const DevModel &dm, const mc2_hset *q, const mc2_hset *d)
{
	static const bool off = getenv("MC2_SWEEP_LEGACY") != nullptr;
	if (off) {
		return false;
	}
	const bool shape = q->eb == 1 && d->eb == 1 && q->N == 1024 && d->N == 1024 && q->max_sum < 65536 && d->max_sum < 65536;
	const bool model = dm.fast_epi && !dm.regression && dm.bias == 0.0 && !(dm.need & NEED_LOG) && (dm.need & 7) != 0;
	return shape && model && q->n < (1ull << 31) && d->n < (1ull << 31);
}

template <int NEED, bool RAW> static int launch_need(int flush, int grid, cudaStream_t st, const DevModel &dm, const ts::Params &p, const CUtensorMap *m)
{
	const int smem = ts::Smem<NEED>::TOTAL;
#define MC2_TS_GO(F)                                                                                                           \
	{                                                                                                                      \
		MC2_CUDA(cudaFuncSetAttribute(ts::tile_sweep_kernel<NEED, F, RAW>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); \
		ts::tile_sweep_kernel<NEED, F, RAW><<<grid, ts::THREADS, smem, st>>>(dm, p, m[0], m[1], m[2], m[3]);                \
	}
	if constexpr ((NEED & NEED_EMD) != 0) {
		if (flush == 0) MC2_TS_GO(0)
		else if (flush == 1) MC2_TS_GO(1)
		else MC2_TS_GO(2)
	} else {
		MC2_TS_GO(0)
	}
#undef MC2_TS_GO
	return MC2_OK;
}

int launch_tile_sweep(mc2_ctx *ctx, const DevModel &dm, int need, const mc2_hset *q, u64 q0, u64 q1, const mc2_hset *d, u64 d0, u64 d1,
		      int upper_only, double cutoff, u64 max_out, u64 *d_out_q, u64 *d_out_d, double *d_out_score, u64 *d_counters,
		      u32 *raw_dot, u32 *raw_emd, u32 *raw_sad)
{
	if (!ctx->d_sched) {
		MC2_CUDA(cudaMalloc(&ctx->d_sched, (ts::MAX_SUPER + 1) * 4));
	}
	u32 *d_sched = (u32 *)ctx->d_sched;
	const bool raw = raw_dot || raw_emd || raw_sad;
	int rc;
	if (need & NEED_EMD) {
		rc = ensure_cum16(ctx, q);
		if (rc != MC2_OK) return rc;
		rc = ensure_cum16(ctx, d);
		if (rc != MC2_OK) return rc;
		if (!q->cum16_valid || !d->cum16_valid) {
			set_error("tile sweep: cumulative rows unavailable for this set");
			return MC2_ERR_UNSUPPORTED;
		}
	}
	ts::Params p;
	memset(&p, 0, sizeof p);
	p.q0 = q0; p.q1 = q1; p.d0 = d0; p.d1 = d1;
	p.upper_only = upper_only;
	p.cutoff = cutoff;
	p.max_out = max_out;
	p.out_q = d_out_q; p.out_d = d_out_d; p.out_score = d_out_score;
	p.counters = d_counters;
	p.err = ctx->d_err;
	p.sbQ = Sideband{q->mag, q->sum, q->sumsq, q->len};
	p.sbD = Sideband{d->mag, d->sum, d->sumsq, d->len};
	p.csQ = q->cumsum; p.csD = d->cumsum;
	p.nqt = (u32)((q1 - q0 + ts::TQ - 1) / ts::TQ);
	p.ndt = (u32)((d1 - d0 + ts::TD - 1) / ts::TD);
	u32 group = 8;
	while ((p.nqt + group - 1) / group > (u32)ts::MAX_SUPER) {
		group *= 2;
	}
	p.group = group;
	p.n_super = (p.nqt + group - 1) / group;
	p.sched = d_sched;
	const u64 ms = q->max_sum > d->max_sum ? q->max_sum : d->max_sum;
	p.flush_mode = ms <= 4095 ? 0 : (ms <= 16383 ? 1 : 2);
	p.no_screen = getenv("MC2_TS_NOSCREEN") != nullptr;
	if (const char *e = getenv("MC2_TS_FLUSH")) { // experiments: force a (legal) more frequent flush
		const int f = atoi(e);
		if (f > p.flush_mode && f <= 2) p.flush_mode = f;
	}
	p.raw_dot = raw_dot; p.raw_emd = raw_emd; p.raw_sad = raw_sad;
	// the kernel reads the item count from the last prefix entry; it must fit 32 bits
	if ((u64)p.n_super * group * p.ndt >= (1ull << 32)) {
		set_error("tile sweep: more than 2^32 tiles in one call; split the query range");
		return MC2_ERR_ARG;
	}
	ts::sched_kernel<<<1, 1024, 0, ctx->stream>>>(p, d_sched);
	ctx->launches++;
	CUtensorMap maps[4];
	memset(maps, 0, sizeof maps);
	if (need & NEED_EMD) {
		rc = make_map(&maps[0], d->cum16, d->n, 2, ts::TD);
		if (rc != MC2_OK) return rc;
		rc = make_map(&maps[1], q->cum16, q->n, 2, ts::TQ);
		if (rc != MC2_OK) return rc;
	}
	if (need & (NEED_DOT | NEED_MIN)) {
		rc = make_map(&maps[2], d->bins, d->n, 1, ts::TD);
		if (rc != MC2_OK) return rc;
		rc = make_map(&maps[3], q->bins, q->n, 1, ts::TQ);
		if (rc != MC2_OK) return rc;
	}
	const int grid = ctx->sm_count;
	prof_begin(ctx, 3);
	rc = MC2_OK;
	switch (need & 7) {
#define MC2_TS_CASE(n)                                                                                  \
	case n:                                                                                         \
		rc = raw ? launch_need<n, true>(p.flush_mode, grid, ctx->stream, dm, p, maps)           \
			 : launch_need<n, false>(p.flush_mode, grid, ctx->stream, dm, p, maps);         \
		break;
		MC2_TS_CASE(1)
		MC2_TS_CASE(2)
		MC2_TS_CASE(3)
		MC2_TS_CASE(4)
		MC2_TS_CASE(5)
		MC2_TS_CASE(6)
		MC2_TS_CASE(7)
#undef MC2_TS_CASE
	default:
		set_error("tile sweep: the model needs no reduction");
		rc = MC2_ERR_UNSUPPORTED;
	}
	prof_end(ctx);
	if (rc != MC2_OK) return rc;
	ctx->launches++;
	MC2_CUDA(cudaGetLastError());
	return MC2_OK;
}

} // namespace mc2
