// ubench2.cu — issue-rate microbenchmarks for the instruction mixes of the tile sweep (sm_100a), round 2.
// Each kernel runs long enough (several ms) for the clocks to settle; rates are reported per second and per SM clock,
// with the clock taken from clock64() deltas over the event time (both printed so a wrong clock assumption shows).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench2 tools/ubench2.cu ; run under gpurun.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned int u32;
#define CH 16

__device__ __forceinline__ u32 vmin2(u32 a, u32 b) { u32 d; asm volatile("min.u16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
__device__ __forceinline__ u32 add3(u32 a, u32 b, u32 c) { u32 d; asm volatile("{.reg .u32 t; add.u32 t, %1, %2; add.u32 %0, t, %3;}" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
__device__ __forceinline__ u32 dp2a(u32 a, u32 b, u32 c) { u32 d; asm volatile("dp2a.lo.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
__device__ __forceinline__ u32 dp4a(u32 a, u32 b, u32 c) { u32 d; asm volatile("dp4a.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
__device__ __forceinline__ u32 vsad4(u32 a, u32 b, u32 c) { u32 d; asm volatile("vabsdiff4.u32.u32.u32.add %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
__device__ __forceinline__ u32 lop(u32 a, u32 b, u32 c) { u32 d; asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }

// MODE: 0 VIMNMX.U16x2 only | 1 IADD3 only | 2 (2 VIMNMX + 1 IADD3) | 3 (VIMNMX + IDP.2A) | 4 IDP.2A only | 5 LOP3 only
//       6 FFMA only | 7 DFMA only | 8 (2 VIMNMX + IADD3) with an LDS.128 per 24 | 9 VABSDIFF4 only | 10 IDP.4A only
//       11 (VIMNMX + LOP3) same-pipe pair | 12 (IDP.2A + FFMA) | 13 (VIMNMX + FFMA)
template <int MODE> __global__ void __launch_bounds__(256) k(u32 *out, long long *cyc, u32 seed, int iters)
{
	__shared__ uint4 sm[256];
	sm[threadIdx.x] = make_uint4(threadIdx.x, seed, 3, 4);
	__syncthreads();
	u32 a = threadIdx.x * 2654435761u + seed, b = a ^ 0x5bd1e995u;
	u32 c[CH];
	float f[CH];
	double d[CH / 2];
#pragma unroll
	for (int i = 0; i < CH; i++) { c[i] = a + i; f[i] = (float)i; }
#pragma unroll
	for (int i = 0; i < CH / 2; i++) d[i] = (double)i;
	long long t0 = clock64();
	for (int it = 0; it < iters; it++) {
		if (MODE == 8) {
			uint4 v = sm[(threadIdx.x + it) & 255];
			a ^= v.x; b += v.y;
		}
#pragma unroll
		for (int i = 0; i < CH; i++) {
			if (MODE == 0) { if (i & 1) c[i] = vmin2(c[i], b); else asm volatile("max.u16x2 %0, %0, %1;" : "+r"(c[i]) : "r"(a)); }
			else if (MODE == 1) c[i] = add3(c[i], a, b);
			else if (MODE == 2 || MODE == 8) { u32 m1 = vmin2(a, c[i]); u32 m2 = vmin2(b, c[i]); c[i] = add3(c[i], m1, m2); }
			else if (MODE == 3) { u32 m1 = vmin2(a, c[i]); c[i] = dp2a(m1, 0x0101u, c[i]); }
			else if (MODE == 4) c[i] = dp2a(a, b, c[i]);
			else if (MODE == 5) c[i] = lop(a, b, c[i]);
			else if (MODE == 6) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(f[i]) : "f"(1.0001f), "f"(0.5f));
			else if (MODE == 7) { if (i < CH / 2) asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(d[i]) : "d"(1.0001), "d"(0.5)); }
			else if (MODE == 9) c[i] = vsad4(a, b, c[i]);
			else if (MODE == 10) c[i] = dp4a(a, b, c[i]);
			else if (MODE == 11) { u32 m1 = vmin2(a, c[i]); c[i] = lop(m1, b, c[i]); }
			else if (MODE == 12) { c[i] = dp2a(a, b, c[i]); asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(f[i]) : "f"(1.0001f), "f"(0.5f)); }
			else if (MODE == 13) { c[i] = vmin2(c[i], b); asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(f[i]) : "f"(1.0001f), "f"(0.5f)); }
		}
	}
	long long t1 = clock64();
	u32 s = 0;
#pragma unroll
	for (int i = 0; i < CH; i++) s += c[i] + (u32)f[i];
#pragma unroll
	for (int i = 0; i < CH / 2; i++) s += (u32)d[i];
	out[blockIdx.x * blockDim.x + threadIdx.x] = s + a + b;
	if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE> void run(const char *name, double instr_per_iter, int iters)
{
	int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
	const int ctas_per_sm = 4; // 32 warps / SM
	int grid = sms * ctas_per_sm;
	u32 *out; long long *cyc;
	cudaMalloc(&out, (size_t)grid * 256 * 4); cudaMalloc(&cyc, grid * 8);
	k<MODE><<<grid, 256>>>(out, cyc, 1, iters);
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	cudaEventRecord(e0);
	k<MODE><<<grid, 256>>>(out, cyc, 1, iters);
	cudaEventRecord(e1); cudaEventSynchronize(e1);
	float ms; cudaEventElapsedTime(&ms, e0, e1);
	long long *h = new long long[grid];
	cudaMemcpy(h, cyc, grid * 8, cudaMemcpyDeviceToHost);
	double avg = 0; for (int i = 0; i < grid; i++) avg += h[i]; avg /= grid;
	double winstr_per_sm = (double)ctas_per_sm * 8 * (double)iters * instr_per_iter;
	printf("%-40s %8.3f ms  clock64/CTA %10.0f (%.0f MHz)  warp-instr/clk64/SM %6.3f  warp-instr/ns/SM %6.3f  err=%s\n", name, ms, avg,
	       avg / (ms * 1e3), winstr_per_sm / avg, winstr_per_sm / (ms * 1e6), cudaGetErrorString(cudaGetLastError()));
	delete[] h; cudaFree(out); cudaFree(cyc);
}

int main()
{
	cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
	printf("device %s sm_%d%d SMs=%d clockRate=%d kHz\n", p.name, p.major, p.minor, p.multiProcessorCount, p.clockRate);
	const int IT = 40000;
	run<6>("FFMA", CH, IT);
	run<0>("VIMNMX.U16x2", CH, IT);
	run<1>("IADD3 (2 adds fused?)", CH, IT);
	run<2>("2 VIMNMX + IADD3", CH * 3, IT);
	run<3>("VIMNMX + IDP.2A", CH * 2, IT);
	run<4>("IDP.2A", CH, IT);
	run<5>("LOP3", CH, IT);
	run<7>("DFMA (8 chains)", CH / 2, IT);
	run<8>("2 VIMNMX + IADD3 + LDS.128/iter", CH * 3 + 3, IT);
	run<9>("VABSDIFF4.ACC", CH, IT);
	run<10>("IDP.4A", CH, IT);
	run<11>("VIMNMX + LOP3", CH * 2, IT);
	run<12>("IDP.2A + FFMA", CH * 2, IT);
	run<13>("VIMNMX + FFMA", CH * 2, IT);
	return 0;
}
