// ubench.cu — instruction-throughput microbenchmarks for the integer ops the pair kernel is built from (sm_100a).
// Prints warp-instructions per clock per SM for each op at full occupancy (8 CTAs x 256 threads per SM).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench tools/ubench.cu ; run under gpurun.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned int u32;
#define ITERS 4096
#define CHAINS 8

template <int OP> __device__ __forceinline__ u32 op(u32 a, u32 b, u32 c)
{
	u32 d;
	if (OP == 0) asm volatile("vabsdiff4.u32.u32.u32.add %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
	else if (OP == 1) asm volatile("dp4a.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
	else if (OP == 2) asm volatile("vabsdiff.s32.s32.s32.add %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
	else if (OP == 3) asm volatile("prmt.b32 %0, %1, %2, 0x5140;" : "=r"(d) : "r"(c), "r"(b));
	else if (OP == 4) asm volatile("add.u32 %0, %1, %2;" : "=r"(d) : "r"(c), "r"(b));
	else if (OP == 5) asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
	else if (OP == 6) asm volatile("dp2a.lo.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
	else if (OP == 7) asm volatile("min.u16x2 %0, %1, %2;" : "=r"(d) : "r"(c), "r"(b));
	else if (OP == 8) asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
	else if (OP == 9) asm volatile("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
	else if (OP == 10) { d = __reduce_add_sync(0xffffffffu, c) + a; }
	else if (OP == 11) { d = __shfl_up_sync(0xffffffffu, c, 1) + a; }
	else if (OP == 12) asm volatile("abs.s32 %0, %1;" : "=r"(d) : "r"(c));
	else if (OP == 13) asm volatile("shf.l.wrap.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(c), "r"(a), "r"(b));
	else if (OP == 14) asm volatile("popc.b32 %0, %1;" : "=r"(d) : "r"(c));
	else d = c;
	return d;
}

template <int OP> __global__ void __launch_bounds__(256) k_op(u32 *out, long long *cyc, u32 seed)
{
	u32 a = threadIdx.x * 2654435761u + seed, b = a ^ 0x5bd1e995u;
	u32 c[CHAINS];
#pragma unroll
	for (int i = 0; i < CHAINS; i++) c[i] = a + i;
	long long t0 = clock64();
	for (int it = 0; it < ITERS; it++) {
#pragma unroll
		for (int i = 0; i < CHAINS; i++) c[i] = op<OP>(a, b, c[i]);
	}
	long long t1 = clock64();
	u32 s = 0;
#pragma unroll
	for (int i = 0; i < CHAINS; i++) s += c[i];
	out[blockIdx.x * blockDim.x + threadIdx.x] = s;
	if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// mixed: the u8 slab body (12 instr / 4 bins) with independent words, to see the achievable blend
__global__ void __launch_bounds__(256) k_mix(u32 *out, long long *cyc, u32 seed)
{
	u32 p = threadIdx.x * 2654435761u + seed, q = p ^ 0x5bd1e995u;
	u32 sad = 0, dot = 0, e = 0;
	int c = 0;
	long long t0 = clock64();
	for (int it = 0; it < ITERS; it++) {
#pragma unroll
		for (int w = 0; w < 4; w++) {
			u32 pp = p + w * it, qq = q ^ (w + it);
			asm volatile("vabsdiff4.u32.u32.u32.add %0, %1, %2, %3;" : "=r"(sad) : "r"(pp), "r"(qq), "r"(sad));
			asm volatile("dp4a.u32.u32 %0, %1, %2, %3;" : "=r"(dot) : "r"(pp), "r"(qq), "r"(dot));
			u32 w01, w23;
			asm volatile("prmt.b32 %0, %1, %2, 0x5140;" : "=r"(w01) : "r"(pp), "r"(qq));
			asm volatile("prmt.b32 %0, %1, %2, 0x7362;" : "=r"(w23) : "r"(pp), "r"(qq));
			int l0, l1, l2, l3;
			asm volatile("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(l0) : "r"(w01), "r"(0x0000FF01u), "r"(c));
			asm volatile("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(l1) : "r"(w01), "r"(0xFF01FF01u), "r"(c));
			asm volatile("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(l2) : "r"(w23), "r"(0x0000FF01u), "r"(l1));
			asm volatile("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(l3) : "r"(w23), "r"(0xFF01FF01u), "r"(l1));
			c = l3;
			asm volatile("vabsdiff.s32.s32.s32.add %0, %1, %2, %3;" : "=r"(e) : "r"(l0), "r"(7), "r"(e));
			asm volatile("vabsdiff.s32.s32.s32.add %0, %1, %2, %3;" : "=r"(e) : "r"(l1), "r"(7), "r"(e));
			asm volatile("vabsdiff.s32.s32.s32.add %0, %1, %2, %3;" : "=r"(e) : "r"(l2), "r"(7), "r"(e));
			asm volatile("vabsdiff.s32.s32.s32.add %0, %1, %2, %3;" : "=r"(e) : "r"(l3), "r"(7), "r"(e));
		}
	}
	long long t1 = clock64();
	out[blockIdx.x * blockDim.x + threadIdx.x] = sad + dot + e + c;
	if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// shared-memory atomics: random bins in a 1024-entry u32 histogram per warp (the K1 inner op)
template <int MODE> __global__ void __launch_bounds__(256) k_atoms(u32 *out, long long *cyc, u32 seed)
{
	__shared__ u32 h[8 * 1024];
	for (int i = threadIdx.x; i < 8 * 1024; i += 256) h[i] = 0;
	__syncthreads();
	u32 *my = h + (threadIdx.x >> 5) * 1024;
	u32 x = threadIdx.x * 2654435761u + seed + blockIdx.x;
	long long t0 = clock64();
	for (int it = 0; it < ITERS; it++) {
		x = x * 1664525u + 1013904223u;
		u32 idx = MODE == 0 ? (x >> 22) : (MODE == 1 ? (it & 1023) : ((x >> 22) & ~31u) | (threadIdx.x & 31));
		if (MODE == 3) { u32 v = my[idx & 1023]; my[idx & 1023] = v + 1; }
		else atomicAdd(&my[idx], 1u);
	}
	long long t1 = clock64();
	__syncthreads();
	out[blockIdx.x * blockDim.x + threadIdx.x] = h[threadIdx.x];
	if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <typename F> void run(const char *name, F launch, int per_iter)
{
	int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
	int grid = sms * 8;
	u32 *out; long long *cyc;
	cudaMalloc(&out, (size_t)grid * 256 * 4); cudaMalloc(&cyc, grid * 8);
	launch(grid, out, cyc);
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	cudaEventRecord(e0);
	launch(grid, out, cyc);
	cudaEventRecord(e1); cudaEventSynchronize(e1);
	float ms; cudaEventElapsedTime(&ms, e0, e1);
	long long *h = new long long[grid];
	cudaMemcpy(h, cyc, grid * 8, cudaMemcpyDeviceToHost);
	double avg = 0; for (int i = 0; i < grid; i++) avg += h[i]; avg /= grid;
	double winstr_per_sm = 8.0 * 8 /*warps*/ * (double)ITERS * per_iter;
	printf("%-28s %8.3f ms  cycles/CTA %10.0f  warp-instr/clk/SM %6.3f  (clock ~%.0f MHz)  err=%s\n", name, ms, avg,
	       winstr_per_sm / avg, avg / (ms * 1e3), cudaGetErrorString(cudaGetLastError()));
	delete[] h; cudaFree(out); cudaFree(cyc);
}

#define RUN_OP(n, label) run(label, [](int g, u32 *o, long long *c) { k_op<n><<<g, 256>>>(o, c, 1); }, CHAINS)

int main()
{
	cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
	printf("device %s sm_%d%d SMs=%d\n", p.name, p.major, p.minor, p.multiProcessorCount);
	RUN_OP(0, "VABSDIFF4.U8.ACC");
	RUN_OP(1, "IDP.4A.U8.U8");
	RUN_OP(9, "IDP.4A.U8.S8");
	RUN_OP(2, "VABSDIFF (scalar, acc)");
	RUN_OP(3, "PRMT");
	RUN_OP(4, "IADD");
	RUN_OP(5, "IMAD");
	RUN_OP(6, "IDP.2A");
	RUN_OP(7, "VIMNMX.U16x2");
	RUN_OP(8, "LOP3");
	RUN_OP(12, "IABS");
	RUN_OP(13, "SHF (funnel)");
	RUN_OP(14, "POPC");
	RUN_OP(10, "REDUX.SUM (+IADD)");
	RUN_OP(11, "SHFL.UP (+IADD)");
	run("u8 slab mix (12 instr/4 bins)", [](int g, u32 *o, long long *c) { k_mix<<<g, 256>>>(o, c, 1); }, 4 * 12 + 8);
	run("ATOMS random 1024 bins", [](int g, u32 *o, long long *c) { k_atoms<0><<<g, 256>>>(o, c, 1); }, 1);
	run("ATOMS same bin per warp", [](int g, u32 *o, long long *c) { k_atoms<1><<<g, 256>>>(o, c, 1); }, 1);
	run("ATOMS conflict-free", [](int g, u32 *o, long long *c) { k_atoms<2><<<g, 256>>>(o, c, 1); }, 1);
	run("LDS+STS rmw random", [](int g, u32 *o, long long *c) { k_atoms<3><<<g, 256>>>(o, c, 1); }, 1);
	return 0;
}
