// ubench_int.cu -- issue-rate microbenchmark of the integer instructions the DP kernels are built from (sm_100a).
// Each kernel runs 8 independent dependency chains per thread; the printed figure is warp-instructions per clock per SM
// sub-partition (4 per SM), measured with CUDA events against the SM clock reported by the driver at run time.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_int tools/ubench_int.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>

#define CHAINS 8
#define UNROLL 16

struct OpIADD { static __device__ __forceinline__ uint32_t f(uint32_t a, uint32_t b, uint32_t c) { return a + b; } static const char *name() { return "IADD3 (a+b)"; } static int n() { return 1; } };
struct OpMNMX { static __device__ __forceinline__ uint32_t f(uint32_t a, uint32_t b, uint32_t c) { return (uint32_t)max((int)a, (int)b); } static const char *name() { return "IMNMX s32"; } static int n() { return 1; } };
struct OpLOP3 { static __device__ __forceinline__ uint32_t f(uint32_t a, uint32_t b, uint32_t c) { return (a & b) ^ c; } static const char *name() { return "LOP3"; } static int n() { return 1; } };
struct OpIMAD { static __device__ __forceinline__ uint32_t f(uint32_t a, uint32_t b, uint32_t c) { return a * b + c; } static const char *name() { return "IMAD"; } static int n() { return 1; } };
struct OpVADD { static __device__ __forceinline__ uint32_t f(uint32_t a, uint32_t b, uint32_t c) { return __vadd2(a, b); } static const char *name() { return "VIADD.16x2"; } static int n() { return 1; } };
struct OpVMAX { static __device__ __forceinline__ uint32_t f(uint32_t a, uint32_t b, uint32_t c) { return __vmaxs2(a, b); } static const char *name() { return "VIMNMX.S16x2"; } static int n() { return 1; } };
struct OpVMAX3 { static __device__ __forceinline__ uint32_t f(uint32_t a, uint32_t b, uint32_t c) { return __vimax3_s16x2(a, b, c); } static const char *name() { return "VIMNMX3.S16x2"; } static int n() { return 1; } };
struct OpVADDMAX { static __device__ __forceinline__ uint32_t f(uint32_t a, uint32_t b, uint32_t c) { return __viaddmax_s16x2(a, b, c); } static const char *name() { return "VIADDMNMX.S16x2"; } static int n() { return 1; } };
struct OpPRMT { static __device__ __forceinline__ uint32_t f(uint32_t a, uint32_t b, uint32_t c) { return __byte_perm(a, b, 0x6240); } static const char *name() { return "PRMT"; } static int n() { return 1; } };
struct OpSHF { static __device__ __forceinline__ uint32_t f(uint32_t a, uint32_t b, uint32_t c) { return __funnelshift_l(a, b, 7); } static const char *name() { return "SHF"; } static int n() { return 1; } };
struct OpMIX1 { // VIADD + VIMNMX3 + LOP3 + IMAD: the mix of the packed DP cell
	static __device__ __forceinline__ uint32_t f(uint32_t a, uint32_t b, uint32_t c) { uint32_t t = __vadd2(a, b); t = __vimax3_s16x2(t, b, c); t = (t & b) ^ c; return t * b + c; }
	static const char *name() { return "mix VIADD+VIMNMX3+LOP3+IMAD"; } static int n() { return 4; } };
struct OpMIX2 { // VIADD + VIADDMNMX + VIMNMX3 + LOP3 (no fma-pipe op)
	static __device__ __forceinline__ uint32_t f(uint32_t a, uint32_t b, uint32_t c) { uint32_t t = __vadd2(a, b); t = __viaddmax_s16x2(t, b, c); t = __vimax3_s16x2(t, b, c); return (t & b) ^ c; }
	static const char *name() { return "mix VIADD+VIADDMNMX+VIMNMX3+LOP3"; } static int n() { return 4; } };
struct OpMIX3 { // IADD3 + IMNMX (the int32 cell mix)
	static __device__ __forceinline__ uint32_t f(uint32_t a, uint32_t b, uint32_t c) { return (uint32_t)max((int)(a + b), (int)c); }
	static const char *name() { return "mix IADD3+IMNMX"; } static int n() { return 2; } };


#define MIX2(NAME, LABEL, E1, E2) struct NAME { static __device__ __forceinline__ uint32_t f(uint32_t a, uint32_t b, uint32_t c) { uint32_t t = E1; return E2; } \
	static const char *name() { return LABEL; } static int n() { return 2; } };
MIX2(MxVaddLop, "pair VIADD.16x2 + LOP3", __vadd2(a, b), ((t & b) ^ c))
MIX2(MxVaddImad, "pair VIADD.16x2 + IMAD", __vadd2(a, b), (t * 5u + c))
MIX2(MxVmaxImad, "pair VIMNMX.S16x2 + IMAD", __vmaxs2(a, b), (t * 5u + c))
MIX2(MxVmaxVadd, "pair VIMNMX.S16x2 + VIADD.16x2", __vmaxs2(a, b), __vadd2(t, c))
MIX2(MxLopImad, "pair LOP3 + IMAD", ((a & b) ^ c), (t * 5u + c))
MIX2(MxAdd3Lop, "pair IADD3(a+b+c) + LOP3", (a + b + c), ((t & b) ^ c))
MIX2(MxSubLop, "pair ISUB(a-b) + LOP3", (a - b), ((t & b) ^ c))
MIX2(MxVmax3Vadd, "pair VIMNMX3.S16x2 + VIADD.16x2", __vimax3_s16x2(a, b, c), __vadd2(t, c))
MIX2(MxVaddmaxSub, "pair VIADDMNMX.S16x2 + ISUB", __viaddmax_s16x2(a, b, c), (t - b))
MIX2(MxPrmtImad, "pair PRMT + IMAD", __byte_perm(a, b, 0x6240), (t * 5u + c))
struct OpSHFL { static __device__ __forceinline__ uint32_t f(uint32_t a, uint32_t b, uint32_t c) { return __shfl_up_sync(0xffffffffu, a, 1) + b; } static const char *name() { return "pair SHFL.UP + IADD"; } static int n() { return 2; } };

template <class Op>
__global__ void k_bench(uint32_t *out, int iters, uint32_t b, uint32_t c)
{
	uint32_t a[CHAINS];
	#pragma unroll
	for (int k = 0; k < CHAINS; ++k) a[k] = threadIdx.x * 2654435761u + k;
	#pragma unroll 1
	for (int i = 0; i < iters; ++i) {
		#pragma unroll
		for (int u = 0; u < UNROLL; ++u) {
			uint32_t n[CHAINS];
			#pragma unroll
			for (int k = 0; k < CHAINS; ++k) n[k] = Op::f(a[k], a[(k + 3) & (CHAINS - 1)], a[(k + 5) & (CHAINS - 1)]); // cross-chain operands: nothing folds
			#pragma unroll
			for (int k = 0; k < CHAINS; ++k) a[k] = n[k];
		}
	}
	uint32_t x = 0;
	#pragma unroll
	for (int k = 0; k < CHAINS; ++k) x ^= a[k];
	if (x == 0x12345678u) out[0] = x;
}

template <class Op>
static void run(int sms, double mhz, uint32_t *d, int warps_per_sm)
{
	const int tpb = 256, iters = 2048;
	const int grid = sms * (warps_per_sm * 32 / tpb);
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	k_bench<Op><<<grid, tpb>>>(d, 16, 0x00030005u, 0x00010001u);
	float best = 1e30f;
	for (int r = 0; r < 5; ++r) {
		cudaEventRecord(e0);
		k_bench<Op><<<grid, tpb>>>(d, iters, 0x00030005u + r, 0x00010001u);
		cudaEventRecord(e1); cudaEventSynchronize(e1);
		float ms; cudaEventElapsedTime(&ms, e0, e1);
		if (ms < best) best = ms;
	}
	const double winst = (double)Op::n() * CHAINS * UNROLL * iters * ((double)grid * tpb / 32);
	const double per_s = winst / (best * 1e-3);
	printf("%-36s %8.3f ms  %8.2f Gwarp-inst/s  %6.3f warp-inst/clk/SMSP (at %.0f MHz)  %7.2f Tlane-op/s\n", Op::name(), best, per_s / 1e9,
	       per_s / (sms * 4.0 * mhz * 1e6), mhz, per_s * 32 / 1e12);
}

int main()
{
	cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
	int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
	const double mhz = khz / 1e3;
	printf("device %s, %d SMs, clock %.0f MHz\n", p.name, p.multiProcessorCount, mhz);
	uint32_t *d; cudaMalloc(&d, 64);
	for (int wps : {32}) {
		printf("-- %d warps per SM --\n", wps);
		run<OpIADD>(p.multiProcessorCount, mhz, d, wps);
		run<OpMNMX>(p.multiProcessorCount, mhz, d, wps);
		run<OpLOP3>(p.multiProcessorCount, mhz, d, wps);
		run<OpIMAD>(p.multiProcessorCount, mhz, d, wps);
		run<OpSHF>(p.multiProcessorCount, mhz, d, wps);
		run<OpPRMT>(p.multiProcessorCount, mhz, d, wps);
		run<OpVADD>(p.multiProcessorCount, mhz, d, wps);
		run<OpVMAX>(p.multiProcessorCount, mhz, d, wps);
		run<OpVMAX3>(p.multiProcessorCount, mhz, d, wps);
		run<OpVADDMAX>(p.multiProcessorCount, mhz, d, wps);
		run<OpMIX3>(p.multiProcessorCount, mhz, d, wps);
		run<OpMIX1>(p.multiProcessorCount, mhz, d, wps);
		run<OpMIX2>(p.multiProcessorCount, mhz, d, wps);
		run<MxVaddLop>(p.multiProcessorCount, mhz, d, wps);
		run<MxVaddImad>(p.multiProcessorCount, mhz, d, wps);
		run<MxVmaxImad>(p.multiProcessorCount, mhz, d, wps);
		run<MxVmaxVadd>(p.multiProcessorCount, mhz, d, wps);
		run<MxLopImad>(p.multiProcessorCount, mhz, d, wps);
		run<MxAdd3Lop>(p.multiProcessorCount, mhz, d, wps);
		run<MxSubLop>(p.multiProcessorCount, mhz, d, wps);
		run<MxVmax3Vadd>(p.multiProcessorCount, mhz, d, wps);
		run<MxVaddmaxSub>(p.multiProcessorCount, mhz, d, wps);
		run<MxPrmtImad>(p.multiProcessorCount, mhz, d, wps);
		run<OpSHFL>(p.multiProcessorCount, mhz, d, wps);
	}
	cudaError_t e = cudaDeviceSynchronize();
	if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
	return 0;
}
