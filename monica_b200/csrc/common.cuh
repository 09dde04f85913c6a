// common.cuh -- shared types, error handling, device arena and prefix-scan for the sm_100a mapper.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <vector>
#include <stdexcept>
#include "../../include/monica_b200.h"

#define MB_HD __host__ __device__ __forceinline__
#define MB_D  __device__ __forceinline__

struct mb128 { uint64_t x, y; };

// anchor flag bits (minimap2 mmpriv.h)
#define MB_SEED_LONG_JOIN (1ULL<<40)
#define MB_SEED_IGNORE    (1ULL<<41)
#define MB_SEED_TANDEM    (1ULL<<42)
#define MB_SEED_SELF      (1ULL<<43)
#define MB_SEED_SEG_SHIFT 48
#define MB_PARENT_UNSET   (-1)
#define MB_PARENT_TMP_PRI (-2)

#define MB_KSW_NEG_INF     (-0x40000000)
#define MB_EZ_SCORE_ONLY   0x01
#define MB_EZ_RIGHT        0x02
#define MB_EZ_APPROX_MAX   0x08
#define MB_EZ_APPROX_DROP  0x10
#define MB_EZ_EXTZ_ONLY    0x40
#define MB_EZ_REV_CIGAR    0x80

extern thread_local std::string g_mb_err;

struct mb_error : public std::runtime_error {
	int code;
	mb_error(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};

#define CK(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { \
	char b__[512]; snprintf(b__, sizeof(b__), "CUDA error %s at %s:%d: %s", cudaGetErrorName(e__), __FILE__, __LINE__, cudaGetErrorString(e__)); \
	throw mb_error(MB_ERR_CUDA, b__); } } while (0)

static inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---------------------------------------------------------------------------------------------
// Device arena: one growable slab per calling thread; bump allocation, reset per batch.
// Avoids cudaMalloc/cudaFree (implicit syncs) inside the pipeline.
// ---------------------------------------------------------------------------------------------
struct Arena {
	char *base = nullptr;
	size_t cap = 0, used = 0, high = 0;
	size_t batch_total = 0;     // bytes handed out since the last reset, over all slabs
	std::vector<char*> retired; // slabs outgrown during a batch; freed at reset
	void reset() {
		const bool grew = !retired.empty();
		for (char *p : retired) cudaFree(p);
		retired.clear();
		// A batch that outgrew its slab leaves a last slab that held only the tail of its allocations; the next batch of the
		// same size would outgrow it again (a cudaMalloc / cudaFree of gigabytes inside that batch, and again in the one after).
		// Size the arena for the whole batch at once instead, so that the second batch already runs without any allocation.
		if (grew) {
			const size_t want = ((batch_total + batch_total / 4 + ((size_t)64 << 20)) + 255) & ~(size_t)255;
			if (want > cap) {
				char *nb = nullptr;
				if (base) cudaFree(base);
				base = nullptr; cap = 0;
				if (cudaMalloc(&nb, want) == cudaSuccess) base = nb, cap = want;
				else cudaGetLastError(); // fall back to growing on demand
			}
		}
		used = 0; batch_total = 0;
	}
	// size the (empty) arena for a known demand up front, instead of doubling slab by slab (the retired slabs of a doubling
	// sequence add up to as much again as the final one)
	void reserve(size_t want) {
		if (used != 0 || !retired.empty() || want <= cap) return;
		want = (want + 255) & ~(size_t)255;
		if (base) cudaFree(base);
		base = nullptr; cap = 0;
		char *nb = nullptr;
		if (cudaMalloc(&nb, want) == cudaSuccess) base = nb, cap = want;
		else cudaGetLastError(); // fall back to growing on demand
	}
	void release() {
		reset();
		if (base) cudaFree(base);
		base = nullptr; cap = 0;
	}
	// NB: growing mid-batch keeps old allocations valid (old slab retired, not freed) but they are NOT moved.
	void *alloc(size_t bytes) {
		bytes = (bytes + 255) & ~(size_t)255;
		if (const char *lim = getenv("MB_TEST_ARENA_LIMIT"))   // tests: pretend the device holds this much scratch
			if (batch_total + bytes > (size_t)atoll(lim)) throw mb_error(MB_ERR_NOMEM, "scratch limit (MB_TEST_ARENA_LIMIT) reached");
		if (used + bytes > cap) {
			size_t ncap = cap ? cap * 2 : ((size_t)256 << 20);
			while (ncap < bytes * 2) ncap <<= 1;
			if (base) retired.push_back(base);
			char *nb = nullptr;
			cudaError_t e = cudaMalloc(&nb, ncap);
			if (e != cudaSuccess) { // doubling does not fit: take what is left, if that covers the request with some room
				cudaGetLastError();
				size_t fr = 0, tot = 0;
				if (cudaMemGetInfo(&fr, &tot) == cudaSuccess && fr > ((size_t)2 << 30)) {
					const size_t fit = (fr - ((size_t)1 << 30)) & ~(size_t)255;
					if (fit >= bytes + bytes / 4 && fit < ncap) { ncap = fit; e = cudaMalloc(&nb, ncap); if (e != cudaSuccess) cudaGetLastError(); }
				}
			}
			if (e != cudaSuccess) throw mb_error(MB_ERR_NOMEM, std::string("cudaMalloc failed for ") + std::to_string(ncap) + " bytes: " + cudaGetErrorString(e));
			base = nb; cap = ncap; used = 0;
		}
		void *p = base + used;
		used += bytes; batch_total += bytes;
		if (used > high) high = used;
		return p;
	}
	template <typename T> T *get(size_t n) { return (T*)alloc((n ? n : 1) * sizeof(T)); }
};

// ---------------------------------------------------------------------------------------------
// Exclusive prefix sum (int64 out) over int32/int64 inputs: block scan + recursive block sums.
// ---------------------------------------------------------------------------------------------
#define SCAN_TPB 256
#define SCAN_IPT 8
#define SCAN_TILE (SCAN_TPB * SCAN_IPT)

template <typename Tin>
__global__ void k_scan_tile(const Tin *__restrict__ in, int64_t *__restrict__ out, int64_t *__restrict__ tile_sums, int64_t n)
{
	__shared__ int64_t warp_sums[SCAN_TPB / 32];
	int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_IPT;
	int64_t v[SCAN_IPT], s = 0;
	#pragma unroll
	for (int i = 0; i < SCAN_IPT; ++i) { v[i] = (base + i < n) ? (int64_t)in[base + i] : 0; s += v[i]; }
	int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	int64_t incl = s;
	#pragma unroll
	for (int d = 1; d < 32; d <<= 1) { int64_t t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += t; }
	if (lane == 31) warp_sums[wid] = incl;
	__syncthreads();
	if (wid == 0) {
		int64_t w = lane < SCAN_TPB / 32 ? warp_sums[lane] : 0;
		#pragma unroll
		for (int d = 1; d < 32; d <<= 1) { int64_t t = __shfl_up_sync(0xffffffffu, w, d); if (lane >= d) w += t; }
		if (lane < SCAN_TPB / 32) warp_sums[lane] = w;
	}
	__syncthreads();
	int64_t excl = incl - s + (wid ? warp_sums[wid - 1] : 0);
	#pragma unroll
	for (int i = 0; i < SCAN_IPT; ++i) { if (base + i < n) out[base + i] = excl; excl += v[i]; }
	if (threadIdx.x == SCAN_TPB - 1 && tile_sums) tile_sums[blockIdx.x] = excl;
}

__global__ void k_scan_add(int64_t *__restrict__ out, const int64_t *__restrict__ tile_off, int64_t n)
{
	int64_t i = (int64_t)blockIdx.x * SCAN_TILE + threadIdx.x;
	int64_t add = tile_off[blockIdx.x];
	#pragma unroll
	for (int k = 0; k < SCAN_IPT; ++k, i += SCAN_TPB) if (i < n) out[i] += add;
}

// out[0..n) = exclusive scan, out[n] = total.  `out` must have n+1 entries.
template <typename Tin>
static void exclusive_scan(Arena &ar, cudaStream_t st, const Tin *in, int64_t *out, int64_t n, int64_t *n_launch)
{
	if (n <= 0) { CK(cudaMemsetAsync(out, 0, sizeof(int64_t), st)); return; }
	int64_t tiles = cdiv(n, SCAN_TILE);
	int64_t *sums = ar.get<int64_t>(tiles + 1);
	k_scan_tile<Tin><<<(unsigned)tiles, SCAN_TPB, 0, st>>>(in, out, sums, n);
	if (n_launch) ++*n_launch;
	if (tiles == 1) {
		CK(cudaMemcpyAsync(out + n, sums, sizeof(int64_t), cudaMemcpyDeviceToDevice, st));
		return;
	}
	int64_t *sums_scan = ar.get<int64_t>(tiles + 1);
	exclusive_scan<int64_t>(ar, st, sums, sums_scan, tiles, n_launch);
	k_scan_add<<<(unsigned)tiles, SCAN_TPB, 0, st>>>(out, sums_scan, n);
	if (n_launch) ++*n_launch;
	CK(cudaMemcpyAsync(out + n, sums_scan + tiles, sizeof(int64_t), cudaMemcpyDeviceToDevice, st));
}
