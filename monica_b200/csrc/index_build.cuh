// index_build.cuh -- the bucket/sort/hash half of the index build on the device.
//
// Replaces minimap2-2.17 index.c worker_post() (sort the minimizers of a bucket by hash, group equal hashes, singleton ->
// position in the hash value, otherwise a slice of the position array with ascending positions) and mm_idx_cal_max_occ(),
// reached in the reference through mappy.Aligner(fn_idx_in=<fna.gz>, ...) at monica/genomes/aligner.py:45-46.
//
//   * stable LSD radix sort of the 128-bit minimizer records on the 2k-bit hash, 8 bits per pass: per-tile digit histograms,
//     one scan, and a scatter that ranks the 256 items of a sub-tile with __match_any_sync (stable: the sketch emits contigs
//     in order and positions ascending, so every position list comes out ascending like upstream's),
//   * group boundaries by flag + scan, open-addressing insert with atomicCAS (slot placement differs from a sequential
//     insert; lookups do not depend on it), position lists filled per group,
//   * the occurrence histogram for mid_occ.
#pragma once
#include "common.cuh"
#include "seed.cuh"

#define RS_TPB 256
#define RS_TILE 4096          // items per CTA per pass (16 sub-tiles of 256)

__global__ void __launch_bounds__(RS_TPB)
k_rs_hist(const mb128 *__restrict__ in, int64_t n, int shift, int32_t *__restrict__ hist /* [256][n_tiles] */, int64_t n_tiles)
{
	__shared__ int32_t h[256];
	h[threadIdx.x] = 0;
	__syncthreads();
	const int64_t base = (int64_t)blockIdx.x * RS_TILE;
	for (int i = threadIdx.x; i < RS_TILE; i += RS_TPB) {
		const int64_t g = base + i;
		if (g < n) atomicAdd(&h[(int)((in[g].x >> 8) >> shift & 255)], 1);
	}
	__syncthreads();
	hist[(int64_t)threadIdx.x * n_tiles + blockIdx.x] = h[threadIdx.x];
}

__global__ void __launch_bounds__(RS_TPB)
k_rs_scatter(const mb128 *__restrict__ in, mb128 *__restrict__ out, int64_t n, int shift, const int64_t *__restrict__ off /* scan of hist */, int64_t n_tiles)
{
	__shared__ int64_t dig_off[256];          // running output offset of each digit for this tile
	__shared__ int32_t wcnt[RS_TPB / 32][256]; // per-warp digit counts of the current sub-tile
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	dig_off[threadIdx.x] = off[(int64_t)threadIdx.x * n_tiles + blockIdx.x];
	const int64_t base = (int64_t)blockIdx.x * RS_TILE;
	for (int sub = 0; sub < RS_TILE; sub += RS_TPB) {
		for (int w = 0; w < RS_TPB / 32; ++w) wcnt[w][threadIdx.x] = 0;
		__syncthreads();
		const int64_t g = base + sub + threadIdx.x;
		const bool ok = g < n;
		mb128 v; v.x = v.y = 0;
		int d = 256; // out-of-range lanes form their own group
		if (ok) { v = in[g]; d = (int)((v.x >> 8) >> shift & 255); }
		const unsigned peers = __match_any_sync(0xffffffffu, d);
		const int rank = __popc(peers & ((1u << lane) - 1));
		if (ok && rank == 0) wcnt[wid][d] = __popc(peers);
		__syncthreads();
		// thread t owns digit t: prefix of the per-warp counts, then advance the running offset
		int32_t run = 0, mine_before = 0;
		{
			const int t = threadIdx.x;
			for (int w = 0; w < RS_TPB / 32; ++w) { const int32_t c = wcnt[w][t]; wcnt[w][t] = run; run += c; }
			(void)mine_before;
		}
		__syncthreads();
		if (ok) out[dig_off[d] + wcnt[wid][d] + rank] = v;
		__syncthreads();
		dig_off[threadIdx.x] += run;
		__syncthreads();
	}
}

// sorts a[0..n) by hash = x >> 8 (key_bits significant bits), stable; returns the buffer that holds the result (a or tmp)
static mb128 *radix_sort_minimizers(Arena &ar, cudaStream_t st, mb128 *a, mb128 *tmp, int64_t n, int key_bits)
{
	if (n <= 1) return a;
	const int64_t n_tiles = cdiv(n, RS_TILE);
	int32_t *hist = ar.get<int32_t>(256 * n_tiles);
	int64_t *off = ar.get<int64_t>(256 * n_tiles + 1);
	mb128 *src = a, *dst = tmp;
	for (int shift = 0; shift < key_bits; shift += 8) {
		k_rs_hist<<<(unsigned)n_tiles, RS_TPB, 0, st>>>(src, n, shift, hist, n_tiles);
		exclusive_scan<int32_t>(ar, st, hist, off, 256 * n_tiles, nullptr);
		k_rs_scatter<<<(unsigned)n_tiles, RS_TPB, 0, st>>>(src, dst, n, shift, off, n_tiles);
		mb128 *t_ = src; src = dst; dst = t_;
	}
	return src;
}

__global__ void k_ib_flag(const mb128 *__restrict__ m, int64_t n, int32_t *__restrict__ flag)
{
	const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) flag[i] = (i == 0 || (m[i].x >> 8) != (m[i - 1].x >> 8)) ? 1 : 0;
}

__global__ void k_ib_start(const int32_t *__restrict__ flag, const int64_t *__restrict__ gid, int64_t n, int64_t n_keys, int64_t *__restrict__ start)
{
	const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n && flag[i]) start[gid[i]] = i;
	if (i == 0) start[n_keys] = n;
}

__global__ void k_ib_multi(const int64_t *__restrict__ start, int64_t n_keys, int32_t *__restrict__ mcnt, unsigned int *__restrict__ occ_hist)
{
	const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (g >= n_keys) return;
	const int64_t c = start[g + 1] - start[g];
	mcnt[g] = c > 1 ? (int32_t)c : 0;
	atomicAdd(&occ_hist[c < 65535 ? c : 65535], 1u);
}

__global__ void k_ib_insert(const mb128 *__restrict__ m, const int64_t *__restrict__ start, const int64_t *__restrict__ moff, int64_t n_keys,
                            unsigned long long *__restrict__ hkey, uint64_t *__restrict__ hval, uint64_t hmask, int hshift, uint64_t *__restrict__ pos)
{
	const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (g >= n_keys) return;
	const int64_t s = start[g], c = start[g + 1] - s;
	const uint64_t h = m[s].x >> 8;
	const unsigned long long key = c == 1 ? (h << 1 | 1ULL) : (h << 1);
	uint64_t slot = mb_slot_hash(h, hshift) & hmask;
	while (atomicCAS(&hkey[slot], ~0ULL, key) != ~0ULL) slot = (slot + 1) & hmask;
	if (c == 1) hval[slot] = m[s].y;
	else {
		const int64_t o = moff[g];
		hval[slot] = (uint64_t)o << 32 | (uint32_t)c;
		for (int64_t k = 0; k < c; ++k) pos[o + k] = m[s + k].y;
	}
}
