// seed.cuh -- K2 hash-index seed lookup, anchor emission, K2b per-read anchor sort.
//
// Replaces minimap2-2.17 index.c mm_idx_get() and map.c collect_matches()/collect_seed_hits()
// (+ ksort.h radix_sort_128x) as reached from mappy's Aligner.map
// (/root/reference/monica/genomes/aligner.py:193,215).
//
// HBM layout of the index (one replica per GPU):
//   hkey[cap], hval[cap]  open-addressing table, linear probing, load <= 0.5.
//                         hkey = hash<<1 | singleton, ~0 = empty; hval = position (singleton) or start<<32|n
//   pos[n_multi]          position lists of multi-occurrence minimizers, each list ascending (index.c worker_post)
//   S[sum_len/8]          4-bit packed reference, 8 bases per word (mm_idx_t::S), seq_off[], seq_len[]
// A probe touches one 16-byte slot pair (two 8-byte loads from two arrays -> 2 sectors); singletons need no
// second dependent access.
#pragma once
#include "common.cuh"
#include "radix_emul.cuh"

struct DevIndex {
	uint64_t *hkey = nullptr, *hval = nullptr;
	uint64_t hmask = 0;
	int hshift = 0;
	uint64_t *pos = nullptr;
	uint32_t *S = nullptr;
	uint64_t *seq_off = nullptr;
	uint32_t *seq_len = nullptr;
	int n_seq = 0, k = 15, w = 10, mid_occ = 0;
	uint32_t max_seq_len = 0;   // longest contig (bounds the position bytes of an anchor's x)
};

MB_HD uint64_t mb_slot_hash(uint64_t h, int shift) { return (h * 0x9E3779B97F4A7C15ULL) >> shift; }

// one thread per query minimizer: probe, record occurrence count and value
__global__ void k_seed_lookup(DevIndex ix, const mb128 *__restrict__ mini, int64_t n_mini, int mid_occ,
                              int32_t *__restrict__ occ, uint64_t *__restrict__ val, int32_t *__restrict__ cnt)
{
	int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_mini) return;
	uint64_t h = mini[i].x >> 8;
	uint64_t slot = mb_slot_hash(h, ix.hshift) & ix.hmask;
	int n = 0; uint64_t v = 0;
	for (;;) {
		uint64_t kk = __ldg(ix.hkey + slot);
		if (kk == ~0ULL) break;
		if ((kk >> 1) == h) {
			v = __ldg(ix.hval + slot);
			n = (kk & 1) ? 1 : (int)(uint32_t)v;
			if (kk & 1) v = v; // singleton: v is the position itself
			else v = (v >> 32) | (1ULL << 63); // multi: start index into pos[], tagged
			break;
		}
		slot = (slot + 1) & ix.hmask;
	}
	occ[i] = n;
	val[i] = v;
	cnt[i] = n < mid_occ ? n : 0;
}

// one warp per read: rep_len (collect_matches) -- union length of repetitive-minimizer intervals.  The lanes test 32
// occurrence counts at a time (coalesced); the sequential interval merge only runs for the few minimizers at or above mid_occ.
__global__ void k_rep_len(const mb128 *__restrict__ mini, const int64_t *__restrict__ mini_off, const int32_t *__restrict__ occ,
                          int n_reads, int mid_occ, int32_t *__restrict__ rep_len)
{
	const int r = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
	const int lane = threadIdx.x & 31;
	if (r >= n_reads) return;
	int rep_st = 0, rep_en = 0, rl = 0;
	const int64_t lo = mini_off[r], hi = mini_off[r + 1];
	for (int64_t i0 = lo; i0 < hi; i0 += 32) {
		const int64_t i = i0 + lane;
		const bool rep = i < hi && occ[i] >= mid_occ;
		uint32_t q_pos = 0, q_span = 0;
		if (rep) { q_pos = (uint32_t)mini[i].y, q_span = (uint32_t)(mini[i].x & 0xff); }
		for (unsigned m = __ballot_sync(0xffffffffu, rep); m; m &= m - 1) {
			const int src = __ffs(m) - 1;
			const uint32_t qp = __shfl_sync(0xffffffffu, q_pos, src), qs = __shfl_sync(0xffffffffu, q_span, src);
			const int en = (int)(qp >> 1) + 1, st = en - (int)qs;
			if (st > rep_en) { rl += rep_en - rep_st; rep_st = st, rep_en = en; }
			else rep_en = en;
		}
	}
	rl += rep_en - rep_st;
	if (lane == 0) rep_len[r] = rl;
}

// one thread per query minimizer: emit its anchors (collect_seed_hits) in upstream emission order
__global__ void k_seed_fill(DevIndex ix, const mb128 *__restrict__ mini, const int64_t *__restrict__ mini_off, int64_t n_mini,
                            const int64_t *__restrict__ read_off, const int32_t *__restrict__ cnt, const uint64_t *__restrict__ val,
                            const int64_t *__restrict__ a_off, mb128 *__restrict__ a)
{
	int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_mini) return;
	int n = cnt[i];
	if (n == 0) return;
	mb128 m = mini[i];
	int r = (int)(m.y >> 32);
	uint32_t q_pos = (uint32_t)m.y, q_span = (uint32_t)(m.x & 0xff);
	int qlen = (int)(read_off[r + 1] - read_off[r]);
	bool tandem = false;
	if (i > mini_off[r] && (mini[i - 1].x >> 8) == (m.x >> 8)) tandem = true;
	if (i + 1 < mini_off[r + 1] && (mini[i + 1].x >> 8) == (m.x >> 8)) tandem = true;
	uint64_t v = val[i];
	const uint64_t *cr = (v >> 63) ? ix.pos + (v & 0x7fffffffffffffffULL) : nullptr;
	int64_t o = a_off[i];
	for (int kk = 0; kk < n; ++kk) {
		uint64_t rr = cr ? __ldg(cr + kk) : v;
		uint32_t rpos = (uint32_t)rr >> 1;
		mb128 p;
		if ((rr & 1) == (q_pos & 1)) { // forward strand
			p.x = (rr & 0xffffffff00000000ULL) | rpos;
			p.y = (uint64_t)q_span << 32 | q_pos >> 1;
		} else { // reverse strand
			p.x = 1ULL << 63 | (rr & 0xffffffff00000000ULL) | rpos;
			p.y = (uint64_t)q_span << 32 | (uint32_t)(qlen - (int)((q_pos >> 1) + 1 - q_span) - 1);
		}
		if (tandem) p.y |= MB_SEED_TANDEM;
		a[o + kk] = p;
	}
}

// per-read anchor offsets: a_roff[r] = a_off[mini_off[r]]
__global__ void k_read_anchor_off(const int64_t *__restrict__ mini_off, const int64_t *__restrict__ a_off, int n_reads, int64_t *__restrict__ a_roff)
{
	int r = blockIdx.x * blockDim.x + threadIdx.x;
	if (r > n_reads) return;
	a_roff[r] = a_off[mini_off[r]];
}

// ---- reads in descending order of anchor count ----
// The per-read kernels that follow (anchor sort, chaining, chain backtrack) give a read to one CTA or one warp, and a read's
// cost grows with its anchor count; the longest read of a batch costs milliseconds on its own.  Visiting the reads longest
// first lets that tail run underneath the bulk instead of after it.  Counting sort on n_anchors / 32 (256 buckets).
#define RP_NB 256
MB_D int mb_read_bucket(const int64_t *a_roff, int r) { const int64_t b = (a_roff[r + 1] - a_roff[r]) >> 5; return RP_NB - 1 - (int)(b < RP_NB - 1 ? b : RP_NB - 1); }
__global__ void k_rp_hist(const int64_t *__restrict__ a_roff, int n_reads, int32_t *__restrict__ hist)
{
	const int r = blockIdx.x * blockDim.x + threadIdx.x;
	if (r < n_reads) atomicAdd(&hist[mb_read_bucket(a_roff, r)], 1);
}
__global__ void k_rp_scan(int32_t *__restrict__ hist) // one block of RP_NB threads: counts -> exclusive offsets (then used as cursors)
{
	__shared__ int32_t s[RP_NB];
	const int t = threadIdx.x;
	s[t] = hist[t];
	__syncthreads();
	for (int d = 1; d < RP_NB; d <<= 1) { const int v = t >= d ? s[t - d] : 0; __syncthreads(); s[t] += v; __syncthreads(); }
	hist[t] = s[t] - hist[t];
}
__global__ void k_rp_scatter(const int64_t *__restrict__ a_roff, int n_reads, int32_t *__restrict__ cursor, int32_t *__restrict__ perm)
{
	const int r = blockIdx.x * blockDim.x + threadIdx.x;
	if (r < n_reads) perm[atomicAdd(&cursor[mb_read_bucket(a_roff, r)], 1)] = r;
}

// ---- K2b: one CTA per read; bitonic sort of (x, original index) in shared memory, global scratch beyond ----
#define SORT_TPB 256
#define SORT_SMEM_N 2048   // 2048 * (8+4) = 24 KB static shared memory

__global__ void __launch_bounds__(SORT_TPB)
k_sort_anchors(const mb128 *__restrict__ in, mb128 *__restrict__ out, const int64_t *__restrict__ a_roff, int n_reads,
               int32_t *__restrict__ tie_list, int32_t *__restrict__ n_tie, const int32_t *__restrict__ perm)
{
	__shared__ uint64_t skey[SORT_SMEM_N];
	__shared__ uint32_t sidx[SORT_SMEM_N];
	__shared__ int s_tie;
	for (int ri = blockIdx.x; ri < n_reads; ri += gridDim.x) {
		const int r = perm ? perm[ri] : ri;
		const int64_t base = a_roff[r];
		const int n = (int)(a_roff[r + 1] - base);
		if (n == 0) continue;
		if (n == 1) { if (threadIdx.x == 0) out[base] = in[base]; continue; }
		int np = 1; while (np < n) np <<= 1;
		if (np > SORT_SMEM_N) continue;          // sorted by k_sort_anchors_big (LSD radix in global memory)
		uint64_t *key = skey; uint32_t *idx = sidx;
		if (threadIdx.x == 0) s_tie = 0;
		for (int i = threadIdx.x; i < np; i += SORT_TPB) {
			key[i] = i < n ? in[base + i].x : ~0ULL;
			idx[i] = i;
		}
		__syncthreads();
		// bitonic network.  Sub-steps with partner distance j >= 32 exchange through the key/idx arrays, one compare-exchange
		// pair per thread-iteration; the tail of every stage (j <= 16: partners sit in the same warp) runs in registers with
		// shuffles, one load and one store per element and one barrier for up to five sub-steps.
		const unsigned wmask = np >= 32 ? 0xffffffffu : (1u << np) - 1u;
		for (int k2 = 2; k2 <= np; k2 <<= 1) {
			int j = k2 >> 1;
			for (; j >= 32; j >>= 1) {
				for (int t = threadIdx.x; t < (np >> 1); t += SORT_TPB) {
					const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1)), l = i | j;
					const uint64_t ka = key[i], kb = key[l];
					const uint32_t ia = idx[i], ib = idx[l];
					const bool gt = ka > kb || (ka == kb && ia > ib);
					const bool up = (i & k2) == 0;
					if (gt == up) { key[i] = kb, key[l] = ka, idx[i] = ib, idx[l] = ia; }
				}
				__syncthreads();
			}
			for (int i = threadIdx.x; i < np; i += SORT_TPB) {
				uint64_t k = key[i]; uint32_t ix = idx[i];
				const bool up = (i & k2) == 0;
				for (int jj = j; jj > 0; jj >>= 1) {
					const uint64_t ok = __shfl_xor_sync(wmask, k, jj);
					const uint32_t oi = __shfl_xor_sync(wmask, ix, jj);
					const bool gt = k > ok || (k == ok && ix > oi);
					const bool keep_min = ((i & jj) == 0) == up;
					if (gt == keep_min) k = ok, ix = oi;
				}
				key[i] = k, idx[i] = ix;
			}
			__syncthreads();
		}
		for (int i = threadIdx.x; i < n; i += SORT_TPB) {
			out[base + i] = in[base + idx[i]];
			if (i > 0 && key[i] == key[i - 1]) s_tie = 1;
		}
		__syncthreads();
		if (threadIdx.x == 0 && s_tie) tie_list[atomicAdd(n_tie, 1)] = r;
		__syncthreads();
	}
}

// Reads with more anchors than the shared-memory sort holds (long reads against a large database: tens of thousands of
// anchors each): one CTA per read, stable LSD radix sort on x in global memory, 8 bits per pass, only over the byte positions
// of x that can differ (position bytes up to the longest contig, contig-id bytes up to n_seq, the strand byte).  Each pass:
// digit histogram of the read's segment -> 256-bin scan in shared memory -> in-order scatter of 256-anchor tiles, ranked with
// __match_any_sync like the index build's k_rs_scatter.  The unsorted input is left intact (the exact replay of upstream's
// unstable sort for reads with tied keys starts from it); the passes alternate between `tmp` and `out` so that the last one
// lands in `out`.
#define SB_TPB 256
#define TL_MAX 64     // tie positions kept per read; a read with more is replayed in full
__global__ void __launch_bounds__(SB_TPB)
k_sort_anchors_big(const mb128 *__restrict__ in, mb128 *__restrict__ out, mb128 *__restrict__ tmp, const int64_t *__restrict__ a_roff,
                   const int32_t *__restrict__ big_list, const int32_t *__restrict__ n_big, int32_t *__restrict__ cursor, const int64_t *__restrict__ tmp_roff,
                   unsigned pass_mask /* bit b: byte b of x takes part */, int32_t *__restrict__ tie_list, int32_t *__restrict__ n_tie,
                   int32_t *__restrict__ tie_pos /* [n_reads][TL_MAX] sorted positions i with x[i] == x[i-1] */, int32_t *__restrict__ tie_n)
{
	__shared__ int32_t s_hist[256];
	__shared__ int32_t s_off[256];
	__shared__ int32_t s_wcnt[SB_TPB / 32][256];
	__shared__ int s_item, s_tie;
	const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
	const int n_pass = __popc(pass_mask);
	for (;;) {
		__syncthreads();
		if (tid == 0) s_item = atomicAdd(cursor, 1);
		__syncthreads();
		const int item = s_item;
		if (item >= *n_big) break;
		const int r = big_list[item];
		const int64_t base = a_roff[r];
		const int n = (int)(a_roff[r + 1] - base);
		mb128 *T = tmp + tmp_roff[r], *O = out + base;
		const mb128 *src = in + base;
		int done = 0;
		for (int b = 0; b < 8; ++b) {
			if (!(pass_mask >> b & 1)) continue;
			// pass number `done` of n_pass writes to: the last one to O, the one before to T, ...
			mb128 *dst = ((n_pass - 1 - done) & 1) ? T : O;
			const int shift = b * 8;
			s_hist[tid] = 0;
			__syncthreads();
			for (int i = tid; i < n; i += SB_TPB) atomicAdd(&s_hist[(int)(src[i].x >> shift & 255)], 1);
			__syncthreads();
			if (wid == 0) { // exclusive scan of the 256 bins: 8 per lane
				int c[8], tot = 0;
				#pragma unroll
				for (int u = 0; u < 8; ++u) { c[u] = s_hist[lane * 8 + u]; tot += c[u]; }
				int incl = tot;
				#pragma unroll
				for (int d = 1; d < 32; d <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += v; }
				int run = incl - tot;
				#pragma unroll
				for (int u = 0; u < 8; ++u) { s_off[lane * 8 + u] = run; run += c[u]; }
			}
			__syncthreads();
			for (int t0 = 0; t0 < n; t0 += SB_TPB) {
				for (int w = 0; w < SB_TPB / 32; ++w) s_wcnt[w][tid] = 0;
				__syncthreads();
				const int i = t0 + tid;
				const bool ok = i < n;
				mb128 v; v.x = v.y = 0;
				int d = 256;
				if (ok) { v = src[i]; d = (int)(v.x >> shift & 255); }
				const unsigned peers = __match_any_sync(0xffffffffu, d);
				const int rank = __popc(peers & ((1u << lane) - 1));
				if (ok && rank == 0) s_wcnt[wid][d] = __popc(peers);
				__syncthreads();
				int run = 0;
				for (int w = 0; w < SB_TPB / 32; ++w) { const int cc = s_wcnt[w][tid]; s_wcnt[w][tid] = run; run += cc; }   // thread t owns digit t
				__syncthreads();
				if (ok) dst[s_off[d] + s_wcnt[wid][d] + rank] = v;
				__syncthreads();
				s_off[tid] += run;
				__syncthreads();
			}
			src = dst;
			++done;
		}
		if (n_pass == 0) for (int i = tid; i < n; i += SB_TPB) O[i] = in[base + i];
		__syncthreads();
		// tied keys -> the exact replay (k_sort_emul) redoes this read from the unsorted input
		if (tid == 0) s_tie = 0;
		__syncthreads();
		for (int i = tid + 1; i < n; i += SB_TPB)
			if (O[i].x == O[i - 1].x) { const int j = atomicAdd(&s_tie, 1); if (j < TL_MAX) tie_pos[(size_t)r * TL_MAX + j] = i; }
		__syncthreads();
		if (tid == 0 && s_tie) { tie_list[atomicAdd(n_tie, 1)] = r; tie_n[r] = s_tie; }
	}
}

// list the reads that take the big sort and give each its slice of the ping-pong buffer
__global__ void k_big_sort_list(const int64_t *__restrict__ a_roff, int n_reads, int32_t *__restrict__ sz, int32_t *__restrict__ big_list, int32_t *__restrict__ n_big)
{
	int r = blockIdx.x * blockDim.x + threadIdx.x;
	if (r >= n_reads) return;
	int64_t n = a_roff[r + 1] - a_roff[r];
	int np = 1; while (np < n) np <<= 1;
	const bool big = np > SORT_SMEM_N;
	sz[r] = big ? (int32_t)n : 0;
	if (big) big_list[atomicAdd(n_big, 1)] = r;
}

// ---- exact replay for LONG reads with tied keys ----
// What has to be reproduced is the order in which upstream's in-place MSD radix sort (ksort.h rs_sort: cycle-leader
// permutation per byte level, insertion sort below 65 elements) leaves anchors of EQUAL key; everything else is fixed by the
// keys and already sits in `out` from the stable sort.  A range of the recursion that holds no tied pair ends up as its stable
// order, so the replay only descends into ranges that hold one: per tied pair one chain of nested ranges, top level = the
// whole read.  The cycle-leader walk of a range is inherently sequential (which slot an element lands in depends on every
// element visited before it), so the parallelism is across ranges: a device-wide queue of (read, range, shift) items, ONE
// LANE per item, each lane a small state machine that advances by one element per iteration, so that the lanes of a warp stay
// in lock step whatever mix of items they hold and nobody spins while a warp-mate works (sub-ranges are pushed by the lanes
// themselves; `pending` counts items pushed and not yet finished).  The 256 moving bucket heads of a lane live in shared
// memory (1 KB per lane, 192 lanes per SM); the working copy W of a read is the radix sort's ping-pong slice; leaves and
// last-level buckets that hold ties are copied from W into `out`.  A long read costs ~1.7 walks of its length, all long tied
// reads of a batch walk concurrently.
#define SQ_TPB 192
struct EmulQ { int4 *items; int *ready; int *ctr; int cap; };   // ctr: [0] tail, [1] head, [2] pending

static __device__ __forceinline__ mb128 sq_ld(const mb128 *p) { const ulonglong2 v = __ldcg(reinterpret_cast<const ulonglong2*>(p)); mb128 r; r.x = v.x, r.y = v.y; return r; }
static __device__ __forceinline__ void sq_st(mb128 *p, const mb128 &v) { __stcg(reinterpret_cast<ulonglong2*>(p), make_ulonglong2(v.x, v.y)); }
static __device__ __forceinline__ uint64_t sq_ldx(const mb128 *p) { return __ldcg(reinterpret_cast<const unsigned long long*>(&p->x)); }

static __device__ __forceinline__ void sq_push(const EmulQ &q, int r, int lo, int hi, int s)
{
	atomicAdd(&q.ctr[2], 1);
	const int sl = atomicAdd(&q.ctr[0], 1);
	if (sl < q.cap) { __stcg(&q.items[sl], make_int4(r, lo, hi, s)); __threadfence(); *((volatile int*)&q.ready[sl]) = 1; }
	else atomicSub(&q.ctr[2], 1);   // cannot happen: cap covers n/65 ranges per level on 8 levels
}

// Start of the replay for one long tied read, one CTA per read: working copy W of the unsorted anchors and the first item.
// A level with exactly TWO occupied buckets (the usual first one: the strand byte) has a closed form and is applied here in
// parallel instead of being walked: with bucket 0 = [0, c0), p_1 < p_2 < ... the elements of bucket 0's region that belong
// to bucket 1 and q_1 < q_2 < ... those of bucket 1's region that belong to bucket 0, the i-th cycle of the walk puts X(p_i)
// at the head of bucket 1, shifts the elements up to q_i - 1 right by one and closes with Y(q_i) going to p_i.
#define SQS_TPB 256
__global__ void __launch_bounds__(SQS_TPB)
k_sort_emul_q_seed(const mb128 *__restrict__ in, mb128 *__restrict__ wbuf, mb128 *__restrict__ out, const int64_t *__restrict__ a_roff, const int64_t *__restrict__ w_roff,
                   const int32_t *__restrict__ tie_list, const int32_t *__restrict__ n_tie, int32_t *__restrict__ tie_pos, const int32_t *__restrict__ tie_n,
                   EmulQ q, int n_lo, int32_t *__restrict__ scr_pool /* one int per anchor of the long reads */)
{
	__shared__ int s_hist[256];
	__shared__ unsigned long long s_diff;
	__shared__ int s_w[SQS_TPB / 32], s_d[2];
	const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
	const unsigned FULL = 0xffffffffu;
	for (int t = blockIdx.x; t < *n_tie; t += gridDim.x) {
		const int r = tie_list[t];
		const int64_t base = a_roff[r];
		const int n = (int)(a_roff[r + 1] - base);
		if (n <= n_lo) continue;   // short reads: k_sort_emul, in shared memory
		mb128 *W = wbuf + w_roff[r], *O = out + base;
		const mb128 *src = in + base;
		int32_t *T = tie_pos + (size_t)r * TL_MAX;
		const int nt = tie_n[r];
		__syncthreads();
		if (tid == 0) {
			if (nt <= TL_MAX) // ascending positions (collected in arbitrary order)
				for (int i = 1; i < nt; ++i) { const int v = T[i]; int j = i; while (j > 0 && T[j - 1] > v) { T[j] = T[j - 1]; --j; } T[j] = v; }
			s_diff = 0; s_d[0] = 256; s_d[1] = -1;
		}
		s_hist[tid] = 0;
		__syncthreads();
		{ // first byte level on which the keys differ
			const uint64_t k0 = src[0].x; uint64_t d = 0;
			for (int i = tid; i < n; i += SQS_TPB) d |= src[i].x ^ k0;
			#pragma unroll
			for (int o = 16; o > 0; o >>= 1) d |= __shfl_xor_sync(FULL, d, o);
			if (lane == 0 && d) atomicOr(&s_diff, (unsigned long long)d);
		}
		__syncthreads();
		int s = 56;
		{ const uint64_t diff = s_diff; while (s > 0 && !(diff >> s & 255)) s -= 8; }
		for (int i = tid; i < n; i += SQS_TPB) atomicAdd(&s_hist[(int)(src[i].x >> s & 255)], 1);
		__syncthreads();
		const int n_ne = __syncthreads_count(s_hist[tid] > 0);
		if (n_ne != 2 || s == 0) { // general first level: walked like every other range
			for (int i = tid; i < n; i += SQS_TPB) W[i] = src[i];
			if (tid == 0) sq_push(q, r, 0, n, 56);
			continue;
		}
		if (s_hist[tid] > 0) { atomicMin(&s_d[0], tid); atomicMax(&s_d[1], tid); }
		__syncthreads();
		const int d0 = s_d[0], d1 = s_d[1], c0 = s_hist[d0];
		const int m_max = c0 < n - c0 ? c0 : n - c0;
		int32_t *ppos = scr_pool + w_roff[r], *qpos = ppos + m_max;
		// ranks of the misplaced elements of either region; sweep() calls f(i, flag, exclusive rank) for lo <= i < hi in order
		auto sweep = [&](int lo, int hi, int bad, auto f) {
			int carry = 0;
			for (int c = lo; c < hi; c += SQS_TPB) {
				const int i = c + tid;
				const bool flag = i < hi && (int)(src[i].x >> s & 255) == bad;
				const unsigned bal = __ballot_sync(FULL, flag);
				if (lane == 0) s_w[wid] = __popc(bal);
				__syncthreads();
				int woff = 0, tot = 0;
				#pragma unroll
				for (int w = 0; w < SQS_TPB / 32; ++w) { const int v = s_w[w]; woff += w < wid ? v : 0; tot += v; }
				if (i < hi) f(i, flag, carry + woff + __popc(bal & ((1u << lane) - 1)));
				__syncthreads();
				carry += tot;
			}
			return carry;
		};
		const int m = sweep(0, c0, d1, [&](int i, bool flag, int rk) { if (flag) ppos[rk] = i; });
		sweep(c0, n, d0, [&](int i, bool flag, int rk) { if (flag) qpos[rk] = i; });
		__syncthreads();
		sweep(0, c0, d1, [&](int i, bool flag, int rk) { W[i] = flag ? src[qpos[rk]] : src[i]; });
		sweep(c0, n, d0, [&](int i, bool, int e) {
			const bool first = i == c0 || (int)(src[i - 1].x >> s & 255) == d0;
			W[i] = e >= m ? src[i] : first ? src[ppos[e]] : src[i - 1];
		});
		__syncthreads();
		if (tid == 0) { // the two buckets: descend where a tied pair sits
			const int s2 = s > 8 ? s - 8 : 0;
			for (int h = 0; h < 2; ++h) {
				const int lo = h ? c0 : 0, hi = h ? n : c0;
				if (hi - lo < 2) continue;
				bool has = nt > TL_MAX;
				for (int j = 0; j < nt && j < TL_MAX && !has; ++j) has = T[j] > lo && T[j] < hi;
				if (!has) continue;
				if (hi - lo > MB_RS_MIN_SIZE) sq_push(q, r, lo, hi, s2);
				else for (int a = lo; a < hi; ++a) { // leaf: stable order of the current arrangement
					const mb128 v = W[a];
					int rank = 0;
					for (int bb2 = lo; bb2 < hi; ++bb2) { const uint64_t xb = W[bb2].x; rank += (xb < v.x) || (xb == v.x && bb2 < a); }
					O[lo + rank] = v;
				}
			}
		}
	}
}

__global__ void __launch_bounds__(SQ_TPB)
k_sort_emul_q(mb128 *__restrict__ out, mb128 *__restrict__ wbuf, const int64_t *__restrict__ a_roff, const int64_t *__restrict__ w_roff,
              const int32_t *__restrict__ tie_pos, const int32_t *__restrict__ tie_n, EmulQ q, int *__restrict__ ends_pool)
{
	extern __shared__ int sq_bb[];                 // [256][SQ_TPB]: bucket heads, one column per lane
	int *bb = sq_bb + threadIdx.x;
	#define BB(k) bb[(k) * SQ_TPB]
	int *be = ends_pool + (size_t)(blockIdx.x * SQ_TPB + threadIdx.x) * 256;   // bucket ends of the lane's current item
	volatile int *ready = q.ready; volatile int *ctr = q.ctr;
	const unsigned FULL = 0xffffffffu;
	const int lane = threadIdx.x & 31;
	// the first slots go to lane 0 of every warp, the next to lane 1, ...: the long first-level walks of a batch spread over
	// all SMs instead of filling the first few CTAs; later claims take the slots behind those
	const int n_thr = gridDim.x * SQ_TPB, n_warp = n_thr >> 5;
	int state = 0, slot = lane * n_warp + ((blockIdx.x * SQ_TPB + threadIdx.x) >> 5);
	mb128 *W = nullptr, *O = nullptr; const int32_t *T = nullptr;
	int r = 0, nt = 0, beg = 0, end = 0, s = 0, i = 0, k = 0, bek = 0, l = 0, ti = 0, prev_end = 0;
	bool in_cycle = false, all_ranges = false, warp_idle = true;
	mb128 tmp; tmp.x = tmp.y = 0;
	uint64_t k0 = 0, diff = 0;
	for (unsigned iter = 0;; ++iter) {
		bool want_leaf = false; int leaf_lo = 0, leaf_hi = 0;
		if (state == 0) { // take an item (polled every 8th iteration while warp-mates work: a poll is two trips to L2)
			if (slot < 0) slot = n_thr + atomicAdd(&q.ctr[1], 1);
			if (slot >= q.cap) state = 9;
			else if (warp_idle || (iter & 7) == 0) {
				if (ready[slot]) {
					__threadfence();
					const int4 it = __ldcg(&q.items[slot]);
					r = it.x, beg = it.y, end = it.z, s = it.w; slot = -1;
					W = wbuf + w_roff[r]; O = out + a_roff[r];
					T = tie_pos + (size_t)r * TL_MAX; nt = tie_n[r]; all_ranges = nt > TL_MAX;
					k0 = sq_ldx(W + beg); diff = 0; i = beg + 1; state = 1;
				} else if (ctr[2] == 0 && !ready[slot]) state = 9;
			}
		} else if (state == 1) { // byte levels on which the whole range agrees move nothing: skip them
			#pragma unroll
			for (int u = 0; u < 8; ++u) if (i < end) { diff |= sq_ldx(W + i) ^ k0; ++i; }
			if (i >= end) {
				while (s > 0 && !(diff >> s & 255)) s -= 8;
				for (int kk = 0; kk < 256; ++kk) BB(kk) = 0;
				i = beg; state = 2;
			}
		} else if (state == 2) { // digit histogram
			#pragma unroll
			for (int u = 0; u < 8; ++u) if (i < end) { BB((int)(sq_ldx(W + i) >> s & 255)) += 1; ++i; }
			if (i >= end) {
				int run = beg;
				for (int kk = 0; kk < 256; ++kk) { const int c = BB(kk); BB(kk) = run; run += c; be[kk] = run; }
				k = 0; bek = be[0]; in_cycle = false; state = 3;
			}
		} else if (state == 3) { // the cycle-leader walk, one element per iteration
			if (!in_cycle) {
				const int bk = BB(k);
				if (bk == bek) {
					if (++k == 256) { k = 0; ti = 0; prev_end = beg; state = 4; }
					else bek = be[k];
				} else {
					tmp = sq_ld(W + bk); l = (int)(tmp.x >> s & 255);
					if (l == k) BB(k) = bk + 1; else in_cycle = true;
				}
			} else {
				const int p = BB(l); BB(l) = p + 1;
				const mb128 nx = sq_ld(W + p);
				sq_st(W + p, tmp);
				tmp = nx; l = (int)(tmp.x >> s & 255);
				if (l == k) { const int bk = BB(k); sq_st(W + bk, tmp); BB(k) = bk + 1; in_cycle = false; }
			}
		} else if (state == 4) { // one bucket per iteration: descend where a tied pair sits, final order there -> out
			const int lo = prev_end, hi = be[k], len = hi - lo;
			prev_end = hi;
			if (len > 1) {
				while (ti < nt && ti < TL_MAX && T[ti] <= lo) ++ti;
				const bool has = all_ranges || (ti < nt && T[ti] < hi);
				if (has) {
					if (s > 0 && len > MB_RS_MIN_SIZE) sq_push(q, r, lo, hi, s > 8 ? s - 8 : 0);
					else if (s > 0) want_leaf = true, leaf_lo = lo, leaf_hi = hi;
					else for (int a = lo; a < hi; ++a) O[a] = sq_ld(W + a);
				}
			}
			if (++k == 256) state = 5;
		}
		// leaves (<= 64 elements; upstream's insertion sort = stable order of the current arrangement), by the whole warp:
		// two elements per lane, ranks by comparing against all of them
		unsigned need = __ballot_sync(FULL, want_leaf);
		while (need) {
			const int src = __ffs(need) - 1; need &= need - 1;
			const mb128 *Wl = reinterpret_cast<const mb128*>(__shfl_sync(FULL, (unsigned long long)W, src));
			mb128 *Ol = reinterpret_cast<mb128*>(__shfl_sync(FULL, (unsigned long long)O, src));
			const int lo = __shfl_sync(FULL, leaf_lo, src), len = __shfl_sync(FULL, leaf_hi, src) - lo;
			mb128 v0, v1; v0.x = v0.y = v1.x = v1.y = 0;
			if (lane < len) v0 = sq_ld(Wl + lo + lane);
			if (lane + 32 < len) v1 = sq_ld(Wl + lo + lane + 32);
			int r0 = 0, r1 = 0;
			for (int b = 0; b < len; ++b) {
				const uint64_t xa = __shfl_sync(FULL, v0.x, b & 31), xc = __shfl_sync(FULL, v1.x, b & 31);
				const uint64_t xb = b < 32 ? xa : xc;
				r0 += (xb < v0.x) || (xb == v0.x && b < lane);
				r1 += (xb < v1.x) || (xb == v1.x && b < lane + 32);
			}
			if (lane < len) Ol[lo + r0] = v0;
			if (lane + 32 < len) Ol[lo + r1] = v1;
		}
		if (state == 5) { __threadfence(); atomicSub(&q.ctr[2], 1); state = 0; }
		if (__all_sync(FULL, state == 9)) break;
		warp_idle = __all_sync(FULL, state == 0 || state == 9);
		if (warp_idle) __nanosleep(300);
	}
	#undef BB
}

// reads whose anchors tie on x: reproduce upstream's unstable radix permutation exactly.  One warp per read, spread over
// the whole GPU: the anchors are staged in shared memory (global scratch beyond SE_SMEM_N), lane 0 replays the sequential
// cycle-leader passes there (byte levels on which the whole range agrees are no-ops and are skipped), and the <= 64-element
// leaves are insertion-sorted by one lane each.
#define SE_SMEM_N 2560             // 2560 * 16 B = 40 KB of anchors per warp
#define SE_STACK  2304
#define SE_SMEM_BYTES (SE_SMEM_N * 16 + 512 * 4 + 64)
#define SE_MID_N 12800             // second launch, one warp per SM with 200 KB: reads up to 12,800 anchors are replayed at shared-memory latency
#define SE_MID_BYTES (SE_MID_N * 16 + 512 * 4 + 64)

__global__ void __launch_bounds__(32)
k_sort_emul(const mb128 *__restrict__ in, mb128 *__restrict__ out, const int64_t *__restrict__ a_roff,
            const int32_t *__restrict__ tie_list, const int32_t *__restrict__ n_tie, int32_t *__restrict__ cursor,
            int *__restrict__ ws_pool /* per CTA: 3*SE_STACK ints of range stack + leaf list */, int smem_n, int n_lo)
{
	// this launch takes the reads with n_lo < n (and stages those with n <= smem_n in shared memory); smaller ones belong to
	// the other launch
	extern __shared__ __align__(16) uint8_t se_smem[];
	mb128 *sa = reinterpret_cast<mb128*>(se_smem);
	int *bb = reinterpret_cast<int*>(se_smem + (size_t)smem_n * 16), *be = bb + 256;
	int *sh = be + 256; // [0] = n_leaf, [1] = work item
	int *stk = ws_pool + (size_t)blockIdx.x * (6 * SE_STACK);
	int *leaf = stk + 3 * SE_STACK; // (beg, end) pairs
	const int lane = threadIdx.x;
	const unsigned FULL = 0xffffffffu;
	for (;;) {
		if (lane == 0) sh[1] = atomicAdd(cursor, 1);
		__syncwarp();
		const int t = sh[1];
		__syncwarp();
		if (t >= *n_tie) break;
		const int r = tie_list[t];
		const int64_t base = a_roff[r];
		const int n = (int)(a_roff[r + 1] - base);
		if (n <= n_lo || n > smem_n) continue; // another launch's read
		mb128 *a = n <= smem_n ? sa : out + base;
		for (int i = lane; i < n; i += 32) a[i] = in[base + i];
		__syncwarp();
		KeyX key;
		if (n <= MB_RS_MIN_SIZE) {
			if (lane == 0) mb_insertsort(a, a + n, key);
		} else {
			// The range stack is walked by the whole warp.  Of one partition pass (radix_emul.cuh mb_rs_partition) only the
			// cycle-leader permutation is order-dependent and stays with lane 0; the digit histogram, the 256-entry prefix sums
			// and the hand-out of sub-ranges are spread over the lanes.  Sub-ranges are disjoint, so the order in which they are
			// pushed and sorted does not change the result.
			int sp = 1, n_leaf = 0;
			if (lane == 0) stk[0] = 0, stk[1] = n, stk[2] = 56;
			__syncwarp();
			while (sp > 0) {
				--sp;
				const int beg = stk[3 * sp], end = stk[3 * sp + 1];
				int s = stk[3 * sp + 2];
				__syncwarp();
				// skip byte levels on which every key of the range agrees: such a pass moves nothing
				const uint64_t k0 = a[beg].x;
				uint32_t dlo = 0, dhi = 0;
				for (int i = beg + 1 + lane; i < end; i += 32) { const uint64_t d = a[i].x ^ k0; dlo |= (uint32_t)d, dhi |= (uint32_t)(d >> 32); }
				#pragma unroll
				for (int d = 16; d > 0; d >>= 1) dlo |= __shfl_xor_sync(FULL, dlo, d), dhi |= __shfl_xor_sync(FULL, dhi, d);
				const uint64_t diff = (uint64_t)dhi << 32 | dlo;
				while (s > 0 && !(diff >> s & 255)) s -= 8;
				// digit histogram and bucket ranges: be[k] = end of bucket k, bb[k] = its start (moving cursor during the permutation)
				for (int k = lane; k < 256; k += 32) be[k] = 0;
				__syncwarp();
				for (int i = beg + lane; i < end; i += 32) atomicAdd(&be[(int)(a[i].x >> s & 255)], 1);
				__syncwarp();
				{
					int c[8], tot = 0;
					#pragma unroll
					for (int u = 0; u < 8; ++u) { c[u] = be[lane * 8 + u]; tot += c[u]; }
					int incl = tot;
					#pragma unroll
					for (int d = 1; d < 32; d <<= 1) { const int v = __shfl_up_sync(FULL, incl, d); if (lane >= d) incl += v; }
					int run = beg + incl - tot;
					__syncwarp();
					#pragma unroll
					for (int u = 0; u < 8; ++u) { bb[lane * 8 + u] = run; run += c[u]; be[lane * 8 + u] = run; }
				}
				__syncwarp();
				if (lane == 0) {
					for (int k = 0; k < 256;) {
						if (bb[k] != be[k]) {
							int l = (int)(a[bb[k]].x >> s & 255);
							if (l != k) {
								mb128 tmp = a[bb[k]], swap;
								do {
									swap = tmp; tmp = a[bb[l]]; a[bb[l]++] = swap;
									l = (int)(tmp.x >> s & 255);
								} while (l != k);
								a[bb[k]++] = tmp;
							} else ++bb[k];
						} else ++k;
					}
				}
				__syncwarp();
				if (s) {
					const int s2 = s > 8 ? s - 8 : 0;
					// bucket k spans [k ? be[k-1] : beg, be[k]); larger than 64 -> back on the stack, 2..64 -> leaf list
					int lo8[8], hi8[8], n_big = 0, n_lf = 0;
					#pragma unroll
					for (int u = 0; u < 8; ++u) {
						const int k = lane * 8 + u;
						lo8[u] = k ? be[k - 1] : beg, hi8[u] = be[k];
						const int len = hi8[u] - lo8[u];
						n_big += len > MB_RS_MIN_SIZE, n_lf += len > 1 && len <= MB_RS_MIN_SIZE;
					}
					int ib = n_big, il = n_lf;
					#pragma unroll
					for (int d = 1; d < 32; d <<= 1) {
						const int vb = __shfl_up_sync(FULL, ib, d), vl = __shfl_up_sync(FULL, il, d);
						if (lane >= d) ib += vb, il += vl;
					}
					const int tot_b = __shfl_sync(FULL, ib, 31), tot_l = __shfl_sync(FULL, il, 31);
					int pb = sp + ib - n_big, pl = n_leaf + il - n_lf;
					#pragma unroll
					for (int u = 0; u < 8; ++u) {
						const int len = hi8[u] - lo8[u];
						if (len > MB_RS_MIN_SIZE) { stk[3 * pb] = lo8[u], stk[3 * pb + 1] = hi8[u], stk[3 * pb + 2] = s2; ++pb; } // depth <= 8 levels x 255 siblings < SE_STACK
						else if (len > 1) {
							if (2 * pl + 2 <= 3 * SE_STACK) { leaf[2 * pl] = lo8[u], leaf[2 * pl + 1] = hi8[u]; }
							else mb_insertsort(a + lo8[u], a + hi8[u], key);
							++pl;
						}
					}
					sp += tot_b, n_leaf += tot_l;
				}
				__syncwarp();
			}
			if (n_leaf > (3 * SE_STACK) / 2) n_leaf = (3 * SE_STACK) / 2; // the overflow was sorted in place above
			for (int l = lane; l < n_leaf; l += 32) mb_insertsort(a + leaf[2 * l], a + leaf[2 * l + 1], key);
		}
		__syncwarp();
		if (a == sa) for (int i = lane; i < n; i += 32) out[base + i] = sa[i];
		__syncwarp();
	}
}

struct SeedOut {
	mb128 *a = nullptr;            // sorted anchors, per read contiguous
	mb128 *a_unsorted = nullptr;
	int64_t *a_roff = nullptr;     // [n_reads+1]
	int32_t *rep_len = nullptr;    // [n_reads]
	int32_t *read_perm = nullptr;  // [n_reads] reads by descending anchor count (nullptr for small batches)
	int64_t n_a = 0;
};


static void run_seed(Arena &ar, cudaStream_t st, const DevIndex &ix, int mid_occ, const mb128 *mini, const int64_t *mini_off, int64_t n_mini,
                     const int64_t *d_read_off, int n_reads, SeedOut &o, int64_t *n_launch, int num_sms)
{
	o.a_roff = ar.get<int64_t>(n_reads + 1);
	o.rep_len = ar.get<int32_t>(n_reads);
	if (n_mini == 0) {
		CK(cudaMemsetAsync(o.a_roff, 0, (n_reads + 1) * sizeof(int64_t), st));
		CK(cudaMemsetAsync(o.rep_len, 0, (n_reads ? n_reads : 1) * sizeof(int32_t), st));
		o.a = o.a_unsorted = ar.get<mb128>(1); o.n_a = 0;
		return;
	}
	const bool dbg = getenv("MB_DEBUG") != nullptr;
	cudaEvent_t dbg_ev[2] = {nullptr, nullptr};
	if (dbg) { cudaEventCreate(&dbg_ev[0]); cudaEventCreate(&dbg_ev[1]); cudaEventRecord(dbg_ev[0], st); }
	auto mark = [&](const char *what) { // MB_DEBUG: time since the previous mark (synchronises)
		if (!dbg) return;
		cudaEventRecord(dbg_ev[1], st); cudaEventSynchronize(dbg_ev[1]);
		float ms = 0; cudaEventElapsedTime(&ms, dbg_ev[0], dbg_ev[1]);
		fprintf(stderr, "[mb]   seed: %-22s %9.3f ms\n", what, ms);
		cudaEventRecord(dbg_ev[0], st);
	};
	int32_t *occ = ar.get<int32_t>(n_mini), *cnt = ar.get<int32_t>(n_mini);
	uint64_t *val = ar.get<uint64_t>(n_mini);
	int64_t *a_off = ar.get<int64_t>(n_mini + 1);
	k_seed_lookup<<<(unsigned)cdiv(n_mini, 256), 256, 0, st>>>(ix, mini, n_mini, mid_occ, occ, val, cnt); ++*n_launch;
	k_rep_len<<<(unsigned)cdiv((int64_t)n_reads * 32, 128), 128, 0, st>>>(mini, mini_off, occ, n_reads, mid_occ, o.rep_len); ++*n_launch;
	exclusive_scan<int32_t>(ar, st, cnt, a_off, n_mini, n_launch);
	k_read_anchor_off<<<(unsigned)cdiv(n_reads + 1, 128), 128, 0, st>>>(mini_off, a_off, n_reads, o.a_roff); ++*n_launch;
	mark("lookup+scan");
	int64_t n_a = 0;
	CK(cudaMemcpyAsync(&n_a, a_off + n_mini, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
	CK(cudaStreamSynchronize(st));
	o.n_a = n_a;
	o.a_unsorted = ar.get<mb128>(n_a + 1);
	o.a = ar.get<mb128>(n_a + 1);
	if (n_a == 0) return;
	k_seed_fill<<<(unsigned)cdiv(n_mini, 256), 256, 0, st>>>(ix, mini, mini_off, n_mini, d_read_off, cnt, val, a_off, o.a_unsorted); ++*n_launch;
	mark("fill");
	// reads too large for the shared-memory sort: listed, and given a slice of a ping-pong buffer for the global-memory radix sort
	int32_t *big_sz = ar.get<int32_t>(n_reads);
	int64_t *big_off = ar.get<int64_t>(n_reads + 1);
	int32_t *big_list = ar.get<int32_t>(n_reads);
	int32_t *ctr = ar.get<int32_t>(4);     // [0] reads with tied keys, [1] cursor of k_sort_emul, [2] big reads, [3] cursor of k_sort_anchors_big
	CK(cudaMemsetAsync(ctr, 0, 4 * sizeof(int32_t), st));
	k_big_sort_list<<<(unsigned)cdiv(n_reads, 128), 128, 0, st>>>(o.a_roff, n_reads, big_sz, big_list, ctr + 2); ++*n_launch;
	exclusive_scan<int32_t>(ar, st, big_sz, big_off, n_reads, n_launch);
	int64_t big_total = 0;
	CK(cudaMemcpyAsync(&big_total, big_off + n_reads, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
	CK(cudaStreamSynchronize(st));
	mb128 *big_tmp = ar.get<mb128>(big_total + 1);
	int32_t *tie_list = ar.get<int32_t>(n_reads);
	int32_t *tie_pos = ar.get<int32_t>(big_total > 0 ? (size_t)n_reads * TL_MAX : 1), *tie_n = ar.get<int32_t>(n_reads);
	if (n_reads > 1024) {
		int32_t *hist = ar.get<int32_t>(RP_NB);
		o.read_perm = ar.get<int32_t>(n_reads);
		CK(cudaMemsetAsync(hist, 0, RP_NB * sizeof(int32_t), st));
		k_rp_hist<<<(unsigned)cdiv(n_reads, 256), 256, 0, st>>>(o.a_roff, n_reads, hist);
		k_rp_scan<<<1, RP_NB, 0, st>>>(hist);
		k_rp_scatter<<<(unsigned)cdiv(n_reads, 256), 256, 0, st>>>(o.a_roff, n_reads, hist, o.read_perm);
		*n_launch += 3;
	}
	if (big_total > 0) { // byte positions of x that can differ: position bytes, contig-id bytes, the strand byte
		unsigned pass_mask = 0x80u;
		uint32_t max_len = ix.max_seq_len ? ix.max_seq_len : 0xffffffffu;
		for (int b = 0; b < 4; ++b) if (b == 0 || (max_len >> (8 * b)) != 0) pass_mask |= 1u << b;
		uint32_t max_rid = ix.n_seq > 0 ? (uint32_t)(ix.n_seq - 1) : 0;
		for (int b = 0; b < 4; ++b) if ((b == 0 && max_rid) || (b > 0 && (max_rid >> (8 * b)) != 0)) pass_mask |= 1u << (4 + b);
		k_sort_anchors_big<<<num_sms * 4, SB_TPB, 0, st>>>(o.a_unsorted, o.a, big_tmp, o.a_roff, big_list, ctr + 2, ctr + 3, big_off, pass_mask, tie_list, ctr, tie_pos, tie_n); ++*n_launch;
		mark("radix sort (long)");
	}
	int grid = n_reads < num_sms * 8 ? n_reads : num_sms * 8;
	k_sort_anchors<<<grid, SORT_TPB, 0, st>>>(o.a_unsorted, o.a, o.a_roff, n_reads, tie_list, ctr, o.read_perm); ++*n_launch;
	mark("bitonic sort (short)");
	static bool se_attr = false;
	if (!se_attr) {
		CK(cudaFuncSetAttribute(k_sort_emul, cudaFuncAttributeMaxDynamicSharedMemorySize, SE_MID_BYTES));
		CK(cudaFuncSetAttribute(k_sort_emul_q, cudaFuncAttributeMaxDynamicSharedMemorySize, 256 * SQ_TPB * (int)sizeof(int)));
		se_attr = true;
	}
	if (big_total > 0) { // mid-size reads with ties (the long ones of a 50 Mb database): whole replay in 200 KB of shared memory, one warp per SM
		int *ws_mid = ar.get<int>((size_t)num_sms * 6 * SE_STACK);
		int32_t *cur_mid = ar.get<int32_t>(1);
		CK(cudaMemsetAsync(cur_mid, 0, sizeof(int32_t), st));
		k_sort_emul<<<num_sms, 32, SE_MID_BYTES, st>>>(o.a_unsorted, o.a, o.a_roff, tie_list, ctr, cur_mid, ws_mid, SE_MID_N, SE_SMEM_N); ++*n_launch;
		mark("replay (mid, shared memory)");
	}
	if (big_total > 0) { // long reads with ties: queue of ranges, one lane per range (the long walks start first)
		EmulQ q;
		q.cap = (int)std::min<int64_t>(big_total / 8 + n_reads + 2 * (int64_t)num_sms * SQ_TPB + 4096, 0x7fffff00);
		q.items = ar.get<int4>(q.cap); q.ready = ar.get<int>(q.cap); q.ctr = ar.get<int>(4);
		int *ends_pool = ar.get<int>((size_t)num_sms * SQ_TPB * 256);
		int32_t *scr_pool = ar.get<int32_t>(big_total + 1);
		CK(cudaMemsetAsync(q.ready, 0, (size_t)q.cap * sizeof(int), st));
		CK(cudaMemsetAsync(q.ctr, 0, 4 * sizeof(int), st));
		k_sort_emul_q_seed<<<num_sms * 4, SQS_TPB, 0, st>>>(o.a_unsorted, big_tmp, o.a, o.a_roff, big_off, tie_list, ctr, tie_pos, tie_n, q, SE_MID_N, scr_pool);
		k_sort_emul_q<<<num_sms, SQ_TPB, 256 * SQ_TPB * sizeof(int), st>>>(o.a, big_tmp, o.a_roff, big_off, tie_pos, tie_n, q, ends_pool);
		*n_launch += 2;
		mark("replay (long, queued)");
	}
	const int se_grid = num_sms * 5;   // 43 KB of shared memory per one-warp CTA: five fit an SM
	int *ws_pool = ar.get<int>((size_t)se_grid * 6 * SE_STACK);
	k_sort_emul<<<se_grid, 32, SE_SMEM_BYTES, st>>>(o.a_unsorted, o.a, o.a_roff, tie_list, ctr, ctr + 1, ws_pool, SE_SMEM_N, 0); ++*n_launch;
	mark("replay (short)");
	if (dbg) { cudaEventDestroy(dbg_ev[0]); cudaEventDestroy(dbg_ev[1]); }
	if (getenv("MB_DEBUG")) {
		int32_t h = 0; cudaMemcpyAsync(&h, ctr, sizeof(h), cudaMemcpyDeviceToHost, st); cudaStreamSynchronize(st);
		fprintf(stderr, "[mb] reads with tied anchor keys: %d of %d\n", h, n_reads);
	}
}
