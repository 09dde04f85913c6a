// chain.cuh -- K3: anchor chaining DP (minimap2-2.17 chain.c mm_chain_dp) on sm_100a.
//
// Reached in the reference through mappy's Aligner.map (/root/reference/monica/genomes/aligner.py:193,215
// -> mm_map_frag -> mm_chain_dp).
//
// One warp per read.  Anchors i are processed in order (the recurrence on f[] is sequential in i); for each i
// the predecessor window j = i-1 .. st is evaluated 32 at a time.  All 32 candidate scores are independent;
// what is sequential in upstream's inner loop is the (max_f, n_skip, t[]) automaton that implements the
// max_skip early exit.  It is resolved per 32-chunk with warp primitives:
//   * t[]-marks of the chunk are written first (a mark can only target a j visited later, and marks written by
//     lanes past the break position are never consulted, so writing all of them up front is equivalent),
//   * an exclusive prefix max gives, per lane, whether it would have improved max_f,
//   * n_skip is a saturating counter: each lane applies n -> max(n+a, b) with (a,b) = (-1,0) on improve,
//     (+1,-inf) on a marked non-improving j, identity otherwise; these compose associatively, so one warp scan
//     yields n_skip after every lane and a ballot finds the first lane where it exceeds max_skip.
// Results (f, p, v) are bit-identical to the sequential loop.
#pragma once
#include "common.cuh"
#include "radix_emul.cuh"

#define CH_WARPS 4

MB_D int mb_ilog2_32(uint32_t v) { return 31 - __clz(v); }

__global__ void __launch_bounds__(CH_WARPS * 32)
k_chain_dp(const mb128 *__restrict__ a, const int64_t *__restrict__ a_roff, int n_reads,
           int max_dist_x, int max_dist_y, int bw, int max_skip, int max_iter,
           int32_t *__restrict__ f, int32_t *__restrict__ p, int32_t *__restrict__ v, int32_t *__restrict__ t,
           int32_t *__restrict__ work_ctr, unsigned long long *__restrict__ cells_out,
           const int32_t *__restrict__ perm)
{
	// `perm`: the order in which the warps draw the reads (longest first, see k_rp_hist), or nullptr
	const int lane = threadIdx.x & 31;
	const unsigned FULL = 0xffffffffu;
	unsigned long long cells = 0;
	for (;;) {
		int r = 0;
		if (lane == 0) r = atomicAdd(work_ctr, 1);
		r = __shfl_sync(FULL, r, 0);
		if (r >= n_reads) break;
		if (perm) r = perm[r];
		const int64_t base = a_roff[r];
		const int n = (int)(a_roff[r + 1] - base);
		if (n == 0) continue;
		const mb128 *A = a + base;
		int32_t *F = f + base, *P = p + base, *V = v + base, *T = t + base;
		// avg_qspan = (float)sum_qspan / n
		unsigned long long sum = 0;
		for (int i = lane; i < n; i += 32) { sum += (A[i].y >> 32) & 0xff; T[i] = 0; }
		#pragma unroll
		for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(FULL, sum, d);
		const float avg_qspan = __fdiv_rn((float)sum, (float)(int64_t)n);
		const double avg_d = (double)avg_qspan;
		__syncwarp();
		int st = 0;
		for (int i = 0; i < n; ++i) {
			const mb128 ai = A[i];
			const uint64_t ri = ai.x;
			const int32_t qi = (int32_t)ai.y, q_span = (int32_t)(ai.y >> 32 & 0xff);
			// while (st < i && ri > a[st].x + max_dist_x) ++st;   (monotone predicate -> count leading trues)
			for (;;) {
				int idx = st + lane;
				bool c = idx < i && ri > A[idx].x + (uint64_t)max_dist_x;
				unsigned m = __ballot_sync(FULL, c);
				if (m == FULL) st += 32;
				else { st += __ffs(~m) - 1; break; }
			}
			if (i - st > max_iter) st = i - max_iter;
			int32_t max_f = q_span, max_j = -1, n_skip = 0;
			bool done = false;
			for (int jb = i - 1; jb >= st && !done; jb -= 32) {
				const int j = jb - lane;
				const bool act = j >= st;
				bool valid = false;
				int32_t sc = INT32_MIN, pj = -1;
				if (act) {
					const mb128 aj = A[j];
					const int64_t dr = (int64_t)(ri - aj.x);
					const int32_t dq = qi - (int32_t)aj.y;
					pj = P[j];
					if (!(dr == 0 || dq <= 0) && !(dq > max_dist_y || dq > max_dist_x)) {
						const int32_t dd = dr > dq ? (int32_t)(dr - dq) : (int32_t)(dq - dr);
						if (dd <= bw) {
							const int32_t min_d = dq < dr ? dq : (int32_t)dr;
							valid = true;
							sc = min_d > q_span ? q_span : min_d;
							const int log_dd = dd ? mb_ilog2_32((uint32_t)dd) : 0;
							sc -= (int)__dmul_rn(__dmul_rn((double)dd, .01), avg_d) + (log_dd >> 1);
							sc += F[j];
						}
					}
				}
				if (valid && pj >= 0) T[pj] = i;
				__syncwarp();
				const bool mark = valid && T[j] == i;
				// exclusive prefix max over lanes (lane 0 = first visited)
				int32_t incl = sc;
				#pragma unroll
				for (int d = 1; d < 32; d <<= 1) { int32_t o = __shfl_up_sync(FULL, incl, d); if (lane >= d) incl = incl > o ? incl : o; }
				int32_t excl = __shfl_up_sync(FULL, incl, 1);
				if (lane == 0) excl = INT32_MIN;
				excl = excl > max_f ? excl : max_f;
				const bool improve = valid && sc > excl;
				// saturating n_skip counter as composable (a,b): n -> max(n+a, b)
				int32_t ca = improve ? -1 : (mark ? 1 : 0), cb = improve ? 0 : -(1 << 28);
				#pragma unroll
				for (int d = 1; d < 32; d <<= 1) {
					int32_t oa = __shfl_up_sync(FULL, ca, d), ob = __shfl_up_sync(FULL, cb, d);
					if (lane >= d) { int32_t nb = ob + ca; cb = nb > cb ? nb : cb; ca = oa + ca; }
				}
				int32_t n_after = n_skip + ca; n_after = n_after > cb ? n_after : cb;
				const bool brk = valid && !improve && mark && n_after > max_skip;
				const unsigned bm = __ballot_sync(FULL, brk);
				const int last = bm ? __ffs(bm) - 1 : 31;
				const int32_t best = __shfl_sync(FULL, incl, last);
				if (best > max_f) {
					unsigned wm = __ballot_sync(FULL, valid && sc == best && lane <= last);
					max_j = jb - (__ffs(wm) - 1);
					max_f = best;
				}
				n_skip = __shfl_sync(FULL, n_after, last);
				int n_act = jb - st + 1; if (n_act > 32) n_act = 32;
				cells += (unsigned)(last + 1 < n_act ? last + 1 : n_act);
				if (bm) done = true;
			}
			if (lane == 0) {
				F[i] = max_f, P[i] = max_j;
				V[i] = (max_j >= 0 && V[max_j] > max_f) ? V[max_j] : max_f;
			}
			__syncwarp();
		}
	}
	if (lane == 0 && cells_out && cells) atomicAdd(cells_out, cells);
}

// ---- exact sorts used by the per-read glue ----
// Upstream sorts these small arrays with radix_sort_128x / radix_sort_64, which is an insertion sort up to 64 elements and an
// UNSTABLE in-place MSD radix sort beyond; where keys can tie, the permutation of equal keys is part of the result.  Reads
// with more than 64 chains / regions are therefore routed to the "big" variants of the per-read kernels (one-warp CTAs that
// own MB_RS_WS_INTS ints of shared-memory scratch, `ws`), which replay upstream's radix passes exactly (radix_emul.cuh);
// everything else never leaves the insertion-sort range and passes ws = nullptr.
template <typename T, typename K>
MB_HD void mb_sort_exact(T *a, int n, K key, int *ws)
{
	if (n <= MB_RS_MIN_SIZE || ws == nullptr) mb_insertsort(a, a + n, key);
	else mb_radix_sort_emul(a, n, ws, key);
}

// chain.c mm_chain_dp(), part 2: chain ends, greedy backtrack, emit chains ordered by first-anchor x.
// In: a (sorted anchors), f, p, v; scratch: t (int32[n]), b (mb128[n]), u (u64[n]), scr (u64[3n+3]).
// Out: a overwritten with chained anchors (chain by chain), u[0..n_u) = score<<32|cnt; returns n_u.
// Warp-cooperative: the array sweeps (marking, chain-end collection, anchor copies) are spread over
// the lanes, the two inherently sequential parts (greedy backtrack over the score-sorted chain ends, the small sorts) run on
// lane 0.  One warp per read also keeps a long read from stalling 31 neighbours in lock-step.  Same outputs as the scalar
// routine: the chain ends are collected in arbitrary order, but their keys are distinct, so the sort that follows is unique.
// Returns -1 (nothing modified that a re-run would not recompute) when the read has more than 64 chain ends and no `ws`:
// the caller then hands the read to the big variant of its kernel.
MB_D int mb_chain_backtrack_warp(int n, mb128 *a, const int32_t *f, const int32_t *p, int32_t *v, int32_t *t, mb128 *b, uint64_t *u, uint64_t *scr,
                                 int min_cnt, int min_sc, int *ws, int lane)
{
	const unsigned FULL = 0xffffffffu;
	if (n == 0) return 0;
	for (int i = lane; i < n; i += 32) t[i] = 0;
	__syncwarp();
	for (int i = lane; i < n; i += 32) if (p[i] >= 0) t[p[i]] = 1;
	__syncwarp();
	int n_u = 0;
	for (int i0 = 0; i0 < n; i0 += 32) {
		const int i = i0 + lane;
		uint64_t key = 0; bool have = false;
		if (i < n && t[i] == 0 && v[i] >= min_sc) {
			int j = i;
			while (j >= 0 && f[j] < v[j]) j = p[j];
			if (j < 0) j = i;
			key = (uint64_t)(uint32_t)f[j] << 32 | (uint32_t)j; have = true;
		}
		const unsigned m = __ballot_sync(FULL, have);
		if (have) u[n_u + __popc(m & ((1u << lane) - 1))] = key;
		n_u += __popc(m);
	}
	__syncwarp();
	if (n_u == 0) return 0;
	if (n_u > MB_RS_MIN_SIZE && ws == nullptr) return -1;
	if (lane == 0) {
		mb_sort_exact(u, n_u, KeyU64(), ws);
		for (int i = 0; i < n_u >> 1; ++i) { uint64_t tt = u[i]; u[i] = u[n_u - i - 1], u[n_u - i - 1] = tt; }
	}
	for (int i = lane; i < n; i += 32) t[i] = 0;
	__syncwarp();
	int k_chains = 0;
	if (lane == 0) {
		int n_v = 0, k = 0;
		for (int i = 0; i < n_u; ++i) {
			const int n_v0 = n_v, k0 = k;
			int j = (int32_t)u[i];
			do {
				v[n_v++] = j;
				t[j] = 1;
				j = p[j];
			} while (j >= 0 && t[j] == 0);
			if (j < 0) {
				if (n_v - n_v0 >= min_cnt) u[k++] = u[i] >> 32 << 32 | (uint32_t)(n_v - n_v0);
			} else if ((int32_t)(u[i] >> 32) - f[j] >= min_sc) {
				if (n_v - n_v0 >= min_cnt) u[k++] = (uint64_t)((u[i] >> 32) - (uint64_t)(int64_t)f[j]) << 32 | (uint32_t)(n_v - n_v0);
			}
			if (k0 == k) n_v = n_v0;
		}
		k_chains = k;
	}
	k_chains = __shfl_sync(FULL, k_chains, 0);
	n_u = k_chains;
	__syncwarp();
	// b = anchors chain by chain (each chain reversed into ascending order)
	{
		int k = 0;
		for (int i = 0; i < n_u; ++i) {
			const int ni = (int32_t)u[i];
			for (int j = lane; j < ni; j += 32) b[k + j] = a[v[k + (ni - j - 1)]];
			k += ni;
		}
	}
	__syncwarp();
	mb128 *w = (mb128*)scr;
	uint64_t *u2 = scr + 2 * (size_t)n_u;
	if (lane == 0) {
		int k = 0;
		for (int i = 0; i < n_u; ++i) {
			w[i].x = b[k].x, w[i].y = (uint64_t)k << 32 | (uint32_t)i;
			k += (int32_t)u[i];
		}
		mb_sort_exact(w, n_u, KeyX(), ws);
		for (int i = 0; i < n_u; ++i) u2[i] = u[(int32_t)w[i].y];
	}
	__syncwarp();
	{
		int k = 0;
		for (int i = 0; i < n_u; ++i) {
			const int nn = (int32_t)u2[i];
			const mb128 *src = b + (w[i].y >> 32);
			for (int j = lane; j < nn; j += 32) a[k + j] = src[j];
			k += nn;
		}
	}
	__syncwarp();
	for (int i = lane; i < n_u; i += 32) u[i] = u2[i];
	__syncwarp();
	return n_u;
}
