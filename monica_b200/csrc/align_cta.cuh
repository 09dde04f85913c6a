// align_cta.cuh -- k_dp_cta: the block-exact ksw_extd2 kernel of align.cuh (k_dp) with one CTA of 256 threads per task.
//
// k_dp gives a task to ONE warp and walks its anti-diagonals 32 cells at a time; a band-limited extension over a 5 kb
// window is ~10,000 dependent diagonals of up to 24 such steps each, tens of milliseconds for a single warp, and those
// few long tasks are what the other DP kernels end up waiting for (second passes, Z-drop split rounds, the inversion
// pass: phases in which nothing else is running).  Here the cells of a diagonal are spread over 256 threads (three cells
// per thread at w = 751), the per-diagonal state lives in the same shared-memory arrays with upstream's layout, and a
// diagonal costs two barriers: every thread first LOADS what its cells need (own column state, the left neighbour's x, v,
// x2 from before this diagonal, the stored score), then computes and STORES.  The semantics are k_dp's, statement for
// statement: 16-aligned block ranges, stale scores outside the fill range, exact per-diagonal maximum with upstream's
// SSE-lane tie order, Z-drop, approximate H0 walk.  The traceback is done by warp 0, which fetches 32 cells of the current
// diagonal per round trip.
#pragma once
#include "align.cuh"
#include "dp_fast.cuh"

#define DPC_THREADS 256
#define DPC_SMEM_MAX (100 * 1024)
// WIN: the per-diagonal state only lives inside the band, [st - 1, en + 16] of the current diagonal, and both ends only move
// right: with WIN = true the state arrays (and the target / query codes) are circular buffers of DPC_WIN entries indexed by
// t & (DPC_WIN - 1) (query: j & (DPC_WIN - 1)), topped up with pristine entries one diagonal ahead.  A 5 kb x 10 kb end
// extension then keeps its state in 14 KB of shared memory instead of 125 KB of global memory.  Needs w + 64 < DPC_WIN.
#define DPC_WIN 1024
#define DPC_WIN_SMEM (DPC_WIN * 9 + DPC_WIN * 4 + 64)
#define DPC2_THREADS 224             // k_dp_cta2: 4 cells per thread, 224 x 4 = 896 >= the widest block range (w + 1 + 30 + 16)

template <bool WIN>
__global__ void __launch_bounds__(DPC_THREADS)
k_dp_cta(DpTask *__restrict__ tasks, const int32_t *__restrict__ order, const int32_t *__restrict__ n_order, int32_t *__restrict__ work_ctr,
         const uint8_t *__restrict__ codes, const uint32_t *__restrict__ S, const uint8_t *__restrict__ pool,
         uint8_t *__restrict__ p_scr, size_t p_stride, int8_t *__restrict__ g_ws, size_t g_stride, int32_t *__restrict__ h_scr, size_t h_stride,
         uint32_t *__restrict__ cigar_pool, DpScoring sc, unsigned long long *__restrict__ cells_out, int smem_bytes)
{
	extern __shared__ __align__(16) int8_t dpc_smem[];
	__shared__ long long s_red[DPC_THREADS / 32];
	__shared__ int s_task;
	const unsigned FULL = 0xffffffffu;
	const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
	uint8_t *P = p_scr + (size_t)blockIdx.x * p_stride;
	unsigned long long cells = 0;
	const int n_total = *n_order;
	for (;;) {
		__syncthreads();
		if (tid == 0) s_task = atomicAdd(work_ctr, 1);
		__syncthreads();
		const int oi = s_task;
		if (oi >= n_total) break;
		DpTask &T = tasks[order[oi]];
		const int qlen = T.qlen, tlen = T.tlen, flag = T.flag, zdrop = T.zdrop, end_bonus = T.end_bonus;
		if (tid == 0) dp_reset(T);
		if (T.skip) { if (tid == 0) T.zdropped = 1; continue; }
		if (qlen <= 0 || tlen <= 0) continue;
		int8_t q = sc.q, e = sc.e, q2 = sc.q2, e2 = sc.e2;
		if (q2 + e2 < q + e) { int8_t t_ = q; q = q2, q2 = t_; t_ = e; e = e2, e2 = t_; }
		const int qe = q + e, qe2 = q2 + e2;
		int w = T.w;
		if (w < 0) w = tlen > qlen ? tlen : qlen;
		const int tlen_ = (tlen + 15) / 16, qlen_ = (qlen + 15) / 16;
		int n_col_ = qlen < tlen ? qlen : tlen;
		n_col_ = ((n_col_ < w + 1 ? n_col_ : w + 1) + 15) / 16 + 1;
		const int ncol16 = n_col_ * 16;
		const int T16 = tlen_ * 16;
		const bool with_exact = !(flag & MB_EZ_APPROX_MAX);
		const bool right = (flag & MB_EZ_RIGHT) != 0;
		int long_thres = e != e2 ? (q2 - q) / (e - e2) - 1 : 0;
		if (q2 + e2 + long_thres * e2 > q + e + long_thres * e) ++long_thres;
		const int long_diff = long_thres * (e - e2) - (q2 - q) - e2;
		const size_t ws_need = (size_t)T16 * 8 + (size_t)qlen_ * 16 + 16;
		const size_t h_need = with_exact ? (size_t)T16 * 4 : 0;
		const bool in_smem = WIN || ((ws_need + 15) & ~(size_t)15) + h_need <= (size_t)smem_bytes;
		const int A = WIN ? DPC_WIN : T16;   // entries per state array
		int8_t *ws = in_smem ? dpc_smem : g_ws + (size_t)blockIdx.x * g_stride;
		int32_t *H = WIN ? reinterpret_cast<int32_t*>(dpc_smem + DPC_WIN * 9)
		           : (in_smem && with_exact) ? reinterpret_cast<int32_t*>(ws + ((ws_need + 15) & ~(size_t)15)) : h_scr + (size_t)blockIdx.x * h_stride;
		int8_t *u = ws, *v = u + A, *x = v + A, *y = x + A, *x2 = y + A, *y2 = x2 + A, *s = y2 + A;
		uint8_t *sf = (uint8_t*)(s + A), *qr = sf + A;   // WIN: qr holds the query by position j (not reversed), circular
		const int8_t I1 = (int8_t)(-q - e), I2 = (int8_t)(-q2 - e2);
		QView qv; qv.codes = T.q_comp == 2 ? pool : codes; qv.idx0 = T.q_idx0; qv.step = T.q_step; qv.comp = T.q_comp == 1;
		TView tv; tv.S = S; tv.bytes = pool; tv.idx0 = T.t_idx0; tv.step = T.t_step; tv.packed = T.t_packed;
		auto ix = [](int t) { return WIN ? (t & (DPC_WIN - 1)) : t; };
		auto fresh_t = [&](int t) { // a pristine entry, as upstream's arrays are before a diagonal touches them
			const int k = ix(t);
			u[k] = I1, v[k] = I1, x[k] = I1, y[k] = I1, x2[k] = I2, y2[k] = I2, s[k] = 0;
			sf[k] = t < tlen ? (uint8_t)tv.at(t) : 0;
			if (with_exact) H[k] = MB_KSW_NEG_INF;
		};
		int hi_t = (WIN && T16 > DPC_WIN ? DPC_WIN : T16) - 1;            // entries 0..hi_t are set up
		int hi_j = WIN ? (qlen > DPC_WIN ? DPC_WIN : qlen) - 1 : 0;
		for (int t = tid; t <= hi_t; t += DPC_THREADS) fresh_t(t);
		if (WIN) { for (int j = tid; j <= hi_j; j += DPC_THREADS) qr[j] = (uint8_t)qv.at(j); }
		else for (int t = tid; t < qlen_ * 16 + 16; t += DPC_THREADS) qr[t] = t < qlen ? (uint8_t)qv.at(qlen - 1 - t) : 0;
		__syncthreads();
		// ez state: every thread keeps the same copy
		int ez_max = 0, ez_max_t = -1, ez_max_q = -1, ez_mqe = MB_KSW_NEG_INF, ez_mqe_t = -1, ez_mte = MB_KSW_NEG_INF, ez_score = MB_KSW_NEG_INF, ez_zdropped = 0;
		int32_t H0 = 0, last_H0_t = 0;
		int last_st = -1, last_en = -1;
		const int n_rows = qlen + tlen - 1;
		auto bnd = [&](int r) { return (int8_t)(r == 0 ? -q - e : r < long_thres ? -e : r == long_thres ? long_diff : -e2); };
		for (int r = 0; r < n_rows; ++r) {
			int st = 0, en = tlen - 1;
			if (st < r - qlen + 1) st = r - qlen + 1;
			if (en > r) en = r;
			if (st < (r - w + 1) >> 1) st = (r - w + 1) >> 1;
			if (en > (r + w) >> 1) en = (r + w) >> 1;
			if (st > en) { ez_zdropped = 1; break; }
			const int st0 = st, en0 = en;
			st = st / 16 * 16, en = (en + 16) / 16 * 16 - 1;
			const int fill_end = min(st0 + ((en0 - st0) / 16 + 1) * 16, T16);   // scores are refreshed for [st0, fill_end)
			const int span_end = max(en + 1, fill_end);
			const uint8_t *qrr = qr + (qlen - 1 - r);
			// ---- phase A: loads (everything a cell reads is from before this diagonal) ----
			constexpr int KMAX = 4;  // (752 + 30) / 256 cells per thread, rounded up
			int8_t uo[KMAX], yo[KMAX], y2o[KMAX], xl[KMAX], vl[KMAX], x2l[KMAX], zz[KMAX];
			int32_t hold[KMAX], hprev = 0;
			bool fresh[KMAX];
			#pragma unroll
			for (int k = 0; k < KMAX; ++k) {
				const int t = st + tid + k * DPC_THREADS;
				fresh[k] = false; uo[k] = yo[k] = y2o[k] = xl[k] = vl[k] = x2l[k] = zz[k] = 0; hold[k] = 0;
				if (t >= span_end) continue;
				if (t >= st0 && t < fill_end) {
					const int jq = r - t;
					const uint8_t sq = sf[ix(t)], sq2 = WIN ? ((jq >= 0 && jq < qlen) ? qr[jq & (DPC_WIN - 1)] : (uint8_t)0) : qrr[t];
					zz[k] = (sq == 4 || sq2 == 4) ? sc.sc_N : (sq == sq2 ? sc.sc_mch : sc.sc_mis);
					fresh[k] = true;
				} else if (t <= en) zz[k] = s[ix(t)];
				if (t > en) continue;
				uo[k] = u[ix(t)], yo[k] = y[ix(t)], y2o[k] = y2[ix(t)];
				if (t > st) xl[k] = x[ix(t - 1)], vl[k] = v[ix(t - 1)], x2l[k] = x2[ix(t - 1)];
				else if (st > 0) {
					if (st - 1 >= last_st && st - 1 <= last_en) xl[k] = x[ix(st - 1)], x2l[k] = x2[ix(st - 1)], vl[k] = v[ix(st - 1)];
					else xl[k] = I1, x2l[k] = I2, vl[k] = I1;
				} else { xl[k] = I1, x2l[k] = I2; vl[k] = bnd(r); }
				if (en >= r && t == r) { yo[k] = I1, y2o[k] = I2; uo[k] = bnd(r); }
				if (with_exact && r > 0) {
					if (t >= st0 && t <= en0) hold[k] = H[ix(t)];
					if (t == en0 && en0 > 0) hprev = H[ix(en0 - 1)];
				}
			}
			__syncthreads();
			// ---- phase B: compute and store ----
			long long best = ((long long)MB_KSW_NEG_INF << 32);
			#pragma unroll
			for (int k = 0; k < KMAX; ++k) {
				const int t = st + tid + k * DPC_THREADS;
				if (t >= span_end) continue;
				if (fresh[k]) s[ix(t)] = zz[k];
				if (t > en) continue;
				int8_t z = zz[k];
				const int8_t xt1 = xl[k], vt1 = vl[k], x2t1 = x2l[k], ut = uo[k];
				int8_t a = (int8_t)(xt1 + vt1), b = (int8_t)(yo[k] + ut), a2 = (int8_t)(x2t1 + vt1), b2 = (int8_t)(y2o[k] + ut);
				int8_t d;
				if (!right) {
					d = a > z ? 1 : 0;   z = z > a ? z : a;
					d = b > z ? 2 : d;   z = z > b ? z : b;
					d = a2 > z ? 3 : d;  z = z > a2 ? z : a2;
					d = b2 > z ? 4 : d;  z = z > b2 ? z : b2;
				} else {
					d = z > a ? 0 : 1;   z = z > a ? z : a;
					d = z > b ? d : 2;   z = z > b ? z : b;
					d = z > a2 ? d : 3;  z = z > a2 ? z : a2;
					d = z > b2 ? d : 4;  z = z > b2 ? z : b2;
				}
				z = z < sc.sc_mch ? z : sc.sc_mch;
				const int8_t un = (int8_t)(z - vt1), vn = (int8_t)(z - ut);
				u[ix(t)] = un, v[ix(t)] = vn;
				int8_t tmp = (int8_t)(z - q);
				a = (int8_t)(a - tmp), b = (int8_t)(b - tmp);
				tmp = (int8_t)(z - q2);
				a2 = (int8_t)(a2 - tmp), b2 = (int8_t)(b2 - tmp);
				if (!right) {
					x[ix(t)]  = (int8_t)((a  > 0 ? a  : 0) - qe);  if (a  > 0) d |= 0x08;
					y[ix(t)]  = (int8_t)((b  > 0 ? b  : 0) - qe);  if (b  > 0) d |= 0x10;
					x2[ix(t)] = (int8_t)((a2 > 0 ? a2 : 0) - qe2); if (a2 > 0) d |= 0x20;
					y2[ix(t)] = (int8_t)((b2 > 0 ? b2 : 0) - qe2); if (b2 > 0) d |= 0x40;
				} else {
					x[ix(t)]  = (int8_t)((0 > a  ? 0 : a)  - qe);  if (!(0 > a))  d |= 0x08;
					y[ix(t)]  = (int8_t)((0 > b  ? 0 : b)  - qe);  if (!(0 > b))  d |= 0x10;
					x2[ix(t)] = (int8_t)((0 > a2 ? 0 : a2) - qe2); if (!(0 > a2)) d |= 0x20;
					y2[ix(t)] = (int8_t)((0 > b2 ? 0 : b2) - qe2); if (!(0 > b2)) d |= 0x40;
				}
				P[(size_t)r * ncol16 + (t - st)] = (uint8_t)d;
				if (with_exact) {
					if (r > 0) {
						if (t >= st0 && t < en0) {
							const int32_t h = hold[k] + (int32_t)vn;
							H[ix(t)] = h;
							const int en1 = st0 + (en0 - st0) / 4 * 4;
							const unsigned rank = t < en1 ? 1u + ((unsigned)((t - st0) & 3) << 22) + (unsigned)((t - st0) >> 2) : 1u + (4u << 22) + (unsigned)(t - en1);
							const long long key = ((long long)h << 32) | (unsigned)(0x7fffffffu - rank);
							best = key > best ? key : best;
						}
						if (t == en0) {
							const int32_t he = en0 > 0 ? hprev + (int32_t)un : hold[k] + (int32_t)vn;
							H[ix(en0)] = he;
							const long long key = ((long long)he << 32) | (unsigned)0x7fffffffu;
							best = key > best ? key : best;
						}
					} else if (t == 0) H[0] = (int32_t)vn - qe;   // ix(0) == 0
				}
			}
			if (with_exact && r > 0) {
				#pragma unroll
				for (int dlt = 16; dlt > 0; dlt >>= 1) { const long long o = __shfl_xor_sync(FULL, best, dlt); best = o > best ? o : best; }
				if (lane == 0) s_red[wid] = best;
			}
			if (WIN && r + 1 < n_rows) { // pristine entries for what the next diagonal reaches beyond the entries set up so far
				const int r1 = r + 1;
				int st1 = 0, en1 = tlen - 1;
				if (st1 < r1 - qlen + 1) st1 = r1 - qlen + 1;
				if (en1 > r1) en1 = r1;
				if (st1 < (r1 - w + 1) >> 1) st1 = (r1 - w + 1) >> 1;
				if (en1 > (r1 + w) >> 1) en1 = (r1 + w) >> 1;
				int need_t = (en1 + 16) / 16 * 16 + 31; if (need_t > T16 - 1) need_t = T16 - 1;
				for (int t = hi_t + 1 + tid; t <= need_t; t += DPC_THREADS) fresh_t(t);
				if (need_t > hi_t) hi_t = need_t;
				int need_j = r1 - st1 + 16; if (need_j > qlen - 1) need_j = qlen - 1;
				for (int j = hi_j + 1 + tid; j <= need_j; j += DPC_THREADS) qr[j & (DPC_WIN - 1)] = (uint8_t)qv.at(j);
				if (need_j > hi_j) hi_j = need_j;
			}
			__syncthreads();
			cells += tid == 0 ? (unsigned)(en0 - st0 + 1) : 0u;
			if (with_exact) {
				int32_t max_H, max_t;
				if (r > 0) {
					long long bb = s_red[0];
					#pragma unroll
					for (int k = 1; k < DPC_THREADS / 32; ++k) { const long long o = s_red[k]; bb = o > bb ? o : bb; }
					max_H = (int32_t)(bb >> 32);
					const int en1 = st0 + (en0 - st0) / 4 * 4;
					const unsigned rank = 0x7fffffffu - (unsigned)(bb & 0xffffffffu);
					if (rank == 0) max_t = en0;
					else if (rank < 1u + (4u << 22)) { const unsigned rr = rank - 1; max_t = st0 + (int)((rr & ((1u << 22) - 1)) << 2) + (int)(rr >> 22); }
					else max_t = en1 + (int)(rank - 1u - (4u << 22));
				} else { max_H = (int32_t)v[0] - qe, max_t = 0; }
				if (en0 == tlen - 1) { const int32_t h = H[ix(en0)]; if (h > ez_mte) ez_mte = h; }
				if (r - st0 == qlen - 1) { const int32_t h = H[ix(st0)]; if (h > ez_mqe) ez_mqe = h, ez_mqe_t = st0; }
				bool brk = false;
				if (max_H > ez_max) { ez_max = max_H, ez_max_t = max_t, ez_max_q = r - max_t; }
				else if (max_t >= ez_max_t && r - max_t >= ez_max_q) {
					const int tl = max_t - ez_max_t, ql = (r - max_t) - ez_max_q;
					const int l = tl > ql ? tl - ql : ql - tl;
					if (zdrop >= 0 && ez_max - max_H > zdrop + l * e2) { ez_zdropped = 1; brk = true; }
				}
				if (brk) break;
				if (r == qlen + tlen - 2 && en0 == tlen - 1) ez_score = H[ix(tlen - 1)];
			} else {
				if (r > 0) {
					if (last_H0_t >= st0 && last_H0_t <= en0 && last_H0_t + 1 >= st0 && last_H0_t + 1 <= en0) {
						const int32_t d0 = v[ix(last_H0_t)], d1 = u[ix(last_H0_t + 1)];
						if (d0 > d1) H0 += d0;
						else H0 += d1, ++last_H0_t;
					} else if (last_H0_t >= st0 && last_H0_t <= en0) {
						H0 += v[ix(last_H0_t)];
					} else {
						++last_H0_t, H0 += u[ix(last_H0_t)];
					}
				} else H0 = v[0] - qe, last_H0_t = 0;
				if (r == qlen + tlen - 2 && en0 == tlen - 1) ez_score = H0;
			}
			last_st = st, last_en = en;
		}
		__syncthreads();
		// ---- backtrack: warp 0; 32 cells of the current diagonal per fetch, ksw_backtrack's state machine replayed through shuffles ----
		if (wid == 0) {
			int reach_end = 0, n_cigar = 0;
			int i0 = -1, j0 = -1;
			const bool rev_cigar = flag & MB_EZ_REV_CIGAR;
			if (!ez_zdropped && !(flag & MB_EZ_EXTZ_ONLY)) i0 = tlen - 1, j0 = qlen - 1;
			else if (!ez_zdropped && (flag & MB_EZ_EXTZ_ONLY) && ez_mqe + end_bonus > ez_max) reach_end = 1, i0 = ez_mqe_t, j0 = qlen - 1;
			else if (ez_max_t >= 0 && ez_max_q >= 0) i0 = ez_max_t, j0 = ez_max_q;
			uint32_t *cigar = cigar_pool + T.cigar_off;
			uint32_t cur_op = 0; int cur_len = 0;
			auto push = [&](uint32_t op, int len) {
				if (cur_len > 0 && op != cur_op) { if (lane == 0) cigar[n_cigar] = (uint32_t)cur_len << 4 | cur_op; ++n_cigar; cur_len = 0; }
				cur_op = op, cur_len += len;
			};
			auto range = [&](int r, int &st, int &en) {
				st = 0, en = tlen - 1;
				if (st < r - qlen + 1) st = r - qlen + 1;
				if (en > r) en = r;
				if (st < (r - w + 1) >> 1) st = (r - w + 1) >> 1;
				if (en > (r + w) >> 1) en = (r + w) >> 1;
				st = st / 16 * 16, en = (en + 16) / 16 * 16 - 1;
			};
			if (i0 >= 0 && j0 >= 0) {
				int i = i0, j = j0, state = 0;
				while (i >= 0 && j >= 0) {
					const int ci = i - lane, cj = j - lane;
					uint32_t raw = 0xffu; // off the matrix or outside the computed range: stops a run
					if (ci >= 0 && cj >= 0) {
						int st, en; range(ci + cj, st, en);
						if (ci >= st && ci <= en) raw = P[(size_t)(ci + cj) * ncol16 + ci - st];
					}
					int k = 0;
					if (state == 0) { // leading cells of the diagonal whose direction is "match" (d == 0)
						const unsigned stop = __ballot_sync(FULL, (raw & 7u) != 0u || raw == 0xffu);
						k = stop ? __ffs(stop) - 1 : 32;
						if (k) { push(0, k); i -= k, j -= k; }
					}
					if (k < 32 && i >= 0 && j >= 0) {
						const uint32_t rk = __shfl_sync(FULL, raw, k);
						int st, en, force = -1;
						range(i + j, st, en);
						if (i < st) force = 2;
						if (i > en) force = 1;
						const uint32_t tmp = force < 0 ? rk : 0u;
						if (state == 0) state = tmp & 7;
						else if (!(tmp >> (state + 2) & 1)) state = 0;
						if (state == 0) state = tmp & 7;
						if (force >= 0) state = force;
						if (state == 0) { push(0, 1); --i, --j; }
						else if (state == 1 || state == 3) { push(2, 1); --i; }
						else { push(1, 1); --j; }
					}
				}
				if (i >= 0) push(2, i + 1);
				if (j >= 0) push(1, j + 1);
				if (cur_len > 0) { if (lane == 0) cigar[n_cigar] = (uint32_t)cur_len << 4 | cur_op; ++n_cigar; }
				__syncwarp();
				if (!rev_cigar)
					for (int k = lane; k < n_cigar >> 1; k += 32) { uint32_t t_ = cigar[k]; cigar[k] = cigar[n_cigar - 1 - k], cigar[n_cigar - 1 - k] = t_; }
			}
			if (lane == 0) {
				T.score = ez_score, T.max = ez_max, T.max_q = ez_max_q, T.max_t = ez_max_t, T.mqe = ez_mqe, T.mqe_t = ez_mqe_t;
				T.zdropped = ez_zdropped, T.reach_end = reach_end, T.n_cigar = n_cigar;
			}
		}
	}
	if (tid == 0 && cells_out && cells) atomicAdd(cells_out, cells);
}

// k_dp_cta2: the same kernel with the cells of a diagonal computed four per thread in packed 16-bit lanes (see phase B).
__global__ void __launch_bounds__(DPC2_THREADS, 5)
k_dp_cta2(DpTask *__restrict__ tasks, const int32_t *__restrict__ order, const int32_t *__restrict__ n_order, int32_t *__restrict__ work_ctr,
         const uint8_t *__restrict__ codes, const uint32_t *__restrict__ S, const uint8_t *__restrict__ pool,
         uint8_t *__restrict__ p_scr, size_t p_stride, int8_t *__restrict__ g_ws, size_t g_stride, int32_t *__restrict__ h_scr, size_t h_stride,
         uint32_t *__restrict__ cigar_pool, DpScoring sc, unsigned long long *__restrict__ cells_out, int smem_bytes)
{
	constexpr bool WIN = true;
	extern __shared__ __align__(16) int8_t dpc_smem[];
	__shared__ uint32_t s_red[DPC2_THREADS / 32];
	__shared__ int s_task, s_brk;   // s_brk: warp 0 found the Z-drop break on the previous diagonal
	const unsigned FULL = 0xffffffffu;
	const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
	uint8_t *P = p_scr + (size_t)blockIdx.x * p_stride;
	unsigned long long cells = 0;
	const int n_total = *n_order;
	for (;;) {
		__syncthreads();
		if (tid == 0) s_task = atomicAdd(work_ctr, 1);
		__syncthreads();
		const int oi = s_task;
		if (oi >= n_total) break;
		DpTask &T = tasks[order[oi]];
		const int qlen = T.qlen, tlen = T.tlen, flag = T.flag, zdrop = T.zdrop, end_bonus = T.end_bonus;
		if (tid == 0) { dp_reset(T); s_brk = 0; }
		if (T.skip) { if (tid == 0) T.zdropped = 1; continue; }
		if (qlen <= 0 || tlen <= 0) continue;
		int8_t q = sc.q, e = sc.e, q2 = sc.q2, e2 = sc.e2;
		if (q2 + e2 < q + e) { int8_t t_ = q; q = q2, q2 = t_; t_ = e; e = e2, e2 = t_; }
		const int qe = q + e, qe2 = q2 + e2;
		int w = T.w;
		if (w < 0) w = tlen > qlen ? tlen : qlen;
		const int tlen_ = (tlen + 15) / 16, qlen_ = (qlen + 15) / 16;
		int n_col_ = qlen < tlen ? qlen : tlen;
		n_col_ = ((n_col_ < w + 1 ? n_col_ : w + 1) + 15) / 16 + 1;
		const int ncol16 = n_col_ * 16;
		const int T16 = tlen_ * 16;
		const bool with_exact = !(flag & MB_EZ_APPROX_MAX);
		const bool right = (flag & MB_EZ_RIGHT) != 0;
		int long_thres = e != e2 ? (q2 - q) / (e - e2) - 1 : 0;
		if (q2 + e2 + long_thres * e2 > q + e + long_thres * e) ++long_thres;
		const int long_diff = long_thres * (e - e2) - (q2 - q) - e2;
		const size_t ws_need = (size_t)T16 * 8 + (size_t)qlen_ * 16 + 16;
		const size_t h_need = with_exact ? (size_t)T16 * 4 : 0;
		const bool in_smem = WIN || ((ws_need + 15) & ~(size_t)15) + h_need <= (size_t)smem_bytes;
		const int A = WIN ? DPC_WIN : T16;   // entries per state array
		int8_t *ws = in_smem ? dpc_smem : g_ws + (size_t)blockIdx.x * g_stride;
		int32_t *H = WIN ? reinterpret_cast<int32_t*>(dpc_smem + DPC_WIN * 9)
		           : (in_smem && with_exact) ? reinterpret_cast<int32_t*>(ws + ((ws_need + 15) & ~(size_t)15)) : h_scr + (size_t)blockIdx.x * h_stride;
		int8_t *u = ws, *v = u + A, *x = v + A, *y = x + A, *x2 = y + A, *y2 = x2 + A, *s = y2 + A;
		uint8_t *sf = (uint8_t*)(s + A), *qr = sf + A;   // WIN: qr holds the query by position j (not reversed), circular
		const int8_t I1 = (int8_t)(-q - e), I2 = (int8_t)(-q2 - e2);
		QView qv; qv.codes = T.q_comp == 2 ? pool : codes; qv.idx0 = T.q_idx0; qv.step = T.q_step; qv.comp = T.q_comp == 1;
		TView tv; tv.S = S; tv.bytes = pool; tv.idx0 = T.t_idx0; tv.step = T.t_step; tv.packed = T.t_packed;
		auto ix = [](int t) { return WIN ? (t & (DPC_WIN - 1)) : t; };
		auto fresh_t = [&](int t) { // a pristine entry, as upstream's arrays are before a diagonal touches them
			const int k = ix(t);
			u[k] = I1, v[k] = I1, x[k] = I1, y[k] = I1, x2[k] = I2, y2[k] = I2, s[k] = 0;
			sf[k] = t < tlen ? (uint8_t)tv.at(t) : 0;
			if (with_exact) H[k] = MB_KSW_NEG_INF;
		};
		int hi_t = (WIN && T16 > DPC_WIN ? DPC_WIN : T16) - 1;            // entries 0..hi_t are set up
		int hi_j = WIN ? (qlen > DPC_WIN ? DPC_WIN : qlen) - 1 : 0;
		for (int t = tid; t <= hi_t; t += DPC2_THREADS) fresh_t(t);
		if (WIN) { for (int j = tid; j <= hi_j; j += DPC2_THREADS) qr[(-j) & (DPC_WIN - 1)] = (uint8_t)qv.at(j); }
		else for (int t = tid; t < qlen_ * 16 + 16; t += DPC2_THREADS) qr[t] = t < qlen ? (uint8_t)qv.at(qlen - 1 - t) : 0;
		__syncthreads();
		// ez state: every thread keeps the same copy
		int ez_max = 0, ez_max_t = -1, ez_max_q = -1, ez_mqe = MB_KSW_NEG_INF, ez_mqe_t = -1, ez_mte = MB_KSW_NEG_INF, ez_score = MB_KSW_NEG_INF, ez_zdropped = 0;
		int32_t H0 = 0, last_H0_t = 0;
		int last_st = -1, last_en = -1;
		const int n_rows = qlen + tlen - 1;
		// packed constants: int8 value in the high byte of each 16-bit lane, candidate priority tag in the low byte
		// (diag, x, y, x2, y2) = (4, 3, 2, 1, 0): ties go to the earlier candidate; reversed with KSW_EZ_RIGHT
		const uint32_t TZ = right ? 0u : 0x00040004u, T1 = right ? 0x00010001u : 0x00030003u, T2 = 0x00020002u, T3 = right ? 0x00030003u : 0x00010001u, T4 = right ? 0x00040004u : 0u;
		const uint32_t MCH8 = (uint32_t)(uint8_t)sc.sc_mch * 0x01000100u, Q8 = (uint32_t)(uint8_t)q * 0x01000100u, Q28 = (uint32_t)(uint8_t)q2 * 0x01000100u;
		const uint32_t QE8 = (uint32_t)(uint8_t)qe * 0x01000100u, QE28 = (uint32_t)(uint8_t)qe2 * 0x01000100u, C256 = 0x01000100u;
		const uint32_t LUT = (uint32_t)(uint8_t)sc.sc_mch | (uint32_t)(uint8_t)sc.sc_mis << 8 | (uint32_t)(uint8_t)sc.sc_N << 16 | (uint32_t)(uint8_t)sc.sc_N << 24;
		auto bnd = [&](int r) { return (int8_t)(r == 0 ? -q - e : r < long_thres ? -e : r == long_thres ? long_diff : -e2); };
		for (int r = 0; r < n_rows; ++r) {
			int st = 0, en = tlen - 1;
			if (st < r - qlen + 1) st = r - qlen + 1;
			if (en > r) en = r;
			if (st < (r - w + 1) >> 1) st = (r - w + 1) >> 1;
			if (en > (r + w) >> 1) en = (r + w) >> 1;
			if (st > en) { ez_zdropped = 1; break; }
			const int st0 = st, en0 = en;
			st = st / 16 * 16, en = (en + 16) / 16 * 16 - 1;
			const int fill_end = min(st0 + ((en0 - st0) / 16 + 1) * 16, T16);   // scores are refreshed for [st0, fill_end)
			const int span_end = max(en + 1, fill_end);
			const uint8_t *qrr = qr + (qlen - 1 - r);
			// ---- phase A: loads (everything a cell reads is from before this diagonal); four consecutive cells per thread ----
			const int t0 = st + 4 * tid;
			const bool dp_on = t0 <= en;                       // block ranges are 16-aligned: a group of four is inside or outside as a whole
			const bool any_on = t0 < span_end;
			uint32_t Wu = 0, Wy = 0, Wy2 = 0, Ws = 0, Wx = 0, Wv = 0, Wx2 = 0, Wsf = 0, Wq = 0;
			uint32_t ex = 0, ev = 0, ex2 = 0;
			int4 hold = make_int4(0, 0, 0, 0); int32_t hprev = 0;
			const int k0 = t0 & (DPC_WIN - 1);
			if (any_on) {
				Ws = *reinterpret_cast<const uint32_t*>(s + k0);
				Wsf = *reinterpret_cast<const uint32_t*>(sf + k0);
				const int aq = (t0 - r) & (DPC_WIN - 1);
				const uint32_t q0 = *reinterpret_cast<const uint32_t*>(qr + (aq & ~3)), q1 = *reinterpret_cast<const uint32_t*>(qr + ((aq + 4) & (DPC_WIN - 1) & ~3));
				Wq = __funnelshift_r(q0, q1, (aq & 3) * 8);
				if (t0 + 3 > r) { // columns right of the first row (j < 0): upstream reads its zero padding there
					const int nz = r - t0 + 1;                 // cells 0..nz-1 are real
					Wq = nz <= 0 ? 0u : (Wq & (0xffffffffu >> (8 * (4 - nz))));
				}
			}
			if (dp_on) {
				Wu = *reinterpret_cast<const uint32_t*>(u + k0), Wy = *reinterpret_cast<const uint32_t*>(y + k0), Wy2 = *reinterpret_cast<const uint32_t*>(y2 + k0);
				Wx = *reinterpret_cast<const uint32_t*>(x + k0), Wv = *reinterpret_cast<const uint32_t*>(v + k0), Wx2 = *reinterpret_cast<const uint32_t*>(x2 + k0);
				if (t0 > st) ex = (uint8_t)x[ix(t0 - 1)], ev = (uint8_t)v[ix(t0 - 1)], ex2 = (uint8_t)x2[ix(t0 - 1)];
				else if (st > 0) {
					if (st - 1 >= last_st && st - 1 <= last_en) ex = (uint8_t)x[ix(st - 1)], ex2 = (uint8_t)x2[ix(st - 1)], ev = (uint8_t)v[ix(st - 1)];
					else ex = (uint8_t)I1, ex2 = (uint8_t)I2, ev = (uint8_t)I1;
				} else { ex = (uint8_t)I1, ex2 = (uint8_t)I2; ev = (uint8_t)bnd(r); }
				if (en >= r && r >= t0 && r < t0 + 4) { // the cell of the first row (t == r): its y, y2, u are the boundary values
					const int sh = 8 * (r - t0);
					const uint32_t m = ~(0xffu << sh);
					Wy = (Wy & m) | ((uint32_t)(uint8_t)I1 << sh), Wy2 = (Wy2 & m) | ((uint32_t)(uint8_t)I2 << sh), Wu = (Wu & m) | ((uint32_t)(uint8_t)bnd(r) << sh);
				}
				if (with_exact && r > 0) {
					hold = *reinterpret_cast<const int4*>(H + k0);
					if (en0 >= t0 && en0 < t0 + 4 && en0 > 0) hprev = H[ix(en0 - 1)];
				}
			}
			__syncthreads();
			if (s_brk) break;   // (set by warp 0 in the bookkeeping of the previous diagonal; nothing of this diagonal has been stored yet)
			// ---- phase B: compute (two cells per 32-bit register: int8 value in the high byte of a 16-bit lane, so that 16-bit
			//      wrap-around IS upstream's int8 wrap-around; candidate priority tags in the low byte) and store ----
			// per-diagonal maximum of H with upstream's order among equal scores: key = (H + 2^20) << 11 | (2047 - rank); |H| stays far
			// below 2^20 for the task sizes this kernel takes (lengths <= max_gap windows, a few ten thousand at most)
			uint32_t best = 0u;
			if (any_on) {
				// scores of the fresh cells [st0, fill_end); the others keep the stored one
				uint32_t Wz = Ws;
				{
					const int lo = st0 - t0 < 0 ? 0 : (st0 - t0 > 4 ? 4 : st0 - t0), hi = fill_end - t0 < 0 ? 0 : (fill_end - t0 > 4 ? 4 : fill_end - t0);
					if (hi > lo) {
						const uint32_t X = (Wsf ^ Wq) & 0x07070707u;   // (cells outside the fresh range may hold stale bytes: no carries between bytes)
						const uint32_t neq = ((X + 0x7f7f7f7fu) & 0x80808080u) >> 7, amb = ((Wsf | Wq) & 0x04040404u) >> 1;   // 1: differ, 2: ambiguous
						const uint32_t I = neq | amb;
						const uint32_t sel = (I & 0xfu) | (I >> 4 & 0xf0u) | (I >> 8 & 0xf00u) | (I >> 12 & 0xf000u);
						const uint32_t Sn = dpf_prmt(LUT, 0u, sel);
						const uint32_t fm = (hi >= 4 ? 0xffffffffu : ~(0xffffffffu << (8 * hi))) & (lo >= 4 ? 0u : (0xffffffffu << (8 * lo)));
						Wz = (Sn & fm) | (Ws & ~fm);
						*reinterpret_cast<uint32_t*>(s + k0) = Wz;
					}
				}
				if (dp_on) {
					const uint32_t Mh = 0xff00ff00u;
					uint32_t dW[2], nu[2], nv[2], nx[2], ny[2], nx2[2], ny2[2];
					#pragma unroll
					for (int p2 = 0; p2 < 2; ++p2) {
						// lanes: cells 2*p2 (low) and 2*p2+1 (high)
						const uint32_t selv = p2 ? 0x3424u : 0x1404u;
						const uint32_t Z0 = dpf_prmt(Wz, 0u, selv) | TZ;
						const uint32_t U_ = dpf_prmt(Wu, 0u, selv), Y_ = dpf_prmt(Wy, 0u, selv) | T2, Y2_ = dpf_prmt(Wy2, 0u, selv) | T4;
						// left neighbours: cells t-1 (edge byte for the first cell of the group)
						const uint32_t sell = p2 ? 0x2414u : 0x0444u;                                  // p2: bytes (x1, x2); else (edge, x0)
						const uint32_t XL = (p2 ? dpf_prmt(Wx, 0u, sell) : (dpf_prmt(Wx, 0u, 0x0444u) | ex << 8)) | T1;
						const uint32_t VL = (p2 ? dpf_prmt(Wv, 0u, sell) : (dpf_prmt(Wv, 0u, 0x0444u) | ev << 8));
						const uint32_t X2L = (p2 ? dpf_prmt(Wx2, 0u, sell) : (dpf_prmt(Wx2, 0u, 0x0444u) | ex2 << 8)) | T3;
						const uint32_t a = __vadd2(XL, VL), b = __vadd2(Y_, U_), a2 = __vadd2(X2L, VL), b2 = __vadd2(Y2_, U_);
						const uint32_t zt = __vimax3_s16x2(__vimax3_s16x2(Z0, a, b), a2, b2);
						const uint32_t tag = zt & 0x00ff00ffu;
						const uint32_t zc = __vmins2(zt & Mh, MCH8);
						nu[p2] = __vsub2(zc, VL), nv[p2] = __vsub2(zc, U_);
						const uint32_t nz1 = __vsub2(Q8, zc), nz2 = __vsub2(Q28, zc);
						const uint32_t s1 = __vadd2(a, nz1), s2 = __vadd2(b, nz1), s3 = __vadd2(a2, nz2), s4 = __vadd2(b2, nz2);
						const uint32_t r1 = __vmaxs2(s1, T1), r2 = __vmaxs2(s2, T2), r3 = __vmaxs2(s3, T3), r4 = __vmaxs2(s4, T4);   // max(.., 0) keeping the tag
						nx[p2] = __vsub2(r1, QE8), ny[p2] = __vsub2(r2, QE8), nx2[p2] = __vsub2(r3, QE28), ny2[p2] = __vsub2(r4, QE28);
						uint32_t g1, g2, g3, g4;
						if (!right) { // continue the gap iff the value is > 0
							g1 = __vmins2(__vsub2(r1, T1), C256), g2 = __vmins2(__vsub2(r2, T2), C256), g3 = __vmins2(__vsub2(r3, T3), C256), g4 = __vmins2(__vsub2(r4, T4), C256);
						} else {      // ... iff it is >= 0
							g1 = (~s1 >> 7) & C256, g2 = (~s2 >> 7) & C256, g3 = (~s3 >> 7) & C256, g4 = (~s4 >> 7) & C256;
						}
						const uint32_t dir = right ? tag : 0x00040004u - tag;
						dW[p2] = dir | g1 >> 5 | g2 >> 4 | g3 >> 3 | g4 >> 2;
					}
					*reinterpret_cast<uint32_t*>(u + k0) = dpf_prmt(nu[0], nu[1], 0x7531u);
					*reinterpret_cast<uint32_t*>(v + k0) = dpf_prmt(nv[0], nv[1], 0x7531u);
					*reinterpret_cast<uint32_t*>(x + k0) = dpf_prmt(nx[0], nx[1], 0x7531u);
					*reinterpret_cast<uint32_t*>(y + k0) = dpf_prmt(ny[0], ny[1], 0x7531u);
					*reinterpret_cast<uint32_t*>(x2 + k0) = dpf_prmt(nx2[0], nx2[1], 0x7531u);
					*reinterpret_cast<uint32_t*>(y2 + k0) = dpf_prmt(ny2[0], ny2[1], 0x7531u);
					*reinterpret_cast<uint32_t*>(P + (size_t)r * ncol16 + (t0 - st)) = dpf_prmt(dW[0], dW[1], 0x6420u);
					if (with_exact) {
						if (r > 0) {
							int32_t hv[4] = { hold.x, hold.y, hold.z, hold.w };
							const int en1 = st0 + (en0 - st0) / 4 * 4;
							#pragma unroll
							for (int c = 0; c < 4; ++c) {
								const int t = t0 + c;
								const int32_t vn = (int32_t)(int16_t)((c & 1) ? nv[c >> 1] >> 16 : nv[c >> 1] & 0xffffu) >> 8;
								if (t >= st0 && t < en0) {
									const int32_t h = hv[c] + vn;
									hv[c] = h;
									const unsigned rank = t < en1 ? 1u + ((unsigned)((t - st0) & 3) << 8) + (unsigned)((t - st0) >> 2) : 1025u + (unsigned)(t - en1);
									const int32_t hc = h < -0xfffff ? -0xfffff : (h > 0xfffff ? 0xfffff : h);
									const uint32_t key = (uint32_t)(hc + 0x100000) << 11 | (2047u - rank);
									best = key > best ? key : best;
								} else if (t == en0) {
									const int32_t un = (int32_t)(int16_t)((c & 1) ? nu[c >> 1] >> 16 : nu[c >> 1] & 0xffffu) >> 8;
									const int32_t he = en0 > 0 ? hprev + un : hv[c] + vn;
									hv[c] = he;
									const int32_t hc = he < -0xfffff ? -0xfffff : (he > 0xfffff ? 0xfffff : he);
									const uint32_t key = (uint32_t)(hc + 0x100000) << 11 | 2047u;
									best = key > best ? key : best;
								}
							}
							*reinterpret_cast<int4*>(H + k0) = make_int4(hv[0], hv[1], hv[2], hv[3]);
						} else if (t0 == 0) H[0] = ((int32_t)(int16_t)(nv[0] & 0xffffu) >> 8) - qe;
					}
				}
			}
			if (with_exact && r > 0) {
				best = __reduce_max_sync(FULL, best);
				if (lane == 0) s_red[wid] = best;
			}
			if ((r & 15) == 15 && r + 1 < n_rows) { // every 16th diagonal: pristine entries for what the next 16 diagonals reach
				const int r1 = r + 16 < n_rows ? r + 16 : n_rows - 1;
				int st1 = 0, en1 = tlen - 1;
				if (st1 < r1 - qlen + 1) st1 = r1 - qlen + 1;
				if (en1 > r1) en1 = r1;
				if (st1 < (r1 - w + 1) >> 1) st1 = (r1 - w + 1) >> 1;
				if (en1 > (r1 + w) >> 1) en1 = (r1 + w) >> 1;
				int need_t = (en1 + 16) / 16 * 16 + 31; if (need_t > T16 - 1) need_t = T16 - 1;
				for (int t = hi_t + 1 + tid; t <= need_t; t += DPC2_THREADS) fresh_t(t);
				if (need_t > hi_t) hi_t = need_t;
				int need_j = r1 - st1 + 16; if (need_j > qlen - 1) need_j = qlen - 1;
				for (int j = hi_j + 1 + tid; j <= need_j; j += DPC2_THREADS) qr[(-j) & (DPC_WIN - 1)] = (uint8_t)qv.at(j);
				if (need_j > hi_j) hi_j = need_j;
			}
			__syncthreads();
			cells += tid == 0 ? (unsigned)(en0 - st0 + 1) : 0u;
			// the per-diagonal bookkeeping (ez state, Z-drop) is warp 0's alone: it also does the traceback; the other warps go on to
			// the next diagonal's loads and learn of a break at its first barrier
			if (wid != 0) { last_st = st, last_en = en; continue; }
			if (with_exact) {
				int32_t max_H, max_t;
				if (r > 0) {
					uint32_t bb = s_red[0];
					#pragma unroll
					for (int k = 1; k < DPC2_THREADS / 32; ++k) { const uint32_t o = s_red[k]; bb = o > bb ? o : bb; }
					max_H = (int32_t)(bb >> 11) - 0x100000;
					const int en1 = st0 + (en0 - st0) / 4 * 4;
					const unsigned rank = 2047u - (bb & 2047u);
					if (rank == 0) max_t = en0;
					else if (rank < 1025u) { const unsigned rr = rank - 1; max_t = st0 + (int)((rr & 255u) << 2) + (int)(rr >> 8); }
					else max_t = en1 + (int)(rank - 1025u);
				} else { max_H = (int32_t)v[0] - qe, max_t = 0; }
				if (en0 == tlen - 1) { const int32_t h = H[ix(en0)]; if (h > ez_mte) ez_mte = h; }
				if (r - st0 == qlen - 1) { const int32_t h = H[ix(st0)]; if (h > ez_mqe) ez_mqe = h, ez_mqe_t = st0; }
				bool brk = false;
				if (max_H > ez_max) { ez_max = max_H, ez_max_t = max_t, ez_max_q = r - max_t; }
				else if (max_t >= ez_max_t && r - max_t >= ez_max_q) {
					const int tl = max_t - ez_max_t, ql = (r - max_t) - ez_max_q;
					const int l = tl > ql ? tl - ql : ql - tl;
					if (zdrop >= 0 && ez_max - max_H > zdrop + l * e2) { ez_zdropped = 1; brk = true; }
				}
				if (brk) { if (lane == 0) s_brk = 1; }
				else if (r == qlen + tlen - 2 && en0 == tlen - 1) ez_score = H[ix(tlen - 1)];
			} else {
				if (r > 0) {
					if (last_H0_t >= st0 && last_H0_t <= en0 && last_H0_t + 1 >= st0 && last_H0_t + 1 <= en0) {
						const int32_t d0 = v[ix(last_H0_t)], d1 = u[ix(last_H0_t + 1)];
						if (d0 > d1) H0 += d0;
						else H0 += d1, ++last_H0_t;
					} else if (last_H0_t >= st0 && last_H0_t <= en0) {
						H0 += v[ix(last_H0_t)];
					} else {
						++last_H0_t, H0 += u[ix(last_H0_t)];
					}
				} else H0 = v[0] - qe, last_H0_t = 0;
				if (r == qlen + tlen - 2 && en0 == tlen - 1) ez_score = H0;
			}
			last_st = st, last_en = en;
		}
		__syncthreads();
		// ---- backtrack: warp 0; 32 cells of the current diagonal per fetch, ksw_backtrack's state machine replayed through shuffles ----
		if (wid == 0) {
			int reach_end = 0, n_cigar = 0;
			int i0 = -1, j0 = -1;
			const bool rev_cigar = flag & MB_EZ_REV_CIGAR;
			if (!ez_zdropped && !(flag & MB_EZ_EXTZ_ONLY)) i0 = tlen - 1, j0 = qlen - 1;
			else if (!ez_zdropped && (flag & MB_EZ_EXTZ_ONLY) && ez_mqe + end_bonus > ez_max) reach_end = 1, i0 = ez_mqe_t, j0 = qlen - 1;
			else if (ez_max_t >= 0 && ez_max_q >= 0) i0 = ez_max_t, j0 = ez_max_q;
			uint32_t *cigar = cigar_pool + T.cigar_off;
			uint32_t cur_op = 0; int cur_len = 0;
			auto push = [&](uint32_t op, int len) {
				if (cur_len > 0 && op != cur_op) { if (lane == 0) cigar[n_cigar] = (uint32_t)cur_len << 4 | cur_op; ++n_cigar; cur_len = 0; }
				cur_op = op, cur_len += len;
			};
			auto range = [&](int r, int &st, int &en) {
				st = 0, en = tlen - 1;
				if (st < r - qlen + 1) st = r - qlen + 1;
				if (en > r) en = r;
				if (st < (r - w + 1) >> 1) st = (r - w + 1) >> 1;
				if (en > (r + w) >> 1) en = (r + w) >> 1;
				st = st / 16 * 16, en = (en + 16) / 16 * 16 - 1;
			};
			if (i0 >= 0 && j0 >= 0) {
				int i = i0, j = j0, state = 0;
				while (i >= 0 && j >= 0) {
					const int ci = i - lane, cj = j - lane;
					uint32_t raw = 0xffu; // off the matrix or outside the computed range: stops a run
					if (ci >= 0 && cj >= 0) {
						int st, en; range(ci + cj, st, en);
						if (ci >= st && ci <= en) raw = P[(size_t)(ci + cj) * ncol16 + ci - st];
					}
					int k = 0;
					if (state == 0) { // leading cells of the diagonal whose direction is "match" (d == 0)
						const unsigned stop = __ballot_sync(FULL, (raw & 7u) != 0u || raw == 0xffu);
						k = stop ? __ffs(stop) - 1 : 32;
						if (k) { push(0, k); i -= k, j -= k; }
					}
					if (k < 32 && i >= 0 && j >= 0) {
						const uint32_t rk = __shfl_sync(FULL, raw, k);
						int st, en, force = -1;
						range(i + j, st, en);
						if (i < st) force = 2;
						if (i > en) force = 1;
						const uint32_t tmp = force < 0 ? rk : 0u;
						if (state == 0) state = tmp & 7;
						else if (!(tmp >> (state + 2) & 1)) state = 0;
						if (state == 0) state = tmp & 7;
						if (force >= 0) state = force;
						if (state == 0) { push(0, 1); --i, --j; }
						else if (state == 1 || state == 3) { push(2, 1); --i; }
						else { push(1, 1); --j; }
					}
				}
				if (i >= 0) push(2, i + 1);
				if (j >= 0) push(1, j + 1);
				if (cur_len > 0) { if (lane == 0) cigar[n_cigar] = (uint32_t)cur_len << 4 | cur_op; ++n_cigar; }
				__syncwarp();
				if (!rev_cigar)
					for (int k = lane; k < n_cigar >> 1; k += 32) { uint32_t t_ = cigar[k]; cigar[k] = cigar[n_cigar - 1 - k], cigar[n_cigar - 1 - k] = t_; }
			}
			if (lane == 0) {
				T.score = ez_score, T.max = ez_max, T.max_q = ez_max_q, T.max_t = ez_max_t, T.mqe = ez_mqe, T.mqe_t = ez_mqe_t;
				T.zdropped = ez_zdropped, T.reach_end = reach_end, T.n_cigar = n_cigar;
			}
		}
	}
	if (tid == 0 && cells_out && cells) atomicAdd(cells_out, cells);
}
