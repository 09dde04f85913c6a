// align2.cuh -- the bookkeeping kernels around k_dp: mm_test_zdrop, CIGAR stitching / Z-drop splitting,
// mm_update_extra (mlen, blen, NM, dp_max), and the per-read finalisation (mm_filter_regs, mm_hit_sort, mm_set_parent,
// mm_select_sub, mm_set_sam_pri, mm_set_mapq) plus hit emission in mappy's field layout (python/cmappy.h mm_reg2hitpy).
// Reference call sites: /root/reference/monica/genomes/aligner.py:193-195,215-217.
#pragma once
#include "align.cuh"

#include "ll.cuh"

// a gap fill whose CIGAR walk dropped by more than zdrop_inv: the inversion test (a local alignment of the dropped stretch
// against its reverse complement) is run by k_ztest_ll, one warp per candidate
struct ZCand { int32_t task, q_end, q_len, t_st, t_len, max_zdrop; };

MB_D void mb_update_max_zdrop(int32_t score, int i, int j, int32_t *mx, int *max_i, int *max_j, int e, int *max_zdrop, int pos[2][2])
{
	if (score < *mx) {
		int li = i - *max_i, lj = j - *max_j;
		int diff = li > lj ? li - lj : lj - li;
		int z = *mx - score - diff * e;
		if (z > *max_zdrop) {
			*max_zdrop = z;
			pos[0][0] = *max_i, pos[0][1] = *max_j;
			pos[1][0] = i, pos[1][1] = j;
		}
	} else *mx = score, *max_i = i, *max_j = j;
}

MB_D int mb_mat(int ct, int cq, const mb_opt_t &o)
{
	if (ct > 3 || cq > 3) return -(o.sc_ambi > 0 ? o.sc_ambi : -o.sc_ambi);
	return ct == cq ? (o.a < 0 ? -o.a : o.a) : (o.b > 0 ? -o.b : o.b);
}

// one thread per task of this round; gap fills only
// Cheap screen before the walk: the largest drop mm_test_zdrop can see is bounded by the total penalty mass of the path,
// and that follows from the CIGAR and the DP score alone -- with Lm aligned columns, G2 the two-piece gap costs and S the
// score, a*Lm - S - G2 = (a+b) * mismatches (an ambiguous base only makes the estimate larger as long as b >= sc_ambi), so
// penalty <= b * ceil((a*Lm - S - G2) / (a+b)) + sum(q + e*len).  At or below min(zdrop, zdrop_inv) the task cannot Z-drop
// and is done without touching the sequences; the others are listed (in task order) for the walk.
__global__ void k_ztest_screen(AlignCtx c, DpTask *__restrict__ tasks, int64_t n_tasks, const uint32_t *__restrict__ cigar_pool,
                               int32_t *__restrict__ walk_list, int32_t *__restrict__ n_walk)
{
	const int64_t ti = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	const int lane = threadIdx.x & 31;
	bool walk = false;
	if (ti < n_tasks && tasks[ti].kind == 1) {
		DpTask &T = tasks[ti];
		const mb_opt_t &opt = c.opt;
		const int thr = opt.zdrop < opt.zdrop_inv ? opt.zdrop : opt.zdrop_inv;
		walk = true;
		if (opt.a > 0 && opt.b > 0 && opt.b >= (opt.sc_ambi > 0 ? opt.sc_ambi : -opt.sc_ambi) && thr > 0) {
			const uint32_t *cigar = cigar_pool + T.cigar_off;
			int64_t lm = 0, g2 = 0, g1 = 0;
			bool odd = false;
			for (int k = 0; k < T.n_cigar; ++k) {
				const uint32_t op = cigar[k] & 0xf; const int64_t len = cigar[k] >> 4;
				if (op == 0) lm += len;
				else if (op == 1 || op == 2 || op == 3) {
					const int64_t c1 = opt.q + opt.e * len, c2 = opt.q2 + opt.e2 * len;
					g2 += c1 < c2 ? c1 : c2;
					g1 += c1;
				} else odd = true;
			}
			const int64_t num = (int64_t)opt.a * lm - (int64_t)T.score - g2;
			if (!odd && num >= 0) {
				const int64_t x_ub = (num + opt.a + opt.b - 1) / (opt.a + opt.b);
				if (opt.b * x_ub + g1 <= thr) { walk = false; T.zdrop_code = 0; }
			}
		}
	}
	const unsigned m = __ballot_sync(0xffffffffu, walk);
	if (m) {
		const int leader = __ffs(m) - 1;
		int base = 0;
		if (lane == leader) base = atomicAdd(n_walk, __popc(m));
		base = __shfl_sync(0xffffffffu, base, leader);
		if (walk) walk_list[base + __popc(m & ((1u << lane) - 1))] = (int32_t)ti;
	}
}

// One thread per task, walked as a flat event loop -- every iteration consumes exactly one event (a base of an M run or a
// whole gap run) -- over the tasks k_ztest_screen listed (`order`, `n_order`), so the lanes do not diverge over nested
// per-run loops.  Tasks that need the inversion test are handed to k_ztest_ll through `cand`; the others are finished here.
MB_D void mb_ztest_finish(DpTask &T, int code, const mb_opt_t &opt, int32_t *pass2_list, int32_t *n_pass2, int32_t ti)
{
	T.zdrop_code = code;
	if (code) {
		T.flag = 0; // second pass: exact max, real Z-drop
		T.zdrop = code == 2 ? opt.zdrop_inv : opt.zdrop;
		pass2_list[atomicAdd(n_pass2, 1)] = ti;
	}
}

__global__ void k_ztest(AlignCtx c, DpTask *__restrict__ tasks, const int32_t *__restrict__ order, const int32_t *__restrict__ n_order, int64_t n_tasks, const uint32_t *__restrict__ cigar_pool,
                        int32_t *__restrict__ pass2_list, int32_t *__restrict__ n_pass2, ZCand *__restrict__ cand, int32_t *__restrict__ n_cand)
{
	int64_t ti = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (ti >= n_tasks || (n_order && ti >= *n_order)) return;
	if (order) ti = order[ti];
	DpTask &T = tasks[ti];
	if (T.kind != 1) return;
	const mb_opt_t &opt = c.opt;
	QView qv; qv.codes = c.codes; qv.idx0 = T.q_idx0; qv.step = T.q_step; qv.comp = T.q_comp == 1;
	TView tv; tv.S = c.ix.S; tv.bytes = nullptr; tv.idx0 = T.t_idx0; tv.step = T.t_step; tv.packed = 1;
	const uint32_t *cigar = cigar_pool + T.cigar_off;
	const int n_cigar = T.n_cigar;
	int32_t score = 0, mx = INT32_MIN, max_i = -1, max_j = -1, i = 0, j = 0, max_zdrop = 0;
	int pos[2][2] = {{-1, -1}, {-1, -1}};
	const int gap_o = opt.q, gap_e = opt.e;
	int k = 0; uint32_t op = 0, len = 0, rem = 0;
	for (;;) {
		if (rem == 0) {
			if (k >= n_cigar) break;
			const uint32_t cg = cigar[k++];
			op = cg & 0xf, len = cg >> 4;
			if (op == 0) { rem = len; if (rem == 0) continue; }
			else if (op > 3) continue;
		}
		int ev_i, ev_j;
		if (op == 0) {
			score += mb_mat(tv.at(i), qv.at(j), opt);
			ev_i = i, ev_j = j;
			++i, ++j, --rem;
		} else {
			score -= gap_o + gap_e * (int)len;
			if (op == 1) j += len; else i += len;
			ev_i = i, ev_j = j;
		}
		mb_update_max_zdrop(score, ev_i, ev_j, &mx, &max_i, &max_j, gap_e, &max_zdrop, pos);
	}
	const int q_len = pos[1][1] - pos[0][1], t_len = pos[1][0] - pos[0][0];
	if (max_zdrop > opt.zdrop_inv && q_len < opt.max_gap && t_len < opt.max_gap) {
		if (q_len > 0 && t_len > 0) {
			ZCand z; z.task = (int32_t)ti, z.q_end = pos[1][1], z.q_len = q_len, z.t_st = pos[0][0], z.t_len = t_len, z.max_zdrop = max_zdrop;
			cand[atomicAdd(n_cand, 1)] = z;
			return;
		}
		// an empty side: upstream's local alignment scores 0 there (0 >= the thresholds only for degenerate options)
		if (0 >= opt.min_chain_score * opt.a && 0 >= opt.min_dp_max) { mb_ztest_finish(T, 2, opt, pass2_list, n_pass2, (int32_t)ti); return; }
	}
	mb_ztest_finish(T, max_zdrop > opt.zdrop ? 1 : 0, opt, pass2_list, n_pass2, (int32_t)ti);
}

// mm_test_zdrop's inversion test for the candidates of k_ztest: local alignment (ll.cuh) of the target stretch against the
// reverse complement of the query stretch.  Persistent one-warp CTAs over the candidate list; `scr` = 4*LL_MAX_LEN ints per CTA.
__global__ void __launch_bounds__(32)
k_ztest_ll(AlignCtx c, DpTask *__restrict__ tasks, const ZCand *__restrict__ cand, const int32_t *__restrict__ n_cand, int32_t *__restrict__ cursor,
           int32_t *__restrict__ pass2_list, int32_t *__restrict__ n_pass2, int *__restrict__ scr_pool)
{
	const int lane = threadIdx.x;
	int *scr = scr_pool + (size_t)blockIdx.x * (4 * LL_MAX_LEN);
	const mb_opt_t &opt = c.opt;
	const int n = *n_cand;
	for (;;) {
		int k = 0;
		if (lane == 0) k = atomicAdd(cursor, 1);
		k = __shfl_sync(0xffffffffu, k, 0);
		if (k >= n) break;
		const ZCand z = cand[k];
		DpTask &T = tasks[z.task];
		QView qv; qv.codes = c.codes; qv.idx0 = T.q_idx0; qv.step = T.q_step; qv.comp = T.q_comp == 1;
		TView tv; tv.S = c.ix.S; tv.bytes = nullptr; tv.idx0 = T.t_idx0; tv.step = T.t_step; tv.packed = 1;
		int qe, te;
		// qseq2[i] = complement of qseq[q_end - 1 - i]
		const int sc = mb_ll_warp([&](int col) { const int b = qv.at(z.q_end - 1 - col); return b >= 4 ? 4 : 3 - b; },
		                          [&](int row) { return tv.at(z.t_st + row); }, z.q_len, z.t_len, opt, scr, false, &qe, &te, lane,
		                          max(opt.min_chain_score * opt.a, opt.min_dp_max));   // only `score >= both thresholds` is asked
		__syncwarp();
		if (lane == 0) {
			int code = (sc >= opt.min_chain_score * opt.a && sc >= opt.min_dp_max) ? 2 : 0;
			if (code == 0) code = z.max_zdrop > opt.zdrop ? 1 : 0;
			mb_ztest_finish(T, code, opt, pass2_list, n_pass2, z.task);
		}
		__syncwarp();
	}
}

// ---- region CIGAR assembled in place at the first task's slot ----
MB_D void mb_append_cigar(uint32_t *dst, int &n_dst, const uint32_t *src, int n_src)
{
	if (n_src == 0) return;
	int k = 0;
	if (n_dst > 0 && (dst[n_dst - 1] & 0xf) == (src[0] & 0xf)) { dst[n_dst - 1] += (src[0] >> 4) << 4; k = 1; }
	for (; k < n_src; ++k) dst[n_dst++] = src[k];
}

// warp-cooperative mb_append_cigar: every lane runs the (uniform) bookkeeping, the copy is spread over the lanes.  The region
// CIGAR is compacted in place at the first task's slot, so dst + n_dst never passes src: a chunk is read completely before
// it is written.
MB_D void mb_append_cigar_warp(uint32_t *dst, int &n_dst, const uint32_t *src, int n_src, int lane)
{
	if (n_src == 0) return;
	int k = 0;
	__syncwarp();
	if (n_dst > 0 && (dst[n_dst - 1] & 0xf) == (src[0] & 0xf)) {
		const uint32_t add = (src[0] >> 4) << 4;
		__syncwarp();
		if (lane == 0) dst[n_dst - 1] += add;
		k = 1;
	}
	for (int b = k; b < n_src; b += 32) {
		const int i = b + lane;
		const uint32_t v = i < n_src ? src[i] : 0u;
		__syncwarp();
		if (i < n_src) dst[n_dst + (i - k)] = v;
		__syncwarp();
	}
	n_dst += n_src - k;
}

MB_D void mb_fix_cigar(Reg *r, uint32_t *cigar, const QView &qv, const TView &tv, int *qshift, int *tshift)
{
	int32_t toff = 0, qoff = 0, to_shrink = 0;
	int n_cigar = r->n_cigar;
	*qshift = *tshift = 0;
	if (n_cigar <= 1) return;
	for (int k = 0; k < n_cigar; ++k) {
		const uint32_t op = cigar[k] & 0xf, len = cigar[k] >> 4;
		if (len == 0) to_shrink = 1;
		if (op == 0) {
			toff += len, qoff += len;
		} else if (op == 1 || op == 2) {
			if (k > 0 && k < n_cigar - 1 && (cigar[k - 1] & 0xf) == 0 && (cigar[k + 1] & 0xf) == 0) {
				int l, prev_len = (int)(cigar[k - 1] >> 4);
				if (op == 1) {
					for (l = 0; l < prev_len; ++l)
						if (qv.at(qoff - 1 - l) != qv.at(qoff + (int)len - 1 - l)) break;
				} else {
					for (l = 0; l < prev_len; ++l)
						if (tv.at(toff - 1 - l) != tv.at(toff + (int)len - 1 - l)) break;
				}
				if (l > 0) cigar[k - 1] -= (uint32_t)l << 4, cigar[k + 1] += (uint32_t)l << 4, qoff -= l, toff -= l;
				if (l == prev_len) to_shrink = 1;
			}
			if (op == 1) qoff += len; else toff += len;
		} else if (op == 3) toff += len;
	}
	if (to_shrink) {
		int l = 0;
		for (int k = 0; k < n_cigar; ++k)
			if (cigar[k] >> 4 != 0) cigar[l++] = cigar[k];
		n_cigar = l;
		l = 0;
		for (int k = 0; k < n_cigar; ++k)
			if (k == n_cigar - 1 || (cigar[k] & 0xf) != (cigar[k + 1] & 0xf)) cigar[l++] = cigar[k];
			else cigar[k + 1] += cigar[k] >> 4 << 4;
		n_cigar = l;
	}
	if ((cigar[0] & 0xf) == 1 || (cigar[0] & 0xf) == 2) {
		const int32_t l = (int32_t)(cigar[0] >> 4);
		if ((cigar[0] & 0xf) == 1) {
			if (r->rev) r->qe -= l; else r->qs += l;
			*qshift = l;
		} else r->rs += l, *tshift = l;
		--n_cigar;
		for (int k = 0; k < n_cigar; ++k) cigar[k] = cigar[k + 1];
	}
	r->n_cigar = n_cigar;
}

// mm_fix_cigar, warp-cooperative.  What upstream does per indel op k flanked by two match ops: shift it left by
// l_k = min(prev_len_k, run_k), where run_k is how far the bases before the gap repeat the gap's own tail (sequence only:
// the offsets at which an op starts never change) and prev_len_k is the current length of the preceding match op, i.e. its
// original length plus the shift of the indel two ops earlier.  So: (A) every lane computes run_k for the indels of its
// block of ops (the only part that reads sequence), (B) lane 0 resolves the short recurrence over the stored runs, (C) the
// lanes apply -l / +l to the neighbouring match ops.  Zero-length ops and a leading indel (both rare) are finished by lane
// 0 exactly as upstream.  `scr` = 2 * n_cigar ints of scratch.
MB_D void mb_fix_cigar_warp(Reg *r, uint32_t *cigar, int32_t *scr, const QView &qv, const TView &tv, int *qshift, int *tshift, int lane)
{
	const unsigned FULL = 0xffffffffu;
	int n_cigar = r->n_cigar;
	*qshift = *tshift = 0;
	if (n_cigar <= 1) return;
	// plain (cached) accesses: every hand-over between lanes below is separated by a __syncwarp(), which orders the warp's
	// global-memory accesses; volatile would send each access to L2
	uint32_t *cg = cigar;
	int32_t *run = scr;
	const int per = (n_cigar + 31) / 32;
	const int lo = min(lane * per, n_cigar), hi = min(lo + per, n_cigar);
	int qsum = 0, tsum = 0;
	for (int k = lo; k < hi; ++k) {
		const uint32_t op = cg[k] & 0xf, len = cg[k] >> 4;
		if (op == 0) qsum += len, tsum += len;
		else if (op == 1) qsum += len;
		else if (op == 2 || op == 3) tsum += len;
	}
	int qoff = qsum, toff = tsum;
	#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		const int oq = __shfl_up_sync(FULL, qoff, d), ot = __shfl_up_sync(FULL, toff, d);
		if (lane >= d) qoff += oq, toff += ot;
	}
	qoff -= qsum, toff -= tsum;
	// (A) runs.  l_k = min(len(k-1) + l_{k-2}, run_k): wherever run_k < len(k-1) the carry from k-2 cannot matter and l_k = run_k
	// is final; only the (few) positions with run_k >= len(k-1) depend on their predecessor and are resolved in order below.
	int n_bound = 0;
	for (int k = lo; k < hi; ++k) {
		const uint32_t op = cg[k] & 0xf; const int len = (int)(cg[k] >> 4);
		int rn = -1; // -1: not an eligible indel
		if ((op == 1 || op == 2) && k > 0 && k < n_cigar - 1 && (cg[k - 1] & 0xf) == 0 && (cg[k + 1] & 0xf) == 0) {
			int l = 0;
			if (op == 1) { const int cap = qoff; while (l < cap && qv.at(qoff - 1 - l) == qv.at(qoff + len - 1 - l)) ++l; }
			else { const int cap = toff; while (l < cap && tv.at(toff - 1 - l) == tv.at(toff + len - 1 - l)) ++l; }
			rn = l;
			if (l >= (int)(cg[k - 1] >> 4)) ++n_bound;
		}
		run[k] = rn;
		if (op == 0) qoff += len, toff += len;
		else if (op == 1) qoff += len;
		else if (op == 2 || op == 3) toff += len;
	}
	// ordered list of the carry-dependent positions
	int b_off = n_bound;
	#pragma unroll
	for (int d = 1; d < 32; d <<= 1) { const int o = __shfl_up_sync(FULL, b_off, d); if (lane >= d) b_off += o; }
	const int n_bound_all = __shfl_sync(FULL, b_off, 31);
	b_off -= n_bound;
	int32_t *blist = scr + n_cigar;
	if (n_bound) for (int k = lo; k < hi; ++k) { const int rn = run[k]; if (rn >= 0 && rn >= (int)(cg[k - 1] >> 4)) blist[b_off++] = k; }
	int zero_len = 0;
	for (int k = lo; k < hi; ++k) if ((cg[k] >> 4) == 0) zero_len = 1; // (upstream tests the current length; a spurious flag only runs a no-op compaction)
	__syncwarp();
	// (B) resolve the carry-dependent positions in order
	int to_shrink = 0;
	if (lane == 0) {
		for (int b = 0; b < n_bound_all; ++b) {
			const int k = blist[b];
			const int carry = (k >= 2 && run[k - 2] > 0) ? run[k - 2] : 0; // final: k-2 is non-dependent or was resolved one step earlier
			const int prev_len = (int)(cg[k - 1] >> 4) + carry;
			const int rn = run[k];
			const int l = rn < prev_len ? rn : prev_len;
			if (l == prev_len) to_shrink = 1;
			run[k] = l;
		}
	}
	to_shrink = __any_sync(FULL, to_shrink | zero_len) ? 1 : 0;
	__syncwarp();
	// (C) apply
	for (int k = lo; k < hi; ++k) {
		const uint32_t w = cg[k];
		if ((w & 0xf) == 0) {
			int len = (int)(w >> 4);
			if (k + 1 < n_cigar) { const int l = run[k + 1]; if (l > 0) len -= l; }
			if (k > 0) { const int l = run[k - 1]; if (l > 0) len += l; }
			cg[k] = (uint32_t)len << 4;
		}
	}
	__syncwarp();
	if (to_shrink) {
		// (D) drop the ops whose length became zero, then merge neighbours of the same kind -- two warp-wide compactions
		// through the second scratch block (upstream: two sequential passes over the whole CIGAR)
		uint32_t *tmp = reinterpret_cast<uint32_t*>(scr + n_cigar); // the carry list of (B) is dead by now
		int cnt = 0;
		for (int k = lo; k < hi; ++k) cnt += (cg[k] >> 4) != 0;
		int off = cnt;
		#pragma unroll
		for (int d = 1; d < 32; d <<= 1) { const int o = __shfl_up_sync(FULL, off, d); if (lane >= d) off += o; }
		const int n1 = __shfl_sync(FULL, off, 31);
		off -= cnt;
		for (int k = lo; k < hi; ++k) { const uint32_t w = cg[k]; if (w >> 4) tmp[off++] = w; }
		__syncwarp();
		const int per1 = (n1 + 31) / 32;
		const int lo1 = min(lane * per1, n1), hi1 = min(lo1 + per1, n1);
		int heads = 0;
		for (int k = lo1; k < hi1; ++k) heads += k == 0 || (tmp[k] & 0xf) != (tmp[k - 1] & 0xf);
		int off2 = heads;
		#pragma unroll
		for (int d = 1; d < 32; d <<= 1) { const int o = __shfl_up_sync(FULL, off2, d); if (lane >= d) off2 += o; }
		const int n2 = __shfl_sync(FULL, off2, 31);
		off2 -= heads;
		for (int k = lo1; k < hi1; ++k) {
			const uint32_t w = tmp[k];
			if (k == 0 || (w & 0xf) != (tmp[k - 1] & 0xf)) { // head of a run of equal ops: the run's lengths add up
				uint32_t len = w >> 4;
				for (int k2 = k + 1; k2 < n1 && (tmp[k2] & 0xf) == (w & 0xf); ++k2) len += tmp[k2] >> 4;
				cigar[off2++] = len << 4 | (w & 0xf);
			}
		}
		n_cigar = n2;
		__syncwarp();
	}
	if (lane == 0) {
		if ((cigar[0] & 0xf) == 1 || (cigar[0] & 0xf) == 2) {
			const int32_t l = (int32_t)(cigar[0] >> 4);
			if ((cigar[0] & 0xf) == 1) {
				if (r->rev) r->qe -= l; else r->qs += l;
				*qshift = l;
			} else r->rs += l, *tshift = l;
			--n_cigar;
			for (int k = 0; k < n_cigar; ++k) cigar[k] = cigar[k + 1];
		}
		r->n_cigar = n_cigar;
	}
	__syncwarp();
}

// mm_update_extra for the regions stitched in this round, one WARP per region.  mm_fix_cigar (a sequential pass over the
// CIGAR ops) is done by lane 0; the base-level scan that yields blen / mlen / n_ambi / dp_max is split over the lanes by
// blocks of CIGAR ops.  dp_max is the maximum of the running score s <- max(s + d, 0): a lane summarises its block as the
// pair of max-plus maps  s_out = max(s_in + A, B)  and  block_max = max(s_in + C, D), which compose associatively, so one
// pass per lane plus a 32-step combine gives the exact sequential result.
#define UE_NEG (-(1 << 29))
__global__ void __launch_bounds__(128)
k_update_extra(AlignCtx c, ReadArrays ra, const int2 *__restrict__ work, int n_work, const RegPlan *__restrict__ plans, const DpTask *__restrict__ tasks,
               uint32_t *__restrict__ cigar_pool, const int32_t *__restrict__ perm)
{
	const unsigned FULL = 0xffffffffu;
	const int lane = threadIdx.x & 31;
	int wi = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
	if (wi >= n_work) return;
	if (perm) wi = perm[wi];
	const int read = work[wi].x, slot = work[wi].y;
	Reg *r = ra.regs + ra.reg_off[read] + slot;
	if ((r->cnt == 0 && !r->inv) || !r->has_p) return;
	const mb_opt_t &opt = c.opt;
	const int64_t roff = c.read_off[read];
	const int qlen = (int)(c.read_off[read + 1] - roff);
	const int rev = r->rev;
	uint32_t *cigar = cigar_pool + r->cigar_off;
	const int qs1 = rev ? qlen - r->qe : r->qs, rs1 = r->rs;
	QView qv; qv.codes = c.codes;
	if (!rev) qv.idx0 = roff + qs1, qv.step = 1, qv.comp = 0;
	else qv.idx0 = roff + qlen - 1 - qs1, qv.step = -1, qv.comp = 1;
	TView tv; tv.S = c.ix.S; tv.bytes = nullptr; tv.idx0 = (int64_t)c.ix.seq_off[r->rid] + rs1; tv.step = 1; tv.packed = 1;
	int qshift = 0, tshift = 0;
	__syncwarp();
	{
		// scratch for the per-op runs: the tail of this region's own CIGAR slots (the per-task CIGARs parked there have been
		// consumed by k_stitch); fall back to the sequential routine in the unlikely case that it does not fit
		const RegPlan &pl = plans[wi];
		const DpTask &tl = tasks[pl.task0 + pl.n_tasks - 1];
		const int64_t cap = tl.cigar_off + tl.qlen + tl.tlen + 1 - r->cigar_off;
		const int nc = r->n_cigar;
		if (pl.n_tasks > 0 && (int64_t)3 * nc + 2 <= cap) mb_fix_cigar_warp(r, cigar, reinterpret_cast<int32_t*>(cigar + nc), qv, tv, &qshift, &tshift, lane);
		else if (lane == 0) mb_fix_cigar(r, cigar, qv, tv, &qshift, &tshift);
	}
	__syncwarp();
	qshift = __shfl_sync(FULL, qshift, 0), tshift = __shfl_sync(FULL, tshift, 0);
	qv.idx0 += (int64_t)qshift * qv.step, tv.idx0 += (int64_t)tshift * tv.step;
	const int n_cigar = r->n_cigar;   // written by lane 0 before the __syncwarp() above
	const uint32_t *cg = cigar;
	// Work split: the CIGAR's base events (every base of an M, I or D run) are cut into 32 equal shares, so the lanes run the
	// same number of iterations whatever the run lengths are.  To find where a share starts, each lane first sums a block of
	// ops (events, query and reference bases), the warp scans the sums, and the lane walks at most one block of ops from the
	// start of the block that holds its first event.
	const int per = (n_cigar + 31) / 32;
	const int lo = min(lane * per, n_cigar), hi = min(lo + per, n_cigar);
	int qsum = 0, tsum = 0, esum = 0;
	for (int k = lo; k < hi; ++k) {
		const uint32_t op = cg[k] & 0xf, len = cg[k] >> 4;
		if (op == 0) qsum += len, tsum += len, esum += len;
		else if (op == 1) qsum += len, esum += len;
		else if (op == 2) tsum += len, esum += len;
		else if (op == 3) tsum += len;
	}
	int qoff = qsum, toff = tsum, eoff = esum;
	#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		const int oq = __shfl_up_sync(FULL, qoff, d), ot = __shfl_up_sync(FULL, toff, d), oe = __shfl_up_sync(FULL, eoff, d);
		if (lane >= d) qoff += oq, toff += ot, eoff += oe;
	}
	const int n_ev = __shfl_sync(FULL, eoff, 31);
	qoff -= qsum, toff -= tsum, eoff -= esum;
	const int share = (n_ev + 31) / 32;
	const int e0 = min(lane * share, n_ev), my_ev = min(share, n_ev - e0);
	// the block that holds event e0: the last one whose first event is <= e0 (empty blocks share their successor's offset)
	int blk = 0;
	#pragma unroll
	for (int L = 1; L < 32; ++L) if (__shfl_sync(FULL, eoff, L) <= e0) blk = L;
	int k = min(blk * per, n_cigar);
	int q = __shfl_sync(FULL, qoff, blk), t = __shfl_sync(FULL, toff, blk), ev = __shfl_sync(FULL, eoff, blk);
	uint32_t op = 0, len = 0, rem = 0;
	if (my_ev > 0) {
		for (;;) { // whole runs that end at or before e0
			const uint32_t cgk = cg[k];
			op = cgk & 0xf, len = cgk >> 4;
			const int l = op <= 2 ? (int)len : 0;
			if (l > 0 && ev + l > e0) break;
			ev += l;
			if (op == 0) q += len, t += len; else if (op == 1) q += len; else if (op == 2 || op == 3) t += len;
			++k;
		}
		++k;
		const int skip = e0 - ev;
		rem = len - skip;
		if (op != 2) q += skip;
		if (op != 1) t += skip;
	}
	const int sc_a = opt.a < 0 ? -opt.a : opt.a, sc_b = opt.b > 0 ? -opt.b : opt.b, sc_n = -(opt.sc_ambi > 0 ? opt.sc_ambi : -opt.sc_ambi);
	const int gap_o = opt.q, gap_e = opt.e;
	int blen = 0, mlen = 0, n_ambi_t = 0;
	int A = 0, B = UE_NEG, Cm = UE_NEG, D = UE_NEG;
	auto apply = [&](int d) { A += d; B = max(B + d, 0); Cm = max(Cm, A); D = max(D, B); };
	// one base per iteration.  A gap run charges its penalty with its last base; its other bases apply d = 0, which changes
	// nothing (the running score is never negative, so an extra clamp at 0 and an extra max candidate are harmless).
	// The kernel is bound by L2 sector traffic (every lane walks its own stretch of both sequences and the lines do not
	// survive in L1), so each sequence is read a 32-bit word at a time -- 4 query codes / 8 packed reference bases per load --
	// and the word is kept in a register until the walk leaves it.
	const uint8_t *qp = qv.codes + qv.idx0 + (int64_t)q * qv.step;
	const int64_t qstep = qv.step;
	const int cmask = qv.comp ? 3 : 0;          // 3 - c for c < 4; an ambiguous 4 becomes 7, still > 3
	int64_t tp = tv.idx0 + t;
	uintptr_t qwa = 0; uint32_t qw = 0;         // address and content of the query word held
	int64_t twi = -1; uint32_t tw = 0;          // index and content of the reference word held
	for (int it = 0; it < share; ++it) {
		__syncwarp();
		if (it < my_ev) {
			while (rem == 0) {
				const uint32_t cgk = cg[k++];
				op = cgk & 0xf, len = cgk >> 4, rem = len;
				if (op == 3) tp += len, rem = 0;
				else if (op > 3) rem = 0;
			}
			--rem;
			int cq = 0, ct = 0;
			if (op != 2) {
				const uintptr_t pa = reinterpret_cast<uintptr_t>(qp), wa = pa & ~(uintptr_t)3;
				if (wa != qwa) { qwa = wa; qw = *reinterpret_cast<const uint32_t*>(wa); }
				cq = (int)(qw >> ((pa & 3) << 3) & 0xffu) ^ cmask;
				qp += qstep;
			}
			if (op != 1) {
				const int64_t wi = tp >> 3;
				if (wi != twi) { twi = wi; tw = tv.S[wi]; }
				ct = (int)(tw >> ((tp & 7) << 2) & 0xfu);
				++tp;
			}
			const int ambi = (ct > 3 || cq > 3) ? 1 : 0;
			n_ambi_t += ambi, blen += 1 - ambi;
			int d;
			if (op == 0) {
				d = ambi ? sc_n : ct != cq ? sc_b : sc_a;
				mlen += (ambi == 0 && ct == cq) ? 1 : 0;
			} else d = rem == 0 ? -(gap_o + gap_e * (int)len) : 0;
			apply(d);
		}
	}
	// combine the blocks in lane order
	int s_run = 0, mx = 0;
	for (int L = 0; L < 32; ++L) {
		const int a = __shfl_sync(FULL, A, L), b = __shfl_sync(FULL, B, L), cm = __shfl_sync(FULL, Cm, L), dm = __shfl_sync(FULL, D, L);
		mx = max(mx, max(s_run + cm, dm));
		s_run = max(s_run + a, b);
	}
	#pragma unroll
	for (int d = 16; d > 0; d >>= 1) {
		blen += __shfl_xor_sync(FULL, blen, d), mlen += __shfl_xor_sync(FULL, mlen, d), n_ambi_t += __shfl_xor_sync(FULL, n_ambi_t, d);
	}
	if (lane == 0) r->blen = blen, r->mlen = mlen, r->n_ambi += n_ambi_t, r->dp_max = mx;
}

// one thread per region of this round: the part of mm_align1 after each mm_align_pair
__global__ void k_stitch(AlignCtx c, ReadArrays ra, const int2 *__restrict__ work, int n_work, const RegPlan *__restrict__ plans,
                         DpTask *__restrict__ tasks, uint32_t *__restrict__ cigar_pool,
                         int2 *__restrict__ next_work, int32_t *__restrict__ n_next, int *__restrict__ err, const int32_t *__restrict__ perm,
                         int4 *__restrict__ inv_list, int32_t *__restrict__ n_inv)
{
	// one WARP per region: every lane runs the same (uniform) control flow, lane 0 alone writes region state, and the CIGAR
	// copies -- the bulk of the work -- are spread over the lanes
	int wi = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
	const int lane = threadIdx.x & 31;
	if (wi >= n_work) return;
	if (perm) wi = perm[wi];
	const int read = work[wi].x, slot = work[wi].y;
	Reg *regs = ra.regs + ra.reg_off[read];
	Reg *r = regs + slot;
	const mb128 *a = ra.a + ra.a_roff[read];
	const RegPlan pl = plans[wi];
	const mb_opt_t &opt = c.opt;
	const int64_t roff = c.read_off[read];
	const int qlen = (int)(c.read_off[read + 1] - roff);
	const int r_cnt = r->cnt, r_as = r->as;
	__syncwarp();
	if (lane == 0) r->aligned = 1;
	if (r_cnt == 0) return;
	const int rev = r->rev, rid = r->rid;
	const int k2 = c.ix.k >> 1;
	DpTask *T = tasks + pl.task0;
	int ti = 0, n_cig = 0, dropped = 0;
	uint32_t *cig = pl.n_tasks > 0 ? cigar_pool + T[0].cigar_off : nullptr;
	int32_t rs = pl.rs, qs = pl.qs, re = pl.re, qe = pl.qe, rs1, qs1, re1, qe1;
	int32_t dp_score = 0;
	if (pl.has_left) {
		const DpTask &t = T[ti++];
		if (t.n_cigar > 0) { mb_append_cigar_warp(cig, n_cig, cigar_pool + t.cigar_off, t.n_cigar, lane); dp_score += t.max; }
		rs1 = rs - (t.reach_end ? t.mqe_t + 1 : t.max_t + 1);
		qs1 = qs - (t.reach_end ? qs - pl.qs0 : t.max_q + 1);
	} else rs1 = rs, qs1 = qs;
	re1 = rs, qe1 = qs;
	uint64_t bx = 0, by = 0; // the warp holds 32 anchors at a time (one coalesced load), handed out by shuffle
	for (int i = 1; i < pl.cnt1; ++i) {
		const int bi = (i - 1) & 31;
		if (bi == 0) {
			const int g = i + lane;
			if (g < pl.cnt1) { const mb128 v = a[pl.as1 + g]; bx = v.x, by = v.y; }
		}
		mb128 ai;
		ai.x = __shfl_sync(0xffffffffu, bx, bi), ai.y = __shfl_sync(0xffffffffu, by, bi);
		if ((ai.y & (MB_SEED_IGNORE | MB_SEED_TANDEM)) && i != pl.cnt1 - 1) continue;
		re = (int32_t)ai.x - k2, qe = (int32_t)ai.y - k2;
		re1 = re, qe1 = qe;
		if (i == pl.cnt1 - 1 || (ai.y & MB_SEED_LONG_JOIN) || (qe - qs >= opt.min_ksw_len && re - rs >= opt.min_ksw_len)) {
			const DpTask &t = T[ti++];
			if (t.n_cigar > 0) mb_append_cigar_warp(cig, n_cig, cigar_pool + t.cigar_off, t.n_cigar, lane);
			if (t.zdropped) {
				int j;
				for (j = i - 1; j >= 0; --j)
					if ((int32_t)a[pl.as1 + j].x <= rs + t.max_t) break;
				dropped = 1;
				if (j < 0) j = 0;
				dp_score += t.max;
				re1 = rs + (t.max_t + 1);
				qe1 = qs + (t.max_q + 1);
				if (lane == 0 && pl.cnt1 - (j + 1) >= opt.min_cnt) {
					const int n_split = pl.as1 + j + 1 - r_as;
					if (n_split > 0 && n_split < r_cnt) {
						const int ns = atomicAdd(&ra.n_regs[read], 1);
						if (ns >= (int)(ra.reg_off[read + 1] - ra.reg_off[read])) { *err = 2; atomicSub(&ra.n_regs[read], 1); } // unreachable: k_round_need sized the pool
						else {
							Reg *r2 = regs + ns;
							mb_split_reg(r, r2, n_split, qlen, a);
							r2->slot = ns;
							if (t.zdrop_code == 2) { // mm_align1_inv is tried between r and r2 once both are aligned (k_inv_*)
								r2->split_inv = 1, r2->inv_state = 1;
								inv_list[atomicAdd(n_inv, 1)] = make_int4(read, ns, slot, 0);
							}
							r->next_split = ns;
							next_work[atomicAdd(n_next, 1)] = make_int2(read, ns);
						}
					}
				}
				break;
			} else dp_score += t.score;
			rs = re, qs = qe;
		}
	}
	if (!dropped && pl.has_right) {
		const DpTask &t = T[pl.n_tasks - 1];
		if (t.n_cigar > 0) { mb_append_cigar_warp(cig, n_cig, cigar_pool + t.cigar_off, t.n_cigar, lane); dp_score += t.max; }
		re1 = re + (t.reach_end ? t.mqe_t + 1 : t.max_t + 1);
		qe1 = qe + (t.reach_end ? pl.qe0 - qe : t.max_q + 1);
	}
	__syncwarp();
	if (lane == 0) {
		r->rs = rs1, r->re = re1;
		if (rev) r->qs = qlen - qe1, r->qe = qlen - qs1;
		else r->qs = qs1, r->qe = qe1;
		if (n_cig > 0) {
			r->has_p = 1, r->n_cigar = n_cig, r->cigar_off = T[0].cigar_off, r->dp_score = dp_score; // mm_update_extra: k_update_extra
		}
	}
}

// ---- per-read kernels before / after the alignment rounds ----
struct ReadScratch {
	mb128 *b;          // mb128[n_a]  (same offsets as anchors)
	uint64_t *u;       // u64[n_a]
	uint64_t *scr;     // u64[3*n_a + 3*n_reads]: offset 3*a_roff[r] + 3*r
	int32_t *f, *p, *v, *t;
	Reg *regs_tmp;     // same offsets as regs
};

// G1a: chain backtrack, one warp per read; writes n_u per read.  Reads with more than 64 chain ends are listed in `big`
// (their upstream sorts leave the insertion-sort range) and redone by k_chain_bt_big with the exact radix replay.
__global__ void __launch_bounds__(128)
k_chain_bt(ReadArrays ra, ReadScratch rs, int n_reads, int min_cnt, int min_sc, int32_t *__restrict__ n_u, const int32_t *__restrict__ perm,
           int32_t *__restrict__ big, int32_t *__restrict__ n_big)
{
	int r = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
	const int lane = threadIdx.x & 31;
	if (r >= n_reads) return;
	if (perm) r = perm[r];
	const int64_t base = ra.a_roff[r];
	const int n = (int)(ra.a_roff[r + 1] - base);
	const int k = mb_chain_backtrack_warp(n, ra.a + base, rs.f + base, rs.p + base, rs.v + base, rs.t + base, rs.b + base, rs.u + base,
	                                      rs.scr + 3 * base + 3 * r, min_cnt, min_sc, nullptr, lane);
	if (lane == 0) {
		if (k < 0) { n_u[r] = 0; big[atomicAdd(n_big, 1)] = r; }
		else n_u[r] = k;
	}
}

#define BIG_SMEM_BYTES (MB_RS_WS_INTS * 4)
__global__ void __launch_bounds__(32)
k_chain_bt_big(ReadArrays ra, ReadScratch rs, int min_cnt, int min_sc, int32_t *__restrict__ n_u,
               const int32_t *__restrict__ big, const int32_t *__restrict__ n_big, int32_t *__restrict__ cursor)
{
	extern __shared__ __align__(16) int big_ws[];
	const int lane = threadIdx.x;
	const int n = *n_big;
	for (;;) {
		int k = 0;
		if (lane == 0) k = atomicAdd(cursor, 1);
		k = __shfl_sync(0xffffffffu, k, 0);
		if (k >= n) break;
		const int r = big[k];
		const int64_t base = ra.a_roff[r];
		const int na = (int)(ra.a_roff[r + 1] - base);
		const int nu = mb_chain_backtrack_warp(na, ra.a + base, rs.f + base, rs.p + base, rs.v + base, rs.t + base, rs.b + base, rs.u + base,
		                                       rs.scr + 3 * base + 3 * r, min_cnt, min_sc, big_ws, lane);
		if (lane == 0) n_u[r] = nu;
		__syncwarp();
	}
}

// region capacity per read: what the chains need now plus room for the first round of Z-drop splits (a region splits at
// most once per round; k_round_need checks every round and the host grows the pool when a read would run out)
__global__ void k_reg_cap(const int32_t *__restrict__ n_u, int n_reads, int tight, int32_t *__restrict__ cap)
{
	int r = blockIdx.x * blockDim.x + threadIdx.x;
	if (r >= n_reads) return;
	cap[r] = n_u[r] ? (tight ? n_u[r] : 2 * n_u[r] + 4) : 0;
}

// ---- region pool bookkeeping ----
// need[read] += 1 per work item (each region aligned in a round can split off one new region; each inversion candidate can
// add one inversion hit); flag the batch when a read's pool would overflow
__global__ void k_round_need(const int2 *__restrict__ work, int n_work, const int4 *__restrict__ inv_list, ReadArrays ra, int32_t *__restrict__ need, int32_t *__restrict__ flag)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_work) return;
	const int read = work ? work[i].x : inv_list[i].x;
	const int before = atomicAdd(&need[read], 1);
	if (ra.n_regs[read] + before + 1 > (int)(ra.reg_off[read + 1] - ra.reg_off[read])) *flag = 1;
}
__global__ void k_regs_newcap(ReadArrays ra, const int32_t *__restrict__ need, int n_reads, int32_t *__restrict__ cap)
{
	const int r = blockIdx.x * blockDim.x + threadIdx.x;
	if (r >= n_reads) return;
	const int old = (int)(ra.reg_off[r + 1] - ra.reg_off[r]), want = ra.n_regs[r] + 2 * need[r] + 2;
	cap[r] = need[r] && want > old ? want : old;
}
__global__ void k_regs_move(const Reg *__restrict__ src, const int64_t *__restrict__ src_off, Reg *__restrict__ dst, const int64_t *__restrict__ dst_off,
                            const int32_t *__restrict__ n_regs, int n_reads)
{
	const int r = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
	if (r >= n_reads) return;
	const int n = n_regs[r];
	const uint32_t *s = reinterpret_cast<const uint32_t*>(src + src_off[r]);
	uint32_t *d = reinterpret_cast<uint32_t*>(dst + dst_off[r]);
	const int words = n * (int)(sizeof(Reg) / 4);
	for (int i = lane; i < words; i += 32) d[i] = s[i];
}


// G1b: mm_gen_regs + chain_post (mm_set_parent, mm_select_sub, mm_join_long) + mm_squeeze_a of mm_align_skeleton, for one read
MB_D void mb_gen_regs_read(const AlignCtx &c, const ReadArrays &ra, const ReadScratch &rs, int r, int n, int32_t *n_a_sq, int2 *work, int32_t *n_work, int *ws)
{
	const mb_opt_t &opt = c.opt;
	const int64_t base = ra.a_roff[r];
	const int qlen = (int)(c.read_off[r + 1] - c.read_off[r]);
	mb128 *a = ra.a + base;
	Reg *regs = ra.regs + ra.reg_off[r];
	uint64_t *scr = rs.scr + 3 * base + 3 * r;
	int *iscr = ra.iscr + base;
	uint32_t hash = 0;
	hash ^= mb_wang32((uint32_t)qlen) + mb_wang32((uint32_t)opt.seed);
	hash = mb_wang32(hash);
	mb_gen_regs(hash, qlen, n, rs.u + base, a, regs, (mb128*)scr, ws);
	mb_set_parent(opt.mask_level, n, regs, opt.a * 2 + opt.b, scr, iscr, ws);
	mb_select_sub(opt.pri_ratio, c.ix.k * 2, opt.best_n, &n, regs, iscr);
	mb_join_long(&opt, qlen, &n, regs, a, scr, iscr, ws);
	// mm_align_skeleton: n_a = mm_squeeze_a(...)
	n_a_sq[r] = mb_squeeze_a(n, regs, a, scr, ws);
	for (int i = 0; i < n; ++i) regs[i].slot = i, regs[i].next_split = -1, regs[i].inv_after = -1, regs[i].inv_state = 0, regs[i].aligned = 0;
	ra.n_regs[r] = n;
	if (n > 0) {
		int w0 = atomicAdd(n_work, n);
		for (int i = 0; i < n; ++i) work[w0 + i] = make_int2(r, i);
	}
}

// one thread per read; reads with more than 64 chains go to k_gen_regs_big (exact radix replay of upstream's sorts)
__global__ void k_gen_regs(AlignCtx c, ReadArrays ra, ReadScratch rs, int n_reads, const int32_t *__restrict__ n_u, int32_t *__restrict__ n_a_sq,
                           int2 *__restrict__ work, int32_t *__restrict__ n_work, int32_t *__restrict__ big, int32_t *__restrict__ n_big)
{
	int r = blockIdx.x * blockDim.x + threadIdx.x;
	if (r >= n_reads) return;
	const int n = n_u[r];
	ra.n_regs[r] = 0;
	n_a_sq[r] = 0;
	if (n == 0) return;
	if (n > MB_RS_MIN_SIZE) { big[atomicAdd(n_big, 1)] = r; return; }
	mb_gen_regs_read(c, ra, rs, r, n, n_a_sq, work, n_work, nullptr);
}

__global__ void __launch_bounds__(32)
k_gen_regs_big(AlignCtx c, ReadArrays ra, ReadScratch rs, const int32_t *__restrict__ n_u, int32_t *__restrict__ n_a_sq,
               int2 *__restrict__ work, int32_t *__restrict__ n_work, const int32_t *__restrict__ big, const int32_t *__restrict__ n_big, int32_t *__restrict__ cursor)
{
	extern __shared__ __align__(16) int big_ws[];
	const int lane = threadIdx.x;
	const int n = *n_big;
	for (;;) {
		int k = 0;
		if (lane == 0) k = atomicAdd(cursor, 1);
		k = __shfl_sync(0xffffffffu, k, 0);
		if (k >= n) break;
		if (lane == 0) { const int r = big[k]; mb_gen_regs_read(c, ra, rs, r, n_u[r], n_a_sq, work, n_work, big_ws); }
		__syncwarp();
	}
}

// ---- work-list permutation: longest regions first ----
// The warp-per-region kernels (stitch, mm_update_extra) cost time in proportion to the region's length; in read order a launch
// ends with whatever long region happened to be scheduled late while most SMs sit idle.  A counting sort by descending query
// span (256-base buckets) gives the order in which those kernels visit the work list; the thread-per-region plan kernels keep
// read order (neighbouring threads then touch neighbouring anchors).
#define WORK_NB 256
MB_D int mb_work_bucket(const Reg *r) { const int b = (r->qe - r->qs) >> 8; return WORK_NB - 1 - (b < WORK_NB - 1 ? (b < 0 ? 0 : b) : WORK_NB - 1); }
__global__ void k_work_hist(const int2 *__restrict__ work, int n, ReadArrays ra, int32_t *__restrict__ hist)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	atomicAdd(&hist[mb_work_bucket(ra.regs + ra.reg_off[work[i].x] + work[i].y)], 1);
}
__global__ void k_work_scan(int32_t *__restrict__ hist) // one block of WORK_NB threads: counts -> exclusive offsets (used as cursors)
{
	__shared__ int32_t s[WORK_NB];
	const int t = threadIdx.x;
	s[t] = hist[t];
	__syncthreads();
	for (int d = 1; d < WORK_NB; d <<= 1) { const int v = t >= d ? s[t - d] : 0; __syncthreads(); s[t] += v; __syncthreads(); }
	hist[t] = s[t] - hist[t];
}
__global__ void k_work_scatter(const int2 *__restrict__ work, int n, ReadArrays ra, int32_t *__restrict__ cursor, int32_t *__restrict__ perm)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const int2 w = work[i];
	perm[atomicAdd(&cursor[mb_work_bucket(ra.regs + ra.reg_off[w.x] + w.y)], 1)] = i;
}

// after the alignment rounds: restore upstream's region order (a split-off region sits right after its source, an inversion
// hit right after the split_inv region it belongs to), then mm_filter_regs, mm_hit_sort, mm_set_parent, mm_select_sub,
// mm_set_sam_pri, mm_set_mapq (+ mm_set_inv_mapq), for one read
MB_D void mb_finish_read(const AlignCtx &c, const ReadArrays &ra, const ReadScratch &rs, int r, int n0, int rep_len, int32_t *n_hits, int32_t *n_hit_cigar, int *ws)
{
	const mb_opt_t &opt = c.opt;
	const int64_t base = ra.a_roff[r];
	const int qlen = (int)(c.read_off[r + 1] - c.read_off[r]);
	Reg *regs = ra.regs + ra.reg_off[r], *tmp = rs.regs_tmp + ra.reg_off[r];
	uint64_t *scr = rs.scr + 3 * base + 3 * r;
	int *iscr = ra.iscr + base;
	int n = 0;
	for (int i = 0; i < n0; ++i) {
		int s = i;
		while (s >= 0) {
			tmp[n++] = regs[s];
			if (regs[s].inv_after >= 0) tmp[n++] = regs[regs[s].inv_after];
			s = regs[s].next_split;
		}
	}
	for (int i = 0; i < n; ++i) regs[i] = tmp[i];
	mb_filter_regs(&opt, qlen, &n, regs);
	mb_hit_sort(&n, regs, (mb128*)scr, tmp, ws);
	mb_set_parent(opt.mask_level, n, regs, opt.a * 2 + opt.b, scr, iscr, ws);
	mb_select_sub(opt.pri_ratio, c.ix.k * 2, opt.best_n, &n, regs, iscr);
	mb_set_sam_pri(n, regs);
	mb_set_mapq(n, regs, opt.min_chain_score, opt.a, rep_len);
	mb_set_inv_mapq(n, regs, (mb128*)scr, ws);
	ra.n_regs[r] = n;
	n_hits[r] = n;
	int nc = 0;
	for (int i = 0; i < n; ++i) nc += regs[i].has_p ? regs[i].n_cigar : 0;
	n_hit_cigar[r] = nc;
}

// scratch note: scr holds 3*(n_a + 1) u64 per read and iscr n_a ints; a read's regions never outnumber its anchors (every
// region keeps at least one anchor; an inversion hit belongs to a split-off region of >= min_cnt anchors)
__global__ void k_finish(AlignCtx c, ReadArrays ra, ReadScratch rs, int n_reads, const int32_t *__restrict__ n0_regs, const int32_t *__restrict__ rep_len,
                         int32_t *__restrict__ n_hits, int32_t *__restrict__ n_hit_cigar, int32_t *__restrict__ big, int32_t *__restrict__ n_big)
{
	int r = blockIdx.x * blockDim.x + threadIdx.x;
	if (r >= n_reads) return;
	const int n_all = ra.n_regs[r];
	n_hits[r] = 0, n_hit_cigar[r] = 0;
	if (n_all == 0) return;
	if (n_all > MB_RS_MIN_SIZE) { big[atomicAdd(n_big, 1)] = r; return; }
	mb_finish_read(c, ra, rs, r, n0_regs[r], rep_len[r], n_hits, n_hit_cigar, nullptr);
}

__global__ void __launch_bounds__(32)
k_finish_big(AlignCtx c, ReadArrays ra, ReadScratch rs, const int32_t *__restrict__ n0_regs, const int32_t *__restrict__ rep_len,
             int32_t *__restrict__ n_hits, int32_t *__restrict__ n_hit_cigar, const int32_t *__restrict__ big, const int32_t *__restrict__ n_big, int32_t *__restrict__ cursor)
{
	extern __shared__ __align__(16) int big_ws[];
	const int lane = threadIdx.x;
	const int n = *n_big;
	for (;;) {
		int k = 0;
		if (lane == 0) k = atomicAdd(cursor, 1);
		k = __shfl_sync(0xffffffffu, k, 0);
		if (k >= n) break;
		if (lane == 0) { const int r = big[k]; mb_finish_read(c, ra, rs, r, n0_regs[r], rep_len[r], n_hits, n_hit_cigar, big_ws); }
		__syncwarp();
	}
}

// ------------------------------------------------------------------------------------------------
// mm_align1_inv (align.c): after the rounds, for every region s that a Z-drop split flagged as a potential inversion
// (split_inv), align the query stretch between the region before it and s -- read off the OPPOSITE strand -- to the
// reference stretch between them.  Upstream does this inside the mm_align_skeleton loop right after s has been aligned,
// with regs[i-1] as the left neighbour: that is the source region s was split from, unless the source's own inversion hit
// was inserted in between (then the test `r1->split & 1` fails), so candidates are decided in chain order:
//   k_inv_plan  (thread / candidate)  defer if the source's own candidate is undecided; upstream's eligibility tests
//   k_inv_ll    (warp / candidate)    local alignment of the reversed stretches -> start of the inverted block, DP task
//   DP          (k_dp / k_dp_ext)     extension alignment from there
//   k_inv_finish(thread / DP task)    the inversion hit as a new region, linked behind s; mm_update_extra follows
//   k_inv_close (thread / candidate)  mark the candidates of this pass decided
// ------------------------------------------------------------------------------------------------
struct InvTask { int32_t read, slot_s, slot_src, ql, tl; };

__global__ void k_inv_plan(AlignCtx c, ReadArrays ra, const int4 *__restrict__ list, int n, int4 *__restrict__ next_list, int32_t *__restrict__ n_next,
                           int4 *__restrict__ proc_list, int32_t *__restrict__ n_proc, InvTask *__restrict__ ll, int32_t *__restrict__ n_ll)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const int4 e = list[i];
	const mb_opt_t &opt = c.opt;
	const Reg *regs = ra.regs + ra.reg_off[e.x];
	const Reg *r2 = regs + e.y, *r1 = regs + e.z;
	if (r1->split_inv && r1->inv_state == 1) { next_list[atomicAdd(n_next, 1)] = e; return; } // the source's own inversion comes first
	proc_list[atomicAdd(n_proc, 1)] = e;
	if (r1->inv_after >= 0) return;   // the region before s is the source's inversion hit: its split bits are 0
	if (!(r1->split & 1) || !(r2->split & 2)) return;
	if (r1->id != r1->parent && r1->parent != MB_PARENT_TMP_PRI) return;
	if (r2->id != r2->parent && r2->parent != MB_PARENT_TMP_PRI) return;
	if (r1->rid != r2->rid || r1->rev != r2->rev) return;
	const int ql = r1->rev ? r1->qs - r2->qe : r2->qs - r1->qe, tl = r2->rs - r1->re;
	if (ql < opt.min_chain_score || ql > opt.max_gap) return;
	if (tl < opt.min_chain_score || tl > opt.max_gap) return;
	InvTask t; t.read = e.x, t.slot_s = e.y, t.slot_src = e.z, t.ql = ql, t.tl = tl;
	ll[atomicAdd(n_ll, 1)] = t;
}

__global__ void __launch_bounds__(32)
k_inv_ll(AlignCtx c, ReadArrays ra, const InvTask *__restrict__ ll, const int32_t *__restrict__ n_ll, int32_t *__restrict__ cursor, int *__restrict__ scr_pool,
         DpTask *__restrict__ tasks, InvTask *__restrict__ task_inv, int32_t *__restrict__ cig_cap, int32_t *__restrict__ n_dp)
{
	const int lane = threadIdx.x;
	int *scr = scr_pool + (size_t)blockIdx.x * (4 * LL_MAX_LEN);
	const mb_opt_t &opt = c.opt;
	const int n = *n_ll;
	for (;;) {
		int k = 0;
		if (lane == 0) k = atomicAdd(cursor, 1);
		k = __shfl_sync(0xffffffffu, k, 0);
		if (k >= n) break;
		const InvTask it = ll[k];
		const Reg *regs = ra.regs + ra.reg_off[it.read];
		const Reg *r2 = regs + it.slot_s, *r1 = regs + it.slot_src;
		const int64_t roff = c.read_off[it.read];
		const int qlen = (int)(c.read_off[it.read + 1] - roff);
		const int ql = it.ql, tl = it.tl;
		const int strand = r1->rev ? 0 : 1;                       // the inversion hit lies on the other strand
		const int p0 = r1->rev ? r2->qe : qlen - r2->qs;          // start of the stretch in that strand's coordinates
		const int64_t t0 = (int64_t)c.ix.seq_off[r1->rid] + r1->re;
		const uint8_t *codes = c.codes; const uint32_t *S = c.ix.S;
		auto qat = [&](int p) -> int { // base p of the read as seen on `strand`
			if (!strand) return codes[roff + p];
			const int b = codes[roff + qlen - 1 - p];
			return b < 4 ? 3 - b : 4;
		};
		auto tat = [&](int p) -> int { const int64_t g = t0 + p; return (int)(S[g >> 3] >> ((g & 7) << 2) & 0xf); };
		int qe, te;
		// both stretches reversed, as upstream does before ksw_ll_i16
		const int score = (ql > LL_MAX_LEN - 8 || tl > LL_MAX_LEN - 8) ? -1 :
			mb_ll_warp([&](int col) { return qat(p0 + ql - 1 - col); }, [&](int row) { return tat(tl - 1 - row); }, ql, tl, opt, scr, true, &qe, &te, lane);
		__syncwarp();
		if (lane == 0 && score >= opt.min_dp_max) {
			const int q_off = ql - (qe + 1), t_off = tl - (te + 1);  // q_off can be -1..-7 (padding columns, see ll.cuh)
			DpTask t; memset(&t, 0, sizeof(t));
			t.reg = -1, t.kind = 2;
			t.qlen = ql - q_off, t.tlen = tl - t_off, t.w = (int)(opt.bw * 1.5), t.zdrop = opt.zdrop, t.end_bonus = -1, t.flag = MB_EZ_EXTZ_ONLY;
			const int ps = p0 + q_off;
			if (!strand) t.q_idx0 = roff + ps, t.q_step = 1, t.q_comp = 0;
			else t.q_idx0 = roff + qlen - 1 - ps, t.q_step = -1, t.q_comp = 1;
			t.t_idx0 = t0 + t_off, t.t_step = 1, t.t_packed = 1;
			t.anchor_i = -1;
			t.qs = q_off, t.rs = t_off;    // kept for k_inv_finish
			t.skip = (opt.max_sw_mat > 0 && (int64_t)t.tlen * t.qlen > opt.max_sw_mat) ? 1 : 0;
			const int kk = atomicAdd(n_dp, 1);
			tasks[kk] = t; task_inv[kk] = it; cig_cap[kk] = t.qlen + t.tlen + 1;
		}
		__syncwarp();
	}
}

__global__ void k_inv_finish(AlignCtx c, ReadArrays ra, const DpTask *__restrict__ tasks, const InvTask *__restrict__ task_inv, int n_dp,
                             int2 *__restrict__ inv_work, RegPlan *__restrict__ inv_plans, int32_t *__restrict__ n_ok, int *__restrict__ err)
{
	const int k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= n_dp) return;
	const DpTask &T = tasks[k];
	if (T.n_cigar == 0) return; // upstream: "should never be here"
	const InvTask it = task_inv[k];
	Reg *regs = ra.regs + ra.reg_off[it.read];
	Reg *r2 = regs + it.slot_s; const Reg *r1 = regs + it.slot_src;
	const int ns = atomicAdd(&ra.n_regs[it.read], 1);
	if (ns >= (int)(ra.reg_off[it.read + 1] - ra.reg_off[it.read])) { *err = 2; atomicSub(&ra.n_regs[it.read], 1); return; } // unreachable: pool sized before the pass
	Reg v; memset(&v, 0, sizeof(v));
	const int q_off = T.qs, t_off = T.rs;
	v.id = -1, v.parent = MB_PARENT_UNSET, v.inv = 1, v.rev = !r1->rev, v.rid = r1->rid;
	if (v.rev == 0) { v.qs = r2->qe + q_off; v.qe = v.qs + T.max_q + 1; }
	else { v.qe = r2->qs - q_off; v.qs = v.qe - (T.max_q + 1); }
	v.rs = r1->re + t_off, v.re = v.rs + T.max_t + 1;
	v.has_p = 1, v.n_cigar = T.n_cigar, v.cigar_off = T.cigar_off, v.dp_score = T.max;
	v.next_split = -1, v.inv_after = -1, v.slot = ns, v.aligned = 1;
	regs[ns] = v;
	r2->inv_after = ns;
	const int w = atomicAdd(n_ok, 1);
	inv_work[w] = make_int2(it.read, ns);
	RegPlan pl; memset(&pl, 0, sizeof(pl));
	pl.n_tasks = 1, pl.task0 = k;
	inv_plans[w] = pl;
}

__global__ void k_inv_close(ReadArrays ra, const int4 *__restrict__ proc_list, const int32_t *__restrict__ n_proc)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= *n_proc) return;
	ra.regs[ra.reg_off[proc_list[i].x] + proc_list[i].y].inv_state = 2;
}

#define HIT_NF 23
// field order: read_idx rid rev qs qe rs re mapq mlen blen nm dp_max dp_max2 score score0 cnt subsc n_sub id parent is_primary sam_pri n_cigar
__global__ void k_write_hits(ReadArrays ra, int n_reads, const int64_t *__restrict__ hit_off, const int64_t *__restrict__ cig_off, int64_t n_hits_total,
                             int32_t *__restrict__ fields, int64_t *__restrict__ hit_cig_off, const uint32_t *__restrict__ cigar_pool, uint32_t *__restrict__ out_cigar)
{
	// one warp per read: lane 0 writes the hit fields, the CIGAR copy (the bulk of the bytes) is spread over the lanes
	const int r = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
	const int lane = threadIdx.x & 31;
	if (r >= n_reads) return;
	const Reg *regs = ra.regs + ra.reg_off[r];
	const int n = ra.n_regs[r];
	int64_t h = hit_off[r], co = cig_off[r];
	for (int i = 0; i < n; ++i, ++h) {
		const Reg &g = regs[i];
		if (lane == 0) {
			const int32_t vals[HIT_NF] = { r, g.rid, g.rev, g.qs, g.qe, g.rs, g.re, g.mapq, g.mlen, g.blen,
				g.blen - g.mlen + (g.has_p ? g.n_ambi : 0), g.has_p ? g.dp_max : 0, g.has_p ? g.dp_max2 : 0, g.score, g.score0, g.cnt, g.subsc, g.n_sub,
				g.id, g.parent, g.id == g.parent, g.sam_pri, g.has_p ? g.n_cigar : 0 };
			#pragma unroll
			for (int f = 0; f < HIT_NF; ++f) fields[(int64_t)f * n_hits_total + h] = vals[f];
			hit_cig_off[h] = co;
		}
		if (g.has_p) {
			const uint32_t *src = cigar_pool + g.cigar_off;
			const int nc = g.n_cigar;
			for (int k = lane; k < nc; k += 32) out_cigar[co + k] = src[k];
			co += nc;
		}
	}
}

// ---- monica's hit filter + best_hit + per-target sum (aligner.py:193-195,225-263,328-339) ----
// best_hit scans hits left to right with `<=` on NM/mlen and returns 0 (ambiguous) iff the minimum occurs twice or more.
// float(NM)/mlen comparisons are done exactly by cross-multiplication (NM, mlen < 2^31).
__global__ void k_count(int n_reads, const int64_t *__restrict__ hit_off, int64_t n_hits_total, const int32_t *__restrict__ fields,
                        const int64_t *__restrict__ read_off, int mapq_min, int mode,
                        unsigned long long *__restrict__ counts, unsigned long long *__restrict__ n_class, int8_t *__restrict__ read_class, int64_t *__restrict__ read_best)
{
	int r = blockIdx.x * blockDim.x + threadIdx.x;
	if (r >= n_reads) return;
	const int32_t *F_rid = fields + 1 * n_hits_total, *F_mapq = fields + 7 * n_hits_total, *F_mlen = fields + 8 * n_hits_total;
	const int32_t *F_nm = fields + 10 * n_hits_total, *F_pri = fields + 20 * n_hits_total;
	int64_t best = -1; int n_kept = 0; bool tie = false;
	for (int64_t h = hit_off[r]; h < hit_off[r + 1]; ++h) {
		if (!(F_pri[h] && F_mapq[h] >= mapq_min)) continue;
		++n_kept;
		if (best < 0) { best = h; tie = false; continue; }
		// inverse_identity(h) <= inverse_identity(best)  <=>  nm_h * mlen_b <= nm_b * mlen_h
		const long long lhs = (long long)F_nm[h] * F_mlen[best], rhs = (long long)F_nm[best] * F_mlen[h];
		if (lhs <= rhs) { tie = (lhs == rhs); best = h; }
	}
	int cls = 0;
	if (n_kept == 0) cls = 0;
	else if (n_kept >= 2 && tie) cls = 2, best = -1;
	else cls = 1;
	if (read_class) read_class[r] = (int8_t)cls;
	if (read_best) read_best[r] = best;
	atomicAdd(&n_class[cls == 1 ? 0 : cls == 0 ? 1 : 2], 1ULL);
	if (cls == 1 && counts) {
		unsigned long long inc = mode == 0 ? 1ULL : mode == 1 ? (unsigned long long)(read_off[r + 1] - read_off[r]) : mode == 2 ? (unsigned long long)F_mlen[best] : 0ULL;
		if (inc) atomicAdd(&counts[F_rid[best]], inc);
	}
}
