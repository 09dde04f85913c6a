// ll.cuh -- local (Smith-Waterman) single-affine alignment for the inversion logic, one warp per problem.
//
// Replaces minimap2-2.17 ksw2_ll_sse.c ksw_ll_qinit() + ksw_ll_i16() at its two call sites in align.c: mm_test_zdrop (is
// the most-dropped stretch of a gap fill an inversion?  only the score is used) and mm_align1_inv (where does the inverted
// block start?  score, query end and target end are used).  Reached in the reference through
// /root/reference/monica/genomes/aligner.py:193,215 (index.map -> mm_align_skeleton).
//
// Upstream runs Farrar's striped kernel on 8 int16 lanes; what that kernel computes is the plain recurrence
//   H(i,c) = max(H(i-1,c-1) + s(i,c), E(i,c), F(i,c), 0),  E(i+1,c) = max(E(i,c) - e, H(i,c) - q - e, 0),  F likewise along c
// as long as b <= q + 2e (a vertical gap opened on a cell that was reached through a horizontal gap can then never win; the
// striped kernel drops exactly that transition; checked on the host), over P = 8*ceil(qlen/8) columns: the query is padded
// to whole vectors and padding columns score 0 against every base.  The END POSITION upstream reports follows from three
// details that are kept here: the row maximum includes the padding columns (a maximum in the last real column is carried
// diagonally through the padding, so `te` advances by up to 7 rows and `qe` can point past the query end), `imax >= gmax`
// keeps the LAST row that reaches the global maximum, and within that row the LAST memory slot equal to the maximum wins,
// where column c lives in slot (c % slen) * 8 + c / slen of the striped layout.
//
// One warp per problem, row by row; the 32 lanes take 32 consecutive columns per step.  F along a row is a max-plus prefix
// (F(c) = max_{c'<c} h0(c') - q - (c-c') e, where h0 is the cell before F is applied), resolved with a 5-step warp scan.
// H and E of the previous row live in a scratch of 2*P ints per warp.
#pragma once
#include "align.cuh"

#define LL_MAX_LEN 5008   // both lengths are < max_gap (5000) at either call site; 8-column padding on top

static __host__ __device__ inline bool mb_ll_scoring_ok(const mb_opt_t &o)
{
	const int b = o.b > 0 ? o.b : -o.b;
	return o.q >= 0 && o.e > 0 && b <= o.q + 2 * o.e && o.a > 0 && o.a < 64;
}

// QF(c): query code of column c (0 <= c < ql), TF(i): target code of row i (0 <= i < tl).  Returns the maximum score;
// *qe / *te as upstream (or -1).  want_pos = false skips the end-position bookkeeping.  All lanes get the same results.
// `enough`: stop as soon as the maximum reaches it (the returned score is then a lower bound >= enough).
template <typename QF, typename TF>
MB_D int mb_ll_warp(QF qf, TF tf, int ql, int tl, const mb_opt_t &o, int *__restrict__ scr, bool want_pos, int *qe, int *te, int lane, int enough = 0x7fffffff)
{
	const unsigned FULL = 0xffffffffu;
	const int slen = (ql + 7) >> 3, P = slen << 3;
	const int gq = o.q, ge = o.e, gqe = o.q + o.e;
	int *H = scr, *E = scr + P;
	*qe = *te = -1;
	if (ql <= 0) return 0;
	for (int c = lane; c < P; c += 32) H[c] = 0, E[c] = 0;
	__syncwarp();
	int gmax = 0, best_te = -1, best_slot = -1;
	for (int i = 0; i < tl; ++i) {
		const int ct = tf(i);
		int carry_h = 0;                 // H(i-1, cb-1)
		int carry_f = -(1 << 28);        // F entering column cb, stored as F + cb*e
		int row_max = 0, row_slot = -1;  // per lane: best H of this row and the largest slot that holds it
		for (int cb = 0; cb < P; cb += 32) {
			const int c = cb + lane;
			const bool act = c < P;
			const int hp = act ? H[c] : 0;
			int diag = __shfl_up_sync(FULL, hp, 1);
			if (lane == 0) diag = carry_h;
			carry_h = __shfl_sync(FULL, hp, 31);
			const int e = act ? E[c] : 0;
			const int s = (act && c < ql) ? mb_mat(ct, qf(c), o) : 0;
			int h0 = diag + s; h0 = h0 > e ? h0 : e;                 // >= 0: E never goes below 0
			// F(c) + c*e = max(carry_f, max_{cb <= c' < c} (h0(c') - q + c'*e)): exclusive prefix max over the lanes
			int v = act ? h0 - gq + c * ge : -(1 << 28);
			int incl = v;
			#pragma unroll
			for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(FULL, incl, d); if (lane >= d) incl = incl > t ? incl : t; }
			int excl = __shfl_up_sync(FULL, incl, 1);
			if (lane == 0) excl = -(1 << 28);
			excl = excl > carry_f ? excl : carry_f;
			const int f = excl - c * ge;
			int h = h0 > f ? h0 : f; h = h > 0 ? h : 0;
			{ const int nf = __shfl_sync(FULL, incl, 31); carry_f = nf > carry_f ? nf : carry_f; }
			if (act) {
				H[c] = h;
				int en = e - ge, t2 = h - gqe;
				en = en > t2 ? en : t2;
				E[c] = en > 0 ? en : 0;
				if (want_pos) {
					if (h >= row_max) {
						const int slot = (c % slen) * 8 + c / slen;
						if (h > row_max || slot > row_slot) row_slot = slot;
						row_max = h;
					}
				} else row_max = row_max > h ? row_max : h;
			}
		}
		// row maximum (and, among the cells that hold it, the largest slot)
		int key_h = row_max, key_s = row_slot;
		#pragma unroll
		for (int d = 16; d > 0; d >>= 1) {
			const int oh = __shfl_xor_sync(FULL, key_h, d), os = __shfl_xor_sync(FULL, key_s, d);
			if (oh > key_h || (oh == key_h && os > key_s)) key_h = oh, key_s = os;
		}
		if (key_h >= gmax) gmax = key_h, best_te = i, best_slot = key_s;
		if (gmax >= enough) break;   // score-only callers that test a threshold need no more
		__syncwarp();
	}
	if (want_pos && best_slot >= 0) { *te = best_te; *qe = best_slot / 8 + (best_slot % 8) * slen; }
	else if (want_pos) *te = best_te;
	return gmax;
}
