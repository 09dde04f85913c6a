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
// One warp per problem, as a systolic wavefront like the DP kernels: the columns are cut into strips of 256 (32 lanes x 8
// columns held in registers: H and E of the previous row); lane L works on row s - L at step s, so the horizontal-gap value F
// and the diagonal H(i-1, c-1) arrive from lane L-1 by shuffle, one step old.  Between strips, each row's last-column H, its
// outgoing F and the row's running (maximum, best slot) go through a scratch of 4 ints per row.  ~12 integer instructions
// per cell and no scans; a 1,000 x 1,000 problem takes a fraction of a millisecond.
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
// scr: 4 * LL_MAX_LEN ints.
#define LL_C 8
template <typename QF, typename TF>
MB_D int mb_ll_warp(QF qf, TF tf, int ql, int tl, const mb_opt_t &o, int *__restrict__ scr, bool want_pos, int *qe, int *te, int lane, int enough = 0x7fffffff)
{
	const unsigned FULL = 0xffffffffu;
	constexpr int C = LL_C;
	const int slen = (ql + 7) >> 3, P = slen << 3;
	const int ge = o.e, gqe = o.q + o.e;
	const int sa = o.a < 0 ? -o.a : o.a, sb = o.b > 0 ? -o.b : o.b, sn = -(o.sc_ambi > 0 ? o.sc_ambi : -o.sc_ambi);
	int4 *edge = reinterpret_cast<int4*>(scr);   // per row: H(i, last column so far), F leaving it, the row's maximum so far, its best slot
	*qe = *te = -1;
	if (ql <= 0 || tl <= 0) return 0;
	for (int i = lane; i < tl; i += 32) edge[i] = make_int4(0, 0, 0, -1);
	__syncwarp();
	int lmax = 0;
	const int n_steps = tl + 31;
	for (int cb = 0; cb < P; cb += 32 * C) {
		const int c0 = cb + lane * C;
		const bool live = c0 < P, last_strip = cb + 32 * C >= P;
		int qc[C], slot[C], H[C], E[C];
		#pragma unroll
		for (int c = 0; c < C; ++c) {
			const int col = c0 + c;
			qc[c] = col < ql ? qf(col) : 5;                                    // 5: padding column (scores 0 against anything)
			slot[c] = (want_pos && col < P) ? (col % slen) * 8 + col / slen : -1;
			H[c] = 0, E[c] = 0;
		}
		int F_o = 0, Hd_o = 0, rm_o = 0, rs_o = -1, hd_edge = 0;
		for (int s = 0; s < n_steps; ++s) {
			const int i = s - lane;
			int F = __shfl_up_sync(FULL, F_o, 1), Hd = __shfl_up_sync(FULL, Hd_o, 1), rm = __shfl_up_sync(FULL, rm_o, 1), rs = __shfl_up_sync(FULL, rs_o, 1);
			if (lane == 0) {
				Hd = hd_edge;                                                    // H(i-1, cb-1)
				if (s < tl) { const int4 ev = edge[s]; hd_edge = ev.x; F = ev.y; rm = ev.z; rs = ev.w; }
			}
			if (live && i >= 0 && i < tl) {
				const int ct = tf(i);
				int hd = Hd, f = F;
				#pragma unroll
				for (int c = 0; c < C; ++c) {
					const int sc = qc[c] == 5 ? 0 : (qc[c] > 3 || ct > 3) ? sn : (qc[c] == ct ? sa : sb);
					int h = hd + sc;
					h = max(h, E[c]); h = max(h, f); h = max(h, 0);
					hd = H[c];
					H[c] = h;
					const int t = max(h - gqe, 0);
					E[c] = max(E[c] - ge, t);
					f = max(f - ge, t);
					if (want_pos) { if (h > rm || (h == rm && slot[c] > rs)) rm = h, rs = slot[c]; }
					else rm = max(rm, h);
				}
				F_o = f, Hd_o = hd, rm_o = rm, rs_o = rs;
				lmax = max(lmax, rm);
				if (lane == 31 || c0 + C >= P) edge[i] = make_int4(H[C - 1], f, rm, rs);    // the strip's (or the matrix's) last live lane
			}
			if (!want_pos && (s & 31) == 31 && __any_sync(FULL, lmax >= enough)) return enough;
		}
		__syncwarp();
		(void)last_strip;
	}
	// upstream keeps the LAST row that reaches the global maximum, and that row's best slot
	int bh = -1, bi = -1;
	for (int i = lane; i < tl; i += 32) { const int h = edge[i].z; if (h > bh || (h == bh && i > bi)) bh = h, bi = i; }
	#pragma unroll
	for (int d = 16; d > 0; d >>= 1) {
		const int oh = __shfl_xor_sync(FULL, bh, d), oi = __shfl_xor_sync(FULL, bi, d);
		if (oh > bh || (oh == bh && oi > bi)) bh = oh, bi = oi;
	}
	if (want_pos && bi >= 0) {
		const int sl = edge[bi].w;
		*te = bi;
		if (sl >= 0) *qe = sl / 8 + (sl % 8) * slen;
	}
	return bh > 0 ? bh : 0;
}
