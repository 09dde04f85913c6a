// dp_band.cuh -- K4 packed path for LARGE and BAND-LIMITED gap fills (ksw_extd2 with KSW_EZ_APPROX_MAX on windows of up to
// 2048 x 2048, w = 751 or wider), the tasks that k_dp_fast's "band never limits, target <= 768" rule leaves out.  At 15 %
// error / 50 kb reads (BASELINE configs[4]) and on reads with unrelated or inverted blocks these windows hold most of the
// cells that would otherwise go through the block-emulating kernel k_dp (align.cuh), which is ~50x slower per cell.
//
// Same machinery as k_dp_fast (dp_fast.cuh): register-resident systolic wavefront, lane L owns 8 consecutive target
// columns, two tasks per warp in 16x2 packed integer SIMD, values stored as 8*(value + bias) + priority tag.  On top:
//
//   * Column strips, pipelined over the warps of a CTA.  The target is cut into strips of 256 columns (32 lanes x 8); a strip
//     is one pass over all query rows and belongs to one warp.  What strip k+1 needs of strip k -- x, v, x2 of its last column,
//     row by row -- goes through a per-strip buffer in global memory; warp k publishes its progress every 32 steps (a counter
//     in shared memory) and warp k+1 follows a few dozen rows behind, so all strips of a window are in flight at once and a
//     2048 x 2048 window costs about 2048 + 8*32 steps of latency instead of 8 * 2079.
//   * upstream's band, exactly.  ksw_extd2_sse walks anti-diagonals r and computes, for each, the 16-ALIGNED column range
//     [st(r), en(r)] around the in-band range [st0(r), en0(r)]: up to 15 columns on either side of the band are computed
//     too, from whatever their neighbours hold, and in-band cells at the band edge read them.  Which cell is computed on
//     which diagonal is a closed-form predicate of (t, r); a column is computed over one contiguous range of diagonals, so
//     per-column state in registers (updated only when the predicate holds) reproduces upstream's arrays: a neighbour that
//     was not computed on the previous diagonal reads as upstream's initial value (this covers both "never computed yet"
//     and upstream's explicit `st-1 not in [last_st, last_en]` rule).  Substitution scores of the out-of-band lanes follow
//     upstream's fill rule: s[t] is refreshed on every diagonal whose fill range [st0, st0 + 16*ceil((en0-st0+1)/16)) holds
//     t -- computed or not -- and a computed lane outside that range reads the value left behind by the last refresh (one
//     register per column; its initial content is what the rows ABOVE the matrix left there: upstream's first blocks also
//     score columns t > r against the zero padding of the reversed query, i.e. against base 0).  Rows above or below the
//     matrix and padding columns are computed by upstream as well but never feed a real cell; they are skipped.
//   * z is clamped to the match score like upstream does (a no-op inside a consistent matrix, not at a band edge), and all
//     packed subtractions are per-half (edge cells break the "differences are non-negative" invariant k_dp_fast relies on).
//   * the end score is H(corner) summed along a path that stays in the band: the main diagonal (z of the cells (t,t)), then
//     the last row or column; upstream's approximate-max walk telescopes to the same potential.
//   * traceback with upstream's forced moves outside [st(r), en(r)].
// Lanes whose eight cells are all well inside the band take a shortcut without the predicates.
// Eligibility (dpb_class): flag == KSW_EZ_APPROX_MAX, |tlen - qlen| < w (else upstream bails out with st > en), both
// lengths <= 2048, no ambiguous base, scoring fits the packed tables.  Everything else stays with k_dp.
#pragma once
#include "dp_fast.cuh"

#define DPB_C 8                      // columns per lane
#define DPB_CW 4                     // direction words per lane per step
#define DPB_STRIP (32 * DPB_C)       // 256 columns per strip
#define DPB_MAX_LEN 2048
#define DPB_NCLASS 2                 // scratch classes: max(qlen, tlen) <= 1024, <= 2048
#define DPB_EDGE_WORDS (4 * (DPB_MAX_LEN + 64))

struct DpbGeo { int Q, T, w; };
MB_D void dpb_range(const DpbGeo &g, int r, int &st0, int &en0)
{
	int a = r - g.Q + 1, b = (r - g.w + 1) >> 1;
	st0 = a > b ? a : b; st0 = st0 > 0 ? st0 : 0;
	int c = g.T - 1 < r ? g.T - 1 : r, d = (r + g.w) >> 1;
	en0 = c < d ? c : d;
}
// is column t computed on diagonal r (16-aligned block range), and is its score fresh?
MB_D void dpb_flags(const DpbGeo &g, int t, int r, bool &comp, bool &fill, bool &low)
{
	int st0, en0; dpb_range(g, r, st0, en0);
	comp = t >= (st0 & ~15) && t <= (((en0 + 16) & ~15) - 1);
	fill = t >= st0 && t < st0 + ((((en0 - st0) >> 4) + 1) << 4);
	low = t < st0;
}
MB_D uint32_t dpb_blend(uint32_t m, uint32_t a, uint32_t b) { return (a & m) | (b & ~m); }   // per-half select: m = 0xffff per chosen half
MB_D uint32_t dpb_mask(bool a, bool b) { return (a ? 0x0000ffffu : 0u) | (b ? 0xffff0000u : 0u); }

static __host__ __device__ inline int dpb_class(int qlen, int tlen, int w, int flag, int skip)
{
	if (skip || flag != MB_EZ_APPROX_MAX || qlen <= 0 || tlen <= 0) return -1;
	if (w < 0) w = tlen > qlen ? tlen : qlen;
	const int mx = tlen > qlen ? tlen : qlen, df = tlen > qlen ? tlen - qlen : qlen - tlen;
	if (mx > DPB_MAX_LEN || df >= w || w < 64) return -1;
	return mx <= 1024 ? 0 : 1;
}

#define DPB_NW 4                      // warps per CTA; warp w takes strips w, w + 4 (a 2048-column window has 8)
#define DPB_MAX_STRIPS (DPB_MAX_LEN / DPB_STRIP)
__global__ void __launch_bounds__(DPB_NW * 32, 3)
k_dp_band(DpTask *__restrict__ tasks, const int32_t *__restrict__ order, const int32_t *__restrict__ n_order, int32_t *__restrict__ work_ctr,
          const uint8_t *__restrict__ codes, const uint32_t *__restrict__ S, const uint8_t *__restrict__ pool,
          uint32_t *__restrict__ p_scr, size_t p_stride_words, uint32_t *__restrict__ cigar_pool, DpScoring sc, unsigned long long *__restrict__ cells_out)
{
	constexpr int C = DPB_C, CW = DPB_CW;
	const unsigned FULL = 0xffffffffu;
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	uint32_t *EDGE = p_scr + (size_t)blockIdx.x * p_stride_words;      // [strip][row] x (X, V, X2, pad)
	constexpr int NW = DPB_NW;
	uint32_t *P = EDGE + (size_t)DPB_MAX_STRIPS * DPB_EDGE_WORDS;
	__shared__ volatile int s_prog[DPB_MAX_STRIPS];   // rows of strip k whose last-column values are in EDGE
	__shared__ int s_task;
	__shared__ long long s_acc[NW][2];
	unsigned long long cells = 0;
	const int n_total = *n_order;
	int q = sc.q, e = sc.e, q2 = sc.q2, e2 = sc.e2;
	if (q2 + e2 < q + e) { int t_ = q; q = q2, q2 = t_; t_ = e; e = e2, e2 = t_; }
	int long_thres = e != e2 ? (q2 - q) / (e - e2) - 1 : 0;
	if (q2 + e2 + long_thres * e2 > q + e + long_thres * e) ++long_thres;
	const int long_diff = long_thres * (e - e2) - (q2 - q) - e2;
	const int B = q + e;
	const uint32_t X_INIT = dpf_pack2(8 * (-q - e + B) + 3), Y_INIT = dpf_pack2(8 * (-q - e + B) + 2);
	const uint32_t X2_INIT = dpf_pack2(8 * (-q2 - e2 + B) + 1), Y2_INIT = dpf_pack2(8 * (-q2 - e2 + B) + 0);
	const uint32_t V_INIT = dpf_pack2(8 * (-q - e + B)), U_INIT = V_INIT;
	const uint32_t FL1 = X_INIT, FL2 = Y_INIT, FL3 = X2_INIT, FL4 = Y2_INIT;
	const uint32_t NFL1 = dpf_pack2(-(8 * (-q - e + B) + 3)), NFL2 = dpf_pack2(-(8 * (-q - e + B) + 2));
	const uint32_t NFL3 = dpf_pack2(-(8 * (-q2 - e2 + B) + 1)), NFL4 = dpf_pack2(-(8 * (-q2 - e2 + B) + 0));
	const uint32_t KE1 = dpf_pack2(8 * (B - e)), KE2 = dpf_pack2(8 * (B - e2));
	const uint32_t EIGHT = dpf_pack2(8);
	const uint32_t MCHB = (uint32_t)(8 * (sc.sc_mch + 2 * B) + 4), MISB = (uint32_t)(8 * (sc.sc_mis + 2 * B) + 4), NB = (uint32_t)(8 * (sc.sc_N + 2 * B) + 4);
	const uint32_t ZEROB = (uint32_t)(8 * (0 + 2 * B) + 4);          // upstream's initial s[] = 0
	const uint32_t MIS4 = MISB * 0x01010101u, N4 = NB * 0x01010101u, MDIFF = MCHB - MISB;
	const uint32_t MCHC = dpf_pack2(8 * (sc.sc_mch + 2 * B));
	auto rowbnd = [&](int r) { return r == 0 ? -q - e : r < long_thres ? -e : r == long_thres ? long_diff : -e2; };
	for (;;) {
		__syncthreads();
		if (threadIdx.x == 0) s_task = atomicAdd(work_ctr, 1) * 2;
		if (threadIdx.x < DPB_MAX_STRIPS) s_prog[threadIdx.x] = 0;
		__syncthreads();
		const int oi = s_task;
		if (oi >= n_total) break;
		const bool hasB = oi + 1 < n_total;
		DpTask &TA = tasks[order[oi]];
		DpTask &TB = tasks[order[hasB ? oi + 1 : oi]];
		DpbGeo gA, gB;
		gA.Q = TA.qlen, gA.T = TA.tlen, gA.w = TA.w < 0 ? max(TA.qlen, TA.tlen) : TA.w;
		gB.Q = TB.qlen, gB.T = TB.tlen, gB.w = TB.w < 0 ? max(TB.qlen, TB.tlen) : TB.w;
		const int Qm = max(gA.Q, gB.Q), TLm = max(gA.T, gB.T);
		QView qvA; qvA.codes = TA.q_comp == 2 ? pool : codes; qvA.idx0 = TA.q_idx0; qvA.step = TA.q_step; qvA.comp = TA.q_comp == 1;
		QView qvB; qvB.codes = TB.q_comp == 2 ? pool : codes; qvB.idx0 = TB.q_idx0; qvB.step = TB.q_step; qvB.comp = TB.q_comp == 1;
		TView tvA; tvA.S = S; tvA.bytes = pool; tvA.idx0 = TA.t_idx0; tvA.step = TA.t_step; tvA.packed = TA.t_packed;
		TView tvB; tvB.S = S; tvB.bytes = pool; tvB.idx0 = TB.t_idx0; tvB.step = TB.t_step; tvB.packed = TB.t_packed;
		const int n_steps = Qm + 31;
		const int n_strips = (TLm + DPB_STRIP - 1) / DPB_STRIP;
		const size_t strip_words = (size_t)n_steps * 32 * CW;
		long long accA = 0, accB = 0;        // sum of the biased potentials along the score path
		auto row_tables = [&](int r, uint32_t &la, uint32_t &lb) {
			const int a = r < gA.Q ? qvA.at(r) : 0, b = r < gB.Q ? qvB.at(r) : 0;
			la = a < 4 ? MIS4 + (MDIFF << (a * 8)) : N4, lb = b < 4 ? MIS4 + (MDIFF << (b * 8)) : N4;
		};
		for (int strip = wid; strip < n_strips; strip += NW) {
			uint32_t *EDGE_IN = EDGE + (size_t)(strip > 0 ? strip - 1 : 0) * DPB_EDGE_WORDS, *EDGE_OUT = EDGE + (size_t)strip * DPB_EDGE_WORDS;
			const int t0 = strip * DPB_STRIP + lane * C;
			uint32_t SEL[C], U[C], Y[C], Y2[C], SZ[C];
			#pragma unroll
			for (int c = 0; c < C; ++c) {
				const int t = t0 + c;
				const uint32_t a = t < gA.T ? (uint32_t)tvA.at(t) & 3u : 0u, b = t < gB.T ? (uint32_t)tvB.at(t) & 3u : 0u;
				SEL[c] = a | (8u | a) << 4 | (4u + b) << 8 | (12u + b) << 12;
				{ // s[t] as the rows above the matrix leave it: scored against base 0 if some diagonal r < t had t in its fill range, else upstream's initial 0
					bool fa = false, fb = false, c_, l_, f_;
					for (int r = t > 15 ? t - 15 : 0; r < t; ++r) {
						if (t < gA.T) { dpb_flags(gA, t, r, c_, f_, l_); fa |= f_; }
						if (t < gB.T) { dpb_flags(gB, t, r, c_, f_, l_); fb |= f_; }
					}
					SZ[c] = (fa ? (a == 0u ? MCHB : MISB) : ZEROB) | (fb ? (b == 0u ? MCHB : MISB) : ZEROB) << 16;
				}
				// upstream resets u[t] to the column boundary on diagonal r = t only if the block range reaches it (en(t) >= t)
				int st0, en0;
				dpb_range(gA, t, st0, en0); const bool ra = (((en0 + 16) & ~15) - 1) >= t;
				dpb_range(gB, t, st0, en0); const bool rb = (((en0 + 16) & ~15) - 1) >= t;
				const uint32_t ub = dpf_pack2(8 * (rowbnd(t) + B));
				U[c] = dpb_blend(dpb_mask(ra, rb), ub, U_INIT);
				Y[c] = Y_INIT, Y2[c] = Y2_INIT;
			}
			uint32_t XLo = 0, VLo = 0, X2Lo = 0;
			uint32_t LAc, LBc, LAn = 0, LBn = 0, LAo = 0, LBo = 0;
			row_tables(lane, LAc, LBc);
			uint32_t *dst = P + (size_t)strip * strip_words + (size_t)lane * CW;
			const bool lane_live = t0 < TLm;
			for (int s = 0; s < n_steps; ++s, dst += 32 * CW) {
				const int j = s - lane;
				if ((s & 31) == 0) {
					if (s) LAc = LAn, LBc = LBn;
					if (s + 32 < Qm) row_tables(s + 32 + lane, LAn, LBn);
					if (strip > 0 && s < Qm) { // rows s .. s+31 of the previous strip must have been published
						const int need = min(s + 32, Qm);
						while (s_prog[strip - 1] < need) { }
						__threadfence_block();
					}
				}
				const uint32_t LA0 = __shfl_sync(FULL, LAc, s & 31), LB0 = __shfl_sync(FULL, LBc, s & 31);
				uint32_t LA = __shfl_up_sync(FULL, LAo, 1), LB = __shfl_up_sync(FULL, LBo, 1);
				uint32_t XL = __shfl_up_sync(FULL, XLo, 1), VL = __shfl_up_sync(FULL, VLo, 1), X2L = __shfl_up_sync(FULL, X2Lo, 1);
				if (lane == 0) {
					LA = LA0, LB = LB0;
					if (strip > 0 && s < Qm) { const uint4 ev = __ldcg(reinterpret_cast<const uint4*>(EDGE_IN + 4 * s)); XL = ev.x, VL = ev.y, X2L = ev.z; }
				}
				LAo = LA, LBo = LB;
				if (lane_live && j >= 0 && j < Qm) {
					uint32_t wv[CW];
					uint32_t wprev = 0;
					// well inside both bands (or outside a matrix, where anything may be computed): no predicates needed
					const int dlo = t0 - j, dhi = t0 + C - 1 - j;
					const bool inA = j < gA.Q && t0 < gA.T, inB = hasB && j < gB.Q && t0 < gB.T;
					const bool plainA = !inA || (t0 + C <= gA.T && dlo > -(gA.w - 40) && dhi < gA.w - 40);
					const bool plainB = !inB || (t0 + C <= gB.T && dlo > -(gB.w - 40) && dhi < gB.w - 40);
					// beyond the block ranges on either side of the band (the corners of a large window): upstream computes nothing there
					const bool voidA = !inA || dlo > gA.w + 40 || dhi < -(gA.w + 40), voidB = !inB || dlo > gB.w + 40 || dhi < -(gB.w + 40);
					if (voidA && voidB) {
						wv[0] = wv[1] = wv[2] = wv[3] = 0u;
					} else if (plainA && plainB) {
						if (t0 == 0) { XL = X_INIT, X2L = X2_INIT; VL = dpf_pack2(8 * (rowbnd(j) + B)); }   // column 0: the row boundary
						#pragma unroll
						for (int c = 0; c < C; ++c) {
							const uint32_t z0 = dpf_prmt(LA, LB, SEL[c]);
							SZ[c] = z0;     // in-band cells are inside the fill range (a half outside its matrix never reads it again)
							const uint32_t a = __vadd2(XL, VL), b = __vadd2(Y[c], U[c]), a2 = __vadd2(X2L, VL), b2 = __vadd2(Y2[c], U[c]);
							const uint32_t zt = __vimax3_s16x2(__vimax3_s16x2(z0, a, b), a2, b2);
							const uint32_t zc = zt & 0xfff8fff8u;
							const uint32_t zl = __vmins2(zc, MCHC);
							const uint32_t un = __vsub2(zl, VL), vn = __vsub2(zl, U[c]);
							const uint32_t nz1 = __vsub2(KE1, zl), nz2 = __vsub2(KE2, zl);
							const uint32_t r1 = __viaddmax_s16x2(a, nz1, FL1), r2 = __viaddmax_s16x2(b, nz1, FL2);
							const uint32_t r3 = __viaddmax_s16x2(a2, nz2, FL3), r4 = __viaddmax_s16x2(b2, nz2, FL4);
							const uint32_t g1 = __viaddmin_s16x2(r1, NFL1, EIGHT), g2 = __viaddmin_s16x2(r2, NFL2, EIGHT);
							const uint32_t g3 = __viaddmin_s16x2(r3, NFL3, EIGHT), g4 = __viaddmin_s16x2(r4, NFL4, EIGHT);
							const uint32_t wd = (zt - zc) + g1 + 2u * g2 + 4u * g3 + 8u * g4;
							XL = r1, X2L = r3, Y[c] = r2, Y2[c] = r4, U[c] = un, VL = vn;
							if (c & 1) wv[c >> 1] = dpf_prmt(wprev, wd, 0x6240u);
							wprev = wd;
							const int t = t0 + c;
							if (t == j) { if (inA) accA += (int)(zl & 0xffffu); if (inB) accB += (int)(zl >> 16); }
							if (inA && ((j == gA.Q - 1 && t >= gA.Q) )) accA += (int)(int16_t)(un & 0xffffu);
							if (inB && ((j == gB.Q - 1 && t >= gB.Q) )) accB += (int)(int16_t)(un >> 16);
							if (inA && t == gA.T - 1 && j >= gA.T) accA += (int)(int16_t)(vn & 0xffffu);
							if (inB && t == gB.T - 1 && j >= gB.T) accB += (int)(int16_t)(vn >> 16);
						}
					} else {
						// general cells: upstream's computed / fresh-score / initial-neighbour predicates per half
						bool pcA = false, pcB = false;   // was the left neighbour (t-1, j) computed on its diagonal?
						if (t0 > 0) {
							bool f_, l_;
							if (inA || (j < gA.Q && t0 - 1 < gA.T)) dpb_flags(gA, t0 - 1, t0 - 1 + j, pcA, f_, l_);
							if (hasB && j < gB.Q && t0 - 1 < gB.T) dpb_flags(gB, t0 - 1, t0 - 1 + j, pcB, f_, l_);
						}
						#pragma unroll
						for (int c = 0; c < C; ++c) {
							const int t = t0 + c, r = t + j;
							bool cA = false, fA = true, lA = false, cB = false, fB = true, lB = false;
							const bool mA = j < gA.Q && t < gA.T, mB = hasB && j < gB.Q && t < gB.T;
							if (mA) dpb_flags(gA, t, r, cA, fA, lA);
							if (mB) dpb_flags(gB, t, r, cB, fB, lB);
							uint32_t z0 = dpf_prmt(LA, LB, SEL[c]);
							{ // fresh inside the fill range (which also refreshes the column's stored score), else the stored one
								const uint32_t mf = dpb_mask(mA && fA, mB && fB);
								z0 = dpb_blend(mf, z0, SZ[c]);
								SZ[c] = z0;
							}
							// left inputs: the boundary for column 0, the neighbour if it was computed on the previous diagonal, else upstream's initial values
							uint32_t xl = XL, vl = VL, x2l = X2L;
							if (t == 0) { xl = X_INIT, x2l = X2_INIT; vl = dpf_pack2(8 * (rowbnd(j) + B)); }
							else { const uint32_t ml = dpb_mask(pcA, pcB); xl = dpb_blend(ml, XL, X_INIT), vl = dpb_blend(ml, VL, V_INIT), x2l = dpb_blend(ml, X2L, X2_INIT); }
							const uint32_t a = __vadd2(xl, vl), b = __vadd2(Y[c], U[c]), a2 = __vadd2(x2l, vl), b2 = __vadd2(Y2[c], U[c]);
							const uint32_t zt = __vimax3_s16x2(__vimax3_s16x2(z0, a, b), a2, b2);
							const uint32_t zc = zt & 0xfff8fff8u;
							const uint32_t zl = __vmins2(zc, MCHC);
							const uint32_t un = __vsub2(zl, vl), vn = __vsub2(zl, U[c]);
							const uint32_t nz1 = __vsub2(KE1, zl), nz2 = __vsub2(KE2, zl);
							const uint32_t r1 = __viaddmax_s16x2(a, nz1, FL1), r2 = __viaddmax_s16x2(b, nz1, FL2);
							const uint32_t r3 = __viaddmax_s16x2(a2, nz2, FL3), r4 = __viaddmax_s16x2(b2, nz2, FL4);
							const uint32_t g1 = __viaddmin_s16x2(r1, NFL1, EIGHT), g2 = __viaddmin_s16x2(r2, NFL2, EIGHT);
							const uint32_t g3 = __viaddmin_s16x2(r3, NFL3, EIGHT), g4 = __viaddmin_s16x2(r4, NFL4, EIGHT);
							const uint32_t wd = (zt - zc) + g1 + 2u * g2 + 4u * g3 + 8u * g4;
							const uint32_t mc = dpb_mask(cA, cB);
							XL = r1, X2L = r3, VL = vn;                       // consumed by the next column only if this one was computed (pcA / pcB)
							Y[c] = dpb_blend(mc, r2, Y[c]), Y2[c] = dpb_blend(mc, r4, Y2[c]), U[c] = dpb_blend(mc, un, U[c]);
							pcA = cA, pcB = cB;
							if (c & 1) wv[c >> 1] = dpf_prmt(wprev, wd, 0x6240u);
							wprev = wd;
							if (t == j) { if (mA) accA += (int)(zl & 0xffffu); if (mB) accB += (int)(zl >> 16); }
							if (mA && j == gA.Q - 1 && t >= gA.Q) accA += (int)(int16_t)(un & 0xffffu);
							if (mB && j == gB.Q - 1 && t >= gB.Q) accB += (int)(int16_t)(un >> 16);
							if (mA && t == gA.T - 1 && j >= gA.T) accA += (int)(int16_t)(vn & 0xffffu);
							if (mB && t == gB.T - 1 && j >= gB.T) accB += (int)(int16_t)(vn >> 16);
						}
					}
					XLo = XL, VLo = VL, X2Lo = X2L;
					*reinterpret_cast<uint4*>(dst) = make_uint4(wv[0], wv[1], wv[2], wv[3]);
				}
				// the strip's last column, row by row, for the next strip (lane 31 evaluated row s - 31 in this step)
				if (strip + 1 < n_strips) {
					if (lane == 31 && j >= 0 && j < Qm) __stcg(reinterpret_cast<uint4*>(EDGE_OUT + 4 * j), make_uint4(XLo, VLo, X2Lo, 0u));
					if ((s & 31) == 31 || s == n_steps - 1) { // publish: rows <= s - 31 are out
						__threadfence_block();
						__syncwarp();
						if (lane == 31) s_prog[strip] = s == n_steps - 1 ? Qm : s - 30;
					}
				}
			}
		}
		// ---- combine the per-strip partial sums of the score path ----
		#pragma unroll
		for (int dlt = 16; dlt > 0; dlt >>= 1) { accA += __shfl_xor_sync(FULL, accA, dlt); accB += __shfl_xor_sync(FULL, accB, dlt); }
		if (lane == 0) s_acc[wid][0] = accA, s_acc[wid][1] = accB;
		__syncthreads();
		if (wid != 0) continue;     // the rest (cell count, end scores, traceback) is warp 0's
		accA = accB = 0;
		for (int k = 0; k < NW; ++k) accA += s_acc[k][0], accB += s_acc[k][1];
		{ // in-band cells, as the oracle counts them
			unsigned long long cc = 0;
			for (int r = lane; r < gA.Q + gA.T - 1; r += 32) { int st0, en0; dpb_range(gA, r, st0, en0); cc += en0 >= st0 ? (unsigned)(en0 - st0 + 1) : 0u; }
			if (hasB) for (int r = lane; r < gB.Q + gB.T - 1; r += 32) { int st0, en0; dpb_range(gB, r, st0, en0); cc += en0 >= st0 ? (unsigned)(en0 - st0 + 1) : 0u; }
			#pragma unroll
			for (int dlt = 16; dlt > 0; dlt >>= 1) cc += __shfl_xor_sync(FULL, cc, dlt);
			cells += cc;
		}
		// ---- end scores: z along the main diagonal (8*(z+2B) each), then u along the last row / v along the last column (8*(.+B) each) ----
		const int ndA = min(gA.Q, gA.T), ndB = min(gB.Q, gB.T);
		const int scoreA = (int)(accA >> 3) - 2 * B * ndA - B * (max(gA.Q, gA.T) - ndA);
		const int scoreB = (int)(accB >> 3) - 2 * B * ndB - B * (max(gB.Q, gB.T) - ndB);
		__syncwarp();
		// ---- backtrack from the corner (ksw_backtrack with the block ranges as off / off_end): lanes 0-15 task A, 16-31 task B ----
		{
			const int grp = lane >> 4, hl = lane & 15;
			const unsigned gmask = grp ? 0xffff0000u : 0x0000ffffu;
			if (grp == 0 || hasB) {
				DpTask &T = grp ? TB : TA;
				const DpbGeo g = grp ? gB : gA;
				uint32_t *cigar = cigar_pool + T.cigar_off;
				int n_cigar = 0, i = g.T - 1, jj = g.Q - 1, state = 0;
				uint32_t cur_op = 0; int cur_len = 0;
				const uint8_t *Pb = reinterpret_cast<const uint8_t*>(P) + (grp ? 2 : 0);
				auto push = [&](uint32_t op, int len) {
					if (cur_len > 0 && op != cur_op) { if (hl == 0) cigar[n_cigar] = (uint32_t)cur_len << 4 | cur_op; ++n_cigar; cur_len = 0; }
					cur_op = op, cur_len += len;
				};
				auto in_range = [&](int ci, int cj, int &st, int &en) { int st0, en0; dpb_range(g, ci + cj, st0, en0); st = st0 & ~15, en = ((en0 + 16) & ~15) - 1; return ci >= st && ci <= en; };
				while (i >= 0 && jj >= 0) {
					const int ci = i - hl, cj = jj - hl;
					uint32_t raw = 0xffu; // off the matrix or outside the computed range: not a match, stops a run
					int st_, en_;
					if (ci >= 0 && cj >= 0 && in_range(ci, cj, st_, en_)) {
						const int sp = ci / DPB_STRIP, tl = ci - sp * DPB_STRIP, L = tl / C, c = tl - L * C;
						raw = Pb[(((size_t)sp * n_steps + (size_t)(cj + L)) * 32 + L) * (CW * 4) + (size_t)(c >> 1) * 4 + (c & 1)];
					}
					int k = 0;
					if (state == 0) {
						const unsigned stop = (__ballot_sync(gmask, (raw & 7u) != 4u) >> (grp << 4)) & 0xffffu;
						k = stop ? __ffs(stop) - 1 : 16;
						if (k) { push(0, k); i -= k, jj -= k; }
					}
					if (k < 16 && i >= 0 && jj >= 0) { // one general step of ksw_backtrack on the next fetched cell
						const uint32_t rk = __shfl_sync(gmask, raw, (grp << 4) + k);
						int st, en, force = -1;
						in_range(i, jj, st, en);
						if (i < st) force = 2;
						if (i > en) force = 1;
						const uint32_t tmp = force < 0 ? ((4u - (rk & 7u)) | (rk & 0x78u)) : 0u;
						if (state == 0) state = tmp & 7;
						else if (!(tmp >> (state + 2) & 1)) state = 0;
						if (state == 0) state = tmp & 7;
						if (force >= 0) state = force;
						if (state == 0) { push(0, 1); --i, --jj; }
						else if (state == 1 || state == 3) { push(2, 1); --i; }
						else { push(1, 1); --jj; }
					}
				}
				if (i >= 0) push(2, i + 1);
				if (jj >= 0) push(1, jj + 1);
				if (cur_len > 0) { if (hl == 0) cigar[n_cigar] = (uint32_t)cur_len << 4 | cur_op; ++n_cigar; }
				__syncwarp(gmask);
				for (int k = hl; k < n_cigar >> 1; k += 16) { uint32_t t_ = cigar[k]; cigar[k] = cigar[n_cigar - 1 - k], cigar[n_cigar - 1 - k] = t_; }
				if (hl == 0) {
					T.max_q = T.max_t = T.mqe_t = -1;
					T.max = 0, T.mqe = MB_KSW_NEG_INF;
					T.score = grp ? scoreB : scoreA, T.zdropped = 0, T.reach_end = 0, T.n_cigar = n_cigar;
				}
			}
		}
		__syncwarp();
	}
	if (threadIdx.x == 0 && cells_out && cells) atomicAdd(cells_out, cells);
}
