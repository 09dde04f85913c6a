// dp_ext.cuh -- K4 fast path for END EXTENSIONS (ksw_extd2 with KSW_EZ_EXTZ_ONLY [| KSW_EZ_RIGHT | KSW_EZ_REV_CIGAR]).
//
// mm_align1 extends every region to the left (on reversed sequences, gaps right-aligned) and to the right with the exact
// variant of ksw_extd2: 32-bit H tracking, per-anti-diagonal maximum with the SSE kernel's lane order as tie-break,
// Z-drop, best score at the query end, traceback from the maximum.  When the band (w = 751) does not limit the matrix the
// cell values are plain two-piece affine DP, so the same register-resident packed systolic wavefront as k_dp_fast applies
// (two tasks per warp, 16x2 SIMD, see dp_fast.cuh).  On top of it:
//   * H(t,j) is carried along the row (H += u, packed, scaled by 8), lane to lane by shuffle,
//   * every cell folds key = (H + 32768) << 12 | (4095 - rank) into diag[t+j] with a shared-memory atomicMax, where rank
//     encodes upstream's order among equal scores on one anti-diagonal (H[en0] first, then the four interleaved SSE lanes
//     in order, then the scalar tail), so diag[r] ends up holding exactly upstream's (max_H, max_t) of diagonal r,
//   * H of the last query row is kept (mqe / score),
//   * after the wavefront one lane per task replays ksw_apply_zdrop over the diagonals in order: the break diagonal, ez.max,
//     max_t, max_q, mqe, mqe_t, score and zdropped come out exactly as if the DP had stopped there (cells beyond the break
//     are never consulted by the traceback, which starts on or before it),
//   * KSW_EZ_RIGHT flips the tie-break priority and makes the gap-continuation test non-strict: both are per-half constants
//     (tags, floors, deltas), so a left and a right extension can share a warp.
// Everything else (band-limited extensions, ambiguous bases, longer windows) stays with k_dp.
#pragma once
#include "dp_fast.cuh"

#define DPX_WARPS 2
#define DPX_MAX_LEN 752                 // max(qlen, tlen): w + 1 for w = 751
#define DPX_NDIAG (2 * DPX_MAX_LEN)
#define DPX_R 256                       // rank stride of one SSE lane ((t - st0) >> 2 < 256)

struct DpxRes { int max, max_t, max_q, mqe, mqe_t, score, zdropped, cells; };

// replay of the per-diagonal bookkeeping of ksw_extd2_sse's exact branch + ksw_apply_zdrop (ksw2_extd2_sse.c, ksw2.h)
MB_D DpxRes dpx_scan(const uint32_t *diag, const int16_t *hlast, int Q, int T, int zdrop, int e2)
{
	DpxRes z; z.max = 0, z.max_t = z.max_q = -1, z.mqe = MB_KSW_NEG_INF, z.mqe_t = -1, z.score = MB_KSW_NEG_INF, z.zdropped = 0, z.cells = 0;
	for (int r = 0; r < Q + T - 1; ++r) {
		const int st0 = r - Q + 1 > 0 ? r - Q + 1 : 0, en0 = r < T - 1 ? r : T - 1;
		z.cells += en0 - st0 + 1; // cells upstream visits: whole diagonals up to and including the break
		const uint32_t key = diag[r];
		const int H = (int)(key >> 12) - 32768, rank = 4095 - (int)(key & 4095u);
		int mt;
		if (rank == 0) mt = en0;
		else if (rank < 1 + 4 * DPX_R) { const int rr = rank - 1; mt = st0 + ((rr & (DPX_R - 1)) << 2) + (rr >> 8); }
		else mt = st0 + (((en0 - st0) >> 2) << 2) + (rank - 1 - 4 * DPX_R);
		if (r - st0 == Q - 1) { const int h = hlast[st0]; if (h > z.mqe) z.mqe = h, z.mqe_t = st0; }
		if (H > z.max) z.max = H, z.max_t = mt, z.max_q = r - mt;
		else if (mt >= z.max_t && r - mt >= z.max_q) {
			const int tl = mt - z.max_t, ql = (r - mt) - z.max_q, l = tl > ql ? tl - ql : ql - tl;
			if (zdrop >= 0 && z.max - H > zdrop + l * e2) { z.zdropped = 1; break; }
		}
		if (r == Q + T - 2 && en0 == T - 1) z.score = hlast[T - 1];
	}
	return z;
}

template <int C>
__global__ void __launch_bounds__(DPX_WARPS * 32)
k_dp_ext(DpTask *__restrict__ tasks, const int32_t *__restrict__ order, const int32_t *__restrict__ n_order, int32_t *__restrict__ work_ctr,
         const uint8_t *__restrict__ codes, const uint32_t *__restrict__ S, const uint8_t *__restrict__ pool,
         uint32_t *__restrict__ p_scr, size_t p_stride_words, uint32_t *__restrict__ cigar_pool, DpScoring sc, unsigned long long *__restrict__ cells_out)
{
	constexpr int CW = (C + 1) / 2;
	__shared__ uint32_t s_diag[DPX_WARPS][2][DPX_NDIAG];
	__shared__ int16_t s_hlast[DPX_WARPS][2][DPF_MAX_T];
	const unsigned FULL = 0xffffffffu;
	const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
	const int gw = blockIdx.x * DPX_WARPS + wib;
	uint32_t *P = p_scr + (size_t)gw * p_stride_words;
	uint32_t *diagA = s_diag[wib][0], *diagB = s_diag[wib][1];
	int16_t *hlA = s_hlast[wib][0], *hlB = s_hlast[wib][1];
	const int n_total = *n_order;
	int q = sc.q, e = sc.e, q2 = sc.q2, e2 = sc.e2;
	if (q2 + e2 < q + e) { int t_ = q; q = q2, q2 = t_; t_ = e; e = e2, e2 = t_; }
	int long_thres = e != e2 ? (q2 - q) / (e - e2) - 1 : 0;
	if (q2 + e2 + long_thres * e2 > q + e + long_thres * e) ++long_thres;
	const int long_diff = long_thres * (e - e2) - (q2 - q) - e2;
	const int B = q + e;
	const uint32_t K1 = 0x00010000u + dpf_pack2(8 * (B - e)), K2 = 0x00010000u + dpf_pack2(8 * (B - e2));
	const uint32_t EIGHT = 0x00080008u, NEG8B = dpf_pack2(-8 * B);
	auto pack_ab = [](int lo, int hi) { return (uint32_t)(uint16_t)(int16_t)lo | (uint32_t)(uint16_t)(int16_t)hi << 16; };
	for (;;) {
		int oi = 0;
		if (lane == 0) oi = atomicAdd(work_ctr, 1) * 2;
		oi = __shfl_sync(FULL, oi, 0);
		if (oi >= n_total) break;
		const bool hasB = oi + 1 < n_total;
		DpTask &TA = tasks[order[oi]];
		DpTask &TB = tasks[order[hasB ? oi + 1 : oi]];
		const int QA = TA.qlen, TLA = TA.tlen, QB = TB.qlen, TLB = TB.tlen;
		const int Qm = QA > QB ? QA : QB, TLm = TLA > TLB ? TLA : TLB;
		const int rtA = (TA.flag & MB_EZ_RIGHT) ? 1 : 0, rtB = (TB.flag & MB_EZ_RIGHT) ? 1 : 0;
		// per-half constants: priority tags (diag, E, F, E2, F2) = (4,3,2,1,0), reversed with KSW_EZ_RIGHT
		const int tzA = rtA ? 0 : 4, tzB = rtB ? 0 : 4;
		const int t1A = rtA ? 1 : 3, t1B = rtB ? 1 : 3, t2A = 2, t2B = 2, t3A = rtA ? 3 : 1, t3B = rtB ? 3 : 1, t4A = rtA ? 4 : 0, t4B = rtB ? 4 : 0;
		const int f1 = 8 * (-q - e + B), f2 = 8 * (-q2 - e2 + B);
		const uint32_t X_INIT = pack_ab(f1 + t1A, f1 + t1B), Y_INIT = pack_ab(f1 + t2A, f1 + t2B);
		const uint32_t X2_INIT = pack_ab(f2 + t3A, f2 + t3B), Y2_INIT = pack_ab(f2 + t4A, f2 + t4B);
		const uint32_t FL1 = X_INIT, FL2 = Y_INIT, FL3 = X2_INIT, FL4 = Y2_INIT;
		// flag deltas: continue iff s - floor (+8 with KSW_EZ_RIGHT: non-strict) >= 8
		const uint32_t DF1 = pack_ab(8 * rtA - (f1 + t1A), 8 * rtB - (f1 + t1B)), DF2 = pack_ab(8 * rtA - (f1 + t2A), 8 * rtB - (f1 + t2B));
		const uint32_t DF3 = pack_ab(8 * rtA - (f2 + t3A), 8 * rtB - (f2 + t3B)), DF4 = pack_ab(8 * rtA - (f2 + t4A), 8 * rtB - (f2 + t4B));
		const uint32_t mchA = (uint32_t)(8 * (sc.sc_mch + 2 * B) + tzA), misA = (uint32_t)(8 * (sc.sc_mis + 2 * B) + tzA), nA = (uint32_t)(8 * (sc.sc_N + 2 * B) + tzA);
		const uint32_t mchB = (uint32_t)(8 * (sc.sc_mch + 2 * B) + tzB), misB = (uint32_t)(8 * (sc.sc_mis + 2 * B) + tzB), nB = (uint32_t)(8 * (sc.sc_N + 2 * B) + tzB);
		QView qvA; qvA.codes = TA.q_comp == 2 ? pool : codes; qvA.idx0 = TA.q_idx0; qvA.step = TA.q_step; qvA.comp = TA.q_comp == 1;
		QView qvB; qvB.codes = TB.q_comp == 2 ? pool : codes; qvB.idx0 = TB.q_idx0; qvB.step = TB.q_step; qvB.comp = TB.q_comp == 1;
		TView tvA; tvA.S = S; tvA.bytes = pool; tvA.idx0 = TA.t_idx0; tvA.step = TA.t_step; tvA.packed = TA.t_packed;
		TView tvB; tvB.S = S; tvB.bytes = pool; tvB.idx0 = TB.t_idx0; tvB.step = TB.t_step; tvB.packed = TB.t_packed;
		const int t0 = lane * C;
		uint32_t SEL[C], U[C], Y[C], Y2[C];
		#pragma unroll
		for (int c = 0; c < C; ++c) {
			const int t = t0 + c;
			const uint32_t a = t < TLA ? (uint32_t)tvA.at(t) & 3u : 0u, b = t < TLB ? (uint32_t)tvB.at(t) & 3u : 0u;
			SEL[c] = a | (8u | a) << 4 | (4u + b) << 8 | (12u + b) << 12;
			U[c] = dpf_pack2(8 * (dpf_bnd(t, q, e, e2, long_thres, long_diff) + B));
			Y[c] = Y_INIT, Y2[c] = Y2_INIT;
		}
		for (int r = lane; r < QA + TLA - 1; r += 32) diagA[r] = 0;
		for (int r = lane; r < QB + TLB - 1; r += 32) diagB[r] = 0;
		__syncwarp();
		uint32_t XLo = 0, VLo = 0, X2Lo = 0, HLo = 0;
		const int n_steps = Qm + 31;
		const bool lane_live = t0 < TLm;
		auto row_tables = [&](int r, uint32_t &la, uint32_t &lb) {
			const int a = r < QA ? qvA.at(r) : 0, b = r < QB ? qvB.at(r) : 0;
			la = a < 4 ? misA * 0x01010101u + ((mchA - misA) << (a * 8)) : nA * 0x01010101u;
			lb = b < 4 ? misB * 0x01010101u + ((mchB - misB) << (b * 8)) : nB * 0x01010101u;
		};
		uint32_t LAc, LBc, LAn = 0, LBn = 0, LAo = 0, LBo = 0;
		row_tables(lane, LAc, LBc);
		int hb = 0; // H(-1, j): first-column boundary, running sum of bnd()
		uint32_t *dst = P + (size_t)lane * CW;
		for (int s = 0; s < n_steps; ++s, dst += 32 * CW) {
			const int j = s - lane;
			if ((s & 31) == 0) {
				if (s) LAc = LAn, LBc = LBn;
				if (s + 32 < Qm) row_tables(s + 32 + lane, LAn, LBn);
			}
			const uint32_t LA0 = __shfl_sync(FULL, LAc, s & 31), LB0 = __shfl_sync(FULL, LBc, s & 31);
			uint32_t LA = __shfl_up_sync(FULL, LAo, 1), LB = __shfl_up_sync(FULL, LBo, 1);
			uint32_t XL = __shfl_up_sync(FULL, XLo, 1), VL = __shfl_up_sync(FULL, VLo, 1), X2L = __shfl_up_sync(FULL, X2Lo, 1);
			uint32_t H8 = __shfl_up_sync(FULL, HLo, 1);
			if (lane == 0) {
				const int bv = dpf_bnd(s, q, e, e2, long_thres, long_diff);
				hb += bv;
				LA = LA0, LB = LB0, XL = X_INIT, X2L = X2_INIT, VL = dpf_pack2(8 * (bv + B)), H8 = dpf_pack2(8 * hb);
			}
			LAo = LA, LBo = LB;
			if (lane_live && j >= 0 && j < Qm) {
				uint32_t wv[CW];
				uint32_t wprev = 0;
				#pragma unroll
				for (int c = 0; c < C; ++c) {
					const uint32_t z0 = dpf_prmt(LA, LB, SEL[c]);
					const uint32_t a = __vadd2(XL, VL), b = __vadd2(Y[c], U[c]), a2 = __vadd2(X2L, VL), b2 = __vadd2(Y2[c], U[c]);
					const uint32_t zt = __vimax3_s16x2(__vimax3_s16x2(z0, a, b), a2, b2);
					const uint32_t zc = zt & 0xfff8fff8u;
					const uint32_t un = zc - VL, vn = zc - U[c];
					const uint32_t nz1 = K1 - zc, nz2 = K2 - zc;
					const uint32_t s1 = __vadd2(a, nz1), s2 = __vadd2(b, nz1), s3 = __vadd2(a2, nz2), s4 = __vadd2(b2, nz2);
					const uint32_t r1 = __vmaxs2(s1, FL1), r2 = __vmaxs2(s2, FL2), r3 = __vmaxs2(s3, FL3), r4 = __vmaxs2(s4, FL4);
					const uint32_t g1 = __vmins2(__viaddmax_s16x2(s1, DF1, 0u), EIGHT), g2 = __vmins2(__viaddmax_s16x2(s2, DF2, 0u), EIGHT);
					const uint32_t g3 = __vmins2(__viaddmax_s16x2(s3, DF3, 0u), EIGHT), g4 = __vmins2(__viaddmax_s16x2(s4, DF4, 0u), EIGHT);
					const uint32_t wd = (zt - zc) + g1 + 2u * g2 + 4u * g3 + 8u * g4;
					XL = r1, X2L = r3, Y[c] = r2, Y2[c] = r4, U[c] = un, VL = vn;
					if (c & 1) wv[c >> 1] = dpf_prmt(wprev, wd, 0x6240u);
					else if (c == C - 1) wv[c >> 1] = dpf_prmt(wd, 0u, 0x6240u);
					wprev = wd;
					// H(t, j) = H(t-1, j) + u(t, j); fold it into the maximum of its anti-diagonal
					H8 = __vadd2(__vadd2(H8, un), NEG8B);
					const int t = t0 + c, r = t + j;
					if (t < TLA && j < QA) {
						const int H = (int)(int16_t)(H8 & 0xffffu) >> 3;
						const int st0 = r - QA + 1 > 0 ? r - QA + 1 : 0, en0 = r < TLA - 1 ? r : TLA - 1, o = t - st0, e1 = ((en0 - st0) >> 2) << 2;
						const int rank = t == en0 ? 0 : o < e1 ? 1 + ((o & 3) << 8) + (o >> 2) : 1 + 4 * DPX_R + (o - e1);
						atomicMax(&diagA[r], (uint32_t)(H + 32768) << 12 | (uint32_t)(4095 - rank));
						if (j == QA - 1) hlA[t] = (int16_t)H;
					}
					if (t < TLB && j < QB) {
						const int H = (int)(int16_t)(H8 >> 16) >> 3;
						const int st0 = r - QB + 1 > 0 ? r - QB + 1 : 0, en0 = r < TLB - 1 ? r : TLB - 1, o = t - st0, e1 = ((en0 - st0) >> 2) << 2;
						const int rank = t == en0 ? 0 : o < e1 ? 1 + ((o & 3) << 8) + (o >> 2) : 1 + 4 * DPX_R + (o - e1);
						atomicMax(&diagB[r], (uint32_t)(H + 32768) << 12 | (uint32_t)(4095 - rank));
						if (j == QB - 1) hlB[t] = (int16_t)H;
					}
				}
				XLo = XL, VLo = VL, X2Lo = X2L, HLo = H8;
				if (CW % 4 == 0) {
					#pragma unroll
					for (int k = 0; k < CW / 4; ++k) reinterpret_cast<uint4*>(dst)[k] = make_uint4(wv[4 * k], wv[4 * k + 1], wv[4 * k + 2], wv[4 * k + 3]);
				} else if (CW % 2 == 0) {
					#pragma unroll
					for (int k = 0; k < CW / 2; ++k) reinterpret_cast<uint2*>(dst)[k] = make_uint2(wv[2 * k], wv[2 * k + 1]);
				} else {
					#pragma unroll
					for (int k = 0; k < CW; ++k) dst[k] = wv[k];
				}
			}
		}
		__syncwarp();
		// ---- per-diagonal replay (lane 0: task A, lane 16: task B), then the traceback by half-warps ----
		{
			const int grp = lane >> 4, hl = lane & 15;
			const unsigned gmask = grp ? 0xffff0000u : 0x0000ffffu;
			if (grp == 0 || hasB) {
				DpTask &T = grp ? TB : TA;
				const int TL = grp ? TLB : TLA, Q = grp ? QB : QA, right = grp ? rtB : rtA;
				DpxRes z; z.max = 0, z.max_t = z.max_q = -1, z.mqe = MB_KSW_NEG_INF, z.mqe_t = -1, z.score = MB_KSW_NEG_INF, z.zdropped = 0, z.cells = 0;
				if (hl == 0) z = dpx_scan(grp ? diagB : diagA, grp ? hlB : hlA, Q, TL, T.zdrop, e2);
				const int src = grp << 4;
				z.max = __shfl_sync(gmask, z.max, src), z.max_t = __shfl_sync(gmask, z.max_t, src), z.max_q = __shfl_sync(gmask, z.max_q, src);
				z.mqe = __shfl_sync(gmask, z.mqe, src), z.mqe_t = __shfl_sync(gmask, z.mqe_t, src), z.zdropped = __shfl_sync(gmask, z.zdropped, src);
				int reach_end = 0, i = -1, jj = -1;
				if (!z.zdropped && z.mqe + T.end_bonus > z.max) reach_end = 1, i = z.mqe_t, jj = Q - 1;
				else if (z.max_t >= 0 && z.max_q >= 0) i = z.max_t, jj = z.max_q;
				uint32_t *cigar = cigar_pool + T.cigar_off;
				int n_cigar = 0, state = 0;
				uint32_t cur_op = 0; int cur_len = 0;
				const uint8_t *Pb = reinterpret_cast<const uint8_t*>(P) + (grp ? 2 : 0); // plain loads: the __syncwarp() above orders them after the forward pass
				auto push = [&](uint32_t op, int len) {
					if (cur_len > 0 && op != cur_op) { if (hl == 0) cigar[n_cigar] = (uint32_t)cur_len << 4 | cur_op; ++n_cigar; cur_len = 0; }
					cur_op = op, cur_len += len;
				};
				const bool any = i >= 0 && jj >= 0;
				while (i >= 0 && jj >= 0) {
					const int ci = i - hl, cj = jj - hl;
					uint32_t raw = 0xffu; // off the matrix: not a match, stops a run
					if (ci >= 0 && cj >= 0) {
						const int L = ci / C, c = ci - L * C;
						raw = Pb[((((size_t)(cj + L)) * 32 + L) * CW + (c >> 1)) * 4 + (c & 1)];
					}
					int k = 0;
					if (state == 0) { // a run of diagonal moves is consumed at once: leading cells whose direction is "match"
						const unsigned stop = (__ballot_sync(gmask, (raw & 7u) != (right ? 0u : 4u)) >> (grp << 4)) & 0xffffu;
						k = stop ? __ffs(stop) - 1 : 16;
						if (k) { push(0, k); i -= k, jj -= k; }
					}
					if (k < 16 && i >= 0 && jj >= 0) { // one general step of ksw_backtrack on the next fetched cell
						const uint32_t rk = __shfl_sync(gmask, raw, (grp << 4) + k);
						const uint32_t tmp = (right ? (rk & 7u) : 4u - (rk & 7u)) | (rk & 0x78u);
						if (state == 0) state = tmp & 7;
						else if (!(tmp >> (state + 2) & 1)) state = 0;
						if (state == 0) state = tmp & 7;
						if (state == 0) { push(0, 1); --i, --jj; }
						else if (state == 1 || state == 3) { push(2, 1); --i; }
						else { push(1, 1); --jj; }
					}
				}
				if (any) {
					if (i >= 0) push(2, i + 1);
					if (jj >= 0) push(1, jj + 1);
				}
				if (cur_len > 0) { if (hl == 0) cigar[n_cigar] = (uint32_t)cur_len << 4 | cur_op; ++n_cigar; }
				__syncwarp(gmask);
				if (!(T.flag & MB_EZ_REV_CIGAR))
					for (int k = hl; k < n_cigar >> 1; k += 16) { uint32_t t_ = cigar[k]; cigar[k] = cigar[n_cigar - 1 - k], cigar[n_cigar - 1 - k] = t_; }
				if (hl == 0) {
					if (cells_out) atomicAdd(cells_out, (unsigned long long)z.cells);
					T.score = z.score, T.max = z.max, T.max_q = z.max_q, T.max_t = z.max_t, T.mqe = z.mqe, T.mqe_t = z.mqe_t;
					T.zdropped = z.zdropped, T.reach_end = reach_end, T.n_cigar = n_cigar;
				}
			}
		}
		__syncwarp();
	}
}

// eligibility of an extension task for k_dp_ext: class index by columns per lane, or -1
static __host__ __device__ inline int dpx_class(int qlen, int tlen, int w, int flag, int skip)
{
	if (skip || qlen <= 0 || tlen <= 0) return -1;
	if (flag != MB_EZ_EXTZ_ONLY && flag != (MB_EZ_EXTZ_ONLY | MB_EZ_RIGHT | MB_EZ_REV_CIGAR)) return -1;
	if (w < 0) w = tlen > qlen ? tlen : qlen;
	const int mx = tlen > qlen ? tlen : qlen;
	if (mx > w + 1 || mx > DPX_MAX_LEN) return -1; // the band must never limit the matrix
	const int c = (tlen + 31) / 32;
	const int cls[DPF_NCLASS] = { 4, 5, 6, 7, 8, 9, 10, 11, 12, 14, 16, 20, 24 };
	for (int k = 0; k < DPF_NCLASS; ++k) if (c <= cls[k]) return k;
	return -1;
}
