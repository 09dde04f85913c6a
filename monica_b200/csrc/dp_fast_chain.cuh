// dp_fast_chain.cuh -- k_dp_fast with the task pairs of a warp CHAINED through the lanes.
//
// k_dp_fast (dp_fast.cuh) runs one pair of gap-fill tasks per pass of the systolic wavefront: lane L works on query row s - L,
// so a pass over Q rows takes Q + 31 steps and the warp issues every instruction of the 31 extra steps for lanes that have
// not started or have already finished.  At the typical window of a map-ont gap fill (Q ~ 200 rows) that is 13 % of the
// kernel.  Here a warp takes a GROUP of up to DPF_GMAX consecutive pairs of its class list and feeds their query rows into
// lane 0 back to back: when lane L has finished the last row of pair g it starts row 0 of pair g + 1 at the next step, while
// the lanes behind it are still on pair g.  The ramp is paid once per group instead of once per pair.
//
//   * what differs per pair lives per LANE now (each lane is in its own pair for up to 31 steps): the row counter, the rows
//     left, the column masks for the end-score capture and the column state.  A lane's switch to the next pair is a handful
//     of register moves: the target selectors of all pairs of the group are staged in shared memory when the group starts
//     (sSEL[g][c][lane], conflict-free), u / y / y2 restart from their constants,
//   * the query rows reach lane 0 as one stream (the rows of pair 0, then of pair 1, ...); the 32-row table loads map a
//     stream position to (pair, row) with the prefix sums of the pairs' row counts,
//   * every pair keeps its own [step][lane][C/2] region of direction bytes, the regions back to back, so a lane's store
//     pointer advances by one step per step and by 31 more at a switch; the backtracks of the group run after its forward
//     pass with the code of k_dp_fast,
//   * each lane leaves its share of the end-score sums (sum of u over the last row) in shared memory at its switch.
// The cell arithmetic is k_dp_fast's, statement for statement; results are identical by construction and by test
// (tests/test_gpu_parity.py::test_dp_fast_path_pairs runs both kernels).
#pragma once
#include "dp_fast.cuh"

#define DPF_GMAX 4

template <int C>
__global__ void __launch_bounds__(32, (C <= 8 ? 20 : C <= 12 ? 16 : 4))
k_dp_fast_chain(DpTask *__restrict__ tasks, const int32_t *__restrict__ order, const int32_t *__restrict__ n_order, int32_t *__restrict__ work_ctr,
                const uint8_t *__restrict__ codes, const uint32_t *__restrict__ S, const uint8_t *__restrict__ pool,
                uint32_t *__restrict__ p_scr, size_t p_stride_words, uint32_t *__restrict__ cigar_pool, DpScoring sc, unsigned long long *__restrict__ cells_out, int G)
{
	constexpr int CW = (C + 1) / 2;
	__shared__ uint32_t sSEL[DPF_GMAX][C][32];
	__shared__ int sQA[DPF_GMAX], sQB[DPF_GMAX], sTLA[DPF_GMAX], sTLB[DPF_GMAX], sQm[DPF_GMAX], sBase[DPF_GMAX + 1];
	__shared__ int4 sMeta[DPF_GMAX];                     // what a lane needs when it enters pair g: rows, last row of A, last row of B
	__shared__ uint32_t sCMA[DPF_GMAX][32], sCMB[DPF_GMAX][32];   // ... and which of its columns belong to task A / B
	__shared__ int sSum[DPF_GMAX][2][32];                // each lane's share of the end-score sums of pair g, written when it leaves the pair
	__shared__ int sIdA[DPF_GMAX], sIdB[DPF_GMAX];
	__shared__ QView sqA[DPF_GMAX], sqB[DPF_GMAX];
	const unsigned FULL = 0xffffffffu;
	const int lane = threadIdx.x & 31;
	uint32_t *P = p_scr + (size_t)blockIdx.x * p_stride_words;
	unsigned long long cells = 0;
	const int n_total = *n_order;
	int q = sc.q, e = sc.e, q2 = sc.q2, e2 = sc.e2;
	if (q2 + e2 < q + e) { int t_ = q; q = q2, q2 = t_; t_ = e; e = e2, e2 = t_; }
	int long_thres = e != e2 ? (q2 - q) / (e - e2) - 1 : 0;
	if (q2 + e2 + long_thres * e2 > q + e + long_thres * e) ++long_thres;
	const int long_diff = long_thres * (e - e2) - (q2 - q) - e2;
	const int B = q + e;
	// packed constants (dp_fast.cuh)
	const uint32_t X_INIT = dpf_pack2(8 * (-q - e + B) + 3), Y_INIT = dpf_pack2(8 * (-q - e + B) + 2);
	const uint32_t X2_INIT = dpf_pack2(8 * (-q2 - e2 + B) + 1), Y2_INIT = dpf_pack2(8 * (-q2 - e2 + B) + 0);
	const uint32_t FL1 = X_INIT, FL2 = Y_INIT, FL3 = X2_INIT, FL4 = Y2_INIT;
	const uint32_t NFL1 = dpf_pack2(-(8 * (-q - e + B) + 3)), NFL2 = dpf_pack2(-(8 * (-q - e + B) + 2));
	const uint32_t NFL3 = dpf_pack2(-(8 * (-q2 - e2 + B) + 1)), NFL4 = dpf_pack2(-(8 * (-q2 - e2 + B) + 0));
	const uint32_t K1 = 0x00010000u + dpf_pack2(8 * (B - e)), K2 = 0x00010000u + dpf_pack2(8 * (B - e2));
	const uint32_t NEG1 = 0xffffffffu + (uint32_t)sc.pad;
	const uint32_t EIGHT = dpf_pack2(8 + sc.pad);
	const uint32_t MCHB = (uint32_t)(8 * (sc.sc_mch + 2 * B) + 4), MISB = (uint32_t)(8 * (sc.sc_mis + 2 * B) + 4), NB = (uint32_t)(8 * (sc.sc_N + 2 * B) + 4);
	const uint32_t MIS4 = MISB * 0x01010101u, N4 = NB * 0x01010101u, MDIFF = MCHB - MISB;
	const uint32_t VB0 = dpf_pack2(8 * (dpf_bnd(0, q, e, e2, long_thres, long_diff) + B)), VB1 = dpf_pack2(8 * (-e + B));
	const uint32_t VB2 = dpf_pack2(8 * (long_diff + B)), VB3 = dpf_pack2(8 * (-e2 + B));
	const int t0 = lane * C;
	__shared__ uint32_t sU0[C][32];   // u of row -1 (the first row's upper neighbour) for this lane's columns: the same for every pair
	#pragma unroll
	for (int c = 0; c < C; ++c) sU0[c][lane] = dpf_pack2(8 * (dpf_bnd(t0 + c, q, e, e2, long_thres, long_diff) + B));
	if (G < 1) G = 1;
	if (G > DPF_GMAX) G = DPF_GMAX;
	for (;;) {
		int gi = 0;
		if (lane == 0) gi = atomicAdd(work_ctr, 1);
		gi = __shfl_sync(FULL, gi, 0);
		const long long oi0 = (long long)gi * 2 * G;
		if (oi0 >= n_total) break;
		const int n_g = (int)(((long long)n_total - oi0 + 1) / 2 < G ? ((long long)n_total - oi0 + 1) / 2 : G);
		// ---- the group's pairs: one lane fills one slot ----
		if (lane < DPF_GMAX) {
			if (lane < n_g) {
				const int oi = (int)oi0 + 2 * lane;
				const int ia = order[oi], ib = order[oi + 1 < n_total ? oi + 1 : oi];
				const DpTask &TA = tasks[ia], &TB = tasks[ib];
				sIdA[lane] = ia, sIdB[lane] = oi + 1 < n_total ? ib : -1;
				sQA[lane] = TA.qlen, sQB[lane] = TB.qlen, sTLA[lane] = TA.tlen, sTLB[lane] = TB.tlen;
				sQm[lane] = TA.qlen > TB.qlen ? TA.qlen : TB.qlen;
				QView va; va.codes = TA.q_comp == 2 ? pool : codes; va.idx0 = TA.q_idx0; va.step = TA.q_step; va.comp = TA.q_comp == 1;
				QView vb; vb.codes = TB.q_comp == 2 ? pool : codes; vb.idx0 = TB.q_idx0; vb.step = TB.q_step; vb.comp = TB.q_comp == 1;
				sqA[lane] = va, sqB[lane] = vb;
				sMeta[lane] = make_int4(sQm[lane], TA.qlen - 1, TB.qlen - 1, 0);
			} else sQm[lane] = 0;
		}
		__syncwarp();
		if (lane == 0) { // where each pair's rows start in the stream that feeds lane 0
			int b = 0;
			for (int g = 0; g <= DPF_GMAX; ++g) { sBase[g] = g <= n_g ? b : 0x3fffffff; if (g < DPF_GMAX) b += sQm[g]; }
		}
		__syncwarp();
		const int total_rows = sBase[n_g];
		// target selectors of every pair of the group, this lane's columns
		for (int g = 0; g < n_g; ++g) {
			const DpTask &TA = tasks[sIdA[g]], &TB = tasks[sIdB[g] >= 0 ? sIdB[g] : sIdA[g]];
			TView tvA; tvA.S = S; tvA.bytes = pool; tvA.idx0 = TA.t_idx0; tvA.step = TA.t_step; tvA.packed = TA.t_packed;
			TView tvB; tvB.S = S; tvB.bytes = pool; tvB.idx0 = TB.t_idx0; tvB.step = TB.t_step; tvB.packed = TB.t_packed;
			const int TLA = sTLA[g], TLB = sTLB[g];
			#pragma unroll
			for (int c = 0; c < C; ++c) {
				const int t = t0 + c;
				const uint32_t a = t < TLA ? (uint32_t)tvA.at(t) & 3u : 0u, b = t < TLB ? (uint32_t)tvB.at(t) & 3u : 0u;
				sSEL[g][c][lane] = a | (8u | a) << 4 | (4u + b) << 8 | (12u + b) << 12;
			}
			const int na = TLA - t0, nb = TLB - t0;
			sCMA[g][lane] = na <= 0 ? 0u : na >= C ? (1u << C) - 1u : (1u << na) - 1u;
			sCMB[g][lane] = nb <= 0 ? 0u : nb >= C ? (1u << C) - 1u : (1u << nb) - 1u;
		}
		// (a lane reads back only what it wrote itself: no barrier needed for sSEL)
		uint32_t SEL[C], U[C], Y[C], Y2[C];
		int g_cur = 0, j = -lane, rows, qa1, qb1;
		uint32_t cmA, cmB;   // bit c: column t0 + c belongs to task A / B of the current pair
		bool live;
		auto enter_pair = [&](int g) {
			const int4 m = sMeta[g];
			rows = m.x, qa1 = m.y, qb1 = m.z;
			cmA = sCMA[g][lane], cmB = sCMB[g][lane];
			live = (cmA | cmB) != 0u;
			#pragma unroll
			for (int c = 0; c < C; ++c) {
				SEL[c] = sSEL[g][c][lane];
				U[c] = sU0[c][lane];
				Y[c] = Y_INIT, Y2[c] = Y2_INIT;
			}
		};
		enter_pair(0);
		int sumA = 0, sumB = 0;
		uint32_t XLo = 0, VLo = 0, X2Lo = 0;
		// substitution tables of the query rows, as a stream over the pairs of the group
		auto row_tables = [&](int p, uint32_t &la, uint32_t &lb) {
			int g = 0;
			#pragma unroll
			for (int i = 1; i < DPF_GMAX; ++i) g += p >= sBase[i] ? 1 : 0;
			const int r = p - sBase[g];
			const bool in = p < total_rows;   // (slots beyond the group's pairs hold stale views)
			const int a = in && r < sQA[g] ? sqA[g].at(r) : 0, b = in && r < sQB[g] ? sqB[g].at(r) : 0;
			la = a < 4 ? MIS4 + (MDIFF << (a * 8)) : N4, lb = b < 4 ? MIS4 + (MDIFF << (b * 8)) : N4;
		};
		uint32_t LAc, LBc, LAn = 0, LBn = 0, LAo = 0, LBo = 0;
		row_tables(lane, LAc, LBc);
		uint32_t *dst = P + (size_t)lane * CW;
		const int n_steps = total_rows + 31;
		#pragma unroll 1
		for (int s = 0; s < n_steps; ++s, dst += 32 * CW) {
			if ((s & 31) == 0) {
				if (s) LAc = LAn, LBc = LBn;
				if (s + 32 < total_rows) row_tables(s + 32 + lane, LAn, LBn);
			}
			const uint32_t LA0 = __shfl_sync(FULL, LAc, s & 31), LB0 = __shfl_sync(FULL, LBc, s & 31);
			uint32_t LA = __shfl_up_sync(FULL, LAo, 1), LB = __shfl_up_sync(FULL, LBo, 1);
			uint32_t XL = __shfl_up_sync(FULL, XLo, 1), VL = __shfl_up_sync(FULL, VLo, 1), X2L = __shfl_up_sync(FULL, X2Lo, 1);
			if (lane == 0) {
				LA = LA0, LB = LB0, XL = X_INIT, X2L = X2_INIT;
				VL = j == 0 ? VB0 : j < long_thres ? VB1 : j == long_thres ? VB2 : VB3;   // lane 0's row within its pair
			}
			LAo = LA, LBo = LB;
			if (live && j >= 0) {
				uint32_t wv[CW];
				uint32_t wprev = 0;
				#pragma unroll
				for (int c = 0; c < C; ++c) {
					const uint32_t z0 = dpf_prmt(LA, LB, SEL[c]);
					const uint32_t a = __vadd2(XL, VL), b = __vadd2(Y[c], U[c]), a2 = __vadd2(X2L, VL), b2 = __vadd2(Y2[c], U[c]);
					const uint32_t zt = __vimax3_s16x2(__vimax3_s16x2(z0, a, b), a2, b2);
					const uint32_t zc = dpf_and(zt, 0xfff8fff8u);
					const uint32_t un = zc - VL, vn = zc - U[c];
					const uint32_t nz1 = dpf_mad(zc, NEG1, K1), nz2 = dpf_mad(zc, NEG1, K2);
					const uint32_t r1 = __viaddmax_s16x2(a, nz1, FL1), r2 = __viaddmax_s16x2(b, nz1, FL2);
					const uint32_t r3 = __viaddmax_s16x2(a2, nz2, FL3), r4 = __viaddmax_s16x2(b2, nz2, FL4);
					const uint32_t g1 = __viaddmin_s16x2(r1, NFL1, EIGHT), g2 = __viaddmin_s16x2(r2, NFL2, EIGHT);
					const uint32_t g3 = __viaddmin_s16x2(r3, NFL3, EIGHT), g4 = __viaddmin_s16x2(r4, NFL4, EIGHT);
					const uint32_t wd = dpf_mad(g4, 8u, dpf_mad(g3, 4u, dpf_mad(g2, 2u, dpf_mad(zc, NEG1, zt) + g1)));
					XL = r1, X2L = r3, Y[c] = r2, Y2[c] = r4, U[c] = un, VL = vn;
					if (c & 1) wv[c >> 1] = dpf_prmt(wprev, wd, 0x6240u);
					else if (c == C - 1) wv[c >> 1] = dpf_prmt(wd, 0u, 0x6240u);
					wprev = wd;
				}
				XLo = XL, VLo = VL, X2Lo = X2L;
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
				if (j == qa1 || j == qb1) { // this lane's last row of task A / B: its columns' u enter the end score
					if (j == qa1) {
						#pragma unroll
						for (int c = 0; c < C; ++c) if (cmA >> c & 1u) sumA += (int)(U[c] & 0xffffu);
					}
					if (j == qb1) {
						#pragma unroll
						for (int c = 0; c < C; ++c) if (cmB >> c & 1u) sumB += (int)(U[c] >> 16);
					}
				}
			}
			if (++j == rows) { // this lane moves on to the next pair of the group
				sSum[g_cur][0][lane] = sumA, sSum[g_cur][1][lane] = sumB;
				sumA = sumB = 0;
				++g_cur, j = 0;
				if (g_cur < n_g) enter_pair(g_cur);
				else rows = 0x7fffffff, live = false;
				dst += 31 * 32 * CW;   // the next pair's region starts 31 steps further on (regions are (rows + 31) steps long)
			}
		}
		__syncwarp();
		// ---- end scores and backtracks of the group's pairs (k_dp_fast's code per pair) ----
		for (int g = 0; g < n_g; ++g) {
			const bool hasB = sIdB[g] >= 0;
			DpTask &TA = tasks[sIdA[g]];
			DpTask &TB = tasks[hasB ? sIdB[g] : sIdA[g]];
			const int QA = sQA[g], QB = sQB[g], TLA = sTLA[g], TLB = sTLB[g], Qm = sQm[g];
			cells += (unsigned long long)TLA * (unsigned)QA + (hasB ? (unsigned long long)TLB * (unsigned)QB : 0ULL);
			int bsA = 0, bsB = 0, suA = sSum[g][0][lane], suB = sSum[g][1][lane];
			for (int r = lane; r < Qm; r += 32) {
				const int bv = dpf_bnd(r, q, e, e2, long_thres, long_diff);
				if (r < QA) bsA += bv;
				if (r < QB) bsB += bv;
			}
			#pragma unroll
			for (int dlt = 16; dlt > 0; dlt >>= 1) {
				bsA += __shfl_xor_sync(FULL, bsA, dlt), bsB += __shfl_xor_sync(FULL, bsB, dlt);
				suA += __shfl_xor_sync(FULL, suA, dlt), suB += __shfl_xor_sync(FULL, suB, dlt);
			}
			const int scoreA = bsA + (suA >> 3) - B * TLA, scoreB = bsB + (suB >> 3) - B * TLB;
			const uint32_t *Pg = P + ((size_t)sBase[g] + (size_t)31 * g) * 32 * CW;
			const int grp = lane >> 4, hl = lane & 15;
			const unsigned gmask = grp ? 0xffff0000u : 0x0000ffffu;
			if (grp == 0 || hasB) {
				DpTask &T = grp ? TB : TA;
				const int TL = grp ? TLB : TLA, Q = grp ? QB : QA;
				uint32_t *cigar = cigar_pool + T.cigar_off;
				int n_cigar = 0, i = TL - 1, jj = Q - 1, state = 0;
				uint32_t cur_op = 0; int cur_len = 0;
				const uint8_t *Pb = reinterpret_cast<const uint8_t*>(Pg) + (grp ? 2 : 0);
				auto push = [&](uint32_t op, int len) {
					if (cur_len > 0 && op != cur_op) { if (hl == 0) cigar[n_cigar] = (uint32_t)cur_len << 4 | cur_op; ++n_cigar; cur_len = 0; }
					cur_op = op, cur_len += len;
				};
				while (i >= 0 && jj >= 0) {
					const int ci = i - hl, cj = jj - hl;
					uint32_t raw = 0xffu;
					if (ci >= 0 && cj >= 0) {
						const int L = ci / C, c = ci - L * C;
						raw = Pb[((((size_t)(cj + L)) * 32 + L) * CW + (c >> 1)) * 4 + (c & 1)];
					}
					int k = 0;
					if (state == 0) {
						const unsigned stop = (__ballot_sync(gmask, (raw & 7u) != 4u) >> (grp << 4)) & 0xffffu;
						k = stop ? __ffs(stop) - 1 : 16;
						if (k) { push(0, k); i -= k, jj -= k; }
					}
					if (k < 16 && i >= 0 && jj >= 0) {
						const uint32_t rk = __shfl_sync(gmask, raw, (grp << 4) + k);
						const uint32_t tmp = (4u - (rk & 7u)) | (rk & 0x78u);
						if (state == 0) state = tmp & 7;
						else if (!(tmp >> (state + 2) & 1)) state = 0;
						if (state == 0) state = tmp & 7;
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
			__syncwarp();
		}
	}
	if (lane == 0 && cells_out && cells) atomicAdd(cells_out, cells);
}
