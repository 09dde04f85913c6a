// dp_fast.cuh -- K4 fast path: register-resident systolic wavefront, two tasks per warp in packed 16x2 integer SIMD.
//
// ~95 % of all DP cells of a map-ont batch come from mm_align1's first-pass gap fills (ksw_extd2 with KSW_EZ_APPROX_MAX:
// global alignment, no Z-drop bookkeeping, backtrack from the corner) on ~200-350 bp windows, where the band (w=751) never
// limits the matrix.  For those tasks the recurrence is plain two-piece affine DP and upstream's 16-lane block artefacts
// cannot be observed (k_dp in align.cuh keeps handling every other task, including all band-limited ones), so the cells
// can be evaluated in any dependency-respecting order:
//
//   * lane L of the warp owns target columns [L*C, L*C+C) (C = ceil(tlen/32), a template constant -> per-column state
//     lives in registers); at step s the lane evaluates query row j = s - L (systolic skew), so the (t-1, j) operands of
//     a lane's first column are what lane L-1 produced one step earlier and arrive by __shfl_up,
//   * a warp works on TWO tasks at once: every 32-bit register holds task A in its low and task B in its high 16 bits,
//     and the recurrence runs on sm_100a's packed integer SIMD (VIADD.16x2 on the FMA pipe; VIMNMX3.S16x2,
//     VIADDMNMX.S16x2 and PRMT on the ALU pipe) -- 12 ALU-pipe + ~14 FMA-pipe instructions per PAIR of cells.
//
// Number representation (per 16-bit half).  Suzuki-Kasahara difference values are small (u,v in [-(q+e), mch+q+e] ...),
// so every quantity is stored as 8*(value + bias) + tag:
//   * bias B = q+e on u, v, x, y, x2, y2 and 2B on z/a/b/a2/b2 makes u, v, z non-negative, so `z - v` and `z - u` are
//     plain 32-bit subtractions without borrow between the halves, and `0x10000 - z` negates both halves at once
//     (z > 0 always: z >= the substitution score, and 8*(mis+2B) > 0 is checked on the host),
//   * the low 3 bits carry a priority tag (diag 4 > E 3 > F 2 > E2 1 > F2 0): one max over the five tagged candidates
//     yields the winner AND upstream's tie-break (`d = a > z ? 1 : d`, strict, in that order), and the tag of the max is
//     the direction,
//   * the substitution score of a cell is one PRMT: the row's 4-entry byte table (one per task) indexed by a per-column
//     selector built from the two target bases,
//   * x' = max(a - z, -q) - e is one VIADDMNMX against the tagged floor constant; "gap continues" flags are
//     min(x' - floor, 8) (VIADDMNMX again), and the direction byte is tag + flags, i.e. 4 - tag is upstream's d.
// Direction bytes are streamed to the warp's scratch as [step][lane][C/2] words (bytes A:c A:c+1 B:c B:c+1), so the 32
// lanes of a step write one contiguous run; the backtrack (lanes 0-15 for task A, 16-31 for task B) indexes that layout
// directly and fetches 16 cells of the current diagonal per memory round trip.  No shared memory.
// The end score is H(tlen-1, qlen-1) = sum_{j<qlen} bnd(j) + sum_{t<tlen} u(t, qlen-1), which is what upstream's
// approximate H0 tracking telescopes to.  Tasks whose sequences contain an ambiguous base never get here (k_dp_classify).
#pragma once
#include <type_traits>
#include "align.cuh"

#define DPF_WARPS 1
#define DPF_MAX_Q 1024
#define DPF_MAX_T 768

MB_HD int dpf_bnd(int r, int q, int e, int e2, int long_thres, int long_diff)
{
	return r == 0 ? -q - e : r < long_thres ? -e : r == long_thres ? long_diff : -e2;
}

// host+device: does the scoring fit the packed representation?  (defaults: mch 2, mis -4, N -1, q+e = 6 -> yes)
static __host__ __device__ inline bool dpf_scoring_ok(const DpScoring &sc)
{
	int q = sc.q, e = sc.e, q2 = sc.q2, e2 = sc.e2;
	if (q2 + e2 < q + e) { int t_ = q; q = q2, q2 = t_; t_ = e; e = e2, e2 = t_; }
	const int B = q + e;
	if (sc.sc_mch <= 0 || sc.sc_mis >= 0 || q < 0 || e <= 0 || q2 < 0 || e2 <= 0) return false;
	if (8 * (sc.sc_mch + 2 * B) + 7 >= 128) return false;      // substitution bytes must stay below 128 (PRMT sign replicate)
	if (sc.sc_mis + 2 * B <= 0 || sc.sc_N + 2 * B <= 0) return false; // z > 0
	if (sc.sc_mis + 2 * B <= B - e || sc.sc_N + 2 * B <= B - e || sc.sc_mis + 2 * B <= B - e2 || sc.sc_N + 2 * B <= B - e2) return false; // z > D: K - z always borrows
	if (q2 + e2 + sc.sc_mch + 2 * B > 2000) return false;
	return true;
}

MB_D uint32_t dpf_pack2(int v) { return (uint32_t)(uint16_t)(int16_t)v * 0x10001u; }
MB_D uint32_t dpf_and(uint32_t a, uint32_t b) { uint32_t d; asm("and.b32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
MB_D uint32_t dpf_mad(uint32_t a, uint32_t b, uint32_t c) { uint32_t d; asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
// PRMT in its default mode: selector nibble bit 3 replicates the sign of the selected byte (used to produce zero bytes)
MB_D uint32_t dpf_sel(bool p, uint32_t a, uint32_t b) { uint32_t d; asm("{ .reg .pred q; setp.ne.u32 q, %3, 0; selp.b32 %0, %1, %2, q; }" : "=r"(d) : "r"(a), "r"(b), "r"((uint32_t)p)); return d; }
MB_D uint32_t dpf_prmt(uint32_t a, uint32_t b, uint32_t sel) { uint32_t d; asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel)); return d; }

// does a contiguous run of nt4 bytes / 4-bit packed bases hold a code >= 4?  One masked pass over the aligned 32-bit words
// with no early exit: the threads of a warp (one task each) stay converged, which an exit-on-first-hit loop does not.
MB_D bool dpf_bytes_ambig(const uint8_t *p, int64_t lo, int n)
{
	if (n <= 0) return false;
	const int64_t a0 = (int64_t)(reinterpret_cast<uintptr_t>(p) + lo), a1 = a0 + n - 1; // byte addresses
	const uint32_t *w = reinterpret_cast<const uint32_t*>(static_cast<uintptr_t>(a0 & ~(int64_t)3));
	const int nw = (int)((a1 >> 2) - (a0 >> 2)) + 1;
	const uint32_t first = 0x04040404u << ((a0 & 3) << 3), last = 0x04040404u >> ((3 - (a1 & 3)) << 3);
	uint32_t acc = 0;
	for (int i = 0; i < nw; ++i) {
		uint32_t m = 0x04040404u;
		if (i == 0) m &= first;
		if (i == nw - 1) m &= last;
		acc |= w[i] & m;
	}
	return acc != 0;
}
MB_D bool dpf_nibbles_ambig(const uint32_t *S, int64_t lo, int n)
{
	if (n <= 0) return false;
	const int64_t hi = lo + n - 1;
	const uint32_t *w = S + (lo >> 3);
	const int nw = (int)((hi >> 3) - (lo >> 3)) + 1;
	const uint32_t first = 0x44444444u << ((lo & 7) << 2), last = 0x44444444u >> ((7 - (hi & 7)) << 2);
	uint32_t acc = 0;
	for (int i = 0; i < nw; ++i) {
		uint32_t m = 0x44444444u;
		if (i == 0) m &= first;
		if (i == nw - 1) m &= last;
		acc |= w[i] & m;
	}
	return acc != 0;
}
// warp-cooperative forms of the two scans: the 32 lanes read consecutive words of ONE range (coalesced), uniform result
MB_D bool dpf_bytes_ambig_warp(const uint8_t *p, int64_t lo, int n, int lane)
{
	if (n <= 0) return false;
	const int64_t a0 = (int64_t)(reinterpret_cast<uintptr_t>(p) + lo), a1 = a0 + n - 1;
	const uint32_t *w = reinterpret_cast<const uint32_t*>(static_cast<uintptr_t>(a0 & ~(int64_t)3));
	const int nw = (int)((a1 >> 2) - (a0 >> 2)) + 1;
	const uint32_t first = 0x04040404u << ((a0 & 3) << 3), last = 0x04040404u >> ((3 - (a1 & 3)) << 3);
	uint32_t acc = 0;
	for (int i = lane; i < nw; i += 32) {
		uint32_t m = 0x04040404u;
		if (i == 0) m &= first;
		if (i == nw - 1) m &= last;
		acc |= w[i] & m;
	}
	return __any_sync(0xffffffffu, acc != 0);
}
MB_D bool dpf_nibbles_ambig_warp(const uint32_t *S, int64_t lo, int n, int lane)
{
	if (n <= 0) return false;
	const int64_t hi = lo + n - 1;
	const uint32_t *w = S + (lo >> 3);
	const int nw = (int)((hi >> 3) - (lo >> 3)) + 1;
	const uint32_t first = 0x44444444u << ((lo & 7) << 2), last = 0x44444444u >> ((7 - (hi & 7)) << 2);
	uint32_t acc = 0;
	for (int i = lane; i < nw; i += 32) {
		uint32_t m = 0x44444444u;
		if (i == 0) m &= first;
		if (i == nw - 1) m &= last;
		acc |= w[i] & m;
	}
	return __any_sync(0xffffffffu, acc != 0);
}

MB_D bool dpf_task_ambig(const DpTask &t, const uint8_t *codes, const uint32_t *S, const uint8_t *pool)
{
	const uint8_t *qc = t.q_comp == 2 ? pool : codes;
	const int64_t qlo = t.q_step > 0 ? t.q_idx0 : t.q_idx0 - (t.qlen - 1);
	if (dpf_bytes_ambig(qc, qlo, t.qlen)) return true;
	const int64_t tlo = t.t_step > 0 ? t.t_idx0 : t.t_idx0 - (t.tlen - 1);
	return t.t_packed ? dpf_nibbles_ambig(S, tlo, t.tlen) : dpf_bytes_ambig(pool, tlo, t.tlen);
}

template <int C>
__global__ void __launch_bounds__(DPF_WARPS * 32, (C <= 8 ? 20 : C <= 12 ? 16 : 4))
k_dp_fast(DpTask *__restrict__ tasks, const int32_t *__restrict__ order, const int32_t *__restrict__ n_order, int32_t *__restrict__ work_ctr,
          const uint8_t *__restrict__ codes, const uint32_t *__restrict__ S, const uint8_t *__restrict__ pool,
          uint32_t *__restrict__ p_scr, size_t p_stride_words, uint32_t *__restrict__ cigar_pool, DpScoring sc, unsigned long long *__restrict__ cells_out)
{
	constexpr int CW = (C + 1) / 2; // 32-bit words of direction bytes per lane per step (two tasks x C columns)
	const unsigned FULL = 0xffffffffu;
	const int lane = threadIdx.x & 31;
	const bool lane0 = lane == 0;
	__shared__ uint2 ring_all[DPF_WARPS][128];   // substitution tables of the query rows in flight (see the main loop)
	uint2 *ring = ring_all[threadIdx.x >> 5];
	const int gw = blockIdx.x * DPF_WARPS + (threadIdx.x >> 5);
	uint32_t *P = p_scr + (size_t)gw * p_stride_words;
	unsigned long long cells = 0;
	const int n_total = *n_order;
	int q = sc.q, e = sc.e, q2 = sc.q2, e2 = sc.e2;
	if (q2 + e2 < q + e) { int t_ = q; q = q2, q2 = t_; t_ = e; e = e2, e2 = t_; }
	int long_thres = e != e2 ? (q2 - q) / (e - e2) - 1 : 0;
	if (q2 + e2 + long_thres * e2 > q + e + long_thres * e) ++long_thres;
	const int long_diff = long_thres * (e - e2) - (q2 - q) - e2;
	const int B = q + e;
	// packed constants
	const uint32_t X_INIT = dpf_pack2(8 * (-q - e + B) + 3), Y_INIT = dpf_pack2(8 * (-q - e + B) + 2);
	const uint32_t X2_INIT = dpf_pack2(8 * (-q2 - e2 + B) + 1), Y2_INIT = dpf_pack2(8 * (-q2 - e2 + B) + 0);
	const uint32_t FL1 = X_INIT, FL2 = Y_INIT, FL3 = X2_INIT, FL4 = Y2_INIT;            // floors: opening a new gap
	const uint32_t NFL1 = dpf_pack2(-(8 * (-q - e + B) + 3)), NFL2 = dpf_pack2(-(8 * (-q - e + B) + 2));
	const uint32_t NFL3 = dpf_pack2(-(8 * (-q2 - e2 + B) + 1)), NFL4 = dpf_pack2(-(8 * (-q2 - e2 + B) + 0));
	const uint32_t K1 = 0x00010000u + dpf_pack2(8 * (B - e)), K2 = 0x00010000u + dpf_pack2(8 * (B - e2));
	const uint32_t NEG1 = 0xffffffffu + (uint32_t)sc.pad; // sc.pad == 0, opaque to the compiler
	const uint32_t EIGHT = dpf_pack2(8 + sc.pad);   // sc.pad == 0: a run-time value stays in a register instead of being re-materialised per use
	const uint32_t MCHB = (uint32_t)(8 * (sc.sc_mch + 2 * B) + 4), MISB = (uint32_t)(8 * (sc.sc_mis + 2 * B) + 4), NB = (uint32_t)(8 * (sc.sc_N + 2 * B) + 4);
	const uint32_t MIS4 = MISB * 0x01010101u, N4 = NB * 0x01010101u, MDIFF = MCHB - MISB;
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
		int sumA = 0, sumB = 0;
		uint32_t XLo = 0, VLo = 0, X2Lo = 0;
		const int n_steps = Qm + 31;
		const bool lane_live = t0 < TLm;
		// Substitution tables of the query rows: every 32 steps the warp loads the next 32 rows (lane k -> row s+32+k, one
		// coalesced access per task) and turns them into the PRMT tables LA|LB; each step lane 0 picks its row's tables by
		// a broadcast shuffle and the tables then travel down the lanes with the systolic skew (shfl_up), like XL/VL/X2L.
		auto row_tables = [&](int r, uint32_t &la, uint32_t &lb) {
			const int a = r < QA ? qvA.at(r) : 0, b = r < QB ? qvB.at(r) : 0;
			la = a < 4 ? MIS4 + (MDIFF << (a * 8)) : N4, lb = b < 4 ? MIS4 + (MDIFF << (b * 8)) : N4;
		};
		// The tables of rows [32k, 32k + 32) are computed by the 32 lanes one block ahead (registers LAn | LBn) and parked in a
		// 128-entry ring in shared memory when their block starts; lane L then reads row s - L with one 8-byte load per step.
		uint32_t LAn = 0, LBn = 0;
		row_tables(lane, LAn, LBn);
		__syncwarp();   // the previous pair's last reads of the ring
		const uint32_t VB0 = dpf_pack2(8 * (dpf_bnd(0, q, e, e2, long_thres, long_diff) + B)), VB1 = dpf_pack2(8 * (-e + B));
		const uint32_t VB2 = dpf_pack2(8 * (long_diff + B)), VB3 = dpf_pack2(8 * (-e2 + B));
		uint32_t *dst = P + (size_t)lane * CW;
		// One step of the wavefront.  CAPTURE: this step may be some lane's last query row (j == QA-1 or QB-1), where the lane adds
		// its columns' u to the end-score sum; the steps before min(QA, QB) - 1 cannot be, and run without that code.
		auto step = [&](int s, auto capture) {
			constexpr bool CAPTURE = decltype(capture)::value;
			const int j = s - lane;
			if ((s & 31) == 0) {
				ring[(s + lane) & 127] = make_uint2(LAn, LBn);
				__syncwarp();
				if (s + 32 < Qm) row_tables(s + 32 + lane, LAn, LBn);
			}
			const uint2 LAB = ring[(s - lane) & 127];
			const uint32_t LA = LAB.x, LB = LAB.y;
			uint32_t XL = __shfl_up_sync(FULL, XLo, 1), VL = __shfl_up_sync(FULL, VLo, 1), X2L = __shfl_up_sync(FULL, X2Lo, 1);
			// lane 0 takes the matrix's left boundary instead of a neighbour's values (selects on a loop-invariant predicate; the
			// boundary value of row s is warp-uniform)
			const uint32_t VB = s == 0 ? VB0 : s < long_thres ? VB1 : s == long_thres ? VB2 : VB3;
			XL = dpf_sel(lane0, X_INIT, XL), X2L = dpf_sel(lane0, X2_INIT, X2L), VL = dpf_sel(lane0, VB, VL);
			if (lane_live && j >= 0 && j < Qm) {
				uint32_t wv[CW];
				uint32_t wprev = 0;
				#pragma unroll
				for (int c = 0; c < C; ++c) {
					const uint32_t z0 = dpf_prmt(LA, LB, SEL[c]);
					const uint32_t a = __vadd2(XL, VL), b = __vadd2(Y[c], U[c]), a2 = __vadd2(X2L, VL), b2 = __vadd2(Y2[c], U[c]);
					const uint32_t zt = __vimax3_s16x2(__vimax3_s16x2(z0, a, b), a2, b2);
					const uint32_t zc = dpf_and(zt, 0xfff8fff8u);          // opaque: keeps `zt - zc` a subtraction (FMA pipe) instead of a 2nd LOP3
					const uint32_t un = zc - VL, vn = zc - U[c];          // halves are non-negative: no borrow
					// per-half D - zc: the low half always borrows (zc > D, checked on the host), which the 0x10000 pays back
					const uint32_t nz1 = dpf_mad(zc, NEG1, K1), nz2 = dpf_mad(zc, NEG1, K2); // K - zc as a register-operand IMAD: FMA pipe (ptxas turns the immediate form into an ALU-pipe IADD3)
					const uint32_t r1 = __viaddmax_s16x2(a, nz1, FL1), r2 = __viaddmax_s16x2(b, nz1, FL2);
					const uint32_t r3 = __viaddmax_s16x2(a2, nz2, FL3), r4 = __viaddmax_s16x2(b2, nz2, FL4);
					const uint32_t g1 = __viaddmin_s16x2(r1, NFL1, EIGHT), g2 = __viaddmin_s16x2(r2, NFL2, EIGHT);
					const uint32_t g3 = __viaddmin_s16x2(r3, NFL3, EIGHT), g4 = __viaddmin_s16x2(r4, NFL4, EIGHT);
					// tag | x-cont 0x08 | y-cont 0x10 | x2-cont 0x20 | y2-cont 0x40, as a chain of 2-input multiply-adds (FMA pipe)
					const uint32_t wd = dpf_mad(g4, 8u, dpf_mad(g3, 4u, dpf_mad(g2, 2u, dpf_mad(zc, NEG1, zt) + g1)));
					XL = r1, X2L = r3, Y[c] = r2, Y2[c] = r4, U[c] = un, VL = vn;
					if (c & 1) wv[c >> 1] = dpf_prmt(wprev, wd, 0x6240u);        // bytes A:c-1 A:c B:c-1 B:c
					else if (c == C - 1) wv[c >> 1] = dpf_prmt(wd, 0u, 0x6240u);
					wprev = wd;
				}
				XLo = XL, VLo = VL, X2Lo = X2L;
				// [step][lane][CW]: the 32 lanes of a step write one contiguous run
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
				if (CAPTURE) {
					if (j == QA - 1) {
						#pragma unroll
						for (int c = 0; c < C; ++c) if (t0 + c < TLA) sumA += (int)(U[c] & 0xffffu);
					}
					if (j == QB - 1) {
						#pragma unroll
						for (int c = 0; c < C; ++c) if (t0 + c < TLB) sumB += (int)(U[c] >> 16);
					}
				}
			}
		};
		{
			int s = 0;
			const int s_plain = (QA < QB ? QA : QB) - 1;   // steps [0, s_plain): no lane is on its last query row yet
			#pragma unroll 2
			for (; s < s_plain; ++s, dst += 32 * CW) step(s, std::false_type());
			#pragma unroll 1
			for (; s < n_steps; ++s, dst += 32 * CW) step(s, std::true_type());
		}
		cells += (unsigned long long)TLA * (unsigned)QA + (hasB ? (unsigned long long)TLB * (unsigned)QB : 0ULL); // warp-uniform; lane 0 reports
		// ---- end scores ----
		int bsA = 0, bsB = 0;
		for (int r = lane; r < Qm; r += 32) {
			const int bv = dpf_bnd(r, q, e, e2, long_thres, long_diff);
			if (r < QA) bsA += bv;
			if (r < QB) bsB += bv;
		}
		#pragma unroll
		for (int dlt = 16; dlt > 0; dlt >>= 1) {
			bsA += __shfl_xor_sync(FULL, bsA, dlt), bsB += __shfl_xor_sync(FULL, bsB, dlt);
			sumA += __shfl_xor_sync(FULL, sumA, dlt), sumB += __shfl_xor_sync(FULL, sumB, dlt);
		}
		const int scoreA = bsA + (sumA >> 3) - B * TLA, scoreB = bsB + (sumB >> 3) - B * TLB;
		__syncwarp();
		// ---- backtrack from the corner: lanes 0-15 -> task A, lanes 16-31 -> task B.  The walk is a serial state machine,
		// but the path mostly runs along a diagonal: each round the 16 lanes of a half fetch the bytes of the next 16 cells
		// on the current diagonal in parallel, every lane replays the (uniform) state machine on them through shuffles until
		// the path leaves the diagonal, and lane 0 of the half writes the CIGAR.  One memory latency per diagonal run.
		{
			const int grp = lane >> 4, hl = lane & 15;
			const unsigned gmask = grp ? 0xffff0000u : 0x0000ffffu;
			if (grp == 0 || hasB) {
				DpTask &T = grp ? TB : TA;
				const int TL = grp ? TLB : TLA, Q = grp ? QB : QA;
				uint32_t *cigar = cigar_pool + T.cigar_off;
				int n_cigar = 0, i = TL - 1, jj = Q - 1, state = 0;
				uint32_t cur_op = 0; int cur_len = 0;
				const uint8_t *Pb = reinterpret_cast<const uint8_t*>(P) + (grp ? 2 : 0); // plain loads: the __syncwarp() above orders them after the forward pass
				auto push = [&](uint32_t op, int len) {
					if (cur_len > 0 && op != cur_op) { if (hl == 0) cigar[n_cigar] = (uint32_t)cur_len << 4 | cur_op; ++n_cigar; cur_len = 0; }
					cur_op = op, cur_len += len;
				};
				while (i >= 0 && jj >= 0) {
					const int ci = i - hl, cj = jj - hl;
					uint32_t raw = 0xffu; // off the matrix: not a match, stops a run
					if (ci >= 0 && cj >= 0) {
						const int L = ci / C, c = ci - L * C;
						raw = Pb[((((size_t)(cj + L)) * 32 + L) * CW + (c >> 1)) * 4 + (c & 1)];
					}
					int k = 0;
					if (state == 0) { // a run of diagonal moves is consumed at once: leading cells whose direction is "match"
						const unsigned stop = (__ballot_sync(gmask, (raw & 7u) != 4u) >> (grp << 4)) & 0xffffu;
						k = stop ? __ffs(stop) - 1 : 16;
						if (k) { push(0, k); i -= k, jj -= k; }
					}
					if (k < 16 && i >= 0 && jj >= 0) { // one general step of ksw_backtrack on the next fetched cell
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
		}
		__syncwarp();
	}
	if (lane == 0 && cells_out && cells) atomicAdd(cells_out, cells);
}

// task classes of the fast path: columns per lane
#define DPF_NCLASS 13
static const int DPF_C[DPF_NCLASS] = { 4, 5, 6, 7, 8, 9, 10, 11, 12, 14, 16, 20, 24 };

// is this task eligible for the fast path, and in which class?  (-1: no).  The caller also checks dpf_scoring_ok()
// and that neither sequence holds an ambiguous base.
static __host__ __device__ inline int dpf_class(int qlen, int tlen, int w, int flag, int skip)
{
	if (skip || flag != MB_EZ_APPROX_MAX || qlen <= 0 || tlen <= 0) return -1;
	if (w < 0) w = tlen > qlen ? tlen : qlen;
	const int mx = tlen > qlen ? tlen : qlen;
	if (mx > w + 1 || tlen > DPF_MAX_T || qlen > DPF_MAX_Q) return -1; // the band must never limit the matrix
	const int c = (tlen + 31) / 32;
	const int cls[DPF_NCLASS] = { 4, 5, 6, 7, 8, 9, 10, 11, 12, 14, 16, 20, 24 };
	for (int k = 0; k < DPF_NCLASS; ++k) if (c <= cls[k]) return k;
	return -1;
}
