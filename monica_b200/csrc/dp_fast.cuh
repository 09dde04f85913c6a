// dp_fast.cuh -- K4 fast path: register-resident systolic wavefront for first-pass gap fills.
//
// ~95 % of all DP cells of a map-ont batch come from mm_align1's first-pass gap fills (ksw_extd2 with KSW_EZ_APPROX_MAX:
// global alignment, no Z-drop bookkeeping, backtrack from the corner) on ~200-350 bp windows, where the band (w=751) never
// limits the matrix.  For those tasks the recurrence is plain two-piece affine DP and upstream's 16-lane block artefacts
// cannot be observed (k_dp in align.cuh keeps handling every other task, including all band-limited ones), so the cells
// can be evaluated in any dependency-respecting order:
//
//   * lane L of the warp owns target columns [L*C, L*C+C) (C = ceil(tlen/32), a template constant -> per-column state
//     u, y, y2 and the lane's target bases live in registers),
//   * at step s the lane evaluates query row j = s - L: a systolic skew, so the (t-1, j) operands of a lane's first
//     column are exactly what lane L-1 produced one step earlier and arrive through one packed __shfl_up,
//   * direction bytes are streamed to a per-lane contiguous region of the warp's scratch ([lane][step][C]), one or more
//     32-bit stores per step, and the backtrack indexes that layout directly.
// No shared memory, no per-cell loads.  The end score is H(tlen-1, qlen-1) = sum_t bnd(t) + sum_j v(tlen-1, j), which is
// what upstream's approximate H0 tracking telescopes to.
#pragma once
#include "align.cuh"

#define DPF_WARPS 4

MB_D int dpf_bnd(int r, int q, int e, int e2, int long_thres, int long_diff)
{
	return r == 0 ? -q - e : r < long_thres ? -e : r == long_thres ? long_diff : -e2;
}

template <int C>
__global__ void __launch_bounds__(DPF_WARPS * 32)
k_dp_fast(DpTask *__restrict__ tasks, const int32_t *__restrict__ order, const int32_t *__restrict__ n_order, int32_t *__restrict__ work_ctr,
          const uint8_t *__restrict__ codes, const uint32_t *__restrict__ S, const uint8_t *__restrict__ pool,
          uint32_t *__restrict__ p_scr, size_t p_stride_words, uint32_t *__restrict__ cigar_pool, DpScoring sc, unsigned long long *__restrict__ cells_out)
{
	constexpr int CW = (C + 3) / 4; // 32-bit words of direction bytes per lane per step
	const unsigned FULL = 0xffffffffu;
	const int lane = threadIdx.x & 31;
	const int gw = blockIdx.x * DPF_WARPS + (threadIdx.x >> 5);
	uint32_t *P = p_scr + (size_t)gw * p_stride_words;
	unsigned long long cells = 0;
	const int n_total = *n_order;
	int q = sc.q, e = sc.e, q2 = sc.q2, e2 = sc.e2;
	if (q2 + e2 < q + e) { int t_ = q; q = q2, q2 = t_; t_ = e; e = e2, e2 = t_; }
	int long_thres = e != e2 ? (q2 - q) / (e - e2) - 1 : 0;
	if (q2 + e2 + long_thres * e2 > q + e + long_thres * e) ++long_thres;
	const int long_diff = long_thres * (e - e2) - (q2 - q) - e2;
	const int mch = sc.sc_mch, mis = sc.sc_mis, scN = sc.sc_N;
	for (;;) {
		int oi = 0;
		if (lane == 0) oi = atomicAdd(work_ctr, 1);
		oi = __shfl_sync(FULL, oi, 0);
		if (oi >= n_total) break;
		DpTask &T = tasks[order[oi]];
		const int Q = T.qlen, TL = T.tlen;
		QView qv; qv.codes = T.q_comp == 2 ? pool : codes; qv.idx0 = T.q_idx0; qv.step = T.q_step; qv.comp = T.q_comp == 1;
		TView tv; tv.S = S; tv.bytes = pool; tv.idx0 = T.t_idx0; tv.step = T.t_step; tv.packed = T.t_packed;
		const int t0 = lane * C;
		int tb[C], u[C], y[C], y2[C];
		#pragma unroll
		for (int c = 0; c < C; ++c) {
			const int t = t0 + c;
			tb[c] = t < TL ? tv.at(t) : 4;
			u[c] = dpf_bnd(t, q, e, e2, long_thres, long_diff);
			y[c] = -q - e, y2[c] = -q2 - e2;
		}
		const int c_last = TL - 1 - t0; // column index of tlen-1 inside this lane (if 0 <= c_last < C)
		int hacc = 0;
		int xo = 0, vo = 0, x2o = 0;
		const int n_steps = Q + 31;
		uint32_t *Pl = P + (size_t)lane * n_steps * CW;
		const bool lane_live = t0 < TL;
		for (int s = 0; s < n_steps; ++s) {
			const int j = s - lane;
			const int mine = (xo & 0xff) | (vo & 0xff) << 8 | (x2o & 0xff) << 16;
			const int prev = __shfl_up_sync(FULL, mine, 1);
			if (lane_live && j >= 0 && j < Q) {
				int xl, vl, x2l;
				if (lane == 0) xl = -q - e, x2l = -q2 - e2, vl = dpf_bnd(j, q, e, e2, long_thres, long_diff);
				else xl = (int)(int8_t)(prev & 0xff), vl = (int)(int8_t)(prev >> 8 & 0xff), x2l = (int)(int8_t)(prev >> 16 & 0xff);
				const int qb = qv.at(j);
				uint32_t w[CW];
				#pragma unroll
				for (int k = 0; k < CW; ++k) w[k] = 0;
				#pragma unroll
				for (int c = 0; c < C; ++c) {
					int z = (tb[c] == qb) ? mch : mis;
					if ((tb[c] | qb) & 4) z = scN;
					const int a = xl + vl, b = y[c] + u[c], a2 = x2l + vl, b2 = y2[c] + u[c];
					int d = 0;
					if (a > z) z = a, d = 1;
					if (b > z) z = b, d = 2;
					if (a2 > z) z = a2, d = 3;
					if (b2 > z) z = b2, d = 4;
					z = z < mch ? z : mch;
					const int un = z - vl, vn = z - u[c];
					const int t1 = a - z, t2 = b - z, t3 = a2 - z, t4 = b2 - z;
					if (t1 > -q) d |= 0x08;
					if (t2 > -q) d |= 0x10;
					if (t3 > -q2) d |= 0x20;
					if (t4 > -q2) d |= 0x40;
					xl = max(t1, -q) - e;
					y[c] = max(t2, -q) - e;
					x2l = max(t3, -q2) - e2;
					y2[c] = max(t4, -q2) - e2;
					u[c] = un, vl = vn;
					if (c == c_last) hacc += vn;
					w[c >> 2] |= (uint32_t)d << ((c & 3) * 8);
				}
				xo = xl, vo = vl, x2o = x2l;
				#pragma unroll
				for (int k = 0; k < CW; ++k) Pl[(size_t)s * CW + k] = w[k];
			}
		}
		cells += (unsigned long long)TL * (unsigned)Q; // warp-uniform; lane 0 reports
		// ---- end score: lane owning column tlen-1 ----
		const int owner = (TL - 1) / C;
		int score = 0;
		{
			int bsum = 0;
			for (int r = lane; r < TL; r += 32) bsum += dpf_bnd(r, q, e, e2, long_thres, long_diff);
			#pragma unroll
			for (int dlt = 16; dlt > 0; dlt >>= 1) bsum += __shfl_xor_sync(FULL, bsum, dlt);
			score = bsum + __shfl_sync(FULL, hacc, owner);
		}
		__syncwarp();
		// ---- backtrack from the corner (lane 0) ----
		if (lane == 0) {
			uint32_t *cigar = cigar_pool + T.cigar_off;
			int n_cigar = 0, i = TL - 1, jj = Q - 1, state = 0;
			const uint8_t *Pb = reinterpret_cast<const uint8_t*>(P);
			while (i >= 0 && jj >= 0) {
				const int L = i / C, c = i - L * C;
				const uint32_t tmp = Pb[(((size_t)L * n_steps + (size_t)(jj + L)) * CW) * 4 + c];
				if (state == 0) state = tmp & 7;
				else if (!(tmp >> (state + 2) & 1)) state = 0;
				if (state == 0) state = tmp & 7;
				if (state == 0) dp_push_cigar(cigar, n_cigar, 0, 1), --i, --jj;
				else if (state == 1 || state == 3) dp_push_cigar(cigar, n_cigar, 2, 1), --i;
				else dp_push_cigar(cigar, n_cigar, 1, 1), --jj;
			}
			if (i >= 0) dp_push_cigar(cigar, n_cigar, 2, i + 1);
			if (jj >= 0) dp_push_cigar(cigar, n_cigar, 1, jj + 1);
			for (int k = 0; k < n_cigar >> 1; ++k) { uint32_t t_ = cigar[k]; cigar[k] = cigar[n_cigar - 1 - k], cigar[n_cigar - 1 - k] = t_; }
			T.max_q = T.max_t = T.mqe_t = -1;
			T.max = 0, T.mqe = MB_KSW_NEG_INF;
			T.score = score, T.zdropped = 0, T.reach_end = 0, T.n_cigar = n_cigar;
		}
		__syncwarp();
	}
	if (lane == 0 && cells_out && cells) atomicAdd(cells_out, cells);
}

// task classes of the fast path: columns per lane
#define DPF_NCLASS 7
static const int DPF_C[DPF_NCLASS] = { 4, 6, 8, 10, 12, 16, 24 };

// is this task eligible for the fast path, and in which class?  (-1: no)
static __host__ __device__ inline int dpf_class(int qlen, int tlen, int w, int flag, int skip)
{
	if (skip || flag != MB_EZ_APPROX_MAX || qlen <= 0 || tlen <= 0) return -1;
	if (w < 0) w = tlen > qlen ? tlen : qlen;
	const int mx = tlen > qlen ? tlen : qlen;
	if (mx > w + 1 || tlen > 768 || qlen > 8192) return -1; // the band must never limit the matrix
	const int c = (tlen + 31) / 32;
	const int cls[DPF_NCLASS] = { 4, 6, 8, 10, 12, 16, 24 };
	for (int k = 0; k < DPF_NCLASS; ++k) if (c <= cls[k]) return k;
	return -1;
}
