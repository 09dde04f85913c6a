// align.cuh -- K4: base-level alignment of chained regions on sm_100a.
//
// Replaces minimap2-2.17 align.c (mm_align_skeleton / mm_align1 / mm_align_pair / mm_test_zdrop / mm_fix_bad_ends /
// mm_filter_bad_seeds[_alt] / mm_fix_cigar / mm_update_extra) and ksw2_extd2_sse.c (ksw_extd2_sse, + ksw2.h
// ksw_backtrack / ksw_apply_zdrop), reached in the reference through mappy's Aligner.map
// (/root/reference/monica/genomes/aligner.py:193,215; mappy always sets MM_F_CIGAR).
//
// Structure.  mm_align1 is sequential per region, but every ksw call's INPUT depends only on the anchors (left
// extension, one gap fill per >=200 bp anchor pair, right extension); only the bookkeeping between calls is
// sequential.  So a batch is aligned in rounds of flat kernels:
//   plan1  (thread/region)  anchor filtering, extension windows, number of DP tasks
//   plan2  (thread/region)  emit DP task descriptors
//   k_dp   (warp/task)      two-piece affine banded DP + traceback           <- the hot kernel
//   ztest  (thread/task)    mm_test_zdrop on gap-fill CIGARs -> tasks that need the exact second pass
//   k_dp   (warp/task)      second pass for those
//   stitch (thread/region)  concatenate CIGARs, Z-drop split -> new regions (next round), mm_update_extra
// Regions split off by Z-drop are aligned in the next round; rounds repeat until no region is pending.
//
// The DP kernel reproduces ksw_extd2_sse exactly, including what the SSE kernel computes OUTSIDE the band: upstream
// processes 16-lane blocks from the 16-aligned st..en, so lanes beyond [st0,en0] hold well-defined values that in-band
// cells and the backtrack can observe when the band (w=751) actually limits the matrix.  The device kernel therefore
// keeps upstream's per-diagonal state arrays (u,v,x,y,x2,y2,s as int8 indexed by target position, contiguous with sf|qr
// like upstream's single allocation) in shared memory, walks anti-diagonals r = 0..qlen+tlen-2, and lets the 32 lanes
// of a warp cover t = st..en of one diagonal 32 cells at a time; the (t-1) operands come from the neighbouring lane
// by shuffle.  Direction bytes go to a per-warp scratch matrix in HBM with upstream's [r][t-st] layout.
#pragma once
#include "common.cuh"
#include "glue.cuh"
#include "seed.cuh"

struct RegPlan {
	int32_t as1, cnt1;
	int32_t rs, qs, re, qe;       // adjusted first / last anchor
	int32_t rs0, qs0, re0, qe0;   // extension windows
	int32_t n_tasks, has_left, has_right, pad;
	int64_t task0;                // first task index (after scan)
};

struct DpTask {
	int32_t reg;                  // global region index (product path) or -1
	int32_t kind;                 // 0 left extension, 1 gap fill, 2 right extension
	int32_t qlen, tlen, w, zdrop, end_bonus, flag;
	int64_t q_idx0, t_idx0;       // first element; see QView/TView
	int8_t  q_step, q_comp, t_step, t_packed;
	int32_t anchor_i;             // gap fill: index i within [as1, as1+cnt1)
	int32_t qs, rs;               // gap fill: window start in oriented query / reference coordinates
	int32_t skip;                 // 1: tlen*qlen > max_sw_mat -> zdropped without DP
	int64_t cigar_off;            // into cigar pool, capacity qlen+tlen+1
	// results
	int32_t score, max, max_q, max_t, mqe, mqe_t, zdropped, reach_end, n_cigar, zdrop_code;
};

struct QView {
	const uint8_t *codes; int64_t idx0; int step, comp;
	MB_D int at(int j) const { int c = codes[idx0 + (int64_t)j * step]; return comp ? (c < 4 ? 3 - c : 4) : c; }
};
struct TView {
	const uint32_t *S; const uint8_t *bytes; int64_t idx0; int step, packed;
	MB_D int at(int i) const {
		int64_t p = idx0 + (int64_t)i * step;
		return packed ? (int)(S[p >> 3] >> ((p & 7) << 2) & 0xf) : (int)bytes[p];
	}
};

// Sequential readers for the per-thread CIGAR walks (k_ztest, k_stitch): cache the aligned 32-bit word that holds the
// current base, so a run of consecutive positions costs one load per 4 query bytes / 8 packed reference bases.
struct QCur {
	const uint8_t *codes; int64_t idx0; int step, comp; int64_t wi; uint32_t w;
	MB_D void init(const QView &v) { codes = v.codes, idx0 = v.idx0, step = v.step, comp = v.comp, wi = -1, w = 0; }
	MB_D int at(int j) {
		const int64_t p = idx0 + (int64_t)j * step, k = p >> 2;
		if (k != wi) { wi = k; w = *reinterpret_cast<const uint32_t*>(codes + (k << 2)); }
		const int c = (int)(w >> ((p & 3) << 3) & 0xff);
		return comp ? (c < 4 ? 3 - c : 4) : c;
	}
};
struct TCur {
	const uint32_t *S; int64_t idx0; int step; int64_t wi; uint32_t w;
	MB_D void init(const TView &v) { S = v.S, idx0 = v.idx0, step = v.step, wi = -1, w = 0; } // packed views only
	MB_D int at(int i) {
		const int64_t p = idx0 + (int64_t)i * step, k = p >> 3;
		if (k != wi) { wi = k; w = S[k]; }
		return (int)(w >> ((p & 7) << 2) & 0xf);
	}
};

struct AlignCtx {
	const uint8_t *codes;         // nt4 read codes
	const int64_t *read_off;
	DevIndex ix;
	mb_opt_t opt;
};

// ------------------------------------------------------------------------------------------------
// plan1: mm_align1 up to (and excluding) the first mm_align_pair
// ------------------------------------------------------------------------------------------------
MB_D int mb_collect_long_gaps(int as1, int cnt1, const mb128 *a, int min_gap, int *K)
{
	// one thread walks its region's anchors: four loads in flight per round trip instead of one
	int n = 0;
	int32_t py = (int32_t)a[as1].y, px = (int32_t)a[as1].x;
	int i = 1;
	for (; i + 4 <= cnt1; i += 4) {
		mb128 v[4];
		#pragma unroll
		for (int u = 0; u < 4; ++u) v[u] = a[as1 + i + u];
		#pragma unroll
		for (int u = 0; u < 4; ++u) {
			const int32_t cy = (int32_t)v[u].y, cx = (int32_t)v[u].x;
			const int gap = (cy - py) - (cx - px);
			if (gap < -min_gap || gap > min_gap) K[n++] = i + u;
			py = cy, px = cx;
		}
	}
	for (; i < cnt1; ++i) {
		const int32_t cy = (int32_t)a[as1 + i].y, cx = (int32_t)a[as1 + i].x;
		const int gap = (cy - py) - (cx - px);
		if (gap < -min_gap || gap > min_gap) K[n++] = i;
		py = cy, px = cx;
	}
	return n <= 1 ? 0 : n;
}

MB_D void mb_fix_bad_ends(const Reg *r, const mb128 *a, int bw, int min_match, int32_t *as, int32_t *cnt)
{
	int32_t i, l, m;
	*as = r->as, *cnt = r->cnt;
	if (r->cnt < 3) return;
	m = l = (int32_t)(a[r->as].y >> 32 & 0xff);
	for (i = r->as + 1; i < r->as + r->cnt - 1; ++i) {
		int32_t lq, lr, mn, mx;
		int32_t q_span = (int32_t)(a[i].y >> 32 & 0xff);
		if (a[i].y & MB_SEED_LONG_JOIN) break;
		lr = (int32_t)a[i].x - (int32_t)a[i - 1].x;
		lq = (int32_t)a[i].y - (int32_t)a[i - 1].y;
		mn = lr < lq ? lr : lq;
		mx = lr > lq ? lr : lq;
		if (mx - mn > l >> 1) *as = i;
		l += mn;
		m += mn < q_span ? mn : q_span;
		if (l >= bw << 1 || (m >= min_match && m >= bw) || m >= r->mlen >> 1) break;
	}
	*cnt = r->as + r->cnt - *as;
	m = l = (int32_t)(a[r->as + r->cnt - 1].y >> 32 & 0xff);
	for (i = r->as + r->cnt - 2; i > *as; --i) {
		int32_t lq, lr, mn, mx;
		int32_t q_span = (int32_t)(a[i + 1].y >> 32 & 0xff);
		if (a[i + 1].y & MB_SEED_LONG_JOIN) break;
		lr = (int32_t)a[i + 1].x - (int32_t)a[i].x;
		lq = (int32_t)a[i + 1].y - (int32_t)a[i].y;
		mn = lr < lq ? lr : lq;
		mx = lr > lq ? lr : lq;
		if (mx - mn > l >> 1) *cnt = i + 1 - *as;
		l += mn;
		m += mn < q_span ? mn : q_span;
		if (l >= bw << 1 || (m >= min_match && m >= bw) || m >= r->mlen >> 1) break;
	}
}

MB_D void mb_filter_bad_seeds(int as1, int cnt1, mb128 *a, int min_gap, int diff_thres, int max_ext_len, int max_ext_cnt, int *K)
{
	int max_st, max_en, n, i, k, mx;
	n = mb_collect_long_gaps(as1, cnt1, a, min_gap, K);
	if (n == 0) return;
	mx = 0, max_st = max_en = -1;
	for (k = 0;; ++k) {
		int gap, l, n_ins = 0, n_del = 0, qs, rs, max_diff = 0, max_diff_l = -1;
		if (k == n || k >= max_en) {
			if (max_en > 0)
				for (i = K[max_st]; i < K[max_en]; ++i)
					a[as1 + i].y |= MB_SEED_IGNORE;
			mx = 0, max_st = max_en = -1;
			if (k == n) break;
		}
		i = K[k];
		gap = ((int32_t)a[as1 + i].y - (int32_t)a[as1 + i - 1].y) - (int32_t)(a[as1 + i].x - a[as1 + i - 1].x);
		if (gap > 0) n_ins += gap;
		else n_del += -gap;
		qs = (int32_t)a[as1 + i - 1].y;
		rs = (int32_t)a[as1 + i - 1].x;
		for (l = k + 1; l < n && l <= k + max_ext_cnt; ++l) {
			int j = K[l], diff;
			if ((int32_t)a[as1 + j].y - qs > max_ext_len || (int32_t)a[as1 + j].x - rs > max_ext_len) break;
			gap = ((int32_t)a[as1 + j].y - (int32_t)a[as1 + j - 1].y) - (int32_t)(a[as1 + j].x - a[as1 + j - 1].x);
			if (gap > 0) n_ins += gap;
			else n_del += -gap;
			int ad = n_ins - n_del; if (ad < 0) ad = -ad;
			diff = n_ins + n_del - ad;
			if (max_diff < diff) max_diff = diff, max_diff_l = l;
		}
		if (max_diff > diff_thres && max_diff > mx)
			mx = max_diff, max_st = k, max_en = max_diff_l;
	}
}

MB_D void mb_filter_bad_seeds_alt(int as1, int cnt1, mb128 *a, int min_gap, int max_ext, int *K)
{
	int n, k;
	n = mb_collect_long_gaps(as1, cnt1, a, min_gap, K);
	if (n == 0) return;
	for (k = 0; k < n;) {
		int i = K[k], l;
		int gap1 = ((int32_t)a[as1 + i].y - (int32_t)a[as1 + i - 1].y) - ((int32_t)a[as1 + i].x - (int32_t)a[as1 + i - 1].x);
		int re1 = (int32_t)a[as1 + i].x;
		int qe1 = (int32_t)a[as1 + i].y;
		gap1 = gap1 > 0 ? gap1 : -gap1;
		for (l = k + 1; l < n; ++l) {
			int j = K[l], gap2, q_span_pre, rs2, qs2, m;
			if ((int32_t)a[as1 + j].y - qe1 > max_ext || (int32_t)a[as1 + j].x - re1 > max_ext) break;
			gap2 = ((int32_t)a[as1 + j].y - (int32_t)a[as1 + j - 1].y) - (int32_t)(a[as1 + j].x - a[as1 + j - 1].x);
			q_span_pre = (int)(a[as1 + j - 1].y >> 32 & 0xff);
			rs2 = (int32_t)a[as1 + j - 1].x + q_span_pre;
			qs2 = (int32_t)a[as1 + j - 1].y + q_span_pre;
			m = rs2 - re1 < qs2 - qe1 ? rs2 - re1 : qs2 - qe1;
			gap2 = gap2 > 0 ? gap2 : -gap2;
			if (m > gap1 + gap2) break;
			re1 = (int32_t)a[as1 + j].x;
			qe1 = (int32_t)a[as1 + j].y;
			gap1 = gap2;
		}
		if (l > k + 1) {
			int j, end = K[l - 1];
			for (j = K[k]; j < end; ++j) a[as1 + j].y |= MB_SEED_IGNORE;
			a[as1 + end].y |= MB_SEED_LONG_JOIN;
		}
		k = l;
	}
}

// walk the gap-filling loop of mm_align1; EMIT=false counts tasks, EMIT=true fills descriptors
template <bool EMIT>
MB_D int mb_walk_tasks(const AlignCtx &c, const Reg *r, int reg_idx, int read, const RegPlan &pl, const mb128 *a, DpTask *tasks)
{
	const mb_opt_t &opt = c.opt;
	const int k2 = c.ix.k >> 1;
	const int64_t roff = c.read_off[read];
	const int qlen = (int)(c.read_off[read + 1] - roff);
	const int rev = r->rev, rid = r->rid;
	const int64_t toff = (int64_t)c.ix.seq_off[rid];
	const int bw = (int)(opt.bw * 1.5 + 1.);
	int n = 0;
	auto set_q = [&](DpTask &t, int qs, int len, bool reversed) {
		if (!rev) {
			if (!reversed) t.q_idx0 = roff + qs, t.q_step = 1, t.q_comp = 0;
			else t.q_idx0 = roff + qs + len - 1, t.q_step = -1, t.q_comp = 0;
		} else {
			if (!reversed) t.q_idx0 = roff + qlen - 1 - qs, t.q_step = -1, t.q_comp = 1;
			else t.q_idx0 = roff + qlen - qs - len, t.q_step = 1, t.q_comp = 1;
		}
	};
	auto set_t = [&](DpTask &t, int rs, int len, bool reversed) {
		t.t_packed = 1;
		if (!reversed) t.t_idx0 = toff + rs, t.t_step = 1;
		else t.t_idx0 = toff + rs + len - 1, t.t_step = -1;
	};
	auto init = [&](DpTask &t, int kind, int ql, int tl, int w, int zdrop, int end_bonus, int flag) {
		t.reg = reg_idx, t.kind = kind, t.qlen = ql, t.tlen = tl, t.w = w, t.zdrop = zdrop, t.end_bonus = end_bonus, t.flag = flag;
		t.anchor_i = -1, t.qs = t.rs = 0, t.zdrop_code = 0, t.n_cigar = 0;
		t.skip = (opt.max_sw_mat > 0 && (int64_t)tl * ql > opt.max_sw_mat) ? 1 : 0;
	};
	if (pl.has_left) {
		if (EMIT) {
			DpTask &t = tasks[n];
			init(t, 0, pl.qs - pl.qs0, pl.rs - pl.rs0, bw, r->split_inv ? opt.zdrop_inv : opt.zdrop, opt.end_bonus, MB_EZ_EXTZ_ONLY | MB_EZ_RIGHT | MB_EZ_REV_CIGAR);
			set_q(t, pl.qs0, pl.qs - pl.qs0, true);
			set_t(t, pl.rs0, pl.rs - pl.rs0, true);
		}
		++n;
	}
	int rs = pl.rs, qs = pl.qs, re, qe;
	// anchors are read four at a time (one round trip per four anchors), then handled in order
	auto anchor_step = [&](int i, const mb128 &ai) {
		if ((ai.y & (MB_SEED_IGNORE | MB_SEED_TANDEM)) && i != pl.cnt1 - 1) return;
		re = (int32_t)ai.x - k2, qe = (int32_t)ai.y - k2;
		if (i == pl.cnt1 - 1 || (ai.y & MB_SEED_LONG_JOIN) || (qe - qs >= opt.min_ksw_len && re - rs >= opt.min_ksw_len)) {
			if (EMIT) {
				int bw1 = bw;
				if (ai.y & MB_SEED_LONG_JOIN) bw1 = qe - qs > re - rs ? qe - qs : re - rs;
				DpTask &t = tasks[n];
				init(t, 1, qe - qs, re - rs, bw1, opt.zdrop, -1, MB_EZ_APPROX_MAX);
				t.anchor_i = i, t.qs = qs, t.rs = rs;
				set_q(t, qs, qe - qs, false);
				set_t(t, rs, re - rs, false);
			}
			++n;
			rs = re, qs = qe;
		}
	};
	for (int i0 = 1; i0 < pl.cnt1; i0 += 4) {
		mb128 pre[4];
		#pragma unroll
		for (int u = 0; u < 4; ++u) if (i0 + u < pl.cnt1) pre[u] = a[pl.as1 + i0 + u];
		#pragma unroll
		for (int u = 0; u < 4; ++u) if (i0 + u < pl.cnt1) anchor_step(i0 + u, pre[u]);
	}
	if (pl.has_right) {
		if (EMIT) {
			DpTask &t = tasks[n];
			init(t, 2, pl.qe0 - pl.qe, pl.re0 - pl.re, bw, opt.zdrop, opt.end_bonus, MB_EZ_EXTZ_ONLY);
			set_q(t, pl.qe, pl.qe0 - pl.qe, false);
			set_t(t, pl.re, pl.re0 - pl.re, false);
		}
		++n;
	}
	return n;
}

// per-read view of the batch arrays handed to the region kernels
struct ReadArrays {
	mb128 *a;                 // chained anchors (all reads)
	const int64_t *a_roff;    // [n_reads+1]
	const int32_t *n_a_sq;    // [n_reads] anchors left after mm_squeeze_a
	Reg *regs;                // region pool
	const int64_t *reg_off;   // [n_reads+1] capacity offsets
	int32_t *n_regs;          // [n_reads] live regions per read
	int32_t *iscr;            // int scratch, same offsets as anchors
};

// work item = (read, slot)
__global__ void k_plan1(AlignCtx c, ReadArrays ra, const int2 *__restrict__ work, int n_work, RegPlan *__restrict__ plans)
{
	int wi = blockIdx.x * blockDim.x + threadIdx.x;
	if (wi >= n_work) return;
	const int read = work[wi].x, slot = work[wi].y;
	Reg *r = ra.regs + ra.reg_off[read] + slot;
	mb128 *a = ra.a + ra.a_roff[read];
	const int n_a = ra.n_a_sq[read];
	const mb_opt_t &opt = c.opt;
	const int qlen = (int)(c.read_off[read + 1] - c.read_off[read]);
	RegPlan pl;
	memset(&pl, 0, sizeof(pl));
	if (r->cnt == 0) { plans[wi] = pl; return; }
	const int32_t rid = (int32_t)(a[r->as].x << 1 >> 33);
	const int32_t tlen_ref = (int32_t)c.ix.seq_len[rid];
	const int k2 = c.ix.k >> 1;
	int32_t as1, cnt1, i, l, rs0, re0, qs0, qe0, rs, re, qs, qe, rs1, qs1, re1, qe1;
	int *K = ra.iscr + ra.a_roff[read] + r->as;
	mb_fix_bad_ends(r, a, opt.bw, opt.min_chain_score * 2, &as1, &cnt1);
	mb_filter_bad_seeds(as1, cnt1, a, 10, 40, opt.max_gap >> 1, 10, K);
	mb_filter_bad_seeds_alt(as1, cnt1, a, 30, opt.max_gap >> 1, K);
	rs = (int32_t)a[as1].x - k2, qs = (int32_t)a[as1].y - k2;
	re = (int32_t)a[as1 + cnt1 - 1].x - k2, qe = (int32_t)a[as1 + cnt1 - 1].y - k2;
	// compute rs0 and qs0
	rs0 = (int32_t)a[r->as].x + 1 - (int32_t)(a[r->as].y >> 32 & 0xff);
	qs0 = (int32_t)a[r->as].y + 1 - (int32_t)(a[r->as].y >> 32 & 0xff);
	if (rs0 < 0) rs0 = 0;
	rs1 = qs1 = 0;
	for (i = r->as - 1, l = 0; i >= 0 && a[i].x >> 32 == a[r->as].x >> 32; --i) {
		int32_t x = (int32_t)a[i].x + 1 - (int32_t)(a[i].y >> 32 & 0xff);
		int32_t y = (int32_t)a[i].y + 1 - (int32_t)(a[i].y >> 32 & 0xff);
		if (x < rs0 && y < qs0) {
			if (++l > opt.min_cnt) {
				l = rs0 - x > qs0 - y ? rs0 - x : qs0 - y;
				rs1 = rs0 - l, qs1 = qs0 - l;
				if (rs1 < 0) rs1 = 0;
				break;
			}
		}
	}
	if (qs > 0 && rs > 0) {
		l = qs < opt.max_gap ? qs : opt.max_gap;
		qs1 = qs1 > qs - l ? qs1 : qs - l;
		qs0 = qs0 < qs1 ? qs0 : qs1;
		l += l * opt.a > opt.q ? (l * opt.a - opt.q) / opt.e : 0;
		l = l < opt.max_gap ? l : opt.max_gap;
		l = l < rs ? l : rs;
		rs1 = rs1 > rs - l ? rs1 : rs - l;
		rs0 = rs0 < rs1 ? rs0 : rs1;
		rs0 = rs0 < rs ? rs0 : rs;
	} else rs0 = rs, qs0 = qs;
	// compute re0 and qe0
	re0 = (int32_t)a[r->as + r->cnt - 1].x + 1;
	qe0 = (int32_t)a[r->as + r->cnt - 1].y + 1;
	re1 = tlen_ref, qe1 = qlen;
	for (i = r->as + r->cnt, l = 0; i < n_a && a[i].x >> 32 == a[r->as].x >> 32; ++i) {
		int32_t x = (int32_t)a[i].x + 1;
		int32_t y = (int32_t)a[i].y + 1;
		if (x > re0 && y > qe0) {
			if (++l > opt.min_cnt) {
				l = x - re0 > y - qe0 ? x - re0 : y - qe0;
				re1 = re0 + l, qe1 = qe0 + l;
				break;
			}
		}
	}
	if (qe < qlen && re < tlen_ref) {
		l = qlen - qe < opt.max_gap ? qlen - qe : opt.max_gap;
		qe1 = qe1 < qe + l ? qe1 : qe + l;
		qe0 = qe0 > qe1 ? qe0 : qe1;
		l += l * opt.a > opt.q ? (l * opt.a - opt.q) / opt.e : 0;
		l = l < opt.max_gap ? l : opt.max_gap;
		l = l < tlen_ref - re ? l : tlen_ref - re;
		re1 = re1 < re + l ? re1 : re + l;
		re0 = re0 > re1 ? re0 : re1;
	} else re0 = re, qe0 = qe;
	pl.as1 = as1, pl.cnt1 = cnt1, pl.rs = rs, pl.qs = qs, pl.re = re, pl.qe = qe;
	pl.rs0 = rs0, pl.qs0 = qs0, pl.re0 = re0, pl.qe0 = qe0;
	pl.has_left = (qs > 0 && rs > 0);
	pl.has_right = (qe < qe0 && re < re0); // evaluated with the last anchor's qe/re; skipped at stitch time if Z-dropped
	pl.n_tasks = mb_walk_tasks<false>(c, r, 0, read, pl, a, nullptr);
	plans[wi] = pl;
}

__global__ void k_plan_ntasks(const RegPlan *__restrict__ plans, int n, int32_t *__restrict__ out)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) out[i] = plans[i].n_tasks;
}

__global__ void k_plan2(AlignCtx c, ReadArrays ra, const int2 *__restrict__ work, int n_work, RegPlan *__restrict__ plans,
                        const int64_t *__restrict__ task_off, DpTask *__restrict__ tasks, int32_t *__restrict__ cig_cap)
{
	int wi = blockIdx.x * blockDim.x + threadIdx.x;
	if (wi >= n_work) return;
	const int read = work[wi].x, slot = work[wi].y;
	Reg *r = ra.regs + ra.reg_off[read] + slot;
	const mb128 *a = ra.a + ra.a_roff[read];
	RegPlan pl = plans[wi];
	pl.task0 = task_off[wi];
	plans[wi] = pl;
	if (pl.n_tasks == 0) return;
	int n = mb_walk_tasks<true>(c, r, wi, read, pl, a, tasks + pl.task0);
	for (int i = 0; i < n; ++i) {
		const DpTask &t = tasks[pl.task0 + i];
		cig_cap[pl.task0 + i] = t.qlen + t.tlen + 1;
	}
}

__global__ void k_set_cigar_off(DpTask *__restrict__ tasks, const int64_t *__restrict__ cig_off, int64_t n_tasks, int64_t base)
{
	int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n_tasks) tasks[i].cigar_off = base + cig_off[i];
}

// ------------------------------------------------------------------------------------------------
// k_dp: ksw_extd2_sse, one warp per task
// ------------------------------------------------------------------------------------------------
#define DP_WARPS 4
#define DP_SMEM_PER_WARP 13568    // covers tlen,qlen <= ~1000: 8*tlen16 + qlen16 + 16 bytes of state + 4*tlen16 of H
#define DP_SMEM_MAX (96 * 1024)   // one-warp CTAs of the long-task classes may use up to this much

struct DpScoring { int8_t q, e, q2, e2, sc_mch, sc_mis, sc_N, pad; };

MB_D void dp_reset(DpTask &t)
{
	t.max_q = t.max_t = t.mqe_t = -1;
	t.max = 0, t.score = t.mqe = MB_KSW_NEG_INF;
	t.n_cigar = 0, t.zdropped = 0, t.reach_end = 0;
}

MB_D void dp_push_cigar(uint32_t *cigar, int &n_cigar, uint32_t op, int len)
{
	if (n_cigar == 0 || op != (cigar[n_cigar - 1] & 0xf)) cigar[n_cigar++] = (uint32_t)len << 4 | op;
	else cigar[n_cigar - 1] += (uint32_t)len << 4;
}

// order[]: task ids, longest first.  p_scr/g_ws/h_scr: per-warp scratch slices (strides in bytes / ints).
__global__ void __launch_bounds__(DP_WARPS * 32)
k_dp(DpTask *__restrict__ tasks, const int32_t *__restrict__ order, const int32_t *__restrict__ n_order, int32_t *__restrict__ work_ctr,
     const uint8_t *__restrict__ codes, const uint32_t *__restrict__ S, const uint8_t *__restrict__ pool,
     uint8_t *__restrict__ p_scr, size_t p_stride, int8_t *__restrict__ g_ws, size_t g_stride, int32_t *__restrict__ h_scr, size_t h_stride,
     uint32_t *__restrict__ cigar_pool, DpScoring sc, unsigned long long *__restrict__ cells_out, int smem_per_warp)
{
	extern __shared__ __align__(16) int8_t dp_smem[];
	const int warps_per_cta = blockDim.x >> 5;   // 4 for the small classes; 1 with a large shared-memory slice for long tasks
	const unsigned FULL = 0xffffffffu;
	const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
	const int gw = blockIdx.x * warps_per_cta + wib;
	uint8_t *P = p_scr + (size_t)gw * p_stride;
	int32_t *H = h_scr + (size_t)gw * h_stride;
	unsigned long long cells = 0;
	const int n_total = *n_order;
	for (;;) {
		int oi = 0;
		if (lane == 0) oi = atomicAdd(work_ctr, 1);
		oi = __shfl_sync(FULL, oi, 0);
		if (oi >= n_total) break;
		DpTask &T = tasks[order[oi]];
		const int qlen = T.qlen, tlen = T.tlen, flag = T.flag, zdrop = T.zdrop, end_bonus = T.end_bonus;
		__syncwarp();
		if (lane == 0) dp_reset(T);
		if (T.skip) { if (lane == 0) T.zdropped = 1; __syncwarp(); continue; }
		if (qlen <= 0 || tlen <= 0) { __syncwarp(); continue; }
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
		int long_thres = e != e2 ? (q2 - q) / (e - e2) - 1 : 0;
		if (q2 + e2 + long_thres * e2 > q + e + long_thres * e) ++long_thres;
		const int long_diff = long_thres * (e - e2) - (q2 - q) - e2;
		const size_t ws_need = (size_t)T16 * 8 + (size_t)qlen_ * 16 + 16;
		// state arrays and (exact mode) the 32-bit H row live in this warp's shared-memory slice when they fit
		const size_t h_need = with_exact ? (size_t)T16 * 4 : 0;
		const bool in_smem = ((ws_need + 15) & ~(size_t)15) + h_need <= (size_t)smem_per_warp;
		int8_t *ws = in_smem ? dp_smem + (size_t)wib * smem_per_warp : g_ws + (size_t)gw * g_stride;
		if (in_smem && with_exact) H = reinterpret_cast<int32_t*>(ws + ((ws_need + 15) & ~(size_t)15));
		else H = h_scr + (size_t)gw * h_stride;
		int8_t *u = ws, *v = u + T16, *x = v + T16, *y = x + T16, *x2 = y + T16, *y2 = x2 + T16, *s = y2 + T16;
		uint8_t *sf = (uint8_t*)(s + T16), *qr = sf + T16;
		// ---- init ----
		{
			QView qv; qv.codes = T.q_comp == 2 ? pool : codes; qv.idx0 = T.q_idx0; qv.step = T.q_step; qv.comp = T.q_comp == 1;
			TView tv; tv.S = S; tv.bytes = pool; tv.idx0 = T.t_idx0; tv.step = T.t_step; tv.packed = T.t_packed;
			const int8_t i1 = (int8_t)(-q - e), i2 = (int8_t)(-q2 - e2);
			for (int t = lane; t < T16; t += 32) {
				u[t] = i1, v[t] = i1, x[t] = i1, y[t] = i1, x2[t] = i2, y2[t] = i2, s[t] = 0;
				sf[t] = t < tlen ? (uint8_t)tv.at(t) : 0;
				if (with_exact) H[t] = MB_KSW_NEG_INF;
			}
			for (int t = lane; t < qlen_ * 16 + 16; t += 32)
				qr[t] = t < qlen ? (uint8_t)qv.at(qlen - 1 - t) : 0;
		}
		__syncwarp();
		// ---- ez state (uniform across lanes) ----
		int ez_max = 0, ez_max_t = -1, ez_max_q = -1, ez_mqe = MB_KSW_NEG_INF, ez_mqe_t = -1, ez_mte = MB_KSW_NEG_INF, ez_score = MB_KSW_NEG_INF;
		int ez_zdropped = 0;
		int32_t H0 = 0, last_H0_t = 0;
		int last_st = -1, last_en = -1;
		const int n_rows = qlen + tlen - 1;
		for (int r = 0; r < n_rows; ++r) {
			int st = 0, en = tlen - 1;
			if (st < r - qlen + 1) st = r - qlen + 1;
			if (en > r) en = r;
			if (st < (r - w + 1) >> 1) st = (r - w + 1) >> 1;
			if (en > (r + w) >> 1) en = (r + w) >> 1;
			if (st > en) { ez_zdropped = 1; break; }
			const int st0 = st, en0 = en;
			st = st / 16 * 16, en = (en + 16) / 16 * 16 - 1;
			int8_t x1, x21, v1;
			if (st > 0) {
				if (st - 1 >= last_st && st - 1 <= last_en) x1 = x[st - 1], x21 = x2[st - 1], v1 = v[st - 1];
				else x1 = (int8_t)(-q - e), x21 = (int8_t)(-q2 - e2), v1 = (int8_t)(-q - e);
			} else {
				x1 = (int8_t)(-q - e), x21 = (int8_t)(-q2 - e2);
				v1 = (int8_t)(r == 0 ? -q - e : r < long_thres ? -e : r == long_thres ? long_diff : -e2);
			}
			__syncwarp();
			if (en >= r && lane == 0) {
				y[r] = (int8_t)(-q - e), y2[r] = (int8_t)(-q2 - e2);
				u[r] = (int8_t)(r == 0 ? -q - e : r < long_thres ? -e : r == long_thres ? long_diff : -e2);
			}
			// scores: 16-byte chunks starting at st0 (upstream's unaligned vector loop); stores past T16 fall into sf[] upstream
			// and are never read again, so they are dropped here
			{
				const uint8_t *qrr = qr + (qlen - 1 - r);
				const int nfill = ((en0 - st0) / 16 + 1) * 16;
				for (int o = lane; o < nfill; o += 32) {
					const int t = st0 + o;
					const uint8_t sq = sf[t], sq2 = qrr[t];
					const int8_t scv = (sq == 4 || sq2 == 4) ? sc.sc_N : (sq == sq2 ? sc.sc_mch : sc.sc_mis);
					if (t < T16) s[t] = scv;
				}
			}
			__syncwarp();
			// core: t = st..en, 32 cells per step
			{
				uint8_t *pr = P + (size_t)r * ncol16 - st;
				int carry = ((int)(uint8_t)x1) | ((int)(uint8_t)v1) << 8 | ((int)(uint8_t)x21) << 16;
				for (int tb = st; tb <= en; tb += 32) {
					const int t = tb + lane;
					const bool act = t <= en;
					int8_t z = 0, xo = 0, vo = 0, x2o = 0, ut = 0, yt = 0, y2t = 0;
					if (act) { z = s[t], xo = x[t], vo = v[t], x2o = x2[t], ut = u[t], yt = y[t], y2t = y2[t]; }
					const int mine = ((int)(uint8_t)xo) | ((int)(uint8_t)vo) << 8 | ((int)(uint8_t)x2o) << 16;
					int prev = __shfl_up_sync(FULL, mine, 1);
					if (lane == 0) prev = carry;
					carry = __shfl_sync(FULL, mine, 31);
					if (act) {
						const int8_t xt1 = (int8_t)(prev & 0xff), vt1 = (int8_t)(prev >> 8 & 0xff), x2t1 = (int8_t)(prev >> 16 & 0xff);
						int8_t a = (int8_t)(xt1 + vt1), b = (int8_t)(yt + ut), a2 = (int8_t)(x2t1 + vt1), b2 = (int8_t)(y2t + ut);
						int8_t d;
						if (!(flag & MB_EZ_RIGHT)) {
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
						u[t] = (int8_t)(z - vt1);
						v[t] = (int8_t)(z - ut);
						int8_t tmp = (int8_t)(z - q);
						a = (int8_t)(a - tmp), b = (int8_t)(b - tmp);
						tmp = (int8_t)(z - q2);
						a2 = (int8_t)(a2 - tmp), b2 = (int8_t)(b2 - tmp);
						if (!(flag & MB_EZ_RIGHT)) {
							x[t]  = (int8_t)((a  > 0 ? a  : 0) - qe);  if (a  > 0) d |= 0x08;
							y[t]  = (int8_t)((b  > 0 ? b  : 0) - qe);  if (b  > 0) d |= 0x10;
							x2[t] = (int8_t)((a2 > 0 ? a2 : 0) - qe2); if (a2 > 0) d |= 0x20;
							y2[t] = (int8_t)((b2 > 0 ? b2 : 0) - qe2); if (b2 > 0) d |= 0x40;
						} else {
							x[t]  = (int8_t)((0 > a  ? 0 : a)  - qe);  if (!(0 > a))  d |= 0x08;
							y[t]  = (int8_t)((0 > b  ? 0 : b)  - qe);  if (!(0 > b))  d |= 0x10;
							x2[t] = (int8_t)((0 > a2 ? 0 : a2) - qe2); if (!(0 > a2)) d |= 0x20;
							y2[t] = (int8_t)((0 > b2 ? 0 : b2) - qe2); if (!(0 > b2)) d |= 0x40;
						}
						pr[t] = (uint8_t)d;
					}
				}
			}
			__syncwarp();
			cells += (unsigned)(en0 - st0 + 1);
			if (with_exact) {
				int32_t max_H, max_t;
				if (r > 0) {
					const int32_t he = en0 > 0 ? H[en0 - 1] + u[en0] : H[en0] + v[en0];
					__syncwarp();
					const int en1 = st0 + (en0 - st0) / 4 * 4;
					long long best = ((long long)he << 32) | (unsigned)(0x7fffffff - 0);
					for (int t = st0 + lane; t < en0; t += 32) {
						const int32_t h = H[t] + (int32_t)v[t];
						H[t] = h;
						const unsigned rank = t < en1 ? 1u + ((unsigned)((t - st0) & 3) << 22) + (unsigned)((t - st0) >> 2) : 1u + (4u << 22) + (unsigned)(t - en1);
						const long long key = ((long long)h << 32) | (unsigned)(0x7fffffffu - rank);
						best = key > best ? key : best;
					}
					if (lane == 0) H[en0] = he;
					#pragma unroll
					for (int d = 16; d > 0; d >>= 1) { long long o = __shfl_xor_sync(FULL, best, d); best = o > best ? o : best; }
					max_H = (int32_t)(best >> 32);
					const unsigned rank = 0x7fffffffu - (unsigned)(best & 0xffffffffu);
					if (rank == 0) max_t = en0;
					else if (rank < 1u + (4u << 22)) { const unsigned rr = rank - 1; max_t = st0 + (int)((rr & ((1u << 22) - 1)) << 2) + (int)(rr >> 22); }
					else max_t = en1 + (int)(rank - 1u - (4u << 22));
				} else {
					if (lane == 0) H[0] = v[0] - qe;
					max_H = v[0] - qe, max_t = 0;
				}
				__syncwarp();
				if (en0 == tlen - 1) { const int32_t h = H[en0]; if (h > ez_mte) ez_mte = h; }
				if (r - st0 == qlen - 1) { const int32_t h = H[st0]; if (h > ez_mqe) ez_mqe = h, ez_mqe_t = st0; }
				// ksw_apply_zdrop
				bool brk = false;
				if (max_H > ez_max) { ez_max = max_H, ez_max_t = max_t, ez_max_q = r - max_t; }
				else if (max_t >= ez_max_t && r - max_t >= ez_max_q) {
					const int tl = max_t - ez_max_t, ql = (r - max_t) - ez_max_q;
					const int l = tl > ql ? tl - ql : ql - tl;
					if (zdrop >= 0 && ez_max - max_H > zdrop + l * e2) { ez_zdropped = 1; brk = true; }
				}
				if (brk) break;
				if (r == qlen + tlen - 2 && en0 == tlen - 1) ez_score = H[tlen - 1];
			} else {
				if (r > 0) {
					if (last_H0_t >= st0 && last_H0_t <= en0 && last_H0_t + 1 >= st0 && last_H0_t + 1 <= en0) {
						const int32_t d0 = v[last_H0_t], d1 = u[last_H0_t + 1];
						if (d0 > d1) H0 += d0;
						else H0 += d1, ++last_H0_t;
					} else if (last_H0_t >= st0 && last_H0_t <= en0) {
						H0 += v[last_H0_t];
					} else {
						++last_H0_t, H0 += u[last_H0_t];
					}
				} else H0 = v[0] - qe, last_H0_t = 0;
				if (r == qlen + tlen - 2 && en0 == tlen - 1) ez_score = H0;
			}
			last_st = st, last_en = en;
		}
		__syncwarp();
		// ---- backtrack (lane 0) ----
		if (lane == 0) {
			int reach_end = 0, n_cigar = 0;
			int i0 = -1, j0 = -1;
			const bool rev_cigar = flag & MB_EZ_REV_CIGAR;
			if (!ez_zdropped && !(flag & MB_EZ_EXTZ_ONLY)) i0 = tlen - 1, j0 = qlen - 1;
			else if (!ez_zdropped && (flag & MB_EZ_EXTZ_ONLY) && ez_mqe + end_bonus > ez_max) reach_end = 1, i0 = ez_mqe_t, j0 = qlen - 1;
			else if (ez_max_t >= 0 && ez_max_q >= 0) i0 = ez_max_t, j0 = ez_max_q;
			uint32_t *cigar = cigar_pool + T.cigar_off;
			if (i0 >= 0 && j0 >= 0) {
				int i = i0, j = j0, state = 0;
				while (i >= 0 && j >= 0) {
					int force_state = -1;
					const int r = i + j;
					int st = 0, en = tlen - 1;
					if (st < r - qlen + 1) st = r - qlen + 1;
					if (en > r) en = r;
					if (st < (r - w + 1) >> 1) st = (r - w + 1) >> 1;
					if (en > (r + w) >> 1) en = (r + w) >> 1;
					st = st / 16 * 16, en = (en + 16) / 16 * 16 - 1;
					if (i < st) force_state = 2;
					if (i > en) force_state = 1;
					const uint32_t tmp = force_state < 0 ? P[(size_t)r * ncol16 + i - st] : 0;
					if (state == 0) state = tmp & 7;
					else if (!(tmp >> (state + 2) & 1)) state = 0;
					if (state == 0) state = tmp & 7;
					if (force_state >= 0) state = force_state;
					if (state == 0) dp_push_cigar(cigar, n_cigar, 0, 1), --i, --j;
					else if (state == 1 || state == 3) dp_push_cigar(cigar, n_cigar, 2, 1), --i;
					else dp_push_cigar(cigar, n_cigar, 1, 1), --j;
				}
				if (i >= 0) dp_push_cigar(cigar, n_cigar, 2, i + 1);
				if (j >= 0) dp_push_cigar(cigar, n_cigar, 1, j + 1);
				if (!rev_cigar)
					for (int k = 0; k < n_cigar >> 1; ++k) { uint32_t t_ = cigar[k]; cigar[k] = cigar[n_cigar - 1 - k], cigar[n_cigar - 1 - k] = t_; }
			}
			T.score = ez_score, T.max = ez_max, T.max_q = ez_max_q, T.max_t = ez_max_t, T.mqe = ez_mqe, T.mqe_t = ez_mqe_t;
			T.zdropped = ez_zdropped, T.reach_end = reach_end, T.n_cigar = n_cigar;
		}
		__syncwarp();
	}
	if (lane == 0 && cells_out && cells) atomicAdd(cells_out, cells);
}
