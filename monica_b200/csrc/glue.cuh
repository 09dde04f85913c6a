// glue.cuh -- chain -> region logic and MAPQ, written as host/device functions that run one thread per read.
//
// Product-side implementation of minimap2-2.17 hit.c (mm_gen_regs, mm_set_parent, mm_select_sub, mm_sync_regs,
// mm_join_long, mm_filter_regs, mm_hit_sort, mm_set_sam_pri, mm_set_mapq, mm_split_reg, mm_squeeze_a), i.e. the
// part of mm_map_frag between chaining and base-level alignment and after it.  Reached in the reference through
// /root/reference/monica/genomes/aligner.py:193,215 (index.map) and read back at :194-195,216-217.
//
// These are irregular, data-dependent and tiny per read (a handful of regions), so they are not kernels of their
// own merit: they stay on the device only so that the batch never round-trips to the host between the chaining
// and extension kernels.
#pragma once
#include "common.cuh"
#include "chain.cuh"

struct Reg {
	int32_t id, cnt, rid, score;
	int32_t qs, qe, rs, re;
	int32_t parent, subsc, as, mlen, blen, n_sub, score0;
	uint32_t hash;
	uint8_t mapq, split, rev, inv, sam_pri, split_inv, has_p, aligned;
	int32_t dp_score, dp_max, dp_max2, n_ambi;
	int32_t n_cigar;
	int64_t cigar_off;     // into the batch cigar pool
	int32_t next_split;    // index (within the read's reg array) of the region split off this one, or -1
	int32_t slot;          // this region's own index in the read's reg array
	int32_t inv_after;     // index of the inversion hit (mm_align1_inv) that upstream inserts right after this region, or -1
	int32_t inv_state;     // split_inv regions: 1 while their inversion alignment is pending, 2 once it has been decided
};

MB_HD uint64_t mb_hash64_full(uint64_t key)
{
	key = (~key + (key << 21));
	key = key ^ key >> 24;
	key = ((key + (key << 3)) + (key << 8));
	key = key ^ key >> 14;
	key = ((key + (key << 2)) + (key << 4));
	key = key ^ key >> 28;
	key = (key + (key << 31));
	return key;
}

MB_HD uint32_t mb_wang32(uint32_t key)
{
	key += ~(key << 15);
	key ^=  (key >> 10);
	key +=  (key << 3);
	key ^=  (key >> 6);
	key += ~(key << 11);
	key ^=  (key >> 16);
	return key;
}

MB_HD void mb_reg_set_coor(Reg *r, int32_t qlen, const mb128 *a)
{
	int32_t k = r->as, q_span = (int32_t)(a[k].y >> 32 & 0xff);
	r->rev = (uint8_t)(a[k].x >> 63);
	r->rid = (int32_t)(a[k].x << 1 >> 33);
	r->rs = (int32_t)a[k].x + 1 > q_span ? (int32_t)a[k].x + 1 - q_span : 0;
	r->re = (int32_t)a[k + r->cnt - 1].x + 1;
	if (!r->rev) {
		r->qs = (int32_t)a[k].y + 1 - q_span;
		r->qe = (int32_t)a[k + r->cnt - 1].y + 1;
	} else {
		r->qs = qlen - ((int32_t)a[k + r->cnt - 1].y + 1);
		r->qe = qlen - ((int32_t)a[k].y + 1 - q_span);
	}
	// mm_cal_fuzzy_len
	r->mlen = r->blen = 0;
	if (r->cnt <= 0) return;
	r->mlen = r->blen = q_span;
	for (int i = r->as + 1; i < r->as + r->cnt; ++i) {
		int span = (int)(a[i].y >> 32 & 0xff);
		int tl = (int32_t)a[i].x - (int32_t)a[i - 1].x;
		int ql = (int32_t)a[i].y - (int32_t)a[i - 1].y;
		r->blen += tl > ql ? tl : ql;
		r->mlen += tl > span && ql > span ? span : tl < ql ? tl : ql;
	}
}

// mm_gen_regs: z scratch = mb128[n_u]
MB_HD void mb_gen_regs(uint32_t hash, int qlen, int n_u, const uint64_t *u, const mb128 *a, Reg *r, mb128 *z, int *ws)
{
	int i, k;
	for (i = k = 0; i < n_u; ++i) {
		uint32_t h = (uint32_t)mb_hash64_full((mb_hash64_full(a[k].x) + mb_hash64_full(a[k].y)) ^ hash);
		z[i].x = u[i] ^ h;
		z[i].y = (uint64_t)k << 32 | (uint32_t)(int32_t)u[i];
		k += (int32_t)u[i];
	}
	mb_sort_exact(z, n_u, KeyX(), ws);
	for (i = 0; i < n_u >> 1; ++i) { mb128 tmp = z[i]; z[i] = z[n_u - 1 - i], z[n_u - 1 - i] = tmp; }
	for (i = 0; i < n_u; ++i) {
		Reg *ri = &r[i];
		memset(ri, 0, sizeof(Reg));
		ri->id = i;
		ri->parent = MB_PARENT_UNSET;
		ri->score = ri->score0 = (int32_t)(z[i].x >> 32);
		ri->hash = (uint32_t)z[i].x;
		ri->cnt = (int32_t)z[i].y;
		ri->as = (int32_t)(z[i].y >> 32);
		ri->next_split = -1;
		ri->inv_after = -1;
		ri->slot = i;
		mb_reg_set_coor(ri, qlen, a);
	}
}

MB_HD void mb_set_sam_pri(int n, Reg *r)
{
	int n_pri = 0;
	for (int i = 0; i < n; ++i)
		if (r[i].id == r[i].parent) { ++n_pri; r[i].sam_pri = (n_pri == 1); }
		else r[i].sam_pri = 0;
}

// tmp scratch: int[max_id+2]
MB_HD void mb_sync_regs(int n_regs, Reg *regs, int *tmp)
{
	int i, max_id = -1, n_tmp;
	if (n_regs <= 0) return;
	for (i = 0; i < n_regs; ++i) max_id = max_id > regs[i].id ? max_id : regs[i].id;
	n_tmp = max_id + 1;
	for (i = 0; i < n_tmp; ++i) tmp[i] = -1;
	for (i = 0; i < n_regs; ++i)
		if (regs[i].id >= 0) tmp[regs[i].id] = i;
	for (i = 0; i < n_regs; ++i) {
		Reg *r = &regs[i];
		r->id = i;
		if (r->parent == MB_PARENT_TMP_PRI) r->parent = i;
		else if (r->parent >= 0 && tmp[r->parent] >= 0) r->parent = tmp[r->parent];
		else r->parent = MB_PARENT_UNSET;
	}
	mb_set_sam_pri(n_regs, regs);
}

// cov scratch: u64[n]; w scratch: int[n]
MB_HD void mb_set_parent(float mask_level, int n, Reg *r, int sub_diff, uint64_t *cov, int *w, int *ws)
{
	int i, j, k;
	if (n <= 0) return;
	for (i = 0; i < n; ++i) r[i].id = i;
	w[0] = 0, r[0].parent = 0;
	for (i = 1, k = 1; i < n; ++i) {
		Reg *ri = &r[i];
		int si = ri->qs, ei = ri->qe, n_cov = 0, uncov_len = 0;
		bool is_new = false;
		for (j = 0; j < k; ++j) {
			Reg *rp = &r[w[j]];
			int sj = rp->qs, ej = rp->qe;
			if (ej <= si || sj >= ei) continue;
			if (sj < si) sj = si;
			if (ej > ei) ej = ei;
			cov[n_cov++] = (uint64_t)(uint32_t)sj << 32 | (uint32_t)ej;
		}
		if (n_cov == 0) {
			is_new = true;
		} else {
			int x = si;
			mb_sort_exact(cov, n_cov, KeyU64(), ws); // equal keys are identical values: any order gives the same array
			for (int jj = 0; jj < n_cov; ++jj) {
				if ((int)(cov[jj] >> 32) > x) uncov_len += (int)(cov[jj] >> 32) - x;
				x = (int32_t)cov[jj] > x ? (int32_t)cov[jj] : x;
			}
			if (ei > x) uncov_len += ei - x;
			for (j = 0; j < k; ++j) {
				Reg *rp = &r[w[j]];
				int sj = rp->qs, ej = rp->qe, mn, mx, ol;
				if (ej <= si || sj >= ei) continue;
				mn = ej - sj < ei - si ? ej - sj : ei - si;
				mx = ej - sj > ei - si ? ej - sj : ei - si;
				ol = si < sj ? (ei < sj ? 0 : ei < ej ? ei - sj : ej - sj) : (ej < si ? 0 : ej < ei ? ej - si : ei - si);
#ifdef __CUDA_ARCH__
				float lhs = __fsub_rn(__fdiv_rn((float)ol, (float)mn), __fdiv_rn((float)uncov_len, (float)mx));
#else
				float lhs = (float)ol / mn - (float)uncov_len / mx;
#endif
				if (lhs > mask_level) {
					int cnt_sub = 0;
					ri->parent = rp->parent;
					rp->subsc = rp->subsc > ri->score ? rp->subsc : ri->score;
					if (ri->cnt >= rp->cnt) cnt_sub = 1;
					if (rp->has_p && ri->has_p && (rp->rid != ri->rid || rp->rs != ri->rs || rp->re != ri->re || ol != mn)) {
						int sci = ri->dp_max;
						rp->dp_max2 = rp->dp_max2 > sci ? rp->dp_max2 : sci;
						if (rp->dp_max - ri->dp_max <= sub_diff) cnt_sub = 1;
					}
					if (cnt_sub) ++rp->n_sub;
					break;
				}
			}
			if (j == k) is_new = true;
		}
		if (is_new) w[k++] = i, ri->parent = i, ri->n_sub = 0;
	}
}

// tmp scratch: int[n+2]
MB_HD void mb_select_sub(float pri_ratio, int min_diff, int best_n, int *n_, Reg *r, int *tmp)
{
	if (pri_ratio > 0.0f && *n_ > 0) {
		int i, k, n = *n_, n_2nd = 0;
		for (i = k = 0; i < n; ++i) {
			int p = r[i].parent;
			if (p == i || r[i].inv) {
				r[k++] = r[i];
			} else {
#ifdef __CUDA_ARCH__
				float thr = __fmul_rn((float)r[p].score, pri_ratio);
#else
				float thr = r[p].score * pri_ratio;
#endif
				if (((float)r[i].score >= thr || r[i].score + min_diff >= r[p].score) && n_2nd < best_n) {
					if (!(r[i].qs == r[p].qs && r[i].qe == r[p].qe && r[i].rid == r[p].rid && r[i].rs == r[p].rs && r[i].re == r[p].re))
						r[k++] = r[i], ++n_2nd;
				}
			}
		}
		if (k != n) mb_sync_regs(k, r, tmp);
		*n_ = k;
	}
}

MB_HD void mb_filter_regs(const mb_opt_t *opt, int qlen, int *n_regs, Reg *regs)
{
	int i, k;
	for (i = k = 0; i < *n_regs; ++i) {
		Reg *r = &regs[i];
		int flt = 0;
		if (!r->inv && r->cnt < opt->min_cnt) flt = 1;
		if (r->has_p) {
			if (r->mlen < opt->min_chain_score) flt = 1;
			else if (r->dp_max < opt->min_dp_max) flt = 1;
			else {
#ifdef __CUDA_ARCH__
				float lim = __fmul_rn((float)qlen, opt->max_clip_ratio);
#else
				float lim = qlen * opt->max_clip_ratio;
#endif
				if ((float)r->qs > lim && (float)(qlen - r->qe) > lim) flt = 1;
			}
		}
		if (flt) continue;
		if (k < i) regs[k++] = regs[i];
		else ++k;
	}
	*n_regs = k;
}

// aux scratch: mb128[n]; t scratch: Reg[n]
MB_HD void mb_hit_sort(int *n_regs, Reg *r, mb128 *aux, Reg *t, int *ws)
{
	int i, n_aux, n = *n_regs;
	if (n <= 1) return;
	for (i = n_aux = 0; i < n; ++i) {
		if (r[i].inv || r[i].cnt > 0) {
			int score = r[i].has_p ? r[i].dp_max : r[i].score;
			aux[n_aux].x = (uint64_t)(uint32_t)score << 32 | r[i].hash;
			aux[n_aux++].y = (uint64_t)i;
		}
	}
	mb_sort_exact(aux, n_aux, KeyX(), ws);
	for (i = n_aux - 1; i >= 0; --i) t[n_aux - 1 - i] = r[aux[i].y];
	for (i = 0; i < n_aux; ++i) r[i] = t[i];
	*n_regs = n_aux;
}

// aux scratch: u64[n_regs]
MB_HD int mb_squeeze_a(int n_regs, Reg *regs, mb128 *a, uint64_t *aux, int *ws)
{
	int i, as = 0;
	for (i = 0; i < n_regs; ++i) aux[i] = (uint64_t)(uint32_t)regs[i].as << 32 | (uint32_t)i;
	mb_sort_exact(aux, n_regs, KeyU64(), ws);
	for (i = 0; i < n_regs; ++i) {
		Reg *r = &regs[(int32_t)aux[i]];
		if (r->as != as) {
			for (int j = 0; j < r->cnt; ++j) a[as + j] = a[r->as + j]; // memmove to a lower address
			r->as = as;
		}
		as += r->cnt;
	}
	return as;
}

// aux scratch: u64[n_regs]; tmp: int[n_regs+2]
MB_HD void mb_join_long(const mb_opt_t *opt, int qlen, int *n_regs_, Reg *regs, mb128 *a, uint64_t *aux, int *tmp, int *ws)
{
	int i, n_aux, n_regs = *n_regs_, n_drop = 0;
	if (n_regs < 2) return;
	mb_squeeze_a(n_regs, regs, a, aux, ws);
	for (i = n_aux = 0; i < n_regs; ++i)
		if (regs[i].parent == i || regs[i].parent < 0)
			aux[n_aux++] = (uint64_t)(uint32_t)regs[i].as << 32 | (uint32_t)i;
	mb_sort_exact(aux, n_aux, KeyU64(), ws);
	for (i = n_aux - 1; i >= 1; --i) {
		Reg *r0 = &regs[(int32_t)aux[i - 1]], *r1 = &regs[(int32_t)aux[i]];
		const mb128 *a0e, *a1s;
		int max_gap, min_gap, sc_thres, min_flank_len;
		if (r0->as + r0->cnt != r1->as) continue;
		if (r0->rid != r1->rid || r0->rev != r1->rev) continue;
		a0e = &a[r0->as + r0->cnt - 1];
		a1s = &a[r1->as];
		if (a1s->x <= a0e->x || (int32_t)a1s->y <= (int32_t)a0e->y) continue;
		max_gap = min_gap = (int32_t)a1s->y - (int32_t)a0e->y;
		int64_t dx = (int64_t)(a1s->x - a0e->x);
		max_gap = max_gap > dx ? max_gap : (int)dx;
		min_gap = min_gap < dx ? min_gap : (int)dx;
		if (max_gap > opt->max_join_long || min_gap > opt->max_join_short) continue;
#ifdef __CUDA_ARCH__
		sc_thres = (int)__dadd_rn((double)__fmul_rn(__fdiv_rn((float)opt->min_join_flank_sc, (float)opt->max_join_long), (float)max_gap), .499);
		min_flank_len = (int)__fmul_rn((float)max_gap, opt->min_join_flank_ratio);
#else
		sc_thres = (int)((float)opt->min_join_flank_sc / opt->max_join_long * max_gap + .499);
		min_flank_len = (int)(max_gap * opt->min_join_flank_ratio);
#endif
		if (r0->score < sc_thres || r1->score < sc_thres) continue;
		if (r0->re - r0->rs < min_flank_len || r0->qe - r0->qs < min_flank_len) continue;
		if (r1->re - r1->rs < min_flank_len || r1->qe - r1->qs < min_flank_len) continue;
		a[r1->as].y |= MB_SEED_LONG_JOIN;
		r0->cnt += r1->cnt, r0->score += r1->score;
		mb_reg_set_coor(r0, qlen, a);
		r1->cnt = 0;
		r1->parent = r0->id;
		++n_drop;
	}
	if (n_drop > 0) {
		for (i = 0; i < n_regs; ++i) {
			Reg *r = &regs[i];
			if (r->parent >= 0 && r->id != r->parent) {
				if (regs[r->parent].parent >= 0 && regs[r->parent].parent != r->parent)
					r->parent = regs[r->parent].parent;
			}
		}
		mb_filter_regs(opt, qlen, n_regs_, regs);
		mb_sync_regs(*n_regs_, regs, tmp);
	}
}

// mm_split_reg: r2 receives the tail of r starting n anchors in
MB_HD bool mb_split_reg(Reg *r, Reg *r2, int n, int qlen, const mb128 *a)
{
	if (n <= 0 || n >= r->cnt) return false;
	*r2 = *r;
	r2->id = -1;
	r2->sam_pri = 0;
	r2->has_p = 0; r2->n_cigar = 0; r2->dp_score = r2->dp_max = r2->dp_max2 = r2->n_ambi = 0; r2->aligned = 0; r2->next_split = -1;
	r2->split_inv = 0; r2->inv_after = -1; r2->inv_state = 0;
	r2->cnt = r->cnt - n;
#ifdef __CUDA_ARCH__
	r2->score = (int32_t)__dadd_rn((double)__fmul_rn((float)r->score, __fdiv_rn((float)r2->cnt, (float)r->cnt)), .499);
#else
	r2->score = (int32_t)(r->score * ((float)r2->cnt / r->cnt) + .499);
#endif
	r2->as = r->as + n;
	if (r->parent == r->id) r2->parent = MB_PARENT_TMP_PRI;
	mb_reg_set_coor(r2, qlen, a);
	r->cnt -= r2->cnt;
	r->score -= r2->score;
	mb_reg_set_coor(r, qlen, a);
	r->split |= 1, r2->split |= 2;
	return true;
}

// logf with glibc's exact algorithm (sysdeps/ieee754/flt-32/e_logf.c + logf_data.c, i.e. ARM optimized-routines logf:
// 16-entry table, degree-3 polynomial evaluated in double, rounded once to float).  mm_set_mapq truncates a float product
// containing logf(), so MAPQ parity needs the same float as the host libm the CPU path links; glibc's logf is NOT
// correctly rounded, so (float)log((double)x) would differ on a few percent of inputs.  mb_logf_sweep (monica_b200.cu) compares
// this routine with the system logf bit for bit: tests/test_gpu_parity.py sweeps every positive normal float on the device,
// tests/test_boundary_cpu.py the host twin over [1, 2^24) and a stride through the rest.
MB_HD float mb_logf(float x)
{
	const double T[16][2] = {
		{ 0x1.661ec79f8f3bep+0, -0x1.57bf7808caadep-2 }, { 0x1.571ed4aaf883dp+0, -0x1.2bef0a7c06ddbp-2 },
		{ 0x1.49539f0f010bp+0,  -0x1.01eae7f513a67p-2 }, { 0x1.3c995b0b80385p+0, -0x1.b31d8a68224e9p-3 },
		{ 0x1.30d190c8864a5p+0, -0x1.6574f0ac07758p-3 }, { 0x1.25e227b0b8eap+0,  -0x1.1aa2bc79c81p-3 },
		{ 0x1.1bb4a4a1a343fp+0, -0x1.a4e76ce8c0e5ep-4 }, { 0x1.12358f08ae5bap+0, -0x1.1973c5a611cccp-4 },
		{ 0x1.0953f419900a7p+0, -0x1.252f438e10c1ep-5 }, { 0x1p+0, 0x0p+0 },
		{ 0x1.e608cfd9a47acp-1,  0x1.aa5aa5df25984p-5 }, { 0x1.ca4b31f026aap-1,   0x1.c5e53aa362eb4p-4 },
		{ 0x1.b2036576afce6p-1,  0x1.526e57720db08p-3 }, { 0x1.9c2d163a1aa2dp-1,  0x1.bc2860d22477p-3 },
		{ 0x1.886e6037841edp-1,  0x1.1058bc8a07ee1p-2 }, { 0x1.767dcf5534862p-1,  0x1.4043057b6ee09p-2 },
	};
	const double Ln2 = 0x1.62e42fefa39efp-1;
	const double A0 = -0x1.00ea348b88334p-2, A1 = 0x1.5575b0be00b6ap-2, A2 = -0x1.ffffef20a4123p-2;
	uint32_t ix, iz, tmp;
#ifdef __CUDA_ARCH__
	ix = __float_as_uint(x);
#else
	memcpy(&ix, &x, 4);
#endif
	if (ix == 0x3f800000u) return 0.0f;
	if (ix - 0x00800000u >= 0x7f800000u - 0x00800000u) { // zero, subnormal, negative, inf, nan: not reachable from mm_set_mapq
		return (float)log((double)x);
	}
	tmp = ix - 0x3f330000u;
	int i = (int)((tmp >> 19) % 16);
	int k = (int32_t)tmp >> 23;
	iz = ix - (tmp & 0x1ffu << 23);
	float fz;
#ifdef __CUDA_ARCH__
	fz = __uint_as_float(iz);
	double z = (double)fz, invc = T[i][0], logc = T[i][1];
	double r = __dadd_rn(__dmul_rn(z, invc), -1.0);
	double y0 = __dadd_rn(logc, __dmul_rn((double)k, Ln2));
	double r2 = __dmul_rn(r, r);
	double y = __dadd_rn(__dmul_rn(A1, r), A2);
	y = __dadd_rn(__dmul_rn(A0, r2), y);
	y = __dadd_rn(__dmul_rn(y, r2), __dadd_rn(y0, r));
	return __double2float_rn(y);
#else
	memcpy(&fz, &iz, 4);
	double z = (double)fz, invc = T[i][0], logc = T[i][1];
	double r = z * invc - 1;
	double y0 = logc + (double)k * Ln2;
	double r2 = r * r;
	double y = A1 * r + A2;
	y = A0 * r2 + y;
	y = y * r2 + (y0 + r);
	return (float)y;
#endif
}

MB_HD void mb_set_mapq(int n_regs, Reg *regs, int min_chain_sc, int match_sc, int rep_len)
{
	const float q_coef = 40.0f;
	int64_t sum_sc = 0;
	float uniq_ratio;
	int i;
	if (n_regs == 0) return;
	for (i = 0; i < n_regs; ++i)
		if (regs[i].parent == regs[i].id) sum_sc += regs[i].score;
#ifdef __CUDA_ARCH__
	#define FM(a, b) __fmul_rn((a), (b))
	#define FD(a, b) __fdiv_rn((a), (b))
	#define FS(a, b) __fsub_rn((a), (b))
	#define FA(a, b) __fadd_rn((a), (b))
#else
	#define FM(a, b) ((a) * (b))
	#define FD(a, b) ((a) / (b))
	#define FS(a, b) ((a) - (b))
	#define FA(a, b) ((a) + (b))
#endif
	uniq_ratio = FD((float)sum_sc, (float)(sum_sc + rep_len));
	for (i = 0; i < n_regs; ++i) {
		Reg *r = &regs[i];
		if (r->inv) {
			r->mapq = 0;
		} else if (r->parent == r->id) {
			int mapq, subsc;
			float pen_s1 = FM((r->score > 100 ? 1.0f : FM(0.01f, (float)r->score)), uniq_ratio);
			float pen_cm = r->cnt > 10 ? 1.0f : FM(0.1f, (float)r->cnt);
			pen_cm = pen_s1 < pen_cm ? pen_s1 : pen_cm;
			subsc = r->subsc > min_chain_sc ? r->subsc : min_chain_sc;
			if (r->has_p && r->dp_max2 > 0 && r->dp_max > 0) {
				float identity = FD((float)r->mlen, (float)r->blen);
				float x = FD(FD(FM((float)r->dp_max2, (float)subsc), (float)r->dp_max), (float)r->score0);
				mapq = (int)FM(FM(FM(FM(identity, pen_cm), q_coef), FS(1.0f, FM(x, x))), mb_logf(FD((float)r->dp_max, (float)match_sc)));
				int mapq_alt = (int)FA(FD(FM(FM(FM(6.02f, identity), identity), (float)(r->dp_max - r->dp_max2)), (float)match_sc), .499f);
				mapq = mapq < mapq_alt ? mapq : mapq_alt;
			} else {
				float x = FD((float)subsc, (float)r->score0);
				if (r->has_p) {
					float identity = FD((float)r->mlen, (float)r->blen);
					mapq = (int)FM(FM(FM(FM(identity, pen_cm), q_coef), FS(1.0f, x)), mb_logf(FD((float)r->dp_max, (float)match_sc)));
				} else {
					mapq = (int)FM(FM(FM(pen_cm, q_coef), FS(1.0f, x)), mb_logf((float)r->score));
				}
			}
			mapq -= (int)FA(FM(4.343f, mb_logf((float)(r->n_sub + 1))), .499f);
			mapq = mapq > 0 ? mapq : 0;
			r->mapq = (uint8_t)(mapq < 60 ? mapq : 60);
			if (r->has_p && r->dp_max > r->dp_max2 && r->mapq == 0) r->mapq = 1;
		} else r->mapq = 0;
	}
	#undef FM
	#undef FD
	#undef FS
	#undef FA
}

// hit.c mm_set_inv_mapq(): an inversion hit takes the smaller MAPQ of its two neighbours in reference order.
// aux scratch: mb128[n_regs]
MB_HD void mb_set_inv_mapq(int n_regs, Reg *regs, mb128 *aux, int *ws)
{
	int i, n_aux;
	if (n_regs < 3) return;
	for (i = 0; i < n_regs; ++i)
		if (regs[i].inv) break;
	if (i == n_regs) return;
	for (i = n_aux = 0; i < n_regs; ++i)
		if (regs[i].parent == i || regs[i].parent < 0)
			aux[n_aux].y = (uint64_t)i, aux[n_aux++].x = (uint64_t)(uint32_t)regs[i].rid << 32 | (uint32_t)regs[i].rs;
	mb_sort_exact(aux, n_aux, KeyX(), ws);
	for (i = 1; i < n_aux - 1; ++i) {
		Reg *inv = &regs[aux[i].y];
		if (inv->inv) {
			const Reg *l = &regs[aux[i - 1].y], *r = &regs[aux[i + 1].y];
			inv->mapq = l->mapq < r->mapq ? l->mapq : r->mapq;
		}
	}
}
