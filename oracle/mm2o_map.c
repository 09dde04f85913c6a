/*
 * mm2o_map.c -- CPU ORACLE (test infrastructure): chain->region logic, base-level alignment
 * control flow, MAPQ, and the per-read driver.
 * Restates minimap2 v2.17 hit.c (mm_gen_regs, mm_set_parent, mm_select_sub, mm_sync_regs,
 * mm_join_long, mm_filter_regs, mm_hit_sort, mm_set_sam_pri, mm_set_mapq, mm_split_reg,
 * mm_squeeze_a), align.c (mm_align_skeleton, mm_align1, mm_align_pair, mm_test_zdrop,
 * mm_fix_bad_ends, mm_filter_bad_seeds[_alt], mm_fix_cigar, mm_update_extra, mm_append_cigar),
 * map.c (mm_map_frag, chain_post, align_regs) and python/cmappy.h (mm_reg2hitpy).
 * mm_align1_inv (inversion re-alignment between the two halves of a split_inv Z-drop split) and mm_set_inv_mapq are
 * restated below.  NOT restated: mm_est_err (sets mm_reg1_t::div, which mappy does not expose).
 * PARITY UNPINNED -- see mm2o.h.  Call sites in the reference:
 * /root/reference/monica/genomes/aligner.py:193,215 (index.map) and :194-195,216-217 (fields read).
 */
#include <stdlib.h>
#include <string.h>
#include <assert.h>
#include <math.h>
#include <pthread.h>
#include "mm2o.h"

mm128_t *mm2o_collect_seed_hits(const mm2o_opt_t *opt, int max_occ, const mm2o_idx_t *mi, const mm128_v *mv, int qlen, int64_t *n_a, int *rep_len);
const unsigned char *mm2o_nt4_table(void);

typedef struct {
	uint32_t capacity;
	int32_t dp_score, dp_max, dp_max2;
	uint32_t n_ambi:30, trans_strand:2;
	uint32_t n_cigar;
	uint32_t *cigar;
} mm_extra_t;

typedef struct {
	int32_t id;
	int32_t cnt;
	int32_t rid;
	int32_t score;
	int32_t qs, qe, rs, re;
	int32_t parent, subsc;
	int32_t as;
	int32_t mlen, blen;
	int32_t n_sub;
	int32_t score0;
	uint32_t mapq:8, split:2, rev:1, inv:1, sam_pri:1, proper_frag:1, pe_thru:1, seg_split:1, seg_id:8, split_inv:1, dummy:7;
	uint32_t hash;
	float div;
	mm_extra_t *p;
} mm_reg1_t;

static inline uint64_t hash64(uint64_t key)
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

static inline uint32_t wang_hash32(uint32_t key) /* khash.h __ac_Wang_hash */
{
	key += ~(key << 15);
	key ^=  (key >> 10);
	key +=  (key << 3);
	key ^=  (key >> 6);
	key += ~(key << 11);
	key ^=  (key >> 16);
	return key;
}

/* ---------------- hit.c ---------------- */
static inline void mm_cal_fuzzy_len(mm_reg1_t *r, const mm128_t *a)
{
	int i;
	r->mlen = r->blen = 0;
	if (r->cnt <= 0) return;
	r->mlen = r->blen = a[r->as].y>>32&0xff;
	for (i = r->as + 1; i < r->as + r->cnt; ++i) {
		int span = a[i].y>>32&0xff;
		int tl = (int32_t)a[i].x - (int32_t)a[i-1].x;
		int ql = (int32_t)a[i].y - (int32_t)a[i-1].y;
		r->blen += tl > ql? tl : ql;
		r->mlen += tl > span && ql > span? span : tl < ql? tl : ql;
	}
}

static inline void mm_reg_set_coor(mm_reg1_t *r, int32_t qlen, const mm128_t *a)
{
	int32_t k = r->as, q_span = (int32_t)(a[k].y>>32&0xff);
	r->rev = a[k].x>>63;
	r->rid = a[k].x<<1>>33;
	r->rs = (int32_t)a[k].x + 1 > q_span? (int32_t)a[k].x + 1 - q_span : 0;
	r->re = (int32_t)a[k + r->cnt - 1].x + 1;
	if (!r->rev) {
		r->qs = (int32_t)a[k].y + 1 - q_span;
		r->qe = (int32_t)a[k + r->cnt - 1].y + 1;
	} else {
		r->qs = qlen - ((int32_t)a[k + r->cnt - 1].y + 1);
		r->qe = qlen - ((int32_t)a[k].y + 1 - q_span);
	}
	mm_cal_fuzzy_len(r, a);
}

static mm_reg1_t *mm_gen_regs(uint32_t hash, int qlen, int n_u, uint64_t *u, mm128_t *a)
{
	mm128_t *z, tmp;
	mm_reg1_t *r;
	int i, k;
	if (n_u == 0) return 0;
	z = (mm128_t*)malloc(n_u * 16);
	for (i = k = 0; i < n_u; ++i) {
		uint32_t h;
		h = (uint32_t)hash64((hash64(a[k].x) + hash64(a[k].y)) ^ hash);
		z[i].x = u[i] ^ h;
		z[i].y = (uint64_t)k << 32 | (int32_t)u[i];
		k += (int32_t)u[i];
	}
	mm2o_radix_sort_128x(z, z + n_u);
	for (i = 0; i < n_u>>1; ++i)
		tmp = z[i], z[i] = z[n_u-1-i], z[n_u-1-i] = tmp;
	r = (mm_reg1_t*)calloc(n_u, sizeof(mm_reg1_t));
	for (i = 0; i < n_u; ++i) {
		mm_reg1_t *ri = &r[i];
		ri->id = i;
		ri->parent = MM_PARENT_UNSET;
		ri->score = ri->score0 = z[i].x >> 32;
		ri->hash = (uint32_t)z[i].x;
		ri->cnt = (int32_t)z[i].y;
		ri->as = z[i].y >> 32;
		ri->div = -1.0f;
		mm_reg_set_coor(ri, qlen, a);
	}
	free(z);
	return r;
}

static int mm_set_sam_pri(int n, mm_reg1_t *r)
{
	int i, n_pri = 0;
	for (i = 0; i < n; ++i)
		if (r[i].id == r[i].parent) {
			++n_pri;
			r[i].sam_pri = (n_pri == 1);
		} else r[i].sam_pri = 0;
	return n_pri;
}

static void mm_sync_regs(int n_regs, mm_reg1_t *regs)
{
	int *tmp, i, max_id = -1, n_tmp;
	if (n_regs <= 0) return;
	for (i = 0; i < n_regs; ++i)
		max_id = max_id > regs[i].id? max_id : regs[i].id;
	n_tmp = max_id + 1;
	tmp = (int*)malloc((n_tmp + 1) * sizeof(int));
	for (i = 0; i < n_tmp; ++i) tmp[i] = -1;
	for (i = 0; i < n_regs; ++i)
		if (regs[i].id >= 0) tmp[regs[i].id] = i;
	for (i = 0; i < n_regs; ++i) {
		mm_reg1_t *r = &regs[i];
		r->id = i;
		if (r->parent == MM_PARENT_TMP_PRI)
			r->parent = i;
		else if (r->parent >= 0 && tmp[r->parent] >= 0)
			r->parent = tmp[r->parent];
		else r->parent = MM_PARENT_UNSET;
	}
	free(tmp);
	mm_set_sam_pri(n_regs, regs);
}

static void mm_set_parent(float mask_level, int n, mm_reg1_t *r, int sub_diff)
{
	int i, j, k, *w;
	uint64_t *cov;
	if (n <= 0) return;
	for (i = 0; i < n; ++i) r[i].id = i;
	cov = (uint64_t*)malloc(n * sizeof(uint64_t));
	w = (int*)malloc(n * sizeof(int));
	w[0] = 0, r[0].parent = 0;
	for (i = 1, k = 1; i < n; ++i) {
		mm_reg1_t *ri = &r[i];
		int si = ri->qs, ei = ri->qe, n_cov = 0, uncov_len = 0;
		for (j = 0; j < k; ++j) { /* traverse existing primary hits to find overlapping hits */
			mm_reg1_t *rp = &r[w[j]];
			int sj = rp->qs, ej = rp->qe;
			if (ej <= si || sj >= ei) continue;
			if (sj < si) sj = si;
			if (ej > ei) ej = ei;
			cov[n_cov++] = (uint64_t)sj<<32 | ej;
		}
		if (n_cov == 0) {
			goto set_parent_test; /* no overlapping primary hits; then i is a new primary hit */
		} else if (n_cov > 0) {
			int jj, x = si;
			mm2o_radix_sort_64(cov, cov + n_cov);
			for (jj = 0; jj < n_cov; ++jj) {
				if ((int)(cov[jj]>>32) > x) uncov_len += (cov[jj]>>32) - x;
				x = (int32_t)cov[jj] > x? (int32_t)cov[jj] : x;
			}
			if (ei > x) uncov_len += ei - x;
		}
		for (j = 0; j < k; ++j) { /* traverse existing primary hits again */
			mm_reg1_t *rp = &r[w[j]];
			int sj = rp->qs, ej = rp->qe, min, max, ol;
			if (ej <= si || sj >= ei) continue;
			min = ej - sj < ei - si? ej - sj : ei - si;
			max = ej - sj > ei - si? ej - sj : ei - si;
			ol = si < sj? (ei < sj? 0 : ei < ej? ei - sj : ej - sj) : (ej < si? 0 : ej < ei? ej - si : ei - si);
			if ((float)ol / min - (float)uncov_len / max > mask_level) {
				int cnt_sub = 0;
				ri->parent = rp->parent;
				rp->subsc = rp->subsc > ri->score? rp->subsc : ri->score;
				if (ri->cnt >= rp->cnt) cnt_sub = 1;
				if (rp->p && ri->p && (rp->rid != ri->rid || rp->rs != ri->rs || rp->re != ri->re || ol != min)) {
					int sci = ri->p->dp_max;
					rp->p->dp_max2 = rp->p->dp_max2 > sci? rp->p->dp_max2 : sci;
					if (rp->p->dp_max - ri->p->dp_max <= sub_diff) cnt_sub = 1;
				}
				if (cnt_sub) ++rp->n_sub;
				break;
			}
		}
set_parent_test:
		if (j == k) w[k++] = i, ri->parent = i, ri->n_sub = 0;
	}
	free(cov);
	free(w);
}

static void reg_free_p(mm_reg1_t *r)
{
	if (r->p) { free(r->p->cigar); free(r->p); r->p = 0; }
}

static void mm_select_sub(float pri_ratio, int min_diff, int best_n, int *n_, mm_reg1_t *r)
{
	if (pri_ratio > 0.0f && *n_ > 0) {
		int i, k, n = *n_, n_2nd = 0;
		for (i = k = 0; i < n; ++i) {
			int p = r[i].parent;
			if (p == i || r[i].inv) {
				r[k++] = r[i];
			} else if ((r[i].score >= r[p].score * pri_ratio || r[i].score + min_diff >= r[p].score) && n_2nd < best_n) {
				if (!(r[i].qs == r[p].qs && r[i].qe == r[p].qe && r[i].rid == r[p].rid && r[i].rs == r[p].rs && r[i].re == r[p].re))
					r[k++] = r[i], ++n_2nd;
				else reg_free_p(&r[i]);
			} else reg_free_p(&r[i]);
		}
		if (k != n) mm_sync_regs(k, r);
		*n_ = k;
	}
}

static void mm_filter_regs(const mm2o_opt_t *opt, int qlen, int *n_regs, mm_reg1_t *regs)
{
	int i, k;
	for (i = k = 0; i < *n_regs; ++i) {
		mm_reg1_t *r = &regs[i];
		int flt = 0;
		if (!r->inv && !r->seg_split && r->cnt < opt->min_cnt) flt = 1;
		if (r->p) {
			if (r->mlen < opt->min_chain_score) flt = 1;
			else if (r->p->dp_max < opt->min_dp_max) flt = 1;
			else if (r->qs > qlen * opt->max_clip_ratio && qlen - r->qe > qlen * opt->max_clip_ratio) flt = 1;
			if (flt) reg_free_p(r);
		}
		if (flt) continue;
		if (k < i) regs[k++] = regs[i];
		else ++k;
	}
	*n_regs = k;
}

static void mm_hit_sort(int *n_regs, mm_reg1_t *r)
{
	int32_t i, n_aux, n = *n_regs;
	mm128_t *aux;
	mm_reg1_t *t;
	if (n <= 1) return;
	aux = (mm128_t*)malloc(n * 16);
	t = (mm_reg1_t*)malloc(n * sizeof(mm_reg1_t));
	for (i = n_aux = 0; i < n; ++i) {
		if (r[i].inv || r[i].cnt > 0) {
			int score;
			if (r[i].p) score = r[i].p->dp_max;
			else score = r[i].score;
			aux[n_aux].x = (uint64_t)score << 32 | r[i].hash;
			aux[n_aux++].y = i;
		} else if (r[i].p) {
			reg_free_p(&r[i]);
		}
	}
	mm2o_radix_sort_128x(aux, aux + n_aux);
	for (i = n_aux - 1; i >= 0; --i)
		t[n_aux - 1 - i] = r[aux[i].y];
	memcpy(r, t, sizeof(mm_reg1_t) * n_aux);
	*n_regs = n_aux;
	free(aux);
	free(t);
}

static int mm_squeeze_a(int n_regs, mm_reg1_t *regs, mm128_t *a)
{
	int i, as = 0;
	uint64_t *aux;
	aux = (uint64_t*)malloc((n_regs + 1) * 8);
	for (i = 0; i < n_regs; ++i)
		aux[i] = (uint64_t)regs[i].as << 32 | i;
	mm2o_radix_sort_64(aux, aux + n_regs);
	for (i = 0; i < n_regs; ++i) {
		mm_reg1_t *r = &regs[(int32_t)aux[i]];
		if (r->as != as) {
			memmove(&a[as], &a[r->as], r->cnt * 16);
			r->as = as;
		}
		as += r->cnt;
	}
	free(aux);
	return as;
}

static void mm_join_long(const mm2o_opt_t *opt, int qlen, int *n_regs_, mm_reg1_t *regs, mm128_t *a)
{
	int i, n_aux, n_regs = *n_regs_, n_drop = 0;
	uint64_t *aux;
	if (n_regs < 2) return;
	mm_squeeze_a(n_regs, regs, a);
	aux = (uint64_t*)malloc(n_regs * 8);
	for (i = n_aux = 0; i < n_regs; ++i)
		if (regs[i].parent == i || regs[i].parent < 0)
			aux[n_aux++] = (uint64_t)regs[i].as << 32 | i;
	mm2o_radix_sort_64(aux, aux + n_aux);
	for (i = n_aux - 1; i >= 1; --i) {
		mm_reg1_t *r0 = &regs[(int32_t)aux[i-1]], *r1 = &regs[(int32_t)aux[i]];
		mm128_t *a0e, *a1s;
		int max_gap, min_gap, sc_thres, min_flank_len;
		if (r0->as + r0->cnt != r1->as) continue;
		if (r0->rid != r1->rid || r0->rev != r1->rev) continue;
		a0e = &a[r0->as + r0->cnt - 1];
		a1s = &a[r1->as];
		if (a1s->x <= a0e->x || (int32_t)a1s->y <= (int32_t)a0e->y) continue;
		max_gap = min_gap = (int32_t)a1s->y - (int32_t)a0e->y;
		max_gap = max_gap > (int64_t)(a1s->x - a0e->x)? max_gap : (int)(a1s->x - a0e->x);
		min_gap = min_gap < (int64_t)(a1s->x - a0e->x)? min_gap : (int)(a1s->x - a0e->x);
		if (max_gap > opt->max_join_long || min_gap > opt->max_join_short) continue;
		sc_thres = (int)((float)opt->min_join_flank_sc / opt->max_join_long * max_gap + .499);
		if (r0->score < sc_thres || r1->score < sc_thres) continue;
		min_flank_len = (int)(max_gap * opt->min_join_flank_ratio);
		if (r0->re - r0->rs < min_flank_len || r0->qe - r0->qs < min_flank_len) continue;
		if (r1->re - r1->rs < min_flank_len || r1->qe - r1->qs < min_flank_len) continue;
		a[r1->as].y |= MM_SEED_LONG_JOIN;
		r0->cnt += r1->cnt, r0->score += r1->score;
		mm_reg_set_coor(r0, qlen, a);
		r1->cnt = 0;
		r1->parent = r0->id;
		++n_drop;
	}
	free(aux);
	if (n_drop > 0) {
		for (i = 0; i < n_regs; ++i) {
			mm_reg1_t *r = &regs[i];
			if (r->parent >= 0 && r->id != r->parent) {
				if (regs[r->parent].parent >= 0 && regs[r->parent].parent != r->parent)
					r->parent = regs[r->parent].parent;
			}
		}
		mm_filter_regs(opt, qlen, n_regs_, regs);
		mm_sync_regs(*n_regs_, regs);
	}
}

static void mm_split_reg(mm_reg1_t *r, mm_reg1_t *r2, int n, int qlen, mm128_t *a)
{
	if (n <= 0 || n >= r->cnt) return;
	*r2 = *r;
	r2->id = -1;
	r2->sam_pri = 0;
	r2->p = 0;
	r2->split_inv = 0;
	r2->cnt = r->cnt - n;
	r2->score = (int32_t)(r->score * ((float)r2->cnt / r->cnt) + .499);
	r2->as = r->as + n;
	if (r->parent == r->id) r2->parent = MM_PARENT_TMP_PRI;
	mm_reg_set_coor(r2, qlen, a);
	r->cnt -= r2->cnt;
	r->score -= r2->score;
	mm_reg_set_coor(r, qlen, a);
	r->split |= 1, r2->split |= 2;
}

static void mm_set_inv_mapq(int n_regs, mm_reg1_t *regs);

static void mm_set_mapq(int n_regs, mm_reg1_t *regs, int min_chain_sc, int match_sc, int rep_len)
{
	static const float q_coef = 40.0f;
	int64_t sum_sc = 0;
	float uniq_ratio;
	int i;
	if (n_regs == 0) return;
	for (i = 0; i < n_regs; ++i)
		if (regs[i].parent == regs[i].id)
			sum_sc += regs[i].score;
	uniq_ratio = (float)sum_sc / (sum_sc + rep_len);
	for (i = 0; i < n_regs; ++i) {
		mm_reg1_t *r = &regs[i];
		if (r->inv) {
			r->mapq = 0;
		} else if (r->parent == r->id) {
			int mapq, subsc;
			float pen_s1 = (r->score > 100? 1.0f : 0.01f * r->score) * uniq_ratio;
			float pen_cm = r->cnt > 10? 1.0f : 0.1f * r->cnt;
			pen_cm = pen_s1 < pen_cm? pen_s1 : pen_cm;
			subsc = r->subsc > min_chain_sc? r->subsc : min_chain_sc;
			if (r->p && r->p->dp_max2 > 0 && r->p->dp_max > 0) {
				float identity = (float)r->mlen / r->blen;
				float x = (float)r->p->dp_max2 * subsc / r->p->dp_max / r->score0;
				mapq = (int)(identity * pen_cm * q_coef * (1.0f - x * x) * logf((float)r->p->dp_max / match_sc));
				{
					int mapq_alt = (int)(6.02f * identity * identity * (r->p->dp_max - r->p->dp_max2) / match_sc + .499f);
					mapq = mapq < mapq_alt? mapq : mapq_alt;
				}
			} else {
				float x = (float)subsc / r->score0;
				if (r->p) {
					float identity = (float)r->mlen / r->blen;
					mapq = (int)(identity * pen_cm * q_coef * (1.0f - x) * logf((float)r->p->dp_max / match_sc));
				} else {
					mapq = (int)(pen_cm * q_coef * (1.0f - x) * logf(r->score));
				}
			}
			mapq -= (int)(4.343f * logf(r->n_sub + 1) + .499f);
			mapq = mapq > 0? mapq : 0;
			r->mapq = mapq < 60? mapq : 60;
			if (r->p && r->p->dp_max > r->p->dp_max2 && r->mapq == 0) r->mapq = 1;
		} else r->mapq = 0;
	}
	mm_set_inv_mapq(n_regs, regs);
}

/* hit.c mm_set_inv_mapq(): an inversion hit takes the smaller MAPQ of its two neighbours in reference order */
static void mm_set_inv_mapq(int n_regs, mm_reg1_t *regs)
{
	int i, n_aux;
	mm128_t *aux;
	if (n_regs < 3) return;
	for (i = 0; i < n_regs; ++i)
		if (regs[i].inv) break;
	if (i == n_regs) return; /* no inversion hits */
	aux = (mm128_t*)malloc(n_regs * 16);
	for (i = n_aux = 0; i < n_regs; ++i)
		if (regs[i].parent == i || regs[i].parent < 0)
			aux[n_aux].y = i, aux[n_aux++].x = (uint64_t)regs[i].rid << 32 | regs[i].rs;
	mm2o_radix_sort_128x(aux, aux + n_aux);
	for (i = 1; i < n_aux - 1; ++i) {
		mm_reg1_t *inv = &regs[aux[i].y];
		if (inv->inv) {
			mm_reg1_t *l = &regs[aux[i-1].y];
			mm_reg1_t *r = &regs[aux[i+1].y];
			inv->mapq = l->mapq < r->mapq? l->mapq : r->mapq;
		}
	}
	free(aux);
}

/* ---------------- align.c ---------------- */
typedef struct {
	const mm2o_opt_t *opt;
	const mm2o_idx_t *mi;
	mm2o_trace_t *trace;
	int64_t dp_cells, n_dp_calls;
} actx_t;

static inline void mm_seq_rev(uint32_t len, uint8_t *seq)
{
	uint32_t i;
	uint8_t t;
	for (i = 0; i < len>>1; ++i)
		t = seq[i], seq[i] = seq[len - 1 - i], seq[len - 1 - i] = t;
}

static void mm_fix_cigar(mm_reg1_t *r, const uint8_t *qseq, const uint8_t *tseq, int *qshift, int *tshift)
{
	mm_extra_t *p = r->p;
	int32_t toff = 0, qoff = 0, to_shrink = 0;
	uint32_t k;
	*qshift = *tshift = 0;
	if (p->n_cigar <= 1) return;
	for (k = 0; k < p->n_cigar; ++k) { /* indel left alignment */
		uint32_t op = p->cigar[k]&0xf, len = p->cigar[k]>>4;
		if (len == 0) to_shrink = 1;
		if (op == 0) {
			toff += len, qoff += len;
		} else if (op == 1 || op == 2) {
			if (k > 0 && k < p->n_cigar - 1 && (p->cigar[k-1]&0xf) == 0 && (p->cigar[k+1]&0xf) == 0) {
				int l, prev_len = p->cigar[k-1] >> 4;
				if (op == 1) {
					for (l = 0; l < prev_len; ++l)
						if (qseq[qoff - 1 - l] != qseq[qoff + len - 1 - l])
							break;
				} else {
					for (l = 0; l < prev_len; ++l)
						if (tseq[toff - 1 - l] != tseq[toff + len - 1 - l])
							break;
				}
				if (l > 0)
					p->cigar[k-1] -= l<<4, p->cigar[k+1] += l<<4, qoff -= l, toff -= l;
				if (l == prev_len) to_shrink = 1;
			}
			if (op == 1) qoff += len;
			else toff += len;
		} else if (op == 3) {
			toff += len;
		}
	}
	assert(qoff == r->qe - r->qs && toff == r->re - r->rs);
	if (to_shrink) { /* squeeze out zero-length operations */
		int32_t l = 0;
		for (k = 0; k < p->n_cigar; ++k)
			if (p->cigar[k]>>4 != 0)
				p->cigar[l++] = p->cigar[k];
		p->n_cigar = l;
		for (k = l = 0; k < p->n_cigar; ++k) /* merge two adjacent operations if they are the same */
			if (k == p->n_cigar - 1 || (p->cigar[k]&0xf) != (p->cigar[k+1]&0xf))
				p->cigar[l++] = p->cigar[k];
			else p->cigar[k+1] += p->cigar[k]>>4<<4;
		p->n_cigar = l;
	}
	if ((p->cigar[0]&0xf) == 1 || (p->cigar[0]&0xf) == 2) { /* get rid of leading I or D */
		int32_t l = p->cigar[0] >> 4;
		if ((p->cigar[0]&0xf) == 1) {
			if (r->rev) r->qe -= l;
			else r->qs += l;
			*qshift = l;
		} else r->rs += l, *tshift = l;
		--p->n_cigar;
		memmove(p->cigar, p->cigar + 1, p->n_cigar * 4);
	}
}

static void mm_update_extra(mm_reg1_t *r, const uint8_t *qseq, const uint8_t *tseq, const int8_t *mat, int8_t q, int8_t e)
{
	uint32_t k, l;
	int32_t s = 0, max = 0, qshift, tshift, toff = 0, qoff = 0;
	mm_extra_t *p = r->p;
	if (p == 0) return;
	mm_fix_cigar(r, qseq, tseq, &qshift, &tshift);
	qseq += qshift, tseq += tshift;
	r->blen = r->mlen = 0;
	for (k = 0; k < p->n_cigar; ++k) {
		uint32_t op = p->cigar[k]&0xf, len = p->cigar[k]>>4;
		if (op == 0) {
			int n_ambi = 0, n_diff = 0;
			for (l = 0; l < len; ++l) {
				int cq = qseq[qoff + l], ct = tseq[toff + l];
				if (ct > 3 || cq > 3) ++n_ambi;
				else if (ct != cq) ++n_diff;
				s += mat[ct * 5 + cq];
				if (s < 0) s = 0;
				else max = max > s? max : s;
			}
			r->blen += len - n_ambi, r->mlen += len - (n_ambi + n_diff), p->n_ambi += n_ambi;
			toff += len, qoff += len;
		} else if (op == 1) {
			int n_ambi = 0;
			for (l = 0; l < len; ++l)
				if (qseq[qoff + l] > 3) ++n_ambi;
			r->blen += len - n_ambi, p->n_ambi += n_ambi;
			s -= q + e * len;
			if (s < 0) s = 0;
			qoff += len;
		} else if (op == 2) {
			int n_ambi = 0;
			for (l = 0; l < len; ++l)
				if (tseq[toff + l] > 3) ++n_ambi;
			r->blen += len - n_ambi, p->n_ambi += n_ambi;
			s -= q + e * len;
			if (s < 0) s = 0;
			toff += len;
		} else if (op == 3) {
			toff += len;
		}
	}
	p->dp_max = max;
	assert(qoff == r->qe - r->qs && toff == r->re - r->rs);
}

static void mm_append_cigar(mm_reg1_t *r, uint32_t n_cigar, uint32_t *cigar)
{
	mm_extra_t *p;
	if (n_cigar == 0) return;
	if (r->p == 0) {
		r->p = (mm_extra_t*)calloc(1, sizeof(mm_extra_t));
		r->p->capacity = n_cigar + 16;
		r->p->cigar = (uint32_t*)malloc(r->p->capacity * 4);
	} else if (r->p->n_cigar + n_cigar > r->p->capacity) {
		r->p->capacity = r->p->n_cigar + n_cigar;
		r->p->capacity += (r->p->capacity >> 1) + 16;
		r->p->cigar = (uint32_t*)realloc(r->p->cigar, r->p->capacity * 4);
	}
	p = r->p;
	if (p->n_cigar > 0 && (p->cigar[p->n_cigar-1]&0xf) == (cigar[0]&0xf)) { /* same CIGAR op at the boundary */
		p->cigar[p->n_cigar-1] += (cigar[0]>>4)<<4;
		if (n_cigar > 1) memcpy(p->cigar + p->n_cigar, cigar + 1, (n_cigar - 1) * 4);
		p->n_cigar += n_cigar - 1;
	} else {
		memcpy(p->cigar + p->n_cigar, cigar, n_cigar * 4);
		p->n_cigar += n_cigar;
	}
}

static void trace_dp(actx_t *c, int qlen, const uint8_t *qseq, int tlen, const uint8_t *tseq, int w, int end_bonus, int zdrop, int flag, const ksw_extz_t *ez)
{
	mm2o_trace_t *t = c->trace;
	mm2o_dp_rec_t *d;
	if (t == 0 || !t->enabled) return;
	if (t->n_dp == t->m_dp) {
		t->m_dp = t->m_dp? t->m_dp << 1 : 64;
		t->dp = (mm2o_dp_rec_t*)realloc(t->dp, t->m_dp * sizeof(mm2o_dp_rec_t));
	}
	if (t->n_seq + qlen + tlen > t->m_seq) {
		t->m_seq = (t->n_seq + qlen + tlen) * 2 + 1024;
		t->seqpool = (uint8_t*)realloc(t->seqpool, t->m_seq);
	}
	if (t->n_cig + ez->n_cigar > t->m_cig) {
		t->m_cig = (t->n_cig + ez->n_cigar) * 2 + 1024;
		t->cigpool = (uint32_t*)realloc(t->cigpool, t->m_cig * 4);
	}
	d = &t->dp[t->n_dp++];
	d->qlen = qlen, d->tlen = tlen, d->w = w, d->zdrop = zdrop, d->end_bonus = end_bonus, d->flag = flag;
	d->q_off = t->n_seq; memcpy(t->seqpool + t->n_seq, qseq, qlen); t->n_seq += qlen;
	d->t_off = t->n_seq; memcpy(t->seqpool + t->n_seq, tseq, tlen); t->n_seq += tlen;
	d->score = ez->score, d->max = ez->max, d->max_q = ez->max_q, d->max_t = ez->max_t, d->mqe = ez->mqe, d->mqe_t = ez->mqe_t;
	d->zdropped = ez->zdropped, d->reach_end = ez->reach_end, d->n_cigar = ez->n_cigar;
	d->cigar_off = t->n_cig;
	if (ez->n_cigar > 0) memcpy(t->cigpool + t->n_cig, ez->cigar, ez->n_cigar * 4);
	t->n_cig += ez->n_cigar;
}

static void mm_align_pair(actx_t *c, int qlen, const uint8_t *qseq, int tlen, const uint8_t *tseq, const int8_t *mat, int w, int end_bonus, int zdrop, int flag, ksw_extz_t *ez)
{
	const mm2o_opt_t *opt = c->opt;
	if (opt->max_sw_mat > 0 && (int64_t)tlen * qlen > opt->max_sw_mat) {
		ez->max_q = ez->max_t = ez->mqe_t = ez->mte_q = -1;
		ez->max = 0, ez->score = ez->mqe = ez->mte = KSW_NEG_INF;
		ez->n_cigar = 0, ez->zdropped = 0, ez->reach_end = 0;
		ez->zdropped = 1;
	} else {
		/* q != q2 for map-ont => upstream dispatches to ksw_extd2_sse */
		mm2o_ksw_extd2(qlen, qseq, tlen, tseq, 5, mat, opt->q, opt->e, opt->q2, opt->e2, w, zdrop, end_bonus, flag, ez);
		c->dp_cells += mm2o_ksw_cells(qlen, tlen, w);
		++c->n_dp_calls;
	}
	trace_dp(c, qlen, qseq, tlen, tseq, w, end_bonus, zdrop, flag, ez);
}

static inline void update_max_zdrop(int32_t score, int i, int j, int32_t *max, int *max_i, int *max_j, int e, int *max_zdrop, int pos[2][2])
{
	if (score < *max) {
		int li = i - *max_i;
		int lj = j - *max_j;
		int diff = li > lj? li - lj : lj - li;
		int z = *max - score - diff * e;
		if (z > *max_zdrop) {
			*max_zdrop = z;
			pos[0][0] = *max_i, pos[0][1] = *max_j;
			pos[1][0] = i, pos[1][1] = j;
		}
	} else *max = score, *max_i = i, *max_j = j;
}

/* ksw2_ll_sse.c ksw_ll_i16(): local (Smith-Waterman) single-affine score; scalar restatement returns the same max */
static int ksw_ll_score(int qlen, const uint8_t *query, int tlen, const uint8_t *target, const int8_t *mat, int gapo, int gape)
{
	int i, j, gmax = 0;
	int *H = (int*)calloc(qlen + 1, sizeof(int)), *E = (int*)calloc(qlen + 1, sizeof(int));
	for (i = 0; i < tlen; ++i) {
		int f = 0, h_diag = 0; /* H(i-1,j-1) */
		for (j = 0; j < qlen; ++j) {
			int h = h_diag + mat[target[i] * 5 + query[j]];
			int e = E[j + 1];
			h_diag = H[j + 1];
			h = h > e? h : e;
			h = h > f? h : f;
			h = h > 0? h : 0;
			H[j + 1] = h;
			gmax = gmax > h? gmax : h;
			h -= gapo + gape; if (h < 0) h = 0;
			e -= gape; e = e > h? e : h; E[j + 1] = e;
			f -= gape; f = f > h? f : h;
		}
	}
	free(H); free(E);
	return gmax;
}

static int mm_test_zdrop(const mm2o_opt_t *opt, const uint8_t *qseq, const uint8_t *tseq, uint32_t n_cigar, uint32_t *cigar, const int8_t *mat)
{
	uint32_t k;
	int32_t score = 0, max = INT32_MIN, max_i = -1, max_j = -1, i = 0, j = 0, max_zdrop = 0;
	int pos[2][2] = {{-1, -1}, {-1, -1}}, q_len, t_len;

	for (k = 0, score = 0; k < n_cigar; ++k) {
		uint32_t l, op = cigar[k]&0xf, len = cigar[k]>>4;
		if (op == 0) {
			for (l = 0; l < len; ++l) {
				score += mat[tseq[i + l] * 5 + qseq[j + l]];
				update_max_zdrop(score, i+l, j+l, &max, &max_i, &max_j, opt->e, &max_zdrop, pos);
			}
			i += len, j += len;
		} else if (op == 1 || op == 2 || op == 3) {
			score -= opt->q + opt->e * len;
			if (op == 1) j += len;
			else i += len;
			update_max_zdrop(score, i, j, &max, &max_i, &max_j, opt->e, &max_zdrop, pos);
		}
	}

	/* test if there is an inversion in the most dropped region */
	q_len = pos[1][1] - pos[0][1], t_len = pos[1][0] - pos[0][0];
	if (max_zdrop > opt->zdrop_inv && q_len < opt->max_gap && t_len < opt->max_gap) {
		uint8_t *qseq2;
		qseq2 = (uint8_t*)malloc(q_len + 1);
		for (i = 0; i < q_len; ++i) {
			int c = qseq[pos[1][1] - i - 1];
			qseq2[i] = c >= 4? 4 : 3 - c;
		}
		{
			int q_off, t_off; /* upstream passes them and ignores the values here */
			score = mm2o_ksw_ll_i16(q_len, qseq2, t_len, tseq + pos[0][0], 5, mat, opt->q, opt->e, &q_off, &t_off);
			assert(score == ksw_ll_score(q_len, qseq2, t_len, tseq + pos[0][0], mat, opt->q, opt->e)); /* the plain recurrence gives the same maximum */
		}
		free(qseq2);
		if (score >= opt->min_chain_score * opt->a && score >= opt->min_dp_max)
			return 2; /* there is a potential inversion */
	}
	return max_zdrop > opt->zdrop? 1 : 0;
}

static void mm_fix_bad_ends(const mm_reg1_t *r, const mm128_t *a, int bw, int min_match, int32_t *as, int32_t *cnt)
{
	int32_t i, l, m;
	*as = r->as, *cnt = r->cnt;
	if (r->cnt < 3) return;
	m = l = a[r->as].y >> 32 & 0xff;
	for (i = r->as + 1; i < r->as + r->cnt - 1; ++i) {
		int32_t lq, lr, min, max;
		int32_t q_span = a[i].y >> 32 & 0xff;
		if (a[i].y & MM_SEED_LONG_JOIN) break;
		lr = (int32_t)a[i].x - (int32_t)a[i-1].x;
		lq = (int32_t)a[i].y - (int32_t)a[i-1].y;
		min = lr < lq? lr : lq;
		max = lr > lq? lr : lq;
		if (max - min > l >> 1) *as = i;
		l += min;
		m += min < q_span? min : q_span;
		if (l >= bw << 1 || (m >= min_match && m >= bw) || m >= r->mlen >> 1) break;
	}
	*cnt = r->as + r->cnt - *as;
	m = l = a[r->as + r->cnt - 1].y >> 32 & 0xff;
	for (i = r->as + r->cnt - 2; i > *as; --i) {
		int32_t lq, lr, min, max;
		int32_t q_span = a[i+1].y >> 32 & 0xff;
		if (a[i+1].y & MM_SEED_LONG_JOIN) break;
		lr = (int32_t)a[i+1].x - (int32_t)a[i].x;
		lq = (int32_t)a[i+1].y - (int32_t)a[i].y;
		min = lr < lq? lr : lq;
		max = lr > lq? lr : lq;
		if (max - min > l >> 1) *cnt = i + 1 - *as;
		l += min;
		m += min < q_span? min : q_span;
		if (l >= bw << 1 || (m >= min_match && m >= bw) || m >= r->mlen >> 1) break;
	}
}

static int *collect_long_gaps(int as1, int cnt1, mm128_t *a, int min_gap, int *n_)
{
	int i, n, *K;
	*n_ = 0;
	for (i = 1, n = 0; i < cnt1; ++i) {
		int gap = ((int32_t)a[as1 + i].y - a[as1 + i - 1].y) - ((int32_t)a[as1 + i].x - a[as1 + i - 1].x);
		if (gap < -min_gap || gap > min_gap) ++n;
	}
	if (n <= 1) return 0;
	K = (int*)malloc(n * sizeof(int));
	for (i = 1, n = 0; i < cnt1; ++i) {
		int gap = ((int32_t)a[as1 + i].y - a[as1 + i - 1].y) - ((int32_t)a[as1 + i].x - a[as1 + i - 1].x);
		if (gap < -min_gap || gap > min_gap)
			K[n++] = i;
	}
	*n_ = n;
	return K;
}

static void mm_filter_bad_seeds(int as1, int cnt1, mm128_t *a, int min_gap, int diff_thres, int max_ext_len, int max_ext_cnt)
{
	int max_st, max_en, n, i, k, max, *K;
	K = collect_long_gaps(as1, cnt1, a, min_gap, &n);
	if (K == 0) return;
	max = 0, max_st = max_en = -1;
	for (k = 0;; ++k) {
		int gap, l, n_ins = 0, n_del = 0, qs, rs, max_diff = 0, max_diff_l = -1;
		if (k == n || k >= max_en) {
			if (max_en > 0)
				for (i = K[max_st]; i < K[max_en]; ++i)
					a[as1 + i].y |= MM_SEED_IGNORE;
			max = 0, max_st = max_en = -1;
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
			diff = n_ins + n_del - abs(n_ins - n_del);
			if (max_diff < diff)
				max_diff = diff, max_diff_l = l;
		}
		if (max_diff > diff_thres && max_diff > max)
			max = max_diff, max_st = k, max_en = max_diff_l;
	}
	free(K);
}

static void mm_filter_bad_seeds_alt(int as1, int cnt1, mm128_t *a, int min_gap, int max_ext)
{
	int n, k, *K;
	K = collect_long_gaps(as1, cnt1, a, min_gap, &n);
	if (K == 0) return;
	for (k = 0; k < n;) {
		int i = K[k], l;
		int gap1 = ((int32_t)a[as1 + i].y - (int32_t)a[as1 + i - 1].y) - ((int32_t)a[as1 + i].x - (int32_t)a[as1 + i - 1].x);
		int re1 = (int32_t)a[as1 + i].x;
		int qe1 = (int32_t)a[as1 + i].y;
		gap1 = gap1 > 0? gap1 : -gap1;
		for (l = k + 1; l < n; ++l) {
			int j = K[l], gap2, q_span_pre, rs2, qs2, m;
			if ((int32_t)a[as1 + j].y - qe1 > max_ext || (int32_t)a[as1 + j].x - re1 > max_ext) break;
			gap2 = ((int32_t)a[as1 + j].y - (int32_t)a[as1 + j - 1].y) - (int32_t)(a[as1 + j].x - a[as1 + j - 1].x);
			q_span_pre = a[as1 + j - 1].y >> 32 & 0xff;
			rs2 = (int32_t)a[as1 + j - 1].x + q_span_pre;
			qs2 = (int32_t)a[as1 + j - 1].y + q_span_pre;
			m = rs2 - re1 < qs2 - qe1? rs2 - re1 : qs2 - qe1;
			gap2 = gap2 > 0? gap2 : -gap2;
			if (m > gap1 + gap2) break;
			re1 = (int32_t)a[as1 + j].x;
			qe1 = (int32_t)a[as1 + j].y;
			gap1 = gap2;
		}
		if (l > k + 1) {
			int j, end = K[l - 1];
			for (j = K[k]; j < end; ++j)
				a[as1 + j].y |= MM_SEED_IGNORE;
			a[as1 + end].y |= MM_SEED_LONG_JOIN;
		}
		k = l;
	}
	free(K);
}

static inline void mm_adjust_minier(const mm2o_idx_t *mi, mm128_t *a, int32_t *r, int32_t *q)
{
	*r = (int32_t)a->x - (mi->k>>1);
	*q = (int32_t)a->y - (mi->k>>1);
}

static void mm_align1(actx_t *c, int qlen, uint8_t *qseq0[2], mm_reg1_t *r, mm_reg1_t *r2, int n_a, mm128_t *a, ksw_extz_t *ez)
{
	const mm2o_opt_t *opt = c->opt;
	const mm2o_idx_t *mi = c->mi;
	int32_t rid = a[r->as].x<<1>>33, rev = a[r->as].x>>63, as1, cnt1;
	uint8_t *tseq, *qseq;
	int32_t i, l, bw, dropped = 0, extra_flag = 0, rs0, re0, qs0, qe0;
	int32_t rs, re, qs, qe;
	int32_t rs1, qs1, re1, qe1;
	int8_t mat[25];

	r2->cnt = 0;
	if (r->cnt == 0) return;
	mm2o_gen_simple_mat(5, mat, opt->a, opt->b, opt->sc_ambi);
	bw = (int)(opt->bw * 1.5 + 1.);

	mm_fix_bad_ends(r, a, opt->bw, opt->min_chain_score * 2, &as1, &cnt1);
	mm_filter_bad_seeds(as1, cnt1, a, 10, 40, opt->max_gap>>1, 10);
	mm_filter_bad_seeds_alt(as1, cnt1, a, 30, opt->max_gap>>1);
	mm_adjust_minier(mi, &a[as1], &rs, &qs);
	mm_adjust_minier(mi, &a[as1 + cnt1 - 1], &re, &qe);
	assert(cnt1 > 0);

	/* compute rs0 and qs0 */
	rs0 = (int32_t)a[r->as].x + 1 - (int32_t)(a[r->as].y>>32&0xff);
	qs0 = (int32_t)a[r->as].y + 1 - (int32_t)(a[r->as].y>>32&0xff);
	if (rs0 < 0) rs0 = 0;
	assert(qs0 >= 0);
	rs1 = qs1 = 0;
	for (i = r->as - 1, l = 0; i >= 0 && a[i].x>>32 == a[r->as].x>>32; --i) { /* inspect nearby seeds */
		int32_t x = (int32_t)a[i].x + 1 - (int32_t)(a[i].y>>32&0xff);
		int32_t y = (int32_t)a[i].y + 1 - (int32_t)(a[i].y>>32&0xff);
		if (x < rs0 && y < qs0) {
			if (++l > opt->min_cnt) {
				l = rs0 - x > qs0 - y? rs0 - x : qs0 - y;
				rs1 = rs0 - l, qs1 = qs0 - l;
				if (rs1 < 0) rs1 = 0;
				break;
			}
		}
	}
	if (qs > 0 && rs > 0) {
		l = qs < opt->max_gap? qs : opt->max_gap;
		qs1 = qs1 > qs - l? qs1 : qs - l;
		qs0 = qs0 < qs1? qs0 : qs1; /* at least include qs0 */
		l += l * opt->a > opt->q? (l * opt->a - opt->q) / opt->e : 0;
		l = l < opt->max_gap? l : opt->max_gap;
		l = l < rs? l : rs;
		rs1 = rs1 > rs - l? rs1 : rs - l;
		rs0 = rs0 < rs1? rs0 : rs1;
		rs0 = rs0 < rs? rs0 : rs;
	} else rs0 = rs, qs0 = qs;
	/* compute re0 and qe0 */
	re0 = (int32_t)a[r->as + r->cnt - 1].x + 1;
	qe0 = (int32_t)a[r->as + r->cnt - 1].y + 1;
	re1 = mi->seq[rid].len, qe1 = qlen;
	for (i = r->as + r->cnt, l = 0; i < n_a && a[i].x>>32 == a[r->as].x>>32; ++i) { /* inspect nearby seeds */
		int32_t x = (int32_t)a[i].x + 1;
		int32_t y = (int32_t)a[i].y + 1;
		if (x > re0 && y > qe0) {
			if (++l > opt->min_cnt) {
				l = x - re0 > y - qe0? x - re0 : y - qe0;
				re1 = re0 + l, qe1 = qe0 + l;
				break;
			}
		}
	}
	if (qe < qlen && re < (int32_t)mi->seq[rid].len) {
		l = qlen - qe < opt->max_gap? qlen - qe : opt->max_gap;
		qe1 = qe1 < qe + l? qe1 : qe + l;
		qe0 = qe0 > qe1? qe0 : qe1; /* at least include qe0 */
		l += l * opt->a > opt->q? (l * opt->a - opt->q) / opt->e : 0;
		l = l < opt->max_gap? l : opt->max_gap;
		l = l < (int32_t)mi->seq[rid].len - re? l : (int32_t)mi->seq[rid].len - re;
		re1 = re1 < re + l? re1 : re + l;
		re0 = re0 > re1? re0 : re1;
	} else re0 = re, qe0 = qe;

	assert(re0 > rs0);
	tseq = (uint8_t*)malloc(re0 - rs0 + 16);

	if (qs > 0 && rs > 0) { /* left extension */
		qseq = &qseq0[rev][qs0];
		mm2o_idx_getseq(mi, rid, rs0, rs, tseq);
		mm_seq_rev(qs - qs0, qseq);
		mm_seq_rev(rs - rs0, tseq);
		mm_align_pair(c, qs - qs0, qseq, rs - rs0, tseq, mat, bw, opt->end_bonus, r->split_inv? opt->zdrop_inv : opt->zdrop, extra_flag|KSW_EZ_EXTZ_ONLY|KSW_EZ_RIGHT|KSW_EZ_REV_CIGAR, ez);
		if (ez->n_cigar > 0) {
			mm_append_cigar(r, ez->n_cigar, ez->cigar);
			r->p->dp_score += ez->max;
		}
		rs1 = rs - (ez->reach_end? ez->mqe_t + 1 : ez->max_t + 1);
		qs1 = qs - (ez->reach_end? qs - qs0 : ez->max_q + 1);
		mm_seq_rev(qs - qs0, qseq);
	} else rs1 = rs, qs1 = qs;
	re1 = rs, qe1 = qs;
	assert(qs1 >= 0 && rs1 >= 0);

	for (i = 1; i < cnt1; ++i) { /* gap filling */
		if ((a[as1+i].y & (MM_SEED_IGNORE|MM_SEED_TANDEM)) && i != cnt1 - 1) continue;
		mm_adjust_minier(mi, &a[as1 + i], &re, &qe);
		re1 = re, qe1 = qe;
		if (i == cnt1 - 1 || (a[as1+i].y&MM_SEED_LONG_JOIN) || (qe - qs >= opt->min_ksw_len && re - rs >= opt->min_ksw_len)) {
			int j, bw1 = bw, zdrop_code;
			if (a[as1+i].y & MM_SEED_LONG_JOIN)
				bw1 = qe - qs > re - rs? qe - qs : re - rs;
			qseq = &qseq0[rev][qs];
			mm2o_idx_getseq(mi, rid, rs, re, tseq);
			mm_align_pair(c, qe - qs, qseq, re - rs, tseq, mat, bw1, -1, opt->zdrop, extra_flag|KSW_EZ_APPROX_MAX, ez); /* first pass: approximate Z-drop */
			if ((zdrop_code = mm_test_zdrop(opt, qseq, tseq, ez->n_cigar, ez->cigar, mat)) != 0)
				mm_align_pair(c, qe - qs, qseq, re - rs, tseq, mat, bw1, -1, zdrop_code == 2? opt->zdrop_inv : opt->zdrop, extra_flag, ez); /* second pass: lift approximate */
			if (ez->n_cigar > 0)
				mm_append_cigar(r, ez->n_cigar, ez->cigar);
			if (ez->zdropped) { /* truncated by Z-drop */
				for (j = i - 1; j >= 0; --j)
					if ((int32_t)a[as1 + j].x <= rs + ez->max_t)
						break;
				dropped = 1;
				if (j < 0) j = 0;
				if (r->p) r->p->dp_score += ez->max; /* upstream dereferences unguarded; dp_score is not exported */
				re1 = rs + (ez->max_t + 1);
				qe1 = qs + (ez->max_q + 1);
				if (cnt1 - (j + 1) >= opt->min_cnt) {
					mm_split_reg(r, r2, as1 + j + 1 - r->as, qlen, a);
					if (zdrop_code == 2) r2->split_inv = 1;
				}
				break;
			} else if (r->p) r->p->dp_score += ez->score;
			rs = re, qs = qe;
		}
	}

	if (!dropped && qe < qe0 && re < re0) { /* right extension */
		qseq = &qseq0[rev][qe];
		mm2o_idx_getseq(mi, rid, re, re0, tseq);
		mm_align_pair(c, qe0 - qe, qseq, re0 - re, tseq, mat, bw, opt->end_bonus, opt->zdrop, extra_flag|KSW_EZ_EXTZ_ONLY, ez);
		if (ez->n_cigar > 0) {
			mm_append_cigar(r, ez->n_cigar, ez->cigar);
			r->p->dp_score += ez->max;
		}
		re1 = re + (ez->reach_end? ez->mqe_t + 1 : ez->max_t + 1);
		qe1 = qe + (ez->reach_end? qe0 - qe : ez->max_q + 1);
	}
	assert(qe1 <= qlen);

	r->rs = rs1, r->re = re1;
	if (rev) r->qs = qlen - qe1, r->qe = qlen - qs1;
	else r->qs = qs1, r->qe = qe1;

	assert(re1 - rs1 <= re0 - rs0);
	if (r->p) {
		mm2o_idx_getseq(mi, rid, rs1, re1, tseq);
		mm_update_extra(r, &qseq0[r->rev][qs1], tseq, mat, opt->q, opt->e);
	}
	free(tseq);
}

/* align.c mm_align1_inv(): after a Z-drop split flagged as a potential inversion (mm_test_zdrop code 2), align the query
 * segment between the two halves, taken from the OPPOSITE strand, to the reference segment between them: a local alignment
 * of the reversed sequences finds where the inverted block starts (ksw_ll_i16), an extension alignment from there gives the
 * inv region.  r1 = the region before, r2 = the split-off region (both already aligned). */
static int mm_align1_inv(actx_t *c, int qlen, uint8_t *qseq0[2], const mm_reg1_t *r1, const mm_reg1_t *r2, mm_reg1_t *r_inv, ksw_extz_t *ez)
{
	const mm2o_opt_t *opt = c->opt;
	const mm2o_idx_t *mi = c->mi;
	int tl, ql, score, ret = 0, q_off, t_off;
	uint8_t *tseq, *qseq;
	int8_t mat[25];

	memset(r_inv, 0, sizeof(mm_reg1_t));
	if (!(r1->split&1) || !(r2->split&2)) return 0;
	if (r1->id != r1->parent && r1->parent != MM_PARENT_TMP_PRI) return 0;
	if (r2->id != r2->parent && r2->parent != MM_PARENT_TMP_PRI) return 0;
	if (r1->rid != r2->rid || r1->rev != r2->rev) return 0;
	ql = r1->rev? r1->qs - r2->qe : r2->qs - r1->qe;
	tl = r2->rs - r1->re;
	if (ql < opt->min_chain_score || ql > opt->max_gap) return 0;
	if (tl < opt->min_chain_score || tl > opt->max_gap) return 0;

	mm2o_gen_simple_mat(5, mat, opt->a, opt->b, opt->sc_ambi);
	tseq = (uint8_t*)malloc(tl);
	mm2o_idx_getseq(mi, r1->rid, r1->re, r2->rs, tseq);
	qseq = r1->rev? &qseq0[0][r2->qe] : &qseq0[1][qlen - r2->qs];

	mm_seq_rev(ql, qseq);
	mm_seq_rev(tl, tseq);
	score = mm2o_ksw_ll_i16(ql, qseq, tl, tseq, 5, mat, opt->q, opt->e, &q_off, &t_off);
	mm_seq_rev(ql, qseq);
	mm_seq_rev(tl, tseq);
	if (score < opt->min_dp_max) goto end_align1_inv;
	q_off = ql - (q_off + 1), t_off = tl - (t_off + 1); /* q_off can be -1..-7: see mm2o_ksw_ll_i16 on the padded columns */
	mm_align_pair(c, ql - q_off, qseq + q_off, tl - t_off, tseq + t_off, mat, (int)(opt->bw * 1.5), -1, opt->zdrop, KSW_EZ_EXTZ_ONLY, ez);
	if (ez->n_cigar == 0) goto end_align1_inv; /* should never be here */
	mm_append_cigar(r_inv, ez->n_cigar, ez->cigar);
	r_inv->p->dp_score = ez->max;
	r_inv->id = -1;
	r_inv->parent = MM_PARENT_UNSET;
	r_inv->inv = 1;
	r_inv->rev = !r1->rev;
	r_inv->rid = r1->rid;
	r_inv->div = -1.0f;
	if (r_inv->rev == 0) {
		r_inv->qs = r2->qe + q_off;
		r_inv->qe = r_inv->qs + ez->max_q + 1;
	} else {
		r_inv->qe = r2->qs - q_off;
		r_inv->qs = r_inv->qe - (ez->max_q + 1);
	}
	r_inv->rs = r1->re + t_off;
	r_inv->re = r_inv->rs + ez->max_t + 1;
	mm_update_extra(r_inv, &qseq[q_off], &tseq[t_off], mat, opt->q, opt->e);
	ret = 1;
end_align1_inv:
	free(tseq);
	return ret;
}

static mm_reg1_t *mm_insert_reg(const mm_reg1_t *r, int i, int *n_regs, mm_reg1_t *regs)
{
	regs = (mm_reg1_t*)realloc(regs, (*n_regs + 1) * sizeof(mm_reg1_t));
	if (i + 1 != *n_regs)
		memmove(&regs[i + 2], &regs[i + 1], sizeof(mm_reg1_t) * (*n_regs - i - 1));
	regs[i + 1] = *r;
	++*n_regs;
	return regs;
}

static mm_reg1_t *mm_align_skeleton(actx_t *c, int qlen, const char *qstr, int *n_regs_, mm_reg1_t *regs, mm128_t *a)
{
	const unsigned char *nt4 = mm2o_nt4_table();
	int32_t i, n_regs = *n_regs_, n_a;
	uint8_t *qseq0[2];
	ksw_extz_t ez;

	qseq0[0] = (uint8_t*)malloc(qlen * 2 + 2);
	qseq0[1] = qseq0[0] + qlen;
	for (i = 0; i < qlen; ++i) {
		qseq0[0][i] = nt4[(uint8_t)qstr[i]];
		qseq0[1][qlen - 1 - i] = qseq0[0][i] < 4? 3 - qseq0[0][i] : 4;
	}
	n_a = mm_squeeze_a(n_regs, regs, a);
	memset(&ez, 0, sizeof(ksw_extz_t));
	for (i = 0; i < n_regs; ++i) {
		mm_reg1_t r2;
		mm_align1(c, qlen, qseq0, &regs[i], &r2, n_a, a, &ez);
		if (r2.cnt > 0) regs = mm_insert_reg(&r2, i, &n_regs, regs);
		if (i > 0 && regs[i].split_inv) {
			if (mm_align1_inv(c, qlen, qseq0, &regs[i-1], &regs[i], &r2, &ez)) {
				regs = mm_insert_reg(&r2, i, &n_regs, regs);
				++i; /* skip the inserted INV alignment */
			}
		}
	}
	*n_regs_ = n_regs;
	free(qseq0[0]);
	free(ez.cigar);
	mm_filter_regs(c->opt, qlen, n_regs_, regs);
	mm_hit_sort(n_regs_, regs);
	return regs;
}

/* ---------------- map.c mm_map_frag() for one segment ---------------- */
mm2o_result_t *mm2o_map(const mm2o_idx_t *mi, const mm2o_opt_t *opt, const char *seq, int qlen, mm2o_trace_t *trace)
{
	mm2o_result_t *res = (mm2o_result_t*)calloc(1, sizeof(mm2o_result_t));
	int i, rep_len = 0, n_regs0 = 0, max_chain_gap_qry, max_chain_gap_ref;
	uint32_t hash;
	int64_t n_a = 0;
	uint64_t *u = 0;
	mm128_t *a;
	mm128_v mv = {0,0,0};
	mm_reg1_t *regs0;
	actx_t ctx;

	if (qlen <= 0) return res;
	hash  = 0; /* mappy passes qname = NULL */
	hash ^= wang_hash32(qlen) + wang_hash32(opt->seed);
	hash  = wang_hash32(hash);

	mm2o_sketch(seq, qlen, mi->w, mi->k, 0, &mv);
	a = mm2o_collect_seed_hits(opt, opt->mid_occ, mi, &mv, qlen, &n_a, &rep_len);
	res->n_mini = mv.n, res->n_anchor = n_a, res->rep_len = rep_len;
	if (trace && trace->enabled) {
		trace->mini.n = trace->mini.m = mv.n;
		trace->mini.a = (mm128_t*)malloc((mv.n + 1) * 16); if (mv.n) memcpy(trace->mini.a, mv.a, mv.n * 16);
		trace->anchors.n = trace->anchors.m = n_a;
		trace->anchors.a = (mm128_t*)malloc((n_a + 1) * 16); if (n_a) memcpy(trace->anchors.a, a, n_a * 16);
	}
	free(mv.a);

	max_chain_gap_qry = opt->max_gap;
	max_chain_gap_ref = opt->max_gap_ref > 0? opt->max_gap_ref : opt->max_gap;

	a = mm2o_chain_dp(max_chain_gap_ref, max_chain_gap_qry, opt->bw, opt->max_chain_skip, opt->max_chain_iter, opt->min_cnt, opt->min_chain_score,
					  n_a, a, &n_regs0, &u, trace, &res->chain_cells);
	if (trace && trace->enabled && a) {
		int64_t tot = 0;
		for (i = 0; i < n_regs0; ++i) tot += (int32_t)u[i];
		trace->n_u = n_regs0;
		trace->u = (uint64_t*)malloc((n_regs0 + 1) * 8); if (n_regs0) memcpy(trace->u, u, n_regs0 * 8);
		trace->chained.n = trace->chained.m = tot;
		trace->chained.a = (mm128_t*)malloc((tot + 1) * 16); if (tot) memcpy(trace->chained.a, a, tot * 16);
	}

	regs0 = mm_gen_regs(hash, qlen, n_regs0, u, a);
	/* chain_post() */
	mm_set_parent(opt->mask_level, n_regs0, regs0, opt->a * 2 + opt->b);
	mm_select_sub(opt->pri_ratio, mi->k*2, opt->best_n, &n_regs0, regs0);
	mm_join_long(opt, qlen, &n_regs0, regs0, a);
	/* mm_est_err(): sets div only; not restated */

	/* align_regs(): mappy always sets MM_F_CIGAR */
	ctx.opt = opt, ctx.mi = mi, ctx.trace = trace, ctx.dp_cells = 0, ctx.n_dp_calls = 0;
	if (n_regs0 > 0) {
		regs0 = mm_align_skeleton(&ctx, qlen, seq, &n_regs0, regs0, a);
		mm_set_parent(opt->mask_level, n_regs0, regs0, opt->a * 2 + opt->b);
		mm_select_sub(opt->pri_ratio, mi->k*2, opt->best_n, &n_regs0, regs0);
		mm_set_sam_pri(n_regs0, regs0);
	}
	mm_set_mapq(n_regs0, regs0, opt->min_chain_score, opt->a, rep_len);
	res->dp_cells = ctx.dp_cells, res->n_dp_calls = ctx.n_dp_calls;

	/* python/cmappy.h mm_reg2hitpy() */
	res->n_hits = n_regs0;
	res->hits = (mm2o_hit_t*)calloc(n_regs0 + 1, sizeof(mm2o_hit_t));
	for (i = 0; i < n_regs0; ++i) res->n_cigar_pool += regs0[i].p? regs0[i].p->n_cigar : 0;
	res->cigar_pool = (uint32_t*)malloc((res->n_cigar_pool + 1) * 4);
	res->n_cigar_pool = 0;
	for (i = 0; i < n_regs0; ++i) {
		mm_reg1_t *r = &regs0[i];
		mm2o_hit_t *h = &res->hits[i];
		h->rid = r->rid, h->rev = r->rev;
		h->qs = r->qs, h->qe = r->qe, h->rs = r->rs, h->re = r->re;
		h->mapq = r->mapq, h->mlen = r->mlen, h->blen = r->blen;
		h->nm = r->blen - r->mlen + (r->p? (int)r->p->n_ambi : 0);
		h->dp_max = r->p? r->p->dp_max : 0, h->dp_max2 = r->p? r->p->dp_max2 : 0;
		h->score = r->score, h->score0 = r->score0, h->cnt = r->cnt, h->subsc = r->subsc, h->n_sub = r->n_sub;
		h->id = r->id, h->parent = r->parent, h->is_primary = (r->id == r->parent), h->sam_pri = r->sam_pri;
		h->n_cigar = r->p? r->p->n_cigar : 0;
		h->cigar_off = res->n_cigar_pool;
		if (r->p) {
			memcpy(res->cigar_pool + res->n_cigar_pool, r->p->cigar, r->p->n_cigar * 4);
			res->n_cigar_pool += r->p->n_cigar;
			reg_free_p(r);
		}
	}
	free(regs0);
	free(a);
	free(u);
	return res;
}

void mm2o_result_destroy(mm2o_result_t *r)
{
	if (r == 0) return;
	free(r->hits); free(r->cigar_pool); free(r);
}

mm2o_trace_t *mm2o_trace_new(void)
{
	mm2o_trace_t *t = (mm2o_trace_t*)calloc(1, sizeof(mm2o_trace_t));
	t->enabled = 1;
	return t;
}

void mm2o_trace_destroy(mm2o_trace_t *t)
{
	if (t == 0) return;
	free(t->mini.a); free(t->anchors.a); free(t->f); free(t->p); free(t->v); free(t->u); free(t->chained.a);
	free(t->dp); free(t->seqpool); free(t->cigpool);
	free(t);
}

/* ---------------- batch driver (pthreads; one read per work item, like mappy called from a ThreadPool) ---------------- */
typedef struct {
	const mm2o_idx_t *mi; const mm2o_opt_t *opt;
	int n; const char *cat; const int64_t *off;
	mm2o_result_t **results;
	volatile int next;
	pthread_mutex_t mu;
} batch_t;

static void *batch_worker(void *arg)
{
	batch_t *b = (batch_t*)arg;
	for (;;) {
		int i;
		pthread_mutex_lock(&b->mu);
		i = b->next; b->next += 16;
		pthread_mutex_unlock(&b->mu);
		if (i >= b->n) break;
		for (int j = i; j < i + 16 && j < b->n; ++j)
			b->results[j] = mm2o_map(b->mi, b->opt, b->cat + b->off[j], (int)(b->off[j+1] - b->off[j]), 0);
	}
	return 0;
}

void mm2o_map_batch(const mm2o_idx_t *mi, const mm2o_opt_t *opt, int n, const char *cat, const int64_t *off, int n_threads, mm2o_result_t **results)
{
	batch_t b;
	pthread_t *th;
	int i;
	b.mi = mi, b.opt = opt, b.n = n, b.cat = cat, b.off = off, b.results = results, b.next = 0;
	pthread_mutex_init(&b.mu, 0);
	if (n_threads < 1) n_threads = 1;
	th = (pthread_t*)malloc(n_threads * sizeof(pthread_t));
	for (i = 0; i < n_threads; ++i) pthread_create(&th[i], 0, batch_worker, &b);
	for (i = 0; i < n_threads; ++i) pthread_join(th[i], 0);
	free(th);
	pthread_mutex_destroy(&b.mu);
}

/* ---- struct-of-arrays export of a mapped batch (checker side of the full-batch parity runs in bench.py / tests) ----
 * Pass 1 (fields == NULL): returns the number of hits and stores the number of CIGAR words in *n_cigar_words.
 * Pass 2: fields[n_hits * 22] in the column order of mm2o_hit_t (cigar_off excluded), hit_off[n + 1], cigar[n_cigar_words]. */
int64_t mm2o_batch_export(mm2o_result_t **results, int n, int32_t *fields, int64_t *hit_off, uint32_t *cigar, int64_t *n_cigar_words)
{
	int64_t nh = 0, nc = 0;
	int i, j;
	for (i = 0; i < n; ++i) {
		const mm2o_result_t *r = results[i];
		if (hit_off) hit_off[i] = nh;
		for (j = 0; j < r->n_hits; ++j) {
			const mm2o_hit_t *h = &r->hits[j];
			if (fields) {
				memcpy(fields + (nh + j) * 22, h, 22 * sizeof(int32_t));
				memcpy(cigar + nc, r->cigar_pool + h->cigar_off, (size_t)h->n_cigar * 4);
			}
			nc += h->n_cigar;
		}
		nh += r->n_hits;
	}
	if (hit_off) hit_off[n] = nh;
	if (n_cigar_words) *n_cigar_words = nc;
	return nh;
}

/* libm's logf over a run of float bit patterns: what mm_set_mapq's `logf(...)` (hit.c, minimap2-2.17) evaluates to on this
 * host.  Used to sweep the device's logf against it (tests/test_gpu_parity.py). */
void mm2o_logf_range(uint32_t first_bits, int64_t n, float *out)
{
	int64_t i;
	for (i = 0; i < n; ++i) {
		uint32_t b = first_bits + (uint32_t)i;
		float x;
		memcpy(&x, &b, 4);
		out[i] = logf(x);
	}
}
