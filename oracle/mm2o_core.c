/*
 * mm2o_core.c -- CPU ORACLE (test infrastructure): sketch, sorts, index, seeds, chaining.
 * Restates minimap2 v2.17 sketch.c / ksort.h / index.c / map.c / chain.c (upstream, absent
 * from /root/reference; reached via mappy at /root/reference/monica/genomes/aligner.py:45,59,193,215).
 * PARITY UNPINNED -- see mm2o.h.
 */
#include <stdlib.h>
#include <pthread.h>
#include <string.h>
#include <assert.h>
#include <stdio.h>
#include "mm2o.h"

/* ---------------- options: upstream options.c mm_mapopt_init ---------------- */
void mm2o_opt_init(mm2o_opt_t *opt)
{
	memset(opt, 0, sizeof(*opt));
	opt->seed = 11;
	opt->mid_occ_frac = 2e-4f;
	opt->min_cnt = 3;
	opt->min_chain_score = 40;
	opt->bw = 500;
	opt->max_gap = 5000;
	opt->max_gap_ref = -1;
	opt->max_chain_skip = 25;
	opt->max_chain_iter = 5000;
	opt->mask_level = 0.5f;
	opt->pri_ratio = 0.8f;
	opt->best_n = 5;
	opt->max_join_long = 20000;
	opt->max_join_short = 2000;
	opt->min_join_flank_sc = 1000;
	opt->min_join_flank_ratio = 0.5f;
	opt->a = 2, opt->b = 4, opt->q = 4, opt->e = 2, opt->q2 = 24, opt->e2 = 1;
	opt->sc_ambi = 1;
	opt->zdrop = 400, opt->zdrop_inv = 200;
	opt->end_bonus = -1;
	opt->min_dp_max = opt->min_chain_score * opt->a;
	opt->min_ksw_len = 200;
	opt->max_clip_ratio = 1.0f;
	opt->max_sw_mat = 100000000;
	opt->mid_occ = 0;
}

/* ---------------- sketch.c ---------------- */
static const unsigned char seq_nt4_table[256] = {
	0, 1, 2, 3,  4, 4, 4, 4,  4, 4, 4, 4,  4, 4, 4, 4,
	4, 4, 4, 4,  4, 4, 4, 4,  4, 4, 4, 4,  4, 4, 4, 4,
	4, 4, 4, 4,  4, 4, 4, 4,  4, 4, 4, 4,  4, 4, 4, 4,
	4, 4, 4, 4,  4, 4, 4, 4,  4, 4, 4, 4,  4, 4, 4, 4,
	4, 0, 4, 1,  4, 4, 4, 2,  4, 4, 4, 4,  4, 4, 4, 4,
	4, 4, 4, 4,  3, 3, 4, 4,  4, 4, 4, 4,  4, 4, 4, 4,
	4, 0, 4, 1,  4, 4, 4, 2,  4, 4, 4, 4,  4, 4, 4, 4,
	4, 4, 4, 4,  3, 3, 4, 4,  4, 4, 4, 4,  4, 4, 4, 4,
	4, 4, 4, 4,  4, 4, 4, 4,  4, 4, 4, 4,  4, 4, 4, 4,
	4, 4, 4, 4,  4, 4, 4, 4,  4, 4, 4, 4,  4, 4, 4, 4,
	4, 4, 4, 4,  4, 4, 4, 4,  4, 4, 4, 4,  4, 4, 4, 4,
	4, 4, 4, 4,  4, 4, 4, 4,  4, 4, 4, 4,  4, 4, 4, 4,
	4, 4, 4, 4,  4, 4, 4, 4,  4, 4, 4, 4,  4, 4, 4, 4,
	4, 4, 4, 4,  4, 4, 4, 4,  4, 4, 4, 4,  4, 4, 4, 4,
	4, 4, 4, 4,  4, 4, 4, 4,  4, 4, 4, 4,  4, 4, 4, 4,
	4, 4, 4, 4,  4, 4, 4, 4,  4, 4, 4, 4,  4, 4, 4, 4
};

const unsigned char *mm2o_nt4_table(void) { return seq_nt4_table; }

/* sketch.c hash64(): Thomas Wang's invertible integer hash, masked */
uint64_t mm2o_hash64(uint64_t key, uint64_t mask)
{
	key = (~key + (key << 21)) & mask;
	key = key ^ key >> 24;
	key = ((key + (key << 3)) + (key << 8)) & mask;
	key = key ^ key >> 14;
	key = ((key + (key << 2)) + (key << 4)) & mask;
	key = key ^ key >> 28;
	key = (key + (key << 31)) & mask;
	return key;
}

static inline void v_push(mm128_v *p, mm128_t v)
{
	if (p->n == p->m) {
		p->m = p->m ? p->m << 1 : 256;
		p->a = (mm128_t*)realloc(p->a, p->m * sizeof(mm128_t));
	}
	p->a[p->n++] = v;
}

/* sketch.c mm_sketch(), non-HPC branch */
void mm2o_sketch(const char *str, int len, int w, int k, uint32_t rid, mm128_v *p)
{
	uint64_t shift1 = 2 * (k - 1), mask = (1ULL<<2*k) - 1, kmer[2] = {0,0};
	int i, j, l, buf_pos, min_pos, kmer_span = 0;
	mm128_t buf[256], min = { UINT64_MAX, UINT64_MAX };

	assert(len > 0 && (w > 0 && w < 256) && (k > 0 && k <= 28));
	memset(buf, 0xff, w * 16);

	for (i = l = buf_pos = min_pos = 0; i < len; ++i) {
		int c = seq_nt4_table[(uint8_t)str[i]];
		mm128_t info = { UINT64_MAX, UINT64_MAX };
		if (c < 4) {
			int z;
			kmer_span = l + 1 < k? l + 1 : k;
			kmer[0] = (kmer[0] << 2 | c) & mask;
			kmer[1] = (kmer[1] >> 2) | (3ULL^c) << shift1;
			if (kmer[0] == kmer[1]) continue; /* symmetric k-mer: strand unknown */
			z = kmer[0] < kmer[1]? 0 : 1;
			++l;
			if (l >= k && kmer_span < 256) {
				info.x = mm2o_hash64(kmer[z], mask) << 8 | kmer_span;
				info.y = (uint64_t)rid<<32 | (uint32_t)i<<1 | z;
			}
		} else l = 0, kmer_span = 0;
		buf[buf_pos] = info;
		if (l == w + k - 1 && min.x != UINT64_MAX) { /* first window: identical k-mers not stored yet */
			for (j = buf_pos + 1; j < w; ++j)
				if (min.x == buf[j].x && buf[j].y != min.y) v_push(p, buf[j]);
			for (j = 0; j < buf_pos; ++j)
				if (min.x == buf[j].x && buf[j].y != min.y) v_push(p, buf[j]);
		}
		if (info.x <= min.x) { /* a new minimum; write the old min */
			if (l >= w + k && min.x != UINT64_MAX) v_push(p, min);
			min = info, min_pos = buf_pos;
		} else if (buf_pos == min_pos) { /* old min has moved outside the window */
			if (l >= w + k - 1 && min.x != UINT64_MAX) v_push(p, min);
			for (j = buf_pos + 1, min.x = UINT64_MAX; j < w; ++j)
				if (min.x >= buf[j].x) min = buf[j], min_pos = j;
			for (j = 0; j <= buf_pos; ++j)
				if (min.x >= buf[j].x) min = buf[j], min_pos = j;
			if (l >= w + k - 1 && min.x != UINT64_MAX) { /* write identical k-mers */
				for (j = buf_pos + 1; j < w; ++j)
					if (min.x == buf[j].x && min.y != buf[j].y) v_push(p, buf[j]);
				for (j = 0; j <= buf_pos; ++j)
					if (min.x == buf[j].x && min.y != buf[j].y) v_push(p, buf[j]);
			}
		}
		if (++buf_pos == w) buf_pos = 0;
	}
	if (min.x != UINT64_MAX)
		v_push(p, min);
}

int64_t mm2o_sketch_buf(const char *str, int len, int w, int k, uint32_t rid, uint64_t *out_xy, int64_t cap)
{
	mm128_v v = {0,0,0};
	int64_t i, n;
	if (len > 0) mm2o_sketch(str, len, w, k, rid, &v);
	n = (int64_t)v.n;
	for (i = 0; i < n && i < cap; ++i) out_xy[2*i] = v.a[i].x, out_xy[2*i+1] = v.a[i].y;
	free(v.a);
	return n;
}

/* ---------------- ksort.h KRADIX_SORT_INIT (RS_MIN_SIZE 64, RS_MAX_BITS 8) ---------------- */
#define RS_MIN_SIZE 64
#define RS_MAX_BITS 8

#define RADIX_IMPL(name, rstype_t, rskey, sizeof_key) \
	typedef struct { rstype_t *b, *e; } rsbucket_##name##_t; \
	static void rs_insertsort_##name(rstype_t *beg, rstype_t *end) \
	{ \
		rstype_t *i; \
		for (i = beg + 1; i < end; ++i) \
			if (rskey(*i) < rskey(*(i - 1))) { \
				rstype_t *j, tmp = *i; \
				for (j = i; j > beg && rskey(tmp) < rskey(*(j-1)); --j) \
					*j = *(j - 1); \
				*j = tmp; \
			} \
	} \
	static void rs_sort_##name(rstype_t *beg, rstype_t *end, int n_bits, int s) \
	{ \
		rstype_t *i; \
		int size = 1<<n_bits, m = size - 1; \
		rsbucket_##name##_t *k, b[1<<RS_MAX_BITS], *be = b + size; \
		for (k = b; k != be; ++k) k->b = k->e = beg; \
		for (i = beg; i != end; ++i) ++b[rskey(*i)>>s&m].e; \
		for (k = b + 1; k != be; ++k) \
			k->e += (k-1)->e - beg, k->b = (k-1)->e; \
		for (k = b; k != be;) { \
			if (k->b != k->e) { \
				rsbucket_##name##_t *l; \
				if ((l = b + (rskey(*k->b)>>s&m)) != k) { \
					rstype_t tmp = *k->b, swap; \
					do { \
						swap = tmp; tmp = *l->b; *l->b++ = swap; \
						l = b + (rskey(tmp)>>s&m); \
					} while (l != k); \
					*k->b++ = tmp; \
				} else ++k->b; \
			} else ++k; \
		} \
		for (b->b = beg, k = b + 1; k != be; ++k) k->b = (k-1)->e; \
		if (s) { \
			s = s > n_bits? s - n_bits : 0; \
			for (k = b; k != be; ++k) \
				if (k->e - k->b > RS_MIN_SIZE) rs_sort_##name(k->b, k->e, n_bits, s); \
				else if (k->e - k->b > 1) rs_insertsort_##name(k->b, k->e); \
		} \
	} \
	void mm2o_radix_sort_##name(rstype_t *beg, rstype_t *end) \
	{ \
		if (end - beg <= RS_MIN_SIZE) rs_insertsort_##name(beg, end); \
		else rs_sort_##name(beg, end, RS_MAX_BITS, (sizeof_key - 1) * RS_MAX_BITS); \
	}

#define sort_key_128x(a) ((a).x)
RADIX_IMPL(128x, mm128_t, sort_key_128x, 8)
#define sort_key_64(x) (x)
RADIX_IMPL(64, uint64_t, sort_key_64, 8)

/* ---------------- index.c ---------------- */
static int cmp_hy(const void *a_, const void *b_)
{
	const mm128_t *a = (const mm128_t*)a_, *b = (const mm128_t*)b_;
	uint64_t ha = a->x >> 8, hb = b->x >> 8;
	if (ha != hb) return ha < hb ? -1 : 1;
	if (a->y != b->y) return a->y < b->y ? -1 : 1;
	return 0;
}

/* index.c mm_idx_gen()/worker_post(): one record per distinct hash, positions sorted by y.
 * The bucket/khash layout of upstream is an implementation detail; lookups return the same
 * (n, sorted position list) and mm_idx_cal_max_occ sees the same multiset of counts. */
/* sketch the contigs on several threads (contig i -> its own vector), then concatenate in contig order: the same array a
 * sequential pass produces */
typedef struct { int n_seq; const char **seqs; const int64_t *lens; int w, k; mm128_v *per; volatile int next; pthread_mutex_t mu; } sk_job_t;
static void *sk_worker(void *arg)
{
	sk_job_t *j = (sk_job_t*)arg;
	for (;;) {
		int s;
		pthread_mutex_lock(&j->mu); s = j->next++; pthread_mutex_unlock(&j->mu);
		if (s >= j->n_seq) break;
		if (j->lens[s] > 0) mm2o_sketch(j->seqs[s], (int)j->lens[s], j->w, j->k, (uint32_t)s, &j->per[s]);
	}
	return 0;
}

/* stable LSD radix sort of the minimizers on the hash (x >> 8): equal hashes keep their (contig, position) order, which is
 * the order by y that upstream's worker_post establishes with its own sort */
static void sort_by_hash(mm128_t *a, size_t n, int key_bits)
{
	mm128_t *tmp, *src = a, *dst;
	int shift;
	if (n < 2) return;
	tmp = (mm128_t*)malloc(n * sizeof(mm128_t));
	dst = tmp;
	for (shift = 0; shift < key_bits; shift += 11) {
		size_t cnt[2049], i;
		memset(cnt, 0, sizeof(cnt));
		for (i = 0; i < n; ++i) ++cnt[((src[i].x >> 8) >> shift & 2047) + 1];
		for (i = 1; i <= 2048; ++i) cnt[i] += cnt[i - 1];
		for (i = 0; i < n; ++i) dst[cnt[(src[i].x >> 8) >> shift & 2047]++] = src[i];
		{ mm128_t *t = src; src = dst; dst = t; }
	}
	if (src != a) memcpy(a, src, n * sizeof(mm128_t));
	free(tmp);
}

mm2o_idx_t *mm2o_idx_build_mt(int n_seq, const char **names, const char **seqs, const int64_t *lens, int w, int k, int n_threads)
{
	mm2o_idx_t *mi = (mm2o_idx_t*)calloc(1, sizeof(mm2o_idx_t));
	mm128_v a = {0,0,0};
	uint64_t sum_len = 0, i, j, cap;
	int s;
	mi->k = k, mi->w = w, mi->b = 14, mi->flag = 0;
	mi->n_seq = n_seq;
	mi->seq = (mm2o_seq_t*)calloc(n_seq, sizeof(mm2o_seq_t));
	for (s = 0; s < n_seq; ++s) {
		mi->seq[s].name = strdup(names[s]);
		mi->seq[s].len = (uint32_t)lens[s];
		mi->seq[s].offset = sum_len;
		sum_len += lens[s];
	}
	mi->S = (uint32_t*)calloc((sum_len + 7) / 8 + 1, 4);
	for (s = 0; s < n_seq; ++s) {
		uint64_t o = mi->seq[s].offset;
		for (j = 0; j < (uint64_t)lens[s]; ++j) {
			uint32_t c = seq_nt4_table[(uint8_t)seqs[s][j]];
			uint64_t pos = o + j;
			mi->S[pos>>3] |= c << ((pos & 7) << 2); /* mm_seq4_set */
		}
	}
	{
		sk_job_t job; pthread_t th[256]; int t; size_t tot = 0, at = 0;
		if (n_threads < 1) n_threads = 1;
		if (n_threads > 256) n_threads = 256;
		if (n_threads > n_seq) n_threads = n_seq > 0 ? n_seq : 1;
		job.n_seq = n_seq, job.seqs = seqs, job.lens = lens, job.w = w, job.k = k, job.next = 0;
		job.per = (mm128_v*)calloc(n_seq > 0 ? n_seq : 1, sizeof(mm128_v));
		pthread_mutex_init(&job.mu, 0);
		for (t = 0; t < n_threads; ++t) pthread_create(&th[t], 0, sk_worker, &job);
		for (t = 0; t < n_threads; ++t) pthread_join(th[t], 0);
		pthread_mutex_destroy(&job.mu);
		for (s = 0; s < n_seq; ++s) tot += job.per[s].n;
		a.n = a.m = tot;
		a.a = (mm128_t*)malloc((tot + 1) * sizeof(mm128_t));
		for (s = 0; s < n_seq; ++s) {
			if (job.per[s].n) memcpy(a.a + at, job.per[s].a, job.per[s].n * sizeof(mm128_t));
			at += job.per[s].n;
			free(job.per[s].a);
		}
		free(job.per);
	}
	sort_by_hash(a.a, a.n, 2 * k);
	mi->n_pos = a.n;
	for (i = 0, mi->n_keys = 0; i < a.n; ++i)
		if (i == 0 || (a.a[i].x>>8) != (a.a[i-1].x>>8)) ++mi->n_keys;
	mi->keys = (uint64_t*)malloc((mi->n_keys + 1) * 8);
	mi->key_off = (uint64_t*)malloc((mi->n_keys + 1) * 8);
	mi->pos = (uint64_t*)malloc((a.n + 1) * 8);
	for (i = 0, j = 0; i < a.n; ++i) {
		if (i == 0 || (a.a[i].x>>8) != (a.a[i-1].x>>8)) {
			mi->keys[j] = a.a[i].x >> 8;
			mi->key_off[j++] = i;
		}
		mi->pos[i] = a.a[i].y;
	}
	mi->key_off[mi->n_keys] = a.n;
	free(a.a);
	for (cap = 16; cap < mi->n_keys * 2 + 2; cap <<= 1) {}
	mi->tab_mask = cap - 1;
	mi->tab = (uint32_t*)calloc(cap, 4);
	for (i = 0; i < mi->n_keys; ++i) {
		uint64_t h = (mi->keys[i] * 0x9E3779B97F4A7C15ULL) >> 20 & mi->tab_mask;
		while (mi->tab[h]) h = (h + 1) & mi->tab_mask;
		mi->tab[h] = (uint32_t)(i + 1);
	}
	return mi;
}

mm2o_idx_t *mm2o_idx_build(int n_seq, const char **names, const char **seqs, const int64_t *lens, int w, int k)
{
	return mm2o_idx_build_mt(n_seq, names, seqs, lens, w, k, 1);
}

void mm2o_idx_destroy(mm2o_idx_t *mi)
{
	uint32_t i;
	if (mi == 0) return;
	for (i = 0; i < mi->n_seq; ++i) free(mi->seq[i].name);
	free(mi->seq); free(mi->S); free(mi->keys); free(mi->key_off); free(mi->pos); free(mi->tab);
	free(mi);
}

/* index.c mm_idx_get() */
const uint64_t *mm2o_idx_get(const mm2o_idx_t *mi, uint64_t minier, int *n)
{
	uint64_t h = (minier * 0x9E3779B97F4A7C15ULL) >> 20 & mi->tab_mask;
	*n = 0;
	while (mi->tab[h]) {
		uint64_t i = mi->tab[h] - 1;
		if (mi->keys[i] == minier) {
			*n = (int)(mi->key_off[i+1] - mi->key_off[i]);
			return &mi->pos[mi->key_off[i]];
		}
		h = (h + 1) & mi->tab_mask;
	}
	return 0;
}

static int cmp_u32(const void *a, const void *b)
{
	uint32_t x = *(const uint32_t*)a, y = *(const uint32_t*)b;
	return x < y ? -1 : x > y;
}

/* index.c mm_idx_cal_max_occ(): ks_ksmall(counts, (1-f)*n) + 1 */
int32_t mm2o_idx_cal_max_occ(const mm2o_idx_t *mi, float f)
{
	uint64_t i, n = mi->n_keys;
	uint32_t *a, thres;
	if (f <= 0.) return INT32_MAX;
	if (n == 0) return 1; /* upstream would read an empty array; never reached with a real index */
	a = (uint32_t*)malloc(n * 4);
	for (i = 0; i < n; ++i) a[i] = (uint32_t)(mi->key_off[i+1] - mi->key_off[i]);
	qsort(a, n, 4, cmp_u32);
	thres = a[(uint32_t)((1. - f) * n)] + 1;
	free(a);
	return (int32_t)thres;
}

/* options.c mm_mapopt_update() */
void mm2o_mapopt_update(mm2o_opt_t *opt, const mm2o_idx_t *mi)
{
	if (opt->mid_occ <= 0)
		opt->mid_occ = mm2o_idx_cal_max_occ(mi, opt->mid_occ_frac);
}

/* index.c mm_idx_getseq() */
int mm2o_idx_getseq(const mm2o_idx_t *mi, uint32_t rid, uint32_t st, uint32_t en, uint8_t *seq)
{
	uint64_t i, st1, en1;
	if (rid >= mi->n_seq || st >= mi->seq[rid].len) return -1;
	if (en > mi->seq[rid].len) en = mi->seq[rid].len;
	st1 = mi->seq[rid].offset + st;
	en1 = mi->seq[rid].offset + en;
	for (i = st1; i < en1; ++i)
		seq[i - st1] = mi->S[i>>3] >> ((i&7)<<2) & 0xf;
	return (int)(en - st);
}

/* ---------------- map.c: collect_matches() + collect_seed_hits() ---------------- */
typedef struct {
	uint32_t n;
	uint32_t q_pos, q_span;
	uint32_t seg_id:31, is_tandem:1;
	const uint64_t *cr;
} mm_match_t;

mm128_t *mm2o_collect_seed_hits(const mm2o_opt_t *opt, int max_occ, const mm2o_idx_t *mi, const mm128_v *mv, int qlen, int64_t *n_a, int *rep_len)
{
	int rep_st = 0, rep_en = 0, n_m, i;
	size_t ii;
	mm_match_t *m;
	mm128_t *a;
	(void)opt;
	m = (mm_match_t*)malloc((mv->n + 1) * sizeof(mm_match_t));
	for (ii = 0, n_m = 0, *rep_len = 0, *n_a = 0; ii < mv->n; ++ii) {
		const uint64_t *cr;
		mm128_t *p = &mv->a[ii];
		uint32_t q_pos = (uint32_t)p->y, q_span = p->x & 0xff;
		int t;
		cr = mm2o_idx_get(mi, p->x>>8, &t);
		if (t >= max_occ) {
			int en = (q_pos >> 1) + 1, st = en - q_span;
			if (st > rep_en) {
				*rep_len += rep_en - rep_st;
				rep_st = st, rep_en = en;
			} else rep_en = en;
		} else {
			mm_match_t *q = &m[n_m++];
			q->q_pos = q_pos, q->q_span = q_span, q->cr = cr, q->n = t, q->seg_id = p->y >> 32;
			q->is_tandem = 0;
			if (ii > 0 && p->x>>8 == mv->a[ii - 1].x>>8) q->is_tandem = 1;
			if (ii < mv->n - 1 && p->x>>8 == mv->a[ii + 1].x>>8) q->is_tandem = 1;
			*n_a += q->n;
		}
	}
	*rep_len += rep_en - rep_st;
	a = (mm128_t*)malloc((*n_a + 1) * sizeof(mm128_t));
	for (i = 0, *n_a = 0; i < n_m; ++i) {
		mm_match_t *q = &m[i];
		const uint64_t *r = q->cr;
		uint32_t k;
		for (k = 0; k < q->n; ++k) {
			int32_t rpos = (uint32_t)r[k] >> 1;
			mm128_t *p = &a[(*n_a)++];
			if ((r[k]&1) == (q->q_pos&1)) { /* forward strand */
				p->x = (r[k]&0xffffffff00000000ULL) | rpos;
				p->y = (uint64_t)q->q_span << 32 | q->q_pos >> 1;
			} else { /* reverse strand */
				p->x = 1ULL<<63 | (r[k]&0xffffffff00000000ULL) | rpos;
				p->y = (uint64_t)q->q_span << 32 | (qlen - ((q->q_pos>>1) + 1 - q->q_span) - 1);
			}
			p->y |= (uint64_t)q->seg_id << MM_SEED_SEG_SHIFT;
			if (q->is_tandem) p->y |= MM_SEED_TANDEM;
		}
	}
	free(m);
	mm2o_radix_sort_128x(a, a + (*n_a));
	return a;
}

/* ---------------- chain.c mm_chain_dp() (v2.17 form) ---------------- */
static inline int ilog2_32(uint32_t v)
{
	int r = 0;
	while (v >>= 1) ++r; /* == upstream LogTable256 lookup for v > 0 */
	return r;
}

mm128_t *mm2o_chain_dp(int max_dist_x, int max_dist_y, int bw, int max_skip, int max_iter, int min_cnt, int min_sc,
					   int64_t n, mm128_t *a, int *n_u_, uint64_t **_u, mm2o_trace_t *trace, int64_t *cells)
{
	int32_t k, *f, *p, *t, *v, n_u, n_v;
	int64_t i, j, st = 0, n_cells = 0;
	uint64_t *u, *u2, sum_qspan = 0;
	float avg_qspan;
	mm128_t *b, *w;

	if (_u) *_u = 0, *n_u_ = 0;
	if (n == 0 || a == 0) { free(a); return 0; }
	f = (int32_t*)malloc(n * 4);
	p = (int32_t*)malloc(n * 4);
	t = (int32_t*)malloc(n * 4);
	v = (int32_t*)malloc(n * 4);
	memset(t, 0, n * 4);

	for (i = 0; i < n; ++i) sum_qspan += a[i].y>>32&0xff;
	avg_qspan = (float)sum_qspan / n;
	for (i = 0; i < n; ++i) {
		uint64_t ri = a[i].x;
		int64_t max_j = -1;
		int32_t qi = (int32_t)a[i].y, q_span = a[i].y>>32&0xff;
		int32_t max_f = q_span, n_skip = 0, min_d;
		while (st < i && ri > a[st].x + max_dist_x) ++st;
		if (i - st > max_iter) st = i - max_iter;
		for (j = i - 1; j >= st; --j) {
			int64_t dr = ri - a[j].x;
			int32_t dq = qi - (int32_t)a[j].y, dd, sc, log_dd;
			++n_cells;
			if (dr == 0 || dq <= 0) continue;
			if (dq > max_dist_y || dq > max_dist_x) continue;
			dd = dr > dq? dr - dq : dq - dr;
			if (dd > bw) continue;
			min_d = dq < dr? dq : dr;
			sc = min_d > q_span? q_span : dq < dr? dq : dr;
			log_dd = dd? ilog2_32(dd) : 0;
			sc -= (int)(dd * .01 * avg_qspan) + (log_dd>>1);
			sc += f[j];
			if (sc > max_f) {
				max_f = sc, max_j = j;
				if (n_skip > 0) --n_skip;
			} else if (t[j] == i) {
				if (++n_skip > max_skip)
					break;
			}
			if (p[j] >= 0) t[p[j]] = i;
		}
		f[i] = max_f, p[i] = max_j;
		v[i] = max_j >= 0 && v[max_j] > max_f? v[max_j] : max_f;
	}
	if (cells) *cells += n_cells;
	if (trace && trace->enabled) {
		trace->n_chain_arr = n;
		trace->f = (int32_t*)malloc(n * 4); memcpy(trace->f, f, n * 4);
		trace->p = (int32_t*)malloc(n * 4); memcpy(trace->p, p, n * 4);
		trace->v = (int32_t*)malloc(n * 4); memcpy(trace->v, v, n * 4);
	}

	/* find the ending positions of chains */
	memset(t, 0, n * 4);
	for (i = 0; i < n; ++i)
		if (p[i] >= 0) t[p[i]] = 1;
	for (i = n_u = 0; i < n; ++i)
		if (t[i] == 0 && v[i] >= min_sc)
			++n_u;
	if (n_u == 0) {
		free(a); free(f); free(p); free(t); free(v);
		return 0;
	}
	u = (uint64_t*)malloc(n_u * 8);
	for (i = n_u = 0; i < n; ++i) {
		if (t[i] == 0 && v[i] >= min_sc) {
			j = i;
			while (j >= 0 && f[j] < v[j]) j = p[j]; /* find the peak that maximizes f[] */
			if (j < 0) j = i;
			u[n_u++] = (uint64_t)f[j] << 32 | j;
		}
	}
	mm2o_radix_sort_64(u, u + n_u);
	for (i = 0; i < n_u>>1; ++i) { /* reverse: highest score first */
		uint64_t tt = u[i];
		u[i] = u[n_u - i - 1], u[n_u - i - 1] = tt;
	}

	/* backtrack */
	memset(t, 0, n * 4);
	for (i = n_v = k = 0; i < n_u; ++i) {
		int32_t n_v0 = n_v, k0 = k;
		j = (int32_t)u[i];
		do {
			v[n_v++] = j;
			t[j] = 1;
			j = p[j];
		} while (j >= 0 && t[j] == 0);
		if (j < 0) {
			if (n_v - n_v0 >= min_cnt) u[k++] = u[i]>>32<<32 | (n_v - n_v0);
		} else if ((int32_t)(u[i]>>32) - f[j] >= min_sc) {
			if (n_v - n_v0 >= min_cnt) u[k++] = ((u[i]>>32) - f[j]) << 32 | (n_v - n_v0);
		}
		if (k0 == k) n_v = n_v0; /* no new chain added, reset */
	}
	*n_u_ = n_u = k, *_u = u;
	free(f); free(p); free(t);

	/* write the result to b[] */
	b = (mm128_t*)malloc((n_v + 1) * sizeof(mm128_t));
	for (i = 0, k = 0; i < n_u; ++i) {
		int32_t k0 = k, ni = (int32_t)u[i];
		for (j = 0; j < ni; ++j)
			b[k] = a[v[k0 + (ni - j - 1)]], ++k;
	}
	free(v);

	/* sort u[] and a[] by a[].x, such that adjacent chains may be joined */
	w = (mm128_t*)malloc((n_u + 1) * sizeof(mm128_t));
	for (i = k = 0; i < n_u; ++i) {
		w[i].x = b[k].x, w[i].y = (uint64_t)k<<32|i;
		k += (int32_t)u[i];
	}
	mm2o_radix_sort_128x(w, w + n_u);
	u2 = (uint64_t*)malloc((n_u + 1) * 8);
	for (i = k = 0; i < n_u; ++i) {
		int32_t jj = (int32_t)w[i].y, nn = (int32_t)u[jj];
		u2[i] = u[jj];
		memcpy(&a[k], &b[w[i].y>>32], nn * sizeof(mm128_t));
		k += nn;
	}
	if (n_u) memcpy(u, u2, n_u * 8);
	memcpy(b, a, k * sizeof(mm128_t));
	free(a); free(w); free(u2);
	return b;
}
