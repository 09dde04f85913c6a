/*
 * mm2o.h -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the minimap2 v2.17 map-ont pipeline that monica reaches
 * through mappy==2.17 (/root/reference/requirements.txt:3; call sites
 * /root/reference/monica/genomes/aligner.py:45-46,59,193,215).
 *
 * PARITY UNPINNED: mappy / minimap2 sources are absent from /root/reference and from the
 * build container, and the reference's own tests hold no golden vectors for this path
 * (SURVEY.md section 4, 8c).  Every function below restates the published minimap2-2.17
 * algorithm from the upstream file named in its comment; none of it could be checked
 * against a real mappy run.  Parity claims in this repo therefore read
 * "bit-exact vs this restatement", never "vs mappy".
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product path (monica_b200/) never links or calls it.
 */
#ifndef MM2O_H
#define MM2O_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { uint64_t x, y; } mm128_t;
typedef struct { size_t n, m; mm128_t *a; } mm128_v;

/* anchor flag bits (upstream mmpriv.h) */
#define MM_SEED_LONG_JOIN (1ULL<<40)
#define MM_SEED_IGNORE    (1ULL<<41)
#define MM_SEED_TANDEM    (1ULL<<42)
#define MM_SEED_SELF      (1ULL<<43)
#define MM_SEED_SEG_SHIFT 48
#define MM_SEED_SEG_MASK  (0xffULL<<(MM_SEED_SEG_SHIFT))

#define MM_PARENT_UNSET   (-1)
#define MM_PARENT_TMP_PRI (-2)

/* ksw2 flags (upstream ksw2.h) */
#define KSW_NEG_INF       (-0x40000000)
#define KSW_EZ_SCORE_ONLY  0x01
#define KSW_EZ_RIGHT       0x02
#define KSW_EZ_GENERIC_SC  0x04
#define KSW_EZ_APPROX_MAX  0x08
#define KSW_EZ_APPROX_DROP 0x10
#define KSW_EZ_EXTZ_ONLY   0x40
#define KSW_EZ_REV_CIGAR   0x80

typedef struct {
	uint32_t max:31, zdropped:1;
	int max_q, max_t;
	int mqe, mqe_t;
	int mte, mte_q;
	int score;
	int m_cigar, n_cigar;
	int reach_end;
	uint32_t *cigar;
} ksw_extz_t;

/* mapping options: upstream mm_mapopt_init() defaults (options.c) == map-ont apart from k */
typedef struct {
	int seed;
	float mid_occ_frac;
	int min_cnt, min_chain_score, bw, max_gap, max_gap_ref, max_chain_skip, max_chain_iter;
	float mask_level, pri_ratio;
	int best_n;
	int max_join_long, max_join_short, min_join_flank_sc;
	float min_join_flank_ratio;
	int a, b, q, e, q2, e2, sc_ambi, zdrop, zdrop_inv, end_bonus, min_dp_max, min_ksw_len;
	float max_clip_ratio;
	int64_t max_sw_mat;
	int mid_occ;
} mm2o_opt_t;

typedef struct {
	char *name;
	uint64_t offset;
	uint32_t len;
} mm2o_seq_t;

typedef struct {
	int32_t k, w, b, flag;
	uint32_t n_seq;
	mm2o_seq_t *seq;
	uint32_t *S;            /* 4-bit packed reference, 8 bases per word (upstream mm_idx_t::S) */
	/* lookup structure: distinct minimizer hashes with their sorted position lists */
	uint64_t n_keys;        /* number of distinct hashes */
	uint64_t n_pos;         /* total number of (hash,pos) records */
	uint64_t *keys;         /* [n_keys] hash values, ascending */
	uint64_t *key_off;      /* [n_keys+1] offsets into pos */
	uint64_t *pos;          /* [n_pos] y values (rid<<32|pos<<1|strand), ascending within a key */
	uint64_t tab_mask;      /* open-addressing table capacity-1 */
	uint32_t *tab;          /* [cap] index into keys + 1, 0 = empty */
} mm2o_idx_t;

/* one alignment record: the fields mappy exposes (python/cmappy.h mm_reg2hitpy) + dp_max */
typedef struct {
	int32_t rid, rev;
	int32_t qs, qe, rs, re;
	int32_t mapq, mlen, blen, nm;
	int32_t dp_max, dp_max2, score, score0, cnt, subsc, n_sub;
	int32_t id, parent, is_primary, sam_pri;
	int32_t n_cigar;
	int32_t cigar_off;      /* offset into the result's cigar pool */
} mm2o_hit_t;

typedef struct {
	int32_t n_hits;
	mm2o_hit_t *hits;
	int32_t n_cigar_pool;
	uint32_t *cigar_pool;
	int32_t rep_len;
	int64_t n_mini, n_anchor, chain_cells, dp_cells, n_dp_calls;
} mm2o_result_t;

/* optional stage trace (tests compare the CUDA stages against these) */
typedef struct {
	/* one record per ksw call */
	int32_t qlen, tlen, w, zdrop, end_bonus, flag;
	int64_t q_off, t_off;         /* offsets into trace q/t byte pools */
	int32_t score, max, max_q, max_t, mqe, mqe_t, zdropped, reach_end, n_cigar;
	int64_t cigar_off;
} mm2o_dp_rec_t;

typedef struct {
	int enabled;
	mm128_v mini;                 /* query minimizers */
	mm128_v anchors;              /* sorted anchors fed to chaining */
	int32_t *f, *p, *v;           /* chaining arrays, length anchors.n */
	int64_t n_chain_arr;
	int32_t n_u; uint64_t *u;     /* chains after mm_chain_dp: score<<32|cnt */
	mm128_v chained;              /* anchors after mm_chain_dp, chain by chain */
	/* DP calls */
	int64_t n_dp, m_dp; mm2o_dp_rec_t *dp;
	int64_t n_seq, m_seq; uint8_t *seqpool;
	int64_t n_cig, m_cig; uint32_t *cigpool;
} mm2o_trace_t;

/* ---- C-ABI of the oracle (loaded with ctypes by tests / bench baseline only) ---- */
void mm2o_opt_init(mm2o_opt_t *opt);
uint64_t mm2o_hash64(uint64_t key, uint64_t mask);
void mm2o_sketch(const char *str, int len, int w, int k, uint32_t rid, mm128_v *p);
int64_t mm2o_sketch_buf(const char *str, int len, int w, int k, uint32_t rid, uint64_t *out_xy, int64_t cap);

mm2o_idx_t *mm2o_idx_build(int n_seq, const char **names, const char **seqs, const int64_t *lens, int w, int k);
mm2o_idx_t *mm2o_idx_build_mt(int n_seq, const char **names, const char **seqs, const int64_t *lens, int w, int k, int n_threads);
void mm2o_idx_destroy(mm2o_idx_t *mi);
const uint64_t *mm2o_idx_get(const mm2o_idx_t *mi, uint64_t minier, int *n);
int32_t mm2o_idx_cal_max_occ(const mm2o_idx_t *mi, float f);
int mm2o_idx_getseq(const mm2o_idx_t *mi, uint32_t rid, uint32_t st, uint32_t en, uint8_t *seq);
void mm2o_mapopt_update(mm2o_opt_t *opt, const mm2o_idx_t *mi);

mm2o_result_t *mm2o_map(const mm2o_idx_t *mi, const mm2o_opt_t *opt, const char *seq, int qlen, mm2o_trace_t *trace);
void mm2o_result_destroy(mm2o_result_t *r);
mm2o_trace_t *mm2o_trace_new(void);
void mm2o_trace_destroy(mm2o_trace_t *t);

/* batch mapping over n reads with n_threads pthreads; results[i] allocated, free each with mm2o_result_destroy */
void mm2o_map_batch(const mm2o_idx_t *mi, const mm2o_opt_t *opt, int n, const char *cat, const int64_t *off, int n_threads, mm2o_result_t **results);

/* struct-of-arrays export of a mapped batch (two passes: sizes, then data); see mm2o_map.c */
int64_t mm2o_batch_export(mm2o_result_t **results, int n, int32_t *fields, int64_t *hit_off, uint32_t *cigar, int64_t *n_cigar_words);

/* stand-alone DP entry for kernel parity tests */
void mm2o_ksw_set_simd(int on);   /* 1: SSE4.1 16-lane core (default), 0: scalar statement of the same lanes */
void mm2o_ksw_extd2(int qlen, const uint8_t *query, int tlen, const uint8_t *target, int8_t m, const int8_t *mat,
					int8_t q, int8_t e, int8_t q2, int8_t e2, int w, int zdrop, int end_bonus, int flag, ksw_extz_t *ez);
void mm2o_gen_simple_mat(int m, int8_t *mat, int8_t a, int8_t b, int8_t sc_ambi);
/* ksw2_ll_sse.c ksw_ll_qinit + ksw_ll_i16 (striped local alignment: score, query end, target end) */
int mm2o_ksw_ll_i16(int qlen, const uint8_t *query, int tlen, const uint8_t *target, int m, const int8_t *mat, int gapo, int gape, int *qe, int *te);
int64_t mm2o_ksw_cells(int qlen, int tlen, int w);

/* radix sorts with upstream's exact (unstable) permutation (ksort.h KRADIX_SORT_INIT) */
void mm2o_radix_sort_128x(mm128_t *beg, mm128_t *end);
void mm2o_radix_sort_64(uint64_t *beg, uint64_t *end);

/* chaining entry for parity tests: returns chained anchors (malloc'd) and u[] */
mm128_t *mm2o_chain_dp(int max_dist_x, int max_dist_y, int bw, int max_skip, int max_iter, int min_cnt, int min_sc,
					   int64_t n, mm128_t *a, int *n_u_, uint64_t **_u, mm2o_trace_t *trace, int64_t *cells);

/* the host libm's logf -- the function mm_set_mapq (hit.c) calls -- over n consecutive float bit patterns from first_bits */
void mm2o_logf_range(uint32_t first_bits, int64_t n, float *out);

#ifdef __cplusplus
}
#endif
#endif
