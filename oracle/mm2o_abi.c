/*
 * mm2o_abi.c -- the CPU oracle behind the SAME C ABI as the CUDA library (include/monica_b200.h).
 *
 * TEST INFRASTRUCTURE ONLY (PARITY UNPINNED, see mm2o.h).  SURVEY.md 8(b) asks for the boundary to be implemented twice,
 * once by the CUDA library and once by the CPU oracle, so that one harness can drive either: this file wraps the oracle
 * (mm2o_*.c, the restatement of minimap2-2.17's map-ont path) in the entry points of include/monica_b200.h that make up
 * the hot path -- index build, batch mapping, hit arrays, monica's count, and the per-stage sketch / extension / local
 * alignment entries -- with the same names, argument meaning, ownership rules and error codes.  Built into
 * oracle/_build/libmonica_b200_oracle.so; only tests/ load it.  Everything that is device plumbing in the CUDA library
 * (resident reads, packed uploads, NCCL, streams) is not exported here.
 *
 * Reference sites (paths relative to /root/reference):
 *   mb_index_build            mappy.Aligner(fn_idx_in=..., preset='map-ont', best_n=15)   monica/genomes/aligner.py:45-46
 *   mb_map_batch[_ex]         for hit in index.map(str(seq_record.seq))                   monica/genomes/aligner.py:193,215
 *   mb_hits_field             hit.is_primary .mapq .ctg .NM .mlen                         monica/genomes/aligner.py:194-195,216-217
 *   mb_count                  the hit filter, best_hit and the Counter update             monica/genomes/aligner.py:194,216,225-263,328-339
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include "mm2o.h"
#include "../include/monica_b200.h"

#define ABI_NF 23
static const char *ABI_NAMES[ABI_NF] = { "read_idx", "rid", "rev", "qs", "qe", "rs", "re", "mapq", "mlen", "blen", "nm", "dp_max", "dp_max2",
	"score", "score0", "cnt", "subsc", "n_sub", "id", "parent", "is_primary", "sam_pri", "n_cigar" };

struct mb_index { mm2o_idx_t *mi; };
struct mb_hits {
	int64_t n; int32_t n_reads;
	int32_t *fields;      /* [ABI_NF][n] */
	int64_t *cigar_off;   /* [n] */
	uint32_t *cigar; int64_t n_cigar;
	int32_t *rep_len;     /* [n_reads] */
	int64_t *hit_off;     /* [n_reads+1] */
	int64_t *read_off;    /* [n_reads+1] */
};

static __thread char g_err[256];
static int fail(int code, const char *msg) { snprintf(g_err, sizeof(g_err), "%s", msg); return code; }

const char *mb_last_error(void) { return g_err; }
int mb_device_count(void) { return 0; }   /* the oracle needs no device */

int mb_opt_init(mb_opt_t *opt)
{
	if (!opt) return fail(MB_ERR_ARG, "bad arguments");
	/* the two option structs have the same layout (tests/test_boundary_cpu.py::test_opt_defaults_match_oracle) */
	mm2o_opt_init((mm2o_opt_t*)opt);
	return MB_OK;
}

/* ---- index ---- */
int mb_index_build(int device, int n_seq, const char *const *names, const uint8_t *const *seqs, const int64_t *lens, int w, int k, mb_index_t **out)
{
	int n_threads = (int)sysconf(_SC_NPROCESSORS_ONLN);
	(void)device;
	if (n_seq <= 0 || !names || !seqs || !lens || !out) return fail(MB_ERR_ARG, "bad arguments");
	*out = (mb_index_t*)calloc(1, sizeof(mb_index_t));
	(*out)->mi = mm2o_idx_build_mt(n_seq, (const char**)names, (const char**)seqs, lens, w, k, n_threads < 1 ? 1 : n_threads);
	if (!(*out)->mi) { free(*out); *out = 0; return fail(MB_ERR_ARG, "index building failed"); }
	return MB_OK;
}
void mb_index_free(mb_index_t *ix) { if (ix) { mm2o_idx_destroy(ix->mi); free(ix); } }
int mb_index_n_seq(const mb_index_t *ix) { return ix ? (int)ix->mi->n_seq : 0; }
const char *mb_index_seq_name(const mb_index_t *ix, int rid) { return ix && rid >= 0 && (uint32_t)rid < ix->mi->n_seq ? ix->mi->seq[rid].name : 0; }
int64_t mb_index_seq_len(const mb_index_t *ix, int rid) { return ix && rid >= 0 && (uint32_t)rid < ix->mi->n_seq ? (int64_t)ix->mi->seq[rid].len : -1; }
int mb_index_mid_occ(const mb_index_t *ix)
{
	mm2o_opt_t o;
	if (!ix) return 0;
	mm2o_opt_init(&o);
	return mm2o_idx_cal_max_occ(ix->mi, o.mid_occ_frac);
}
int mb_index_kw(const mb_index_t *ix, int *k, int *w)
{
	if (!ix) return fail(MB_ERR_ARG, "bad arguments");
	if (k) *k = ix->mi->k;
	if (w) *w = ix->mi->w;
	return MB_OK;
}

/* ---- batch mapping ---- */
void mb_hits_free(mb_hits_t *h)
{
	if (!h) return;
	free(h->fields); free(h->cigar_off); free(h->cigar); free(h->rep_len); free(h->hit_off); free(h->read_off);
	free(h);
}

int mb_map_batch_ex(mb_index_t *ix, const mb_opt_t *opt, const uint8_t *cat, const int64_t *off, int32_t n_reads, int want, mb_hits_t **out, mb_stats_t *stats)
{
	mm2o_opt_t o;
	mm2o_result_t **res;
	mb_hits_t *H;
	int32_t *rows;
	int64_t n, nc = 0, i, c;
	int f, n_threads = (int)sysconf(_SC_NPROCESSORS_ONLN);
	(void)want;
	if (!ix || !opt || !off || !out || n_reads < 0 || (n_reads > 0 && (off[0] != 0 || (!cat && off[n_reads] > 0)))) return fail(MB_ERR_ARG, "bad arguments");
	for (i = 0; i < n_reads; ++i) if (off[i + 1] < off[i]) return fail(MB_ERR_ARG, "offsets must be non-decreasing");
	memcpy(&o, opt, sizeof(o));
	if (o.mid_occ <= 0) mm2o_mapopt_update(&o, ix->mi);
	res = (mm2o_result_t**)calloc(n_reads > 0 ? n_reads : 1, sizeof(*res));
	mm2o_map_batch(ix->mi, &o, n_reads, (const char*)cat, off, n_threads < 1 ? 1 : n_threads, res);
	n = mm2o_batch_export(res, n_reads, 0, 0, 0, &nc);
	H = (mb_hits_t*)calloc(1, sizeof(*H));
	H->n = n, H->n_reads = n_reads, H->n_cigar = nc;
	rows = (int32_t*)malloc((size_t)(n > 0 ? n : 1) * 22 * 4);
	H->fields = (int32_t*)malloc((size_t)(n > 0 ? n : 1) * ABI_NF * 4);
	H->cigar_off = (int64_t*)malloc((size_t)(n > 0 ? n : 1) * 8);
	H->cigar = (uint32_t*)malloc((size_t)(nc > 0 ? nc : 1) * 4);
	H->rep_len = (int32_t*)malloc((size_t)(n_reads > 0 ? n_reads : 1) * 4);
	H->hit_off = (int64_t*)malloc((size_t)(n_reads + 1) * 8);
	H->read_off = (int64_t*)malloc((size_t)(n_reads + 1) * 8);
	mm2o_batch_export(res, n_reads, rows, H->hit_off, H->cigar, &nc);
	memcpy(H->read_off, off, (size_t)(n_reads + 1) * 8);
	if (n_reads == 0) H->read_off[0] = 0;
	/* row-major oracle records -> the ABI's struct of arrays; read_idx first, CIGAR offsets as a running sum */
	for (i = 0, c = 0; i < n; ++i) {
		for (f = 0; f < 22; ++f) H->fields[(size_t)(f + 1) * n + i] = rows[i * 22 + f];
		H->cigar_off[i] = c;
		c += rows[i * 22 + 21];
	}
	for (i = 0; i < n_reads; ++i) {
		int64_t h;
		H->rep_len[i] = res[i]->rep_len;
		for (h = H->hit_off[i]; h < H->hit_off[i + 1]; ++h) H->fields[h] = (int32_t)i;
	}
	if (stats) {
		memset(stats, 0, sizeof(*stats));
		stats->n_reads = n_reads, stats->n_bases = n_reads > 0 ? off[n_reads] : 0, stats->n_hits = n;
		for (i = 0; i < n_reads; ++i) {
			stats->n_mini += res[i]->n_mini, stats->n_anchor += res[i]->n_anchor, stats->chain_cells += res[i]->chain_cells;
			stats->dp_cells += res[i]->dp_cells, stats->n_dp_tasks += res[i]->n_dp_calls;
		}
	}
	for (i = 0; i < n_reads; ++i) mm2o_result_destroy(res[i]);
	free(res); free(rows);
	*out = H;
	return MB_OK;
}
int mb_map_batch(mb_index_t *ix, const mb_opt_t *opt, const uint8_t *cat, const int64_t *off, int32_t n_reads, mb_hits_t **out, mb_stats_t *stats)
{
	return mb_map_batch_ex(ix, opt, cat, off, n_reads, 3, out, stats);
}

int64_t mb_hits_n(const mb_hits_t *h) { return h ? h->n : 0; }
const int32_t *mb_hits_field(const mb_hits_t *h, const char *name)
{
	int f;
	if (!h || !name) return 0;
	for (f = 0; f < ABI_NF; ++f) if (strcmp(name, ABI_NAMES[f]) == 0) return h->fields + (size_t)f * h->n;
	return 0;
}
const int64_t *mb_hits_cigar_off(const mb_hits_t *h) { return h ? h->cigar_off : 0; }
const uint32_t *mb_hits_cigar_pool(const mb_hits_t *h, int64_t *n) { if (!h) return 0; if (n) *n = h->n_cigar; return h->cigar; }
const int32_t *mb_hits_rep_len(const mb_hits_t *h, int64_t *n_reads) { if (!h) return 0; if (n_reads) *n_reads = h->n_reads; return h->rep_len; }

/* ---- monica's counting, stated the way aligner.py does it: keep primary hits with mapq >= mapping_quality (:194,216); no hit ->
 * unmapped (:209-211); one hit -> it; several -> best_hit (:328-339): the smallest float(NM)/mlen, ambiguous (0) iff that
 * minimum occurs twice or more; then Counter[accession] += 1 | len(read) | mlen by mode (:247-263), nothing for any other mode. */
int mb_count(mb_index_t *ix, const mb_hits_t *h, int32_t mapq_min, int mode, int64_t *counts, int64_t *n_class, int8_t *read_class, int64_t *read_best)
{
	const int32_t *rid, *mapq, *mlen, *nm, *pri;
	int64_t cls_n[3] = { 0, 0, 0 };
	int32_t r;
	if (!ix || !h) return fail(MB_ERR_ARG, "bad arguments");
	rid = mb_hits_field(h, "rid"), mapq = mb_hits_field(h, "mapq"), mlen = mb_hits_field(h, "mlen"), nm = mb_hits_field(h, "nm"), pri = mb_hits_field(h, "is_primary");
	if (counts) memset(counts, 0, (size_t)ix->mi->n_seq * 8);
	for (r = 0; r < h->n_reads; ++r) {
		int64_t k, best = -1;
		int n_kept = 0, n_min = 0, cls;
		double vmin = 0.0;
		for (k = h->hit_off[r]; k < h->hit_off[r + 1]; ++k) {
			double v;
			if (!(pri[k] && mapq[k] >= mapq_min)) continue;
			v = (double)nm[k] / (double)mlen[k];
			if (n_kept == 0 || v < vmin) vmin = v, n_min = 1, best = k;
			else if (v == vmin) ++n_min, best = k;
			++n_kept;
		}
		if (n_kept == 0) cls = 0;
		else if (n_kept >= 2 && n_min >= 2) cls = 2, best = -1;
		else cls = 1;
		if (read_class) read_class[r] = (int8_t)cls;
		if (read_best) read_best[r] = best;
		++cls_n[cls == 1 ? 0 : cls == 0 ? 1 : 2];
		if (cls == 1 && counts) {
			if (mode == 0) counts[rid[best]] += 1;
			else if (mode == 1) counts[rid[best]] += h->read_off[r + 1] - h->read_off[r];
			else if (mode == 2) counts[rid[best]] += mlen[best];
		}
	}
	if (n_class) memcpy(n_class, cls_n, sizeof(cls_n));
	return MB_OK;
}

/* ---- per-stage entry points ---- */
int mb_sketch(int device, const uint8_t *cat, const int64_t *off, int32_t n_reads, int w, int k, uint64_t *out_xy, int64_t cap, int64_t *out_off)
{
	int64_t n = 0;
	int32_t i;
	(void)device;
	if (!off || !out_xy || !out_off || n_reads < 0) return fail(MB_ERR_ARG, "bad arguments");
	for (i = 0; i < n_reads; ++i) {
		int64_t m = mm2o_sketch_buf((const char*)cat + off[i], (int)(off[i + 1] - off[i]), w, k, (uint32_t)i, out_xy + 2 * n, cap - n);
		out_off[i] = n;
		if (m < 0 || n + m > cap) return fail(MB_ERR_OVERFLOW, "minimizer buffer too small");
		n += m;
	}
	out_off[n_reads] = n;
	return MB_OK;
}

int mb_dp_batch(int device, const mb_opt_t *opt, mb_dp_task_t *tasks, int64_t n_tasks, const uint8_t *seqpool, int64_t n_seqpool, uint32_t *cigar_pool, int64_t n_cigar_pool)
{
	int8_t mat[25];
	int64_t i;
	(void)device; (void)n_seqpool;
	if (!opt || (n_tasks > 0 && (!tasks || !seqpool || !cigar_pool))) return fail(MB_ERR_ARG, "bad arguments");
	mm2o_gen_simple_mat(5, mat, (int8_t)opt->a, (int8_t)opt->b, (int8_t)opt->sc_ambi);
	for (i = 0; i < n_tasks; ++i) {
		mb_dp_task_t *t = &tasks[i];
		ksw_extz_t ez;
		memset(&ez, 0, sizeof(ez));
		mm2o_ksw_extd2(t->qlen, seqpool + t->q_off, t->tlen, seqpool + t->t_off, 5, mat, (int8_t)opt->q, (int8_t)opt->e, (int8_t)opt->q2, (int8_t)opt->e2,
		               t->w, t->zdrop, t->end_bonus, t->flag, &ez);
		t->score = ez.score, t->max = ez.max, t->max_q = ez.max_q, t->max_t = ez.max_t, t->mqe = ez.mqe, t->mqe_t = ez.mqe_t;
		t->zdropped = ez.zdropped, t->reach_end = ez.reach_end, t->n_cigar = ez.n_cigar;
		if (t->cigar_off + ez.n_cigar > n_cigar_pool) { free(ez.cigar); return fail(MB_ERR_OVERFLOW, "CIGAR pool too small"); }
		if (ez.n_cigar) memcpy(cigar_pool + t->cigar_off, ez.cigar, (size_t)ez.n_cigar * 4);
		free(ez.cigar);
	}
	return MB_OK;
}

int mb_ll_batch(int device, const mb_opt_t *opt, mb_ll_task_t *tasks, int64_t n_tasks, const uint8_t *seqpool, int64_t n_seqpool)
{
	int8_t mat[25];
	int64_t i;
	(void)device; (void)n_seqpool;
	if (!opt || (n_tasks > 0 && (!tasks || !seqpool))) return fail(MB_ERR_ARG, "bad arguments");
	mm2o_gen_simple_mat(5, mat, (int8_t)opt->a, (int8_t)opt->b, (int8_t)opt->sc_ambi);
	for (i = 0; i < n_tasks; ++i) {
		mb_ll_task_t *t = &tasks[i];
		int qe = -1, te = -1;
		t->score = mm2o_ksw_ll_i16(t->qlen, seqpool + t->q_off, t->tlen, seqpool + t->t_off, 5, mat, opt->q, opt->e, &qe, &te);
		t->qe = qe, t->te = te;
	}
	return MB_OK;
}
