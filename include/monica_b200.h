/*
 * monica_b200.h -- C ABI of the B200-native mapper that replaces mappy==2.17 under monica's aligner.
 *
 * Loaded with ctypes (monica_b200/_lib.py).  Plain pointers and sizes only; no torch types.
 * Every function returns 0 on success or a negative error code; mb_last_error() returns the
 * thread-local message.  There is NO CPU fallback: every compute entry point needs a CUDA device
 * and fails with MB_ERR_CUDA otherwise.
 *
 * What each entry point replaces in the reference (paths relative to /root/reference):
 *   mb_index_build / mb_index_build_fasta / mb_index_save
 *        mappy.Aligner(fn_idx_in=<fna.gz>, preset='map-ont', best_n=15, fn_idx_out=<mmi>)
 *        monica/genomes/aligner.py:45-46 (indexer, :31-53)
 *   mb_index_load
 *        mappy.Aligner(fn_idx_in=<mmi>)            monica/genomes/aligner.py:59 (index_loader, :56-62)
 *   mb_map_batch / mb_map_packed (+ mb_hits_*)
 *        for hit in index.map(str(seq_record.seq)) monica/genomes/aligner.py:193,215
 *        fields hit.is_primary .mapq .ctg .NM .mlen monica/genomes/aligner.py:194-195,216-217
 *   mb_count
 *        best_hit + taxon/accession Counter update monica/genomes/aligner.py:225-263,328-339
 *   mb_allreduce_counts
 *        the additive merge of per-shard tallies    monica/genomes/aligner.py:282-302 (alignment_update)
 *   mb_sketch / mb_seed / mb_chain / mb_dp_batch / mb_ll_batch
 *        per-stage entry points for parity tests (no reference counterpart; stages of mm_map_frag)
 */
#ifndef MONICA_B200_H
#define MONICA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MB_OK            0
#define MB_ERR_ARG      -1
#define MB_ERR_CUDA     -2
#define MB_ERR_IO       -3
#define MB_ERR_OVERFLOW -4
#define MB_ERR_NOMEM    -5

typedef struct mb_index mb_index_t;   /* opaque: host metadata + HBM-resident index */
typedef struct mb_hits  mb_hits_t;    /* opaque: library-owned result of one mb_map_batch call */

/* Mapping options; mb_opt_init() sets minimap2-2.17 mm_mapopt_init() defaults (== map-ont apart from k). */
typedef struct {
	int32_t seed;
	float   mid_occ_frac;
	int32_t min_cnt, min_chain_score, bw, max_gap, max_gap_ref, max_chain_skip, max_chain_iter;
	float   mask_level, pri_ratio;
	int32_t best_n;
	int32_t max_join_long, max_join_short, min_join_flank_sc;
	float   min_join_flank_ratio;
	int32_t a, b, q, e, q2, e2, sc_ambi, zdrop, zdrop_inv, end_bonus, min_dp_max, min_ksw_len;
	float   max_clip_ratio;
	int64_t max_sw_mat;
	int32_t mid_occ;          /* <=0: derive from the index (mm_mapopt_update) */
} mb_opt_t;

/* Work counters of one batch (the algorithmic units bench.py's roofline uses). */
typedef struct {
	int64_t n_reads, n_bases, n_mini, n_anchor, n_regs, n_dp_tasks, n_dp_pass2, dp_cells, n_hits, n_rounds;
	float   ms_sketch, ms_seed, ms_sort, ms_chain, ms_glue, ms_dp, ms_post, ms_total, ms_h2d, ms_d2h;
	int64_t n_launches;       /* kernels launched for this batch */
	float   ms_kdp;           /* CUDA-event time of the k_dp launches alone (the roofline kernel) */
	int32_t n_kdp;            /* number of DP kernel launches */
	float   ms_kdp_fast;      /* wall time (CUDA events) of the k_dp_fast launches, which overlap each other on 3 streams */
	float   ms_kdp_exact;     /* ... and the exact ksw_extd2 block emulation (k_dp) */
	int64_t n_fast_tasks, n_exact_tasks;
	int64_t chain_cells;      /* predecessors visited by the chaining DP (the oracle counts the same loop) */
	int64_t dp_cells_exact;   /* DP cells evaluated by k_dp (the rest of dp_cells went through k_dp_fast) */
	int64_t n_kdp_fast;       /* k_dp_fast launches (one per non-empty column class and pass) */
	int64_t n_ext_tasks, dp_cells_ext; /* end extensions through k_dp_ext */
	float   ms_kdp_ext; int32_t n_inv; /* inversion hits produced (mm_align1_inv) */
	int64_t arena_bytes;      /* device scratch handed out for this batch (largest piece) */
	int32_t n_pieces, pad_;   /* sequential pieces the batch was cut into (memory budget) */
	int64_t n_band_tasks, dp_cells_band; /* large / band-limited gap fills through k_dp_band */
	float   ms_kdp_band; int32_t pad2_;
} mb_stats_t;

const char *mb_last_error(void);
int  mb_device_count(void);
int  mb_opt_init(mb_opt_t *opt);

/* ---- index ---- */
int  mb_index_build(int device, int n_seq, const char *const *names, const uint8_t *const *seqs, const int64_t *lens,
                    int w, int k, mb_index_t **out);
int  mb_index_build_fasta(int device, const char *fasta_gz_path, int w, int k, mb_index_t **out);
int  mb_index_save(const mb_index_t *idx, const char *mmi_path);
int  mb_index_load(int device, const char *mmi_path, mb_index_t **out);
void mb_index_free(mb_index_t *idx);
int  mb_index_n_seq(const mb_index_t *idx);
const char *mb_index_seq_name(const mb_index_t *idx, int rid);
int64_t mb_index_seq_len(const mb_index_t *idx, int rid);
int  mb_index_mid_occ(const mb_index_t *idx);
int  mb_index_kw(const mb_index_t *idx, int *k, int *w);
int64_t mb_index_n_minimizers(const mb_index_t *idx);
int64_t mb_index_hbm_bytes(const mb_index_t *idx);

/* ---- batch mapping: host buffers in, library-owned host hits out ----
 * cat: concatenated ASCII reads; off[n_reads+1]: byte offsets.  Thread-safe per (index, calling thread):
 * each calling thread gets its own CUDA stream + scratch. */
int  mb_map_batch(mb_index_t *idx, const mb_opt_t *opt, const uint8_t *cat, const int64_t *off, int32_t n_reads,
                  mb_hits_t **out, mb_stats_t *stats);
/* same with a choice of what is copied back: want bit 0 = hit fields, bit 1 = CIGARs (mb_map_batch == want 3).  monica reads only
 * (ctg, NM, mlen) of each hit (aligner.py:195,217), so its path skips the CIGAR copy */
int  mb_map_batch_ex(mb_index_t *idx, const mb_opt_t *opt, const uint8_t *cat, const int64_t *off, int32_t n_reads, int want,
                     mb_hits_t **out, mb_stats_t *stats);
/* device-resident variant used by bench.py's `value` leg: reads must already be uploaded with mb_reads_upload */
typedef struct mb_reads mb_reads_t;
int  mb_reads_upload(mb_index_t *idx, const uint8_t *cat, const int64_t *off, int32_t n_reads, mb_reads_t **out);
void mb_reads_free(mb_reads_t *r);
int  mb_map_resident(mb_index_t *idx, const mb_opt_t *opt, mb_reads_t *reads, int want_hits, mb_hits_t **out, mb_stats_t *stats);

/* packed reads: the batch reduced ONCE on the host (n_threads threads; <= 0: all cores) to 2-bit words -- base i of the
 * concatenation in bits [2(i&15), 2(i&15)+2) of word i>>4, nt4 codes A/a 0, C/c 1, G/g 2, T/t/U/u 3 -- plus the (start, length)
 * runs of every other character, in page-locked memory when a device is present.  mb_map_packed == mb_map_batch_ex on such
 * a batch: it uploads 0.25 B/base instead of 1 and expands the words on the device to the byte codes the kernels read, so
 * results are identical.  mappy takes a Python str per read (aligner.py:193,215); nothing after the nt4 table sees the
 * characters.  mb_reads_pack / mb_fastq_pack need no device. */
typedef struct mb_packed mb_packed_t;
typedef struct mb_fastq mb_fastq_t;
int  mb_reads_pack(const uint8_t *cat, const int64_t *off, int32_t n_reads, int n_threads, mb_packed_t **out);
int  mb_fastq_pack(const mb_fastq_t *fq, int n_threads, mb_packed_t **out);
void mb_packed_free(mb_packed_t *p);
int64_t mb_packed_upload_bytes(const mb_packed_t *p);   /* bytes one mb_map_packed moves host -> device */
const uint32_t *mb_packed_words(const mb_packed_t *p, int64_t *n_words, const int64_t **intervals, int64_t *n_intervals);
int  mb_map_packed(mb_index_t *idx, const mb_opt_t *opt, const mb_packed_t *reads, int want, mb_hits_t **out, mb_stats_t *stats);

/* hits: struct-of-arrays, n = mb_hits_n(); arrays stay valid until mb_hits_free */
int64_t mb_hits_n(const mb_hits_t *h);
const int32_t *mb_hits_field(const mb_hits_t *h, const char *name);
/* names: read_idx rid rev qs qe rs re mapq mlen blen nm dp_max dp_max2 score score0 cnt subsc n_sub
 *        id parent is_primary sam_pri n_cigar */
const int64_t *mb_hits_cigar_off(const mb_hits_t *h);   /* [n] offsets into the cigar pool */
const uint32_t *mb_hits_cigar_pool(const mb_hits_t *h, int64_t *n);
const int32_t *mb_hits_rep_len(const mb_hits_t *h, int64_t *n_reads); /* per read */
void mb_hits_free(mb_hits_t *h);

/* ---- counting: monica's hit filter + best_hit + per-target sum, on the device ----
 * mode: 0 basic (+1), 1 query_length (+len(read)), 2 matching (+mlen)   (aligner.py:247-263)
 * counts[n_seq] int64, indexed by rid (contigs of one genome share a name; the host folds by name).
 * n_class[3] = {mapped, unmapped, ambiguous} reads.  read_class[n_reads]: 0 unmapped, 1 mapped, 2 ambiguous;
 * read_best[n_reads]: hit index chosen or -1.  Any pointer may be NULL. */
int  mb_count(mb_index_t *idx, const mb_hits_t *h, int32_t mapq_min, int mode, int64_t *counts, int64_t *n_class,
              int8_t *read_class, int64_t *read_best);
/* same, on the device-resident hits of the calling thread's LAST mb_map_batch / mb_map_resident on this index (no host
 * round trip of the hit arrays; valid until that thread's next mapping call) */
int  mb_count_last(mb_index_t *idx, int32_t mapq_min, int mode, int64_t *counts, int64_t *n_class);
/* device pointer to the int64[n_seq] count vector of the LAST mb_count on this thread (for the NCCL allreduce) */
void *mb_count_device_ptr(mb_index_t *idx);
int  mb_count_fetch(mb_index_t *idx, int64_t *counts);
/* normalizer(alignment, genomes_length)  monica/genomes/aligner.py:305-319, on the device-resident count vector of the LAST
 * mb_count on this thread (after the multi-GPU all-reduce, if any).  group[n_seq] maps each contig to its genome (accession)
 * or -1, group_len[n_groups] = genome lengths, order[n_order] = the groups in the order the reference's nested dict visits
 * them (float addition order of the sample total).  bpm[n_groups] = (count/len) / sum(count/len). */
int  mb_normalize_last(mb_index_t *idx, const int32_t *group, int32_t n_groups, const double *group_len,
                       const int32_t *order, int32_t n_order, double *bpm);

/* ---- multi-GPU: the one collective of the path ----
 * Reads shard over one process per GPU with the index replicated; per-sample tallies merge additively (Counter.update,
 * monica/genomes/aligner.py:288-292), so the only exchange is ONE all-reduce of the int64[n_seq] count vector (+ the three
 * read-class counters), done with NCCL on the mapping stream on the device-resident vector of the calling thread's last
 * mb_count / mb_count_last.  libnccl is bound at run time (dlopen "libnccl.so.2"); without it these calls fail loudly.
 * mb_comm_unique_id: rank 0 makes the 128-byte NCCL id, the caller hands it to the other ranks over any channel. */
typedef struct mb_comm mb_comm_t;
int  mb_comm_unique_id(uint8_t id[128]);
int  mb_comm_init(int device, int rank, int world, const uint8_t id[128], mb_comm_t **out);
void mb_comm_free(mb_comm_t *comm);
int  mb_allreduce_counts(mb_index_t *idx, mb_comm_t *comm, int64_t *counts, int64_t *n_class);

/* ---- FASTQ ingest and routed writers (host side, no device needed) ----
 * mb_fastq_load      for seq_record in SeqIO.parse(sample, 'fastq')            monica/genomes/aligner.py:191,212
 * mb_fastq_route     SeqIO.write(seq_record, <mapped|unmapped|ambiguous|focus>) monica/genomes/aligner.py:232,236,243,265
 * The sequences come back in the concatenated layout mb_map_batch takes. */
int  mb_fastq_load(const char *path, mb_fastq_t **out);              /* plain or gzip; multi-line records accepted */
/* MB_ERR_IO with a message starting "malformed FASTQ" where Bio.SeqIO raises ValueError (no '+' line, unequal sequence and
 * quality lengths, a record that does not start with '@'), and for a gzip stream that ends before its trailer. */
int64_t mb_fastq_n(const mb_fastq_t *fq);
const uint8_t *mb_fastq_seqs(const mb_fastq_t *fq, const int64_t **off);
const char *mb_fastq_header(const mb_fastq_t *fq, int64_t i, int64_t *len, int32_t *id_len);
int  mb_fastq_ids_unique(const mb_fastq_t *fq);
int  mb_fastq_route(const mb_fastq_t *fq, const int8_t *dest, const char *const *new_id, const uint8_t *focus,
                    const char *mapped_path, const char *unmapped_path, const char *ambiguous_path, const char *focus_path);
/* the same with the mapped reads' new ids given as a table: read i takes ids[target[i]] (one entry per contig / tax unit) */
int  mb_fastq_route_targets(const mb_fastq_t *fq, const int8_t *dest, const int32_t *target, const char *const *ids, int32_t n_ids, const uint8_t *focus,
                            const char *mapped_path, const char *unmapped_path, const char *ambiguous_path, const char *focus_path);
void mb_fastq_free(mb_fastq_t *fq);

/* ---- database builder (host side, no device needed) ----
 * mb_db_build        builder(genomes_chunk, databases_path, database_name, database_number)  monica/genomes/database.py:52-67
 * Writes one database<N>.fna.gz: every FASTA record of genome i re-headed new_headers[i] ("<tax_unit>:<accession>") the way
 * the reference's SeqIO.parse -> SeqIO.write round trip does (title "<new id> <old title>", 60-column lines);
 * genome_len[i] = bases of genome i (current_genomes_length.pkl). */
int  mb_db_build(const char *out_path, int32_t n_genomes, const char *const *genome_paths, const char *const *new_headers,
                 int64_t *genome_len);

/* ---- per-stage entry points (parity tests) ---- */
/* minimizers of each read: out_xy[2*cap], out_off[n_reads+1]; y carries the read index in its high 32 bits */
int  mb_sketch(int device, const uint8_t *cat, const int64_t *off, int32_t n_reads, int w, int k,
               uint64_t *out_xy, int64_t cap, int64_t *out_off);
/* sorted anchors of each read */
int  mb_seed(mb_index_t *idx, const mb_opt_t *opt, const uint8_t *cat, const int64_t *off, int32_t n_reads,
             uint64_t *out_xy, int64_t cap, int64_t *out_off, int32_t *rep_len);
/* chaining DP arrays f,p,v for caller-provided sorted anchors (per read), plus chains */
int  mb_chain(int device, const mb_opt_t *opt, const uint64_t *anchors_xy, const int64_t *off, int32_t n_reads,
              int32_t *f, int32_t *p, int32_t *v,
              uint64_t *chained_xy, int64_t *chained_off, uint64_t *u, int64_t *u_off);
/* batch of stand-alone ksw_extd2 problems on nt4-coded sequences */
typedef struct {
	int32_t qlen, tlen, w, zdrop, end_bonus, flag;
	int64_t q_off, t_off;        /* into seqpool */
	/* outputs */
	int32_t score, max, max_q, max_t, mqe, mqe_t, zdropped, reach_end, n_cigar;
	int64_t cigar_off;           /* into cigar_pool; capacity qlen+tlen+1 per task, set by caller */
} mb_dp_task_t;
int  mb_dp_batch(int device, const mb_opt_t *opt, mb_dp_task_t *tasks, int64_t n_tasks, const uint8_t *seqpool, int64_t n_seqpool,
                 uint32_t *cigar_pool, int64_t n_cigar_pool);

/* batch of stand-alone local alignments (ksw_ll_i16 as mm_test_zdrop / mm_align1_inv call it) on nt4-coded sequences:
 * score, query end and target end with upstream's striped-layout tie rules */
typedef struct {
	int32_t qlen, tlen;
	int64_t q_off, t_off;        /* into seqpool */
	int32_t score, qe, te, pad_; /* outputs */
} mb_ll_task_t;
int  mb_ll_batch(int device, const mb_opt_t *opt, mb_ll_task_t *tasks, int64_t n_tasks, const uint8_t *seqpool, int64_t n_seqpool);

/* ---- measurement helpers ----
 * mb_stream: the cudaStream_t on which the calling thread's batches of this index run (so a caller can bracket K batches
 * with CUDA events recorded on the launching stream, or enqueue its NCCL all-reduce of mb_count_device_ptr() behind them). */
void *mb_stream(mb_index_t *idx);
/* sustained INT32 add/max issue rate of the device (roofline denominator for K3/K4) ---- */
int  mb_int_peak(int device, double *tera_int_ops_per_s);   /* reference no counterpart: roofline denominator for K3/K4 */

/* mm_set_mapq's logf (the float that minimap2's `(int)(... * logf(...))` truncates) against the caller's libm, bit for bit:
 * x runs over the n consecutive float bit patterns from first_bits, expected[i] = logf(x_i) of the host libm the CPU path
 * links.  device >= 0: evaluated by the GPU routine of the MAPQ kernel; device < 0: by its host twin.  n_bad = values whose
 * bits differ, first_bad = the smallest differing input pattern (0xffffffff if none).  No reference counterpart (test hook). */
int  mb_logf_sweep(int device, uint32_t first_bits, int64_t n, const float *expected, int64_t *n_bad, uint32_t *first_bad);

#ifdef __cplusplus
}
#endif
#endif
