/*
 * mbsynth.c -- seeded synthetic inputs for bench.py at BASELINE.json's sizes (measurement infrastructure, not product code,
 * not part of the oracle).  numpy generates ~10 MB/s of simulated reads per core; the 1 M-read / 4 Gb configurations need
 * gigabytes, so the same kind of data is produced here by a counter-based generator, multi-threaded, deterministic in
 * (seed, item index) and independent of the thread count.
 *
 *   mbs_genome     uniform-random ACGT
 *   mbs_mutate     a "strain" copy: i.i.d. substitutions / insertions / deletions
 *   mbs_reads      ONT-like reads: log-normal lengths, i.i.d. sub / ins / del errors, random strand, plus a stated mix of
 *                  read classes that reach the rarely-taken paths of the mapper (SURVEY.md 8(d) asks for a workload where
 *                  MAPQ < 60, best_hit ties and Z-drop splits occur):
 *                    class 1  a block of unrelated sequence replaces 500-1500 bp in the middle (Z-drop split, 2nd DP pass)
 *                    class 2  an internal block of 600-1500 bp is reverse-complemented (inversion hit)
 *                    class 3  two error-free fragments of different loci joined (two MAPQ-60 primaries with NM = 0:
 *                             monica's best_hit calls the read ambiguous)
 *                    class 4  unrelated sequence only (unmapped)
 * Build: gcc -O2 -fPIC -shared -pthread (see __graft_entry__.build).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <pthread.h>

static inline uint64_t splitmix(uint64_t *s)
{
	uint64_t z = (*s += 0x9E3779B97F4A7C15ULL);
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
	return z ^ (z >> 31);
}
static inline uint64_t mix2(uint64_t a, uint64_t b) { uint64_t s = a * 0xD6E8FEB86659FD93ULL + b; return splitmix(&s); }
static inline double u01(uint64_t *s) { return (double)(splitmix(s) >> 11) * (1.0 / 9007199254740992.0); }
static const char ACGT[4] = { 'A', 'C', 'G', 'T' };
static inline char comp(char c) { return c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : c == 'T' ? 'A' : c; }

/* ---- generic parallel-for over blocks ---- */
typedef struct { void (*fn)(void*, int64_t, int64_t); void *arg; int64_t n, grain; volatile int64_t next; pthread_mutex_t mu; } pf_t;
static void *pf_worker(void *a_)
{
	pf_t *p = (pf_t*)a_;
	for (;;) {
		int64_t lo;
		pthread_mutex_lock(&p->mu); lo = p->next; p->next += p->grain; pthread_mutex_unlock(&p->mu);
		if (lo >= p->n) break;
		p->fn(p->arg, lo, lo + p->grain < p->n ? lo + p->grain : p->n);
	}
	return 0;
}
static void parallel_for(int threads, int64_t n, int64_t grain, void (*fn)(void*, int64_t, int64_t), void *arg)
{
	pf_t p; pthread_t th[256]; int i;
	if (threads < 1) threads = 1;
	if (threads > 256) threads = 256;
	p.fn = fn, p.arg = arg, p.n = n, p.grain = grain < 1 ? 1 : grain, p.next = 0;
	pthread_mutex_init(&p.mu, 0);
	for (i = 0; i < threads; ++i) pthread_create(&th[i], 0, pf_worker, &p);
	for (i = 0; i < threads; ++i) pthread_join(th[i], 0);
	pthread_mutex_destroy(&p.mu);
}

/* ---- genomes ---- */
typedef struct { uint64_t seed; char *out; int64_t len; } gen_t;
static void gen_block(void *a_, int64_t lo, int64_t hi)
{
	gen_t *g = (gen_t*)a_;
	int64_t b;
	for (b = lo; b < hi; ++b) { /* 32 bases per block from one 64-bit draw */
		uint64_t r = mix2(g->seed, (uint64_t)b);
		int64_t i = b * 32, e = i + 32 < g->len ? i + 32 : g->len;
		for (; i < e; ++i, r >>= 2) g->out[i] = ACGT[r & 3];
	}
}
void mbs_genome(uint64_t seed, int64_t len, char *out, int threads)
{
	gen_t g; g.seed = seed, g.out = out, g.len = len;
	parallel_for(threads, (len + 31) / 32, 1 << 15, gen_block, &g);
}

/* strain copy; returns the new length (out needs len * (1 + ins) + slack) */
int64_t mbs_mutate(uint64_t seed, const char *src, int64_t len, double sub, double ins, double del, char *out, int64_t cap)
{
	uint64_t s = mix2(seed, 0x5157A1B5ULL);
	int64_t i, n = 0;
	for (i = 0; i < len && n + 2 < cap; ++i) {
		double r = u01(&s);
		if (r < sub) out[n++] = ACGT[((src[i] == 'A' ? 0 : src[i] == 'C' ? 1 : src[i] == 'G' ? 2 : 3) + 1 + (int)(splitmix(&s) % 3)) & 3];
		else if (r < sub + del) continue;
		else {
			out[n++] = src[i];
			if (r < sub + del + ins) out[n++] = ACGT[splitmix(&s) & 3];
		}
	}
	return n;
}

/* ---- reads ---- */
typedef struct {
	uint64_t seed;
	const char *gcat; const int64_t *goff; int n_genomes;
	int64_t n_reads;
	double mu, sigma; int64_t min_len, max_len;
	double sub, ins, del;
	double f_junkins, f_inv, f_chim, f_junk;
	char *out; int64_t *off;      /* off[i+1] - off[i] = length of read i (pass 1 writes lengths into off[i+1]) */
	int64_t first;                /* pass 2 writes reads first .. first+n_reads-1, the first of them at out[0] */
	int8_t *cls;                  /* read class, may be NULL */
	int pass;
} reads_t;

static inline int64_t draw_len(uint64_t *s, const reads_t *p)
{
	double u1 = u01(s), u2 = u01(s), z, L;
	if (u1 < 1e-300) u1 = 1e-300;
	z = sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
	L = exp(p->mu + p->sigma * z);
	if (L < (double)p->min_len) L = (double)p->min_len;
	if (p->max_len > 0 && L > (double)p->max_len) L = (double)p->max_len;
	return (int64_t)L;
}
/* a locus of `len` bases inside one genome, genomes weighted by length */
static inline void draw_locus(uint64_t *s, const reads_t *p, int64_t *len, int64_t *pos)
{
	const int64_t total = p->goff[p->n_genomes];
	for (;;) {
		int64_t g0 = (int64_t)(u01(s) * (double)total);
		int lo = 0, hi = p->n_genomes; /* goff[lo] <= g0 < goff[hi] */
		while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (p->goff[mid] <= g0) lo = mid; else hi = mid; }
		{
			int64_t glen = p->goff[lo + 1] - p->goff[lo];
			if (*len > glen) *len = glen;
			*pos = p->goff[lo] + (int64_t)(u01(s) * (double)(glen - *len + 1));
			if (*pos + *len <= p->goff[lo + 1]) return;
		}
	}
}
/* emit src[0..n) (optionally reverse-complemented) with errors; w == NULL counts only */
static inline int64_t emit_noisy(uint64_t *s, const reads_t *p, const char *src, int64_t n, int rev, char *w)
{
	int64_t i, k = 0;
	const double sub = p->sub, del = p->del, ins = p->ins;
	for (i = 0; i < n; ++i) {
		char c = rev ? comp(src[n - 1 - i]) : src[i];
		double r = u01(s);
		if (r < sub) { uint64_t t = splitmix(s) % 3; int b = (c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : 3); if (w) w[k] = ACGT[(b + 1 + (int)t) & 3]; ++k; }
		else if (r < sub + del) continue;
		else {
			if (w) w[k] = c; ++k;
			if (r < sub + del + ins) { uint64_t t = splitmix(s); if (w) w[k] = ACGT[t & 3]; ++k; }
		}
	}
	return k;
}
static inline int64_t emit_random(uint64_t *s, int64_t n, char *w)
{
	int64_t i;
	for (i = 0; i < n; i += 32) {
		uint64_t r = splitmix(s);
		int64_t e = i + 32 < n ? i + 32 : n, j;
		for (j = i; j < e; ++j, r >>= 2) if (w) w[j] = ACGT[r & 3];
	}
	return n;
}
static int64_t one_read(const reads_t *p, int64_t idx, char *w, int8_t *cls_out)
{
	uint64_t s = mix2(p->seed, (uint64_t)idx);
	double c = u01(&s);
	int64_t L = draw_len(&s, p), pos, k = 0;
	int rev = (int)(splitmix(&s) & 1), cls = 0;
	if (c < p->f_junk) cls = 4;
	else if (c < p->f_junk + p->f_chim) cls = 3;
	else if (c < p->f_junk + p->f_chim + p->f_inv) cls = 2;
	else if (c < p->f_junk + p->f_chim + p->f_inv + p->f_junkins) cls = 1;
	if ((cls == 1 || cls == 2) && L < 4000) cls = 0;
	if (cls_out) *cls_out = (int8_t)cls;
	if (cls == 4) return emit_random(&s, L, w);
	if (cls == 3) { /* two exact fragments of 1-4 kb from different loci, each on a random strand */
		int part;
		for (part = 0; part < 2; ++part) {
			int64_t l2 = 1000 + (int64_t)(u01(&s) * 3000.0), p2, i;
			int rv = (int)(splitmix(&s) & 1);
			draw_locus(&s, p, &l2, &p2);
			for (i = 0; i < l2; ++i) if (w) w[k + i] = rv ? comp(p->gcat[p2 + l2 - 1 - i]) : p->gcat[p2 + i];
			k += l2;
		}
		return k;
	}
	draw_locus(&s, p, &L, &pos);
	if (cls == 0 || L < 4000) return emit_noisy(&s, p, p->gcat + pos, L, rev, w);
	{ /* classes 1 and 2: a block in the middle third of the fragment is replaced / inverted, then errors are applied */
		int64_t bl = (cls == 1 ? 500 : 600) + (int64_t)(u01(&s) * (cls == 1 ? 1000.0 : 900.0));
		int64_t a = L / 3 + (int64_t)(u01(&s) * (double)(L / 3)), b;
		const char *src = p->gcat + pos;
		if (a + bl > L - 1000) bl = L - 1000 - a;
		if (bl < 200) return emit_noisy(&s, p, src, L, rev, w);
		b = a + bl;
		if (!rev) {
			k += emit_noisy(&s, p, src, a, 0, w);
			if (cls == 1) k += emit_random(&s, bl, w ? w + k : 0); else k += emit_noisy(&s, p, src + a, bl, 1, w ? w + k : 0);
			k += emit_noisy(&s, p, src + b, L - b, 0, w ? w + k : 0);
		} else { /* the whole read on the reverse strand */
			k += emit_noisy(&s, p, src + b, L - b, 1, w);
			if (cls == 1) k += emit_random(&s, bl, w ? w + k : 0); else k += emit_noisy(&s, p, src + a, bl, 0, w ? w + k : 0);
			k += emit_noisy(&s, p, src, a, 1, w ? w + k : 0);
		}
		return k;
	}
}
static void reads_block(void *a_, int64_t lo, int64_t hi)
{
	reads_t *p = (reads_t*)a_;
	int64_t i;
	for (i = lo; i < hi; ++i) {
		if (p->pass == 1) p->off[i + 1] = one_read(p, i, 0, p->cls ? &p->cls[i] : 0);
		else one_read(p, p->first + i, p->out + (p->off[p->first + i] - p->off[p->first]), 0);
	}
}
/* pass 1 (out == NULL): fills off[0..n] (prefix sums) and cls for all n_reads; returns total bases.
 * pass 2: writes the bases of reads [first, first + count) into out (the shard of one rank), off as filled by pass 1. */
int64_t mbs_reads(uint64_t seed, const char *gcat, const int64_t *goff, int n_genomes, int64_t n_reads, double n50, double sigma,
                  int64_t min_len, int64_t max_len, double sub, double ins, double del,
                  double f_junkins, double f_inv, double f_chim, double f_junk, int threads, char *out, int64_t *off, int8_t *cls,
                  int64_t first, int64_t count)
{
	reads_t p; int64_t i;
	p.seed = seed, p.gcat = gcat, p.goff = goff, p.n_genomes = n_genomes, p.n_reads = n_reads;
	p.sigma = sigma, p.mu = log(n50) - sigma * sigma; /* length-weighted median = n50 */
	p.min_len = min_len, p.max_len = max_len, p.sub = sub, p.ins = ins, p.del = del;
	p.f_junkins = f_junkins, p.f_inv = f_inv, p.f_chim = f_chim, p.f_junk = f_junk;
	p.out = out, p.off = off, p.cls = cls, p.first = 0;
	if (out == 0) {
		p.pass = 1; off[0] = 0;
		parallel_for(threads, n_reads, 256, reads_block, &p);
		for (i = 0; i < n_reads; ++i) off[i + 1] += off[i];
		return off[n_reads];
	}
	p.pass = 2, p.first = first;
	parallel_for(threads, count, 256, reads_block, &p);
	return off[first + count] - off[first];
}
