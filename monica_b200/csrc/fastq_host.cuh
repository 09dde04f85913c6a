// fastq_host.cuh -- FASTQ ingest and routed writers (host side).
//
// Replaces the per-record Biopython loop of /root/reference/monica/genomes/aligner.py:191,212 (SeqIO.parse) and
// :232,236,243,265 (SeqIO.write): a whole file is parsed in one pass into the concatenated-reads layout mb_map_batch takes,
// and the mapped / unmapped / ambiguous / focus files are appended with Biopython's header rule ('@' + description when
// it starts with the id, else '@' + id + ' ' + description).  No device is needed for these calls.
//
// Built for throughput, because once mapping runs on the GPU this is what a FASTQ file's wall time is made of:
//   * a plain file is mmap'ed (no read copy); gzip input is inflated into one buffer,
//   * the record scan runs on several threads: the file is cut at arbitrary byte positions, every thread finds the first
//     record start at or after its cut ('@' line whose line+2 starts with '+': a quality line that happens to start with
//     '@' fails the test because the line two below it is a sequence line) and parses from there; the pieces must meet
//     exactly, otherwise -- and for anything that is not plain four-line FASTQ -- the sequential parser takes the file,
//   * sequences are gathered into ONE buffer that is page-locked when a CUDA device is present, so mb_map_batch uploads
//     straight from it,
//   * routed records are written with writev straight from the input pages (a record that is plain four-line FASTQ is
//     byte-identical in the output apart from the mapped reads' new id), one thread per sink.
#pragma once
#include <chrono>
#include <mutex>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <sys/uio.h>
#include <unistd.h>
#include <thread>
#include <zlib.h>
#include "common.cuh"
#include "pinpool.cuh"

struct mb_fastq {
	const char *raw = nullptr; size_t raw_len = 0;   // the (decompressed) file
	void *map_base = nullptr; size_t map_len = 0;    // mmap backing, if any
	std::string owned;                               // gzip input / fallback read
	std::vector<int64_t> head, head_len;             // header line without '@'
	std::vector<int32_t> id_len;                     // up to the first blank / tab
	std::vector<int64_t> qual, qual_len;             // quality string (single line, or joined copy in `extra`)
	std::vector<int64_t> rec_end;                    // canonical records: one past the record's final '\n' in raw; else -1
	uint8_t *cat = nullptr; size_t cat_len = 0, cat_cap = 0; bool cat_pinned = false; // concatenated sequences
	std::vector<int64_t> off;                        // [n+1]
	std::string extra;                               // joined quality strings of multi-line records
	std::vector<uint8_t> qual_in_extra;
	~mb_fastq() {
		if (cat) { if (cat_pinned) fqh::pin_give(cat, cat_cap); else free(cat); }
		if (map_base) munmap(map_base, map_len);
	}
};

namespace fqh {

struct Rec { int64_t head, head_len, seq, seq_len, qual, qual_len, end; int32_t id_len; };

static inline int64_t line_end(const char *s, int64_t N, int64_t p)
{
	const void *q = memchr(s + p, '\n', (size_t)(N - p));
	return q ? (int64_t)((const char*)q - s) : N;
}

// strict four-line records with '\n' endings, from `p` while the record starts before `stop`; false on anything else
static bool parse_strict(const char *s, int64_t N, int64_t p, int64_t stop, std::vector<Rec> &out, int64_t *next)
{
	while (p < stop && p < N) {
		if (s[p] != '@') return false;
		const int64_t e1 = line_end(s, N, p); if (e1 >= N) return false;
		const int64_t e2 = line_end(s, N, e1 + 1); if (e2 >= N) return false;
		if (e2 + 2 >= N || s[e2 + 1] != '+' || s[e2 + 2] != '\n') return false;
		const int64_t qb = e2 + 3, e4 = line_end(s, N, qb);
		if (e4 >= N) return false;                                  // the last line must end with '\n' too
		Rec r; r.head = p + 1, r.head_len = e1 - (p + 1), r.seq = e1 + 1, r.seq_len = e2 - (e1 + 1), r.qual = qb, r.qual_len = e4 - qb, r.end = e4 + 1;
		if (r.qual_len != r.seq_len || r.seq_len == 0 || r.head_len <= 0) return false;
		if (s[e1 - 1] == '\r' || s[e2 - 1] == '\r' || s[e4 - 1] == '\r') return false;
		if (s[e1 - 1] == ' ' || s[e1 - 1] == '\t' || s[e1 - 1] == '\v' || s[e1 - 1] == '\f') return false;   // Bio strips the title: general parser
		int32_t idl = 0;
		while (idl < r.head_len && s[r.head + idl] != ' ' && s[r.head + idl] != '\t') ++idl;
		r.id_len = idl;
		out.push_back(r);
		p = r.end;
	}
	*next = p;
	return true;
}

// first record start at or after `p`: a line that starts with '@' and whose line + 2 starts with '+'
static int64_t find_record_start(const char *s, int64_t N, int64_t p)
{
	if (p > 0) { p = line_end(s, N, p - 1) + 1; }                 // start of the next line (or p itself when p begins a line)
	for (int tries = 0; tries < 8 && p < N; ++tries) {
		if (s[p] == '@') {
			const int64_t e1 = line_end(s, N, p);
			if (e1 < N) { const int64_t e2 = line_end(s, N, e1 + 1); if (e2 + 1 < N && s[e2 + 1] == '+') return p; }
		}
		p = line_end(s, N, p) + 1;
	}
	return -1;
}

} // namespace fqh

static void fastq_parse_sequential(mb_fastq *fq, const char *path);

static void fastq_alloc_cat(mb_fastq *fq, size_t n)
{
	fq->cat_len = n;
	if (n == 0) n = 1;
	void *p = nullptr;
	int ndev = 0;
	if (cudaGetDeviceCount(&ndev) == cudaSuccess && ndev > 0 && (p = fqh::pin_take(n + 32, &fq->cat_cap)) != nullptr) { fq->cat = (uint8_t*)p; fq->cat_pinned = true; return; }
	cudaGetLastError();
	fq->cat = (uint8_t*)malloc(n + 32);
	if (!fq->cat) throw mb_error(MB_ERR_NOMEM, "out of host memory for the read buffer");
}

static bool fastq_parse_parallel(mb_fastq *fq)
{
	const char *s = fq->raw; const int64_t N = (int64_t)fq->raw_len;
	if (N == 0) return false;
	int T = (int)std::thread::hardware_concurrency(); if (T < 1) T = 1; if (T > 16) T = 16;
	if (N < ((int64_t)8 << 20)) T = 1;
	std::vector<int64_t> start(T + 1, N);
	start[0] = 0;
	for (int t = 1; t < T; ++t) { const int64_t p = fqh::find_record_start(s, N, N / T * t); start[t] = p < 0 ? N : p; }
	for (int t = 1; t <= T; ++t) if (start[t] < start[t - 1]) start[t] = start[t - 1];
	std::vector<std::vector<fqh::Rec>> recs(T);
	std::vector<int64_t> next(T, 0);
	std::vector<char> ok(T, 0);
	{
		std::vector<std::thread> th;
		for (int t = 0; t < T; ++t) th.emplace_back([&, t]() {
			recs[t].reserve((size_t)((start[t + 1] - start[t]) / 4000 + 16));
			ok[t] = fqh::parse_strict(s, N, start[t], start[t + 1], recs[t], &next[t]) ? 1 : 0;
		});
		for (auto &x : th) x.join();
	}
	for (int t = 0; t < T; ++t) {
		if (!ok[t]) return false;
		if (start[t] < start[t + 1] && next[t] != start[t + 1]) return false;       // the pieces must meet exactly
	}
	size_t n = 0;
	for (int t = 0; t < T; ++t) n += recs[t].size();
	fq->head.resize(n); fq->head_len.resize(n); fq->id_len.resize(n); fq->qual.resize(n); fq->qual_len.resize(n); fq->rec_end.resize(n);
	fq->qual_in_extra.assign(n, 0); fq->off.resize(n + 1);
	std::vector<size_t> base(T + 1, 0); std::vector<int64_t> sbase(T + 1, 0);
	for (int t = 0; t < T; ++t) {
		int64_t sl = 0;
		for (const fqh::Rec &r : recs[t]) sl += r.seq_len;
		base[t + 1] = base[t] + recs[t].size(); sbase[t + 1] = sbase[t] + sl;
	}
	const auto t_a0 = std::chrono::steady_clock::now();
	fastq_alloc_cat(fq, (size_t)sbase[T]);
	if (getenv("MB_DEBUG")) fprintf(stderr, "[mb] fastq read buffer (%s) %.1f ms\n", fq->cat_pinned ? "page-locked" : "pageable", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_a0).count());
	{
		std::vector<std::thread> th;
		for (int t = 0; t < T; ++t) th.emplace_back([&, t]() {
			size_t i = base[t]; int64_t o = sbase[t];
			for (const fqh::Rec &r : recs[t]) {
				fq->head[i] = r.head, fq->head_len[i] = r.head_len, fq->id_len[i] = r.id_len, fq->qual[i] = r.qual, fq->qual_len[i] = r.qual_len, fq->rec_end[i] = r.end;
				fq->off[i] = o;
				memcpy(fq->cat + o, s + r.seq, (size_t)r.seq_len);
				o += r.seq_len; ++i;
			}
		});
		for (auto &x : th) x.join();
	}
	fq->off[n] = sbase[T];
	return true;
}

extern "C" int mb_fastq_load(const char *path, mb_fastq_t **out)
{
	API_BEGIN
	if (!path || !out) throw mb_error(MB_ERR_ARG, "bad arguments");
	std::unique_ptr<mb_fastq> fq(new mb_fastq());
	{
		const int fd = open(path, O_RDONLY);
		if (fd < 0) throw mb_error(MB_ERR_IO, std::string("cannot open ") + path);
		unsigned char magic[2] = {0, 0};
		const ssize_t got = pread(fd, magic, 2, 0);
		struct stat sb;
		if (fstat(fd, &sb) != 0) { close(fd); throw mb_error(MB_ERR_IO, std::string("cannot stat ") + path); }
		if (got == 2 && magic[0] == 0x1f && magic[1] == 0x8b) { // gzip: stream through zlib
			close(fd);
			gzFile fp = gzopen(path, "rb");
			if (!fp) throw mb_error(MB_ERR_IO, std::string("cannot open ") + path);
			gzbuffer(fp, 1 << 20);
			std::vector<char> buf(1 << 22);
			int n;
			while ((n = gzread(fp, buf.data(), (unsigned)buf.size())) > 0) fq->owned.append(buf.data(), (size_t)n);
			const bool bad = n < 0 || !mb_gz_clean_eof(fp);   // a truncated .gz ends with n == 0 and Z_BUF_ERROR
			gzclose(fp);
			if (bad) throw mb_error(MB_ERR_IO, std::string("read error in ") + path);
			fq->raw = fq->owned.data(), fq->raw_len = fq->owned.size();
		} else if (sb.st_size > 0) {
			void *m = mmap(nullptr, (size_t)sb.st_size, PROT_READ, MAP_PRIVATE | MAP_POPULATE, fd, 0);
			if (m != MAP_FAILED) { fq->map_base = m, fq->map_len = (size_t)sb.st_size; fq->raw = (const char*)m, fq->raw_len = (size_t)sb.st_size; }
			else { // not mappable (pipe, odd filesystem): read it
				fq->owned.resize((size_t)sb.st_size);
				size_t done = 0;
				while (done < fq->owned.size()) { const ssize_t r = pread(fd, &fq->owned[done], fq->owned.size() - done, (off_t)done); if (r <= 0) break; done += (size_t)r; }
				if (done != fq->owned.size()) { close(fd); throw mb_error(MB_ERR_IO, std::string("read error in ") + path); }
				fq->raw = fq->owned.data(), fq->raw_len = fq->owned.size();
			}
			close(fd);
		} else close(fd);
	}
	const bool fq_dbg = getenv("MB_DEBUG") != nullptr;
	const auto fq_t0 = std::chrono::steady_clock::now();
	if (!fastq_parse_parallel(fq.get())) {
		fq->head.clear(); fq->head_len.clear(); fq->id_len.clear(); fq->qual.clear(); fq->qual_len.clear(); fq->rec_end.clear(); fq->qual_in_extra.clear(); fq->off.clear();
		if (fq->cat) { if (fq->cat_pinned) fqh::pin_give(fq->cat, fq->cat_cap); else free(fq->cat); fq->cat = nullptr; fq->cat_len = 0; fq->cat_pinned = false; }
		fastq_parse_sequential(fq.get(), path);
	}
	if (fq_dbg) fprintf(stderr, "[mb] fastq parse %.1f ms (%zu bytes)\n", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - fq_t0).count(), fq->raw_len);
	*out = fq.release();
	API_END
}

// the general parser: blank lines, '\r\n', multi-line records, a repeated header on the '+' line, a missing final newline
static void fastq_parse_sequential(mb_fastq *fq, const char *path)
{
	const char *s = fq->raw;
	const int64_t N = (int64_t)fq->raw_len;
	std::vector<uint8_t> cat;
	cat.reserve((size_t)N / 2 + 16);
	auto line_end = [&](int64_t p) { return fqh::line_end(s, N, p); };
	auto rstrip = [&](int64_t b, int64_t e) { while (e > b && (s[e - 1] == '\r' || s[e - 1] == '\n')) --e; return e; };
	fq->off.push_back(0);
	int64_t p = 0;
	while (p < N) {
		int64_t e = line_end(p), le = rstrip(p, e);
		if (le == p) { p = e + 1; continue; }                          // blank line
		if (s[p] != '@') throw mb_error(MB_ERR_IO, std::string("malformed FASTQ: records should start with '@': ") + path);
		const int64_t rec_begin = p;
		const int64_t hb = p + 1;
		int64_t hl = le - hb;
		bool canonical = le == e;                                       // no '\r'
		while (hl > 0 && (s[hb + hl - 1] == ' ' || s[hb + hl - 1] == '\t' || s[hb + hl - 1] == '\v' || s[hb + hl - 1] == '\f' || s[hb + hl - 1] == '\r')) {
			--hl; canonical = false;                                    // Bio.SeqIO right-strips the title line
		}
		int32_t idl = 0;
		while (idl < hl && s[hb + idl] != ' ' && s[hb + idl] != '\t') ++idl;
		p = e + 1;
		// sequence lines until '+'
		int64_t seq_len = 0; int n_seq_lines = 0;
		for (;;) {
			if (p >= N) break;
			e = line_end(p);
			if (s[p] == '+' && n_seq_lines > 0) break;                  // the line after the title is sequence whatever it starts with (Bio.SeqIO)
			le = rstrip(p, e);
			cat.insert(cat.end(), (const uint8_t*)s + p, (const uint8_t*)s + le);
			seq_len += le - p; ++n_seq_lines;
			if (le != e) canonical = false;
			p = e + 1;
		}
		if (n_seq_lines != 1) canonical = false;
		// Bio.SeqIO raises ValueError for a record without its '+' line or with unequal sequence / quality lengths (the
		// reference's SeqIO.parse, aligner.py:191,212): a cut-off file must not be mapped as if it were whole
		if (p >= N) throw mb_error(MB_ERR_IO, std::string("malformed FASTQ: end of file without quality information (record '") + std::string(s + hb, (size_t)std::min<int64_t>(hl, 80)) + "'): " + path);
		e = line_end(p); if (e != p + 1) canonical = false; p = e + 1;   // skip the '+' line
		// quality lines until as long as the sequence
		int64_t qb = p, ql = 0; bool multi = false; size_t xb = fq->extra.size();
		int n_lines = 0; int64_t last_e = p;
		while (ql < seq_len && p < N) {
			e = line_end(p); le = rstrip(p, e);
			if (n_lines == 1) { multi = true; fq->extra.append(s, (size_t)qb, (size_t)ql); }
			if (multi) fq->extra.append(s, (size_t)p, (size_t)(le - p));
			if (le != e) canonical = false;
			ql += le - p; ++n_lines; last_e = e;
			p = e + 1;
		}
		if (ql != seq_len) throw mb_error(MB_ERR_IO, std::string("malformed FASTQ: lengths of sequence and quality values differ for '") + std::string(s + hb, (size_t)std::min<int64_t>(hl, 80)) + "' (" + std::to_string(seq_len) + " and " + std::to_string(ql) + "): " + path);
		if (multi || n_lines != 1 || last_e >= N || (p - rec_begin) != (hl + 1 + 1) + (seq_len + 1) + 2 + (ql + 1)) canonical = false;
		fq->head.push_back(hb); fq->head_len.push_back(hl); fq->id_len.push_back(idl);
		if (multi) { fq->qual.push_back((int64_t)xb); fq->qual_in_extra.push_back(1); }
		else { fq->qual.push_back(qb); fq->qual_in_extra.push_back(0); }
		fq->qual_len.push_back(ql);
		fq->rec_end.push_back(canonical ? p : -1);
		fq->off.push_back((int64_t)cat.size());
	}
	fastq_alloc_cat(fq, cat.size());
	if (!cat.empty()) memcpy(fq->cat, cat.data(), cat.size());
}

extern "C" int64_t mb_fastq_n(const mb_fastq_t *fq) { return fq ? (int64_t)fq->head.size() : 0; }
extern "C" const uint8_t *mb_fastq_seqs(const mb_fastq_t *fq, const int64_t **off)
{
	if (!fq) return nullptr;
	if (off) *off = fq->off.data();
	return fq->cat;
}
extern "C" const char *mb_fastq_header(const mb_fastq_t *fq, int64_t i, int64_t *len, int32_t *id_len)
{
	if (!fq || i < 0 || i >= (int64_t)fq->head.size()) return nullptr;
	if (len) *len = fq->head_len[i];
	if (id_len) *id_len = fq->id_len[i];
	return fq->raw + fq->head[i];
}
/* 1 if every record id of the file is distinct (the vectorised aligner path needs that; duplicates take the reference's
 * per-record dictionary semantics in Python) */
extern "C" int mb_fastq_ids_unique(const mb_fastq_t *fq)
{
	if (!fq) return 0;
	const size_t n = fq->head.size();
	// hash the ids, sort the hashes, compare the ids only where hashes collide
	std::vector<std::pair<uint64_t, uint32_t>> hs(n);
	for (size_t i = 0; i < n; ++i) {
		const unsigned char *p = (const unsigned char*)fq->raw + fq->head[i];
		uint64_t h = 1469598103934665603ULL;
		for (int32_t k = 0; k < fq->id_len[i]; ++k) h = (h ^ p[k]) * 1099511628211ULL;
		hs[i] = std::make_pair(h, (uint32_t)i);
	}
	std::sort(hs.begin(), hs.end());
	for (size_t i = 1; i < n; ++i) {
		if (hs[i].first != hs[i - 1].first) continue;
		for (size_t j = i; j > 0 && hs[j - 1].first == hs[i].first; --j) {
			const uint32_t a = hs[i].second, b = hs[j - 1].second;
			if (fq->id_len[a] == fq->id_len[b] && memcmp(fq->raw + fq->head[a], fq->raw + fq->head[b], (size_t)fq->id_len[a]) == 0) return 0;
		}
	}
	return 1;
}

// ---- routed writers ----
namespace fqh {
struct Sink {
	const char *path = nullptr;
	std::vector<struct iovec> iov;
	std::string small;              // new ids + separators live here (reserved up front: pointers stay valid)
	bool ok = true; std::string err;
	void add(const void *p, size_t n) { if (n) { struct iovec v; v.iov_base = const_cast<void*>(p); v.iov_len = n; iov.push_back(v); } }
	// write iov[lo, hi) at file offset `pos` (pwritev: several threads fill disjoint ranges of one sink)
	bool write_range(int fd, size_t lo, size_t hi, off_t pos) {
		size_t i = lo;
		while (i < hi) {
			const int cnt = (int)std::min<size_t>(hi - i, 1024);
			ssize_t w = pwritev(fd, &iov[i], cnt, pos);
			if (w < 0) return false;
			pos += w;
			while (w > 0 && i < hi) { // advance over fully written entries; finish a partially written one
				if ((size_t)w >= iov[i].iov_len) { w -= (ssize_t)iov[i].iov_len; ++i; }
				else { iov[i].iov_base = (char*)iov[i].iov_base + w; iov[i].iov_len -= (size_t)w; w = 0; }
			}
		}
		return true;
	}
	void flush_to_file() {
		// appended like the reference does (open mode 'a'): the records go behind whatever the file holds already
		const int fd = open(path, O_WRONLY | O_CREAT, 0666);
		if (fd < 0) { ok = false; err = std::string("cannot append to ") + path; return; }
		struct stat sb;
		if (fstat(fd, &sb) != 0) { ok = false; err = std::string("cannot stat ") + path; close(fd); return; }
		size_t total = 0;
		for (const struct iovec &v : iov) total += v.iov_len;
		int T = total > ((size_t)32 << 20) ? 4 : 1;
		if (T > 1 && ftruncate(fd, sb.st_size + (off_t)total) != 0) T = 1;
		if (T == 1) ok = write_range(fd, 0, iov.size(), sb.st_size);
		else {
			std::vector<size_t> cut(T + 1, iov.size()); std::vector<off_t> pos(T + 1, 0);
			cut[0] = 0; pos[0] = sb.st_size;
			size_t acc = 0; int k = 1;
			for (size_t i = 0; i < iov.size() && k < T; ++i) {
				if (acc >= total / T * k) { cut[k] = i; pos[k] = sb.st_size + (off_t)acc; ++k; }
				acc += iov[i].iov_len;
			}
			for (; k <= T; ++k) { cut[k] = iov.size(); pos[k] = sb.st_size + (off_t)total; }
			std::vector<char> good(T, 1);
			std::vector<std::thread> th;
			for (int t = 0; t < T; ++t) th.emplace_back([&, t]() { good[t] = write_range(fd, cut[t], cut[t + 1], pos[t]) ? 1 : 0; });
			for (auto &x : th) x.join();
			for (int t = 0; t < T; ++t) if (!good[t]) ok = false;
		}
		if (!ok && err.empty()) err = std::string("write failed: ") + path;
		if (close(fd) != 0 && ok) { ok = false; err = std::string("write failed: ") + path; }
	}
};
}

/* dest[i]: 0 unmapped, 1 mapped, 2 ambiguous, anything else: skip.  Mapped reads take the id ids[target[i]] in place of the
 * record id (aligner.py:242); focus[i] != 0: also append the ORIGINAL record to focus_path (aligner.py:235-236).  Files are
 * opened in append mode like the reference does; a NULL path skips that sink. */
static void fastq_route_impl(const mb_fastq *fq, const int8_t *dest, const int32_t *target, const char *const *ids, int32_t n_ids, const char *const *per_read_id,
                             const uint8_t *focus, const char *mapped_path, const char *unmapped_path, const char *ambiguous_path, const char *focus_path)
{
	static const char AT = '@', NL = '\n', SP = ' ';
	static const char PLUS[3] = { '\n', '+', '\n' };
	fqh::Sink sink[4];
	const char *paths[4] = { unmapped_path, mapped_path, ambiguous_path, focus_path };
	const int64_t n = (int64_t)fq->head.size();
	std::vector<size_t> id_len_of(n_ids > 0 ? n_ids : 0);
	for (int32_t k = 0; k < n_ids; ++k) id_len_of[k] = ids[k] ? strlen(ids[k]) : 0;
	size_t cnt[4] = {0, 0, 0, 0};
	for (int64_t i = 0; i < n; ++i) {
		const int d = dest[i];
		if (d < 0 || d > 2) continue;
		++cnt[d];
		if (focus && focus[i] && focus_path) ++cnt[3];
	}
	for (int k = 0; k < 4; ++k) { sink[k].path = paths[k]; if (paths[k]) sink[k].iov.reserve(cnt[k] * (k == 1 ? 4 : 1) + 16); }
	auto emit = [&](fqh::Sink &o, int64_t i, const char *rid, size_t rl) {
		const char *h = fq->raw + fq->head[i];
		const bool same_id = rid && (size_t)fq->id_len[i] == rl && memcmp(h, rid, rl) == 0;  // Bio's writer: the description already starts with the id
		if (fq->rec_end[i] >= 0) {                 // plain four-line record: everything from the header on is already in output form
			if (!rid || same_id) { o.add(h - 1, (size_t)(fq->rec_end[i] - (fq->head[i] - 1))); return; }
			o.add(&AT, 1); o.add(rid, rl); o.add(&SP, 1);
			o.add(h, (size_t)(fq->rec_end[i] - fq->head[i]));
			return;
		}
		o.add(&AT, 1);
		if (rid && !same_id) { o.add(rid, rl); o.add(&SP, 1); }
		o.add(h, (size_t)fq->head_len[i]);
		o.add(&NL, 1);
		o.add(fq->cat + fq->off[i], (size_t)(fq->off[i + 1] - fq->off[i]));
		o.add(PLUS, 3);
		const char *q = fq->qual_in_extra[i] ? fq->extra.data() + fq->qual[i] : fq->raw + fq->qual[i];
		o.add(q, (size_t)fq->qual_len[i]);
		o.add(&NL, 1);
	};
	for (int64_t i = 0; i < n; ++i) {
		const int d = dest[i];
		if (d < 0 || d > 2) continue;
		if (focus && focus[i] && focus_path) emit(sink[3], i, nullptr, 0);
		if (!paths[d]) continue;
		if (d == 1) {
			const char *rid = nullptr; size_t rl = 0;
			if (per_read_id) { rid = per_read_id[i]; rl = rid ? strlen(rid) : 0; }
			else if (target && target[i] >= 0 && target[i] < n_ids) { rid = ids[target[i]]; rl = id_len_of[target[i]]; }
			if (!rid) throw mb_error(MB_ERR_ARG, "mapped read without a new id");
			emit(sink[1], i, rid, rl);
		} else emit(sink[d], i, nullptr, 0);
	}
	{
		std::vector<std::thread> th;
		for (int k = 0; k < 4; ++k) if (paths[k]) th.emplace_back([&, k]() { sink[k].flush_to_file(); });
		for (auto &t : th) t.join();
	}
	for (int k = 0; k < 4; ++k) if (paths[k] && !sink[k].ok) throw mb_error(MB_ERR_IO, sink[k].err);
}

extern "C" int mb_fastq_route(const mb_fastq_t *fq, const int8_t *dest, const char *const *new_id, const uint8_t *focus,
                              const char *mapped_path, const char *unmapped_path, const char *ambiguous_path, const char *focus_path)
{
	API_BEGIN
	if (!fq || !dest) throw mb_error(MB_ERR_ARG, "bad arguments");
	fastq_route_impl(fq, dest, nullptr, nullptr, 0, new_id, focus, mapped_path, unmapped_path, ambiguous_path, focus_path);
	API_END
}

extern "C" int mb_fastq_route_targets(const mb_fastq_t *fq, const int8_t *dest, const int32_t *target, const char *const *ids, int32_t n_ids, const uint8_t *focus,
                                      const char *mapped_path, const char *unmapped_path, const char *ambiguous_path, const char *focus_path)
{
	API_BEGIN
	if (!fq || !dest || (n_ids > 0 && (!ids || !target))) throw mb_error(MB_ERR_ARG, "bad arguments");
	fastq_route_impl(fq, dest, target, ids, n_ids, nullptr, focus, mapped_path, unmapped_path, ambiguous_path, focus_path);
	API_END
}

extern "C" void mb_fastq_free(mb_fastq_t *fq) { delete fq; }
