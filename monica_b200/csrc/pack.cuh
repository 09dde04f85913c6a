// pack.cuh -- reads as 2-bit words + ambiguity intervals: packed once on the host, uploaded at a quarter of the bytes.
//
// A read batch reaches mb_map_batch as concatenated ASCII (1 B/base over PCIe).  mappy's own input is a Python str per
// read (/root/reference/monica/genomes/aligner.py:193,215: index.map(str(seq_record.seq))); nothing downstream of the nt4
// table (A/a 0, C/c 1, G/g 2, T/t/U/u 3, everything else 4) ever looks at the characters again.  mb_reads_pack / mb_fastq_pack
// therefore reduce the batch ONCE, on the host and on several threads, to
//   * words[]: base i of the concatenation in bits [2(i&15), 2(i&15)+2) of word i>>4 (ambiguous bases as 0),
//   * intervals[]: (start, length) of every maximal run of ambiguous bases, ascending,
// in page-locked memory, and mb_map_packed uploads that (0.25 B/base + 16 B per run of N) and expands it on the device to the
// nt4 byte codes the kernels read -- the same bytes k_encode_nt4 produces from the ASCII, so every later stage is unchanged.
#pragma once
#include <emmintrin.h>
#include <thread>
#include <vector>
#include <memory>
#include "common.cuh"
#include "pinpool.cuh"

struct mb_packed {
	int32_t n_reads = 0;
	int64_t total = 0;
	std::vector<int64_t> off;        // [n_reads+1]
	uint32_t *words = nullptr;       // [n_words]: ceil(total/16) + 2 (zero padding: the device reads one word ahead)
	size_t n_words = 0, words_cap = 0;
	bool pinned = false;
	std::vector<int64_t> iv;         // 2 per interval: start, length
	~mb_packed();
};

namespace pk {

// 16 ASCII bases -> one word of 2-bit codes; returns the mask of positions that are not A/C/G/T/U in either case
static inline uint32_t pack16(const uint8_t *s, uint32_t *bad)
{
	const __m128i v = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s));
	const __m128i u = _mm_and_si128(v, _mm_set1_epi8((char)0xDF));   // fold case (only 0x61..0x7a reach 0x41..0x5a this way)
	const __m128i isA = _mm_cmpeq_epi8(u, _mm_set1_epi8('A')), isC = _mm_cmpeq_epi8(u, _mm_set1_epi8('C')), isG = _mm_cmpeq_epi8(u, _mm_set1_epi8('G'));
	const __m128i isT = _mm_or_si128(_mm_cmpeq_epi8(u, _mm_set1_epi8('T')), _mm_cmpeq_epi8(u, _mm_set1_epi8('U')));
	*bad = ~(uint32_t)_mm_movemask_epi8(_mm_or_si128(_mm_or_si128(isA, isC), _mm_or_si128(isG, isT))) & 0xffffu;
	__m128i c = _mm_or_si128(_mm_and_si128(isC, _mm_set1_epi8(1)), _mm_or_si128(_mm_and_si128(isG, _mm_set1_epi8(2)), _mm_and_si128(isT, _mm_set1_epi8(3))));
	// bytes b0..b15 (2 bits each) -> 32 bits: fold pairs of bytes, then pairs of 16-bit, then pairs of 32-bit lanes
	c = _mm_and_si128(_mm_or_si128(c, _mm_srli_epi16(c, 6)), _mm_set1_epi16(0x000F));
	c = _mm_and_si128(_mm_or_si128(c, _mm_srli_epi32(c, 12)), _mm_set1_epi32(0x000000FF));
	c = _mm_or_si128(c, _mm_srli_epi64(c, 24));
	return ((uint32_t)_mm_cvtsi128_si32(c) & 0xffffu) | ((uint32_t)_mm_extract_epi16(c, 4) << 16);
}

// words [w0, w1) of the concatenation `cat` (total bases); the runs of ambiguous bases found are appended to iv
static void pack_range(const uint8_t *cat, int64_t total, int64_t w0, int64_t w1, uint32_t *words, std::vector<int64_t> &iv)
{
	for (int64_t w = w0; w < w1; ++w) {
		const int64_t i = w << 4;
		uint32_t bad, word;
		if (i + 16 <= total) word = pack16(cat + i, &bad);
		else { // the last, partial word of the batch
			uint8_t tmp[16];
			const int n = (int)(total - i);
			for (int j = 0; j < 16; ++j) tmp[j] = j < n ? cat[i + j] : (uint8_t)'A';
			word = pack16(tmp, &bad);
		}
		words[w] = word;
		while (bad) {
			const int j = __builtin_ctz(bad);
			int len = __builtin_ctz(~(bad >> j));          // run of set bits from j (bad has 16 significant bits: ~ ends it)
			const int64_t s = i + j;
			if (!iv.empty() && iv[iv.size() - 2] + iv.back() == s) iv.back() += len;
			else { iv.push_back(s); iv.push_back(len); }
			bad = j + len >= 32 ? 0u : bad & ~(((1u << len) - 1u) << j);
		}
	}
}

static mb_packed *pack(const uint8_t *cat, const int64_t *off, int32_t n_reads, int n_threads)
{
	if (n_reads < 0 || !off || (n_reads > 0 && off[0] != 0)) throw mb_error(MB_ERR_ARG, "offsets must start at 0");
	for (int32_t i = 0; i < n_reads; ++i) if (off[i + 1] < off[i]) throw mb_error(MB_ERR_ARG, "offsets must be non-decreasing");
	std::unique_ptr<mb_packed> P(new mb_packed());
	P->n_reads = n_reads;
	P->total = n_reads > 0 ? off[n_reads] : 0;
	if (P->total > 0 && !cat) throw mb_error(MB_ERR_ARG, "bad arguments");
	P->off.assign(off, off + n_reads + 1);
	if (n_reads == 0) P->off.assign(1, 0);
	const int64_t nw = (P->total + 15) >> 4;
	P->n_words = (size_t)nw + 2;
	int ndev = 0;
	void *p = nullptr;
	if (cudaGetDeviceCount(&ndev) == cudaSuccess && ndev > 0 && (p = fqh::pin_take(P->n_words * 4, &P->words_cap)) != nullptr) P->pinned = true;
	else {
		cudaGetLastError();
		p = malloc(P->n_words * 4);
		if (!p) throw mb_error(MB_ERR_NOMEM, "out of host memory for the packed reads");
	}
	P->words = (uint32_t*)p;
	P->words[nw] = P->words[nw + 1] = 0;
	int T = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
	if (T < 1) T = 1;
	if (T > 64) T = 64;
	if ((int64_t)T > (nw >> 16) + 1) T = (int)((nw >> 16) + 1);   // at least 1 Mbase per thread
	std::vector<std::vector<int64_t>> ivs(T);
	if (T == 1) pack_range(cat, P->total, 0, nw, P->words, ivs[0]);
	else {
		std::vector<std::thread> th;
		for (int t = 0; t < T; ++t) {
			const int64_t w0 = nw * t / T, w1 = nw * (t + 1) / T;
			th.emplace_back([&, t, w0, w1]() { pack_range(cat, P->total, w0, w1, P->words, ivs[t]); });
		}
		for (auto &t : th) t.join();
	}
	for (int t = 0; t < T; ++t) { // in order; a run that crosses a thread boundary is joined
		const std::vector<int64_t> &v = ivs[t];
		for (size_t k = 0; k < v.size(); k += 2) {
			if (!P->iv.empty() && P->iv[P->iv.size() - 2] + P->iv.back() == v[k]) P->iv.back() += v[k + 1];
			else { P->iv.push_back(v[k]); P->iv.push_back(v[k + 1]); }
		}
	}
	return P.release();
}

} // namespace pk

inline mb_packed::~mb_packed()
{
	if (words) { if (pinned) fqh::pin_give(words, words_cap); else free(words); }
}

// ---- device side ----
// W[t], W[t+1] hold the 16 bases of output group t and the first `sh/2` of the next word's (the piece starts sh/2 bases into
// its first word); 16 codes per thread, one 16-byte store
__global__ void k_unpack_nt4(const uint32_t *__restrict__ W, int sh, uint8_t *__restrict__ codes, int64_t n)
{
	const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	const int64_t i = t * 16;
	if (i >= n) return;
	const uint32_t w = __funnelshift_r(W[t], W[t + 1], sh);
	uint32_t o[4];
	#pragma unroll
	for (int j = 0; j < 4; ++j) {
		const uint32_t b = w >> (8 * j) & 0xffu;
		o[j] = (b & 3u) | (b >> 2 & 3u) << 8 | (b >> 4 & 3u) << 16 | (b >> 6) << 24;
	}
	if (i + 16 <= n) *reinterpret_cast<uint4*>(codes + i) = make_uint4(o[0], o[1], o[2], o[3]);
	else for (int j = 0; i + j < n; ++j) codes[i + j] = (uint8_t)(o[j >> 2] >> (8 * (j & 3)));
}

// runs of ambiguous bases (global coordinates) -> code 4 in codes[lo, hi) (coordinates relative to g0); one warp per run
__global__ void k_apply_amb(const int64_t *__restrict__ iv, int64_t n_iv, int64_t g0, uint8_t *__restrict__ codes, int64_t lo, int64_t hi)
{
	const int64_t k = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	if (k >= n_iv) return;
	int64_t s = iv[2 * k] - g0, e = s + iv[2 * k + 1];
	if (s < lo) s = lo;
	if (e > hi) e = hi;
	for (int64_t x = s + (threadIdx.x & 31); x < e; x += 32) codes[x] = 4;
}

// what a piece of a batch is uploaded from: ASCII (mb_map_batch) or packed words (mb_map_packed)
struct HostSrc {
	const uint8_t *ascii = nullptr;   // concatenated ASCII of the whole batch
	const mb_packed *pk = nullptr;
	bool any() const { return ascii || pk; }
};
