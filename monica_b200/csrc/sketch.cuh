// sketch.cuh -- K1: (w,k)-minimizer sketching of a concatenated read batch on sm_100a.
//
// Replaces minimap2-2.17 sketch.c mm_sketch() as reached from mappy's Aligner.map
// (/root/reference/monica/genomes/aligner.py:193,215 -> mm_map_frag -> collect_minimizers).
//
// Parallelisation: the batch is one long nt4-coded byte array cut into fixed 256-base chunks, one
// thread per chunk.  mm_sketch is a sequential automaton, but its state after a position is a pure
// function of the last w window entries plus whether l (valid bases since the last N) has
// saturated, so a thread that warms up on the 64 bases before its chunk reproduces the sequential
// state exactly (odd k only: a full odd-length k-mer can never equal its reverse complement, so the
// "symmetric k-mer" stall cannot fire during warm-up).  If the warm-up window holds an N, k is even,
// or w+k > 48, the thread replays the read from its first base instead.  Emission is two-pass
// (count -> exclusive scan -> write) so minimizers land in read-major, position order.
//
// Memory: a CTA stages its 128 chunks (+64 warm-up bytes) into shared memory with coalesced 16-byte
// loads; chunk rows are padded by one word so the per-thread strided reads are bank-conflict free.
#pragma once
#include "common.cuh"

#define SK_CHUNK 256
#define SK_TPB   128
#define SK_WARM  64
#define SK_SPAN  (SK_CHUNK * SK_TPB)
#define SK_SMEM_BYTES ((SK_SPAN + SK_WARM) + 4 * ((SK_SPAN + SK_WARM) / 256 + 2))

MB_HD uint64_t mb_hash64(uint64_t key, uint64_t mask)
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

// ASCII -> nt4 (A/C/G/T -> 0..3 case-insensitive, U -> 3, everything else 4), 16 bytes per thread-iteration
__device__ __constant__ uint8_t c_nt4[256];

static const uint8_t h_nt4[256] = {
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

__global__ void k_encode_nt4(const uint8_t *__restrict__ ascii, uint8_t *__restrict__ codes, int64_t n)
{
	int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 16;
	if (i + 16 <= n) {
		uint4 v = *reinterpret_cast<const uint4*>(ascii + i);
		uint32_t w[4] = {v.x, v.y, v.z, v.w}, o[4];
		#pragma unroll
		for (int j = 0; j < 4; ++j)
			o[j] = (uint32_t)c_nt4[w[j] & 0xff] | (uint32_t)c_nt4[w[j] >> 8 & 0xff] << 8 | (uint32_t)c_nt4[w[j] >> 16 & 0xff] << 16 | (uint32_t)c_nt4[w[j] >> 24] << 24;
		*reinterpret_cast<uint4*>(codes + i) = make_uint4(o[0], o[1], o[2], o[3]);
	} else {
		for (; i < n; ++i) codes[i] = c_nt4[ascii[i]];
	}
}

// ---- the mm_sketch automaton with a shift-register window (newest entry at index W-1) ----
template <int W>
struct SketchState {
	uint64_t kmer0, kmer1, minx;
	uint64_t wx[W];
	uint32_t wy[W];
	uint32_t miny;
	int l, min_age;
	MB_D void fresh() {
		kmer0 = kmer1 = 0; l = 0; minx = UINT64_MAX; miny = 0xffffffffu; min_age = W - 1;
		#pragma unroll
		for (int j = 0; j < W; ++j) wx[j] = UINT64_MAX, wy[j] = 0xffffffffu;
	}
	// one base; `i` is the position within the read; emit(x, ylow) called in output order
	template <typename E>
	MB_D void step(int c, uint32_t i, int w, int k, uint64_t mask, uint64_t shift1, E &emit) {
		uint64_t ix = UINT64_MAX; uint32_t iy = 0xffffffffu;
		if (c < 4) {
			int kmer_span = l + 1 < k ? l + 1 : k;
			kmer0 = (kmer0 << 2 | (uint64_t)c) & mask;
			kmer1 = (kmer1 >> 2) | (3ULL ^ (uint64_t)c) << shift1;
			if (kmer0 == kmer1) return; // symmetric k-mer: no window advance (upstream `continue`)
			int z = kmer0 < kmer1 ? 0 : 1;
			++l;
			if (l >= k) {
				ix = mb_hash64(z ? kmer1 : kmer0, mask) << 8 | (uint64_t)kmer_span;
				iy = i << 1 | (uint32_t)z;
			}
		} else l = 0;
		#pragma unroll
		for (int j = 0; j < W - 1; ++j) wx[j] = wx[j + 1], wy[j] = wy[j + 1];
		wx[W - 1] = ix, wy[W - 1] = iy;
		--min_age;
		if (l == w + k - 1 && minx != UINT64_MAX) { // first window: identical k-mers not stored yet
			#pragma unroll
			for (int j = 0; j < W - 1; ++j)
				if (minx == wx[j] && wy[j] != miny) emit(wx[j], wy[j]);
		}
		if (ix <= minx) { // a new minimum; write the old min
			if (l >= w + k && minx != UINT64_MAX) emit(minx, miny);
			minx = ix, miny = iy, min_age = W - 1;
		} else if (min_age < 0) { // old min has moved outside the window
			if (l >= w + k - 1 && minx != UINT64_MAX) emit(minx, miny);
			minx = UINT64_MAX;
			#pragma unroll
			for (int j = 0; j < W; ++j)
				if (minx >= wx[j]) minx = wx[j], miny = wy[j], min_age = j;
			if (l >= w + k - 1 && minx != UINT64_MAX) {
				#pragma unroll
				for (int j = 0; j < W; ++j)
					if (minx == wx[j] && miny != wy[j]) emit(wx[j], wy[j]);
			}
		}
	}
};

struct EmitCount { int n; MB_D void operator()(uint64_t, uint32_t) { ++n; } };
// count pass that also parks the minimizers of its chunk in a fixed-capacity staging row, so that the second run of the
// automaton (the write pass) is only needed when some chunk overflowed its row
#define SK_STAGE_CAP 80
struct EmitStage {
	mb128 *row; int n; int base; uint64_t rid_hi;
	MB_D void operator()(uint64_t x, uint32_t y) { const int k = base + n; if (k < SK_STAGE_CAP) { row[k].x = x; row[k].y = rid_hi | y; } ++n; }
};
struct EmitWrite {
	mb128 *out; int64_t pos; uint64_t rid_hi;
	MB_D void operator()(uint64_t x, uint32_t y) { out[pos].x = x; out[pos].y = rid_hi | y; ++pos; }
};

// read index containing global position g (skipping empty reads): largest r with off[r] <= g
MB_D int sk_find_read(const int64_t *__restrict__ off, int n_reads, int64_t g)
{
	int lo = 0, hi = n_reads; // invariant: off[lo] <= g < off[hi]
	while (hi - lo > 1) {
		int mid = (lo + hi) >> 1;
		if (off[mid] <= g) lo = mid; else hi = mid;
	}
	return lo;
}

template <int W, bool WRITE>
__global__ void __launch_bounds__(SK_TPB)
k_sketch(const uint8_t *__restrict__ codes, const int64_t *__restrict__ off, int n_reads, int64_t total,
         int w, int k, int32_t *__restrict__ chunk_cnt, int32_t *__restrict__ read_cnt,
         const int64_t *__restrict__ chunk_off, mb128 *__restrict__ out, mb128 *__restrict__ stage, int *__restrict__ overflow, int64_t cta0)
{
	extern __shared__ __align__(16) uint8_t sm[];
	const int64_t cta_base = ((int64_t)blockIdx.x + cta0) * SK_SPAN;
	const int64_t win_lo = cta_base - SK_WARM; // logical smem index 0 <-> global win_lo (may be negative)
	// ---- stage [win_lo, cta_base + SK_SPAN) with coalesced 16-byte loads; row padding of one word per 256 bytes ----
	{
		const int n_vec = (SK_SPAN + SK_WARM) / 16;
		for (int v = threadIdx.x; v < n_vec; v += SK_TPB) {
			int64_t g = win_lo + (int64_t)v * 16;
			uint4 val = make_uint4(0x04040404u, 0x04040404u, 0x04040404u, 0x04040404u);
			if (g >= 0 && g + 16 <= total) val = *reinterpret_cast<const uint4*>(codes + g);
			else if (g + 16 > 0 && g < total) {
				uint8_t tmp[16];
				#pragma unroll
				for (int b = 0; b < 16; ++b) tmp[b] = (g + b >= 0 && g + b < total) ? codes[g + b] : 4;
				val = *reinterpret_cast<uint4*>(tmp);
			}
			int L = v * 16;
			int P = L + 4 * (L >> 8);
			uint32_t *d = reinterpret_cast<uint32_t*>(sm + P); // a 16-byte group never straddles a 256-byte row
			d[0] = val.x, d[1] = val.y, d[2] = val.z, d[3] = val.w;
		}
	}
	__syncthreads();
	const int64_t s = cta_base + (int64_t)threadIdx.x * SK_CHUNK;
	const int64_t chunk = ((int64_t)blockIdx.x + cta0) * SK_TPB + threadIdx.x;
	if (s >= total) { if (!WRITE && chunk_cnt && chunk * SK_CHUNK < total + SK_CHUNK) {} return; }
	const int64_t e = s + SK_CHUNK < total ? s + SK_CHUNK : total;
	const uint64_t mask = (1ULL << 2 * k) - 1, shift1 = 2 * (k - 1);
	auto code_at = [&](int64_t g) -> int {
		int L = (int)(g - win_lo);
		if (g >= win_lo && L < SK_SPAN + SK_WARM) return sm[L + 4 * (L >> 8)];
		return codes[g];
	};
	SketchState<W> st;
	int r = sk_find_read(off, n_reads, s);
	int64_t pos = s;
	if (pos < off[0]) pos = off[0]; // a sub-batch may start a few bytes into its (aligned) code array
	int total_cnt = 0;
	int64_t wpos = WRITE ? chunk_off[chunk] : 0;
	while (pos < e) {
		const int64_t rs = off[r], re = off[r + 1];
		const int64_t seg_end = re < e ? re : e;
		EmitCount nul; nul.n = 0;
		// ---- bring the automaton to `pos` ----
		bool warm = false;
		if (pos > rs && (k & 1) && w + k <= 48 && pos - rs > SK_WARM) {
			warm = true;
			for (int64_t g = pos - SK_WARM; g < pos; ++g) if (code_at(g) >= 4) { warm = false; break; }
		}
		st.fresh();
		if (warm) {
			int64_t g = pos - SK_WARM;
			for (int j = 0; j < k - 1; ++j, ++g) { // pre-fill the k-mer registers
				uint64_t c = (uint64_t)code_at(g);
				st.kmer0 = (st.kmer0 << 2 | c) & mask;
				st.kmer1 = (st.kmer1 >> 2) | (3ULL ^ c) << shift1;
			}
			st.l = k - 1;
			for (; g < pos; ++g) st.step(code_at(g), (uint32_t)(g - rs), w, k, mask, shift1, nul);
		} else {
			for (int64_t g = rs; g < pos; ++g) st.step(code_at(g), (uint32_t)(g - rs), w, k, mask, shift1, nul);
		}
		// ---- emit for [pos, seg_end) ----
		int seg_cnt;
		if (WRITE) {
			EmitWrite ew; ew.out = out; ew.pos = wpos; ew.rid_hi = (uint64_t)(uint32_t)r << 32;
			for (int64_t g = pos; g < seg_end; ++g) st.step(code_at(g), (uint32_t)(g - rs), w, k, mask, shift1, ew);
			if (seg_end == re && st.minx != UINT64_MAX) ew(st.minx, st.miny);
			seg_cnt = (int)(ew.pos - wpos);
			wpos = ew.pos;
		} else if (stage) {
			EmitStage es; es.row = stage + chunk * SK_STAGE_CAP; es.n = 0; es.base = total_cnt; es.rid_hi = (uint64_t)(uint32_t)r << 32;
			for (int64_t g = pos; g < seg_end; ++g) st.step(code_at(g), (uint32_t)(g - rs), w, k, mask, shift1, es);
			if (seg_end == re && st.minx != UINT64_MAX) es(st.minx, st.miny);
			seg_cnt = es.n;
			if (seg_cnt) atomicAdd(&read_cnt[r], seg_cnt);
		} else {
			EmitCount ec; ec.n = 0;
			for (int64_t g = pos; g < seg_end; ++g) st.step(code_at(g), (uint32_t)(g - rs), w, k, mask, shift1, ec);
			if (seg_end == re && st.minx != UINT64_MAX) ++ec.n;
			seg_cnt = ec.n;
			if (seg_cnt) atomicAdd(&read_cnt[r], seg_cnt);
		}
		total_cnt += seg_cnt;
		pos = seg_end;
		if (pos == re) { ++r; while (r < n_reads && off[r + 1] == off[r]) ++r; }
	}
	if (!WRITE) {
		chunk_cnt[chunk] = total_cnt;
		if (stage && total_cnt > SK_STAGE_CAP) *overflow = 1;
	}
}

// staged minimizers -> their final, read-major positions: 8 lanes per chunk, 16-byte copies
__global__ void k_sketch_compact(const mb128 *__restrict__ stage, const int32_t *__restrict__ chunk_cnt, const int64_t *__restrict__ chunk_off,
                                 int64_t n_chunks, mb128 *__restrict__ out)
{
	const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	const int64_t chunk = t >> 3; const int sub = (int)(t & 7);
	if (chunk >= n_chunks) return;
	const int n = chunk_cnt[chunk] < SK_STAGE_CAP ? chunk_cnt[chunk] : SK_STAGE_CAP;
	const uint4 *src = reinterpret_cast<const uint4*>(stage + chunk * SK_STAGE_CAP);
	uint4 *dst = reinterpret_cast<uint4*>(out + chunk_off[chunk]);
	for (int i = sub; i < n; i += 8) dst[i] = src[i];
}

struct SketchOut {
	mb128 *mini = nullptr;       // [n_mini] x = hash<<8|span, y = read<<32 | pos<<1 | strand
	int64_t *mini_off = nullptr; // [n_reads+1]
	int64_t n_mini = 0;
};

// Optional input feed for run_sketch: the ASCII reads are still in (pinned) host memory; they are copied in pieces on a
// second stream and each piece is encoded and sketched as soon as it has landed, so the host->device copy of a batch hides
// behind the sketch kernel instead of preceding it.
struct SketchFeed {
	const uint8_t *h_ascii = nullptr;   // host reads (concatenated ASCII)
	uint8_t *d_ascii = nullptr;         // device staging, total + 32 bytes
	uint8_t *d_codes = nullptr;         // nt4 codes to produce (== `codes` passed to run_sketch)
	cudaStream_t copy_st = nullptr;
	std::vector<cudaEvent_t> *events = nullptr; // pool, grown on demand
};

// codes: device nt4 bytes [total]; d_off: device offsets [n_reads+1].  Synchronises once to learn n_mini.
static void run_sketch(Arena &ar, cudaStream_t st, const uint8_t *codes, const int64_t *d_off, int n_reads, int64_t total,
                       int w, int k, SketchOut &o, int64_t *n_launch, const SketchFeed *feed = nullptr)
{
	if (w != 10) throw mb_error(MB_ERR_ARG, "sketch kernel is instantiated for w=10 (map-ont) only");
	if (k < 1 || k > 28) throw mb_error(MB_ERR_ARG, "k out of range");
	int64_t n_chunks = cdiv(total, SK_CHUNK);
	int64_t n_cta = cdiv(total, SK_SPAN);
	o.mini_off = ar.get<int64_t>(n_reads + 1);
	if (total == 0 || n_reads == 0) {
		CK(cudaMemsetAsync(o.mini_off, 0, (n_reads + 1) * sizeof(int64_t), st));
		o.mini = ar.get<mb128>(1); o.n_mini = 0;
		return;
	}
	int32_t *chunk_cnt = ar.get<int32_t>(n_chunks);
	int32_t *read_cnt = ar.get<int32_t>(n_reads);
	int64_t *chunk_off = ar.get<int64_t>(n_chunks + 1);
	CK(cudaMemsetAsync(read_cnt, 0, n_reads * sizeof(int32_t), st));
	static bool attr_set = false;
	if (!attr_set) {
		CK(cudaFuncSetAttribute(k_sketch<10, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SK_SMEM_BYTES));
		CK(cudaFuncSetAttribute(k_sketch<10, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SK_SMEM_BYTES));
		attr_set = true;
	}
	mb128 *stage = ar.get<mb128>((size_t)n_chunks * SK_STAGE_CAP);
	int *d_ovf = ar.get<int>(1);
	CK(cudaMemsetAsync(d_ovf, 0, sizeof(int), st));
	if (!feed) {
		k_sketch<10, false><<<(unsigned)n_cta, SK_TPB, SK_SMEM_BYTES, st>>>(codes, d_off, n_reads, total, w, k, chunk_cnt, read_cnt, nullptr, nullptr, stage, d_ovf, 0);
		++*n_launch;
	} else {
		// pieces of ~1/8 of the batch, whole CTA spans each; piece p is sketched while piece p+1 is in flight
		int64_t per = cdiv(cdiv(n_cta, 8), 1); if (per < 1) per = 1;
		int p = 0;
		for (int64_t c0 = 0; c0 < n_cta; c0 += per, ++p) {
			const int64_t c1 = c0 + per < n_cta ? c0 + per : n_cta;
			const int64_t b0 = c0 * SK_SPAN, b1 = c1 * SK_SPAN < total ? c1 * SK_SPAN : total;
			while ((int)feed->events->size() <= p) { cudaEvent_t e; CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); feed->events->push_back(e); }
			CK(cudaMemcpyAsync(feed->d_ascii + b0, feed->h_ascii + b0, (size_t)(b1 - b0), cudaMemcpyHostToDevice, feed->copy_st));
			CK(cudaEventRecord((*feed->events)[p], feed->copy_st));
			CK(cudaStreamWaitEvent(st, (*feed->events)[p], 0));
			k_encode_nt4<<<(unsigned)cdiv(cdiv(b1 - b0, 16), 256), 256, 0, st>>>(feed->d_ascii + b0, feed->d_codes + b0, b1 - b0);
			k_sketch<10, false><<<(unsigned)(c1 - c0), SK_TPB, SK_SMEM_BYTES, st>>>(codes, d_off, n_reads, total, w, k, chunk_cnt, read_cnt, nullptr, nullptr, stage, d_ovf, c0);
			*n_launch += 2;
		}
	}
	exclusive_scan<int32_t>(ar, st, chunk_cnt, chunk_off, n_chunks, n_launch);
	exclusive_scan<int32_t>(ar, st, read_cnt, o.mini_off, n_reads, n_launch);
	int64_t n_mini = 0; int h_ovf = 0;
	CK(cudaMemcpyAsync(&n_mini, chunk_off + n_chunks, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
	CK(cudaMemcpyAsync(&h_ovf, d_ovf, sizeof(int), cudaMemcpyDeviceToHost, st));
	CK(cudaStreamSynchronize(st));
	o.n_mini = n_mini;
	o.mini = ar.get<mb128>(n_mini + 1);
	if (!h_ovf) { // the usual case: every chunk fitted its staging row, one 16-byte copy per minimizer finishes the job
		k_sketch_compact<<<(unsigned)cdiv(n_chunks * 8, 256), 256, 0, st>>>(stage, chunk_cnt, chunk_off, n_chunks, o.mini);
	} else {      // some 256-base chunk produced more than SK_STAGE_CAP minimizers (low-complexity sequence): re-run and write in place
		k_sketch<10, true><<<(unsigned)n_cta, SK_TPB, SK_SMEM_BYTES, st>>>(codes, d_off, n_reads, total, w, k, nullptr, nullptr, chunk_off, o.mini, nullptr, nullptr, 0);
	}
	++*n_launch;
}
