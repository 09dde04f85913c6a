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
// Memory: a CTA stages its 128 chunks (+64 warm-up bytes each) into shared memory with coalesced 16-byte
// loads; chunk rows are padded by one word so the per-thread strided reads are bank-conflict free.
//
// Two kernels share the work (odd k <= 15, the map-ont case): chunks that lie in the interior of a read, with no ambiguous
// base nearby, go through k_sketch_par -- one warp per chunk, every position in parallel (see there); the chunks at read
// starts / ends, around N runs and at the edge of a copy piece go through the automaton, one thread per chunk, over a
// compacted list.  Other (w,k) use the automaton for everything.
#pragma once
#include "common.cuh"
#include "pack.cuh"

#define SK_CHUNK 256
#define SK_TPB   128
#define SK_WARM  64
#define SK_SPAN  (SK_CHUNK * SK_TPB)
#define SK_LOOK  16
#define SK_ROW   (SK_WARM + SK_CHUNK + SK_LOOK + 4)   // shared-memory row of one chunk: warm-up + chunk + look-ahead, padded by one word (odd word stride)
#define SK_SMEM_BYTES (SK_TPB * SK_ROW)

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
// passes on the minimizers whose position in the read lies in [lo, hi)
template <typename E>
struct EmitFilter {
	E &e; uint32_t lo, hi;
	MB_D void operator()(uint64_t x, uint32_t y) { const uint32_t i = y >> 1; if (i >= lo && i < hi) e(x, y); }
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
         const int64_t *__restrict__ chunk_off, mb128 *__restrict__ out, mb128 *__restrict__ stage, int *__restrict__ overflow, int64_t cta0,
         const int32_t *__restrict__ list, const int32_t *__restrict__ n_list)
{
	extern __shared__ __align__(16) uint8_t sm[];
	const unsigned FULL = 0xffffffffu;
	const int lane = threadIdx.x & 31;
	// which chunk: the CTA's 128 consecutive chunks, or 128 entries of the boundary-chunk list
	int64_t chunk = -1;
	if (list) {
		const int n = *n_list;
		if ((int64_t)blockIdx.x * SK_TPB >= n) return;
		const int64_t e = (int64_t)blockIdx.x * SK_TPB + threadIdx.x;
		if (e < n) chunk = list[e];
	} else {
		chunk = ((int64_t)blockIdx.x + cta0) * SK_TPB + threadIdx.x;
		if (chunk * SK_CHUNK >= total) chunk = -1;
	}
	// ---- each warp stages the rows of its own 32 chunks: [s - SK_WARM, s + SK_CHUNK + SK_LOOK) as 21 coalesced 16-byte loads per row ----
	{
		constexpr int VPR = (SK_WARM + SK_CHUNK + SK_LOOK) / 16;
		for (int it = 0; it < VPR; ++it) {
			const int idx = it * 32 + lane, row = idx / VPR, v = idx - row * VPR;
			const int64_t ch = __shfl_sync(FULL, chunk, row);
			if (ch < 0) continue;
			const int64_t g = ch * SK_CHUNK - SK_WARM + (int64_t)v * 16;
			uint4 val = make_uint4(0x04040404u, 0x04040404u, 0x04040404u, 0x04040404u);
			if (g >= 0 && g + 16 <= total) val = *reinterpret_cast<const uint4*>(codes + g);
			else if (g + 16 > 0 && g < total) {
				uint8_t tmp[16];
				#pragma unroll
				for (int b = 0; b < 16; ++b) tmp[b] = (g + b >= 0 && g + b < total) ? codes[g + b] : 4;
				val = *reinterpret_cast<uint4*>(tmp);
			}
			uint32_t *d = reinterpret_cast<uint32_t*>(sm + (size_t)((threadIdx.x & ~31) + row) * SK_ROW + v * 16);
			d[0] = val.x, d[1] = val.y, d[2] = val.z, d[3] = val.w;
		}
	}
	__syncwarp();
	if (chunk < 0) return;
	const int64_t s = chunk * SK_CHUNK;
	const int64_t win_lo = s - SK_WARM; // row index 0 <-> global win_lo (may be negative)
	const int64_t e = s + SK_CHUNK < total ? s + SK_CHUNK : total;
	const uint64_t mask = (1ULL << 2 * k) - 1, shift1 = 2 * (k - 1);
	const uint8_t *row_sm = sm + (size_t)threadIdx.x * SK_ROW;
	auto code_at = [&](int64_t g) -> int {
		const int64_t L = g - win_lo;
		if (L >= 0 && L < SK_WARM + SK_CHUNK + SK_LOOK) return row_sm[L];
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
		// A minimizer belongs to the chunk that holds its POSITION (the rule k_sketch_par uses), but the automaton writes it up
		// to w steps later -- when it is replaced, leaves the window, or at the final flush.  So: filter by position, and when the
		// chunk ends inside the read run on for w more steps to collect the late writes of this chunk's positions.
		const uint32_t f_lo = (uint32_t)(pos - rs), f_hi = (uint32_t)(seg_end - rs);
		const int64_t run_end = seg_end == re ? re : (seg_end + W < re ? seg_end + W : re);
		int seg_cnt;
		if (WRITE) {
			EmitWrite ew; ew.out = out; ew.pos = wpos; ew.rid_hi = (uint64_t)(uint32_t)r << 32;
			EmitFilter<EmitWrite> ef{ew, f_lo, f_hi};
			for (int64_t g = pos; g < run_end; ++g) st.step(code_at(g), (uint32_t)(g - rs), w, k, mask, shift1, ef);
			if (run_end == re && st.minx != UINT64_MAX) ef(st.minx, st.miny);
			seg_cnt = (int)(ew.pos - wpos);
			wpos = ew.pos;
		} else if (stage) {
			EmitStage es; es.row = stage + chunk * SK_STAGE_CAP; es.n = 0; es.base = total_cnt; es.rid_hi = (uint64_t)(uint32_t)r << 32;
			EmitFilter<EmitStage> ef{es, f_lo, f_hi};
			for (int64_t g = pos; g < run_end; ++g) st.step(code_at(g), (uint32_t)(g - rs), w, k, mask, shift1, ef);
			if (run_end == re && st.minx != UINT64_MAX) ef(st.minx, st.miny);
			seg_cnt = es.n;
			if (seg_cnt) atomicAdd(&read_cnt[r], seg_cnt);
		} else {
			EmitCount ec; ec.n = 0;
			EmitFilter<EmitCount> ef{ec, f_lo, f_hi};
			for (int64_t g = pos; g < run_end; ++g) st.step(code_at(g), (uint32_t)(g - rs), w, k, mask, shift1, ef);
			if (run_end == re && st.minx != UINT64_MAX) ef(st.minx, st.miny);
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

// ---- position-parallel sketch of interior chunks (odd k <= 15, w = 10) ----
// Away from read ends and ambiguous bases mm_sketch's automaton emits exactly the positions whose hash is minimal (ties
// included) in at least one full window of w consecutive k-mers, in ascending order: a minimum is written when a later
// k-mer replaces it (`<=`, so every tied minimum gets its turn) or when it leaves the window, and the rescan that follows
// writes the tied ones that never became the running minimum.  With M[e] = min h[e-9..e] the test for position p is
// max_{e in [p,p+9]} M[e] == h[p]: two log-step sliding passes in registers.  The quirks of the automaton (first window after
// a read start or an N, the `l >= w+k` emission thresholds, the running minimum lost at an N, the final flush) all live
// within w+k bases of a read end or an N; chunks that come that close are not handled here but listed for the automaton.
//   One warp per 256-base chunk: 19 lanes stage [s-32, s+272) as 2-bit codes; lane L hashes positions s-9+9L .. s-1+9L
// (both strands from one 64-bit window: the reverse strand is the complement of the field, the forward strand its pair
// reversal; the invertible hash runs in 32-bit arithmetic because 2k <= 32), lane L then tests positions s+8L .. s+8L+7 and
// the warp writes its minimizers to the chunk's staging row in position order.
#define SKP_WARPS 4
#define SKP_LO 32                       // bases staged before the chunk
#define SKP_VEC 19                      // 16-byte vectors staged: [s-32, s+272)
#define SKP_NH 288

MB_D uint32_t mb_hash32(uint32_t key, uint32_t mask)
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

// read that holds the first base of every chunk (one thread per chunk: the binary search is then not repeated by 32 lanes)
__global__ void k_chunk_read(const int64_t *__restrict__ off, int n_reads, int64_t n_chunks, int32_t *__restrict__ chunk_read)
{
	const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (c >= n_chunks) return;
	const int64_t s = c * SK_CHUNK;
	chunk_read[c] = s >= off[0] ? sk_find_read(off, n_reads, s) : 0;
}

__global__ void __launch_bounds__(SKP_WARPS * 32)
k_sketch_par(const uint8_t *__restrict__ codes, const int64_t *__restrict__ off, int n_reads, int64_t limit, int k,
             int32_t *__restrict__ chunk_cnt, int32_t *__restrict__ read_cnt, mb128 *__restrict__ stage, int *__restrict__ overflow,
             int32_t *__restrict__ blist, int32_t *__restrict__ bctr, int64_t chunk0, int64_t chunk1, const int32_t *__restrict__ chunk_read)
{
	__shared__ uint32_t s_pk[SKP_WARPS][SKP_VEC + 3];
	__shared__ uint32_t s_h[SKP_WARPS][SKP_NH];
	const unsigned FULL = 0xffffffffu;
	const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
	const int64_t chunk = chunk0 + (int64_t)blockIdx.x * SKP_WARPS + wib;
	if (chunk >= chunk1) return;
	const int64_t s = chunk * SK_CHUNK;
	// interior: [s-32, s+272) inside one read, already landed, free of ambiguous bases
	bool interior = s - SKP_LO >= off[0] && s + SK_CHUNK + 16 <= limit;
	int r = 0; int64_t rs = 0;
	if (interior) {
		r = chunk_read[chunk];
		rs = off[r];
		interior = s - SKP_LO >= rs && s + SK_CHUNK + 16 <= off[r + 1];
	}
	uint32_t amb = 0;
	if (interior && lane < SKP_VEC) {
		const uint4 v = *reinterpret_cast<const uint4*>(codes + s - SKP_LO + lane * 16);
		amb = (v.x | v.y | v.z | v.w) & 0xfcfcfcfcu;
		auto pack4 = [](uint32_t x) { const uint32_t t = x & 0x03030303u; return (t | t >> 6 | t >> 12 | t >> 18) & 0xffu; };
		s_pk[wib][lane] = pack4(v.x) | pack4(v.y) << 8 | pack4(v.z) << 16 | pack4(v.w) << 24;
	}
	if (lane >= SKP_VEC && lane < SKP_VEC + 3) s_pk[wib][lane] = 0;
	if (!interior || __any_sync(FULL, amb != 0)) {
		if (lane == 0) blist[atomicAdd(bctr, 1)] = (int32_t)chunk;
		return;
	}
	__syncwarp();
	// ---- hashes of positions s-9+9L+j, j = 0..8 (only those <= s+264 are used) ----
	const uint32_t mask = k < 16 ? (1u << 2 * k) - 1u : 0xffffffffu;
	{
		const int base0 = SKP_LO - 9 - (k - 1) + 9 * lane;        // staged index of the first base of the first k-mer
		const int wi = base0 >> 4, sh = (base0 & 15) * 2;
		const uint32_t a = s_pk[wib][wi], b = s_pk[wib][wi + 1], c = wi + 2 < SKP_VEC + 3 ? s_pk[wib][wi + 2] : 0u;
		uint64_t win = ((uint64_t)b << 32 | a) >> sh;
		if (sh) win |= (uint64_t)c << (64 - sh);
		#pragma unroll
		for (int j = 0; j < 9; ++j) {
			const uint32_t f = (uint32_t)(win >> (2 * j)) & mask;           // oldest base in the low bits
			const uint32_t k1 = ~f & mask;                                  // reverse strand: complement, newest base on top
			const uint32_t t = __brev(f << (32 - 2 * k));
			const uint32_t k0 = (t >> 1 & 0x55555555u) | (t & 0x55555555u) << 1; // forward strand: oldest base on top
			const uint32_t z = k0 < k1 ? 0u : 1u;
			s_h[wib][9 * lane + j] = mb_hash32(z ? k1 : k0, mask) | z << 31;
		}
	}
	__syncwarp();
	// ---- window test for positions s+8L+c: H[i] = hash of position s+8L-9+i ----
	uint32_t H[26], Z = 0;
	#pragma unroll
	for (int i = 0; i < 26; ++i) {
		const uint32_t v = s_h[wib][8 * lane + i];
		H[i] = v & 0x7fffffffu;
		if (i >= 9 && i < 17) Z |= (v >> 31) << (i - 9);
	}
	uint32_t m2[26], m4[26], m8[26], M[26];
	#pragma unroll
	for (int i = 1; i < 26; ++i) m2[i] = min(H[i], H[i - 1]);
	#pragma unroll
	for (int i = 3; i < 26; ++i) m4[i] = min(m2[i], m2[i - 2]);
	#pragma unroll
	for (int i = 7; i < 26; ++i) m8[i] = min(m4[i], m4[i - 4]);
	#pragma unroll
	for (int i = 9; i < 26; ++i) M[i] = min(m8[i], m2[i - 8]);       // min of H[i-9..i]
	uint32_t x2[26], x4[26], x8[26];
	#pragma unroll
	for (int i = 9; i < 25; ++i) x2[i] = max(M[i], M[i + 1]);
	#pragma unroll
	for (int i = 9; i < 23; ++i) x4[i] = max(x2[i], x2[i + 2]);
	#pragma unroll
	for (int i = 9; i < 19; ++i) x8[i] = max(x4[i], x4[i + 4]);
	uint32_t flags = 0;
	#pragma unroll
	for (int c = 0; c < 8; ++c) {
		const uint32_t X = max(x8[9 + c], x2[9 + c + 8]);               // max of M[9+c .. 18+c]
		if (X == H[9 + c]) flags |= 1u << c;
	}
	// ---- ordered emission into the chunk's staging row ----
	const int cnt = __popc(flags);
	int incl = cnt;
	#pragma unroll
	for (int d = 1; d < 32; d <<= 1) { const int v = __shfl_up_sync(FULL, incl, d); if (lane >= d) incl += v; }
	const int tot = __shfl_sync(FULL, incl, 31);
	int at = incl - cnt;
	mb128 *row = stage + chunk * SK_STAGE_CAP;
	const uint64_t rid_hi = (uint64_t)(uint32_t)r << 32;
	const uint32_t i0 = (uint32_t)(s - rs) + 8u * lane;
	#pragma unroll
	for (int c = 0; c < 8; ++c) {
		if (flags >> c & 1) {
			if (at < SK_STAGE_CAP) {
				row[at].x = (uint64_t)H[9 + c] << 8 | (uint64_t)k;
				row[at].y = rid_hi | (uint64_t)((i0 + c) << 1 | (Z >> c & 1u));
			}
			++at;
		}
	}
	if (lane == 0) {
		chunk_cnt[chunk] = tot;
		if (tot > SK_STAGE_CAP) *overflow = 1;
		if (tot) atomicAdd(&read_cnt[r], tot);
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
	// ... or 2-bit packed words (pack.cuh): h_words[k] / d_words[k] hold bases [16k - sh/2, 16k - sh/2 + 16) of this batch
	const uint32_t *h_words = nullptr;
	uint32_t *d_words = nullptr;
	int sh = 0;                         // 2 x (bases of the first word that precede the batch)
	const int64_t *d_iv = nullptr; int64_t n_iv = 0, g0 = 0; // runs of ambiguous bases that touch the batch (global coordinates; g0 = the batch's first base)
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
	// odd k <= 15: interior chunks position-parallel, the rest (listed in blist) through the automaton
	const bool par = (k & 1) && k <= 15 && n_chunks < INT32_MAX;
	int32_t *blist = par ? ar.get<int32_t>(n_chunks) : nullptr;
	int32_t *bctr = par ? ar.get<int32_t>(1) : nullptr;
	int32_t *chunk_read = par ? ar.get<int32_t>(n_chunks) : nullptr;
	if (par) {
		CK(cudaMemsetAsync(bctr, 0, sizeof(int32_t), st));
		k_chunk_read<<<(unsigned)cdiv(n_chunks, 256), 256, 0, st>>>(d_off, n_reads, n_chunks, chunk_read); ++*n_launch;
	}
	if (!feed) {
		if (par) {
			k_sketch_par<<<(unsigned)cdiv(n_chunks, SKP_WARPS), SKP_WARPS * 32, 0, st>>>(codes, d_off, n_reads, total, k, chunk_cnt, read_cnt, stage, d_ovf, blist, bctr, 0, n_chunks, chunk_read);
			++*n_launch;
		}
		k_sketch<10, false><<<(unsigned)n_cta, SK_TPB, SK_SMEM_BYTES, st>>>(codes, d_off, n_reads, total, w, k, chunk_cnt, read_cnt, nullptr, nullptr, stage, d_ovf, 0, blist, bctr);
		++*n_launch;
	} else {
		// pieces of ~1/8 of the batch, whole CTA spans each; piece p is sketched while piece p+1 is in flight
		int64_t per = cdiv(cdiv(n_cta, 8), 1); if (per < 1) per = 1;
		int p = 0;
		for (int64_t c0 = 0; c0 < n_cta; c0 += per, ++p) {
			const int64_t c1 = c0 + per < n_cta ? c0 + per : n_cta;
			const int64_t b0 = c0 * SK_SPAN, b1 = c1 * SK_SPAN < total ? c1 * SK_SPAN : total;
			while ((int)feed->events->size() <= p) { cudaEvent_t e; CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); feed->events->push_back(e); }
			if (feed->h_words) { // b0 is a multiple of 16: the piece's words start at b0 / 16, the last thread reads one word ahead
				const int64_t k0w = b0 >> 4, k1w = ((feed->sh / 2 + b1 + 15) >> 4) + 1;
				CK(cudaMemcpyAsync(feed->d_words + k0w, feed->h_words + k0w, (size_t)(k1w - k0w) * 4, cudaMemcpyHostToDevice, feed->copy_st));
			} else CK(cudaMemcpyAsync(feed->d_ascii + b0, feed->h_ascii + b0, (size_t)(b1 - b0), cudaMemcpyHostToDevice, feed->copy_st));
			CK(cudaEventRecord((*feed->events)[p], feed->copy_st));
			CK(cudaStreamWaitEvent(st, (*feed->events)[p], 0));
			if (feed->h_words) {
				k_unpack_nt4<<<(unsigned)cdiv(cdiv(b1 - b0, 16), 256), 256, 0, st>>>(feed->d_words + (b0 >> 4), feed->sh, feed->d_codes + b0, b1 - b0);
				if (feed->n_iv) { k_apply_amb<<<(unsigned)cdiv(feed->n_iv * 32, 256), 256, 0, st>>>(feed->d_iv, feed->n_iv, feed->g0, feed->d_codes, b0, b1); ++*n_launch; }
			} else k_encode_nt4<<<(unsigned)cdiv(cdiv(b1 - b0, 16), 256), 256, 0, st>>>(feed->d_ascii + b0, feed->d_codes + b0, b1 - b0);
			if (par) {
				const int64_t k0 = c0 * SK_TPB, k1 = c1 * SK_TPB < n_chunks ? c1 * SK_TPB : n_chunks;
				k_sketch_par<<<(unsigned)cdiv(k1 - k0, SKP_WARPS), SKP_WARPS * 32, 0, st>>>(codes, d_off, n_reads, b1, k, chunk_cnt, read_cnt, stage, d_ovf, blist, bctr, k0, k1, chunk_read);
			}
			*n_launch += 2;
		}
		// every piece has landed: the listed boundary chunks (or, for other k, all chunks: the automaton looks w bases past its
		// chunk) in one launch; CTAs beyond the list return at once
		k_sketch<10, false><<<(unsigned)n_cta, SK_TPB, SK_SMEM_BYTES, st>>>(codes, d_off, n_reads, total, w, k, chunk_cnt, read_cnt, nullptr, nullptr, stage, d_ovf, 0, blist, bctr);
		++*n_launch;
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
		k_sketch<10, true><<<(unsigned)n_cta, SK_TPB, SK_SMEM_BYTES, st>>>(codes, d_off, n_reads, total, w, k, nullptr, nullptr, chunk_off, o.mini, nullptr, nullptr, 0, nullptr, nullptr);
	}
	++*n_launch;
}
