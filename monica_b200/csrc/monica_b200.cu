// monica_b200.cu -- C-ABI implementation (include/monica_b200.h) and the per-batch device pipeline.
//
// Pipeline of one batch (all on the calling thread's stream; the host only learns array sizes between stages):
//   H2D reads -> k_encode_nt4 -> K1 k_sketch -> K2 k_seed_lookup/k_seed_fill -> K2b k_sort_anchors
//   -> K3 k_chain_dp -> k_chain_bt / k_gen_regs (region logic) -> rounds of { k_plan1, k_plan2, K4 k_dp, k_ztest, k_dp, k_stitch }
//   -> k_finish -> k_write_hits -> D2H hits.
// There is no CPU fallback anywhere in this file: every compute entry point throws MB_ERR_CUDA without a device.
#include <algorithm>
#include <chrono>
#include <map>
#include <atomic>
#include <memory>
#include <mutex>
#include <thread>
#include <zlib.h>
#include <dlfcn.h>
#include "common.cuh"
#include "sketch.cuh"
#include "seed.cuh"
#include "index_build.cuh"
#include "chain.cuh"
#include "glue.cuh"
#include "align.cuh"
#include "align2.cuh"
#include "dp_fast.cuh"
#include "dp_fast_chain.cuh"
#include "dp_ext.cuh"
#include "dp_band.cuh"
#include "align_cta.cuh"

// zlib reports a stream that ends before its trailer (a truncated .gz) through gzerror() only: gzread() hands out the bytes
// it could inflate and then returns 0 like at a proper end of file.
static inline bool mb_gz_clean_eof(gzFile fp)
{
	int e = Z_OK;
	gzerror(fp, &e);
	return e == Z_OK || e == Z_STREAM_END;
}

thread_local std::string g_mb_err;
static thread_local std::chrono::steady_clock::time_point g_dbg_t0 = std::chrono::steady_clock::now();

extern "C" const char *mb_last_error(void) { return g_mb_err.c_str(); }

#define API_BEGIN try {
#define API_END   } catch (const mb_error &e) { g_mb_err = e.what(); return e.code; } \
                    catch (const std::exception &e) { g_mb_err = e.what(); return MB_ERR_ARG; } \
                    return MB_OK;

extern "C" int mb_device_count(void)
{
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
	return n;
}

extern "C" int mb_opt_init(mb_opt_t *opt)
{
	if (!opt) return MB_ERR_ARG;
	memset(opt, 0, sizeof(*opt));
	opt->seed = 11;
	opt->mid_occ_frac = 2e-4f;
	opt->min_cnt = 3, opt->min_chain_score = 40, opt->bw = 500, opt->max_gap = 5000, opt->max_gap_ref = -1;
	opt->max_chain_skip = 25, opt->max_chain_iter = 5000;
	opt->mask_level = 0.5f, opt->pri_ratio = 0.8f, opt->best_n = 5;
	opt->max_join_long = 20000, opt->max_join_short = 2000, opt->min_join_flank_sc = 1000, opt->min_join_flank_ratio = 0.5f;
	opt->a = 2, opt->b = 4, opt->q = 4, opt->e = 2, opt->q2 = 24, opt->e2 = 1, opt->sc_ambi = 1;
	opt->zdrop = 400, opt->zdrop_inv = 200, opt->end_bonus = -1;
	opt->min_dp_max = opt->min_chain_score * opt->a;
	opt->min_ksw_len = 200;
	opt->max_clip_ratio = 1.0f;
	opt->max_sw_mat = 100000000;
	opt->mid_occ = 0;
	return MB_OK;
}

// ---------------------------------------------------------------------------------------------
// per-thread device context
// ---------------------------------------------------------------------------------------------
#define MB_NSIDE 4
struct ThreadCtx {
	int device = -1;
	cudaStream_t st = nullptr;
	// high-priority side streams: the long-tailed exact DP launches and the small extension launches overlap the fast DP
	// kernels.  Few of them on purpose: streams beyond the device's hardware queues (8 by default) alias and serialise.
	cudaStream_t st2[MB_NSIDE] = {};
	cudaEvent_t ev_join[MB_NSIDE] = {};
	cudaEvent_t ev_fast_done = nullptr;   // recorded behind the last k_dp_fast launch of a DpRunner::run
	// the k_dp_fast launches of the column classes rotate over st and these two streams, so the drain of one launch (a few
	// warps still on their last task pair) overlaps the ramp-up of the next instead of idling the GPU 13 times per pass
	cudaStream_t stf[2] = {};
	cudaEvent_t ev_f[2] = {}, ev_fork = nullptr;
	cudaEvent_t ev_x[5] = {}, ev_fork2 = nullptr;
	cudaStream_t st_defer[3] = {}; cudaEvent_t ev_defer[3] = {};   // long exact END EXTENSIONS of a first pass: run behind the gap-fill kernels, joined before the stitch   // extension / band launches done (before the gap-fill launches start)
	std::vector<cudaEvent_t> feed_events;   // host->device pieces of mb_map_batch (SketchFeed)
	int64_t piece_bases = 0;                // bases per sequential piece once a batch did not fit the device (0: the default)
	uint64_t piece_index = 0;               // ... for this index (mb_index::uid)
	Arena ar;
	int num_sms = 148;
	uint8_t *h_pin = nullptr; size_t h_pin_cap = 0;   // pinned staging for reads
	unsigned long long *d_counts = nullptr; int n_counts = 0; // last mb_count vector
	// device-resident result of the last mapping call, one entry per sub-batch (lives in the arenas until the next reset)
	struct LastPart { const int32_t *fields; const int64_t *hit_off, *read_off; int64_t n_hits; int32_t n_reads; };
	std::vector<LastPart> last_parts;
	int32_t last_n_reads = -1; uint64_t last_index = 0;   // mb_index::uid of the last mapped batch
	// batches cut into sequential pieces (memory budget): the hit fields / offsets of every piece are kept here (plain device
	// memory, grown on demand, reused by later batches) so that mb_count_last sees the whole batch
	char *store = nullptr; size_t store_cap = 0, store_used = 0;
	// helper contexts (own stream + arena) for the other sub-batches of a call: a batch is cut in MB_NPART pieces that run
	// concurrently, so the latency-bound stages of one piece (sketch, seeding, chaining, region logic, stitching) hide under
	// the issue-bound DP kernels of the other
	std::vector<ThreadCtx*> helpers;
	~ThreadCtx() {
		for (ThreadCtx *h : helpers) delete h;
		if (device >= 0) {
			cudaSetDevice(device);
			ar.release();
			if (h_pin) cudaFreeHost(h_pin);
			if (store) cudaFree(store);
			if (d_counts) cudaFree(d_counts);
			if (st) cudaStreamDestroy(st);
			for (int i = 0; i < MB_NSIDE; ++i) { if (st2[i]) cudaStreamDestroy(st2[i]); if (ev_join[i]) cudaEventDestroy(ev_join[i]); }
			if (ev_fast_done) cudaEventDestroy(ev_fast_done);
			for (int i = 0; i < 2; ++i) { if (stf[i]) cudaStreamDestroy(stf[i]); if (ev_f[i]) cudaEventDestroy(ev_f[i]); }
			if (ev_fork) cudaEventDestroy(ev_fork);
			if (ev_fork2) cudaEventDestroy(ev_fork2);
			for (int i = 0; i < 5; ++i) if (ev_x[i]) cudaEventDestroy(ev_x[i]);
			for (int i = 0; i < 3; ++i) { if (st_defer[i]) cudaStreamDestroy(st_defer[i]); if (ev_defer[i]) cudaEventDestroy(ev_defer[i]); }
			for (cudaEvent_t e : feed_events) cudaEventDestroy(e);
		}
	}
};
// One context per (calling thread, device).  monica's multi_threaded_aligner creates a fresh ThreadPool per call (aligner.py:89):
// a context that outlived its thread would leak its arena (several GB), and one that died with it would make every file of a run
// pay for streams, events and a cold arena again.  So a thread that exits hands its contexts to a small process-wide pool and the
// next new thread takes one from there (warm arena, sized by the batches it has seen); contexts beyond the pool's size are freed.
struct CtxPool { std::mutex mu; std::vector<ThreadCtx*> idle[16]; };
static CtxPool *g_ctx_pool = new CtxPool();   // never destroyed: a context must not be freed after the CUDA runtime has shut down
#define MB_CTX_POOL_MAX 4
struct ThreadCtxMap {
	std::map<int, ThreadCtx*> m;
	~ThreadCtxMap() {
		for (auto &kv : m) {
			ThreadCtx *c = kv.second;
			if (cudaSetDevice(c->device) == cudaSuccess) cudaStreamSynchronize(c->st);
			cudaGetLastError();
			c->last_parts.clear(); c->last_index = 0; c->last_n_reads = 0;
			bool kept = false;
			{
				std::lock_guard<std::mutex> g(g_ctx_pool->mu);
				auto &v = g_ctx_pool->idle[c->device & 15];
				if ((int)v.size() < MB_CTX_POOL_MAX) { v.push_back(c); kept = true; }
			}
			if (!kept) delete c;
		}
	}
};
static thread_local ThreadCtxMap t_ctx_holder;
#define t_ctx (t_ctx_holder.m)

static void ensure_device(int device)
{
	int n = 0;
	cudaError_t e = cudaGetDeviceCount(&n);
	if (e != cudaSuccess || n == 0) { cudaGetLastError(); throw mb_error(MB_ERR_CUDA, "no CUDA device available: monica_b200 has no CPU fallback"); }
	if (device < 0 || device >= n) throw mb_error(MB_ERR_ARG, "bad device ordinal");
	CK(cudaSetDevice(device));
}

static std::atomic<long long> g_scratch_per_base_x16{0};   // arena bytes per read base of the last batch (x16), for a cold arena's first reservation
static std::once_flag g_const_once[16];
static std::mutex g_dp_mutex[16];   // per device: serialises the DP stage of concurrent pieces / calling threads
static int64_t mb_dp_lock_min_bases()
{
	// 0: every batch takes its turn.  Measured (4,000-read batches, 2 calling threads): letting small batches overlap in the DP
	// stage is slower (0.55 vs 0.71 Gbases/s) -- their grids are sized for the whole GPU and get in each other's way
	static const int64_t v = getenv("MB_DP_LOCK_MIN_BASES") ? atoll(getenv("MB_DP_LOCK_MIN_BASES")) : 0;
	return v;
}

static ThreadCtx *make_ctx(int device);
static ThreadCtx &get_ctx(int device)
{
	ensure_device(device);
	auto it = t_ctx.find(device);
	if (it != t_ctx.end()) return *it->second;
	ThreadCtx *c = nullptr;
	{
		std::lock_guard<std::mutex> g(g_ctx_pool->mu);
		auto &v = g_ctx_pool->idle[device & 15];
		for (size_t i = 0; i < v.size(); ++i) if (v[i]->device == device) { c = v[i]; v.erase(v.begin() + i); break; }
	}
	if (!c) c = make_ctx(device);
	t_ctx[device] = c;
	return *c;
}

static ThreadCtx *make_ctx(int device)
{
	ThreadCtx *c = new ThreadCtx();
	c->device = device;
	{
		// three priority levels: side streams (exact / band / extension kernels of the running pass) > the main and gap-fill
		// streams > the deferred exact end extensions, whose CTAs must not queue ahead of the Z-drop test's small kernels
		int lo = 0, hi = 0;
		CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
		const int mid = (lo - hi >= 2) ? lo - 1 : lo;
		CK(cudaStreamCreateWithPriority(&c->st, cudaStreamNonBlocking, mid));
		CK(cudaEventCreateWithFlags(&c->ev_fast_done, cudaEventDisableTiming));
		CK(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
		CK(cudaEventCreateWithFlags(&c->ev_fork2, cudaEventDisableTiming));
		for (int i = 0; i < 3; ++i) { CK(cudaStreamCreateWithPriority(&c->st_defer[i], cudaStreamNonBlocking, lo)); CK(cudaEventCreateWithFlags(&c->ev_defer[i], cudaEventDisableTiming)); }
		for (int i = 0; i < 5; ++i) CK(cudaEventCreateWithFlags(&c->ev_x[i], cudaEventDisableTiming));
		for (int i = 0; i < 2; ++i) { CK(cudaStreamCreateWithPriority(&c->stf[i], cudaStreamNonBlocking, mid)); CK(cudaEventCreateWithFlags(&c->ev_f[i], cudaEventDisableTiming)); }
		for (int i = 0; i < MB_NSIDE; ++i) {
			CK(cudaStreamCreateWithPriority(&c->st2[i], cudaStreamNonBlocking, getenv("MB_SIDE_PRIO0") ? 0 : hi));
			CK(cudaEventCreateWithFlags(&c->ev_join[i], cudaEventDisableTiming));
		}
	}
	cudaDeviceProp prop;
	CK(cudaGetDeviceProperties(&prop, device));
	c->num_sms = prop.multiProcessorCount;
	std::call_once(g_const_once[device & 15], [&]() {
		CK(cudaMemcpyToSymbol(c_nt4, h_nt4, 256));
		CK(cudaFuncSetAttribute(k_dp, cudaFuncAttributeMaxDynamicSharedMemorySize, DP_SMEM_MAX));
		CK(cudaFuncSetAttribute(k_dp_cta<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, DPC_SMEM_MAX));
	});
	return c;
}

template <typename T> static T d2h_scalar(const T *d, cudaStream_t st)
{
	T v; CK(cudaMemcpyAsync(&v, d, sizeof(T), cudaMemcpyDeviceToHost, st)); CK(cudaStreamSynchronize(st)); return v;
}

// ---------------------------------------------------------------------------------------------
// index
// ---------------------------------------------------------------------------------------------
static std::atomic<uint64_t> g_index_uid{1};
struct mb_index {
	const uint64_t uid = g_index_uid.fetch_add(1);   // never reused (a pointer is: a new index may land where a freed one was)
	int device = 0;
	int k = 15, w = 10, b = 14;
	std::vector<std::string> names;
	std::vector<uint32_t> lens;
	std::vector<uint64_t> offs;
	uint64_t sum_len = 0;
	int64_t n_mini = 0, n_keys = 0;
	int mid_occ = 0;
	// host copies (needed by mb_index_save only; filled lazily when the table was built on the device)
	std::vector<uint64_t> h_hkey, h_hval, h_pos;
	std::vector<uint32_t> h_S;
	bool host_copies = true;
	size_t cap = 0, n_pos = 0;
	DevIndex d;
	int64_t hbm_bytes = 0;
};

__global__ void k_pack4(const uint8_t *__restrict__ codes, uint32_t *__restrict__ S, int64_t n)
{
	int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (w * 8 >= n) return;
	uint32_t v = 0;
	#pragma unroll
	for (int j = 0; j < 8; ++j) {
		int64_t i = w * 8 + j;
		uint32_t c = i < n ? codes[i] : 0;
		v |= c << (j << 2);
	}
	S[w] = v;
}

// minimizers (device, grouped arbitrarily, positions of one hash ascending) + 4-bit packed sequence (device) -> HBM-resident index
static void index_finish_device(mb_index *ix, ThreadCtx &c, mb128 *d_mini, int64_t n_m, const uint32_t *d_S, int k)
{
	cudaStream_t st = c.st;
	const uint64_t sum = ix->sum_len;
	int64_t nl = 0;
	ix->n_mini = n_m;
	DevIndex &d = ix->d;
	int64_t n_keys = 0, n_pos = 0;
	mb128 *sorted = d_mini;
	int64_t *start = nullptr, *moff = nullptr;
	unsigned int *occ_hist = c.ar.get<unsigned int>(65536);
	CK(cudaMemsetAsync(occ_hist, 0, 65536 * sizeof(unsigned int), st));
	if (n_m > 0) {
		mb128 *tmp = c.ar.get<mb128>(n_m);
		sorted = radix_sort_minimizers(c.ar, st, d_mini, tmp, n_m, 2 * k);
		int32_t *flag = c.ar.get<int32_t>(n_m);
		int64_t *gid = c.ar.get<int64_t>(n_m + 1);
		k_ib_flag<<<(unsigned)cdiv(n_m, 256), 256, 0, st>>>(sorted, n_m, flag);
		exclusive_scan<int32_t>(c.ar, st, flag, gid, n_m, &nl);
		n_keys = d2h_scalar(gid + n_m, st);
		start = c.ar.get<int64_t>(n_keys + 1);
		k_ib_start<<<(unsigned)cdiv(n_m, 256), 256, 0, st>>>(flag, gid, n_m, n_keys, start);
		int32_t *mcnt = c.ar.get<int32_t>(n_keys);
		moff = c.ar.get<int64_t>(n_keys + 1);
		k_ib_multi<<<(unsigned)cdiv(n_keys, 256), 256, 0, st>>>(start, n_keys, mcnt, occ_hist);
		exclusive_scan<int32_t>(c.ar, st, mcnt, moff, n_keys, &nl);
		n_pos = d2h_scalar(moff + n_keys, st);
	}
	ix->n_keys = n_keys;
	size_t cap = 1024; while (cap < (size_t)n_keys * 2 + 2) cap <<= 1;
	int bits = 0; while (((size_t)1 << bits) < cap) ++bits;
	ix->cap = cap, ix->n_pos = (size_t)n_pos;
	CK(cudaMalloc(&d.hkey, cap * 8)); CK(cudaMalloc(&d.hval, cap * 8));
	CK(cudaMalloc(&d.pos, ((size_t)n_pos + 1) * 8));
	CK(cudaMalloc(&d.S, ((size_t)(sum + 7) / 8 + 1) * 4));
	CK(cudaMalloc(&d.seq_off, (ix->offs.size() + 1) * 8));
	CK(cudaMalloc(&d.seq_len, (ix->lens.size() + 1) * 4));
	CK(cudaMemsetAsync(d.hkey, 0xff, cap * 8, st));
	CK(cudaMemsetAsync(d.hval, 0, cap * 8, st));
	d.hmask = cap - 1; d.hshift = 64 - bits;
	if (n_keys > 0) k_ib_insert<<<(unsigned)cdiv(n_keys, 256), 256, 0, st>>>(sorted, start, moff, n_keys, (unsigned long long*)d.hkey, d.hval, d.hmask, d.hshift, d.pos);
	if (sum) CK(cudaMemcpyAsync(d.S, d_S, ((size_t)(sum + 7) / 8) * 4, cudaMemcpyDeviceToDevice, st));
	CK(cudaMemcpyAsync(d.seq_off, ix->offs.data(), ix->offs.size() * 8, cudaMemcpyHostToDevice, st));
	CK(cudaMemcpyAsync(d.seq_len, ix->lens.data(), ix->lens.size() * 4, cudaMemcpyHostToDevice, st));
	std::vector<unsigned int> h_occ(65536);
	CK(cudaMemcpyAsync(h_occ.data(), occ_hist, 65536 * sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
	CK(cudaStreamSynchronize(st));
	CK(cudaGetLastError());
	// index.c mm_idx_cal_max_occ(mi, 2e-4): ks_ksmall(counts, (1-f)*n) + 1, from the occurrence histogram
	if (n_keys > 0) {
		size_t kth = (size_t)(uint32_t)((1. - (double)2e-4f) * (double)n_keys);
		if (kth >= (size_t)n_keys) kth = (size_t)n_keys - 1;
		size_t cum = 0; int v = 0;
		for (v = 0; v < 65536; ++v) { cum += h_occ[v]; if (cum > kth) break; }
		ix->mid_occ = v + 1;
	} else ix->mid_occ = 1;
	d.n_seq = (int)ix->names.size(); d.k = ix->k; d.w = ix->w; d.mid_occ = ix->mid_occ;
	d.max_seq_len = 0; for (uint32_t l : ix->lens) d.max_seq_len = std::max(d.max_seq_len, l);
	ix->hbm_bytes = (int64_t)(cap * 16 + (size_t)n_pos * 8 + ((size_t)(sum + 7) / 8) * 4 + ix->offs.size() * 12);
	ix->host_copies = false;
}

static mb_index *index_build_impl(int device, int n_seq, const char *const *names, const uint8_t *const *seqs, const int64_t *lens, int w, int k)
{
	ThreadCtx &c = get_ctx(device);
	std::unique_ptr<mb_index> ix(new mb_index());
	ix->device = device, ix->k = k, ix->w = w;
	uint64_t sum = 0;
	for (int i = 0; i < n_seq; ++i) {
		if (lens[i] < 0 || lens[i] > 0x7fffffffLL) throw mb_error(MB_ERR_ARG, "contig longer than 2^31");
		ix->names.push_back(names[i]); ix->lens.push_back((uint32_t)lens[i]); ix->offs.push_back(sum); sum += (uint64_t)lens[i];
	}
	ix->sum_len = sum;
	c.ar.reset();
	// ASCII + codes + 4-bit sequence (2.5 B/base), the sketch's staging rows (5 B/base), ~0.19 minimizers/base x (2 x 16 B records
	// + 32 B of grouping arrays), scans: ~20 B/base
	if (sum > ((uint64_t)64 << 20)) c.ar.reserve((size_t)sum * 20 + ((size_t)256 << 20));
	cudaStream_t st = c.st;
	// upload contigs as one "read batch": the sketch kernel stores the sequence index in y>>32, which is mm_idx's rid
	std::vector<int64_t> off(n_seq + 1, 0);
	for (int i = 0; i < n_seq; ++i) off[i + 1] = off[i] + lens[i];
	uint8_t *d_ascii = c.ar.get<uint8_t>(sum + 16), *d_codes = c.ar.get<uint8_t>(sum + 16);
	int64_t *d_off = c.ar.get<int64_t>(n_seq + 1);
	for (int i = 0; i < n_seq; ++i) if (lens[i]) CK(cudaMemcpyAsync(d_ascii + off[i], seqs[i], lens[i], cudaMemcpyHostToDevice, st));
	CK(cudaMemcpyAsync(d_off, off.data(), (n_seq + 1) * 8, cudaMemcpyHostToDevice, st));
	int64_t nl = 0;
	if (sum) k_encode_nt4<<<(unsigned)cdiv(cdiv((int64_t)sum, 16), 256), 256, 0, st>>>(d_ascii, d_codes, (int64_t)sum);
	SketchOut so;
	run_sketch(c.ar, st, d_codes, d_off, n_seq, (int64_t)sum, w, k, so, &nl);
	// ---- sort, group, hash-insert on the device (index_build.cuh) ----
	uint32_t *d_S = c.ar.get<uint32_t>(sum / 8 + 2);
	if (sum) k_pack4<<<(unsigned)cdiv(cdiv((int64_t)sum, 8), 256), 256, 0, st>>>(d_codes, d_S, (int64_t)sum);
	index_finish_device(ix.get(), c, so.mini, so.n_mini, d_S, k);
	c.ar.reset();
	if (c.ar.cap > ((size_t)8 << 30)) c.ar.release(); // a one-off build of a large index: hand the scratch back
	return ix.release();
}

// host copies of a device-built table (only mb_index_save needs them)
static void index_ensure_host(mb_index *ix)
{
	if (ix->host_copies) return;
	CK(cudaSetDevice(ix->device));
	ix->h_hkey.resize(ix->cap); ix->h_hval.resize(ix->cap); ix->h_pos.resize(ix->n_pos); ix->h_S.resize((ix->sum_len + 7) / 8);
	CK(cudaMemcpy(ix->h_hkey.data(), ix->d.hkey, ix->cap * 8, cudaMemcpyDeviceToHost));
	CK(cudaMemcpy(ix->h_hval.data(), ix->d.hval, ix->cap * 8, cudaMemcpyDeviceToHost));
	if (ix->n_pos) CK(cudaMemcpy(ix->h_pos.data(), ix->d.pos, ix->n_pos * 8, cudaMemcpyDeviceToHost));
	if (!ix->h_S.empty()) CK(cudaMemcpy(ix->h_S.data(), ix->d.S, ix->h_S.size() * 4, cudaMemcpyDeviceToHost));
	ix->host_copies = true;
}

extern "C" int mb_index_build(int device, int n_seq, const char *const *names, const uint8_t *const *seqs, const int64_t *lens, int w, int k, mb_index_t **out)
{
	API_BEGIN
	if (!out || n_seq < 0) throw mb_error(MB_ERR_ARG, "bad arguments");
	*out = index_build_impl(device, n_seq, names, seqs, lens, w, k);
	API_END
}

extern "C" int mb_index_build_fasta(int device, const char *path, int w, int k, mb_index_t **out)
{
	API_BEGIN
	gzFile fp = gzopen(path, "rb");
	if (!fp) throw mb_error(MB_ERR_IO, std::string("cannot open ") + path);
	gzbuffer(fp, 1 << 20);
	std::vector<std::string> names; std::vector<std::string> seqs;
	// whole lines are appended with memchr / append (no per-character work); a line cut by the read buffer is carried over
	std::vector<char> buf(1 << 24);
	std::string carry;
	int n;
	auto take_line = [&](const char *l, size_t len) {
		while (len && (l[len - 1] == '\r' || l[len - 1] == '\n')) --len;
		if (len == 0) return;
		if (l[0] == '>') {
			size_t e = 1;
			while (e < len && l[e] != ' ' && l[e] != '\t') ++e;
			names.emplace_back(l + 1, e - 1);
			seqs.emplace_back();
			seqs.back().reserve(seqs.size() > 1 ? seqs[seqs.size() - 2].size() + 1024 : (size_t)1 << 20);
		} else if (!seqs.empty()) seqs.back().append(l, len);
	};
	while ((n = gzread(fp, buf.data(), (unsigned)buf.size())) > 0) {
		const char *p = buf.data(), *end = p + n;
		while (p < end) {
			const char *nl = (const char*)memchr(p, '\n', (size_t)(end - p));
			if (!nl) { carry.append(p, (size_t)(end - p)); break; }
			if (!carry.empty()) { carry.append(p, (size_t)(nl - p)); take_line(carry.data(), carry.size()); carry.clear(); }
			else take_line(p, (size_t)(nl - p));
			p = nl + 1;
		}
	}
	const bool read_ok = n == 0 && mb_gz_clean_eof(fp);
	if (!carry.empty()) take_line(carry.data(), carry.size());
	gzclose(fp);
	if (!read_ok) throw mb_error(MB_ERR_IO, std::string("read error in ") + path);
	if (names.empty()) throw mb_error(MB_ERR_IO, std::string("no sequences in ") + path);
	std::vector<const char*> np; std::vector<const uint8_t*> sp; std::vector<int64_t> lp;
	for (size_t i = 0; i < names.size(); ++i) { np.push_back(names[i].c_str()); sp.push_back((const uint8_t*)seqs[i].data()); lp.push_back((int64_t)seqs[i].size()); }
	*out = index_build_impl(device, (int)names.size(), np.data(), sp.data(), lp.data(), w, k);
	API_END
}

// ---- on-disk format ----
// Written in minimap2's .mmi layout (index.c mm_idx_dump): "MMI\2", u32 w,k,b,n_seq,flag; per sequence u8 name_len, name,
// u32 len; per bucket (2^b of them) i32 n, u64 p[n], u32 size, (u64 key, u64 val)[size]; then the 4-bit packed sequence.
// Bucket = hash & (2^b-1); key = hash>>b<<1 | singleton; val = position or offset<<32|n into the bucket's p[].
extern "C" int mb_index_save(const mb_index_t *ix, const char *path)
{
	API_BEGIN
	if (!ix || !path) throw mb_error(MB_ERR_ARG, "bad arguments");
	index_ensure_host(const_cast<mb_index*>(ix));
	FILE *fp = fopen(path, "wb");
	if (!fp) throw mb_error(MB_ERR_IO, std::string("cannot write ") + path);
	const int b = ix->b;
	uint32_t x[5] = { (uint32_t)ix->w, (uint32_t)ix->k, (uint32_t)b, (uint32_t)ix->names.size(), 0 };
	fwrite("MMI\2", 1, 4, fp);
	fwrite(x, 4, 5, fp);
	for (size_t i = 0; i < ix->names.size(); ++i) {
		if (ix->names[i].size() > 255) { fclose(fp); throw mb_error(MB_ERR_ARG, "contig name longer than 255 bytes cannot be stored in .mmi"); }
		uint8_t l = (uint8_t)ix->names[i].size();
		fwrite(&l, 1, 1, fp);
		fwrite(ix->names[i].data(), 1, l, fp);
		fwrite(&ix->lens[i], 4, 1, fp);
	}
	// regroup the table by bucket
	struct Ent { uint64_t h, key_val; bool single; uint32_t n; uint64_t start; };
	std::vector<std::vector<Ent>> B((size_t)1 << b);
	for (size_t s = 0; s < ix->h_hkey.size(); ++s) {
		uint64_t kk = ix->h_hkey[s];
		if (kk == ~0ULL) continue;
		Ent e; e.h = kk >> 1; e.single = kk & 1;
		if (e.single) e.key_val = ix->h_hval[s], e.n = 1, e.start = 0;
		else e.key_val = 0, e.n = (uint32_t)ix->h_hval[s], e.start = ix->h_hval[s] >> 32;
		B[e.h & (((uint64_t)1 << b) - 1)].push_back(e);
	}
	for (auto &bk : B) {
		std::sort(bk.begin(), bk.end(), [](const Ent &a, const Ent &c) { return a.h < c.h; });
		std::vector<uint64_t> p;
		std::vector<uint64_t> kv;
		for (auto &e : bk) {
			uint64_t key = e.h >> b << 1 | (e.single ? 1 : 0), val;
			if (e.single) val = e.key_val;
			else {
				val = (uint64_t)p.size() << 32 | e.n;
				for (uint32_t k2 = 0; k2 < e.n; ++k2) p.push_back(ix->h_pos[e.start + k2]);
			}
			kv.push_back(key); kv.push_back(val);
		}
		int32_t n = (int32_t)p.size();
		uint32_t size = (uint32_t)bk.size();
		fwrite(&n, 4, 1, fp);
		if (n) fwrite(p.data(), 8, p.size(), fp);
		fwrite(&size, 4, 1, fp);
		if (size) fwrite(kv.data(), 8, kv.size(), fp);
	}
	fwrite(ix->h_S.data(), 4, (ix->sum_len + 7) / 8, fp);
	if (fclose(fp) != 0) throw mb_error(MB_ERR_IO, "write failed");
	API_END
}

extern "C" int mb_index_load(int device, const char *path, mb_index_t **out)
{
	API_BEGIN
	get_ctx(device);
	struct FileGuard { FILE *f; ~FileGuard() { if (f) fclose(f); } } guard{ fopen(path, "rb") };
	FILE *fp = guard.f;
	if (!fp) throw mb_error(MB_ERR_IO, std::string("cannot open ") + path);
	std::unique_ptr<mb_index> ix(new mb_index());
	auto fail = [&](const char *m) { throw mb_error(MB_ERR_IO, std::string(m) + ": " + path); };
	char magic[4]; uint32_t x[5];
	if (fread(magic, 1, 4, fp) != 4 || memcmp(magic, "MMI\2", 4) != 0) fail("damaged or empty index (bad magic)");
	if (fread(x, 4, 5, fp) != 5) fail("damaged index header");
	ix->device = device, ix->w = (int)x[0], ix->k = (int)x[1], ix->b = (int)x[2];
	if (x[4] & 3) fail("HPC / no-sequence .mmi flags are not supported");
	if (x[2] < 1 || x[2] > 28) fail("damaged index header (bucket bits)");
	if (x[1] < 1 || x[1] > 28 || x[0] < 1 || x[0] > 255) fail("damaged index header (k, w)");
	uint64_t sum = 0;
	for (uint32_t i = 0; i < x[3]; ++i) {
		uint8_t l; char nm[256]; uint32_t len;
		if (fread(&l, 1, 1, fp) != 1) fail("damaged index (names)");
		if (l && fread(nm, 1, l, fp) != l) fail("damaged index (names)");
		if (fread(&len, 4, 1, fp) != 1) fail("damaged index (names)");
		ix->names.emplace_back(nm, l); ix->lens.push_back(len); ix->offs.push_back(sum); sum += len;
	}
	ix->sum_len = sum;
	std::vector<mb128> m;
	const int b = ix->b;
	for (uint64_t bi = 0; bi < ((uint64_t)1 << b); ++bi) {
		int32_t n; uint32_t size;
		if (fread(&n, 4, 1, fp) != 1 || n < 0) fail("damaged index (bucket)");
		std::vector<uint64_t> p(n);
		if (n && fread(p.data(), 8, n, fp) != (size_t)n) fail("damaged index (bucket)");
		if (fread(&size, 4, 1, fp) != 1) fail("damaged index (bucket)");
		for (uint32_t s = 0; s < size; ++s) {
			uint64_t kv[2];
			if (fread(kv, 8, 2, fp) != 2) fail("damaged index (hash)");
			uint64_t h = (kv[0] >> 1) << b | bi;
			if (kv[0] & 1) m.push_back(mb128{ h << 8 | (uint64_t)ix->k, kv[1] });
			else {
				uint64_t st = kv[1] >> 32; uint32_t cnt = (uint32_t)kv[1];
				if (st + cnt > (uint64_t)n) fail("damaged index (offsets)");
				for (uint32_t c2 = 0; c2 < cnt; ++c2) m.push_back(mb128{ h << 8 | (uint64_t)ix->k, p[st + c2] });
			}
		}
	}
	ix->h_S.assign((sum + 7) / 8, 0);
	if (!ix->h_S.empty() && fread(ix->h_S.data(), 4, ix->h_S.size(), fp) != ix->h_S.size()) fail("damaged index (sequence)");
	{ // same device pipeline as a fresh build: the positions of one hash come ascending out of the .mmi, the sort is stable
		ThreadCtx &c = get_ctx(device);
		c.ar.reset();
		mb128 *d_m = c.ar.get<mb128>(m.size() + 1);
		uint32_t *d_S = c.ar.get<uint32_t>(ix->h_S.size() + 2);
		if (!m.empty()) CK(cudaMemcpyAsync(d_m, m.data(), m.size() * sizeof(mb128), cudaMemcpyHostToDevice, c.st));
		if (!ix->h_S.empty()) CK(cudaMemcpyAsync(d_S, ix->h_S.data(), ix->h_S.size() * 4, cudaMemcpyHostToDevice, c.st));
		index_finish_device(ix.get(), c, d_m, (int64_t)m.size(), d_S, ix->k);
		ix->h_S.clear(); ix->h_S.shrink_to_fit();
		c.ar.reset();
	}
	*out = ix.release();
	API_END
}

extern "C" void mb_index_free(mb_index_t *ix)
{
	if (!ix) return;
	cudaSetDevice(ix->device);
	DevIndex &d = ix->d;
	cudaFree(d.hkey); cudaFree(d.hval); cudaFree(d.pos); cudaFree(d.S); cudaFree(d.seq_off); cudaFree(d.seq_len);
	delete ix;
}
extern "C" int mb_index_n_seq(const mb_index_t *ix) { return ix ? (int)ix->names.size() : 0; }
extern "C" const char *mb_index_seq_name(const mb_index_t *ix, int rid) { return (ix && rid >= 0 && rid < (int)ix->names.size()) ? ix->names[rid].c_str() : nullptr; }
extern "C" int64_t mb_index_seq_len(const mb_index_t *ix, int rid) { return (ix && rid >= 0 && rid < (int)ix->lens.size()) ? (int64_t)ix->lens[rid] : (int64_t)-1; }
extern "C" int mb_index_mid_occ(const mb_index_t *ix) { return ix ? ix->mid_occ : 0; }
extern "C" int mb_index_kw(const mb_index_t *ix, int *k, int *w) { if (!ix) return MB_ERR_ARG; if (k) *k = ix->k; if (w) *w = ix->w; return MB_OK; }
extern "C" int64_t mb_index_n_minimizers(const mb_index_t *ix) { return ix ? ix->n_mini : 0; }
extern "C" int64_t mb_index_hbm_bytes(const mb_index_t *ix) { return ix ? ix->hbm_bytes : 0; }

// ---------------------------------------------------------------------------------------------
// reads / hits containers
// ---------------------------------------------------------------------------------------------
struct mb_reads {
	int device = 0; int32_t n_reads = 0; int64_t total = 0;
	uint8_t *d_codes = nullptr; int64_t *d_off = nullptr;
};

// Result arrays live in PINNED host memory so the device->host copies of a batch (hundreds of MB of CIGARs at the bench
// size) run at PCIe speed; the buffers come from a process-wide pool and go back to it in mb_hits_free, so a steady stream
// of batches allocates nothing.
struct PinPool {
	std::mutex m;
	std::vector<std::pair<void*, size_t>> free_list;
	void *get(size_t bytes, size_t *cap) {
		if (bytes == 0) bytes = 1;
		{
			std::lock_guard<std::mutex> g(m);
			int best = -1;
			for (int i = 0; i < (int)free_list.size(); ++i)
				if (free_list[i].second >= bytes && (best < 0 || free_list[i].second < free_list[best].second)) best = i;
			if (best >= 0) { void *p = free_list[best].first; *cap = free_list[best].second; free_list.erase(free_list.begin() + best); return p; }
		}
		size_t c = (bytes + (bytes >> 3) + 4095) & ~(size_t)4095; // a little headroom: the next batch is rarely identical
		void *p = nullptr;
		if (cudaHostAlloc(&p, c, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); throw mb_error(MB_ERR_NOMEM, "cudaHostAlloc failed for " + std::to_string(c) + " bytes"); }
		*cap = c;
		return p;
	}
	void put(void *p, size_t cap) {
		if (!p) return;
		std::lock_guard<std::mutex> g(m);
		if (free_list.size() >= 24) { // keep the pool bounded: drop the smallest buffer
			int small = 0;
			for (int i = 1; i < (int)free_list.size(); ++i) if (free_list[i].second < free_list[small].second) small = i;
			if (free_list[small].second < cap) { cudaFreeHost(free_list[small].first); free_list[small] = std::make_pair(p, cap); }
			else cudaFreeHost(p);
			return;
		}
		free_list.emplace_back(p, cap);
	}
};
static PinPool g_pin_pool;

template <typename T> struct PinVec {
	T *p = nullptr; size_t n = 0, cap = 0;
	bool plain = false;   // pageable memory (results assembled on the host from several pieces: never a DMA target)
	PinVec() {}
	PinVec(const PinVec&) = delete; PinVec &operator=(const PinVec&) = delete;
	~PinVec() { if (plain) free(p); else g_pin_pool.put(p, cap); }
	void resize(size_t k) { // contents are NOT preserved or initialised
		if (plain) { if (k * sizeof(T) > cap) { free(p); cap = k * sizeof(T) + 64; p = (T*)malloc(cap); if (!p) throw mb_error(MB_ERR_NOMEM, "out of host memory"); } n = k; return; }
		if (k * sizeof(T) > cap) { g_pin_pool.put(p, cap); p = nullptr; cap = 0; p = (T*)g_pin_pool.get(k * sizeof(T), &cap); }
		n = k;
	}
	void append_plain(const T *src, size_t k) { // plain mode only: grow geometrically, keep the contents
		if ((n + k) * sizeof(T) > cap) {
			size_t nc = std::max((n + k) * sizeof(T), cap + cap / 2) + 64;
			T *np = (T*)realloc(p, nc);
			if (!np) throw mb_error(MB_ERR_NOMEM, "out of host memory");
			p = np, cap = nc;
		}
		if (k) memcpy(p + n, src, k * sizeof(T));
		n += k;
	}
	void assign(size_t k, T v) { resize(k); for (size_t i = 0; i < k; ++i) p[i] = v; }
	T *data() { return p; } const T *data() const { return p; }
	size_t size() const { return n; }
};

struct mb_hits {
	int64_t n = 0; int32_t n_reads = 0;
	PinVec<int32_t> fields;             // HIT_NF * n
	PinVec<int64_t> cigar_off;
	PinVec<uint32_t> cigar;
	PinVec<int32_t> rep_len;
	PinVec<int64_t> hit_off;            // [n_reads+1]
	std::vector<int64_t> read_off;      // [n_reads+1] (for mb_count's query_length mode)
};

static const char *HIT_NAMES[HIT_NF] = { "read_idx", "rid", "rev", "qs", "qe", "rs", "re", "mapq", "mlen", "blen", "nm", "dp_max", "dp_max2",
	"score", "score0", "cnt", "subsc", "n_sub", "id", "parent", "is_primary", "sam_pri", "n_cigar" };

extern "C" int64_t mb_hits_n(const mb_hits_t *h) { return h ? h->n : 0; }
extern "C" const int32_t *mb_hits_field(const mb_hits_t *h, const char *name)
{
	if (!h || !name) return nullptr;
	for (int f = 0; f < HIT_NF; ++f) if (strcmp(name, HIT_NAMES[f]) == 0) return h->fields.data() + (size_t)f * h->n;
	return nullptr;
}
extern "C" const int64_t *mb_hits_cigar_off(const mb_hits_t *h) { return h ? h->cigar_off.data() : nullptr; }
extern "C" const uint32_t *mb_hits_cigar_pool(const mb_hits_t *h, int64_t *n) { if (!h) return nullptr; if (n) *n = (int64_t)h->cigar.size(); return h->cigar.data(); }
extern "C" const int32_t *mb_hits_rep_len(const mb_hits_t *h, int64_t *n_reads) { if (!h) return nullptr; if (n_reads) *n_reads = h->n_reads; return h->rep_len.data(); }
extern "C" void mb_hits_free(mb_hits_t *h) { delete h; }

// ---------------------------------------------------------------------------------------------
// the pipeline
// ---------------------------------------------------------------------------------------------
struct Timer {
	cudaEvent_t e[2]; cudaStream_t st;
	Timer(cudaStream_t s) : st(s) { cudaEventCreate(&e[0]); cudaEventCreate(&e[1]); }
	~Timer() { cudaEventDestroy(e[0]); cudaEventDestroy(e[1]); }
	void start() { cudaEventRecord(e[0], st); }
	float stop() { cudaEventRecord(e[1], st); cudaEventSynchronize(e[1]); float ms = 0; cudaEventElapsedTime(&ms, e[0], e[1]); return ms; }
};


static DpScoring make_scoring(const mb_opt_t &o)
{
	DpScoring s;
	s.q = (int8_t)o.q, s.e = (int8_t)o.e, s.q2 = (int8_t)o.q2, s.e2 = (int8_t)o.e2;
	s.sc_mch = (int8_t)(o.a < 0 ? -o.a : o.a);
	s.sc_mis = (int8_t)(o.b > 0 ? -o.b : o.b);
	int amb = o.sc_ambi > 0 ? -o.sc_ambi : o.sc_ambi;
	s.sc_N = (int8_t)(amb == 0 ? -o.e2 : amb);
	s.pad = 0;
	return s;
}

// scratch geometry of one DP task
struct DpGeom { size_t p_bytes, ws_bytes, h_ints; };
static __host__ __device__ inline DpGeom dp_geom(int qlen, int tlen, int w)
{
	DpGeom g; g.p_bytes = g.ws_bytes = g.h_ints = 0;
	if (qlen <= 0 || tlen <= 0) return g;
	if (w < 0) w = tlen > qlen ? tlen : qlen;
	int tlen_ = (tlen + 15) / 16, qlen_ = (qlen + 15) / 16;
	int n_col_ = qlen < tlen ? qlen : tlen;
	n_col_ = ((n_col_ < w + 1 ? n_col_ : w + 1) + 15) / 16 + 1;
	g.p_bytes = ((size_t)(qlen + tlen - 1) * n_col_ + 1) * 16;
	g.ws_bytes = (size_t)tlen_ * 16 * 8 + (size_t)qlen_ * 16 + 16;
	g.h_ints = (size_t)tlen_ * 16;
	return g;
}

#define DP_NEXACT 6                 // exact-kernel classes by direction-matrix size: <=64K, <=256K, <=1M, <=4M, <=16M, larger
#define DP_XBASE DPF_NCLASS          // extension fast-path classes follow the gap-fill fast-path classes
#define DP_BBASE (2 * DPF_NCLASS)   // then the packed band kernel's classes (large / band-limited gap fills)
#define DP_EBASE (2 * DPF_NCLASS + DPB_NCLASS)   // then the exact-kernel classes (one warp per task)
#define DP_NCTA 4                    // and the CTA-per-task exact kernel for the large ones (exact classes 2..5)
#define DP_CBASE (DP_EBASE + DP_NEXACT)
#define DP_DBASE (DP_CBASE + DP_NCTA)   // the same for end extensions whose results are not needed before the stitch (deferred launches)
#define DP_NCLS (DP_DBASE + DP_NCTA)
static __host__ __device__ inline int dp_exact_class(size_t p_bytes)
{
	int k = 0;
	for (size_t lim = (size_t)64 << 10; k < DP_NEXACT - 1 && p_bytes > lim; lim <<= 2) ++k;
	return k;
}

// classify tasks: fast path by columns-per-lane class, the rest into small / big scratch classes; record maxima
__global__ void k_dp_classify(const DpTask *__restrict__ tasks, const int32_t *__restrict__ ids, int64_t n, int use_ids,
                              const uint8_t *__restrict__ codes, const uint32_t *__restrict__ S, const uint8_t *__restrict__ pool, int fast_ok, int defer_ext,
                              int32_t *__restrict__ lists /* DP_NCLS x n */, int32_t *__restrict__ ctr /* DP_NCLS */,
                              unsigned long long *__restrict__ maxima /* DP_NCLS x 3 */)
{
	const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	const int lane = threadIdx.x & 31;
	const bool live = i < n;
	int id = 0, cls = -1;
	DpTask t;
	if (live) {
		id = use_ids ? ids[i] : (int)i;
		t = tasks[id];
		cls = fast_ok ? dpf_class(t.qlen, t.tlen, t.w, t.flag, t.skip) : -1;
		if (cls < 0 && fast_ok) { cls = dpx_class(t.qlen, t.tlen, t.w, t.flag, t.skip); if (cls >= 0) cls += DP_XBASE; }
		if (cls < 0 && fast_ok) { cls = dpb_class(t.qlen, t.tlen, t.w, t.flag, t.skip); if (cls >= 0) cls += DP_BBASE; }
	}
	// ambiguous bases send a task to the exact kernel.  The scan of a task's two windows is done by the whole warp, one task
	// after the other: 32 lanes reading consecutive words of one window instead of 32 lanes each walking its own
	const unsigned need = __ballot_sync(0xffffffffu, live && cls >= 0);
	for (unsigned m = need; m; m &= m - 1) {
		const int src = __ffs(m) - 1;
		const int q_comp = __shfl_sync(0xffffffffu, (int)t.q_comp, src), q_step = __shfl_sync(0xffffffffu, (int)t.q_step, src);
		const int t_step = __shfl_sync(0xffffffffu, (int)t.t_step, src), t_packed = __shfl_sync(0xffffffffu, (int)t.t_packed, src);
		const int qlen = __shfl_sync(0xffffffffu, t.qlen, src), tlen = __shfl_sync(0xffffffffu, t.tlen, src);
		const int64_t q_idx0 = __shfl_sync(0xffffffffu, (long long)t.q_idx0, src), t_idx0 = __shfl_sync(0xffffffffu, (long long)t.t_idx0, src);
		const uint8_t *qc = q_comp == 2 ? pool : codes;
		const int64_t qlo = q_step > 0 ? q_idx0 : q_idx0 - (qlen - 1), tlo = t_step > 0 ? t_idx0 : t_idx0 - (tlen - 1);
		bool amb = dpf_bytes_ambig_warp(qc, qlo, qlen, lane);
		if (!amb) amb = t_packed ? dpf_nibbles_ambig_warp(S, tlo, tlen, lane) : dpf_bytes_ambig_warp(pool, tlo, tlen, lane);
		if (amb && lane == src) cls = -1;
	}
	// class counters and maxima are a handful of hot addresses: one atomic per (warp, class) instead of one per task
	unsigned m0 = 0, m1 = 0, m2 = 0;
	if (live) {
		if (cls >= 0) { m0 = (unsigned)t.qlen; if (cls >= DP_BBASE) m1 = (unsigned)t.tlen; }
		else {
			DpGeom g = dp_geom(t.qlen, t.tlen, t.w);
			if (t.skip) g.p_bytes = g.ws_bytes = g.h_ints = 0;
			const int ec = dp_exact_class(g.p_bytes);
			int wd = t.w < 0 ? (t.tlen > t.qlen ? t.tlen : t.qlen) : t.w;
			int dw = t.qlen < t.tlen ? t.qlen : t.tlen; dw = dw < wd + 1 ? dw : wd + 1;      // widest diagonal
			cls = (ec >= 2 && dw + 80 <= DPC2_THREADS * 4) ? ((defer_ext && t.kind != 1) ? DP_DBASE : DP_CBASE) + (ec - 2) : DP_EBASE + ec;
			m0 = (unsigned)g.p_bytes, m1 = (unsigned)g.ws_bytes, m2 = (unsigned)g.h_ints;
		}
	}
	const unsigned peers = __match_any_sync(0xffffffffu, live ? cls : -1);
	if (!live) return;
	// rank among the warp's tasks of this class by query length: the DP kernels pair neighbouring list entries in one warp,
	// and a pair runs for the longer of its two queries
	const int leader = __ffs(peers) - 1;
	int rank = 0;
	for (unsigned m = peers; m; m &= m - 1) {
		const int src = __ffs(m) - 1;
		const int q = __shfl_sync(peers, t.qlen, src);
		rank += (q < t.qlen || (q == t.qlen && src < lane)) ? 1 : 0;
	}
	const unsigned x0 = __reduce_max_sync(peers, m0), x1 = __reduce_max_sync(peers, m1), x2 = __reduce_max_sync(peers, m2);
	int base = 0;
	if (lane == leader) {
		base = atomicAdd(&ctr[cls], __popc(peers));
		atomicMax(&maxima[cls * 3 + 0], (unsigned long long)x0);
		if (cls >= DP_BBASE) { atomicMax(&maxima[cls * 3 + 1], (unsigned long long)x1); atomicMax(&maxima[cls * 3 + 2], (unsigned long long)x2); }
	}
	base = __shfl_sync(peers, base, leader);
	lists[(int64_t)cls * n + base + rank] = id;
}

// Longest task first for the classes whose launches are a few long tasks per CTA (band, CTA exact): the launch then ends
// with short tasks instead of a straggler, and neighbouring list entries (the band kernel pairs them in one warp) are alike.
#define DPL_MAX 4096
__global__ void __launch_bounds__(1024)
k_list_longest_first(const DpTask *__restrict__ tasks, const int32_t *__restrict__ list, int n, int32_t *__restrict__ out)
{
	__shared__ unsigned long long s_key[DPL_MAX];
	for (int i = threadIdx.x; i < n; i += 1024) { const DpTask &t = tasks[list[i]]; s_key[i] = (unsigned long long)((unsigned)(t.qlen + t.tlen)) << 32 | (unsigned)(n - 1 - i); }
	__syncthreads();
	for (int i = threadIdx.x; i < n; i += 1024) {
		const unsigned long long k = s_key[i];
		int rank = 0;
		for (int j = 0; j < n; ++j) rank += s_key[j] > k;
		out[rank] = list[i];
	}
}

// grids of the side-stream kernels (exact, band, extension) relative to their defaults (MB_SIDE_SCALE; tuning knob)
static double mb_side_scale() { static double v = -1; if (v < 0) { const char *e = getenv("MB_SIDE_SCALE"); v = e ? atof(e) : 1.0; if (v <= 0) v = 1.0; } return v; }
static int mb_side_grid(int n) { const int v = (int)(n * mb_side_scale()); return v < 8 ? (n < 8 ? n : 8) : v; }

struct DpRunner {
	ThreadCtx &c; cudaStream_t st; int64_t *nl;
	std::vector<std::pair<cudaEvent_t, cudaEvent_t>> evs; // one pair per DP launch, read back after the batch
	cudaEvent_t ev_base = nullptr;
	std::vector<std::pair<cudaEvent_t, cudaEvent_t>> fast_wall; // fork -> all k_dp_fast launches of a run done (they overlap each other)
	std::vector<int> ev_fast; int64_t n_fast = 0, n_exact = 0, n_ext = 0, n_band = 0; // ev_fast: 1 k_dp_fast, 2 k_dp_ext, 3 k_dp_band, 0 k_dp
	DpRunner(ThreadCtx &c_, int64_t *nl_) : c(c_), st(c_.st), nl(nl_) {}
	~DpRunner() { for (auto &e : evs) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); } for (auto &e : fast_wall) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); } }
	float fast_wall_ms() { float t = 0; for (auto &e : fast_wall) { float ms = 0; if (cudaEventElapsedTime(&ms, e.first, e.second) == cudaSuccess) t += ms; } return t; }
	float total_ms(int which = -1) { float t = 0; for (size_t i = 0; i < evs.size(); ++i) { if (which >= 0 && ev_fast[i] != which) continue; float ms = 0; if (cudaEventElapsedTime(&ms, evs[i].first, evs[i].second) == cudaSuccess) t += ms; } return t; }

	template <int C>
	void launch_fast(DpTask *tasks, const int32_t *list, const int32_t *d_cnt, int64_t cnt, int max_q, const uint8_t *codes, const uint32_t *S, const uint8_t *pool,
	                 uint32_t *cigar_pool, const DpScoring &sc, unsigned long long *d_cells, cudaStream_t st)
	{
		Arena &ar = c.ar;
		constexpr int CW = (C + 1) / 2;
		static int occ = 0; // resident CTAs per SM for this instantiation
		if (occ == 0) {
			CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_dp_fast<C>, DPF_WARPS * 32, 0));
			if (occ < 1) occ = 1;
		}
		const size_t stride_words = ((size_t)32 * (size_t)(max_q + 31) * CW + 63) & ~(size_t)63;
		// leave one CTA slot per SM (registers) to the side-stream kernels (extensions, exact DP) so they co-run instead of
		// queueing behind a full-occupancy grid
		int max_cta = c.num_sms * occ;
		int64_t want = cdiv(cdiv(cnt, 2), DPF_WARPS); // two tasks per warp
		// Experiment, off by default (MB_FAST_CHAIN=G turns it on): the pairs of a warp chained through the lanes in groups of G
		// (dp_fast_chain.cuh), which pays the 31-step ramp of the systolic wavefront once per group.  G shrinks with the pairs a
		// warp gets (a group is the unit of work stealing) and with the direction-byte scratch (G regions per warp, at most 2 GB
		// per launch).  Bit-identical results; measured on configs[1]: 149.5 ms of gap-fill launches at G = 4 (151.8 at G = 2)
		// against 150.6 - 151.1 ms unchained.  A lane's switch to the next pair is executed by one lane at a time, 32 times per
		// boundary (66 instructions each: 3.4 % of a pair), and the backtracks of a group read direction bytes that have left the
		// L2 by then, so most of what the removed ramps save is spent again; 1 % does not pay for 4x the direction-byte scratch
		// (profiles/r02_summary.md).
		const char *chain_env = getenv("MB_FAST_CHAIN");
		const int chain_g = chain_env ? atoi(chain_env) : 0;
		if (chain_g >= 2) {
			static int occ_c = 0;
			if (occ_c == 0) { CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_c, k_dp_fast_chain<C>, 32, 0)); if (occ_c < 1) occ_c = 1; }
			const int max_cta_c = c.num_sms * occ_c;
			int G = chain_g < DPF_GMAX ? chain_g : DPF_GMAX;
			const int64_t per_warp = want / max_cta_c;
			const bool chain_force = getenv("MB_FAST_CHAIN_FORCE") != nullptr;   // tests: chain small classes too
			if (per_warp / 4 < G && !chain_force) G = (int)(per_warp / 4);
			while (G >= 2 && (size_t)max_cta_c * stride_words * 4 * G > ((size_t)2 << 30)) --G;
			if (G >= 2) {
				const int64_t want_c = cdiv(want, G);
				const int n_cta = (int)(want_c < max_cta_c ? want_c : max_cta_c);
				uint32_t *p_scr = ar.get<uint32_t>((size_t)n_cta * stride_words * G);
				int32_t *wc = ar.get<int32_t>(1);
				CK(cudaMemsetAsync(wc, 0, sizeof(int32_t), st));
				cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
				cudaEventRecord(e0, st);
				k_dp_fast_chain<C><<<n_cta, 32, 0, st>>>(tasks, list, d_cnt, wc, codes, S, pool, p_scr, stride_words * G, cigar_pool, sc, d_cells, G);
				cudaEventRecord(e1, st);
				evs.emplace_back(e0, e1); ev_fast.push_back(1); n_fast += cnt;
				++*nl;
				return;
			}
		}
		int n_cta = (int)(want < max_cta ? want : max_cta);
		uint32_t *p_scr = ar.get<uint32_t>((size_t)n_cta * DPF_WARPS * stride_words);
		int32_t *wc = ar.get<int32_t>(1);
		CK(cudaMemsetAsync(wc, 0, sizeof(int32_t), st));
		cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
		cudaEventRecord(e0, st);
		k_dp_fast<C><<<n_cta, DPF_WARPS * 32, 0, st>>>(tasks, list, d_cnt, wc, codes, S, pool, p_scr, stride_words, cigar_pool, sc, d_cells);
		cudaEventRecord(e1, st);
		evs.emplace_back(e0, e1); ev_fast.push_back(1); n_fast += cnt;
		++*nl;
	}

	template <int C>
	void launch_ext(DpTask *tasks, const int32_t *list, const int32_t *d_cnt, int64_t cnt, int max_q, const uint8_t *codes, const uint32_t *S, const uint8_t *pool,
	                uint32_t *cigar_pool, const DpScoring &sc, unsigned long long *d_cells, cudaStream_t st)
	{
		Arena &ar = c.ar;
		constexpr int CW = (C + 1) / 2;
		static int occ = 0;
		if (occ == 0) {
			CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_dp_ext<C>, DPX_WARPS * 32, 0));
			if (occ < 1) occ = 1;
		}
		const size_t stride_words = ((size_t)32 * (size_t)(max_q + 31) * CW + 63) & ~(size_t)63;
		int max_cta = c.num_sms * occ;   // run at full occupancy and leave: the gap-fill launches wait for these
		int64_t want = cdiv(cdiv(cnt, 2), DPX_WARPS);
		int n_cta = (int)(want < max_cta ? want : max_cta);
		uint32_t *p_scr = ar.get<uint32_t>((size_t)n_cta * DPX_WARPS * stride_words);
		int32_t *wc = ar.get<int32_t>(1);
		CK(cudaMemsetAsync(wc, 0, sizeof(int32_t), st));
		cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
		cudaEventRecord(e0, st);
		k_dp_ext<C><<<n_cta, DPX_WARPS * 32, 0, st>>>(tasks, list, d_cnt, wc, codes, S, pool, p_scr, stride_words, cigar_pool, sc, d_cells);
		cudaEventRecord(e1, st);
		evs.emplace_back(e0, e1); ev_fast.push_back(2); n_ext += cnt;
		++*nl;
	}

	// run the DP kernels over `n` tasks (ids[] if use_ids else 0..n-1)
	// defer_ext: the large exact END EXTENSIONS (kind != 1; only the stitch reads their results) are launched on c.st_defer
	// behind the gap-fill kernels and NOT joined: the caller waits for c.ev_defer (join_deferred) before it reads them.
	int deferred = 0;   // bit i: c.st_defer[i] carries a deferred launch
	void join_deferred() { for (int i = 0; i < 3; ++i) if (deferred >> i & 1) CK(cudaStreamWaitEvent(st, c.ev_defer[i], 0)); deferred = 0; }
	void run(DpTask *tasks, const int32_t *ids, int64_t n, bool use_ids, const uint8_t *codes, const uint32_t *S, const uint8_t *pool,
	         uint32_t *cigar_pool, const DpScoring &sc, unsigned long long *d_cells, bool defer_ext = false)
	{
		if (n <= 0) return;
		Arena &ar = c.ar;
		int32_t *lists = ar.get<int32_t>((size_t)DP_NCLS * n);
		int32_t *ctr = ar.get<int32_t>(DP_NCLS);
		unsigned long long *maxima = ar.get<unsigned long long>(DP_NCLS * 3);
		CK(cudaMemsetAsync(ctr, 0, DP_NCLS * sizeof(int32_t), st));
		CK(cudaMemsetAsync(maxima, 0, DP_NCLS * 3 * sizeof(unsigned long long), st));
		k_dp_classify<<<(unsigned)cdiv(n, 256), 256, 0, st>>>(tasks, ids, n, use_ids ? 1 : 0, codes, S, pool, dpf_scoring_ok(sc) ? 1 : 0, defer_ext ? 1 : 0, lists, ctr, maxima); ++*nl;
		int32_t h_ctr[DP_NCLS]; unsigned long long h_max[DP_NCLS * 3];
		CK(cudaMemcpyAsync(h_ctr, ctr, sizeof(h_ctr), cudaMemcpyDeviceToHost, st));
		CK(cudaMemcpyAsync(h_max, maxima, sizeof(h_max), cudaMemcpyDeviceToHost, st));
		CK(cudaStreamSynchronize(st));
		if (getenv("MB_DEBUG")) { auto t1 = std::chrono::steady_clock::now(); fprintf(stderr, "[mb]   classify done (+%.3f ms)\n", std::chrono::duration<double, std::milli>(t1 - g_dbg_t0).count()); g_dbg_t0 = t1; }
		if (getenv("MB_DEBUG") && !ev_base) { cudaEventCreate(&ev_base); cudaEventRecord(ev_base, st); }
		if (getenv("MB_DEBUG_TASKS")) { // which tasks miss the packed kernels, and why (host-side histogram of the task list)
			int64_t n_all = n;
			if (use_ids) { int64_t mx = 0; std::vector<int32_t> hi(n); cudaMemcpy(hi.data(), ids, n * 4, cudaMemcpyDeviceToHost); for (int32_t v : hi) mx = std::max<int64_t>(mx, v + 1); n_all = mx; }
			std::vector<DpTask> ht(n_all);
			cudaMemcpy(ht.data(), tasks, n_all * sizeof(DpTask), cudaMemcpyDeviceToHost);
			std::vector<int32_t> hid(n);
			if (use_ids) cudaMemcpy(hid.data(), ids, n * 4, cudaMemcpyDeviceToHost); else for (int64_t i = 0; i < n; ++i) hid[i] = (int32_t)i;
			struct Cat { int64_t n = 0; double cells = 0; };
			std::map<std::string, Cat> cats;
			for (int64_t i = 0; i < n; ++i) {
				const DpTask &t = ht[hid[i]];
				if (t.skip || t.qlen <= 0 || t.tlen <= 0) continue;
				if (dpf_class(t.qlen, t.tlen, t.w, t.flag, t.skip) >= 0 || dpx_class(t.qlen, t.tlen, t.w, t.flag, t.skip) >= 0) continue;
				int w = t.w < 0 ? std::max(t.qlen, t.tlen) : t.w;
				const int mx = std::max(t.qlen, t.tlen);
				std::string k = t.flag == 0 ? "pass2" : (t.flag & MB_EZ_EXTZ_ONLY) ? "ext" : "fill";
				k += mx > w + 1 ? "/band-limited" : "/band-free";
				k += mx <= 768 ? "/<=768" : mx <= 1536 ? "/<=1536" : mx <= 3072 ? "/<=3072" : "/>3072";
				Cat &c2 = cats[k]; ++c2.n;
				double cells = 0; for (int r = 0; r < t.qlen + t.tlen - 1; ++r) { int st0 = std::max(std::max(0, r - t.qlen + 1), (r - w + 1) >> 1), en0 = std::min(std::min(t.tlen - 1, r), (r + w) >> 1); if (en0 >= st0) cells += en0 - st0 + 1; }
				c2.cells += cells;
			}
			for (auto &kv : cats) fprintf(stderr, "[mb-tasks] %-36s %8lld tasks %10.3f Gcells (full band area)\n", kv.first.c_str(), (long long)kv.second.n, kv.second.cells / 1e9);
		}
		// the exact kernel goes first, on the high-priority side stream (the host synchronised `st` above, so its inputs are
		// complete), largest class first: a handful of long band-limited tasks give each launch a long tail of a few busy
		// warps, which the fast kernels on `st` fill
		auto longest_first = [&](int cls, cudaStream_t s2) -> const int32_t* { // sorted copy of a class list (small classes only)
			const int64_t cnt = h_ctr[cls];
			const int32_t *list = lists + (int64_t)cls * n;
			if (cnt < 3 || cnt > DPL_MAX) return list;
			int32_t *out = ar.get<int32_t>(cnt);
			k_list_longest_first<<<1, 1024, 0, s2>>>(tasks, list, (int)cnt, out); ++*nl;
			return out;
		};
		// The exact kernels start AFTER the band and extension kernels are done: their CTAs hold most of an SM's shared memory for
		// tens of milliseconds (long dependency chains, few tasks) and would keep the short extension / band CTAs from becoming
		// resident.  They run beside the gap-fill launches.
		auto launch_cta = [&](int cls, cudaStream_t st2, int cap_per_sm = 0) { // the large exact tasks: one CTA per task (align_cta.cuh)
			const int64_t cnt = h_ctr[cls];
			if (cnt == 0) return false;
			size_t p_stride = ((size_t)h_max[cls * 3 + 0] + 255) & ~(size_t)255;
			const size_t g_stride = ((size_t)h_max[cls * 3 + 1] + 255) & ~(size_t)255;
			size_t h_stride = ((size_t)h_max[cls * 3 + 2] + 63) & ~(size_t)63;
			if (p_stride == 0) p_stride = 256;
			if (h_stride == 0) h_stride = 64;
			static const bool cta_old = getenv("MB_CTA_OLD") != nullptr;   // debug: one cell at a time in 32-bit registers
			static const bool nowin = getenv("MB_CTA_NOWIN") != nullptr;   // debug: state arrays of full length (shared memory if they fit, else global)
			const size_t need = ((g_stride + 15) & ~(size_t)15) + h_stride * 4 + 64;
			const int smem = nowin ? (int)(need < DPC_SMEM_MAX ? need : DPC_SMEM_MAX) : DPC_WIN_SMEM;
			int per_sm = (200 * 1024) / (smem > 4096 ? smem : 4096); if (per_sm < 1) per_sm = 1; if (per_sm > (nowin ? 4 : 7)) per_sm = nowin ? 4 : 7;
			if (!nowin && !cta_old) { // persistent CTAs: exactly as many as are resident at once (registers), so that none queues ahead of other streams' kernels
				static int occ2 = 0;
				if (occ2 == 0) { CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ2, k_dp_cta2, DPC2_THREADS, DPC_WIN_SMEM)); if (occ2 < 1) occ2 = 1; }
				if (per_sm > occ2) per_sm = occ2;
			}
			if (cap_per_sm > 0 && per_sm > cap_per_sm) per_sm = cap_per_sm;
			int n_cta = (int)std::min<int64_t>(cnt, (int64_t)mb_side_grid(c.num_sms * per_sm));
			const size_t per_cta = p_stride + (nowin ? g_stride + h_stride * 4 : 0);
			while (n_cta > 1 && (size_t)n_cta * per_cta > ((size_t)24 << 30)) n_cta = (n_cta + 1) / 2;
			uint8_t *p_scr = ar.get<uint8_t>((size_t)n_cta * p_stride);
			int8_t *g_ws = ar.get<int8_t>(nowin ? (size_t)n_cta * g_stride + 16 : 16);     // full-length state arrays: only without the window
			int32_t *h_scr = ar.get<int32_t>(nowin ? (size_t)n_cta * h_stride : 16);
			int32_t *wc = ar.get<int32_t>(1);
			CK(cudaMemsetAsync(wc, 0, sizeof(int32_t), st2));
			cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
			cudaEventRecord(e0, st2);
			const int32_t *clist = longest_first(cls, st2);
			if (nowin) k_dp_cta<false><<<n_cta, DPC_THREADS, smem, st2>>>(tasks, clist, ctr + cls, wc, codes, S, pool,
				p_scr, p_stride, g_ws, g_stride, h_scr, h_stride, cigar_pool, sc, d_cells ? d_cells + 1 : nullptr, smem);
			else if (cta_old) k_dp_cta<true><<<n_cta, DPC_THREADS, smem, st2>>>(tasks, clist, ctr + cls, wc, codes, S, pool,
				p_scr, p_stride, g_ws, g_stride, h_scr, h_stride, cigar_pool, sc, d_cells ? d_cells + 1 : nullptr, smem);
			else k_dp_cta2<<<n_cta, DPC2_THREADS, smem, st2>>>(tasks, clist, ctr + cls, wc, codes, S, pool,
				p_scr, p_stride, g_ws, g_stride, h_scr, h_stride, cigar_pool, sc, d_cells ? d_cells + 1 : nullptr, smem);
			{ cudaError_t le = cudaGetLastError(); if (le != cudaSuccess) throw mb_error(MB_ERR_CUDA, std::string("k_dp_cta launch: ") + cudaGetErrorString(le) + " grid " + std::to_string(n_cta) + " smem " + std::to_string(smem)); }
			cudaEventRecord(e1, st2);
			evs.emplace_back(e0, e1); ev_fast.push_back(0); n_exact += cnt;
			++*nl;
			return true;
		};
		bool side[MB_NSIDE] = {};
		bool band_fills = false;   // a band launch with a full grid: worth waiting for (else it is a few long tasks: run beside it)
		const bool serial = getenv("MB_DEBUG_SERIAL") != nullptr;   // every launch on `st`, one after the other: stand-alone durations
		for (int k = DPB_NCLASS - 1; k >= 0; --k) { // large / band-limited gap fills: packed systolic kernel with upstream's band (dp_band.cuh)
			const int cls = DP_BBASE + k;
			const int64_t cnt = h_ctr[cls];
			if (cnt == 0) continue;
			const int mq = (int)h_max[cls * 3], mt = (int)h_max[cls * 3 + 1];
			const int n_strips = (mt + DPB_STRIP - 1) / DPB_STRIP;
			const size_t stride_words = ((size_t)DPB_MAX_STRIPS * DPB_EDGE_WORDS + (size_t)n_strips * (size_t)(mq + 31) * 32 * DPB_CW + 63) & ~(size_t)63;
			static const int band_per_sm = getenv("MB_BAND_PER_SM") ? atoi(getenv("MB_BAND_PER_SM")) : 3;
			int max_cta = c.num_sms * (band_per_sm >= 1 && band_per_sm <= 3 ? band_per_sm : 3);   // 3 = full occupancy: the gap-fill launches wait for the band launches
			const int64_t want = cdiv(cnt, 2);
			const int n_cta = (int)(want < max_cta ? want : max_cta);
			uint32_t *p_scr = ar.get<uint32_t>((size_t)n_cta * stride_words);
			int32_t *wc = ar.get<int32_t>(1);
			cudaStream_t sb = serial ? st : (k & 1) ? c.st2[0] : c.st2[2]; side[(k & 1) ? 0 : 2] = true;   // the two classes side by side
			const int32_t *blist = longest_first(cls, sb);
			static const bool band_nowait = getenv("MB_BAND_NOWAIT") != nullptr;
			if (want >= max_cta && band_per_sm >= 3 && !band_nowait) band_fills = true;
			CK(cudaMemsetAsync(wc, 0, sizeof(int32_t), sb));
			cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
			cudaEventRecord(e0, sb);
			k_dp_band<<<n_cta, DPB_NW * 32, 0, sb>>>(tasks, blist, ctr + cls, wc, codes, S, pool, p_scr, stride_words, cigar_pool, sc, d_cells ? d_cells + 3 : nullptr);
			{ cudaError_t le = cudaGetLastError(); if (le != cudaSuccess) throw mb_error(MB_ERR_CUDA, std::string("k_dp_band launch: ") + cudaGetErrorString(le) + " grid " + std::to_string(n_cta)); }
			cudaEventRecord(e1, sb);
			evs.emplace_back(e0, e1); ev_fast.push_back(3); n_band += cnt;
			++*nl;
		}
		CK(cudaEventRecord(c.ev_fork, st));
		CK(cudaStreamWaitEvent(c.stf[0], c.ev_fork, 0)); CK(cudaStreamWaitEvent(c.stf[1], c.ev_fork, 0));
		int n_ext_launch = 0;
		for (int k = DPF_NCLASS - 1; k >= 0; --k) { // widest class first: its tasks are the longest
			const int64_t cnt = h_ctr[DP_XBASE + k];
			if (cnt == 0) continue;
			const int32_t *list = lists + (int64_t)(DP_XBASE + k) * n;
			const int mq = (int)h_max[(DP_XBASE + k) * 3];
			unsigned long long *xc = d_cells ? d_cells + 2 : nullptr;
			// the extension classes: small launches with tails, spread over four streams (three of them carry gap-fill launches
			// afterwards) so that they overlap each other and are gone within a few milliseconds
			const int si = n_ext_launch++ & 3;
			cudaStream_t sx = serial ? st : (si == 0 ? c.st2[3] : si == 1 ? st : c.stf[si - 2]);
			if (si == 0) side[3] = true;
			switch (DPF_C[k]) {
			case 4:  launch_ext<4>(tasks, list, ctr + DP_XBASE + k, cnt, mq, codes, S, pool, cigar_pool, sc, xc, sx); break;
			case 5:  launch_ext<5>(tasks, list, ctr + DP_XBASE + k, cnt, mq, codes, S, pool, cigar_pool, sc, xc, sx); break;
			case 6:  launch_ext<6>(tasks, list, ctr + DP_XBASE + k, cnt, mq, codes, S, pool, cigar_pool, sc, xc, sx); break;
			case 7:  launch_ext<7>(tasks, list, ctr + DP_XBASE + k, cnt, mq, codes, S, pool, cigar_pool, sc, xc, sx); break;
			case 8:  launch_ext<8>(tasks, list, ctr + DP_XBASE + k, cnt, mq, codes, S, pool, cigar_pool, sc, xc, sx); break;
			case 9:  launch_ext<9>(tasks, list, ctr + DP_XBASE + k, cnt, mq, codes, S, pool, cigar_pool, sc, xc, sx); break;
			case 10: launch_ext<10>(tasks, list, ctr + DP_XBASE + k, cnt, mq, codes, S, pool, cigar_pool, sc, xc, sx); break;
			case 11: launch_ext<11>(tasks, list, ctr + DP_XBASE + k, cnt, mq, codes, S, pool, cigar_pool, sc, xc, sx); break;
			case 12: launch_ext<12>(tasks, list, ctr + DP_XBASE + k, cnt, mq, codes, S, pool, cigar_pool, sc, xc, sx); break;
			case 14: launch_ext<14>(tasks, list, ctr + DP_XBASE + k, cnt, mq, codes, S, pool, cigar_pool, sc, xc, sx); break;
			case 16: launch_ext<16>(tasks, list, ctr + DP_XBASE + k, cnt, mq, codes, S, pool, cigar_pool, sc, xc, sx); break;
			case 20: launch_ext<20>(tasks, list, ctr + DP_XBASE + k, cnt, mq, codes, S, pool, cigar_pool, sc, xc, sx); break;
			default: launch_ext<24>(tasks, list, ctr + DP_XBASE + k, cnt, mq, codes, S, pool, cigar_pool, sc, xc, sx); break;
			}
		}
		// experiment (MB_DEFER_MODE=1): the deferred exact end extensions beside the extension / band launches, at MB_DEFER_PER_SM CTAs per SM
		static const int defer_mode = getenv("MB_DEFER_MODE") ? atoi(getenv("MB_DEFER_MODE")) : 0;
		static const int defer_cap = getenv("MB_DEFER_PER_SM") ? atoi(getenv("MB_DEFER_PER_SM")) : 3;
		bool deferred_early = false;
		if (defer_mode == 1 && !serial) {
			int k = 0;
			for (int b = DP_NCTA - 1; b >= 0; --b) {
				if (h_ctr[DP_DBASE + b] == 0) continue;
				const int i = k++ % 3;
				if (!(deferred >> i & 1)) CK(cudaStreamWaitEvent(c.st_defer[i], c.ev_fork, 0));
				launch_cta(DP_DBASE + b, c.st_defer[i], defer_cap);
				deferred |= 1 << i;
			}
			deferred_early = true;
		}
		if (!serial) { // the gap-fill and exact launches start when the extension and band launches are done
			CK(cudaEventRecord(c.ev_x[0], c.st2[2])); CK(cudaEventRecord(c.ev_x[1], c.st2[3]));
			CK(cudaEventRecord(c.ev_x[2], c.stf[0])); CK(cudaEventRecord(c.ev_x[3], c.stf[1]));
			CK(cudaEventRecord(c.ev_x[4], c.st2[0]));   // the second band class
			for (int i = band_fills ? 0 : 1; i < 4; ++i) CK(cudaStreamWaitEvent(st, c.ev_x[i], 0));
			if (band_fills) CK(cudaStreamWaitEvent(st, c.ev_x[4], 0));
			CK(cudaEventRecord(c.ev_fork2, st));
			CK(cudaStreamWaitEvent(c.stf[0], c.ev_fork2, 0)); CK(cudaStreamWaitEvent(c.stf[1], c.ev_fork2, 0));
			CK(cudaStreamWaitEvent(c.st2[0], c.ev_fork2, 0)); CK(cudaStreamWaitEvent(c.st2[1], c.ev_fork2, 0));
		}
		for (int b = DP_NCTA - 1; b >= 0; --b) // largest class first
			if (launch_cta(DP_CBASE + b, serial ? st : c.st2[b & 1])) side[b & 1] = true;
		for (int b = DP_NEXACT - 1; b >= 0; --b) {
			const int cls = DP_EBASE + b;
			cudaStream_t st2 = serial ? st : c.st2[b & 1]; // exact classes alternate over side streams 0 and 1
			int64_t cnt = h_ctr[cls];
			if (cnt == 0) continue;
			if (getenv("MB_DEBUG")) fprintf(stderr, "[mb] exact class %d: %lld tasks, max p_bytes %llu, ws_bytes %llu, h_ints %llu\n", b, (long long)cnt, h_max[cls * 3], h_max[cls * 3 + 1], h_max[cls * 3 + 2]);
			size_t p_stride = ((size_t)h_max[cls * 3 + 0] + 255) & ~(size_t)255;
			size_t g_stride = ((size_t)h_max[cls * 3 + 1] + 255) & ~(size_t)255;
			size_t h_stride = ((size_t)h_max[cls * 3 + 2] + 63) & ~(size_t)63;
			if (p_stride == 0) p_stride = 256;
			if (h_stride == 0) h_stride = 64;
			// small tasks: 4 warps per CTA with a 9 KB state slice each; tasks whose state arrays need more get one-warp CTAs
			// with a slice of up to 96 KB (the state lives in global memory only beyond that)
			const size_t slice = ((g_stride + 15) & ~(size_t)15) + h_stride * 4 + 64;
			const bool wide = slice > DP_SMEM_PER_WARP;
			const int wpc = wide ? 1 : DP_WARPS;
			const int smem_per_warp = wide ? (int)(slice < DP_SMEM_MAX ? slice : DP_SMEM_MAX) : DP_SMEM_PER_WARP;
			// one-warp CTAs of the long-task classes: as many per SM as their shared-memory slices allow (up to 8)
			int wide_per_sm = (200 * 1024) / smem_per_warp; if (wide_per_sm < 1) wide_per_sm = 1; if (wide_per_sm > 8) wide_per_sm = 8;
			int max_cta = mb_side_grid(c.num_sms * (wide ? wide_per_sm : 6));
			int64_t want_cta = cdiv(cnt, wpc);
			int n_cta = (int)(want_cta < max_cta ? want_cta : max_cta);
			// bound total scratch to ~24 GB
			size_t per_warp = p_stride + g_stride + h_stride * 4;
			size_t budget = (size_t)24 << 30;
			while (n_cta > 1 && (size_t)n_cta * wpc * per_warp > budget) n_cta = (n_cta + 1) / 2;
			size_t n_warps = (size_t)n_cta * wpc;
			uint8_t *p_scr = ar.get<uint8_t>(n_warps * p_stride);
			int8_t *g_ws = ar.get<int8_t>(n_warps * g_stride + 16);
			int32_t *h_scr = ar.get<int32_t>(n_warps * h_stride);
			int32_t *wc = ar.get<int32_t>(1);
			CK(cudaMemsetAsync(wc, 0, sizeof(int32_t), st2));
			cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
			cudaEventRecord(e0, st2);
			k_dp<<<n_cta, wpc * 32, wpc * smem_per_warp, st2>>>(tasks, lists + (int64_t)cls * n, ctr + cls, wc, codes, S, pool,
				p_scr, p_stride, g_ws, g_stride, h_scr, h_stride, cigar_pool, sc, d_cells ? d_cells + 1 : nullptr, smem_per_warp);
			cudaEventRecord(e1, st2);
			evs.emplace_back(e0, e1); ev_fast.push_back(0); n_exact += cnt;
			++*nl; side[b & 1] = true;
		}
		{
			// largest classes first, rotating over three streams
			int ord[DPF_NCLASS];
			for (int k = 0; k < DPF_NCLASS; ++k) ord[k] = k;
			std::sort(ord, ord + DPF_NCLASS, [&](int a, int b) { return (int64_t)h_ctr[a] * DPF_C[a] > (int64_t)h_ctr[b] * DPF_C[b]; });
			cudaEvent_t w0, w1; cudaEventCreate(&w0); cudaEventCreate(&w1);
			cudaEventRecord(w0, st);
			int slot = 0;
			for (int oi = 0; oi < DPF_NCLASS; ++oi) {
				const int k = ord[oi];
				const int64_t cnt = h_ctr[k];
				if (cnt == 0) continue;
				if (getenv("MB_DEBUG")) fprintf(stderr, "[mb] fast class C=%d: %lld tasks, max qlen %d (launch %zu)\n", DPF_C[k], (long long)cnt, (int)h_max[k * 3], evs.size());
				const int32_t *list = lists + (int64_t)k * n;
				const int mq = (int)h_max[k * 3];
				cudaStream_t sf = (slot == 0 || serial) ? st : c.stf[slot - 1];
				slot = (slot + 1) % 3;
				switch (DPF_C[k]) {
				case 4:  launch_fast<4>(tasks, list, ctr + k, cnt, mq, codes, S, pool, cigar_pool, sc, d_cells, sf); break;
				case 5:  launch_fast<5>(tasks, list, ctr + k, cnt, mq, codes, S, pool, cigar_pool, sc, d_cells, sf); break;
				case 6:  launch_fast<6>(tasks, list, ctr + k, cnt, mq, codes, S, pool, cigar_pool, sc, d_cells, sf); break;
				case 7:  launch_fast<7>(tasks, list, ctr + k, cnt, mq, codes, S, pool, cigar_pool, sc, d_cells, sf); break;
				case 8:  launch_fast<8>(tasks, list, ctr + k, cnt, mq, codes, S, pool, cigar_pool, sc, d_cells, sf); break;
				case 9:  launch_fast<9>(tasks, list, ctr + k, cnt, mq, codes, S, pool, cigar_pool, sc, d_cells, sf); break;
				case 10: launch_fast<10>(tasks, list, ctr + k, cnt, mq, codes, S, pool, cigar_pool, sc, d_cells, sf); break;
				case 11: launch_fast<11>(tasks, list, ctr + k, cnt, mq, codes, S, pool, cigar_pool, sc, d_cells, sf); break;
				case 12: launch_fast<12>(tasks, list, ctr + k, cnt, mq, codes, S, pool, cigar_pool, sc, d_cells, sf); break;
				case 14: launch_fast<14>(tasks, list, ctr + k, cnt, mq, codes, S, pool, cigar_pool, sc, d_cells, sf); break;
				case 16: launch_fast<16>(tasks, list, ctr + k, cnt, mq, codes, S, pool, cigar_pool, sc, d_cells, sf); break;
				case 20: launch_fast<20>(tasks, list, ctr + k, cnt, mq, codes, S, pool, cigar_pool, sc, d_cells, sf); break;
				default: launch_fast<24>(tasks, list, ctr + k, cnt, mq, codes, S, pool, cigar_pool, sc, d_cells, sf); break;
				}
			}
			for (int i = 0; i < 2; ++i) { CK(cudaEventRecord(c.ev_f[i], c.stf[i])); CK(cudaStreamWaitEvent(st, c.ev_f[i], 0)); }
			cudaEventRecord(w1, st);
			fast_wall.emplace_back(w0, w1);
		}
		CK(cudaEventRecord(c.ev_fast_done, st));
		{ // deferred end extensions: behind the gap-fill kernels, beside whatever the caller does next (Z-drop test, second pass)
			static const bool defer_early = getenv("MB_DEFER_EARLY") != nullptr;   // experiment: lowest-priority CTAs fill the gaps of the gap-fill launches
			int k = 0;   // one stream per class: a small batch has a few long tasks per class, and the classes then run side by side
			for (int b = DP_NCTA - 1; b >= 0 && !deferred_early; --b) {
				if (h_ctr[DP_DBASE + b] == 0) continue;
				const int i = k++ % 3;
				if (serial) { launch_cta(DP_DBASE + b, st); continue; }
				if (!(deferred >> i & 1)) CK(cudaStreamWaitEvent(c.st_defer[i], defer_early ? c.ev_fork2 : c.ev_fast_done, 0));
				// At full occupancy (6 CTAs per SM) the persistent CTAs of the deferred launches take every SM's registers, and the Z-drop
				// test and second pass that follow on the main stream wait for them to drain: 18 ms on the configs[1] batch.  When the
				// deferred classes are small (a few tasks per CTA: their duration is set by the longest tasks, not by occupancy) they
				// run at 4 CTAs per SM and the main stream's kernels fit beside them (265.8 -> 261.5 ms per step; 5, 3 and 2 per SM
				// measured worse).  Large deferred classes (long reads) keep the full occupancy.  MB_DEFER_CAP / MB_DEFER_CAP_TASKS override.
				static const int defer_cap2 = getenv("MB_DEFER_CAP") ? atoi(getenv("MB_DEFER_CAP")) : 4;
				static const long long defer_cap_tasks = getenv("MB_DEFER_CAP_TASKS") ? atoll(getenv("MB_DEFER_CAP_TASKS")) : 8192;
				long long n_def = 0;
				for (int b2 = 0; b2 < DP_NCTA; ++b2) n_def += h_ctr[DP_DBASE + b2];
				if (getenv("MB_DEBUG")) fprintf(stderr, "[mb] deferred exact class %d: %lld tasks (of %lld deferred)\n", b, (long long)h_ctr[DP_DBASE + b], n_def);
				launch_cta(DP_DBASE + b, c.st_defer[i], n_def <= defer_cap_tasks ? defer_cap2 : 0);
				deferred |= 1 << i;
			}
			for (int i = 0; i < 3; ++i) if (deferred >> i & 1) CK(cudaEventRecord(c.ev_defer[i], c.st_defer[i]));
		}
		for (int b = 0; b < MB_NSIDE; ++b)
			if (side[b]) { CK(cudaEventRecord(c.ev_join[b], c.st2[b])); CK(cudaStreamWaitEvent(st, c.ev_join[b], 0)); }
	}
};

__global__ void k_copy_i32(const int32_t *__restrict__ src, int32_t *__restrict__ dst, int n)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) dst[i] = src[i];
}

__global__ void k_task_cap(const DpTask *__restrict__ tasks, int64_t n, int32_t *__restrict__ cap)
{
	int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) cap[i] = tasks[i].qlen + tasks[i].tlen + 1;
}

static void check_err(int *d_err, cudaStream_t st, const char *where)
{
	// the device flag only guards invariants the host established (pool sizes); read content cannot raise it
	int e = d2h_scalar(d_err, st);
	if (e) throw mb_error(MB_ERR_OVERFLOW, std::string(where) + ": internal error, device consistency flag " + std::to_string(e));
}

// device-resident result of one sub-batch (arrays live in the arena of `c` until its next reset)
struct DevPart {
	ThreadCtx *c = nullptr;
	int32_t read_lo = 0, n_reads = 0;
	const uint8_t *d_codes = nullptr; const int64_t *d_off = nullptr; int64_t total = 0;
	int32_t *d_fields = nullptr; int64_t *d_hcoff = nullptr; uint32_t *d_hcig = nullptr; int64_t *d_hit_off = nullptr; int32_t *d_rep_len = nullptr;
	int64_t n_h = 0, n_c = 0;
	mb_stats_t S;
	std::exception_ptr err;
	const SketchFeed *feed = nullptr;     // reads still on the host: copy, encode and sketch piecewise (single-piece batches)
};

// the whole device pipeline of one sub-batch on the stream / arena of `c`; synchronises c.st before returning
static void map_device_part(mb_index *ix, const mb_opt_t &opt_in, DevPart &part)
{
	ThreadCtx &c = *part.c;
	const uint8_t *d_codes = part.d_codes; const int64_t *d_off = part.d_off;
	const int32_t n_reads = part.n_reads; const int64_t total = part.total;
	cudaStream_t st = c.st;
	Arena &ar = c.ar;
	mb_opt_t opt = opt_in;
	if (opt.mid_occ <= 0) opt.mid_occ = ix->mid_occ;
	if (opt.q + opt.e >= 127 || opt.q2 + opt.e2 >= 127) throw mb_error(MB_ERR_ARG, "gap costs too large for int8 DP");
	if (opt.max_gap > LL_MAX_LEN - 8) throw mb_error(MB_ERR_ARG, "max_gap above 5000 is not supported (inversion-test scratch)");
	if (!mb_ll_scoring_ok(opt)) throw mb_error(MB_ERR_ARG, "scoring outside the range of the local-alignment kernel (needs b <= q + 2e)");
	mb_stats_t &S = part.S; memset(&S, 0, sizeof(S));
	S.n_reads = n_reads, S.n_bases = total;
	int64_t nl = 0;
	Timer tm(st), tall(st);
	tall.start();
	if (n_reads == 0) return;

	// K1
	tm.start();
	SketchOut so;
	run_sketch(ar, st, d_codes, d_off, n_reads, total, ix->w, ix->k, so, &nl, part.feed);
	S.ms_sketch = tm.stop(); S.n_mini = so.n_mini;
	// K2 + K2b
	tm.start();
	SeedOut sd;
	run_seed(ar, st, ix->d, opt.mid_occ, so.mini, so.mini_off, so.n_mini, d_off, n_reads, sd, &nl, c.num_sms);
	S.ms_seed = tm.stop(); S.n_anchor = sd.n_a;
	const int64_t n_a = sd.n_a;
	// K3
	tm.start();
	ReadScratch rs;
	rs.f = ar.get<int32_t>(n_a + 1), rs.p = ar.get<int32_t>(n_a + 1), rs.v = ar.get<int32_t>(n_a + 1), rs.t = ar.get<int32_t>(n_a + 1);
	rs.b = ar.get<mb128>(n_a + 1); rs.u = ar.get<uint64_t>(n_a + 1); rs.scr = ar.get<uint64_t>(3 * n_a + 3 * (int64_t)n_reads + 3);
	int32_t *wc = ar.get<int32_t>(4);
	unsigned long long *d_cells = ar.get<unsigned long long>(6); // chain, k_dp_fast, k_dp, k_dp_ext, k_dp_band
	int *d_err = ar.get<int>(1);
	CK(cudaMemsetAsync(wc, 0, 4 * sizeof(int32_t), st));
	CK(cudaMemsetAsync(d_cells, 0, 6 * sizeof(unsigned long long), st));
	CK(cudaMemsetAsync(d_err, 0, sizeof(int), st));
	const int max_chain_gap_ref = opt.max_gap_ref > 0 ? opt.max_gap_ref : opt.max_gap, max_chain_gap_qry = opt.max_gap;
	if (n_a > 0) {
		int grid = c.num_sms * 8;
		k_chain_dp<<<grid, CH_WARPS * 32, 0, st>>>(sd.a, sd.a_roff, n_reads, max_chain_gap_ref, max_chain_gap_qry, opt.bw, opt.max_chain_skip, opt.max_chain_iter,
			rs.f, rs.p, rs.v, rs.t, wc, d_cells, sd.read_perm); ++nl;
	}
	S.ms_chain = tm.stop();
	// region logic
	tm.start();
	ReadArrays ra;
	ra.a = sd.a; ra.a_roff = sd.a_roff;
	int32_t *n_u = ar.get<int32_t>(n_reads), *cap = ar.get<int32_t>(n_reads), *n_a_sq = ar.get<int32_t>(n_reads), *n0_regs = ar.get<int32_t>(n_reads);
	ra.n_a_sq = n_a_sq;
	ra.n_regs = ar.get<int32_t>(n_reads);
	ra.iscr = ar.get<int32_t>(n_a + 1);
	int64_t *reg_off = ar.get<int64_t>(n_reads + 1);
	ra.reg_off = reg_off;
	// reads with more than 64 chains / regions take the "big" variants of the per-read kernels (see chain.cuh mb_sort_exact)
	int32_t *big = ar.get<int32_t>(n_reads), *big_ctr = ar.get<int32_t>(8);
	CK(cudaMemsetAsync(big_ctr, 0, 8 * sizeof(int32_t), st));
	const int big_grid = c.num_sms;
	const unsigned rb = (unsigned)cdiv(n_reads, 128);
	k_chain_bt<<<(unsigned)cdiv((int64_t)n_reads * 32, 128), 128, 0, st>>>(ra, rs, n_reads, opt.min_cnt, opt.min_chain_score, n_u, sd.read_perm, big, big_ctr + 0); ++nl;
	k_chain_bt_big<<<big_grid, 32, BIG_SMEM_BYTES, st>>>(ra, rs, opt.min_cnt, opt.min_chain_score, n_u, big, big_ctr + 0, big_ctr + 1); ++nl;
	const int tight_regs = getenv("MB_TEST_TIGHT_REGS") ? 1 : 0; // tests: no slack in the region pool, so that every growth path runs
	k_reg_cap<<<rb, 128, 0, st>>>(n_u, n_reads, tight_regs, cap); ++nl;
	exclusive_scan<int32_t>(ar, st, cap, reg_off, n_reads, &nl);
	int64_t reg_total = d2h_scalar(reg_off + n_reads, st);
	ra.regs = ar.get<Reg>(reg_total + 1);
	rs.regs_tmp = ar.get<Reg>(reg_total + 1);
	int2 *work = ar.get<int2>(reg_total + 1), *work2 = ar.get<int2>(reg_total + 1);
	// [0] regions of round 1, [1] regions split off in the current round, [2] inversion candidates so far, [3] pool-overflow flag
	int32_t *n_work = ar.get<int32_t>(4);
	CK(cudaMemsetAsync(n_work, 0, 4 * sizeof(int32_t), st));
	AlignCtx ac; ac.codes = d_codes; ac.read_off = d_off; ac.ix = ix->d; ac.opt = opt;
	k_gen_regs<<<rb, 128, 0, st>>>(ac, ra, rs, n_reads, n_u, n_a_sq, work, n_work, big, big_ctr + 2); ++nl;
	k_gen_regs_big<<<big_grid, 32, BIG_SMEM_BYTES, st>>>(ac, ra, rs, n_u, n_a_sq, work, n_work, big, big_ctr + 2, big_ctr + 3); ++nl;
	k_copy_i32<<<rb, 128, 0, st>>>(ra.n_regs, n0_regs, n_reads); ++nl;
	int32_t h_n_work = d2h_scalar(n_work, st);
	check_err(d_err, st, "region generation");
	S.ms_glue = tm.stop(); S.n_regs = h_n_work;
	// the work lists and the inversion list hold one entry per region: they grow with the pool
	int64_t work_cap = reg_total + 1;
	int4 *inv_list = ar.get<int4>(work_cap);
	int32_t *need = ar.get<int32_t>(n_reads);
	// make room for `n_items` more regions (one per work item of the coming pass); grows the pool when some read would run out
	auto ensure_regs = [&](const int2 *wk, const int4 *il, int n_items) {
		if (n_items <= 0) return;
		CK(cudaMemsetAsync(need, 0, (size_t)n_reads * sizeof(int32_t), st));
		CK(cudaMemsetAsync(n_work + 3, 0, sizeof(int32_t), st));
		k_round_need<<<(unsigned)cdiv(n_items, 256), 256, 0, st>>>(wk, n_items, il, ra, need, n_work + 3); ++nl;
		if (d2h_scalar(n_work + 3, st) == 0) return;
		int32_t *ncap = ar.get<int32_t>(n_reads);
		int64_t *noff = ar.get<int64_t>(n_reads + 1);
		k_regs_newcap<<<rb, 128, 0, st>>>(ra, need, n_reads, ncap); ++nl;
		exclusive_scan<int32_t>(ar, st, ncap, noff, n_reads, &nl);
		const int64_t ntot = d2h_scalar(noff + n_reads, st);
		Reg *nregs = ar.get<Reg>(ntot + 1);
		k_regs_move<<<(unsigned)cdiv((int64_t)n_reads * 32, 128), 128, 0, st>>>(ra.regs, ra.reg_off, nregs, noff, ra.n_regs, n_reads); ++nl;
		ra.regs = nregs, ra.reg_off = noff, reg_off = noff, reg_total = ntot;
		rs.regs_tmp = ar.get<Reg>(ntot + 1);
		if (ntot + 1 > work_cap) { // lists are rebuilt at the new size (work2 / inv_list contents are copied, `work` is being consumed)
			int2 *w2 = ar.get<int2>(ntot + 1), *w1 = ar.get<int2>(ntot + 1);
			int4 *il2 = ar.get<int4>(ntot + 1);
			CK(cudaMemcpyAsync(w1, work, (size_t)work_cap * sizeof(int2), cudaMemcpyDeviceToDevice, st));
			CK(cudaMemcpyAsync(w2, work2, (size_t)work_cap * sizeof(int2), cudaMemcpyDeviceToDevice, st));
			CK(cudaMemcpyAsync(il2, inv_list, (size_t)work_cap * sizeof(int4), cudaMemcpyDeviceToDevice, st));
			work = w1, work2 = w2, inv_list = il2, work_cap = ntot + 1;
		}
	};
	// alignment rounds
	tm.start();
	const DpScoring scoring = make_scoring(opt);
	DpRunner runner(c, &nl);
	int round = 0;
	int32_t h_n_inv = 0;
	const bool dbg = getenv("MB_DEBUG") != nullptr;
	auto phase = [&](const char *name) {
		auto &t0 = g_dbg_t0;
		if (!dbg) return;
		cudaStreamSynchronize(st);
		auto t1 = std::chrono::steady_clock::now();
		fprintf(stderr, "[mb] phase %-14s %8.3f ms\n", name, std::chrono::duration<double, std::milli>(t1 - t0).count());
		t0 = t1;
	};
	phase("pre-align");
	const int ll_grid = c.num_sms * 8;
	int *ll_scr = nullptr;   // scratch of the local-alignment kernels (inversion tests): allocated on first use
	while (h_n_work > 0) {
		++round;
		const unsigned wb = (unsigned)cdiv(h_n_work, 128);
		RegPlan *plans = ar.get<RegPlan>(h_n_work);
		int32_t *nt = ar.get<int32_t>(h_n_work);
		int64_t *task_off = ar.get<int64_t>(h_n_work + 1);
		ensure_regs(work, nullptr, h_n_work);
		k_plan1<<<wb, 128, 0, st>>>(ac, ra, work, h_n_work, plans); ++nl;
		k_plan_ntasks<<<wb, 128, 0, st>>>(plans, h_n_work, nt); ++nl;
		exclusive_scan<int32_t>(ar, st, nt, task_off, h_n_work, &nl);
		const int64_t n_tasks = d2h_scalar(task_off + h_n_work, st);
		DpTask *tasks = ar.get<DpTask>(n_tasks + 1);
		int32_t *cig_cap = ar.get<int32_t>(n_tasks + 1);
		int64_t *cig_off = ar.get<int64_t>(n_tasks + 2);
		k_plan2<<<wb, 128, 0, st>>>(ac, ra, work, h_n_work, plans, task_off, tasks, cig_cap); ++nl;
		phase("plan1+2");
		int64_t cig_total = 0;
		uint32_t *cigar_pool = nullptr;
		if (n_tasks > 0) {
			exclusive_scan<int32_t>(ar, st, cig_cap, cig_off, n_tasks, &nl);
			cig_total = d2h_scalar(cig_off + n_tasks, st);
			cigar_pool = ar.get<uint32_t>(cig_total + 1);
			k_set_cigar_off<<<(unsigned)cdiv(n_tasks, 256), 256, 0, st>>>(tasks, cig_off, n_tasks, (int64_t)((uintptr_t)cigar_pool / 4)); ++nl;
			cigar_pool = nullptr; // offsets are now absolute word addresses (pools of different rounds coexist)
			// One piece at a time in the (issue-bound) DP kernels; the other pieces meanwhile run their latency-bound stages
			// (sketch, seeding, chaining, region logic before; stitching, mm_update_extra, finalisation after) underneath.
			// (MB_DP_LOCK_MIN_BASES: batches below that size skip the turn-taking; default 0 = none do, see mb_dp_lock_min_bases)
			std::unique_lock<std::mutex> dp_token(g_dp_mutex[c.device & 15], std::defer_lock);
			if (total >= mb_dp_lock_min_bases()) dp_token.lock();
			runner.run(tasks, nullptr, n_tasks, false, d_codes, ix->d.S, nullptr, cigar_pool, scoring, d_cells + 1, true);
			CK(cudaEventSynchronize(c.ev_fast_done)); // the long-tailed side-stream launches of this piece may still be running
			if (dp_token.owns_lock()) dp_token.unlock();
			phase("dp pass 1");
			// Z-drop test and second pass
			int32_t *pass2 = ar.get<int32_t>(n_tasks), *zc = ar.get<int32_t>(4); // zc: [0] second-pass tasks, [1] walk list, [2] inversion-test candidates, [3] cursor
			CK(cudaMemsetAsync(zc, 0, 4 * sizeof(int32_t), st));
			int32_t *walk = ar.get<int32_t>(n_tasks);
			ZCand *zcand = ar.get<ZCand>(n_tasks);
			if (!ll_scr) ll_scr = ar.get<int>((size_t)ll_grid * 4 * LL_MAX_LEN);
			k_ztest_screen<<<(unsigned)cdiv(n_tasks, 128), 128, 0, st>>>(ac, tasks, n_tasks, cigar_pool, walk, zc + 1); ++nl;
			k_ztest<<<(unsigned)cdiv(n_tasks, 128), 128, 0, st>>>(ac, tasks, walk, zc + 1, n_tasks, cigar_pool, pass2, zc + 0, zcand, zc + 2); ++nl;
			k_ztest_ll<<<ll_grid, 32, 0, st>>>(ac, tasks, zcand, zc + 2, zc + 3, pass2, zc + 0, ll_scr); ++nl;
			const int32_t h_pass2 = d2h_scalar(zc, st);
			phase("ztest");
			S.n_dp_pass2 += h_pass2;
			if (h_pass2 > 0) runner.run(tasks, pass2, h_pass2, true, d_codes, ix->d.S, nullptr, cigar_pool, scoring, d_cells + 1);
		}
		S.n_dp_tasks += n_tasks;
		CK(cudaMemsetAsync(n_work + 1, 0, sizeof(int32_t), st));
		// the warp-per-region kernels visit the work list longest region first (see k_work_hist)
		int32_t *perm = nullptr;
		if (h_n_work > 1024) {
			int32_t *hist = ar.get<int32_t>(WORK_NB);
			perm = ar.get<int32_t>(h_n_work);
			CK(cudaMemsetAsync(hist, 0, WORK_NB * sizeof(int32_t), st));
			k_work_hist<<<(unsigned)cdiv(h_n_work, 256), 256, 0, st>>>(work, h_n_work, ra, hist);
			k_work_scan<<<1, WORK_NB, 0, st>>>(hist);
			k_work_scatter<<<(unsigned)cdiv(h_n_work, 256), 256, 0, st>>>(work, h_n_work, ra, hist, perm);
			nl += 3;
		}
		runner.join_deferred();   // the long end extensions of the first pass
		k_stitch<<<(unsigned)cdiv((int64_t)h_n_work * 32, 128), 128, 0, st>>>(ac, ra, work, h_n_work, plans, tasks, cigar_pool, work2, n_work + 1, d_err, perm, inv_list, n_work + 2); ++nl;
		phase("stitch");
		k_update_extra<<<(unsigned)cdiv((int64_t)h_n_work * 32, 128), 128, 0, st>>>(ac, ra, work, h_n_work, plans, tasks, cigar_pool, perm); ++nl;
		phase("update_extra");
		{
			int32_t h2[2];
			CK(cudaMemcpyAsync(h2, n_work + 1, sizeof(h2), cudaMemcpyDeviceToHost, st));
			CK(cudaStreamSynchronize(st));
			h_n_work = h2[0], h_n_inv = h2[1];
		}
		check_err(d_err, st, "alignment round");
		std::swap(work, work2);
	}
	// inversion hits (mm_align1_inv), in passes over the candidate chains
	if (h_n_inv > 0) {
		int4 *cur = inv_list, *nxt = ar.get<int4>(h_n_inv), *proc = ar.get<int4>(h_n_inv);
		InvTask *ll = ar.get<InvTask>(h_n_inv), *task_inv = ar.get<InvTask>(h_n_inv);
		DpTask *itasks = ar.get<DpTask>(h_n_inv + 1);
		int32_t *icap = ar.get<int32_t>(h_n_inv + 1);
		int64_t *ioff = ar.get<int64_t>(h_n_inv + 2);
		int2 *iwork = ar.get<int2>(h_n_inv);
		RegPlan *iplans = ar.get<RegPlan>(h_n_inv);
		int32_t *ic = ar.get<int32_t>(8);  // [0] deferred, [1] processed, [2] local alignments, [3] cursor, [4] DP tasks, [5] inversion hits
		if (!ll_scr) ll_scr = ar.get<int>((size_t)ll_grid * 4 * LL_MAX_LEN);
		int n_cur = h_n_inv, guard = 0;
		ensure_regs(nullptr, cur, n_cur);
		while (n_cur > 0) {
			CK(cudaMemsetAsync(ic, 0, 8 * sizeof(int32_t), st));
			k_inv_plan<<<(unsigned)cdiv(n_cur, 128), 128, 0, st>>>(ac, ra, cur, n_cur, nxt, ic + 0, proc, ic + 1, ll, ic + 2); ++nl;
			k_inv_ll<<<ll_grid, 32, 0, st>>>(ac, ra, ll, ic + 2, ic + 3, ll_scr, itasks, task_inv, icap, ic + 4); ++nl;
			int32_t hc[5];
			CK(cudaMemcpyAsync(hc, ic, sizeof(hc), cudaMemcpyDeviceToHost, st));
			CK(cudaStreamSynchronize(st));
			const int n_dp = hc[4];
			if (n_dp > 0) {
				exclusive_scan<int32_t>(ar, st, icap, ioff, n_dp, &nl);
				const int64_t ctot = d2h_scalar(ioff + n_dp, st);
				uint32_t *ipool = ar.get<uint32_t>(ctot + 1);
				k_set_cigar_off<<<(unsigned)cdiv(n_dp, 256), 256, 0, st>>>(itasks, ioff, n_dp, (int64_t)((uintptr_t)ipool / 4)); ++nl;
				std::unique_lock<std::mutex> dp_token(g_dp_mutex[c.device & 15], std::defer_lock);
				if (total >= mb_dp_lock_min_bases()) dp_token.lock();
				runner.run(itasks, nullptr, n_dp, false, d_codes, ix->d.S, nullptr, nullptr, scoring, d_cells + 1);
				CK(cudaEventSynchronize(c.ev_fast_done));
				if (dp_token.owns_lock()) dp_token.unlock();
				k_inv_finish<<<(unsigned)cdiv(n_dp, 128), 128, 0, st>>>(ac, ra, itasks, task_inv, n_dp, iwork, iplans, ic + 5, d_err); ++nl;
				const int32_t n_ok = d2h_scalar(ic + 5, st);
				if (n_ok > 0) { k_update_extra<<<(unsigned)cdiv((int64_t)n_ok * 32, 128), 128, 0, st>>>(ac, ra, iwork, n_ok, iplans, itasks, nullptr, nullptr); ++nl; }
				S.n_dp_tasks += n_dp; S.n_inv += n_ok;
				// the arrays of this pass are consumed; later passes (deeper split chains) get fresh ones
				itasks = ar.get<DpTask>(hc[0] + 1); task_inv = ar.get<InvTask>(hc[0] + 1);
			}
			k_inv_close<<<(unsigned)cdiv(n_cur, 128), 128, 0, st>>>(ra, proc, ic + 1); ++nl;
			CK(cudaStreamSynchronize(st));
			if (hc[1] == 0 && ++guard > 2) throw mb_error(MB_ERR_OVERFLOW, "inversion pass made no progress (internal error)");
			n_cur = hc[0];
			std::swap(cur, nxt);
		}
		check_err(d_err, st, "inversion pass");
	}
	S.n_rounds = round;
	S.ms_dp = tm.stop();
	// finish
	tm.start();
	int32_t *n_hits = ar.get<int32_t>(n_reads), *n_hit_cig = ar.get<int32_t>(n_reads);
	int64_t *hit_off = ar.get<int64_t>(n_reads + 1), *hcig_off = ar.get<int64_t>(n_reads + 1);
	k_finish<<<rb, 128, 0, st>>>(ac, ra, rs, n_reads, n0_regs, sd.rep_len, n_hits, n_hit_cig, big, big_ctr + 4); ++nl;
	k_finish_big<<<big_grid, 32, BIG_SMEM_BYTES, st>>>(ac, ra, rs, n0_regs, sd.rep_len, n_hits, n_hit_cig, big, big_ctr + 4, big_ctr + 5); ++nl;
	exclusive_scan<int32_t>(ar, st, n_hits, hit_off, n_reads, &nl);
	exclusive_scan<int32_t>(ar, st, n_hit_cig, hcig_off, n_reads, &nl);
	const int64_t n_h = d2h_scalar(hit_off + n_reads, st);
	const int64_t n_c = d2h_scalar(hcig_off + n_reads, st);
	check_err(d_err, st, "finalisation");
	int32_t *d_fields = ar.get<int32_t>((size_t)HIT_NF * n_h + 1);
	int64_t *d_hcoff = ar.get<int64_t>(n_h + 1);
	uint32_t *d_hcig = ar.get<uint32_t>(n_c + 1);
	// cigars of surviving regions live in per-round pools that are all still allocated in the arena
	k_write_hits<<<(unsigned)cdiv((int64_t)n_reads * 32, 128), 128, 0, st>>>(ra, n_reads, hit_off, hcig_off, n_h, d_fields, d_hcoff, (const uint32_t*)nullptr, d_hcig); ++nl;
	S.ms_post = tm.stop();
	S.n_hits = n_h;
	part.d_fields = d_fields, part.d_hcoff = d_hcoff, part.d_hcig = d_hcig, part.d_hit_off = hit_off, part.d_rep_len = sd.rep_len;
	part.n_h = n_h, part.n_c = n_c;
	unsigned long long h_cells[6];
	CK(cudaMemcpyAsync(h_cells, d_cells, sizeof(h_cells), cudaMemcpyDeviceToHost, st));
	CK(cudaStreamSynchronize(st));
	CK(cudaGetLastError());
	S.dp_cells = (int64_t)(h_cells[1] + h_cells[2] + h_cells[3] + h_cells[4]); S.dp_cells_exact = (int64_t)h_cells[2]; S.dp_cells_ext = (int64_t)h_cells[3]; S.chain_cells = (int64_t)h_cells[0];
	S.dp_cells_band = (int64_t)h_cells[4]; S.ms_kdp_band = runner.total_ms(3); S.n_band_tasks = runner.n_band;
	S.n_launches = nl;
	if (dbg) for (size_t i = 0; i < runner.evs.size(); ++i) { float ms = 0; cudaEventElapsedTime(&ms, runner.evs[i].first, runner.evs[i].second); float a = 0, b = 0; if (runner.ev_base) { cudaEventElapsedTime(&a, runner.ev_base, runner.evs[i].first); cudaEventElapsedTime(&b, runner.ev_base, runner.evs[i].second); } fprintf(stderr, "[mb] dp launch %2zu kind %d  %8.3f ms  [%8.3f .. %8.3f]\n", i, runner.ev_fast[i], ms, a, b); }
	S.ms_kdp = runner.total_ms(); S.n_kdp = (int32_t)runner.evs.size();
	{ int64_t nf = 0; for (int f : runner.ev_fast) nf += (f == 1); S.n_kdp_fast = nf; }
	S.ms_kdp_fast = runner.fast_wall_ms(); S.ms_kdp_exact = runner.total_ms(0); S.n_fast_tasks = runner.n_fast; S.n_exact_tasks = runner.n_exact;
	S.ms_kdp_ext = runner.total_ms(2); S.n_ext_tasks = runner.n_ext;
	S.arena_bytes = (int64_t)ar.batch_total;
	if (total > ((int64_t)1 << 20)) g_scratch_per_base_x16.store((long long)(ar.batch_total * 16 / (size_t)total));
	S.ms_total = tall.stop();
}

__global__ void k_rebase_offsets(const int64_t *__restrict__ src, int64_t *__restrict__ dst, int64_t n, int64_t base)
{
	int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) dst[i] = src[i] - base;
}

#define MB_NPART_MAX 4
static int mb_n_parts(int32_t n_reads, int64_t total)
{
	const char *e = getenv("MB_PARTS"); // number of concurrent pieces (default 1: measured on B200, 2-4 pieces lose more to smaller launches than they hide); MB_PARTS_MIN_READS lowers the size gate (tests)
	int conf = e ? atoi(e) : 1; if (conf < 1) conf = 1; if (conf > MB_NPART_MAX) conf = MB_NPART_MAX;
	const char *m = getenv("MB_PARTS_MIN_READS");
	const int64_t min_reads = m ? atoll(m) : 1024;
	if (n_reads < min_reads * conf || (!m && total < ((int64_t)4 << 20) * conf)) return 1; // small batches: one piece
	return conf;
}

// Map a device-resident batch.  The batch is cut (at read boundaries, by cumulative bases) into pieces that run
// concurrently on their own streams / arenas / host threads; results are concatenated in read order, so the outcome does
// not depend on the number of pieces.
// cold or too small arena: one reservation from what the last batch of the process needed per base, instead of doubling slab by slab
static void cold_reserve(ThreadCtx &c, int64_t total)
{
	if (total <= 0) return;
	const long long per16 = g_scratch_per_base_x16.load();
	size_t fr = 0, tot = 0;
	if (per16 <= 0 || cudaMemGetInfo(&fr, &tot) != cudaSuccess) { cudaGetLastError(); return; }
	const size_t want = (size_t)((double)per16 / 16.0 * (double)total * 1.15) + ((size_t)256 << 20);
	if (want > c.ar.cap && want + ((size_t)4 << 30) < fr + c.ar.cap) c.ar.reserve(want);   // (also a warm arena that is too small: one slab, not a doubling sequence)
}

static mb_hits *map_device(mb_index *ix, const mb_opt_t &opt, ThreadCtx &c, const uint8_t *d_codes, const int64_t *d_off, const int64_t *h_off,
                           int32_t n_reads, int64_t total, int want_hits /* bit 0: hit fields, bit 1: CIGARs */, mb_stats_t *stats, const SketchFeed *feed = nullptr)
{
	auto t_begin = std::chrono::steady_clock::now();
	std::unique_ptr<mb_hits> H(new mb_hits());
	H->n_reads = n_reads;
	H->read_off.assign(h_off, h_off + n_reads + 1);
	H->hit_off.assign(n_reads + 1, 0);
	H->rep_len.assign(n_reads, 0);
	mb_stats_t S; memset(&S, 0, sizeof(S));
	S.n_reads = n_reads, S.n_bases = total;
	c.last_parts.clear(); c.last_n_reads = n_reads; c.last_index = ix->uid;
	if (n_reads == 0) { if (stats) *stats = S; return H.release(); }
	const int K = mb_n_parts(n_reads, total);
	if (feed && K > 1) { // several pieces read the codes from other streams: finish the upload first
		if (feed->h_words) { // packed input: all words of the batch (one word of look-ahead), then the runs of ambiguous bases
			const int64_t n_w = ((feed->sh / 2 + total + 15) >> 4) + 1;
			CK(cudaMemcpyAsync(feed->d_words, feed->h_words, (size_t)n_w * 4, cudaMemcpyHostToDevice, c.st));
			k_unpack_nt4<<<(unsigned)cdiv(cdiv(total, 16), 256), 256, 0, c.st>>>(feed->d_words, feed->sh, feed->d_codes, total);
			if (feed->n_iv) k_apply_amb<<<(unsigned)cdiv(feed->n_iv * 32, 256), 256, 0, c.st>>>(feed->d_iv, feed->n_iv, feed->g0, feed->d_codes, 0, total);
		} else {
			CK(cudaMemcpyAsync(feed->d_ascii, feed->h_ascii, (size_t)total, cudaMemcpyHostToDevice, c.st));
			k_encode_nt4<<<(unsigned)cdiv(cdiv(total, 16), 256), 256, 0, c.st>>>(feed->d_ascii, feed->d_codes, total);
		}
		feed = nullptr;
	}
	while ((int)c.helpers.size() < K - 1) c.helpers.push_back(make_ctx(c.device));
	std::vector<DevPart> parts(K);
	// cut points: first read whose start offset reaches the k-th share of the bases
	std::vector<int32_t> cut(K + 1, 0);
	cut[K] = n_reads;
	for (int k = 1; k < K; ++k) {
		const int64_t target = total / K * k;
		int32_t i = (int32_t)(std::lower_bound(h_off, h_off + n_reads, target) - h_off);
		cut[k] = std::max(cut[k - 1], std::min(i, n_reads));
	}
	CK(cudaStreamSynchronize(c.st)); // the reads (and anything else queued by the caller) are complete before other streams read them
	for (int k = 0; k < K; ++k) {
		DevPart &p = parts[k];
		p.c = k == 0 ? &c : c.helpers[k - 1];
		p.read_lo = cut[k], p.n_reads = cut[k + 1] - cut[k];
		const int64_t base = h_off[cut[k]] & ~(int64_t)15; // keep the piece's code pointer 16-byte aligned (vector loads); its first
		p.total = h_off[cut[k + 1]] - base;                 // read then starts at offset 0..15 instead of 0
		p.d_codes = d_codes + base;
		p.feed = feed;
		if (k > 0) p.c->ar.reset();
		if (K == 1) p.d_off = d_off;
		else {
			int64_t *o = p.c->ar.get<int64_t>(p.n_reads + 1);
			k_rebase_offsets<<<(unsigned)cdiv(p.n_reads + 1, 256), 256, 0, p.c->st>>>(d_off + cut[k], o, p.n_reads + 1, base);
			p.d_off = o;
		}
	}
	auto run_part = [&](int k) {
		try { CK(cudaSetDevice(c.device)); map_device_part(ix, opt, parts[k]); }
		catch (...) { parts[k].err = std::current_exception(); }
	};
	{
		std::vector<std::thread> th;
		for (int k = 1; k < K; ++k) th.emplace_back(run_part, k);
		run_part(0);
		for (auto &t : th) t.join();
	}
	for (int k = 0; k < K; ++k) if (parts[k].err) { for (int j = 0; j < K; ++j) cudaStreamSynchronize(parts[j].c->st); std::rethrow_exception(parts[k].err); }
	// ---- assemble ----
	int64_t n_h = 0, n_c = 0;
	std::vector<int64_t> hbase(K), cbase(K);
	for (int k = 0; k < K; ++k) { hbase[k] = n_h, cbase[k] = n_c; n_h += parts[k].n_h; n_c += parts[k].n_c; }
	H->n = n_h;
	Timer tm(c.st); tm.start();
	if (want_hits) {
		const bool want_cig = (want_hits & 2) != 0;
		H->fields.resize((size_t)HIT_NF * n_h); H->cigar_off.resize(n_h); H->cigar.resize(want_cig ? n_c : 0);
		for (int k = 0; k < K; ++k) {
			const DevPart &p = parts[k];
			cudaStream_t st = p.c->st;
			for (int f = 0; f < HIT_NF && p.n_h; ++f)
				CK(cudaMemcpyAsync(H->fields.data() + (size_t)f * n_h + hbase[k], p.d_fields + (size_t)f * p.n_h, (size_t)p.n_h * 4, cudaMemcpyDeviceToHost, st));
			if (p.n_h) CK(cudaMemcpyAsync(H->cigar_off.data() + hbase[k], p.d_hcoff, (size_t)p.n_h * 8, cudaMemcpyDeviceToHost, st));
			if (p.n_c && want_cig) CK(cudaMemcpyAsync(H->cigar.data() + cbase[k], p.d_hcig, (size_t)p.n_c * 4, cudaMemcpyDeviceToHost, st));
			if (p.n_reads) {
				CK(cudaMemcpyAsync(H->rep_len.data() + p.read_lo, p.d_rep_len, (size_t)p.n_reads * 4, cudaMemcpyDeviceToHost, st));
				CK(cudaMemcpyAsync(H->hit_off.data() + p.read_lo, p.d_hit_off, (size_t)(p.n_reads + (k == K - 1 ? 1 : 0)) * 8, cudaMemcpyDeviceToHost, st));
			}
		}
		for (int k = 0; k < K; ++k) CK(cudaStreamSynchronize(parts[k].c->st));
		for (int k = 1; k < K; ++k) { // piece-relative indices -> batch indices
			const DevPart &p = parts[k];
			int32_t *ridx = H->fields.data() + hbase[k];
			for (int64_t i = 0; i < p.n_h; ++i) ridx[i] += p.read_lo;
			int64_t *co = H->cigar_off.data() + hbase[k];
			for (int64_t i = 0; i < p.n_h; ++i) co[i] += cbase[k];
			int64_t *ho = H->hit_off.data() + p.read_lo;
			for (int32_t i = 0; i < p.n_reads + (k == K - 1 ? 1 : 0); ++i) ho[i] += hbase[k];
		}
	}
	CK(cudaGetLastError());
	S.ms_d2h = tm.stop();
	for (int k = 0; k < K; ++k) {
		const DevPart &p = parts[k];
		ThreadCtx::LastPart lp; lp.fields = p.d_fields, lp.hit_off = p.d_hit_off, lp.read_off = p.d_off, lp.n_hits = p.n_h, lp.n_reads = p.n_reads;
		c.last_parts.push_back(lp);
		const mb_stats_t &q = p.S;
		S.n_mini += q.n_mini, S.n_anchor += q.n_anchor, S.n_regs += q.n_regs, S.n_dp_tasks += q.n_dp_tasks, S.n_dp_pass2 += q.n_dp_pass2;
		S.dp_cells += q.dp_cells, S.n_hits += q.n_hits, S.n_rounds = std::max(S.n_rounds, q.n_rounds), S.n_launches += q.n_launches;
		S.ms_sketch += q.ms_sketch, S.ms_seed += q.ms_seed, S.ms_sort += q.ms_sort, S.ms_chain += q.ms_chain, S.ms_glue += q.ms_glue;
		S.ms_dp += q.ms_dp, S.ms_post += q.ms_post, S.ms_kdp += q.ms_kdp, S.n_kdp += q.n_kdp, S.ms_kdp_fast += q.ms_kdp_fast, S.ms_kdp_exact += q.ms_kdp_exact;
		S.n_fast_tasks += q.n_fast_tasks, S.n_exact_tasks += q.n_exact_tasks, S.chain_cells += q.chain_cells, S.dp_cells_exact += q.dp_cells_exact;
		S.n_kdp_fast += q.n_kdp_fast, S.n_ext_tasks += q.n_ext_tasks, S.dp_cells_ext += q.dp_cells_ext, S.ms_kdp_ext += q.ms_kdp_ext;
		S.n_inv += q.n_inv, S.arena_bytes = std::max(S.arena_bytes, q.arena_bytes);
		S.dp_cells_band += q.dp_cells_band, S.ms_kdp_band += q.ms_kdp_band, S.n_band_tasks += q.n_band_tasks;
	}
	S.ms_total = (float)std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count();
	if (stats) { float h2d = stats->ms_h2d; *stats = S; stats->ms_h2d = h2d; }
	return H.release();
}

// src: the whole batch on the host (ASCII or packed); g0: the first base of this piece within it; off: the piece's own offsets (from 0)
static void upload_reads(ThreadCtx &c, const HostSrc &src, int64_t g0, const int64_t *off, int32_t n_reads, uint8_t **d_codes, int64_t **d_off, int64_t *total, bool persistent,
                         SketchFeed *feed = nullptr)
{
	cudaStream_t st = c.st;
	*total = n_reads > 0 ? off[n_reads] : 0;
	if (n_reads < 0 || (n_reads > 0 && off[0] != 0)) throw mb_error(MB_ERR_ARG, "offsets must start at 0");
	for (int i = 0; i < n_reads; ++i) {
		if (off[i + 1] < off[i]) throw mb_error(MB_ERR_ARG, "offsets must be non-decreasing");
		if (off[i + 1] - off[i] > 0x3fffffff) throw mb_error(MB_ERR_ARG, "read longer than 2^30");
	}
	uint8_t *d_ascii = nullptr; uint32_t *d_words = nullptr; int64_t *d_iv = nullptr;
	const uint32_t *h_words = nullptr;
	int64_t n_w = 0, n_iv = 0;
	const int r = (int)(g0 & 15);
	if (src.pk) { // words [g0 >> 4, ...) of the batch: the piece starts r bases into its first word
		const mb_packed &P = *src.pk;
		if (g0 + *total > P.total) throw mb_error(MB_ERR_ARG, "piece beyond the packed batch");
		h_words = P.words + (g0 >> 4);
		n_w = ((r + *total + 15) >> 4) + 2;
		if ((size_t)((g0 >> 4) + n_w) > P.n_words) n_w = (int64_t)P.n_words - (g0 >> 4);
		d_words = c.ar.get<uint32_t>((size_t)n_w + 4);
		// the runs of ambiguous bases that touch [g0, g0 + total)
		const int64_t n_all = (int64_t)P.iv.size() / 2;
		int64_t lo = 0, hi = n_all;
		while (lo < hi) { const int64_t m = (lo + hi) >> 1; if (P.iv[2 * m] + P.iv[2 * m + 1] <= g0) lo = m + 1; else hi = m; }
		int64_t k1 = lo;
		while (k1 < n_all && P.iv[2 * k1] < g0 + *total) ++k1;
		n_iv = k1 - lo;
		if (n_iv) {
			d_iv = c.ar.get<int64_t>((size_t)n_iv * 2);
			CK(cudaMemcpyAsync(d_iv, P.iv.data() + 2 * lo, (size_t)n_iv * 16, cudaMemcpyHostToDevice, st));
		}
	} else d_ascii = c.ar.get<uint8_t>(*total + 32);
	if (persistent) { CK(cudaMalloc(d_codes, *total + 32)); CK(cudaMalloc(d_off, (n_reads + 1) * 8)); }
	else { *d_codes = c.ar.get<uint8_t>(*total + 32); *d_off = c.ar.get<int64_t>(n_reads + 1); }
	CK(cudaMemcpyAsync(*d_off, off, (n_reads + 1) * 8, cudaMemcpyHostToDevice, st));
	if (feed) { // the caller overlaps the copy with the sketch (run_sketch, SketchFeed)
		feed->h_ascii = src.pk ? nullptr : src.ascii + g0, feed->d_ascii = d_ascii, feed->d_codes = *d_codes, feed->copy_st = c.stf[0], feed->events = &c.feed_events;
		feed->h_words = h_words, feed->d_words = d_words, feed->sh = 2 * r, feed->d_iv = d_iv, feed->n_iv = n_iv, feed->g0 = g0;
		return;
	}
	if (!*total) return;
	if (src.pk) {
		CK(cudaMemcpyAsync(d_words, h_words, (size_t)n_w * 4, cudaMemcpyHostToDevice, st));
		k_unpack_nt4<<<(unsigned)cdiv(cdiv(*total, 16), 256), 256, 0, st>>>(d_words, 2 * r, *d_codes, *total);
		if (n_iv) k_apply_amb<<<(unsigned)cdiv(n_iv * 32, 256), 256, 0, st>>>(d_iv, n_iv, g0, *d_codes, 0, *total);
	} else {
		CK(cudaMemcpyAsync(d_ascii, src.ascii + g0, *total, cudaMemcpyHostToDevice, st));
		k_encode_nt4<<<(unsigned)cdiv(cdiv(*total, 16), 256), 256, 0, st>>>(d_ascii, *d_codes, *total);
	}
}

// ---- sequential pieces: batches whose scratch would not fit the device are mapped piece by piece ----
static int64_t mb_piece_bases()
{
	const char *e = getenv("MB_PIECE_BASES"); // bases per sequential piece; the default keeps the scratch of a piece around 65 GB
	const int64_t v = e ? atoll(e) : (int64_t)1000000000;
	return v < 1000 ? 1000 : v;
}

static size_t store_put(ThreadCtx &c, const void *d_src, size_t bytes)
{
	bytes = (bytes + 255) & ~(size_t)255;
	if (c.store_used + bytes > c.store_cap) {
		size_t ncap = c.store_cap ? c.store_cap * 2 : ((size_t)64 << 20);
		while (ncap < c.store_used + bytes) ncap <<= 1;
		char *nb = nullptr;
		CK(cudaMalloc(&nb, ncap));
		if (c.store_used) CK(cudaMemcpyAsync(nb, c.store, c.store_used, cudaMemcpyDeviceToDevice, c.st));
		CK(cudaStreamSynchronize(c.st));
		if (c.store) cudaFree(c.store);
		c.store = nb, c.store_cap = ncap;
	}
	const size_t o = c.store_used;
	if (bytes) CK(cudaMemcpyAsync(c.store + o, d_src, bytes, cudaMemcpyDeviceToDevice, c.st));
	c.store_used += bytes;
	return o;
}

template <typename T> static void concat_pin(PinVec<T> &dst, const std::vector<std::pair<const T*, size_t>> &src)
{
	size_t n = 0;
	for (auto &p : src) n += p.second;
	dst.resize(n);
	std::vector<std::thread> th;
	size_t o = 0;
	for (auto &p : src) {
		if (p.second) {
			T *d = dst.data() + o; const T *sp = p.first; const size_t cnt = p.second;
			if (cnt * sizeof(T) > ((size_t)8 << 20)) th.emplace_back([d, sp, cnt]() { memcpy(d, sp, cnt * sizeof(T)); });
			else memcpy(d, sp, cnt * sizeof(T));
		}
		o += p.second;
	}
	for (auto &t : th) t.join();
}

// The scratch a piece needs depends on the database (anchors per base grow with its size and repetitiveness), so the piece
// size is found by doing: a batch (or piece) that runs out of device memory is retried in pieces of half the size, and the
// size that worked is remembered for the following calls of this thread.
static int64_t piece_now(const ThreadCtx &c, const mb_index *ix) { const int64_t d = mb_piece_bases(); return c.piece_index == ix->uid && c.piece_bases > 0 && c.piece_bases < d ? c.piece_bases : d; }
static bool shrink_piece(ThreadCtx &c, const mb_index *ix, const mb_error &e, int64_t total, int32_t n_reads)
{
	const int64_t piece = piece_now(c, ix);
	if (e.code != MB_ERR_NOMEM || n_reads <= 1 || piece <= 1000) return false;
	cudaDeviceSynchronize(); cudaGetLastError();
	c.ar.release();
	c.last_parts.clear();
	c.piece_bases = std::max<int64_t>(1000, std::min(piece, total) / 2); c.piece_index = ix->uid;
	if (getenv("MB_DEBUG")) fprintf(stderr, "[mb] device scratch did not fit: retrying in pieces of %lld bases\n", (long long)c.piece_bases);
	return true;
}

// h_src.any(): reads in host memory (uploaded piece by piece); otherwise d_codes_all / d_off_all hold the resident batch
static mb_hits *map_in_pieces(mb_index *ix, const mb_opt_t &opt, ThreadCtx &c, const HostSrc &h_src, const uint8_t *d_codes_all, const int64_t *d_off_all,
                              const int64_t *h_off, int32_t n_reads, int want, mb_stats_t *stats, int64_t piece)
{
	std::vector<int32_t> cut(1, 0);
	while (cut.back() < n_reads) { // at least one read per piece, then as many as fit
		int32_t lo = cut.back(), hi = lo + 1;
		hi = (int32_t)(std::upper_bound(h_off + hi, h_off + n_reads + 1, h_off[lo] + piece) - h_off) - 1;
		if (hi <= lo) hi = lo + 1;
		cut.push_back(hi);
	}
	const int K = (int)cut.size() - 1;
	struct Kept { size_t fields, hit_off, read_off; int64_t n_hits; int32_t n_reads; };
	std::vector<Kept> kept;
	c.store_used = 0;
	mb_stats_t S; memset(&S, 0, sizeof(S));
	float ms_h2d = 0;
	// The caller's result object is put together on the host piece by piece (pageable memory, grown geometrically), so that a
	// piece's page-locked buffers go back to the pool before the next piece needs them: a 1 M-read batch then page-locks what
	// one piece returns, not what nine do.
	std::unique_ptr<mb_hits> H(new mb_hits());
	H->n_reads = n_reads;
	H->read_off.assign(h_off, h_off + n_reads + 1);
	H->fields.plain = H->cigar_off.plain = H->cigar.plain = H->rep_len.plain = H->hit_off.plain = true;
	H->hit_off.resize(n_reads + 1); H->rep_len.resize(n_reads);
	std::vector<PinVec<int32_t>> fcol(HIT_NF);   // hit fields by column while the total is unknown
	for (auto &v : fcol) v.plain = true;
	int64_t n_h = 0, n_c = 0;
	for (int k = 0; k < K; ++k) {
		const int32_t lo = cut[k], n = cut[k + 1] - lo;
		std::vector<int64_t> po(n + 1);
		for (int32_t i = 0; i <= n; ++i) po[i] = h_off[lo + i] - h_off[lo];
		c.ar.reset();
		cold_reserve(c, po[n]);
		uint8_t *d_codes; int64_t *d_off; int64_t total = po[n];
		SketchFeed feed; bool use_feed = false;
		if (h_src.any()) {
			Timer tm(c.st); tm.start();
			const char *feed_env = getenv("MB_FEED_MIN_BYTES");
			const int64_t feed_min = feed_env ? atoll(feed_env) : ((int64_t)64 << 20);
			use_feed = total >= feed_min;
			upload_reads(c, h_src, h_off[lo], po.data(), n, &d_codes, &d_off, &total, false, use_feed ? &feed : nullptr);
			ms_h2d += tm.stop();
		} else {
			const int64_t base = h_off[lo] & ~(int64_t)15;   // keep the code pointer 16-byte aligned (vector loads)
			for (int32_t i = 0; i <= n; ++i) po[i] = h_off[lo + i] - base;
			total = po[n];
			d_codes = const_cast<uint8_t*>(d_codes_all) + base;
			d_off = c.ar.get<int64_t>(n + 1);
			k_rebase_offsets<<<(unsigned)cdiv(n + 1, 256), 256, 0, c.st>>>(d_off_all + lo, d_off, n + 1, base);
		}
		mb_stats_t st1; memset(&st1, 0, sizeof(st1));
		std::unique_ptr<mb_hits> part(map_device(ix, opt, c, d_codes, d_off, po.data(), n, total, want, &st1, use_feed ? &feed : nullptr));
		// keep what mb_count_last needs beyond the next arena reset
		for (const ThreadCtx::LastPart &lp : c.last_parts) {
			Kept kp; kp.n_hits = lp.n_hits, kp.n_reads = lp.n_reads;
			kp.fields = store_put(c, lp.fields, (size_t)HIT_NF * lp.n_hits * 4);
			kp.hit_off = store_put(c, lp.hit_off, (size_t)(lp.n_reads + 1) * 8);
			kp.read_off = store_put(c, lp.read_off, (size_t)(lp.n_reads + 1) * 8);
			kept.push_back(kp);
		}
		CK(cudaStreamSynchronize(c.st));
		// append this piece to the caller's result
		const int64_t m = part->n;
		if (want) {
			for (int32_t i = 0; i < n; ++i) { H->hit_off.data()[lo + i] = part->hit_off.data()[i] + n_h; H->rep_len.data()[lo + i] = part->rep_len.data()[i]; }
			std::vector<std::thread> th;
			for (int f = 0; f < HIT_NF; ++f) th.emplace_back([&, f]() {
				const size_t at = fcol[f].n;
				fcol[f].append_plain(part->fields.data() + (size_t)f * m, (size_t)m);
				if (f == 0) for (int64_t i = 0; i < m; ++i) fcol[0].p[at + i] += lo;   // read_idx: piece-relative -> batch
			});
			if (k == 0 && K > 1) { // size the result once from the first piece (reads are alike): later appends then do not move it
				const double scale = 1.08 * (double)(h_off[n_reads] - h_off[0]) / (double)std::max<int64_t>(1, po[n] - po[0]);
				H->cigar.append_plain(nullptr, 0); H->cigar_off.append_plain(nullptr, 0);
				auto reserve = [&](auto &v, size_t want) { if (want * sizeof(*v.p) > v.cap) { void *np = realloc(v.p, want * sizeof(*v.p) + 64); if (np) { v.p = (decltype(v.p))np; v.cap = want * sizeof(*v.p) + 64; } } };
				reserve(H->cigar, (size_t)(scale * (double)part->cigar.size()));
				reserve(H->cigar_off, (size_t)(scale * (double)m));
			}
			const size_t at = H->cigar_off.n;
			H->cigar_off.append_plain(part->cigar_off.data(), (size_t)m);
			for (int64_t i = 0; i < m; ++i) H->cigar_off.p[at + i] += n_c;
			H->cigar.append_plain(part->cigar.data(), part->cigar.size());
			for (auto &t : th) t.join();
		}
		n_h += m, n_c += (int64_t)part->cigar.size();
		S.n_reads += st1.n_reads, S.n_bases += st1.n_bases, S.n_mini += st1.n_mini, S.n_anchor += st1.n_anchor, S.n_regs += st1.n_regs;
		S.n_dp_tasks += st1.n_dp_tasks, S.n_dp_pass2 += st1.n_dp_pass2, S.dp_cells += st1.dp_cells, S.n_hits += st1.n_hits;
		S.n_rounds = std::max(S.n_rounds, st1.n_rounds), S.n_launches += st1.n_launches + (h_src.any() ? 1 : 0);
		S.ms_sketch += st1.ms_sketch, S.ms_seed += st1.ms_seed, S.ms_sort += st1.ms_sort, S.ms_chain += st1.ms_chain, S.ms_glue += st1.ms_glue;
		S.ms_dp += st1.ms_dp, S.ms_post += st1.ms_post, S.ms_total += st1.ms_total, S.ms_d2h += st1.ms_d2h;
		S.ms_kdp += st1.ms_kdp, S.n_kdp += st1.n_kdp, S.ms_kdp_fast += st1.ms_kdp_fast, S.ms_kdp_exact += st1.ms_kdp_exact, S.ms_kdp_ext += st1.ms_kdp_ext;
		S.n_fast_tasks += st1.n_fast_tasks, S.n_exact_tasks += st1.n_exact_tasks, S.chain_cells += st1.chain_cells, S.dp_cells_exact += st1.dp_cells_exact;
		S.n_kdp_fast += st1.n_kdp_fast, S.n_ext_tasks += st1.n_ext_tasks, S.dp_cells_ext += st1.dp_cells_ext, S.n_inv += st1.n_inv;
		S.arena_bytes = std::max(S.arena_bytes, st1.arena_bytes);
		S.dp_cells_band += st1.dp_cells_band, S.ms_kdp_band += st1.ms_kdp_band, S.n_band_tasks += st1.n_band_tasks;
	}
	S.ms_h2d = ms_h2d, S.n_pieces = K;
	// device-resident view of the whole batch for mb_count_last
	c.last_parts.clear(); c.last_n_reads = n_reads; c.last_index = ix->uid;
	for (const Kept &kp : kept) {
		ThreadCtx::LastPart lp;
		lp.fields = reinterpret_cast<const int32_t*>(c.store + kp.fields), lp.hit_off = reinterpret_cast<const int64_t*>(c.store + kp.hit_off);
		lp.read_off = reinterpret_cast<const int64_t*>(c.store + kp.read_off), lp.n_hits = kp.n_hits, lp.n_reads = kp.n_reads;
		c.last_parts.push_back(lp);
	}
	H->n = n_h;
	H->hit_off.data()[n_reads] = n_h;
	if (!want) { for (int32_t i = 0; i < n_reads; ++i) H->hit_off.data()[i] = 0, H->rep_len.data()[i] = 0; }
	if (want) { // the SoA block [field][hit]
		H->fields.resize((size_t)HIT_NF * n_h);
		std::vector<std::thread> th;
		for (int f = 0; f < HIT_NF; ++f) th.emplace_back([&, f]() { if (n_h) memcpy(H->fields.data() + (size_t)f * n_h, fcol[f].p, (size_t)n_h * 4); });
		for (auto &t : th) t.join();
	}
	if (stats) *stats = S;
	return H.release();
}

static int map_batch_impl(mb_index_t *ix, const mb_opt_t *opt, const HostSrc &src, const int64_t *off, int32_t n_reads, int want, mb_hits_t **out, mb_stats_t *stats);
extern "C" int mb_map_batch(mb_index_t *ix, const mb_opt_t *opt, const uint8_t *cat, const int64_t *off, int32_t n_reads, mb_hits_t **out, mb_stats_t *stats)
{
	HostSrc src; src.ascii = cat;
	return map_batch_impl(ix, opt, src, off, n_reads, 3, out, stats);
}
extern "C" int mb_map_batch_ex(mb_index_t *ix, const mb_opt_t *opt, const uint8_t *cat, const int64_t *off, int32_t n_reads, int want, mb_hits_t **out, mb_stats_t *stats)
{
	HostSrc src; src.ascii = cat;
	return map_batch_impl(ix, opt, src, off, n_reads, want, out, stats);
}
// ---- packed reads (pack.cuh): packed once on the host, a quarter of the bytes over PCIe ----
extern "C" int mb_reads_pack(const uint8_t *cat, const int64_t *off, int32_t n_reads, int n_threads, mb_packed_t **out)
{
	API_BEGIN
	if (!out) throw mb_error(MB_ERR_ARG, "bad arguments");
	*out = pk::pack(cat, off, n_reads, n_threads);
	API_END
}
extern "C" void mb_packed_free(mb_packed_t *p) { delete p; }
extern "C" int64_t mb_packed_upload_bytes(const mb_packed_t *p)
{
	return p ? (int64_t)(((p->total + 15) >> 4) + 2) * 4 + (int64_t)p->iv.size() * 8 + (int64_t)(p->n_reads + 1) * 8 : 0;
}
extern "C" const uint32_t *mb_packed_words(const mb_packed_t *p, int64_t *n_words, const int64_t **intervals, int64_t *n_intervals)
{
	if (!p) return nullptr;
	if (n_words) *n_words = (int64_t)p->n_words;
	if (intervals) *intervals = p->iv.data();
	if (n_intervals) *n_intervals = (int64_t)p->iv.size() / 2;
	return p->words;
}
extern "C" int mb_map_packed(mb_index_t *ix, const mb_opt_t *opt, const mb_packed_t *reads, int want, mb_hits_t **out, mb_stats_t *stats)
{
	if (!reads) { g_mb_err = "bad arguments"; return MB_ERR_ARG; }
	HostSrc src; src.pk = reads;
	return map_batch_impl(ix, opt, src, reads->off.data(), reads->n_reads, want, out, stats);
}
static int map_batch_impl(mb_index_t *ix, const mb_opt_t *opt, const HostSrc &src, const int64_t *off, int32_t n_reads, int want, mb_hits_t **out, mb_stats_t *stats)
{
	API_BEGIN
	if (!ix || !opt || !off || !out || (n_reads > 0 && !src.any() && off[n_reads] > 0)) throw mb_error(MB_ERR_ARG, "bad arguments");
	ThreadCtx &c = get_ctx(ix->device);
	for (;;) try {
	if (n_reads > 1 && off[n_reads] > piece_now(c, ix)) {
		for (int i = 0; i < n_reads; ++i) {
			if (off[i + 1] < off[i] || off[0] != 0) throw mb_error(MB_ERR_ARG, "offsets must start at 0 and be non-decreasing");
			if (off[i + 1] - off[i] > 0x3fffffff) throw mb_error(MB_ERR_ARG, "read longer than 2^30");
		}
		*out = map_in_pieces(ix, *opt, c, src, nullptr, nullptr, off, n_reads, want, stats, piece_now(c, ix));
		return MB_OK;
	}
	c.ar.reset();
	cold_reserve(c, n_reads > 0 ? off[n_reads] : 0);
	uint8_t *d_codes; int64_t *d_off; int64_t total;
	Timer tm(c.st); tm.start();
	SketchFeed feed;
	// worth overlapping from ~64 MB of reads (measured: smaller batches lose); MB_FEED_MIN_BYTES overrides the threshold (tests)
	const char *feed_env = getenv("MB_FEED_MIN_BYTES");
	const int64_t feed_min = feed_env ? atoll(feed_env) : ((int64_t)64 << 20);
	const bool use_feed = n_reads > 0 && off[n_reads] >= feed_min;
	upload_reads(c, src, 0, off, n_reads, &d_codes, &d_off, &total, false, use_feed ? &feed : nullptr);
	float ms_h2d = tm.stop();
	if (stats) stats->ms_h2d = ms_h2d;
	*out = map_device(ix, *opt, c, d_codes, d_off, off, n_reads, total, want, stats, use_feed ? &feed : nullptr);
	if (stats) { stats->ms_h2d = ms_h2d; stats->n_launches += 1; }
	return MB_OK;
	} catch (const mb_error &e) { if (!shrink_piece(c, ix, e, off[n_reads], n_reads)) throw; }
	API_END
}

extern "C" int mb_reads_upload(mb_index_t *ix, const uint8_t *cat, const int64_t *off, int32_t n_reads, mb_reads_t **out)
{
	API_BEGIN
	if (!ix || !off || !out) throw mb_error(MB_ERR_ARG, "bad arguments");
	ThreadCtx &c = get_ctx(ix->device);
	c.ar.reset();
	std::unique_ptr<mb_reads> r(new mb_reads());
	r->device = ix->device; r->n_reads = n_reads;
	HostSrc src; src.ascii = cat;
	upload_reads(c, src, 0, off, n_reads, &r->d_codes, &r->d_off, &r->total, true);
	CK(cudaStreamSynchronize(c.st));
	c.ar.reset();
	*out = r.release();
	API_END
}

extern "C" void mb_reads_free(mb_reads_t *r)
{
	if (!r) return;
	cudaSetDevice(r->device);
	cudaFree(r->d_codes); cudaFree(r->d_off);
	delete r;
}

extern "C" int mb_map_resident(mb_index_t *ix, const mb_opt_t *opt, mb_reads_t *reads, int want_hits, mb_hits_t **out, mb_stats_t *stats)
{
	API_BEGIN
	if (!ix || !opt || !reads || !out) throw mb_error(MB_ERR_ARG, "bad arguments");
	ThreadCtx &c = get_ctx(ix->device);
	c.ar.reset();
	std::vector<int64_t> h_off(reads->n_reads + 1);
	CK(cudaMemcpyAsync(h_off.data(), reads->d_off, (reads->n_reads + 1) * 8, cudaMemcpyDeviceToHost, c.st));
	CK(cudaStreamSynchronize(c.st));
	if (stats) stats->ms_h2d = 0;
	for (;;) try {
		if (reads->n_reads > 1 && reads->total > piece_now(c, ix)) {
			*out = map_in_pieces(ix, *opt, c, HostSrc(), reads->d_codes, reads->d_off, h_off.data(), reads->n_reads, want_hits ? 3 : 0, stats, piece_now(c, ix));
			return MB_OK;
		}
		c.ar.reset();
		cold_reserve(c, reads->total);
		*out = map_device(ix, *opt, c, reads->d_codes, reads->d_off, h_off.data(), reads->n_reads, reads->total, want_hits ? 3 : 0, stats);
		return MB_OK;
	} catch (const mb_error &e) { if (!shrink_piece(c, ix, e, reads->total, reads->n_reads)) throw; }
	API_END
}

// ---------------------------------------------------------------------------------------------
// counting
// ---------------------------------------------------------------------------------------------
extern "C" int mb_count(mb_index_t *ix, const mb_hits_t *h, int32_t mapq_min, int mode, int64_t *counts, int64_t *n_class, int8_t *read_class, int64_t *read_best)
{
	API_BEGIN
	if (!ix || !h) throw mb_error(MB_ERR_ARG, "bad arguments");
	if ((int64_t)h->fields.size() != (int64_t)HIT_NF * h->n) throw mb_error(MB_ERR_ARG, "hits were produced with want_hits=0");
	ThreadCtx &c = get_ctx(ix->device);
	cudaStream_t st = c.st;
	// (no arena reset here: the scratch below is appended behind the last batch, whose device-resident hits mb_count_last
	// may still be asked for; the next mapping call resets the arena)
	const int n_seq = (int)ix->names.size(), n_reads = h->n_reads;
	if (c.n_counts < n_seq + 4) {
		if (c.d_counts) cudaFree(c.d_counts);
		CK(cudaMalloc(&c.d_counts, (size_t)(n_seq + 4) * 8));
		c.n_counts = n_seq + 4;
	}
	CK(cudaMemsetAsync(c.d_counts, 0, (size_t)(n_seq + 4) * 8, st));
	int32_t *d_fields = c.ar.get<int32_t>((size_t)HIT_NF * h->n + 1);
	int64_t *d_hit_off = c.ar.get<int64_t>(n_reads + 1), *d_read_off = c.ar.get<int64_t>(n_reads + 1);
	int8_t *d_cls = c.ar.get<int8_t>(n_reads + 1);
	int64_t *d_best = c.ar.get<int64_t>(n_reads + 1);
	if (h->n) CK(cudaMemcpyAsync(d_fields, h->fields.data(), (size_t)HIT_NF * h->n * 4, cudaMemcpyHostToDevice, st));
	CK(cudaMemcpyAsync(d_hit_off, h->hit_off.data(), (n_reads + 1) * 8, cudaMemcpyHostToDevice, st));
	CK(cudaMemcpyAsync(d_read_off, h->read_off.data(), (n_reads + 1) * 8, cudaMemcpyHostToDevice, st));
	if (n_reads) k_count<<<(unsigned)cdiv(n_reads, 128), 128, 0, st>>>(n_reads, d_hit_off, h->n, d_fields, d_read_off, mapq_min, mode,
		c.d_counts, c.d_counts + n_seq, d_cls, d_best);
	if (counts) CK(cudaMemcpyAsync(counts, c.d_counts, (size_t)n_seq * 8, cudaMemcpyDeviceToHost, st));
	if (n_class) CK(cudaMemcpyAsync(n_class, c.d_counts + n_seq, 3 * 8, cudaMemcpyDeviceToHost, st));
	if (read_class && n_reads) CK(cudaMemcpyAsync(read_class, d_cls, n_reads, cudaMemcpyDeviceToHost, st));
	if (read_best && n_reads) CK(cudaMemcpyAsync(read_best, d_best, (size_t)n_reads * 8, cudaMemcpyDeviceToHost, st));
	CK(cudaStreamSynchronize(st));
	CK(cudaGetLastError());
	API_END
}

extern "C" int mb_count_last(mb_index_t *ix, int32_t mapq_min, int mode, int64_t *counts, int64_t *n_class)
{
	API_BEGIN
	if (!ix) throw mb_error(MB_ERR_ARG, "bad arguments");
	ThreadCtx &c = get_ctx(ix->device);
	if (c.last_index != ix->uid || c.last_n_reads < 0) throw mb_error(MB_ERR_ARG, "no mapped batch of this index on this thread");
	cudaStream_t st = c.st;
	const int n_seq = (int)ix->names.size();
	if (c.n_counts < n_seq + 4) {
		if (c.d_counts) cudaFree(c.d_counts);
		CK(cudaMalloc(&c.d_counts, (size_t)(n_seq + 4) * 8));
		c.n_counts = n_seq + 4;
	}
	CK(cudaMemsetAsync(c.d_counts, 0, (size_t)(n_seq + 4) * 8, st));
	for (const ThreadCtx::LastPart &lp : c.last_parts)
		if (lp.n_reads) k_count<<<(unsigned)cdiv(lp.n_reads, 128), 128, 0, st>>>(lp.n_reads, lp.hit_off, lp.n_hits, lp.fields, lp.read_off, mapq_min, mode,
			c.d_counts, c.d_counts + n_seq, nullptr, nullptr);
	if (counts) CK(cudaMemcpyAsync(counts, c.d_counts, (size_t)n_seq * 8, cudaMemcpyDeviceToHost, st));
	if (n_class) CK(cudaMemcpyAsync(n_class, c.d_counts + n_seq, 3 * 8, cudaMemcpyDeviceToHost, st));
	CK(cudaStreamSynchronize(st));
	CK(cudaGetLastError());
	API_END
}

// sustained INT32 issue rate: 8 independent add/max chains per thread (IADD3 + IMNMX alternate), all SMs busy
__global__ void k_int_peak(int *out, int iters, int seed)
{
	int a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
	const int b = seed | 1, m = seed + 12345;
	#pragma unroll 1
	for (int i = 0; i < iters; ++i) {
		#pragma unroll
		for (int k = 0; k < 8; ++k) {
			a0 = max(a0 + b, m); a1 = max(a1 + b, m); a2 = max(a2 + b, m); a3 = max(a3 + b, m);
			a4 = max(a4 + b, m); a5 = max(a5 + b, m); a6 = max(a6 + b, m); a7 = max(a7 + b, m);
		}
	}
	if ((a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7) == 0x7fffffff) out[0] = a0;
}

extern "C" int mb_int_peak(int device, double *tops)
{
	API_BEGIN
	ThreadCtx &c = get_ctx(device);
	int *d; CK(cudaMalloc(&d, 64));
	const int iters = 4096, grid = c.num_sms * 8, tpb = 256;
	k_int_peak<<<grid, tpb, 0, c.st>>>(d, 64, 1);
	Timer tm(c.st);
	float best = 1e30f;
	for (int rep = 0; rep < 5; ++rep) {
		tm.start();
		k_int_peak<<<grid, tpb, 0, c.st>>>(d, iters, rep + 2);
		float ms = tm.stop();
		if (ms < best) best = ms;
	}
	CK(cudaGetLastError());
	cudaFree(d);
	// 2 integer ops (add, max) x 8 chains x 8 unrolled x iters per thread
	double ops = 2.0 * 8 * 8 * iters * (double)grid * tpb;
	if (tops) *tops = ops / (best * 1e-3) / 1e12;
	API_END
}

// mm_set_mapq's logf (glue.cuh: mb_logf) against the caller's libm values, bit for bit, over a run of consecutive float bit
// patterns: device >= 0 evaluates on that GPU (the code path k_finish runs), device < 0 on the host.
__global__ void k_logf_sweep(uint32_t first_bits, int64_t n, const uint32_t *__restrict__ want, unsigned long long *__restrict__ n_bad, uint32_t *__restrict__ first_bad)
{
	unsigned long long bad = 0;
	uint32_t fb = 0xffffffffu;
	for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
		const uint32_t b = first_bits + (uint32_t)i;
		if (__float_as_uint(mb_logf(__uint_as_float(b))) != want[i]) { ++bad; if (b < fb) fb = b; }
	}
	if (bad) { atomicAdd(n_bad, bad); atomicMin(first_bad, fb); }
}

extern "C" int mb_logf_sweep(int device, uint32_t first_bits, int64_t n, const float *expected, int64_t *n_bad, uint32_t *first_bad)
{
	API_BEGIN
	if (n < 0 || (n && !expected) || (uint64_t)first_bits + (uint64_t)n > 0x100000000ull) throw mb_error(MB_ERR_ARG, "bad arguments");
	const uint32_t *want = reinterpret_cast<const uint32_t*>(expected);
	unsigned long long bad = 0;
	uint32_t fb = 0xffffffffu;
	if (device < 0) {
		for (int64_t i = 0; i < n; ++i) {
			const uint32_t b = first_bits + (uint32_t)i;
			float x, y; uint32_t yb;
			memcpy(&x, &b, 4); y = mb_logf(x); memcpy(&yb, &y, 4);
			if (yb != want[i]) { ++bad; if (b < fb) fb = b; }
		}
	} else if (n) {
		ThreadCtx &c = get_ctx(device);
		uint32_t *d_want = nullptr; unsigned long long *d_bad = nullptr;
		CK(cudaMalloc(&d_want, (size_t)n * 4));
		if (cudaMalloc(&d_bad, 16) != cudaSuccess) { cudaFree(d_want); throw mb_error(MB_ERR_NOMEM, "out of device memory"); }
		const unsigned long long init[2] = { 0ull, 0xffffffffull };
		cudaError_t e = cudaMemcpyAsync(d_bad, init, 16, cudaMemcpyHostToDevice, c.st);
		if (e == cudaSuccess) e = cudaMemcpyAsync(d_want, want, (size_t)n * 4, cudaMemcpyHostToDevice, c.st);
		if (e == cudaSuccess) {
			k_logf_sweep<<<c.num_sms * 8, 256, 0, c.st>>>(first_bits, n, d_want, d_bad, reinterpret_cast<uint32_t*>(d_bad + 1));
			e = cudaGetLastError();
		}
		unsigned long long out[2] = { 0, 0 };
		if (e == cudaSuccess) e = cudaMemcpyAsync(out, d_bad, 16, cudaMemcpyDeviceToHost, c.st);
		if (e == cudaSuccess) e = cudaStreamSynchronize(c.st);
		cudaFree(d_want); cudaFree(d_bad);
		CK(e);
		bad = out[0], fb = (uint32_t)out[1];
	}
	if (n_bad) *n_bad = (int64_t)bad;
	if (first_bad) *first_bad = fb;
	API_END
}

extern "C" void *mb_stream(mb_index_t *ix)
{
	if (!ix) return nullptr;
	try { return (void*)get_ctx(ix->device).st; } catch (...) { return nullptr; }
}

extern "C" void *mb_count_device_ptr(mb_index_t *ix)
{
	if (!ix) return nullptr;
	auto it = t_ctx.find(ix->device);
	return it == t_ctx.end() ? nullptr : (void*)it->second->d_counts;
}

extern "C" int mb_count_fetch(mb_index_t *ix, int64_t *counts)
{
	API_BEGIN
	if (!ix || !counts) throw mb_error(MB_ERR_ARG, "bad arguments");
	ThreadCtx &c = get_ctx(ix->device);
	if (!c.d_counts) throw mb_error(MB_ERR_ARG, "no count vector on this thread");
	CK(cudaMemcpy(counts, c.d_counts, ix->names.size() * 8, cudaMemcpyDeviceToHost));
	API_END
}

// ---- the one collective: NCCL all-reduce of the count vector ----
// libnccl is bound at run time so that the library loads on hosts without it (index building, CPU-side FASTQ tools).
namespace {
struct NcclId { char internal[128]; };
typedef void *NcclComm;
struct NcclApi {
	void *h = nullptr;
	int (*GetUniqueId)(NcclId*) = nullptr;
	int (*CommInitRank)(NcclComm*, int, NcclId, int) = nullptr;
	int (*AllReduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
	int (*CommDestroy)(NcclComm) = nullptr;
	const char *(*GetErrorString)(int) = nullptr;
};
NcclApi &nccl_api()
{
	static NcclApi api;
	static std::once_flag once;
	std::call_once(once, []() {
		const char *names[] = { "libnccl.so.2", "libnccl.so" };
		for (const char *n : names) { api.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (api.h) break; }
		if (!api.h) return;
		api.GetUniqueId = (int (*)(NcclId*))dlsym(api.h, "ncclGetUniqueId");
		api.CommInitRank = (int (*)(NcclComm*, int, NcclId, int))dlsym(api.h, "ncclCommInitRank");
		api.AllReduce = (int (*)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t))dlsym(api.h, "ncclAllReduce");
		api.CommDestroy = (int (*)(NcclComm))dlsym(api.h, "ncclCommDestroy");
		api.GetErrorString = (const char *(*)(int))dlsym(api.h, "ncclGetErrorString");
	});
	if (!api.h || !api.GetUniqueId || !api.CommInitRank || !api.AllReduce || !api.CommDestroy)
		throw mb_error(MB_ERR_CUDA, "libnccl.so.2 not found (or incomplete): the multi-GPU all-reduce has no fallback");
	return api;
}
void nccl_check(int rc, const char *what)
{
	if (rc == 0) return;
	NcclApi &a = nccl_api();
	throw mb_error(MB_ERR_CUDA, std::string(what) + ": NCCL error " + std::to_string(rc) + (a.GetErrorString ? std::string(" (") + a.GetErrorString(rc) + ")" : std::string()));
}
}
struct mb_comm { NcclComm comm = nullptr; int device = 0, rank = 0, world = 1; };

extern "C" int mb_comm_unique_id(uint8_t id[128])
{
	API_BEGIN
	if (!id) throw mb_error(MB_ERR_ARG, "bad arguments");
	NcclId nid;
	nccl_check(nccl_api().GetUniqueId(&nid), "ncclGetUniqueId");
	memcpy(id, nid.internal, 128);
	API_END
}

extern "C" int mb_comm_init(int device, int rank, int world, const uint8_t id[128], mb_comm_t **out)
{
	API_BEGIN
	if (!id || !out || world < 1 || rank < 0 || rank >= world) throw mb_error(MB_ERR_ARG, "bad arguments");
	ensure_device(device);
	std::unique_ptr<mb_comm> c(new mb_comm());
	c->device = device, c->rank = rank, c->world = world;
	NcclId nid; memcpy(nid.internal, id, 128);
	nccl_check(nccl_api().CommInitRank(&c->comm, world, nid, rank), "ncclCommInitRank");
	*out = c.release();
	API_END
}

extern "C" void mb_comm_free(mb_comm_t *comm)
{
	if (!comm) return;
	try { if (comm->comm) { cudaSetDevice(comm->device); nccl_api().CommDestroy(comm->comm); } } catch (...) {}
	delete comm;
}

extern "C" int mb_allreduce_counts(mb_index_t *ix, mb_comm_t *comm, int64_t *counts, int64_t *n_class)
{
	API_BEGIN
	if (!ix || !comm) throw mb_error(MB_ERR_ARG, "bad arguments");
	ThreadCtx &c = get_ctx(ix->device);
	if (!c.d_counts) throw mb_error(MB_ERR_ARG, "no count vector on this thread (call mb_count / mb_count_last first)");
	const int n_seq = (int)ix->names.size();
	// counts[n_seq] and the three read-class counters are contiguous: one collective
	nccl_check(nccl_api().AllReduce(c.d_counts, c.d_counts, (size_t)n_seq + 3, /*ncclInt64*/ 4, /*ncclSum*/ 0, comm->comm, c.st), "ncclAllReduce");
	if (counts) CK(cudaMemcpyAsync(counts, c.d_counts, (size_t)n_seq * 8, cudaMemcpyDeviceToHost, c.st));
	if (n_class) CK(cudaMemcpyAsync(n_class, c.d_counts + n_seq, 3 * 8, cudaMemcpyDeviceToHost, c.st));
	CK(cudaStreamSynchronize(c.st));
	CK(cudaGetLastError());
	API_END
}

// ---- normaliser on the device-resident count vector: monica/genomes/aligner.py:305-319 ----
// counts are per contig; contigs of one genome share a group (accession).  BPB = group count / genome length, sample total =
// sum of BPB in the caller's order (the reference adds floats in dict order, so the order is part of the result), BPM =
// BPB / total.
__global__ void k_norm_group(const unsigned long long *__restrict__ counts, const int32_t *__restrict__ group, int n_seq, unsigned long long *__restrict__ gsum)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n_seq && group[i] >= 0 && counts[i]) atomicAdd(&gsum[group[i]], counts[i]);
}
__global__ void k_norm_bpb(const unsigned long long *__restrict__ gsum, const double *__restrict__ glen, int n_groups, double *__restrict__ bpb)
{
	const int g = blockIdx.x * blockDim.x + threadIdx.x;
	if (g < n_groups) bpb[g] = __ddiv_rn((double)(long long)gsum[g], glen[g]);
}
__global__ void k_norm_total(const double *__restrict__ bpb, const int32_t *__restrict__ order, int n_order, double *__restrict__ total)
{
	if (blockIdx.x || threadIdx.x) return;
	double t = 0.0;
	for (int i = 0; i < n_order; ++i) t = __dadd_rn(t, bpb[order[i]]); // sequential on purpose: float addition order is the reference's
	*total = t;
}
__global__ void k_norm_bpm(double *__restrict__ bpb, const double *__restrict__ total, int n_groups)
{
	const int g = blockIdx.x * blockDim.x + threadIdx.x;
	if (g < n_groups) bpb[g] = __ddiv_rn(bpb[g], *total);
}

extern "C" int mb_normalize_last(mb_index_t *ix, const int32_t *group, int32_t n_groups, const double *group_len,
                                 const int32_t *order, int32_t n_order, double *bpm)
{
	API_BEGIN
	if (!ix || !group || n_groups <= 0 || !group_len || (n_order && !order) || n_order < 0 || !bpm) throw mb_error(MB_ERR_ARG, "bad arguments");
	ThreadCtx &c = get_ctx(ix->device);
	if (!c.d_counts) throw mb_error(MB_ERR_ARG, "no count vector on this thread");
	const int n_seq = (int)ix->names.size();
	for (int i = 0; i < n_seq; ++i) if (group[i] >= n_groups) throw mb_error(MB_ERR_ARG, "group id out of range");
	for (int i = 0; i < n_order; ++i) if (order[i] < 0 || order[i] >= n_groups) throw mb_error(MB_ERR_ARG, "order entry out of range");
	cudaStream_t st = c.st;
	int32_t *d_group = nullptr, *d_order = nullptr; double *d_len = nullptr, *d_bpb = nullptr; unsigned long long *d_gsum = nullptr;
	CK(cudaMallocAsync(&d_group, (size_t)n_seq * 4, st));
	CK(cudaMallocAsync(&d_order, (size_t)(n_order + 1) * 4, st));
	CK(cudaMallocAsync(&d_len, (size_t)n_groups * 8, st));
	CK(cudaMallocAsync(&d_bpb, (size_t)(n_groups + 1) * 8, st));
	CK(cudaMallocAsync(&d_gsum, (size_t)n_groups * 8, st));
	CK(cudaMemcpyAsync(d_group, group, (size_t)n_seq * 4, cudaMemcpyHostToDevice, st));
	if (n_order) CK(cudaMemcpyAsync(d_order, order, (size_t)n_order * 4, cudaMemcpyHostToDevice, st));
	CK(cudaMemcpyAsync(d_len, group_len, (size_t)n_groups * 8, cudaMemcpyHostToDevice, st));
	CK(cudaMemsetAsync(d_gsum, 0, (size_t)n_groups * 8, st));
	k_norm_group<<<(unsigned)cdiv(n_seq, 256), 256, 0, st>>>(c.d_counts, d_group, n_seq, d_gsum);
	k_norm_bpb<<<(unsigned)cdiv(n_groups, 256), 256, 0, st>>>(d_gsum, d_len, n_groups, d_bpb);
	k_norm_total<<<1, 32, 0, st>>>(d_bpb, d_order, n_order, d_bpb + n_groups);
	k_norm_bpm<<<(unsigned)cdiv(n_groups, 256), 256, 0, st>>>(d_bpb, d_bpb + n_groups, n_groups);
	CK(cudaGetLastError());
	CK(cudaMemcpyAsync(bpm, d_bpb, (size_t)n_groups * 8, cudaMemcpyDeviceToHost, st));
	cudaFreeAsync(d_group, st); cudaFreeAsync(d_order, st); cudaFreeAsync(d_len, st); cudaFreeAsync(d_bpb, st); cudaFreeAsync(d_gsum, st);
	CK(cudaStreamSynchronize(st));
	API_END
}

// ---------------------------------------------------------------------------------------------
// per-stage entry points (parity tests)
// ---------------------------------------------------------------------------------------------
extern "C" int mb_sketch(int device, const uint8_t *cat, const int64_t *off, int32_t n_reads, int w, int k, uint64_t *out_xy, int64_t cap, int64_t *out_off)
{
	API_BEGIN
	ThreadCtx &c = get_ctx(device);
	c.ar.reset();
	uint8_t *d_codes; int64_t *d_off; int64_t total; int64_t nl = 0;
	HostSrc src; src.ascii = cat;
	upload_reads(c, src, 0, off, n_reads, &d_codes, &d_off, &total, false);
	SketchOut so;
	run_sketch(c.ar, c.st, d_codes, d_off, n_reads, total, w, k, so, &nl);
	if (so.n_mini > cap) throw mb_error(MB_ERR_OVERFLOW, "output capacity too small");
	if (so.n_mini) CK(cudaMemcpyAsync(out_xy, so.mini, so.n_mini * 16, cudaMemcpyDeviceToHost, c.st));
	CK(cudaMemcpyAsync(out_off, so.mini_off, (n_reads + 1) * 8, cudaMemcpyDeviceToHost, c.st));
	CK(cudaStreamSynchronize(c.st));
	CK(cudaGetLastError());
	API_END
}

extern "C" int mb_seed(mb_index_t *ix, const mb_opt_t *opt, const uint8_t *cat, const int64_t *off, int32_t n_reads,
                       uint64_t *out_xy, int64_t cap, int64_t *out_off, int32_t *rep_len)
{
	API_BEGIN
	if (!ix || !opt) throw mb_error(MB_ERR_ARG, "bad arguments");
	ThreadCtx &c = get_ctx(ix->device);
	c.ar.reset();
	uint8_t *d_codes; int64_t *d_off; int64_t total; int64_t nl = 0;
	HostSrc src; src.ascii = cat;
	upload_reads(c, src, 0, off, n_reads, &d_codes, &d_off, &total, false);
	SketchOut so;
	run_sketch(c.ar, c.st, d_codes, d_off, n_reads, total, ix->w, ix->k, so, &nl);
	SeedOut sd;
	run_seed(c.ar, c.st, ix->d, opt->mid_occ > 0 ? opt->mid_occ : ix->mid_occ, so.mini, so.mini_off, so.n_mini, d_off, n_reads, sd, &nl, c.num_sms);
	if (sd.n_a > cap) throw mb_error(MB_ERR_OVERFLOW, "output capacity too small");
	if (sd.n_a) CK(cudaMemcpyAsync(out_xy, sd.a, sd.n_a * 16, cudaMemcpyDeviceToHost, c.st));
	CK(cudaMemcpyAsync(out_off, sd.a_roff, (n_reads + 1) * 8, cudaMemcpyDeviceToHost, c.st));
	if (rep_len && n_reads) CK(cudaMemcpyAsync(rep_len, sd.rep_len, n_reads * 4, cudaMemcpyDeviceToHost, c.st));
	CK(cudaStreamSynchronize(c.st));
	CK(cudaGetLastError());
	API_END
}

extern "C" int mb_chain(int device, const mb_opt_t *opt, const uint64_t *anchors_xy, const int64_t *off, int32_t n_reads,
                        int32_t *f, int32_t *p, int32_t *v, uint64_t *chained_xy, int64_t *chained_off, uint64_t *u, int64_t *u_off)
{
	API_BEGIN
	if (!opt || !off) throw mb_error(MB_ERR_ARG, "bad arguments");
	ThreadCtx &c = get_ctx(device);
	c.ar.reset();
	cudaStream_t st = c.st;
	Arena &ar = c.ar;
	const int64_t n_a = off[n_reads];
	mb128 *d_a = ar.get<mb128>(n_a + 1);
	int64_t *d_roff = ar.get<int64_t>(n_reads + 1);
	if (n_a) CK(cudaMemcpyAsync(d_a, anchors_xy, n_a * 16, cudaMemcpyHostToDevice, st));
	CK(cudaMemcpyAsync(d_roff, off, (n_reads + 1) * 8, cudaMemcpyHostToDevice, st));
	ReadScratch rs;
	rs.f = ar.get<int32_t>(n_a + 1), rs.p = ar.get<int32_t>(n_a + 1), rs.v = ar.get<int32_t>(n_a + 1), rs.t = ar.get<int32_t>(n_a + 1);
	rs.b = ar.get<mb128>(n_a + 1); rs.u = ar.get<uint64_t>(n_a + 1); rs.scr = ar.get<uint64_t>(3 * n_a + 3 * (int64_t)n_reads + 3);
	rs.regs_tmp = nullptr;
	int32_t *wc = ar.get<int32_t>(1); int *d_err = ar.get<int>(1);
	CK(cudaMemsetAsync(wc, 0, sizeof(int32_t), st)); CK(cudaMemsetAsync(d_err, 0, sizeof(int), st));
	const int max_chain_gap_ref = opt->max_gap_ref > 0 ? opt->max_gap_ref : opt->max_gap;
	if (n_a) k_chain_dp<<<c.num_sms * 8, CH_WARPS * 32, 0, st>>>(d_a, d_roff, n_reads, max_chain_gap_ref, opt->max_gap, opt->bw, opt->max_chain_skip, opt->max_chain_iter,
		rs.f, rs.p, rs.v, rs.t, wc, nullptr, nullptr);
	if (f && n_a) CK(cudaMemcpyAsync(f, rs.f, n_a * 4, cudaMemcpyDeviceToHost, st));
	if (p && n_a) CK(cudaMemcpyAsync(p, rs.p, n_a * 4, cudaMemcpyDeviceToHost, st));
	if (v && n_a) CK(cudaMemcpyAsync(v, rs.v, n_a * 4, cudaMemcpyDeviceToHost, st));
	CK(cudaStreamSynchronize(st));
	if (chained_xy) {
		ReadArrays ra; memset(&ra, 0, sizeof(ra));
		ra.a = d_a; ra.a_roff = d_roff;
		int32_t *n_u = ar.get<int32_t>(n_reads + 1);
		int32_t *big = ar.get<int32_t>(n_reads + 1), *big_ctr = ar.get<int32_t>(2);
		CK(cudaMemsetAsync(big_ctr, 0, 2 * sizeof(int32_t), st));
		if (n_reads) {
			k_chain_bt<<<(unsigned)cdiv((int64_t)n_reads * 32, 128), 128, 0, st>>>(ra, rs, n_reads, opt->min_cnt, opt->min_chain_score, n_u, nullptr, big, big_ctr);
			k_chain_bt_big<<<c.num_sms, 32, BIG_SMEM_BYTES, st>>>(ra, rs, opt->min_cnt, opt->min_chain_score, n_u, big, big_ctr, big_ctr + 1);
		}
		std::vector<int32_t> h_nu(n_reads);
		std::vector<mb128> h_a(n_a); std::vector<uint64_t> h_u(n_a);
		if (n_reads) CK(cudaMemcpyAsync(h_nu.data(), n_u, n_reads * 4, cudaMemcpyDeviceToHost, st));
		if (n_a) { CK(cudaMemcpyAsync(h_a.data(), d_a, n_a * 16, cudaMemcpyDeviceToHost, st)); CK(cudaMemcpyAsync(h_u.data(), rs.u, n_a * 8, cudaMemcpyDeviceToHost, st)); }
		CK(cudaStreamSynchronize(st));
		check_err(d_err, st, "chain backtrack");
		int64_t co = 0, uo = 0;
		for (int r = 0; r < n_reads; ++r) {
			chained_off[r] = co, u_off[r] = uo;
			int64_t k2 = 0;
			for (int i = 0; i < h_nu[r]; ++i) { u[uo++] = h_u[off[r] + i]; k2 += (int32_t)h_u[off[r] + i]; }
			for (int64_t i = 0; i < k2; ++i) { chained_xy[2 * co] = h_a[off[r] + i].x; chained_xy[2 * co + 1] = h_a[off[r] + i].y; ++co; }
		}
		chained_off[n_reads] = co, u_off[n_reads] = uo;
	}
	CK(cudaGetLastError());
	API_END
}

__global__ void k_tasks_from_api(const mb_dp_task_t *__restrict__ in, DpTask *__restrict__ out, int64_t n, int64_t cig_base)
{
	int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	DpTask t; memset(&t, 0, sizeof(t));
	t.reg = -1, t.kind = (in[i].flag & MB_EZ_EXTZ_ONLY) ? 2 : 1;
	t.qlen = in[i].qlen, t.tlen = in[i].tlen, t.w = in[i].w, t.zdrop = in[i].zdrop, t.end_bonus = in[i].end_bonus, t.flag = in[i].flag;
	t.q_idx0 = in[i].q_off, t.q_step = 1, t.q_comp = 2;
	t.t_idx0 = in[i].t_off, t.t_step = 1, t.t_packed = 0;
	t.cigar_off = cig_base + in[i].cigar_off;
	t.skip = 0;
	out[i] = t;
}

__global__ void k_tasks_to_api(const DpTask *__restrict__ in, mb_dp_task_t *__restrict__ out, int64_t n)
{
	int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	out[i].score = in[i].score, out[i].max = in[i].max, out[i].max_q = in[i].max_q, out[i].max_t = in[i].max_t;
	out[i].mqe = in[i].mqe, out[i].mqe_t = in[i].mqe_t, out[i].zdropped = in[i].zdropped, out[i].reach_end = in[i].reach_end, out[i].n_cigar = in[i].n_cigar;
}

extern "C" int mb_dp_batch(int device, const mb_opt_t *opt, mb_dp_task_t *tasks, int64_t n_tasks, const uint8_t *seqpool, int64_t n_seqpool,
                           uint32_t *cigar_pool, int64_t n_cigar_pool)
{
	API_BEGIN
	if (!opt || !tasks || n_tasks < 0) throw mb_error(MB_ERR_ARG, "bad arguments");
	ThreadCtx &c = get_ctx(device);
	c.ar.reset();
	cudaStream_t st = c.st; Arena &ar = c.ar;
	int64_t nl = 0;
	mb_dp_task_t *d_in = ar.get<mb_dp_task_t>(n_tasks + 1);
	DpTask *d_t = ar.get<DpTask>(n_tasks + 1);
	uint8_t *d_pool = ar.get<uint8_t>(n_seqpool + 16);
	uint32_t *d_cig = ar.get<uint32_t>(n_cigar_pool + 1);
	if (n_tasks) CK(cudaMemcpyAsync(d_in, tasks, n_tasks * sizeof(mb_dp_task_t), cudaMemcpyHostToDevice, st));
	if (n_seqpool) CK(cudaMemcpyAsync(d_pool, seqpool, n_seqpool, cudaMemcpyHostToDevice, st));
	if (n_tasks) k_tasks_from_api<<<(unsigned)cdiv(n_tasks, 256), 256, 0, st>>>(d_in, d_t, n_tasks, (int64_t)((uintptr_t)d_cig / 4));
	DpRunner runner(c, &nl);
	runner.run(d_t, nullptr, n_tasks, false, nullptr, nullptr, d_pool, nullptr, make_scoring(*opt), nullptr);
	if (n_tasks) k_tasks_to_api<<<(unsigned)cdiv(n_tasks, 256), 256, 0, st>>>(d_t, d_in, n_tasks);
	if (n_tasks) CK(cudaMemcpyAsync(tasks, d_in, n_tasks * sizeof(mb_dp_task_t), cudaMemcpyDeviceToHost, st));
	if (n_cigar_pool) CK(cudaMemcpyAsync(cigar_pool, d_cig, n_cigar_pool * 4, cudaMemcpyDeviceToHost, st));
	CK(cudaStreamSynchronize(st));
	CK(cudaGetLastError());
	API_END
}

__global__ void __launch_bounds__(32)
k_ll_batch(mb_opt_t opt, mb_ll_task_t *__restrict__ tasks, int64_t n, int32_t *__restrict__ cursor, const uint8_t *__restrict__ pool, int *__restrict__ scr_pool)
{
	const int lane = threadIdx.x;
	int *scr = scr_pool + (size_t)blockIdx.x * (4 * LL_MAX_LEN);
	for (;;) {
		int k = 0;
		if (lane == 0) k = atomicAdd(cursor, 1);
		k = __shfl_sync(0xffffffffu, k, 0);
		if (k >= n) break;
		mb_ll_task_t &T = tasks[k];
		const uint8_t *q = pool + T.q_off, *t = pool + T.t_off;
		int qe, te;
		const int sc = mb_ll_warp([&](int col) { return (int)q[col]; }, [&](int row) { return (int)t[row]; }, T.qlen, T.tlen, opt, scr, true, &qe, &te, lane);
		__syncwarp();
		if (lane == 0) T.score = sc, T.qe = qe, T.te = te;
		__syncwarp();
	}
}

extern "C" int mb_ll_batch(int device, const mb_opt_t *opt, mb_ll_task_t *tasks, int64_t n_tasks, const uint8_t *seqpool, int64_t n_seqpool)
{
	API_BEGIN
	if (!opt || !tasks || n_tasks < 0) throw mb_error(MB_ERR_ARG, "bad arguments");
	if (!mb_ll_scoring_ok(*opt)) throw mb_error(MB_ERR_ARG, "scoring outside the range of the local-alignment kernel (needs b <= q + 2e)");
	for (int64_t i = 0; i < n_tasks; ++i)
		if (tasks[i].qlen < 0 || tasks[i].tlen < 0 || tasks[i].qlen > LL_MAX_LEN - 8 || tasks[i].tlen > LL_MAX_LEN - 8) throw mb_error(MB_ERR_ARG, "sequence longer than 5000");
	ThreadCtx &c = get_ctx(device);
	c.ar.reset();
	cudaStream_t st = c.st; Arena &ar = c.ar;
	mb_ll_task_t *d_t = ar.get<mb_ll_task_t>(n_tasks + 1);
	uint8_t *d_pool = ar.get<uint8_t>(n_seqpool + 16);
	const int grid = c.num_sms * 4;
	int *scr = ar.get<int>((size_t)grid * 4 * LL_MAX_LEN);
	int32_t *cur = ar.get<int32_t>(1);
	CK(cudaMemsetAsync(cur, 0, sizeof(int32_t), st));
	if (n_tasks) CK(cudaMemcpyAsync(d_t, tasks, n_tasks * sizeof(mb_ll_task_t), cudaMemcpyHostToDevice, st));
	if (n_seqpool) CK(cudaMemcpyAsync(d_pool, seqpool, n_seqpool, cudaMemcpyHostToDevice, st));
	if (n_tasks) k_ll_batch<<<grid, 32, 0, st>>>(*opt, d_t, n_tasks, cur, d_pool, scr);
	if (n_tasks) CK(cudaMemcpyAsync(tasks, d_t, n_tasks * sizeof(mb_ll_task_t), cudaMemcpyDeviceToHost, st));
	CK(cudaStreamSynchronize(st));
	CK(cudaGetLastError());
	API_END
}

#include "fastq_host.cuh"   // FASTQ ingest and routed writers (host side)

// the reads of a loaded FASTQ file in packed form (what mb_map_packed uploads)
extern "C" int mb_fastq_pack(const mb_fastq_t *fq, int n_threads, mb_packed_t **out)
{
	API_BEGIN
	if (!fq || !out) throw mb_error(MB_ERR_ARG, "bad arguments");
	const int64_t zero = 0;
	const int32_t n = fq->off.empty() ? 0 : (int32_t)(fq->off.size() - 1);
	*out = pk::pack(fq->cat, n ? fq->off.data() : &zero, n, n_threads);
	API_END
}

// ---- database builder: monica/genomes/database.py:52-67 builder() ----
// One database<N>.fna.gz = the genomes of a chunk, every record re-headed "<tax_unit>:<accession>" (what the aligner later
// splits on ':', aligner.py:234).  The reference re-serialises each record through Biopython: title = "<new id> <old title>"
// (the old title's first word differs from the new id; a record whose old title already starts with the new id keeps its
// title), sequence re-wrapped at 60 columns with blanks removed.  Decompressed output is byte-identical to that writer;
// genome_len[i] = bases of genome i (the value the reference stores in current_genomes_length.pkl).
extern "C" int mb_db_build(const char *out_path, int32_t n_genomes, const char *const *genome_paths, const char *const *new_headers,
                           int64_t *genome_len)
{
	API_BEGIN
	if (!out_path || n_genomes < 0 || (n_genomes && (!genome_paths || !new_headers))) throw mb_error(MB_ERR_ARG, "bad arguments");
	gzFile out = gzopen(out_path, "wb9");
	if (!out) throw mb_error(MB_ERR_IO, std::string("cannot write ") + out_path);
	gzbuffer(out, 1 << 20);
	std::string rec, title, seq, line;
	std::vector<char> buf(1 << 20);
	bool ok = true;
	auto put = [&](const std::string &v) { if (!v.empty() && gzwrite(out, v.data(), (unsigned)v.size()) != (int)v.size()) ok = false; };
	for (int32_t g = 0; g < n_genomes && ok; ++g) {
		gzFile in = gzopen(genome_paths[g], "rb");
		if (!in) { gzclose(out); throw mb_error(MB_ERR_IO, std::string("cannot open ") + genome_paths[g]); }
		gzbuffer(in, 1 << 20);
		const std::string id = new_headers[g];
		int64_t glen = 0;
		bool have = false;
		auto flush_record = [&]() {
			if (!have) return;
			// Biopython: description (= old title) kept; "id description" unless the description's first word is the id
			size_t b0 = title.find_first_not_of(" \t\v\f\r"), e0 = b0 == std::string::npos ? b0 : title.find_first_of(" \t\v\f\r", b0);
			const std::string first = b0 == std::string::npos ? std::string() : title.substr(b0, e0 == std::string::npos ? std::string::npos : e0 - b0);
			rec.clear(); rec.push_back('>');
			if (title.empty()) rec += id; else if (first == id) rec += title; else { rec += id; rec.push_back(' '); rec += title; }
			rec.push_back('\n');
			for (size_t i = 0; i < seq.size(); i += 60) { rec.append(seq, i, 60); rec.push_back('\n'); }
			put(rec);
			glen += (int64_t)seq.size();
			have = false;
		};
		auto take_line = [&](std::string &l) {
			if (!l.empty() && l[0] == '>') {
				flush_record();
				size_t e = l.find_last_not_of(" \t\v\f\r\n");
				title = e == std::string::npos || e == 0 ? std::string() : l.substr(1, e);
				seq.clear(); have = true;
			} else if (have) {
				const size_t e = l.find_last_not_of(" \t\v\f\r\n"); // line.rstrip(), then blanks and CRs dropped
				for (size_t i = 0; e != std::string::npos && i <= e; ++i) if (l[i] != ' ' && l[i] != '\r') seq.push_back(l[i]);
			}
		};
		int n; line.clear();
		while ((n = gzread(in, buf.data(), (unsigned)buf.size())) > 0) {
			for (int i = 0; i < n; ++i) {
				if (buf[i] == '\n') { take_line(line); line.clear(); } else line.push_back(buf[i]);
			}
		}
		const bool read_ok = n == 0 && mb_gz_clean_eof(in);
		take_line(line); flush_record();
		gzclose(in);
		if (!read_ok) { gzclose(out); throw mb_error(MB_ERR_IO, std::string("read failed: ") + genome_paths[g]); }
		if (genome_len) genome_len[g] = glen;
	}
	if (gzclose(out) != Z_OK || !ok) throw mb_error(MB_ERR_IO, std::string("write failed: ") + out_path);
	API_END
}

