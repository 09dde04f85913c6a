// pinpool.cuh -- a small cache of page-locked host buffers shared by the FASTQ loader and the read packer.
#pragma once
#include <mutex>
#include "common.cuh"

// A few page-locked read buffers are kept between files (the aligner loads FASTQ after FASTQ: page-locking ~130 MB costs tens of
// milliseconds per file, and more while the device is busy).
namespace fqh {
struct PinCache { std::mutex mu; void *p[4] = {}; size_t cap[4] = {}; };   // (buffers still cached at exit are left to the OS)
static PinCache g_pin_cache;
static void *pin_take(size_t n, size_t *cap)
{
	{
		std::lock_guard<std::mutex> g(g_pin_cache.mu);
		int best = -1;
		for (int i = 0; i < 4; ++i) if (g_pin_cache.p[i] && g_pin_cache.cap[i] >= n && (best < 0 || g_pin_cache.cap[i] < g_pin_cache.cap[best])) best = i;
		if (best >= 0) { void *p = g_pin_cache.p[best]; *cap = g_pin_cache.cap[best]; g_pin_cache.p[best] = nullptr; g_pin_cache.cap[best] = 0; return p; }
	}
	void *p = nullptr;
	const size_t want = n + n / 8 + 4096;
	if (cudaHostAlloc(&p, want, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return nullptr; }
	*cap = want;
	return p;
}
static void pin_give(void *p, size_t cap)
{
	void *drop = p;
	if (cap <= ((size_t)2 << 30)) {
		std::lock_guard<std::mutex> g(g_pin_cache.mu);
		int slot = -1;
		for (int i = 0; i < 4; ++i) if (!g_pin_cache.p[i]) { slot = i; break; }
		if (slot < 0) { // full: replace the smallest one if this one is larger
			int mn = 0;
			for (int i = 1; i < 4; ++i) if (g_pin_cache.cap[i] < g_pin_cache.cap[mn]) mn = i;
			if (g_pin_cache.cap[mn] < cap) { drop = g_pin_cache.p[mn]; slot = mn; }
		} else drop = nullptr;
		if (slot >= 0) { g_pin_cache.p[slot] = p; g_pin_cache.cap[slot] = cap; }
	}
	if (drop) cudaFreeHost(drop);
}
}
