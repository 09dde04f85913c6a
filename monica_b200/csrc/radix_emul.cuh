// radix_emul.cuh -- exact single-thread emulation of minimap2's radix_sort_128x / radix_sort_64
// (ksort.h KRADIX_SORT_INIT: in-place MSD byte radix with a cycle-leader permutation, insertion sort
// for ranges of <= 64 elements).  That sort is UNSTABLE, and its permutation of equal keys feeds the
// chaining automaton, so wherever keys can tie the device path has to reproduce this permutation, not
// just "a" sorted order.  Parallel sorts are used when keys are distinct (the order is then unique);
// reads whose anchors tie on x are re-sorted with this routine by one thread.
//
// Iterative (explicit range stack) so it needs no device recursion; `ws` is caller-provided scratch of
// MB_RS_WS_INTS 32-bit words.
#pragma once
#include "common.cuh"

#define MB_RS_MIN_SIZE 64
#define MB_RS_STACK    2304
#define MB_RS_WS_INTS  (512 + 3 * MB_RS_STACK)

struct KeyX   { MB_HD uint64_t operator()(const mb128 &a) const { return a.x; } };
struct KeyU64 { MB_HD uint64_t operator()(const uint64_t &a) const { return a; } };

template <typename T, typename K>
MB_HD void mb_insertsort(T *beg, T *end, K key)
{
	for (T *i = beg + 1; i < end; ++i)
		if (key(*i) < key(*(i - 1))) {
			T *j, tmp = *i;
			for (j = i; j > beg && key(tmp) < key(*(j - 1)); --j)
				*j = *(j - 1);
			*j = tmp;
		}
}

// one partition pass over [beg,end) on byte `s>>3`; fills bb[256]/be[256] with the final bucket ranges (as indices)
template <typename T, typename K>
MB_HD void mb_rs_partition(T *a, int beg, int end, int s, int *bb, int *be, K key)
{
	for (int k = 0; k < 256; ++k) bb[k] = be[k] = beg;
	for (int i = beg; i != end; ++i) ++be[key(a[i]) >> s & 255];
	for (int k = 1; k < 256; ++k)
		be[k] += be[k - 1] - beg, bb[k] = be[k - 1];
	for (int k = 0; k < 256;) {
		if (bb[k] != be[k]) {
			int l = (int)(key(a[bb[k]]) >> s & 255);
			if (l != k) {
				T tmp = a[bb[k]], swap;
				do {
					swap = tmp; tmp = a[bb[l]]; a[bb[l]++] = swap;
					l = (int)(key(tmp) >> s & 255);
				} while (l != k);
				a[bb[k]++] = tmp;
			} else ++bb[k];
		} else ++k;
	}
	bb[0] = beg;
	for (int k = 1; k < 256; ++k) bb[k] = be[k - 1];
}

template <typename T, typename K>
MB_HD void mb_radix_sort_emul(T *a, int n, int *ws, K key)
{
	if (n <= MB_RS_MIN_SIZE) { mb_insertsort(a, a + n, key); return; }
	int *bb = ws, *be = ws + 256, *stk = ws + 512;
	int sp = 0;
	stk[0] = 0, stk[1] = n, stk[2] = 56; sp = 1;
	while (sp > 0) {
		--sp;
		int beg = stk[3 * sp], end = stk[3 * sp + 1], s = stk[3 * sp + 2];
		mb_rs_partition(a, beg, end, s, bb, be, key);
		if (s) {
			int s2 = s > 8 ? s - 8 : 0;
			for (int k = 0; k < 256; ++k) {
				int len = be[k] - bb[k];
				if (len > MB_RS_MIN_SIZE) {
					if (sp < MB_RS_STACK) { stk[3 * sp] = bb[k], stk[3 * sp + 1] = be[k], stk[3 * sp + 2] = s2; ++sp; }
				} else if (len > 1) mb_insertsort(a + bb[k], a + be[k], key);
			}
		}
	}
}
