/*
 * mm2o_ksw.c -- CPU ORACLE (test infrastructure): two-piece gap-affine banded DP.
 * Restates minimap2 v2.17 ksw2_extd2_sse.c (ksw_extd2_sse) and ksw2.h (ksw_backtrack,
 * ksw_apply_zdrop, ksw_reset_extz) as SCALAR code with the SSE kernel's exact data layout:
 * 16-lane blocks are computed from the 16-aligned st..en (so out-of-band lanes inside a
 * block carry the same values the vector code would produce and are visible to in-band
 * cells and to the backtrack exactly as upstream), int8 arithmetic wraps, and the
 * s|sf|qr byte arrays are contiguous and zero padded like upstream's single kcalloc.
 * PARITY UNPINNED -- see mm2o.h.  Reached from /root/reference/monica/genomes/aligner.py:193,215
 * via mappy -> mm_map -> mm_align_skeleton -> mm_align_pair.
 */
#include <stdlib.h>
#include <string.h>
#include <assert.h>
#include "mm2o.h"
#ifdef __SSE4_1__
#include <smmintrin.h>
#endif

/* 1 (default): the anti-diagonal core runs 16 int8 lanes per step with SSE4.1, the way upstream's ksw2_extd2_sse.c does,
 * so the CPU baseline has the speed class of real mappy; 0: the scalar statement of the same lanes.  Both produce identical
 * bytes (tests/test_oracle_cpu.py compares them). */
int mm2o_ksw_simd = 1;
void mm2o_ksw_set_simd(int on) { mm2o_ksw_simd = on; }

void mm2o_gen_simple_mat(int m, int8_t *mat, int8_t a, int8_t b, int8_t sc_ambi)
{
	int i, j;
	a = a < 0? -a : a;
	b = b > 0? -b : b;
	sc_ambi = sc_ambi > 0? -sc_ambi : sc_ambi;
	for (i = 0; i < m - 1; ++i) {
		for (j = 0; j < m - 1; ++j)
			mat[i * m + j] = i == j? a : b;
		mat[i * m + m - 1] = sc_ambi;
	}
	for (j = 0; j < m; ++j)
		mat[(m - 1) * m + j] = sc_ambi;
}

static inline void ksw_reset_extz(ksw_extz_t *ez)
{
	ez->max_q = ez->max_t = ez->mqe_t = ez->mte_q = -1;
	ez->max = 0, ez->score = ez->mqe = ez->mte = KSW_NEG_INF;
	ez->n_cigar = 0, ez->zdropped = 0, ez->reach_end = 0;
}

static inline int ksw_apply_zdrop(ksw_extz_t *ez, int is_rot, int32_t H, int a, int b, int zdrop, int8_t e)
{
	int r, t;
	if (is_rot) r = a, t = b;
	else r = a + b, t = a;
	if (H > (int32_t)ez->max) {
		ez->max = H, ez->max_t = t, ez->max_q = r - t;
	} else if (t >= ez->max_t && r - t >= ez->max_q) {
		int tl = t - ez->max_t, ql = (r - t) - ez->max_q, l;
		l = tl > ql? tl - ql : ql - tl;
		if (zdrop >= 0 && (int32_t)ez->max - H > zdrop + l * e) {
			ez->zdropped = 1;
			return 1;
		}
	}
	return 0;
}

static inline uint32_t *ksw_push_cigar(int *n_cigar, int *m_cigar, uint32_t *cigar, uint32_t op, int len)
{
	if (*n_cigar == 0 || op != (cigar[(*n_cigar) - 1]&0xf)) {
		if (*n_cigar == *m_cigar) {
			*m_cigar = *m_cigar? (*m_cigar)<<1 : 4;
			cigar = (uint32_t*)realloc(cigar, (*m_cigar) << 2);
		}
		cigar[(*n_cigar)++] = len<<4 | op;
	} else cigar[(*n_cigar)-1] += len<<4;
	return cigar;
}

static void ksw_backtrack(int is_rev, const uint8_t *p, const int *off, const int *off_end, int n_col, int i0, int j0,
						  int *m_cigar_, int *n_cigar_, uint32_t **cigar_)
{
	int n_cigar = 0, m_cigar = *m_cigar_, i = i0, j = j0, r, state = 0;
	uint32_t *cigar = *cigar_, tmp;
	while (i >= 0 && j >= 0) {
		int force_state = -1;
		r = i + j;
		if (i < off[r]) force_state = 2;
		if (off_end && i > off_end[r]) force_state = 1;
		tmp = force_state < 0? p[(size_t)r * n_col + i - off[r]] : 0;
		if (state == 0) state = tmp & 7;
		else if (!(tmp >> (state + 2) & 1)) state = 0;
		if (state == 0) state = tmp & 7;
		if (force_state >= 0) state = force_state;
		if (state == 0) cigar = ksw_push_cigar(&n_cigar, &m_cigar, cigar, 0, 1), --i, --j;
		else if (state == 1 || state == 3) cigar = ksw_push_cigar(&n_cigar, &m_cigar, cigar, 2, 1), --i;
		else cigar = ksw_push_cigar(&n_cigar, &m_cigar, cigar, 1, 1), --j;
	}
	if (i >= 0) cigar = ksw_push_cigar(&n_cigar, &m_cigar, cigar, 2, i + 1);
	if (j >= 0) cigar = ksw_push_cigar(&n_cigar, &m_cigar, cigar, 1, j + 1);
	if (!is_rev)
		for (i = 0; i < n_cigar>>1; ++i)
			tmp = cigar[i], cigar[i] = cigar[n_cigar-1-i], cigar[n_cigar-1-i] = tmp;
	*m_cigar_ = m_cigar, *n_cigar_ = n_cigar, *cigar_ = cigar;
}

/* number of in-band cells the kernel visits if it runs to the end: sum over anti-diagonals of en0-st0+1 */
int64_t mm2o_ksw_cells(int qlen, int tlen, int w)
{
	int64_t c = 0;
	int r;
	if (qlen <= 0 || tlen <= 0) return 0;
	if (w < 0) w = tlen > qlen? tlen : qlen;
	for (r = 0; r < qlen + tlen - 1; ++r) {
		int st = 0, en = tlen - 1;
		if (st < r - qlen + 1) st = r - qlen + 1;
		if (en > r) en = r;
		if (st < (r-w+1)>>1) st = (r-w+1)>>1;
		if (en > (r+w)>>1) en = (r+w)>>1;
		if (st > en) break;
		c += en - st + 1;
	}
	return c;
}

#define I8(x) ((int8_t)(x))

void mm2o_ksw_extd2(int qlen, const uint8_t *query, int tlen, const uint8_t *target, int8_t m, const int8_t *mat,
					int8_t q, int8_t e, int8_t q2, int8_t e2, int w, int zdrop, int end_bonus, int flag, ksw_extz_t *ez)
{
	int r, t, qe = q + e, qe2, n_col_, *off = 0, *off_end = 0, tlen_, qlen_, last_st, last_en, wl, wr, max_sc, min_sc, long_thres, long_diff;
	int with_cigar = !(flag&KSW_EZ_SCORE_ONLY), approx_max = !!(flag&KSW_EZ_APPROX_MAX);
	int32_t *H = 0, H0 = 0, last_H0_t = 0;
	uint8_t *qr, *sf, *mem, *p = 0;
	int8_t *u, *v, *x, *y, *x2, *y2, *s;
	int8_t sc_mch, sc_mis, sc_N, m1;

	ksw_reset_extz(ez);
	if (m <= 1 || qlen <= 0 || tlen <= 0) return;

	if (q2 + e2 < q + e) t = q, q = q2, q2 = t, t = e, e = e2, e2 = t; /* make sure q+e no larger than q2+e2 */
	qe = q + e, qe2 = q2 + e2;
	sc_mch = mat[0], sc_mis = mat[1];
	sc_N = mat[m*m-1] == 0? -e2 : mat[m*m-1];
	m1 = m - 1;

	if (w < 0) w = tlen > qlen? tlen : qlen;
	wl = wr = w;
	tlen_ = (tlen + 15) / 16;
	n_col_ = qlen < tlen? qlen : tlen;
	n_col_ = ((n_col_ < w + 1? n_col_ : w + 1) + 15) / 16 + 1;
	qlen_ = (qlen + 15) / 16;
	for (t = 1, max_sc = mat[0], min_sc = mat[1]; t < m * m; ++t) {
		max_sc = max_sc > mat[t]? max_sc : mat[t];
		min_sc = min_sc < mat[t]? min_sc : mat[t];
	}
	if (-min_sc > 2 * (q + e)) return;

	long_thres = e != e2? (q2 - q) / (e - e2) - 1 : 0;
	if (q2 + e2 + long_thres * e2 > q + e + long_thres * e)
		++long_thres;
	long_diff = long_thres * (e - e2) - (q2 - q) - e2;

	/* one zeroed block laid out like upstream: u v x y x2 y2 s sf qr (+16 slack) */
	mem = (uint8_t*)calloc((size_t)tlen_ * 8 + qlen_ + 1, 16);
	u = (int8_t*)mem;
	v = u + tlen_ * 16, x = v + tlen_ * 16, y = x + tlen_ * 16, x2 = y + tlen_ * 16, y2 = x2 + tlen_ * 16;
	s = y2 + tlen_ * 16, sf = (uint8_t*)(s + tlen_ * 16), qr = sf + tlen_ * 16;
	memset(u,  -q  - e,  tlen_ * 16);
	memset(v,  -q  - e,  tlen_ * 16);
	memset(x,  -q  - e,  tlen_ * 16);
	memset(y,  -q  - e,  tlen_ * 16);
	memset(x2, -q2 - e2, tlen_ * 16);
	memset(y2, -q2 - e2, tlen_ * 16);
	if (!approx_max) {
		H = (int32_t*)malloc((size_t)tlen_ * 16 * 4);
		for (t = 0; t < tlen_ * 16; ++t) H[t] = KSW_NEG_INF;
	}
	if (with_cigar) {
		p = (uint8_t*)malloc(((size_t)(qlen + tlen - 1) * n_col_ + 1) * 16);
		off = (int*)malloc((qlen + tlen - 1) * sizeof(int) * 2);
		off_end = off + qlen + tlen - 1;
	}

	for (t = 0; t < qlen; ++t) qr[t] = query[qlen - 1 - t];
	memcpy(sf, target, tlen);

	for (r = 0, last_st = last_en = -1; r < qlen + tlen - 1; ++r) {
		int st = 0, en = tlen - 1, st0, en0;
		int8_t x1, x21, v1;
		uint8_t *qrr = qr + (qlen - 1 - r);
		/* find the boundaries */
		if (st < r - qlen + 1) st = r - qlen + 1;
		if (en > r) en = r;
		if (st < (r-wr+1)>>1) st = (r-wr+1)>>1; /* take the ceil */
		if (en > (r+wl)>>1) en = (r+wl)>>1; /* take the floor */
		if (st > en) {
			ez->zdropped = 1;
			break;
		}
		st0 = st, en0 = en;
		st = st / 16 * 16, en = (en + 16) / 16 * 16 - 1;
		/* set boundary conditions */
		if (st > 0) {
			if (st - 1 >= last_st && st - 1 <= last_en) {
				x1 = x[st - 1], x21 = x2[st - 1], v1 = v[st - 1]; /* (r-1,s-1) calculated in the last round */
			} else {
				x1 = -q - e, x21 = -q2 - e2;
				v1 = -q - e;
			}
		} else {
			x1 = -q - e, x21 = -q2 - e2;
			v1 = r == 0? -q - e : r < long_thres? -e : r == long_thres? long_diff : -e2;
		}
		if (en >= r) {
			y[r] = -q - e, y2[r] = -q2 - e2;
			u[r] = r == 0? -q - e : r < long_thres? -e : r == long_thres? long_diff : -e2;
		}
		/* loop fission: set scores first (16-byte unaligned chunks from st0; may overrun into sf like upstream) */
#ifdef __SSE4_1__
		if (mm2o_ksw_simd) {
			const __m128i m1_ = _mm_set1_epi8(m1), mch_ = _mm_set1_epi8(sc_mch), mis_ = _mm_set1_epi8(sc_mis), scn_ = _mm_set1_epi8(sc_N);
			for (t = st0; t <= en0; t += 16) {
				__m128i sq = _mm_loadu_si128((const __m128i*)(sf + t)), sq2 = _mm_loadu_si128((const __m128i*)(qrr + t));
				__m128i isn = _mm_or_si128(_mm_cmpeq_epi8(sq, m1_), _mm_cmpeq_epi8(sq2, m1_));
				__m128i sc = _mm_blendv_epi8(mis_, mch_, _mm_cmpeq_epi8(sq, sq2));
				_mm_storeu_si128((__m128i*)(s + t), _mm_blendv_epi8(sc, scn_, isn));
			}
		} else
#endif
		for (t = st0; t <= en0; t += 16) {
			int l;
			for (l = 0; l < 16; ++l) {
				uint8_t sq = sf[t + l], sq2 = qrr[t + l];
				int8_t sc = (sq == (uint8_t)m1 || sq2 == (uint8_t)m1)? sc_N : sq == sq2? sc_mch : sc_mis;
				s[t + l] = sc; /* vector store happens after both vector loads: per-chunk loads precede stores */
			}
		}
		/* core loop over the 16-aligned block range */
		{
			int8_t x1c = x1, x21c = x21, v1c = v1;
			uint8_t *pr = with_cigar? p + ((size_t)r * n_col_ - st / 16) * 16 : 0;
			if (with_cigar) off[r] = st, off_end[r] = en;
#ifdef __SSE4_1__
			if (mm2o_ksw_simd) {
				const __m128i zero_ = _mm_setzero_si128(), q_ = _mm_set1_epi8(q), q2_ = _mm_set1_epi8(q2), qe_ = _mm_set1_epi8(qe), qe2_ = _mm_set1_epi8(qe2);
				const __m128i mch_ = _mm_set1_epi8(sc_mch), one_ = _mm_set1_epi8(1), two_ = _mm_set1_epi8(2), three_ = _mm_set1_epi8(3), four_ = _mm_set1_epi8(4);
				const __m128i f08 = _mm_set1_epi8(0x08), f10 = _mm_set1_epi8(0x10), f20 = _mm_set1_epi8(0x20), f40 = _mm_set1_epi8(0x40);
				__m128i x1v = _mm_cvtsi32_si128((uint8_t)x1), x21v = _mm_cvtsi32_si128((uint8_t)x21), v1v = _mm_cvtsi32_si128((uint8_t)v1);
				const int right = !!(flag & KSW_EZ_RIGHT);
				for (t = st; t <= en; t += 16) {
					__m128i z = _mm_loadu_si128((const __m128i*)(s + t));
					__m128i xt = _mm_loadu_si128((const __m128i*)(x + t)), vt = _mm_loadu_si128((const __m128i*)(v + t));
					__m128i x2t = _mm_loadu_si128((const __m128i*)(x2 + t)), ut = _mm_loadu_si128((const __m128i*)(u + t));
					__m128i xt1 = _mm_or_si128(_mm_slli_si128(xt, 1), x1v);     x1v = _mm_srli_si128(xt, 15);   /* lane k sees element t+k-1 */
					__m128i vt1 = _mm_or_si128(_mm_slli_si128(vt, 1), v1v);     v1v = _mm_srli_si128(vt, 15);
					__m128i x2t1 = _mm_or_si128(_mm_slli_si128(x2t, 1), x21v);  x21v = _mm_srli_si128(x2t, 15);
					__m128i a = _mm_add_epi8(xt1, vt1), b = _mm_add_epi8(_mm_loadu_si128((const __m128i*)(y + t)), ut);
					__m128i a2 = _mm_add_epi8(x2t1, vt1), b2 = _mm_add_epi8(_mm_loadu_si128((const __m128i*)(y2 + t)), ut);
					__m128i d, tmp;
					if (!right) {
						d = _mm_and_si128(_mm_cmpgt_epi8(a, z), one_);          z = _mm_max_epi8(z, a);
						d = _mm_blendv_epi8(d, two_, _mm_cmpgt_epi8(b, z));      z = _mm_max_epi8(z, b);
						d = _mm_blendv_epi8(d, three_, _mm_cmpgt_epi8(a2, z));   z = _mm_max_epi8(z, a2);
						d = _mm_blendv_epi8(d, four_, _mm_cmpgt_epi8(b2, z));    z = _mm_max_epi8(z, b2);
					} else {
						d = _mm_andnot_si128(_mm_cmpgt_epi8(z, a), one_);        z = _mm_max_epi8(z, a);
						d = _mm_blendv_epi8(two_, d, _mm_cmpgt_epi8(z, b));      z = _mm_max_epi8(z, b);
						d = _mm_blendv_epi8(three_, d, _mm_cmpgt_epi8(z, a2));   z = _mm_max_epi8(z, a2);
						d = _mm_blendv_epi8(four_, d, _mm_cmpgt_epi8(z, b2));    z = _mm_max_epi8(z, b2);
					}
					z = _mm_min_epi8(z, mch_);
					_mm_storeu_si128((__m128i*)(u + t), _mm_sub_epi8(z, vt1));
					_mm_storeu_si128((__m128i*)(v + t), _mm_sub_epi8(z, ut));
					tmp = _mm_sub_epi8(z, q_);  a = _mm_sub_epi8(a, tmp);   b = _mm_sub_epi8(b, tmp);
					tmp = _mm_sub_epi8(z, q2_); a2 = _mm_sub_epi8(a2, tmp); b2 = _mm_sub_epi8(b2, tmp);
					if (!right) {
						tmp = _mm_cmpgt_epi8(a, zero_);  _mm_storeu_si128((__m128i*)(x + t),  _mm_sub_epi8(_mm_and_si128(tmp, a), qe_));   d = _mm_or_si128(d, _mm_and_si128(tmp, f08));
						tmp = _mm_cmpgt_epi8(b, zero_);  _mm_storeu_si128((__m128i*)(y + t),  _mm_sub_epi8(_mm_and_si128(tmp, b), qe_));   d = _mm_or_si128(d, _mm_and_si128(tmp, f10));
						tmp = _mm_cmpgt_epi8(a2, zero_); _mm_storeu_si128((__m128i*)(x2 + t), _mm_sub_epi8(_mm_and_si128(tmp, a2), qe2_)); d = _mm_or_si128(d, _mm_and_si128(tmp, f20));
						tmp = _mm_cmpgt_epi8(b2, zero_); _mm_storeu_si128((__m128i*)(y2 + t), _mm_sub_epi8(_mm_and_si128(tmp, b2), qe2_)); d = _mm_or_si128(d, _mm_and_si128(tmp, f40));
					} else {
						tmp = _mm_cmpgt_epi8(zero_, a);  _mm_storeu_si128((__m128i*)(x + t),  _mm_sub_epi8(_mm_andnot_si128(tmp, a), qe_));   d = _mm_or_si128(d, _mm_andnot_si128(tmp, f08));
						tmp = _mm_cmpgt_epi8(zero_, b);  _mm_storeu_si128((__m128i*)(y + t),  _mm_sub_epi8(_mm_andnot_si128(tmp, b), qe_));   d = _mm_or_si128(d, _mm_andnot_si128(tmp, f10));
						tmp = _mm_cmpgt_epi8(zero_, a2); _mm_storeu_si128((__m128i*)(x2 + t), _mm_sub_epi8(_mm_andnot_si128(tmp, a2), qe2_)); d = _mm_or_si128(d, _mm_andnot_si128(tmp, f20));
						tmp = _mm_cmpgt_epi8(zero_, b2); _mm_storeu_si128((__m128i*)(y2 + t), _mm_sub_epi8(_mm_andnot_si128(tmp, b2), qe2_)); d = _mm_or_si128(d, _mm_andnot_si128(tmp, f40));
					}
					if (with_cigar) _mm_storeu_si128((__m128i*)(pr + t), d);
				}
			} else
#endif
			for (t = st; t <= en; ++t) {
				int8_t z, a, b, a2, b2, xt1, x2t1, vt1, ut, tmp, d;
				z = s[t];
				xt1 = x1c, x1c = x[t];
				vt1 = v1c, v1c = v[t];
				a = I8(xt1 + vt1);
				ut = u[t];
				b = I8(y[t] + ut);
				x2t1 = x21c, x21c = x2[t];
				a2 = I8(x2t1 + vt1);
				b2 = I8(y2[t] + ut);
				if (!(flag&KSW_EZ_RIGHT)) { /* gap left-alignment */
					d = a > z? 1 : 0;
					z = z > a? z : a;
					d = b > z? 2 : d;
					z = z > b? z : b;
					d = a2 > z? 3 : d;
					z = z > a2? z : a2;
					d = b2 > z? 4 : d;
					z = z > b2? z : b2;
					z = z < sc_mch? z : sc_mch;
				} else { /* gap right-alignment */
					d = z > a? 0 : 1;
					z = z > a? z : a;
					d = z > b? d : 2;
					z = z > b? z : b;
					d = z > a2? d : 3;
					z = z > a2? z : a2;
					d = z > b2? d : 4;
					z = z > b2? z : b2;
					z = z < sc_mch? z : sc_mch;
				}
				u[t] = I8(z - vt1);
				v[t] = I8(z - ut);
				tmp = I8(z - q);
				a = I8(a - tmp);
				b = I8(b - tmp);
				tmp = I8(z - q2);
				a2 = I8(a2 - tmp);
				b2 = I8(b2 - tmp);
				if (!(flag&KSW_EZ_RIGHT)) {
					x[t]  = I8((a  > 0? a  : 0) - qe);  if (a  > 0) d |= 0x08;
					y[t]  = I8((b  > 0? b  : 0) - qe);  if (b  > 0) d |= 0x10;
					x2[t] = I8((a2 > 0? a2 : 0) - qe2); if (a2 > 0) d |= 0x20;
					y2[t] = I8((b2 > 0? b2 : 0) - qe2); if (b2 > 0) d |= 0x40;
				} else {
					x[t]  = I8((0 > a?  0 : a)  - qe);  if (!(0 > a))  d |= 0x08;
					y[t]  = I8((0 > b?  0 : b)  - qe);  if (!(0 > b))  d |= 0x10;
					x2[t] = I8((0 > a2? 0 : a2) - qe2); if (!(0 > a2)) d |= 0x20;
					y2[t] = I8((0 > b2? 0 : b2) - qe2); if (!(0 > b2)) d |= 0x40;
				}
				if (with_cigar) pr[t] = (uint8_t)d;
			}
		}
		if (!approx_max) { /* find the exact max with a 32-bit score array */
			int32_t max_H, max_t;
			if (r > 0) {
				int32_t HH[4], tt[4], en1 = st0 + (en0 - st0) / 4 * 4, i;
				max_H = H[en0] = en0 > 0? H[en0-1] + u[en0] : H[en0] + v[en0]; /* special casing the last element */
				max_t = en0;
				for (i = 0; i < 4; ++i) HH[i] = max_H, tt[i] = max_t;
				for (t = st0; t < en1; t += 4) { /* 4-lane vector loop of upstream */
					for (i = 0; i < 4; ++i) {
						H[t + i] += (int32_t)v[t + i];
						if (H[t + i] > HH[i]) HH[i] = H[t + i], tt[i] = t;
					}
				}
				for (i = 0; i < 4; ++i)
					if (max_H < HH[i]) max_H = HH[i], max_t = tt[i] + i;
				for (; t < en0; ++t) { /* the rest */
					H[t] += (int32_t)v[t];
					if (H[t] > max_H)
						max_H = H[t], max_t = t;
				}
			} else H[0] = v[0] - qe, max_H = H[0], max_t = 0; /* special casing r==0 */
			/* update ez */
			if (en0 == tlen - 1 && H[en0] > ez->mte)
				ez->mte = H[en0], ez->mte_q = r - en0;
			if (r - st0 == qlen - 1 && H[st0] > ez->mqe)
				ez->mqe = H[st0], ez->mqe_t = st0;
			if (ksw_apply_zdrop(ez, 1, max_H, r, max_t, zdrop, e2)) break;
			if (r == qlen + tlen - 2 && en0 == tlen - 1)
				ez->score = H[tlen - 1];
		} else { /* find approximate max; Z-drop might be inaccurate, too. */
			if (r > 0) {
				if (last_H0_t >= st0 && last_H0_t <= en0 && last_H0_t + 1 >= st0 && last_H0_t + 1 <= en0) {
					int32_t d0 = v[last_H0_t];
					int32_t d1 = u[last_H0_t + 1];
					if (d0 > d1) H0 += d0;
					else H0 += d1, ++last_H0_t;
				} else if (last_H0_t >= st0 && last_H0_t <= en0) {
					H0 += v[last_H0_t];
				} else {
					++last_H0_t, H0 += u[last_H0_t];
				}
			} else H0 = v[0] - qe, last_H0_t = 0;
			if ((flag & KSW_EZ_APPROX_DROP) && ksw_apply_zdrop(ez, 1, H0, r, last_H0_t, zdrop, e2)) break;
			if (r == qlen + tlen - 2 && en0 == tlen - 1)
				ez->score = H0;
		}
		last_st = st, last_en = en;
	}
	free(mem);
	if (!approx_max) free(H);
	if (with_cigar) { /* backtrack */
		int rev_cigar = !!(flag & KSW_EZ_REV_CIGAR);
		if (!ez->zdropped && !(flag&KSW_EZ_EXTZ_ONLY)) {
			ksw_backtrack(rev_cigar, p, off, off_end, n_col_*16, tlen-1, qlen-1, &ez->m_cigar, &ez->n_cigar, &ez->cigar);
		} else if (!ez->zdropped && (flag&KSW_EZ_EXTZ_ONLY) && ez->mqe + end_bonus > (int)ez->max) {
			ez->reach_end = 1;
			ksw_backtrack(rev_cigar, p, off, off_end, n_col_*16, ez->mqe_t, qlen-1, &ez->m_cigar, &ez->n_cigar, &ez->cigar);
		} else if (ez->max_t >= 0 && ez->max_q >= 0) {
			ksw_backtrack(rev_cigar, p, off, off_end, n_col_*16, ez->max_t, ez->max_q, &ez->m_cigar, &ez->n_cigar, &ez->cigar);
		}
		free(p); free(off);
	}
}

/*
 * ksw2_ll_sse.c ksw_ll_qinit() + ksw_ll_i16(): local (Smith-Waterman) single-affine alignment in Farrar's striped layout,
 * 8 int16 lanes per vector, restated lane by lane as scalar code so that the upstream quirks that decide the END POSITION
 * are kept: the query is padded to slen*8 columns whose substitution score is 0 (a maximum in the last real column is
 * carried diagonally through the padding, so `te` can advance by up to 7 rows and `qe` can point past the query), the row
 * maximum is taken over all lanes before the lazy-F pass, `imax >= gmax` keeps the LAST row that reaches the maximum, and
 * the final scan keeps the LAST memory slot of that row equal to gmax (slot i <-> query column i/8 + i%8*slen).
 * Callers: mm_test_zdrop (score only) and mm_align1_inv (score, qe, te) in align.c.
 */
int mm2o_ksw_ll_i16(int qlen, const uint8_t *query, int tlen, const uint8_t *target, int m, const int8_t *mat, int gapo, int gape, int *qe, int *te)
{
	const int p = 8, slen = (qlen + p - 1) / p, n8 = slen * p;
	int16_t *qp, *H0, *H1, *E, *Hmax, *tmp;
	int a, i, j, l, gmax = 0, gapoe = gapo + gape;
	*qe = *te = -1;
	if (qlen <= 0) return 0;
	qp = (int16_t*)calloc((size_t)n8 * (m + 4), sizeof(int16_t));
	H0 = qp + (size_t)n8 * m, H1 = H0 + n8, E = H1 + n8, Hmax = E + n8;
	for (a = 0; a < m; ++a) { /* query profile: vector j, lane l holds column j + l*slen */
		int16_t *t = qp + (size_t)a * n8;
		const int8_t *ma = mat + a * m;
		for (i = 0; i < slen; ++i)
			for (j = i; j < n8; j += slen)
				*t++ = (int16_t)(j >= qlen? 0 : ma[query[j]]);
	}
#define SUBS_U16(x, y) ((uint16_t)(x) > (uint16_t)(y)? (int16_t)((uint16_t)(x) - (uint16_t)(y)) : (int16_t)0)
#define ADDS_I16(x, y) ((int)(x) + (int)(y) > 32767? (int16_t)32767 : (int)(x) + (int)(y) < -32768? (int16_t)-32768 : (int16_t)((x) + (y)))
	for (i = 0; i < tlen; ++i) {
		int16_t h[8], e[8], f[8] = {0,0,0,0,0,0,0,0}, mx[8] = {0,0,0,0,0,0,0,0};
		const int16_t *S = qp + (size_t)target[i] * n8;
		int k, imax, done = 0;
		for (l = 7; l >= 1; --l) h[l] = H0[(slen - 1) * 8 + l - 1]; /* _mm_slli_si128(H0[slen-1], 2) */
		h[0] = 0;
		for (j = 0; j < slen; ++j) {
			for (l = 0; l < 8; ++l) {
				int16_t hh = ADDS_I16(h[l], S[j * 8 + l]), t;
				e[l] = E[j * 8 + l];
				hh = hh > e[l]? hh : e[l];
				hh = hh > f[l]? hh : f[l];
				mx[l] = mx[l] > hh? mx[l] : hh;
				H1[j * 8 + l] = hh;
				t = SUBS_U16(hh, gapoe);
				e[l] = SUBS_U16(e[l], gape);
				e[l] = e[l] > t? e[l] : t;
				E[j * 8 + l] = e[l];
				f[l] = SUBS_U16(f[l], gape);
				f[l] = f[l] > t? f[l] : t;
				h[l] = H0[j * 8 + l];
			}
		}
		for (k = 0; k < 8 && !done; ++k) { /* lazy F */
			for (l = 7; l >= 1; --l) f[l] = f[l - 1];
			f[0] = 0;
			for (j = 0; j < slen; ++j) {
				int any = 0;
				for (l = 0; l < 8; ++l) {
					int16_t hh = H1[j * 8 + l];
					hh = hh > f[l]? hh : f[l];
					H1[j * 8 + l] = hh;
					hh = SUBS_U16(hh, gapoe);
					f[l] = SUBS_U16(f[l], gape);
					if (f[l] > hh) any = 1;
				}
				if (!any) { done = 1; break; }
			}
		}
		for (l = 1, imax = mx[0]; l < 8; ++l) imax = imax > mx[l]? imax : mx[l];
		if (imax >= gmax) {
			gmax = imax, *te = i;
			memcpy(Hmax, H1, (size_t)n8 * sizeof(int16_t));
		}
		tmp = H1, H1 = H0, H0 = tmp;
	}
#undef SUBS_U16
#undef ADDS_I16
	for (i = 0; i < n8; ++i)
		if ((int)(uint16_t)Hmax[i] == gmax) *qe = i / 8 + i % 8 * slen;
	free(qp);
	return gmax;
}
