// FM-index primitives on the reference's interleaved occ/BWT layout.
//
// Replaces bwt_occ / bwt_occ4 / bwt_2occ4 / bwt_invPsi / bwt_sa of reference src/bwt_search.cpp:25-119.
// A 64-byte block covers 128 BWT symbols: 4 x uint64 running counts (A,C,G,T before the block)
// followed by 8 x uint32 words holding 16 symbols each, first symbol in the top bits.  One block is
// two 32-byte sectors; a thread fetches it with four 128-bit read-only loads (LDG.E.128.CONSTANT).
// Instead of the reference's byte-wise cnt_table the count of ONE symbol up to a row comes from popcounts:
// the words are XOR-ed with a per-symbol pattern that turns the wanted symbol into binary 11, (w>>1)&w marks
// it on the even bits, a funnel-shift builds the prefix mask, and two words share one POPC (the marks of the
// second word are moved to the odd bits).
#ifndef MC_FMINDEX_H
#define MC_FMINDEX_H

#include "mc_device.h"

MC_HD int mc_nt4(uint8_t c)
{
	// nst_nt4_table (reference src/BWT_Index/bntseq.c:40-57): ACGT / acgt -> 0..3, everything else 4
	// branch-free: bits 0, 2, 6, 19 of 0x80045 mark A, C, G, T among the folded letters; their bits 1..2 spell the code
	const uint32_t u = c & 0xDFu, idx = u - 65u;
	const uint32_t valid = (0x80045u >> (idx < 31u ? idx : 31u)) & 1u;
	return valid ? (int)(((u >> 1) ^ (u >> 2)) & 3u) : 4;
}

MC_HD uint8_t mc_complement(uint8_t c)
{
	// GetComplementaryBase (reference src/tools.cpp:3-17): lower case folds to upper, non-ACGT -> 'N'
	switch (c) { case 'A': case 'a': return 'T'; case 'C': case 'c': return 'G'; case 'G': case 'g': return 'C'; case 'T': case 't': return 'A'; default: return 'N'; }
}

// 2-bit code of RefSequence[p], 0 <= p < 2G (forward half from pac, reverse half = complement of the mirror)
MC_HD int mc_ref_code(const DevIndex& ix, int64_t p)
{
	if (p < ix.G) return (mc_ldg(ix.pac + (p >> 2)) >> ((~p & 3) << 1)) & 3;
	int64_t q = ix.twoG - 1 - p;
	return 3 - ((mc_ldg(ix.pac + (q >> 2)) >> ((~q & 3) << 1)) & 3);
}
MC_HD uint8_t mc_ref_char(const DevIndex& ix, int64_t p) { return (uint8_t)("ACGT"[mc_ref_code(ix, p)]); }

// PosChrIdMap.lower_bound(g): index of the first chromosome end >= g (n_end when none)
MC_HD int mc_chrom_lower_bound(const DevIndex& ix, int64_t g)
{
	int lo = 0, hi = ix.n_end;
	while (lo < hi) { int mid = (lo + hi) >> 1; if (mc_ldg(ix.chrom_end + mid) < g) lo = mid + 1; else hi = mid; }
	return lo;
}

struct OccBlock { mc_u32x4 q0, q1, q2, q3; };

MC_HD void mc_load_block(const DevIndex& ix, uint64_t blk, OccBlock& b)
{
	const uint32_t* p = ix.bwt + (blk << 4);
	b.q0 = mc_gather128(p); b.q1 = mc_gather128(p + 4); b.q2 = mc_gather128(p + 8); b.q3 = mc_gather128(p + 12);
}

// bwt_invPsi (reference src/bwt_search.cpp:101-107): one LF step = one block
MC_HD uint64_t mc_lf_step(const DevIndex& ix, uint64_t k);

// The sampled suffix array in HBM.  The reference keeps SA[k] for every 32nd row (bwt.c:101-123) and walks LF until it meets
// one: 31 dependent random block reads per location on average.  HBM is not the scarce resource on a B200, so the context
// derives a denser sample from the reference's at upload time (mc_sa_dense_body: every 4th row in 32 bits for texts below 2^32
// symbols = one byte per text symbol, every 8th row in 64 bits otherwise) and a location costs 3 (7) steps.  Same values:
// SA[k] = steps + SA[LF^steps(k)] whichever sampled row the walk stops at.
MC_HD bool mc_sa_sampled(const DevIndex& ix, uint64_t k) { return (k & ((1ull << ix.sa_shift) - 1)) == 0; }
MC_HD uint64_t mc_sa_value(const DevIndex& ix, uint64_t k, uint64_t steps)
{
	// 32-bit entries: row 0 holds (uint32)-1 like the reference's sa[0] = -1, and is only ever reached after >= 1 step
	if (ix.sa32) return (uint64_t)(uint32_t)((uint32_t)steps + mc_gather32(ix.sa32 + (k >> ix.sa_shift)));
	return steps + mc_gather64(ix.sa + (k >> ix.sa_shift));
}
// bwt_sa (reference src/bwt_search.cpp:109-119): walk LF until a sampled row
MC_HD uint64_t mc_locate(const DevIndex& ix, uint64_t k, uint32_t* nblk)
{
	uint64_t steps = 0;
	while (!mc_sa_sampled(ix, k)) { steps++; k = mc_lf_step(ix, k); }
	*nblk += (uint32_t)steps;
	return mc_sa_value(ix, k, steps);
}

// Search state.  The reference carries the bi-interval (x0 = rows of the pattern, x1 = rows of its reverse
// complement, x2 = size; src/bwt_search.cpp:121-151).  Extending the pattern forward by base c is a plain backward
// step on the x1 interval with symbol 3-c, and x0 only serves to enumerate the hits at the end.  The text is its own
// reverse complement, so the hits can be enumerated from the x1 interval just as well (an occurrence of the reverse
// complement at q is an occurrence of the pattern at 2G - q - len); the reference's later sort makes the order of
// the hits immaterial (src/ReadMapping.cpp:152).  Dropping x0 removes half of the counting work per step.
struct RcInterval { uint64_t x1, x2; };

// first base of BWT_Search (reference src/bwt_search.cpp:128-131)
MC_HD RcInterval mc_interval_init(const DevIndex& ix, int c)
{
	RcInterval v; v.x1 = ix.L2[3 - c] + 1; v.x2 = ix.L2[c + 1] - ix.L2[c];
	return v;
}

#ifdef MC_HOSTEMU
static inline uint32_t mc_prefix_pairs(int bits) { return bits <= 0 ? 0u : bits >= 32 ? 0x55555555u : (uint32_t)((0x5555555500000000ull >> bits) & 0xFFFFFFFFu); }
#else
// even-bit mask of the top `bits`/2 symbols of a word: low half of 0x55555555:00000000 >> bits, shift clamped to 32
static __device__ __forceinline__ uint32_t mc_prefix_pairs(int bits) { return __funnelshift_rc(0u, 0x55555555u, (uint32_t)max(bits, 0)); }
#endif

// number of symbols equal to the symbol encoded in `flip` among the first nbits/2 symbols of the block
MC_HD int mc_count_in_block(const OccBlock& b, uint32_t flip, int nbits)
{
	const uint32_t w[8] = {b.q2.x, b.q2.y, b.q2.z, b.q2.w, b.q3.x, b.q3.y, b.q3.z, b.q3.w};
	int n = 0;
#pragma unroll
	for (int j = 0; j < 8; j += 2)
	{
		const uint32_t a = w[j] ^ flip, c = w[j + 1] ^ flip;
		const uint32_t e0 = (a >> 1) & a & mc_prefix_pairs(nbits - 32 * j);
		const uint32_t e1 = (c >> 1) & c & mc_prefix_pairs(nbits - 32 * (j + 1));
		n += mc_popc(e0 + (e1 << 1));     // marks of the second word ride on the odd bits: one POPC for two words
	}
	return n;
}
MC_HD uint32_t mc_flip_of(int i) { return ((i & 2) ? 0u : 0xAAAAAAAAu) | ((i & 1) ? 0u : 0x55555555u); }
MC_HD uint64_t mc_block_base(const OccBlock& b, int i)
{
	const uint64_t cA = (uint64_t)b.q0.y << 32 | b.q0.x, cC = (uint64_t)b.q0.w << 32 | b.q0.z;
	const uint64_t cG = (uint64_t)b.q1.y << 32 | b.q1.x, cT = (uint64_t)b.q1.w << 32 | b.q1.z;
	return (i & 2) ? ((i & 1) ? cT : cG) : ((i & 1) ? cC : cA);
}

// ---- compact device layout ---------------------------------------------------------------------------------------
// The on-disk / reference layout spends 32 of every 64 bytes on 64-bit counts and makes a lookup touch two 32-byte sectors
// through four 128-bit loads.  For texts below 2^32 symbols the context re-blocks the index at upload time into 32-byte
// blocks of 64 rows: 4 x uint32 absolute counts, then the 64 symbols as two bit planes (low bits, high bits; row r of the
// block is bit r).  A lookup is ONE sector fetched by ONE 256-bit load (LDG.E.256 is new with sm_100), and the count of a
// symbol among the first n rows is two 3-input logic ops and a POPC per 32 rows.  Rows fit 32 bits, so the search state
// does as well.  (mc_cbwt_build_body below; the 128-row layout stays the path for larger texts.)
struct CBlock { uint32_t c0, c1, c2, c3, lo0, lo1, hi0, hi1; };
#ifdef MC_HOSTEMU
static inline void mc_load_cblock(const DevIndex& ix, uint32_t blk, CBlock& b) { memcpy(&b, ix.cbwt + ((size_t)blk << 3), 32); }
static inline uint32_t mc_prefix_bits(int n) { return n <= 0 ? 0u : n >= 32 ? 0xFFFFFFFFu : (1u << n) - 1u; }
#else
static __device__ __forceinline__ void mc_load_cblock(const DevIndex& ix, uint32_t blk, CBlock& b)
{
	asm volatile("ld.global.nc.L2::64B.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
	             : "=r"(b.c0), "=r"(b.c1), "=r"(b.c2), "=r"(b.c3), "=r"(b.lo0), "=r"(b.lo1), "=r"(b.hi0), "=r"(b.hi1) : "l"(ix.cbwt + ((size_t)blk << 3)));
}
// mask of the low min(n, 32) bits: high word of 0:FFFFFFFF << n, shift clamped to 32
static __device__ __forceinline__ uint32_t mc_prefix_bits(int n) { return __funnelshift_lc(0xFFFFFFFFu, 0u, (uint32_t)max(n, 0)); }
#endif
MC_HD uint32_t mc_cblock_base(const CBlock& b, int i) { return (i & 2) ? ((i & 1) ? b.c3 : b.c2) : ((i & 1) ? b.c1 : b.c0); }
// symbol i among the first n rows of the block (1 <= n <= 64)
MC_HD uint32_t mc_count_in_cblock(const CBlock& b, int i, int n)
{
	const uint32_t fl = (i & 1) ? 0u : 0xFFFFFFFFu, fh = (i & 2) ? 0u : 0xFFFFFFFFu;
	const uint32_t e0 = ((b.lo0 ^ fl) & mc_prefix_bits(n)) & (b.hi0 ^ fh);
	const uint32_t e1 = ((b.lo1 ^ fl) & mc_prefix_bits(n - 32)) & (b.hi1 ^ fh);
	return (uint32_t)(mc_popc(e0) + mc_popc(e1));
}
MC_HD int mc_cblock_symbol(const CBlock& b, uint32_t r)
{
	const uint32_t lo = (r & 32) ? b.lo1 : b.lo0, hi = (r & 32) ? b.hi1 : b.hi0;
	return (int)(((lo >> (r & 31)) & 1u) | (((hi >> (r & 31)) & 1u) << 1));
}
// one compact block from the reference layout: block b covers rows [64 b, 64 b + 64) of source block b >> 1
MC_HD void mc_cbwt_build_body(int64_t b, const uint32_t* src, uint32_t* dst)
{
	const uint32_t* s = src + ((b >> 1) << 4);
	uint32_t c[4] = {(uint32_t)s[0], (uint32_t)s[2], (uint32_t)s[4], (uint32_t)s[6]};   // low halves of the 64-bit counts
	const uint32_t* w = s + 8;
	if (b & 1)
		for (int j = 0; j < 4; j++)
			for (int i = 0; i < 4; i++) { const uint32_t a = w[j] ^ mc_flip_of(i); c[i] += (uint32_t)mc_popc((a >> 1) & a & 0x55555555u); }
	uint32_t lo[2] = {0, 0}, hi[2] = {0, 0};
	for (int r = 0; r < 64; r++)
	{
		const uint32_t sym = (w[((b & 1) << 2) + (r >> 4)] >> ((~(uint32_t)r & 15) << 1)) & 3u;
		lo[r >> 5] |= (sym & 1u) << (r & 31); hi[r >> 5] |= (sym >> 1) << (r & 31);
	}
	uint32_t* d = dst + (b << 3);
	d[0] = c[0]; d[1] = c[1]; d[2] = c[2]; d[3] = c[3]; d[4] = lo[0]; d[5] = lo[1]; d[6] = hi[0]; d[7] = hi[1];
}

// the 48 bytes of a block that a count of symbol i needs: the 8 words and the 16-byte quarter holding count[i]
// (A,C sit in the first quarter, G,T in the second).  Three 128-bit loads instead of four: the L1TEX tag stage, which
// sees one wavefront per lane and load because every lane reads a different line, is what bounds this kernel (profiles/).
struct OccPart { mc_u32x4 qc, q2, q3; };
MC_HD void mc_load_part(const DevIndex& ix, uint64_t blk, int i, OccPart& b)
{
	const uint32_t* p = ix.bwt + (blk << 4);
	b.qc = mc_gather128(p + ((i & 2) << 1)); b.q2 = mc_gather128(p + 8); b.q3 = mc_gather128(p + 12);
}
MC_HD uint64_t mc_part_base(const OccPart& b, int i) { return (i & 1) ? ((uint64_t)b.qc.w << 32 | b.qc.z) : ((uint64_t)b.qc.y << 32 | b.qc.x); }
MC_HD int mc_count_in_part(const OccPart& b, uint32_t flip, int nbits)
{
	const uint32_t w[8] = {b.q2.x, b.q2.y, b.q2.z, b.q2.w, b.q3.x, b.q3.y, b.q3.z, b.q3.w};
	int n = 0;
#pragma unroll
	for (int j = 0; j < 8; j += 2)
	{
		const uint32_t a = w[j] ^ flip, c = w[j + 1] ^ flip;
		const uint32_t e0 = (a >> 1) & a & mc_prefix_pairs(nbits - 32 * j);
		const uint32_t e1 = (c >> 1) & c & mc_prefix_pairs(nbits - 32 * (j + 1));
		n += mc_popc(e0 + (e1 << 1));
	}
	return n;
}

// One forward extension by base code c (reference src/bwt_search.cpp:138-149); false = empty child.
// The two rows k, l fall into the same 128-row block in most steps (the reference's bwt_2occ4 fast path, :73);
// the second block is only fetched when they do not, everything after the fetch is the same code for every lane.
MC_HD bool mc_interval_extend(const DevIndex& ix, RcInterval& v, int c, uint32_t* nblk)
{
	const int i = 3 - c;
	const uint32_t flip = mc_flip_of(i);
	const uint64_t k = v.x1 - 1, l = v.x1 - 1 + v.x2;                 // x1 >= 1, so k never is the (uint64)-1 row
	const uint64_t kk = k - (k >= ix.primary), ll = l - (l >= ix.primary);
	*nblk += (kk >> 7) == (ll >> 7) ? 1u : 2u;                         // blocks the reference algorithm touches
	OccPart bk, bl;
	mc_load_part(ix, kk >> 7, i, bk);
	bl = bk;
	if ((kk >> 7) != (ll >> 7)) mc_load_part(ix, ll >> 7, i, bl);
	const uint64_t occ_k = mc_part_base(bk, i) + (uint64_t)mc_count_in_part(bk, flip, 2 * ((int)(kk & 127) + 1));
	const uint64_t occ_l = mc_part_base(bl, i) + (uint64_t)mc_count_in_part(bl, flip, 2 * ((int)(ll & 127) + 1));
	if (occ_l == occ_k) return false;
	v.x1 = ix.L2[i] + 1 + occ_k; v.x2 = occ_l - occ_k;
	return true;
}
// the same step on the compact layout, all rows 32-bit
struct RcInterval32 { uint32_t x1, x2; };
MC_HD RcInterval32 mc_interval_init32(const DevIndex& ix, int c)
{
	RcInterval32 v; v.x1 = (uint32_t)ix.L2[3 - c] + 1u; v.x2 = (uint32_t)(ix.L2[c + 1] - ix.L2[c]);
	return v;
}
MC_HD bool mc_interval_extend(const DevIndex& ix, RcInterval32& v, int c, uint32_t* nblk)
{
	const int i = 3 - c;
	const uint32_t prim = (uint32_t)ix.primary;
	const uint32_t k = v.x1 - 1u, l = k + v.x2;
	const uint32_t kk = k - (k >= prim), ll = l - (l >= prim);
	*nblk += (kk >> 7) == (ll >> 7) ? 1u : 2u;                         // blocks the reference algorithm touches (its layout)
	CBlock bk, bl;
	mc_load_cblock(ix, kk >> 6, bk);
	bl = bk;
	if ((kk >> 6) != (ll >> 6)) mc_load_cblock(ix, ll >> 6, bl);
	const uint32_t occ_k = mc_cblock_base(bk, i) + mc_count_in_cblock(bk, i, (int)(kk & 63) + 1);
	const uint32_t occ_l = mc_cblock_base(bl, i) + mc_count_in_cblock(bl, i, (int)(ll & 63) + 1);
	if (occ_l == occ_k) return false;
	v.x1 = (uint32_t)ix.L2[i] + 1u + occ_k; v.x2 = occ_l - occ_k;
	return true;
}

// One backward-search step on the compact layout in the form the seed kernel uses for both of its index phases, so that lanes
// extending a pattern and lanes walking towards a sampled row run the SAME instructions: extending {x1, x2} by the symbol
// 3 - c, or - `walk` - by whatever symbol the BWT holds at row x1 (x2 == 1), which is the LF step of that row
// (LF(k) = L2[s] + occ(s, k) = the first row of the extended interval).  false = empty child (never when walking).
MC_HD bool mc_fm_step32(const DevIndex& ix, RcInterval32& v, int c, bool walk, uint32_t* nblk)
{
	const uint32_t prim = (uint32_t)ix.primary;
	const uint32_t k = v.x1 - 1u, l = k + v.x2;
	const uint32_t kk = k - (k >= prim), ll = l - (l >= prim);
	if (!walk) *nblk += (kk >> 7) == (ll >> 7) ? 1u : 2u;              // blocks the reference algorithm touches (its layout)
	CBlock bk, bl;
	mc_load_cblock(ix, kk >> 6, bk);
	bl = bk;
	if ((kk >> 6) != (ll >> 6)) mc_load_cblock(ix, ll >> 6, bl);
	const int i = walk ? mc_cblock_symbol(bl, ll & 63u) : 3 - c;
	const uint32_t occ_k = mc_cblock_base(bk, i) + mc_count_in_cblock(bk, i, (int)(kk & 63) + 1);
	const uint32_t occ_l = mc_cblock_base(bl, i) + mc_count_in_cblock(bl, i, (int)(ll & 63) + 1);
	if (occ_l == occ_k) return false;
	v.x1 = (uint32_t)ix.L2[i] + 1u + occ_k; v.x2 = occ_l - occ_k;
	return true;
}

MC_HD uint64_t mc_lf_step(const DevIndex& ix, uint64_t k)
{
	if (k == ix.primary) return 0;
	const uint64_t x = k - (k > ix.primary);
	if (ix.cbwt)
	{
		const uint32_t x32 = (uint32_t)x;
		CBlock b; mc_load_cblock(ix, x32 >> 6, b);
		const int c = mc_cblock_symbol(b, x32 & 63);
		return (uint64_t)((uint32_t)ix.L2[c] + mc_cblock_base(b, c) + mc_count_in_cblock(b, c, (int)(x32 & 63) + 1));
	}
	OccBlock b; mc_load_block(ix, x >> 7, b);
	const uint32_t w[8] = {b.q2.x, b.q2.y, b.q2.z, b.q2.w, b.q3.x, b.q3.y, b.q3.z, b.q3.w};
	const int c = (int)(w[(x & 127) >> 4] >> ((~(uint32_t)x & 15) << 1)) & 3;
	return ix.L2[c] + mc_block_base(b, c) + (uint64_t)mc_count_in_block(b, mc_flip_of(c), 2 * ((int)(x & 127) + 1));
}

// ---- k-mer start table -------------------------------------------------------------------------------------------
// The first steps of every seed search are the expensive ones: the interval still spans millions of rows, so its two ends
// lie in different blocks (two random lines per step, the reference's bwt_2occ4 slow path) - and they are the same steps for
// every read that starts with the same bases.  The context therefore tabulates, for every k-mer (k ~ log4(text) - 2: 12 at
// 248 Mbp), the search state after its k bases: entry = {x1, x2 | touched << 27} (32-bit) or {x1, x2 | touched << 56} (64-bit),
// `touched` = the number of occ blocks the reference algorithm reads on those k - 1 steps, so that the work counter stays the
// oracle's.  x2 == 0 marks a k-mer the table cannot answer (absent from the text - the search then has to find out WHERE it
// fails -, or more than 2^27 occurrences): the search falls back to stepping, as it does at an N.
#define MC_KTAB_BITS32 27
#define MC_KTAB_BITS64 56
template <class Interval> struct KtabOps;
MC_HD void mc_ktab_build_body(int64_t m, const DevIndex& ix, int k, uint32_t* out32, uint64_t* out64);

// entry j of the denser sample: SA[j << shift] from the reference's every-32nd-row sample (ix.sa, ix.sa_shift == 5 here)
MC_HD void mc_sa_dense_body(int64_t j, const DevIndex& ix, int shift, uint32_t* out32, uint64_t* out64)
{
	uint64_t k = (uint64_t)j << shift, steps = 0;
	while (k & 31) { k = mc_lf_step(ix, k); steps++; }
	const uint64_t v = steps + mc_ldg(ix.sa + (k >> 5));
	if (out32) out32[j] = (uint32_t)v; else out64[j] = v;
}

template <> struct KtabOps<RcInterval32> {
	static MC_HD bool lookup(const DevIndex& ix, uint32_t m, RcInterval32& v, uint32_t* nblk)
	{
		const uint64_t e = mc_gather64((const uint64_t*)ix.ktab32 + m); const uint32_t x1 = (uint32_t)e, y = (uint32_t)(e >> 32);
		if (!y) return false;
		v.x1 = x1; v.x2 = y & ((1u << MC_KTAB_BITS32) - 1); *nblk += y >> MC_KTAB_BITS32;
		return true;
	}
};
template <> struct KtabOps<RcInterval> {
	static MC_HD bool lookup(const DevIndex& ix, uint32_t m, RcInterval& v, uint32_t* nblk)
	{
		const mc_u32x4 e = mc_gather128(ix.ktab64 + 2 * (size_t)m); const uint64_t x1 = (uint64_t)e.y << 32 | e.x, y = (uint64_t)e.w << 32 | e.z;
		if (!y) return false;
		v.x1 = x1; v.x2 = y & ((1ull << MC_KTAB_BITS64) - 1); *nblk += (uint32_t)(y >> MC_KTAB_BITS64);
		return true;
	}
};
// entry m of the table: the k bases of m, first base in the low bits (the order mc_pack8 delivers), searched step by step
MC_HD void mc_ktab_build_body(int64_t m, const DevIndex& ix, int k, uint32_t* out32, uint64_t* out64)
{
	uint32_t nblk = 0; bool ok = true;
	if (out32)
	{
		RcInterval32 v = mc_interval_init32(ix, (int)(m & 3));
		for (int j = 1; j < k && ok; j++) ok = mc_interval_extend(ix, v, (int)((m >> (2 * j)) & 3), &nblk);
		ok = ok && v.x2 > 0 && v.x2 < (1u << MC_KTAB_BITS32);
		out32[2 * m] = ok ? v.x1 : 0u; out32[2 * m + 1] = ok ? (v.x2 | (nblk << MC_KTAB_BITS32)) : 0u;
	}
	else
	{
		RcInterval v = mc_interval_init(ix, (int)(m & 3));
		for (int j = 1; j < k && ok; j++) ok = mc_interval_extend(ix, v, (int)((m >> (2 * j)) & 3), &nblk);
		ok = ok && v.x2 > 0 && v.x2 < (1ull << MC_KTAB_BITS64);
		out64[2 * m] = ok ? v.x1 : 0ull; out64[2 * m + 1] = ok ? (v.x2 | ((uint64_t)nblk << MC_KTAB_BITS64)) : 0ull;
	}
}

#endif
