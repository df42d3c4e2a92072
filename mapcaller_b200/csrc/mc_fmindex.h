// FM-index primitives on the reference's interleaved occ/BWT layout.
//
// Replaces bwt_occ / bwt_occ4 / bwt_2occ4 / bwt_invPsi / bwt_sa of reference src/bwt_search.cpp:25-119.
// A 64-byte block covers 128 BWT symbols: 4 x uint64 running counts (A,C,G,T before the block)
// followed by 8 x uint32 words holding 16 symbols each, first symbol in the top bits.  One block is
// two 32-byte sectors; a thread fetches it with four 128-bit read-only loads (LDG.E.128.CONSTANT).
// Instead of the reference's byte-wise cnt_table the symbol counts come from three popcounts per
// word (pair masks for C, G, T; A is what is left of the symbols scanned).
#ifndef MC_FMINDEX_H
#define MC_FMINDEX_H

#include "mc_device.h"

MC_HD int mc_nt4(uint8_t c)
{
	// nst_nt4_table (reference src/BWT_Index/bntseq.c:40-57): ACGT / acgt -> 0..3, everything else 4
	switch (c & 0xDF) { case 'A': return 0; case 'C': return 1; case 'G': return 2; case 'T': return 3; default: return 4; }
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
	b.q0 = mc_ldg128(p); b.q1 = mc_ldg128(p + 4); b.q2 = mc_ldg128(p + 8); b.q3 = mc_ldg128(p + 12);
}

// occ(A,C,G,T) over symbols [0, k] of the $-less BWT, k already adjusted for primary, block already loaded
MC_HD void mc_occ4_in_block(const OccBlock& b, uint64_t k, uint64_t out[4])
{
	const uint32_t w[8] = {b.q2.x, b.q2.y, b.q2.z, b.q2.w, b.q3.x, b.q3.y, b.q3.z, b.q3.w};
	const int full = (int)((k & 127) >> 4);
	const uint32_t part = ~((1u << ((~(uint32_t)k & 15) << 1)) - 1u);
	int n1 = 0, n2 = 0, n3 = 0;
#pragma unroll
	for (int j = 0; j < 8; j++)
	{
		uint32_t m = j < full ? 0x55555555u : (j == full ? (part & 0x55555555u) : 0u);
		uint32_t lo = w[j], hi = w[j] >> 1;
		n1 += mc_popc(~hi & lo & m); n2 += mc_popc(hi & ~lo & m); n3 += mc_popc(hi & lo & m);
	}
	const int total = (int)(k & 127) + 1;
	out[0] = ((uint64_t)b.q0.y << 32 | b.q0.x) + (uint64_t)(total - n1 - n2 - n3);
	out[1] = ((uint64_t)b.q0.w << 32 | b.q0.z) + (uint64_t)n1;
	out[2] = ((uint64_t)b.q1.y << 32 | b.q1.x) + (uint64_t)n2;
	out[3] = ((uint64_t)b.q1.w << 32 | b.q1.z) + (uint64_t)n3;
}

// bwt_2occ4 (reference src/bwt_search.cpp:68-99).  *nblk += number of 64-byte blocks the reference touches.
MC_HD void mc_occ4_pair(const DevIndex& ix, uint64_t k, uint64_t l, uint64_t ck[4], uint64_t cl[4], uint32_t* nblk)
{
	const bool kn = (k == ~0ull), ln = (l == ~0ull);
	uint64_t kk = k - (k >= ix.primary), ll = l - (l >= ix.primary);
	OccBlock b;
	if (kn) { ck[0] = ck[1] = ck[2] = ck[3] = 0; }
	else { mc_load_block(ix, kk >> 7, b); mc_occ4_in_block(b, kk, ck); (*nblk)++; }
	if (ln) { cl[0] = cl[1] = cl[2] = cl[3] = 0; }
	else
	{
		if (kn || (kk >> 7) != (ll >> 7)) { mc_load_block(ix, ll >> 7, b); (*nblk)++; }
		mc_occ4_in_block(b, ll, cl);
	}
}

// bwt_invPsi (reference src/bwt_search.cpp:101-107): one LF step = one block
MC_HD uint64_t mc_lf_step(const DevIndex& ix, uint64_t k)
{
	if (k == ix.primary) return 0;
	const uint64_t x = k - (k > ix.primary);
	OccBlock b; mc_load_block(ix, x >> 7, b);
	const uint32_t w[8] = {b.q2.x, b.q2.y, b.q2.z, b.q2.w, b.q3.x, b.q3.y, b.q3.z, b.q3.w};
	const int full = (int)((x & 127) >> 4);
	const int c = (int)(w[full] >> ((~(uint32_t)x & 15) << 1)) & 3;
	const uint32_t part = ~((1u << ((~(uint32_t)x & 15) << 1)) - 1u);
	int n = 0;
#pragma unroll
	for (int j = 0; j < 8; j++)
	{
		uint32_t m = j < full ? 0x55555555u : (j == full ? (part & 0x55555555u) : 0u);
		uint32_t lo = w[j], hi = w[j] >> 1;
		n += mc_popc(((c & 2) ? hi : ~hi) & ((c & 1) ? lo : ~lo) & m);
	}
	const uint64_t base = c == 0 ? ((uint64_t)b.q0.y << 32 | b.q0.x) : c == 1 ? ((uint64_t)b.q0.w << 32 | b.q0.z)
	                    : c == 2 ? ((uint64_t)b.q1.y << 32 | b.q1.x) : ((uint64_t)b.q1.w << 32 | b.q1.z);
	return ix.L2[c] + base + (uint64_t)n;
}

// bwt_sa (reference src/bwt_search.cpp:109-119): walk LF until a sampled row
MC_HD uint64_t mc_locate(const DevIndex& ix, uint64_t k, uint32_t* nblk)
{
	uint64_t steps = 0;
	while (k & 31) { steps++; k = mc_lf_step(ix, k); }
	*nblk += (uint32_t)steps;
	return steps + mc_ldg(ix.sa + (k >> 5));
}

struct BiInterval { uint64_t x0, x1, x2; };

// first base of BWT_Search (reference src/bwt_search.cpp:128-131)
MC_HD BiInterval mc_interval_init(const DevIndex& ix, int c)
{
	BiInterval v; v.x0 = ix.L2[c] + 1; v.x1 = ix.L2[3 - c] + 1; v.x2 = ix.L2[c + 1] - ix.L2[c];
	return v;
}

// one forward extension by base code c (reference src/bwt_search.cpp:138-149); false = empty child
MC_HD bool mc_interval_extend(const DevIndex& ix, BiInterval& v, int c, uint32_t* nblk)
{
	uint64_t tk[4], tl[4];
	mc_occ4_pair(ix, v.x1 - 1, v.x1 - 1 + v.x2, tk, tl, nblk);
	const int i = 3 - c;
	const uint64_t n2 = tl[i] - tk[i];
	if (n2 == 0) return false;
	uint64_t n0 = v.x0 + ((v.x1 <= ix.primary && v.x1 + v.x2 - 1 >= ix.primary) ? 1 : 0);
	for (int j = 3; j > i; j--) n0 += tl[j] - tk[j];
	v.x0 = n0; v.x1 = ix.L2[i] + 1 + tk[i]; v.x2 = n2;
	return true;
}

#endif
