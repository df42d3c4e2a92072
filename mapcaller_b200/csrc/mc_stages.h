// Per-work-item bodies of the mapping pipeline (one function = one kernel, see mc_launch.h; a work item is a read, a pair,
// a piece, a rescue window, ... and is handled by a thread, an 8-lane tile, a warp or a thread block as noted at the body).
//
// Stage map (reference function -> body):
//   prep_body      ReverseOrientation + EnCodeReadSeq           src/tools.cpp:45, src/ReadMapping.cpp:404
//   seed_body      IdentifySimplePairs + BWT_Search             src/ReadMapping.cpp:125, src/bwt_search.cpp:121
//   expand_body / locate_body   the bwt_sa loop of BWT_Search   src/bwt_search.cpp:153-161,109-119
//   cluster_body   sort(CompByPosDiff) + SimplePairClustering   src/ReadMapping.cpp:152,194,160
//   pair_body      CheckPairedAlignmentDistance, MaskUnPairedAlnCan, RemoveRedundantAlnCan  :244,305,228
//   rwenum / rwin / rcommit_body   AlignmentRescue              src/AlignmentRescue.cpp:28, src/KmerAnalysis.cpp
//   alnprep_body   ProduceReadAlignment: fragment lists          src/ReadAlignment.cpp:306-342,38-108
//   piece_body     ProcessNormalPair: strings + DP dispatch      src/ReadAlignment.cpp:155-191
//   dp_core (dp_small_body, mc_dp_warp.cuh)   nw_alignment / ksw2_alignment   src/nw_alignment.cpp:18, src/ksw2_alignment.cpp:250
//   alnfin_body    rest of ProduceReadAlignment                 src/ReadAlignment.cpp:343-411
//   pairstat_body  GenCoordinatePair + per-chunk sums           src/ReadMapping.cpp:361,479-539
//   profkey / gate / scatter / profpiece_body   UpdateProfile / UpdateMultiHitCount   src/AlignmentProfile.cpp:41,244
//   search_walk    BWT_Search as an operator (mc_bwt_search_batch)
#ifndef MC_STAGES_H
#define MC_STAGES_H

#include "mc_fmindex.h"

// ------------------------------------------------------------------------------------------------
// Shared argument block: every stage sees the same set of arenas.
// ------------------------------------------------------------------------------------------------
struct ReadSum { int32_t score, sub_score, best_idx, n_live; };
// one window of AlignmentRescue: place the mate of `pair` (dir 0: mate 2, dir 1: mate 1) next to candidate `cand` of the other mate
struct RWin { int32_t pair; int16_t dir; int16_t pad; int32_t cand; int32_t floor; int32_t ok; int32_t score; int32_t pbeg; int32_t n; int32_t lo; int32_t hi; };

struct PipeArgs {
	DevIndex ix;
	DevParams pr;
	DevStats* st;
	int64_t n_reads;
	// reads
	uint8_t* seq;            // ASCII, mate 2 reverse-complemented in place by prep
	const int64_t* roff;     // n_reads + 1
	// seeds: slot s of read r lives in [seed_off[r], seed_off[r+1])
	const int64_t* seed_off; // n_reads + 1
	Seed* seeds;
	uint32_t* slot_freq;     // per slot: number of locations (0 = unused slot)
	const int64_t* slot_loc; // exclusive scan of slot_freq, n_slots + 1
	int64_t n_slots, n_locs;
	// simple pairs: [pair_off(r), ...) with pair_off(r) = slot_loc[seed_off[r]]; rescue seeds appended after n_locs
	SPair* pairs;
	int32_t* npair;          // per read: valid simple pairs after filtering/sorting
	uint8_t* rflag;          // per read: bit0 = a base walked by the seeding is lower case (profile counts upper case only)
	int64_t pair_cap;        // arena capacity (n_locs + rescue region)
	mc_u64* pair_bump;       // next free rescue pair
	// candidates: read r owns [cand_off(r), cand_off(r) + cand_cap(r))
	Cand* cands;
	int32_t* cand_off;       // per read: first candidate slot (written by the cluster stage)
	int32_t* ncand0;         // per read: clusters found (immutable after cluster stage)
	int32_t* ncand;          // per read: clusters + rescued candidates of the current attempt
	int32_t* cscore;         // per cand: live score of the current attempt
	int32_t* cpaired;        // per cand: PairedAlnCanIdx
	int32_t* corient;        // per cand: orientation (1 fwd, 0 rev, -1 dead)
	int32_t* cfrag;          // per cand: first fragment in the frag arena
	int32_t* cnfrag;         // per cand: number of fragments
	int32_t* ctmp;           // per cand: scratch (pairing)
	// per pair / per chunk
	const int32_t* est;      // per chunk: EstiDistance = (int)(avgDist*1.5)
	const uint8_t* active;   // per chunk: 1 = (re)compute in this attempt
	int32_t* pair_flag;      // per pair: bit0 = a result of this batch exists, bit1 = recomputed in the current attempt
	uint8_t* read_redo;      // per read: 1 = (re)computed in the current attempt
	int32_t* est_lo;         // per pair: smallest / largest EstiDistance giving the same outcome
	int32_t* est_hi;
	mc_pair_out* pair_out;   // per pair
	mc_chunk_out* chunk_out; // per chunk
	int32_t* chunk_lo; int32_t* chunk_hi;
	// alignment
	ReadSum* rsum;           // per read
	mc_frag_out* frags; int64_t frag_cap; mc_u64* frag_bump;
	uint8_t* aln; int64_t aln_cap; mc_u64* aln_bump;
	DpTask* tasks; int64_t task_cap; mc_u64* task_bump;
	int32_t* ptask; mc_u64* ptask_bump; int64_t ptask_begin;   // normal pieces whose strings still have to be laid out
	uint8_t* dpws; int64_t dpws_cap; mc_u64* dpws_bump;
	int32_t* rtask;          // rescue task list (pair ids)
	mc_u64* rtask_bump;
	struct RWin* rwin; mc_u64* rwin_bump; int64_t rwin_begin, rwin_cap;   // rescue windows of the current attempt: [rwin_begin, *rwin_bump)
	int32_t* rw_beg;         // per pair: first window of its rescue task
	int64_t rtask_begin, task_begin; // tasks of the current attempt are [begin, *bump)
	// profile
	DevProfile prof;
	int64_t first_read;      // global index of read 0 of this batch (parity of mate, dedup order)
};

MC_HD int64_t pa_pair_off(const PipeArgs& a, int64_t r) { return a.slot_loc[a.seed_off[r]]; }
MC_HD int64_t pa_cand_off_compute(const PipeArgs& a, int64_t r)
{
	if (!a.pr.paired) return pa_pair_off(a, r);
	int64_t r0 = r & ~1ll;
	int64_t base = 2 * pa_pair_off(a, r0);
	return (r & 1) ? base + (pa_pair_off(a, r0 + 2 <= a.n_reads ? r0 + 2 : a.n_reads) - pa_pair_off(a, r0)) : base;
}
MC_HD int64_t pa_cand_off(const PipeArgs& a, int64_t r) { return a.cand_off[r]; }
MC_HD int pa_cand_cap(const PipeArgs& a, int64_t r)
{
	if (!a.pr.paired) return (int)(pa_pair_off(a, r + 1) - pa_pair_off(a, r));
	int64_t r0 = r & ~1ll;
	return (int)(pa_pair_off(a, r0 + 2 <= a.n_reads ? r0 + 2 : a.n_reads) - pa_pair_off(a, r0));
}
MC_HD int pa_chunk_of_read(int64_t r) { return (int)(r / MC_CHUNK_READS); }

// ------------------------------------------------------------------------------------------------
// prep: mate 2 is reverse-complemented in place before seeding (reference src/ReadMapping.cpp:451)
// ------------------------------------------------------------------------------------------------
// one warp per mate 2: lane i swaps (and complements) bytes i and n-1-i, so both accesses of a warp are contiguous
MC_HD void prep_body(int64_t r, int lane, int nl, const PipeArgs& a)
{
	if (!a.pr.paired || !(r & 1)) return;
	uint8_t* s = a.seq + a.roff[r];
	const int n = (int)(a.roff[r + 1] - a.roff[r]);
	for (int i = lane; i < (n >> 1); i += nl) { const int j = n - 1 - i; const uint8_t x = mc_complement(s[j]), y = mc_complement(s[i]); s[i] = x; s[j] = y; }
	if ((n & 1) && lane == 0) s[n >> 1] = mc_complement(s[n >> 1]);
}

// ------------------------------------------------------------------------------------------------
// seeding: one thread walks one read left to right
// ------------------------------------------------------------------------------------------------
// read bases are fetched eight at a time through an aligned 64-bit window, and the following eight are requested as soon as
// a window is entered: the search only moves forward, so the next window is there when the walk reaches it (waiting for a
// read's own bases was 13 % of the stall samples of this kernel)
// The window may lie in shared memory (seed_walk stages a read of up to MC_SEED_STAGE_BASES bases in its lane's row when it opens
// it: the L2 round trip of every window refill was a quarter of the kernel's stall samples), hence plain loads, not __ldg.
struct BaseWindow { const uint64_t* w; uint64_t cur, nxt; int have; int shift0; };
#define MC_SEED_ROW_WORDS 21                                  /* 168 bytes per lane (an odd number of words: the lanes of a half-warp hit distinct banks), 42 KB per 256-thread block: room for a 150-base read at any alignment its offset can have, plus the word the window prefetches */
MC_HD void base_window_init(BaseWindow& bw, const uint8_t* s)
{
	bw.shift0 = (int)((uintptr_t)s & 7); bw.w = (const uint64_t*)(s - bw.shift0); bw.have = 0;
	bw.cur = bw.w[0]; bw.nxt = bw.w[1];     // the read arena is padded: a window past the last read stays inside it
}
MC_HD uint8_t base_at(const uint8_t* s, int p, BaseWindow& bw)
{
	const int q = p + bw.shift0;            // byte offset from the aligned base
	const int word = q >> 3;
	if (word != bw.have)
	{
		if (word == bw.have + 1) bw.cur = bw.nxt; else bw.cur = bw.w[word];
		bw.have = word; bw.nxt = bw.w[word + 1];
	}
	return (uint8_t)(bw.cur >> ((q & 7) << 3));
}
// the eight bytes of the read that start at base p (p lies in the current window: call base_at(s, p, bw) first)
MC_HD uint64_t bases8_at(int p, const BaseWindow& bw)
{
	const int off = ((p + bw.shift0) & 7) << 3;
	return off ? (bw.cur >> off) | (bw.nxt << (64 - off)) : bw.cur;
}
// eight ASCII bases -> sixteen bits of 2-bit codes (base i in bits 2i, 2i + 1; nst_nt4_table for ACGT / acgt); bit i of *bad is
// set when byte i is anything else.  All on the packed word: no per-base loop.
MC_HD uint32_t mc_pack8(uint64_t w, uint32_t* bad)
{
	const uint64_t u = w & 0xDFDFDFDFDFDFDFDFull, L7 = 0x7F7F7F7F7F7F7F7Full, H = 0x8080808080808080ull;
	const uint64_t a = u ^ 0x4141414141414141ull, c = u ^ 0x4343434343434343ull, g = u ^ 0x4747474747474747ull, t = u ^ 0x5454545454545454ull;
	// bit 7 of a byte of ~(((x & 0x7F) + 0x7F) | x) is set iff the byte is zero, exactly (no carries between bytes)
	const uint64_t ok = (~(((a & L7) + L7) | a) | ~(((c & L7) + L7) | c) | ~(((g & L7) + L7) | g) | ~(((t & L7) + L7) | t)) & H;
	*bad = (uint32_t)(((((ok ^ H) >> 7) * 0x0102040810204080ull) >> 56) & 0xFFu);
	uint64_t x = ((w >> 1) ^ (w >> 2)) & 0x0303030303030303ull;
	x = (x | (x >> 6)) & 0x000F000F000F000Full;
	x = (x | (x >> 12)) & 0x000000FF000000FFull;
	return (uint32_t)((x | (x >> 24)) & 0xFFFFu);
}
// four bytes of the packed reference (16 bases, the first one in the top bits of byte 0) -> base t in bits 2t, 2t + 1
MC_HD uint32_t mc_pac_norm(uint32_t x) { return ((x & 0x03030303u) << 6) | ((x & 0x0C0C0C0Cu) << 2) | ((x >> 2) & 0x0C0C0C0Cu) | ((x >> 6) & 0x03030303u); }
// The codes the next eight read bases must have for the pattern to keep occurring: E[i] = 3 - T[tq - 1 - i] in bits 2i, 2i + 1,
// T = forward strand followed by its reverse complement.  *avail = how many of them this call can vouch for (the text may end
// or change halves first).  tq >= 1.
MC_HD uint32_t text_expect8(const DevIndex& ix, int64_t tq, int* avail)
{
	const uint32_t* pac32 = (const uint32_t*)ix.pac;
	if (tq - 1 >= ix.G)
	{
		// reverse-complement half: T[j] = 3 - F[2G - 1 - j], so E[i] = F[(2G - tq) + i]: the forward strand read forwards
		const int64_t p0 = ix.twoG - tq, wi = p0 >> 4;
		const uint64_t x = (uint64_t)mc_pac_norm(mc_ldg(pac32 + wi)) | (uint64_t)mc_pac_norm(mc_ldg(pac32 + wi + 1)) << 32;
		const int64_t left = tq - ix.G; *avail = left < 8 ? (int)left : 8;
		return (uint32_t)(x >> (((uint32_t)p0 & 15u) << 1)) & 0xFFFFu;
	}
	// forward half: E[i] = 3 - F[tq - 1 - i]: the forward strand read backwards, complemented
	const int64_t p1 = tq - 1, a0 = p1 - 7, wi = a0 < 0 ? 0 : a0 >> 4;
	const uint64_t x = (uint64_t)mc_pac_norm(mc_ldg(pac32 + wi)) | (uint64_t)mc_pac_norm(mc_ldg(pac32 + wi + 1)) << 32;
	const int o = (int)(p1 - (wi << 4));                                       // 0..22: where F[p1] sits in x
	uint32_t v = (uint32_t)((x << ((31 - o) << 1)) >> 48);                     // F[p1 - 7 + t] in bits 2t
	v = ((v & 0x3333u) << 2) | ((v >> 2) & 0x3333u); v = ((v & 0x0F0Fu) << 4) | ((v >> 4) & 0x0F0Fu); v = ((v << 8) | (v >> 8)) & 0xFFFFu;   // F[p1 - i] in bits 2i
	*avail = tq < 8 ? (int)tq : 8;
	return v ^ 0xFFFFu;
}

template <class Interval> struct SeedOps;
// step(): extend v by read base c, or - `walk` - move the single row v.x1 one LF step (mc_fm_step32)
template <> struct SeedOps<RcInterval> {
	static MC_HD RcInterval init(const DevIndex& ix, int c) { return mc_interval_init(ix, c); }
	static MC_HD bool step(const DevIndex& ix, RcInterval& v, int c, bool walk, uint32_t* nblk)
	{
		if (!walk) return mc_interval_extend(ix, v, c, nblk);
		v.x1 = mc_lf_step(ix, v.x1); return true;
	}
};
template <> struct SeedOps<RcInterval32> {
	static MC_HD RcInterval32 init(const DevIndex& ix, int c) { return mc_interval_init32(ix, c); }
	static MC_HD bool step(const DevIndex& ix, RcInterval32& v, int c, bool walk, uint32_t* nblk)
	{
		if (walk && v.x1 == (uint32_t)ix.primary) { v.x1 = 0; return true; }      // bwt_invPsi of the row that holds the end marker
		return mc_fm_step32(ix, v, c, walk, nblk);
	}
};

#define MC_SEED_CMP_REPS 4
#define MC_SEED_LOCATED (1ull << 63)   // Seed::x0 of a seed whose single occurrence is already known: the text position, not a BWT row

// One thread walks one read left to right through the reference's greedy scheme (IdentifySimplePairs + BWT_Search,
// src/ReadMapping.cpp:125-158, src/bwt_search.cpp:121-151): from `pos`, extend while the pattern occurs; record it when it is
// >= 16 bases and occurs <= 50 times; continue one base behind where it stopped.
// A seed goes through up to four phases, one load per trip of the flat loop each:
//   start    the first k bases from the k-mer table (mc_fmindex.h), or a single base when the table cannot answer
//   step     one backward-search step on the reverse-complement interval per base - while the pattern occurs more than once
//   locate   the moment ONE occurrence is left (after ~15 bases in a 500 M-symbol text) its text position is looked up once:
//            LF steps to the next sampled row (3 on average with the sample kept in HBM)
//   compare  from then on extending the pattern is comparing the read with the packed reference text itself - sequential,
//            cached, eight bases per trip on packed words - instead of one random index block per base.  The search stops exactly where
//            the reference's does: a single row can only be extended by the base that precedes its suffix in the text.
// A seed that ends in the compare phase already knows its position (Seed::x0 = MC_SEED_LOCATED | position), so the locate
// kernel has nothing to walk for it.  The work counter `seed_blocks` stays the reference algorithm's: the table carries
// the block count of the steps it replaces and every compared base is charged the one block its step reads in the
// reference (an interval of one row straddles two blocks once in 128 steps: that 0.8 % is not counted).
// Lanes are persistent: lane `tid` of `nthreads` takes reads first + tid, first + tid + nthreads, ... and starts its next read in
// the trip after it finished one.  A read that lies in a repeat family keeps stepping through the index for all of its bases
// while a unique read is done after ~25 cheap trips; with one read per thread every warp would wait for its slowest read
// (practically every warp of 32 holds one), with ~26 reads per lane the difference averages out.
template <class Interval> MC_HD void seed_walk(int64_t tid, int64_t nthreads, int64_t first, int64_t n_end, const PipeArgs& a, uint64_t* row)
{
	int64_t r = first + tid;
	const uint8_t* s = nullptr; int rlen = 0, cap = 0, stop = 0; int64_t so = 0;
	BaseWindow bw; bw.w = nullptr; bw.cur = bw.nxt = 0; bw.have = 0; bw.shift0 = 0;
	int ns = 0, pos = 0, p = 0;
	uint32_t lower = 0;
	// phase of this lane: 4 = no reads left, -1 = read not yet opened, 0 between seeds, 1 stepping, 2 locating (v.x1 = the row
	// being walked, v.x2 == 1), 3 comparing
	int mode = r < n_end ? -1 : 4;
	uint32_t nblk = 0, nloc = 0, nsa = 0;
	Interval v; v.x1 = v.x2 = 0;
	uint64_t lsteps = 0;          // locate phase: steps taken
	int64_t tq = 0;               // compare phase: where the reverse complement of read[pos, p) lies in the text
	const bool direct = a.ix.sa_shift < 5;   // the denser suffix-array sample is there (mc_ctx_create)
	// One loop, one load per trip.  The lanes of a warp are in different reads, seeds and phases; every trip the warp runs the
	// ONE phase group most of its lanes are in - start (-1, 0), index step (1, 2) or compare (3) - and the others wait for
	// their turn, so that an instruction is executed for many lanes instead of each group's code for a few (with every present
	// group run in every trip 6 of 32 lanes were active per instruction).
	for (;;)
	{
#if MC_DEV_ONLY
		{
			const unsigned full = 0xffffffffu;
			const unsigned g0 = __ballot_sync(full, mode <= 0), g1 = __ballot_sync(full, mode == 1 || mode == 2), g3 = __ballot_sync(full, mode == 3);
			if (!(g0 | g1 | g3)) break;                                 // every lane is out of reads
			const int n0 = __popc(g0), n1 = __popc(g1), n3 = __popc(g3);
			const int pick = (n1 >= n0 && n1 >= n3) ? 1 : (n3 >= n0 ? 3 : 0);
			const int mine = mode <= 0 ? 0 : (mode == 3 ? 3 : (mode == 4 ? 4 : 1));
			if (mine != pick) continue;
		}
#else
		if (mode == 4) break;
#endif
		bool end = false;
		if (mode <= 0)
		{
			if (mode < 0)
			{
				s = a.seq + a.roff[r]; rlen = (int)(a.roff[r + 1] - a.roff[r]);
				so = a.seed_off[r]; cap = (int)(a.seed_off[r + 1] - so); stop = rlen - MC_MIN_SEED;
				{
					// the read into this lane's row of shared memory, as aligned 64-bit words (the row keeps the read's misalignment)
					const int sh = (int)((uintptr_t)s & 7); const uint64_t* src = (const uint64_t*)(s - sh);
					const int nw = ((sh + rlen + 7) >> 3) + 1;
					if (row && nw <= MC_SEED_ROW_WORDS)
					{
						for (int k = 0; k < nw; k++) row[k] = mc_ldg(src + k);
						s = (const uint8_t*)row + sh;
					}
				}
				base_window_init(bw, s);
				ns = 0; pos = 0; p = 0; lower = 0; mode = 0;
			}
			if (pos >= stop)
			{
				a.rflag[r] = (uint8_t)((lower >> 5) & 1);
				r += nthreads;
				mode = r < n_end ? -1 : 4;
				continue;
			}
			base_at(s, pos, bw);
			const uint64_t b0 = bases8_at(pos, bw);
			uint32_t bad0; const uint32_t c0 = mc_pack8(b0, &bad0);
			if (bad0 & 1u) { pos++; continue; }
			// the k-mer start table answers the first k bases with one load (mc_fmindex.h); at an N, an absent or an over-frequent
			// k-mer the search steps through them as the reference does
			if (a.ix.ktab_k)
			{
				const int K = a.ix.ktab_k;
				uint32_t m = c0, bad = bad0; uint64_t low = b0;
				if (K > 8) { base_at(s, pos + 8, bw); const uint64_t b1 = bases8_at(pos + 8, bw); uint32_t bad1; m |= mc_pack8(b1, &bad1) << 16; bad |= bad1 << 8; low |= b1 & ((1ull << ((K - 8) << 3)) - 1); }
				else if (K < 8) low &= (1ull << (K << 3)) - 1;
				if (!(bad & ((1u << K) - 1)) && KtabOps<Interval>::lookup(a.ix, m & (uint32_t)((1ull << (2 * K)) - 1), v, &nblk))
				{ if (low & 0x2020202020202020ull) lower |= 0x20u; p = pos + K; mode = 1; }
			}
			if (mode == 0) { lower |= (uint32_t)(b0 & 0xFF); v = SeedOps<Interval>::init(a.ix, (int)(c0 & 3u)); p = pos + 1; mode = 1; }
			if (direct && v.x2 == 1) { mode = 2; lsteps = 0; }
		}
		else if (mode <= 2)
		{
			// index step: extend by the next read base (1), or walk the single row towards a sampled one (2)
			const bool walk = mode == 2;
			if (walk && mc_sa_sampled(a.ix, (uint64_t)v.x1)) { tq = (int64_t)mc_sa_value(a.ix, (uint64_t)v.x1, lsteps); mode = 3; }
			else
			{
				int cc = 0; uint8_t ch = 0;
				if (!walk)
				{
					end = p >= rlen;
					if (!end) { ch = base_at(s, p, bw); cc = mc_nt4(ch); end = cc > 3; }
				}
				if (!end)
				{
					const bool ok = SeedOps<Interval>::step(a.ix, v, cc, walk, &nblk);
					if (walk) { lsteps++; nloc++; }
					else if (!ok) end = true;
					else { p++; lower |= ch; if (direct && v.x2 == 1 && p < rlen) { mode = 2; lsteps = 0; } }
				}
			}
		}
		else for (int rep = 0; rep < MC_SEED_CMP_REPS && !end; rep++)   // several 8-base comparisons per turn of the compare group: fewer turns per read, less waiting for the other groups
		{
			if (p >= rlen) end = true;
			else
			{
				base_at(s, p, bw);
				const uint64_t b = bases8_at(p, bw);
				uint32_t bad; const uint32_t rc = mc_pack8(b, &bad);
				if (bad & 1u) end = true;                              // an N ends the seed before the reference reads anything
				else if (tq <= 0) { nblk++; end = true; }              // the text starts here: the step reads its block and finds nothing
				else
				{
					int avail; const uint32_t ex = text_expect8(a.ix, tq, &avail);
					const int rem = rlen - p; int lim = rem < 8 ? rem : 8;
					if (bad) { const int fb = mc_ctz(bad); if (fb < lim) lim = fb; }
					uint32_t d = rc ^ ex; d = (d | (d >> 1)) & 0x5555u;
					int n = d ? mc_ctz(d) >> 1 : 8;                    // bases that agree
					const bool mismatch = n < lim && n < avail;
					if (n > lim) n = lim;
					if (n > avail) n = avail;
					if (n > 0 && (b & 0x2020202020202020ull & (n == 8 ? ~0ull : (1ull << (n << 3)) - 1))) lower |= 0x20u;
					p += n; tq -= n; nblk += (uint32_t)n;
					if (mismatch) { nblk++; end = true; }              // the failing step reads its block in the reference as well
					else if (n == lim && lim < 8) end = true;          // the read ends, or an N follows
				}
			}
		}
		if (end)
		{
			const int len = p - pos;
			if (len >= MC_MIN_SEED && v.x2 <= MC_MAX_OCC && ns < cap)
			{
				Seed sd; sd.x0 = mode == 3 ? (MC_SEED_LOCATED | (uint64_t)tq) : (uint64_t)v.x1; sd.read = (int32_t)r; sd.rpos = (int16_t)pos; sd.len = (int16_t)len;
				a.seeds[so + ns] = sd; a.slot_freq[so + ns] = (uint32_t)v.x2; ns++;
				if (mode == 3) nsa++;
			}
			pos = p + 1; mode = 0;
		}
	}
	if (nblk) mc_stat_add(&a.st->seed_blocks, (uint32_t)(nblk));
	if (direct) { mc_stat_add(&a.st->seed_locate_blocks, nloc); mc_stat_add(&a.st->seed_sa_reads, nsa); }
}
MC_HD void seed_body(int64_t tid, int64_t nthreads, int64_t first, int64_t n_end, const PipeArgs& a, uint64_t* row)
{
	if (a.ix.cbwt) seed_walk<RcInterval32>(tid, nthreads, first, n_end, a, row); else seed_walk<RcInterval>(tid, nthreads, first, n_end, a, row);
}

// BWT_Search as an operator (reference src/bwt_search.cpp:121-164) for independent (codes, start) queries: the search loop of
// seed_walk without the scan around it, followed by the locate of every hit.  The hits come out in the row order of the
// reverse-complement interval (mc_fmindex.h), i.e. as a permutation of the reference's LocArr.
struct SearchArgs { DevIndex ix; const uint8_t* codes; const int64_t* off; const int32_t* start; int32_t* len; int32_t* freq; uint64_t* loc; };
template <class Interval> MC_HD void search_walk(int64_t q, const SearchArgs& a)
{
	const uint8_t* s = a.codes + a.off[q];
	const int stop = (int)(a.off[q + 1] - a.off[q]), start = a.start[q];
	a.len[q] = 0; a.freq[q] = 0;
	if (start < 0 || start >= stop || s[start] > 3) return;
	uint32_t nblk = 0;
	Interval v = SeedOps<Interval>::init(a.ix, s[start]);
	int pos = start + 1;
	for (; pos < stop; pos++) { if (s[pos] > 3 || !mc_interval_extend(a.ix, v, s[pos], &nblk)) break; }
	const int len = pos - start;
	a.len[q] = len;
	if (len < MC_MIN_SEED || v.x2 > MC_MAX_OCC) return;
	a.freq[q] = (int32_t)v.x2;
	for (uint32_t i = 0; i < (uint32_t)v.x2; i++)
		a.loc[q * MC_MAX_OCC + i] = (uint64_t)a.ix.twoG - mc_locate(a.ix, (uint64_t)v.x1 + i, &nblk) - (uint64_t)len;
}
MC_HD void bwtsearch_body(int64_t q, const SearchArgs& a)
{
	if (a.ix.cbwt) search_walk<RcInterval32>(q, a); else search_walk<RcInterval>(q, a);
}

// ------------------------------------------------------------------------------------------------
// read ingest: FASTQ text -> the packed read arrays (reference GetNextEntry / GetNextChunk, FASTQ branch,
// src/GetData.cpp:32-99).  A record is four lines; the read is the second one and its length is the line's length
// without the newline (a '\r' before it would count as a base, as in the reference).
// ------------------------------------------------------------------------------------------------
#define MC_FQ_TILE 64
struct FastqArgs {
	const uint8_t* text[2]; int64_t len[2];       // one text per mate file; text[1] == 0: mates are adjacent records of text[0]
	uint32_t* tile_cnt[2]; const int64_t* tile_off[2]; int64_t* line_end[2]; int64_t n_lines[2];   // newline positions
	int64_t n_reads; uint32_t* rlen; int64_t* rsrc; const int64_t* roff; uint8_t* seq; mc_u64* flag;
	int lpr;                                      // lines per record: 4 (FASTQ) or 2 (FASTA with one sequence line per record)
};
// A tile is 64 bytes = four 128-bit loads per thread (the text buffer is padded to a multiple of 64); the newline bytes
// of a 32-bit word are found with the zero-byte trick on word ^ 0x0A0A0A0A.
MC_HD uint32_t fq_newline_mask(uint32_t w)      // bit 7 of every byte that is a newline
{
	const uint32_t v = w ^ 0x0A0A0A0Au;
	// exact per-byte zero test (no carries between bytes): a byte is zero iff its low seven bits and its top bit are zero
	return ~(((v & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | v) & 0x80808080u;
}
MC_HD void fqcount_body(int64_t t, int f, const FastqArgs& q)
{
	const int64_t b = t * MC_FQ_TILE;
	uint32_t n = 0;
	for (int k = 0; k < MC_FQ_TILE / 16; k++)
	{
		const mc_u32x4 v = mc_ldg128(q.text[f] + b + 16 * k);
		const uint32_t w[4] = {v.x, v.y, v.z, v.w};
		for (int j = 0; j < 4; j++)
		{
			uint32_t m = fq_newline_mask(w[j]);
			const int64_t pos = b + 16 * k + 4 * j;                 // bytes past the end of the text are padding
			if (pos + 4 > q.len[f]) { const int keep = (int)(q.len[f] - pos); m = keep <= 0 ? 0u : (m & (0xFFFFFFFFu >> (32 - 8 * keep))); }
			n += (uint32_t)mc_popc(m);
		}
	}
	q.tile_cnt[f][t] = n;
}
MC_HD void fqlines_body(int64_t t, int f, const FastqArgs& q)
{
	const int64_t b = t * MC_FQ_TILE;
	int64_t k = q.tile_off[f][t];
	if (q.tile_off[f][t + 1] == k) return;                           // no newline in this tile
	for (int c = 0; c < MC_FQ_TILE / 16; c++)
	{
		const mc_u32x4 v = mc_ldg128(q.text[f] + b + 16 * c);
		const uint32_t w[4] = {v.x, v.y, v.z, v.w};
		for (int j = 0; j < 4; j++)
		{
			const uint32_t m = fq_newline_mask(w[j]);
			if (!m) continue;
			const int64_t pos = b + 16 * c + 4 * j;
			for (int i = 0; i < 4; i++) if (((m >> (8 * i + 7)) & 1u) && pos + i < q.len[f]) q.line_end[f][k++] = pos + i;
		}
	}
}
// read r of the batch: record r (one file, mates adjacent) or record r/2 of file r&1; line_end[n_lines] = len (a last line
// without newline) is provided by the host when the block is the end of the file
MC_HD void fqread_body(int64_t r, const FastqArgs& q)
{
	const int f = q.text[1] ? (int)(r & 1) : 0;
	const int64_t rec = q.text[1] ? (r >> 1) : r;
	const int64_t beg = q.line_end[f][q.lpr * rec] + 1, end = q.line_end[f][q.lpr * rec + 1];
	int64_t n = end - beg;
	if (n <= 0 || n > MC_MAX_RLEN) { mc_atomic_or(q.flag, (mc_u64)1 << 56); n = 1; }
	// FASTA: every record is a '>' line and ONE line of bases; a wrapped record shifts a line of bases to where a header should be
	if (q.lpr == 2 && (q.text[f][q.line_end[f][2 * rec - 1] + 1] != '>' || q.text[f][beg] == '>')) mc_atomic_or(q.flag, (mc_u64)1 << 57);
	q.rlen[r] = (uint32_t)n; q.rsrc[r] = beg;
}
MC_HD void fqcopy_body(int64_t r, int lane, int nl, const FastqArgs& q)
{
	const int f = q.text[1] ? (int)(r & 1) : 0;
	const uint8_t* src = q.text[f] + q.rsrc[r];
	uint8_t* dst = q.seq + q.roff[r];
	const int n = (int)(q.roff[r + 1] - q.roff[r]);
	for (int i = lane; i < n; i += nl) dst[i] = src[i];
}

// one thread per slot lays out its locations: read offset, length and - in the place of the genome position - the BWT row
// the location starts from, so that the locate kernel needs a single 16-byte load per location
MC_HD void expand_body(int64_t s, const PipeArgs& a)
{
	const uint32_t f = a.slot_freq[s];
	if (!f) return;
	const int64_t o = a.slot_loc[s];
	const Seed sd = a.seeds[s];
	for (uint32_t i = 0; i < f; i++) { SPair p; p.gpos = (int64_t)(sd.x0 + i); p.rpos = sd.rpos; p.len = sd.len; a.pairs[o + i] = p; }
}

// Locations: every lane walks LF steps (bwt_sa, src/bwt_search.cpp:109-119; mc_sa_value for the sampling in HBM) and, the moment its row is a sampled one,
// finishes that location and picks up its next one (locations tid, tid + nthreads, ...), so the lanes of a warp keep
// stepping together although the walks have very different lengths (0..31+ steps).
MC_HD void locate_body(int64_t tid, int64_t nthreads, const PipeArgs& a)
{
	int64_t t = tid;
	if (t >= a.n_locs) return;
	uint32_t nblk = 0, nsa = 0;
	SPair p = a.pairs[t];
	uint64_t k = (uint64_t)p.gpos, steps = 0;
	bool live = true;
	while (live)
	{
		const bool known = (k & MC_SEED_LOCATED) != 0;   // the seed kernel already compared this seed against the text (seed_walk)
		const bool at = known || mc_sa_sampled(a.ix, k);
		if (at)
		{
			// the seed carries the rows of its reverse complement: an occurrence of that at q is the seed at 2G - q - len
			const uint64_t q = known ? (k & ~MC_SEED_LOCATED) : mc_sa_value(a.ix, k, steps);
			const int64_t g = (int64_t)((uint64_t)a.ix.twoG - q - (uint64_t)p.len);
			p.gpos = g;
			if (g - (int64_t)p.rpos <= 0) p.len = 0;                    // PosDiff <= 0 is dropped (src/ReadMapping.cpp:145)
			a.pairs[t] = p;
			if (!known) nsa++;
			t += nthreads;
			live = t < a.n_locs;
			if (live) { p = a.pairs[t]; k = (uint64_t)p.gpos; steps = 0; }
		}
		if (live && !(k & MC_SEED_LOCATED) && !mc_sa_sampled(a.ix, k)) { k = mc_lf_step(a.ix, k); steps++; nblk++; }
	}
	mc_stat_add(&a.st->locate_blocks, (uint32_t)(nblk));
	mc_stat_add(&a.st->sa_reads, (uint32_t)(nsa));
}

// ------------------------------------------------------------------------------------------------
// clustering: sort the read's simple pairs by (PosDiff, rPos), cut into clusters, apply the
// ratcheting score threshold, resolve tandem repeats
// ------------------------------------------------------------------------------------------------
MC_HD bool sp_less_posdiff(const SPair& x, const SPair& y)
{
	int64_t dx = x.gpos - x.rpos, dy = y.gpos - y.rpos;
	return dx == dy ? x.rpos < y.rpos : dx < dy;
}

MC_HD void cluster_body(int64_t r, const PipeArgs& a)
{
	const int rlen = (int)(a.roff[r + 1] - a.roff[r]);
	const int64_t po = pa_pair_off(a, r);
	const int raw = (int)(pa_pair_off(a, r + 1) - po);
	SPair* v = a.pairs + po;
	int m = 0;
	for (int i = 0; i < raw; i++) if (v[i].len > 0) { if (m != i) v[m] = v[i]; m++; }
	for (int i = 1; i < m; i++) // insertion sort; m is 1-3 for almost every read
	{
		SPair x = v[i]; int j = i - 1;
		while (j >= 0 && sp_less_posdiff(x, v[j])) { v[j + 1] = v[j]; j--; }
		v[j + 1] = x;
	}
	a.npair[r] = m;
	const int64_t co = pa_cand_off_compute(a, r);
	a.cand_off[r] = (int32_t)co;
	int nc = 0;
	if (m > 0)
	{
		int head = 0, score = v[0].len, thr = rlen >> 2;
		int64_t gend = a.ix.chrom_end[mc_chrom_lower_bound(a.ix, v[0].gpos)];
		for (int i = 0, j = 1; j <= m; i++, j++)
		{
			bool cut;
			if (j == m) cut = true; // the terminal pair sits at 2G, beyond every chromosome end
			else
			{
				int64_t d = (v[j].gpos - v[j].rpos) - (v[i].gpos - v[i].rpos);
				if (d < 0) d = -d;
				cut = v[j].gpos > gend || d > a.pr.max_pos_diff;
			}
			if (cut)
			{
				if (score > thr)
				{
					if (thr < (score >> 1)) thr = score >> 1;
					Cand c; c.score = score; c.pbeg = (int32_t)(po + head); c.pend = (int32_t)(po + j);
					if (score >= rlen)
					{
						// tandem repeat: keep only the best run of equal PosDiff, first maximum wins
						int bi = head, bj = head, bs = 0, ri = head, s = v[head].len;
						for (int k = head + 1; k < j; k++)
						{
							if ((v[k].gpos - v[k].rpos) != (v[ri].gpos - v[ri].rpos))
							{
								if (s > bs) { bs = s; bi = ri; bj = k; }
								ri = k; s = v[k].len;
							}
							else s += v[k].len;
						}
						if (s > bs) { bs = s; bi = ri; bj = j; }
						c.score = bs; c.pbeg = (int32_t)(po + bi); c.pend = (int32_t)(po + bj);
					}
					a.cands[co + nc] = c; nc++;
				}
				if (j < m)
				{
					head = j; score = v[j].len;
					gend = a.ix.chrom_end[mc_chrom_lower_bound(a.ix, v[j].gpos)];
				}
			}
			else score += v[j].len;
		}
	}
	a.ncand0[r] = nc;
}

#include "mc_stages_pair.h"
#include "mc_stages_align.h"
#include "mc_stages_profile.h"
#include "mc_stages_vc.h"
#include "mc_stages_sam.h"

#endif
