// Variant-calling scan over the resident profile: the per-column part of the reference's
// VariantCalling() - CalBlockReadDepth (src/VariantCalling.cpp:106-120) and IdentifyVariants (:550-680) with
// DetermineGenotype (:523-548), GetRefCount (:511-521) and CheckDiploidFrequency (:122-127).
//
// One thread walks one block of 100 columns (the reference's BlockSize, so the depth-derived thresholds are constant
// per thread).  What the reference carries from column to column in locals is carried across blocks by scans:
//   gap / dup run lengths   <- inclusive max-scan of "last column that is not a gap (dup) column" per block
//   "last pushed is a NOR"  <- inclusive max-scans of the last normal column and the last event column per block
//   min depth of a NOR run  <- the thread that opens a run walks the following blocks' leading minima until an event
// Records go to slots reserved per block (exclusive scan of an upper bound); unused slots keep VarType 255.
#ifndef MC_STAGES_VC_H
#define MC_STAGES_VC_H
#include <math.h>

#define MC_VC_BLOCK 100        // BlockSize, src/VariantCalling.cpp:4
#define MC_VAR_NIL 255

struct VcCand { int64_t pos; int32_t kind, freq, alt_off, alt_len; };   // an indel window winner (GetAreaIndFrequency != 0)

struct VcArgs {
	mc_vc_params vp; DevIndex ix;
	int64_t G, n_blocks;
	const uint64_t* recs; int64_t tile_beg, tile_end;   // packed MappingRecord_t of [tile_beg, tile_end)
	int32_t* depth;                                       // [n_blocks] BlockDepthArr
	int64_t *last_nongap, *last_nondup;                   // [n_blocks] per block, then inclusive max-scanned
	int64_t *last_normal, *last_event;                    // [n_blocks] per block, then inclusive max-scanned (gvcf)
	int32_t* lead_min;                                    // [n_blocks] min depth of the normal columns before the block's first event
	uint32_t* cnt; const int64_t* off;                    // [n_blocks] upper bound of records per block, its exclusive scan
	const VcCand* cand; int64_t n_cand;                   // sorted by (pos, kind)
	mc_variant_rec* out;
};

// How a walker gets at the packed record of column j of its block.  On the device the 1600 bytes of a block are not read by
// their thread 16 bytes at a time (every lane of a warp in a different line): the thread block stages MC_VC_STAGE columns of
// all its walkers through shared memory with coalesced loads (mc_launch.h: VcFetchShared); sync(j) is called by EVERY thread
// of the block at the top of EVERY one of the MC_VC_BLOCK iterations, whether the walker has a column there or not.
#define MC_VC_STAGE 20
struct VcFetchGlobal {
	const uint64_t* recs;        // record of the walker's column 0
	MC_HD void sync(int) {}
	MC_HD void get(int j, uint64_t& w0, uint64_t& w1) const { w0 = recs[2 * j]; w1 = recs[2 * j + 1]; }
};

MC_HD int vc_cov(uint64_t w) { return (int)((w & 4095) + ((w >> 12) & 4095) + ((w >> 24) & 4095) + ((w >> 36) & 4095)); }

// pass 1: BlockDepthArr and the run carriers
template <class Fetch> MC_HD void vcdepth_walk(int64_t b, bool live, const VcArgs& a, Fetch& f)
{
	const int64_t g0 = b * MC_VC_BLOCK; int64_t g1 = g0 + MC_VC_BLOCK; if (g1 > a.G) g1 = a.G;
	if (!live) g1 = g0;
	int sum = 0; int64_t ng = -1, nd = -1;
	for (int j = 0; j < MC_VC_BLOCK; j++)
	{
		f.sync(j);
		const int64_t g = g0 + j;
		if (g >= g1) continue;
		uint64_t w, w1_unused; f.get(j, w, w1_unused);
		const int cov = vc_cov(w); const int mh = (int)((w >> 48) & 4095);
		sum += cov;
		if (!(cov == 0 && mh == 0)) ng = g;
		if (!(cov == 0 && mh > 0)) nd = g;
	}
	if (!live) return;
	a.depth[b] = sum > 0 ? sum / MC_VC_BLOCK : 0;   // :117, always / BlockSize
	a.last_nongap[b] = ng; a.last_nondup[b] = nd;
}
MC_HD void vcdepth_body(int64_t b, const VcArgs& a)
{
	VcFetchGlobal f; f.recs = a.recs + 2 * (b * MC_VC_BLOCK - a.tile_beg);
	vcdepth_walk(b, true, a, f);
}

// DetermineGenotype, src/VariantCalling.cpp:523-548
MC_HD int vc_genotype(int ploidy, int cov, int alt_reads, int alt_num)
{
	if (ploidy == 1) return alt_reads < (int)(cov * 0.50) ? 1 : 2;
	if (ploidy == 2)
	{
		if (alt_num == 0) return 3;
		if (alt_num == 1) return alt_reads < (int)(cov * 0.50) ? 4 : 5;
		if (alt_num == 2) return 6;
	}
	return 0;
}
// (uint8_t)(int)x as x86-64 does it; a zero divisor gives inf / nan -> "integer indefinite" 0x80000000 -> low byte 0
MC_HD uint8_t vc_q8(double num, double den) { if (den == 0.0) return 0; return (uint8_t)(int)(num / den); }

MC_HD void vc_clear(mc_variant_rec& v) { v.gPos = 0; v.record[0] = v.record[1] = 0; v.alt_off = v.alt_len = 0; v.DP = v.AD_ref = v.AD_alt = 0; v.GenoType = v.qscore = 0; v.VarType = MC_VAR_NIL; v.alt[0] = v.alt[1] = v.alt[2] = 0; }

// passes 2 (emit == false: counts and the gvcf carriers) and 3 (emit == true)
template <class Fetch> MC_HD void vcscan_walk(int64_t b, bool live, const VcArgs& a, bool emit, Fetch& f)
{
	if (!live) b = 0;            // a padding thread of the last block: takes part in the staging only
	const int64_t g0 = b * MC_VC_BLOCK; int64_t g1 = g0 + MC_VC_BLOCK; if (g1 > a.G) g1 = a.G;
	if (!live) g1 = g0;
	const mc_vc_params& P = a.vp;
	const int mad = P.min_allele_depth, depth = a.depth[b];
	int cov_thr = depth >> 1; if (cov_thr < mad) cov_thr = mad;
	if (P.somatic && cov_thr > mad) cov_thr = mad;
	int ins_thr = (int)(cov_thr * 0.25); if (ins_thr < mad) ins_thr = mad;
	int del_thr = (int)(cov_thr * 0.35); if (del_thr < mad) del_thr = mad;
	const double fthr = P.somatic ? 0.01 : (double)P.frequency_thr;
	int64_t gap = 0, dup = 0;
	if (b > 0) { gap = g0 - 1 - a.last_nongap[b - 1]; dup = g0 - 1 - a.last_nondup[b - 1]; }
	// candidates of this block
	int64_t ci; { int64_t lo = 0, hi = a.n_cand; while (lo < hi) { int64_t mid = (lo + hi) >> 1; if (a.cand[mid].pos < g0) lo = mid + 1; else hi = mid; } ci = lo; }
	int64_t next_cand = ci < a.n_cand ? a.cand[ci].pos : -1;   // column of the next candidate: one register compare per column instead of a load
	bool last_nor = false;
	if (emit && P.gvcf && b > 0) { const int64_t ln = a.last_normal[b - 1], le = a.last_event[b - 1]; last_nor = ln >= 0 && ln >= le; }
	int64_t slot = emit ? a.off[b] : 0; const int64_t slot_end = emit ? a.off[b] + a.cnt[b] : 0;
	uint32_t n = 0; int64_t ln_pos = -1, le_pos = -1; int lead = 0x7fffffff;
	int64_t open_slot = -1; int open_min = 0;
	mc_variant_rec v;
#define VC_PUSH() do { if (emit) a.out[slot++] = v; n++; } while (0)
#define VC_EVENT(g) do { le_pos = (g); last_nor = false; if (emit && open_slot >= 0) { a.out[open_slot].AD_alt = (uint16_t)open_min; open_slot = -1; } } while (0)
	for (int j = 0; j < MC_VC_BLOCK; j++)
	{
		f.sync(j);
		const int64_t g = g0 + j;
		if (g >= g1) continue;
		uint64_t w0, w1; f.get(j, w0, w1);
		int c[4] = {(int)(w0 & 4095), (int)((w0 >> 12) & 4095), (int)((w0 >> 24) & 4095), (int)((w0 >> 36) & 4095)};
		const int mh = (int)((w0 >> 48) & 4095), cov = c[0] + c[1] + c[2] + c[3];
		bool normal = true;
		for (; next_cand == g; ci++, next_cand = ci < a.n_cand ? a.cand[ci].pos : -1)   // :571-592, insertion before deletion
		{
			const VcCand k = a.cand[ci];
			if (k.freq < (k.kind ? del_thr : ins_thr)) continue;
			vc_clear(v); v.gPos = g; v.record[0] = w0; v.record[1] = w1; v.VarType = k.kind ? MC_VAR_DEL : MC_VAR_INS;
			v.DP = (uint16_t)depth; v.AD_alt = (uint16_t)k.freq; if (v.DP < v.AD_alt) v.DP = v.AD_alt;
			v.alt_off = k.alt_off; v.alt_len = k.alt_len;
			v.AD_ref = (uint16_t)(v.DP - v.AD_alt); v.GenoType = (uint8_t)vc_genotype(P.ploidy, v.DP, v.AD_alt, 1);
			v.qscore = vc_q8(100.0 * v.AD_alt, (double)cov);
			normal = false; VC_EVENT(g); VC_PUSH();
		}
		if (cov >= cov_thr)   // :594-626
		{
			int freq_thr = (int)ceil(cov * fthr); if (freq_thr < mad) freq_thr = mad;
			const int rb = mc_ref_code(a.ix, g);
			int na = 0, ai[4];
			for (int k = 0; k < 4; k++) if (rb != k && c[k] >= freq_thr) ai[na++] = k;
			if (na == 1 || (na == 2 && c[ai[0]] + c[ai[1]] >= (int)(cov * 0.50)))
			{
				vc_clear(v); v.gPos = g; v.record[0] = w0; v.record[1] = w1; v.VarType = MC_VAR_SUB; v.AD_ref = (uint16_t)c[rb];
				v.DP = (uint16_t)cov; v.AD_alt = (uint16_t)(na == 1 ? c[ai[0]] : c[ai[0]] + c[ai[1]]);
				v.GenoType = (uint8_t)vc_genotype(P.ploidy, cov, v.AD_alt, na);
				if (v.GenoType != 0)
				{
					v.alt[0] = "ACGT"[ai[0]]; if (na == 2) { v.alt[1] = ','; v.alt[2] = "ACGT"[ai[1]]; }
					v.qscore = P.somatic ? vc_q8(35.0 * v.AD_alt, cov * 0.05) : vc_q8(35.0 * v.AD_alt, (double)cov);
					normal = false; VC_EVENT(g); VC_PUSH();
				}
			}
		}
		if (cov == 0 && mh == 0) { normal = false; gap++; }   // :627-636
		else if (gap > 0)
		{
			if (gap >= P.min_unmapped_size) { vc_clear(v); v.VarType = MC_VAR_UMR; v.gPos = g - gap; v.DP = (uint16_t)gap; VC_EVENT(g); VC_PUSH(); }
			gap = 0;
		}
		if (cov == 0 && mh > 0) { normal = false; dup++; }   // :637-646
		else if (dup > 0)
		{
			if (dup > P.min_cnv_size) { vc_clear(v); v.VarType = MC_VAR_CNV; v.gPos = g - dup; v.DP = (uint16_t)dup; VC_EVENT(g); VC_PUSH(); }
			dup = 0;
		}
		if (normal && cov > 0)
		{
			if (P.gvcf)   // :647-658
			{
				ln_pos = g;
				if (le_pos < 0 && cov < lead) lead = cov;
				if (!emit) { if (le_pos < 0 ? n == 0 : !last_nor) { n++; last_nor = true; } }   // the leading run costs one slot whether or not it opens here
				else if (!last_nor)
				{
					vc_clear(v); v.gPos = g; v.record[0] = w0; v.record[1] = w1; v.VarType = MC_VAR_NOR; v.DP = v.AD_alt = (uint16_t)cov;
					open_slot = slot; open_min = cov; a.out[slot++] = v; last_nor = true;
				}
				else if (open_slot >= 0 && cov < open_min) open_min = cov;
			}
			else if (P.monomorphic)   // :659-665
			{
				vc_clear(v); v.gPos = g; v.record[0] = w0; v.record[1] = w1; v.VarType = MC_VAR_MON; v.DP = (uint16_t)cov;
				v.GenoType = (uint8_t)vc_genotype(P.ploidy, cov, 0, 0); v.AD_ref = (uint16_t)c[mc_ref_code(a.ix, g)];
				VC_PUSH();
			}
		}
	}
#undef VC_PUSH
#undef VC_EVENT
	if (!live) return;
	if (!emit) { a.cnt[b] = n; if (P.gvcf) { a.last_normal[b] = ln_pos; a.last_event[b] = le_pos; a.lead_min[b] = lead; } return; }
	if (open_slot >= 0)   // the run is still open at the end of the block: it lasts until the next event
	{
		for (int64_t nb = b + 1; nb < a.n_blocks; nb++)
		{
			const int m = a.lead_min[nb]; if (m < open_min) open_min = m;
			if (a.last_event[nb] > a.last_event[nb - 1]) break;   // inclusive max-scanned: grows iff block nb has an event
		}
		a.out[open_slot].AD_alt = (uint16_t)open_min;
	}
	for (; slot < slot_end; slot++) { vc_clear(v); a.out[slot] = v; }
}

MC_HD void vcscan_body(int64_t b, const VcArgs& a, bool emit)
{
	VcFetchGlobal f; f.recs = a.recs + 2 * (b * MC_VC_BLOCK - a.tile_beg);
	vcscan_walk(b, true, a, emit, f);
}

// The final order of the records (CompByVarPos, src/VariantCalling.cpp:51-55: gPos, then VarType - unique within one scan) is
// made on the device: the slots are in column order except for gap / dup records, which sit where their run ends but carry its
// start, and for the INS, DEL, SUB order of one column; unused slots sort to the end.
MC_HD void vckey_body(int64_t i, const mc_variant_rec* recs, uint64_t* keys, uint32_t* idx, mc_u64* n_valid)
{
	const bool ok = recs[i].VarType != MC_VAR_NIL;
	keys[i] = ok ? ((uint64_t)recs[i].gPos << 8 | (uint64_t)recs[i].VarType) : ~0ull;
	idx[i] = (uint32_t)i;
	mc_stat_add(n_valid, ok ? 1u : 0u);
}
MC_HD void vcgather_body(int64_t j, const mc_variant_rec* in, const uint32_t* idx, mc_variant_rec* out) { out[j] = in[idx[j]]; }

#endif
