// Pile-up profile update (included by mc_stages.h).
//
// Replaces UpdateProfile / UpdateMultiHitCount (reference src/AlignmentProfile.cpp:41-242,244-271).
// The reference serialises the whole update behind ProfileLock; the only part that really is
// order-dependent is the PCR-duplicate gate (readCount[start] < iMaxDuplicate, :76-77), which admits the
// FIRST iMaxDuplicate reads per start position in file order.  Here the gate is decided from a sort of
// (start position, read index) keys; everything else is commutative saturating / wrapping counting and
// is scattered with atomics into the device-resident counters.
#ifndef MC_STAGES_PROFILE_H
#define MC_STAGES_PROFILE_H

#define MC_KEY_SHIFT 28   // key = start << 28 | read index inside the batch

struct ProfArgs {
	uint64_t* keys; mc_u64* key_bump;        // gate candidates of this batch
	uint8_t* accept;                         // per read: 0 skip, 1 UpdateProfile, 2 UpdateMultiHitCount
	int32_t* rnp;                            // per read with a gate key: its live candidate (<< 16) and how many of its pieces have aligned columns
	int64_t n_keys;
	int64_t* bp_pos; mc_u64* bp_bump; int64_t bp_cap;                  // BreakPointMap increments (persistent)
	mc_indel_rec* ind; mc_u64* ind_bump; int64_t ind_cap;              // InsertSeqMap / DeleteSeqMap increments (persistent)
	uint8_t* ind_seq; mc_u64* ind_seq_bump; int64_t ind_seq_cap;
};

// one thread per read: which update applies, clip / break-point bookkeeping, gate key
// the read's gate key (start << 28 | read), or ~0 when it has none
MC_HD uint64_t profkey_of(int64_t r, const PipeArgs& a, const ProfArgs& q)
{
	const uint64_t none = ~(uint64_t)0;
	q.accept[r] = 0;
	const ReadSum sum = a.rsum[r];
	if (sum.score == 0) return none;
	if (sum.n_live != 1) { q.accept[r] = 2; return none; }
	const int rlen = (int)(a.roff[r + 1] - a.roff[r]);
	const int64_t co = pa_cand_off(a, r);
	const int nc = a.ncand[r];
	int ci = 0; while (ci < nc && a.cscore[co + ci] == 0) ci++;
	const mc_frag_out* f = a.frags + a.cfrag[co + ci];
	const int nf = a.cnfrag[co + ci];
	const mc_frag_out first = f[0], last = f[nf - 1];
	if (first.rLen == 0 && first.gLen == 0)
	{
		if (first.rPos > 20)
		{
			int64_t k = (int64_t)mc_atomic_add(q.bp_bump, (mc_u64)1);
			if (k < q.bp_cap) q.bp_pos[k] = first.gPos < a.ix.G ? first.gPos : a.ix.twoG - 1 - first.gPos; else mc_atomic_or(&a.st->overflow, (mc_u64)1 << 48);
		}
		if (first.rPos > a.pr.max_clip) return none;
	}
	if (last.rLen == 0 && last.gLen == 0)
	{
		if (rlen - last.rPos > 20)
		{
			int64_t k = (int64_t)mc_atomic_add(q.bp_bump, (mc_u64)1);
			if (k < q.bp_cap) q.bp_pos[k] = last.gPos < a.ix.G ? last.gPos : a.ix.twoG - 1 - last.gPos; else mc_atomic_or(&a.st->overflow, (mc_u64)1 << 48);
		}
		if (rlen - last.rPos > a.pr.max_clip) return none;
	}
	const int64_t start = a.corient[co + ci] ? first.gPos : a.ix.twoG - (first.gPos + first.gLen);
	if (start < 0 || start >= a.ix.G) return none; // the reference would index outside MappingRecordArr
	// what scatter_body needs before its block-wide cursor bump, so that it does not have to walk read -> candidate -> fragments first
	int np = 0;
	for (int k = 0; k < nf; k++) if (!f[k].bSimple && f[k].gLen != 0 && f[k].rLen != 0) np++;
	q.rnp[r] = (int32_t)(((uint32_t)ci << 16) | (uint32_t)(np > 0xFFFF ? 0xFFFF : np));
	return ((uint64_t)start << MC_KEY_SHIFT) | (uint64_t)r;
}
MC_HD void profkey_body(int64_t r, bool live, const PipeArgs& a, const ProfArgs& q)
{
	const uint64_t key = live ? profkey_of(r, a, q) : ~(uint64_t)0;
	const bool has = key != ~(uint64_t)0;
	const int64_t k = mc_block_bump(q.key_bump, has ? 1u : 0u);
	if (has) q.keys[k] = key;
}

// keys sorted ascending: entry i passes iff fewer than `remaining` earlier entries share its start
MC_HD void gate_body(int64_t i, const PipeArgs& a, const ProfArgs& q)
{
	const uint64_t key = q.keys[i];
	const int64_t s = (int64_t)(key >> MC_KEY_SHIFT);
	const int64_t r = (int64_t)(key & ((1ull << MC_KEY_SHIFT) - 1));
	const int remaining = a.pr.max_dup - (int)a.prof.rcount[s];
	const bool ok = remaining > 0 && (i < remaining || (int64_t)(q.keys[i - remaining] >> MC_KEY_SHIFT) != s);
	q.accept[r] = ok ? 1 : 0;
}

// segment heads bump the persistent per-position count (after every gate_body has read it)
MC_HD void gateupd_body(int64_t i, const PipeArgs& a, const ProfArgs& q)
{
	const int64_t s = (int64_t)(q.keys[i] >> MC_KEY_SHIFT);
	if (i > 0 && (int64_t)(q.keys[i - 1] >> MC_KEY_SHIFT) == s) return;
	int prior = a.prof.rcount[s], n = 0;
	while (prior + n < a.pr.max_dup && i + n < q.n_keys && (int64_t)(q.keys[i + n] >> MC_KEY_SHIFT) == s) n++;
	a.prof.rcount[s] = (uint8_t)(prior + n);
}

// Several GPUs mapping consecutive shards of one library (mc_ctx.cu, ordered exchange): every rank lists how many gate
// candidates it has per start column (heads of the sorted key runs, capped at 15 = the 4-bit readCount) ...
MC_HD void gatecnt_body(int64_t i, const PipeArgs& a, const ProfArgs& q, uint64_t* list, mc_u64* bump)
{
	const int64_t s = (int64_t)(q.keys[i] >> MC_KEY_SHIFT);
	if (i > 0 && (int64_t)(q.keys[i - 1] >> MC_KEY_SHIFT) == s) return;
	int n = 1;
	while (n < 15 && i + n < q.n_keys && (int64_t)(q.keys[i + n] >> MC_KEY_SHIFT) == s) n++;
	const int64_t k = mc_bump_alloc(bump, 1u);
	list[k] = ((uint64_t)s << 8) | (uint64_t)n;
}
// ... and the lists of the ranks before this one in file order are added to readCount before the gate is evaluated, those
// of the ranks after it afterwards: readCount[start] = min(iMaxDuplicate, earlier + own + later), the same on every rank.
// A list holds every start once, so the plain read-modify-write is race free.
MC_HD void gateadd_body(int64_t i, const PipeArgs& a, const uint64_t* list)
{
	const int64_t s = (int64_t)(list[i] >> 8);
	const int v = (int)a.prof.rcount[s] + (int)(list[i] & 0xFF);
	a.prof.rcount[s] = (uint8_t)(v < a.pr.max_dup ? v : a.pr.max_dup);
}

// the same exchange with one byte per column (genomes whose size is comparable to the number of keys): fill ...
MC_HD void gatedense_fill_body(int64_t i, const PipeArgs& a, const ProfArgs& q, uint8_t* dense)
{
	const int64_t s = (int64_t)(q.keys[i] >> MC_KEY_SHIFT);
	if (i > 0 && (int64_t)(q.keys[i - 1] >> MC_KEY_SHIFT) == s) return;
	int n = 1;
	while (n < 15 && i + n < q.n_keys && (int64_t)(q.keys[i + n] >> MC_KEY_SHIFT) == s) n++;
	dense[s] = (uint8_t)n;
}
// ... and add the counts of ranks [r0, r1) to readCount, column by column (streaming)
MC_HD void gatedense_apply_body(int64_t g, const PipeArgs& a, const uint8_t* all, size_t pitch, int r0, int r1)
{
	int v = 0;
	for (int r = r0; r < r1; r++) v += all[(size_t)r * pitch + g];
	if (!v) return;
	v += a.prof.rcount[g];
	a.prof.rcount[g] = (uint8_t)(v < a.pr.max_dup ? v : a.pr.max_dup);
}

MC_HD void prof_base(const PipeArgs& a, int64_t g, int field) // field: 0 A, 1 C, 2 G, 3 T
{
	if (g < 0 || g >= a.ix.G) return;
	mc_atomic_add(a.prof.base16 + g * 2 + (field >> 1), (uint32_t)(1u << ((field & 1) << 4)));
}
// +1 on every column of [g0, g1) of a difference array with `stride` interleaved lanes
MC_HD void prof_range(const PipeArgs& a, int32_t* diff, int stride, int lane, int64_t g0, int64_t g1)
{
	if (g0 < 0) g0 = 0;
	if (g1 > a.ix.G) g1 = a.ix.G;
	if (g0 >= g1) return;
	mc_atomic_add(diff + g0 * stride + lane, 1);
	mc_atomic_add(diff + g1 * stride + lane, -1);
}
MC_HD int base_field(uint8_t c) { return c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : c == 'T' ? 3 : -1; }

MC_HD void indel_emit(const PipeArgs& a, const ProfArgs& q, int kind, int64_t pos, const uint8_t* s, int len)
{
	const int64_t k = mc_bump_alloc(q.ind_bump, 1u);
	const int64_t o = mc_bump_alloc(q.ind_seq_bump, (uint32_t)len);
	if (k >= q.ind_cap || o + len > q.ind_seq_cap) { mc_atomic_or(&a.st->overflow, (mc_u64)1 << 48); return; }
	mc_indel_rec rec; rec.pos = pos; rec.kind = kind; rec.len = len; rec.count = 1; rec.seq_off = (int32_t)o;
	q.ind[k] = rec;
	for (int i = 0; i < len; i++) q.ind_seq[o + i] = s[i];
}

// one thread per read.  An accepted read costs two atomics for its strand coverage, two per exact-match seed and one
// per base of its gapped / mismatching pieces.
MC_HD void scatter_body(int64_t r, bool live, const PipeArgs& a, const ProfArgs& q)
{
	const int mode = live ? q.accept[r] : 0;
	const int rlen = mode ? (int)(a.roff[r + 1] - a.roff[r]) : 0;
	const uint8_t* rs = mode ? a.seq + a.roff[r] : nullptr;
	const int64_t co = mode ? pa_cand_off(a, r) : 0;
	const int nc = mode ? a.ncand[r] : 0;
	if (mode == 2)
	{
		for (int ci = 0; ci < nc; ci++)
		{
			if (a.cscore[co + ci] <= 0) continue;
			const mc_frag_out* f = a.frags + a.cfrag[co + ci];
			const int nf = a.cnfrag[co + ci];
			int64_t g0, g1;
			if (a.corient[co + ci]) { g0 = f[0].gPos; g1 = f[nf - 1].gPos + f[nf - 1].gLen; }
			else { g0 = a.ix.twoG - (f[0].gPos + f[0].gLen); g1 = a.ix.twoG - f[nf - 1].gPos; }
			prof_range(a, a.prof.mdiff, 1, 0, g0, g1);
			if (g1 > g0) mc_stat_add(&a.st->profile_columns, (uint32_t)(g1 - g0));
		}
	}
	// pieces with aligned columns are left to profpiece_body (a tile per piece): their queue slots come from one block-wide
	// cursor bump, in which every thread of the block takes part - with the count profkey_of left behind (one coalesced load), so
	// that no thread waits at the barrier for another one's chain of dependent loads
	const uint32_t hint = mode == 1 ? (uint32_t)q.rnp[r] : 0u;
	int ci = (int)(hint >> 16), np = (int)(hint & 0xFFFFu);
	const mc_frag_out* f = nullptr; int nf = 0;
	if (mode == 1 && np == 0xFFFF)            // (more pieces than the hint can say: count them)
	{
		f = a.frags + a.cfrag[co + ci]; nf = a.cnfrag[co + ci];
		np = 0; for (int k = 0; k < nf; k++) if (!f[k].bSimple && f[k].gLen != 0 && f[k].rLen != 0) np++;
	}
	int64_t pt = mc_block_bump(a.ptask_bump, (uint32_t)np);
	if (mode != 1) return;
	f = a.frags + a.cfrag[co + ci]; nf = a.cnfrag[co + ci];
	const bool fwd = a.corient[co + ci] != 0;
	const bool first_mate = a.pr.paired ? (((a.first_read + r) & 1) == 0) : true;
	const int64_t start = fwd ? f[0].gPos : a.ix.twoG - (f[0].gPos + f[0].gLen);
	const int sfield = first_mate ? (fwd ? 0 : 3) : (fwd ? 1 : 2);   // F1 / R1 / R2 / F2 inside [F1,R2,F2,R1]
	prof_range(a, a.prof.sdiff, 4, sfield, start, start + rlen);
	// seeds found by the FM index match the reference exactly, so they add the reference base over their span; seeds of a
	// rescued candidate may carry shifted labels (src/KmerAnalysis.cpp quirk) and lower-case bases are not counted by
	// the reference at all: both go base by base
	const bool exact = a.cands[co + ci].pbeg < a.n_locs && !(a.rflag[r] & 1);
	int natom = 2;
	for (int k = 0; k < nf; k++)
	{
		const mc_frag_out x = f[k];
		if (x.bSimple)
		{
			if (exact)
			{
				if (fwd) prof_range(a, a.prof.cdiff, 1, 0, x.gPos, x.gPos + x.rLen);
				else prof_range(a, a.prof.cdiff, 1, 0, a.ix.twoG - x.gPos - x.rLen, a.ix.twoG - x.gPos);
				natom += 2;
			}
			else
			{
				for (int j = 0; j < x.rLen; j++)
				{
					const int b = base_field(rs[x.rPos + j]);
					if (b >= 0) { if (fwd) prof_base(a, x.gPos + j, b); else prof_base(a, a.ix.twoG - 1 - x.gPos - j, 3 - b); }
				}
				natom += x.rLen;
			}
			continue;
		}
		const uint8_t* a1 = a.aln + x.aln_off; const uint8_t* a2 = a1 + x.aln_cap;
		// note: an emptied clip piece (rLen == gLen == 0) also lands here and registers an empty insertion string, as the reference does (:121)
		if (x.gLen == 0) { indel_emit(a, q, 0, (fwd ? x.gPos : a.ix.twoG - x.gPos) - 1, a1, x.aln_len); continue; }
		if (x.rLen == 0) { indel_emit(a, q, 1, (fwd ? x.gPos : a.ix.twoG - x.gPos - x.gLen) - 1, a2, x.aln_len); continue; }
		a.ptask[pt++] = a.cfrag[co + ci] + k;
	}
	mc_stat_add(&a.st->profile_columns, (uint32_t)(2 * rlen));
	mc_stat_add(&a.st->profile_atomics, (uint32_t)(natom));
}

// `nl` lanes (a warp) count the columns of one gapped / mismatching piece of an accepted read: pieces without gaps go lane
// by lane over consecutive columns, pieces with gaps are walked by one lane (UpdateProfile, src/AlignmentProfile.cpp:135-166)
MC_HD int profpiece_body(int64_t t, int lane, int nl, const PipeArgs& a, const ProfArgs& q)
{
	const mc_frag_out x = a.frags[a.ptask[t]];
	const bool fwd = x.gPos < a.ix.G;
	const uint8_t* a1 = a.aln + x.aln_off; const uint8_t* a2 = a1 + x.aln_cap;
	int64_t g = fwd ? x.gPos : a.ix.twoG - (x.gPos + x.gLen);
	int natom = 0;
	if (x.aln_len == x.rLen && x.aln_len == x.gLen)
	{
		for (int j = lane; j < x.aln_len; j += nl) { const int b = base_field(a1[j]); if (b >= 0) { prof_base(a, g + j, b); natom++; } }
	}
	else if (lane == 0)
	{
		for (int j = 0; j < x.aln_len;)
		{
			if (a2[j] == '-')
			{
				int e = 1; while (j + e < x.aln_len && a2[j + e] == '-') e++;
				indel_emit(a, q, 0, g - 1, a1 + j, e); j += e;
			}
			else if (a1[j] == '-')
			{
				int e = 1; while (j + e < x.aln_len && a1[j + e] == '-') e++;
				indel_emit(a, q, 1, g - 1, a2 + j, e); j += e; g += e;
			}
			else { int b = base_field(a1[j]); if (b >= 0) { prof_base(a, g, b); natom++; } j++; g++; }
		}
	}
	return natom;   // summed per warp by the caller
}

// ---- read-out: difference arrays -> MappingRecord_t ------------------------------------------------------------
// One warp per MC_PROF_BLOCK columns, a lane per column (coalesced 16-byte / 4-byte loads, 16-byte stores); the running
// sums of the six difference arrays are warp scans chained from one 32-column group to the next.
// Pass 1: totals of the six difference arrays per block (sums[6][n_blocks], scanned afterwards).
MC_HD void profsum_body(int64_t b, int lane, int nl, const DevProfile& p, int64_t G, int64_t n_blocks, int64_t* sums)
{
	const int64_t g0 = b * MC_PROF_BLOCK; int64_t g1 = g0 + MC_PROF_BLOCK; if (g1 > G) g1 = G;
	int t[6] = {0, 0, 0, 0, 0, 0};   // a block's totals are differences of coverages: they fit 32 bits
	for (int64_t g = g0 + lane; g < g1; g += nl)
	{
		const mc_u32x4 q = mc_ldg128(p.sdiff + g * 4);
		t[0] += (int)q.x; t[1] += (int)q.y; t[2] += (int)q.z; t[3] += (int)q.w;
		t[4] += p.cdiff[g]; t[5] += p.mdiff[g];
	}
	for (int k = 0; k < 6; k++) { const int v = mc_warp_sum(t[k]); if (lane == 0) sums[k * n_blocks + b] = v; }
}

// Pass 2: MappingRecord_t image (reference src/structure.h:152-163) of the columns of block b that fall into [beg, end);
// pre[6][n_blocks] holds the exclusive block prefixes
MC_HD void profpack_body(int64_t b, int lane, int nl, const DevIndex& ix, const DevProfile& p, int64_t n_blocks, const int64_t* pre, int64_t beg, int64_t end, uint64_t* out)
{
	const int64_t g0 = b * MC_PROF_BLOCK; int64_t g1 = g0 + MC_PROF_BLOCK; if (g1 > ix.G) g1 = ix.G;
	int64_t t[6]; for (int k = 0; k < 6; k++) t[k] = pre[k * n_blocks + b];
	for (int64_t base = g0; base < g1; base += nl)   // every lane stays in the loop: the scans need the whole warp
	{
		const int64_t g = base + lane; const bool in = g < g1;
		int d[6] = {0, 0, 0, 0, 0, 0};
		if (in) { const mc_u32x4 q = mc_ldg128(p.sdiff + g * 4); d[0] = (int)q.x; d[1] = (int)q.y; d[2] = (int)q.z; d[3] = (int)q.w; d[4] = p.cdiff[g]; d[5] = p.mdiff[g]; }
		int64_t v[6];
		for (int k = 0; k < 6; k++) { const int s = mc_warp_incl_scan(d[k], lane); v[k] = t[k] + s; t[k] += mc_warp_last(s); }
		if (!in || g < beg || g >= end) continue;
		const uint32_t w0 = p.base16[g * 2], w1 = p.base16[g * 2 + 1];
		uint64_t c[4] = {w0 & 0xFFFF, w0 >> 16, w1 & 0xFFFF, w1 >> 16};
		c[mc_ref_code(ix, g)] += (uint64_t)v[4];
		uint64_t M = (uint64_t)v[5], R = p.rcount[g];
		for (int k = 0; k < 4; k++) if (c[k] > 4095) c[k] = 4095;
		if (M > 4095) M = 4095;
		const int64_t i = g - beg;
		out[2 * i] = c[0] | c[1] << 12 | c[2] << 24 | c[3] << 36 | M << 48 | (R & 15) << 60;
		out[2 * i + 1] = ((uint64_t)v[0] & 0xFFFF) | ((uint64_t)v[1] & 0xFFFF) << 16 | ((uint64_t)v[2] & 0xFFFF) << 32 | ((uint64_t)v[3] & 0xFFFF) << 48;
	}
}

// CheckMappingCoverage / ReportDuplicationRate (reference src/ReadMapping.cpp:648-687) over packed records: columns with
// A+C+G+T > 0 and their sum, columns with readCount > 0 and their sum.  acc[4] = {aligned, coverage, sites, reads}
MC_HD void profstat_body(int64_t i, const uint64_t* recs, mc_u64* acc)
{
	const uint64_t w = recs[2 * i];
	const uint32_t cov = (uint32_t)((w & 4095) + ((w >> 12) & 4095) + ((w >> 24) & 4095) + ((w >> 36) & 4095)), rc = (uint32_t)(w >> 60);
	mc_stat_add(acc + 0, cov ? 1u : 0u); mc_stat_add(acc + 1, cov);
	mc_stat_add(acc + 2, rc ? 1u : 0u); mc_stat_add(acc + 3, rc);
}

// ---- aggregation of the raw indel records (mc_profile_indels): the device brings equal (kind, position, sequence) records
// next to each other - sort key = kind | position | hash of the sequence - and lays the records and their sequences out in that
// order, so that the host only counts runs (and orders the few distinct sequences of one position as std::string compares)
MC_HD void indkey_body(int64_t i, const mc_indel_rec* recs, const uint8_t* seq, uint64_t* keys, uint32_t* idx)
{
	const mc_indel_rec r = recs[i];
	uint32_t h = 2166136261u;                                        // FNV-1a over the sequence
	for (int k = 0; k < r.len; k++) h = (h ^ seq[r.seq_off + k]) * 16777619u;
	h = (h ^ (uint32_t)r.len) * 16777619u;
	keys[i] = ((uint64_t)(r.kind & 1) << 63) | ((uint64_t)(uint32_t)(r.pos + 1) << 31) | (uint64_t)(h >> 1);   // positions start at -1
	idx[i] = (uint32_t)i;
}
MC_HD void indlen_body(int64_t j, const mc_indel_rec* recs, const uint32_t* idx, uint32_t* len) { len[j] = (uint32_t)recs[idx[j]].len; }
MC_HD void indgather_body(int64_t j, const mc_indel_rec* recs, const uint8_t* seq, const uint32_t* idx, const int64_t* off, mc_indel_rec* out, uint8_t* out_seq)
{
	mc_indel_rec r = recs[idx[j]];
	const uint8_t* s = seq + r.seq_off;
	r.seq_off = (int32_t)off[j];
	for (int k = 0; k < r.len; k++) out_seq[off[j] + k] = s[k];
	out[j] = r;
}

// fingerprint of the packed profile: every column's record mixed with its position (splitmix64 finaliser), summed into acc[0]
// and xor-ed into acc[1] - independent of the order in which the columns are visited
MC_HD uint64_t prof_mix(uint64_t x)
{
	x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull; x ^= x >> 27; x *= 0x94d049bb133111ebull; x ^= x >> 31;
	return x;
}
MC_HD uint64_t profhash_of(int64_t g, const uint64_t* recs, int64_t i)
{
	return prof_mix(recs[2 * i] + prof_mix(recs[2 * i + 1] + prof_mix((uint64_t)g + 0x9e3779b97f4a7c15ull)));
}

#endif
