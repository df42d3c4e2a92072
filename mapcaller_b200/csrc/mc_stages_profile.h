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
	int64_t n_keys;
	int64_t* bp_pos; mc_u64* bp_bump; int64_t bp_cap;                  // BreakPointMap increments (persistent)
	mc_indel_rec* ind; mc_u64* ind_bump; int64_t ind_cap;              // InsertSeqMap / DeleteSeqMap increments (persistent)
	uint8_t* ind_seq; mc_u64* ind_seq_bump; int64_t ind_seq_cap;
};

// one thread per read: which update applies, clip / break-point bookkeeping, gate key
MC_HD void profkey_body(int64_t r, const PipeArgs& a, const ProfArgs& q)
{
	q.accept[r] = 0;
	const ReadSum sum = a.rsum[r];
	if (sum.score == 0) return;
	if (sum.n_live != 1) { q.accept[r] = 2; return; }
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
		if (first.rPos > a.pr.max_clip) return;
	}
	if (last.rLen == 0 && last.gLen == 0)
	{
		if (rlen - last.rPos > 20)
		{
			int64_t k = (int64_t)mc_atomic_add(q.bp_bump, (mc_u64)1);
			if (k < q.bp_cap) q.bp_pos[k] = last.gPos < a.ix.G ? last.gPos : a.ix.twoG - 1 - last.gPos; else mc_atomic_or(&a.st->overflow, (mc_u64)1 << 48);
		}
		if (rlen - last.rPos > a.pr.max_clip) return;
	}
	const int64_t start = a.corient[co + ci] ? first.gPos : a.ix.twoG - (first.gPos + first.gLen);
	if (start < 0 || start >= a.ix.G) return; // the reference would index outside MappingRecordArr
	const int64_t k = (int64_t)mc_atomic_add(q.key_bump, (mc_u64)1);
	q.keys[k] = ((uint64_t)start << MC_KEY_SHIFT) | (uint64_t)r;
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

MC_HD void prof_add16(const PipeArgs& a, int64_t g, int field) // field: 0 A,1 C,2 G,3 T,4 F1,5 R2,6 F2,7 R1
{
	if (g < 0 || g >= a.ix.G) return;
	mc_atomic_add(a.prof.cnt16 + g * 4 + (field >> 1), (uint32_t)(1u << ((field & 1) << 4)));
}
MC_HD int base_field(uint8_t c) { return c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : c == 'T' ? 3 : -1; }

MC_HD void indel_emit(const PipeArgs& a, const ProfArgs& q, int kind, int64_t pos, const uint8_t* s, int len)
{
	const int64_t k = (int64_t)mc_atomic_add(q.ind_bump, (mc_u64)1);
	const int64_t o = (int64_t)mc_atomic_add(q.ind_seq_bump, (mc_u64)len);
	if (k >= q.ind_cap || o + len > q.ind_seq_cap) { mc_atomic_or(&a.st->overflow, (mc_u64)1 << 48); return; }
	mc_indel_rec rec; rec.pos = pos; rec.kind = kind; rec.len = len; rec.count = 1; rec.seq_off = (int32_t)o;
	q.ind[k] = rec;
	for (int i = 0; i < len; i++) q.ind_seq[o + i] = s[i];
}

// `nl` lanes cooperate on one read (a warp on the GPU); lane-strided loops give coalesced atomics over
// consecutive profile columns
MC_HD void scatter_body(int64_t r, int lane, int nl, const PipeArgs& a, const ProfArgs& q)
{
	const int mode = q.accept[r];
	if (mode == 0) return;
	const int rlen = (int)(a.roff[r + 1] - a.roff[r]);
	const uint8_t* rs = a.seq + a.roff[r];
	const int64_t co = pa_cand_off(a, r);
	const int nc = a.ncand[r];
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
			for (int64_t g = g0 + lane; g < g1; g += nl) if (g >= 0 && g < a.ix.G) mc_atomic_add(a.prof.multi + g, 1u);
			if (lane == 0 && g1 > g0) mc_atomic_add(&a.st->profile_columns, (mc_u64)(g1 - g0));
		}
		return;
	}
	int ci = 0; while (ci < nc && a.cscore[co + ci] == 0) ci++;
	const mc_frag_out* f = a.frags + a.cfrag[co + ci];
	const int nf = a.cnfrag[co + ci];
	const bool fwd = a.corient[co + ci] != 0;
	const bool first_mate = a.pr.paired ? (((a.first_read + r) & 1) == 0) : true;
	const int64_t start = fwd ? f[0].gPos : a.ix.twoG - (f[0].gPos + f[0].gLen);
	const int sfield = first_mate ? (fwd ? 4 : 7) : (fwd ? 5 : 6);   // F1 / R1 / R2 / F2
	for (int i = lane; i < rlen; i += nl) prof_add16(a, start + i, sfield);
	if (lane == 0) mc_atomic_add(&a.st->profile_columns, (mc_u64)(2 * rlen));
	for (int k = 0; k < nf; k++)
	{
		const mc_frag_out x = f[k];
		if (x.bSimple)
		{
			if (fwd) { for (int j = lane; j < x.rLen; j += nl) { int b = base_field(rs[x.rPos + j]); if (b >= 0) prof_add16(a, x.gPos + j, b); } }
			else { for (int j = lane; j < x.rLen; j += nl) { int b = base_field(rs[x.rPos + j]); if (b >= 0) prof_add16(a, a.ix.twoG - 1 - x.gPos - j, 3 - b); } }
			continue;
		}
		if (lane != 0) continue; // gapped / indel pieces are short: one lane walks the columns
		const uint8_t* a1 = a.aln + x.aln_off; const uint8_t* a2 = a1 + x.aln_cap;
		// note: an emptied clip piece (rLen == gLen == 0) also lands here and registers an empty insertion string, as the reference does (:121)
		if (x.gLen == 0) { indel_emit(a, q, 0, (fwd ? x.gPos : a.ix.twoG - x.gPos) - 1, a1, x.aln_len); continue; }
		if (x.rLen == 0) { indel_emit(a, q, 1, (fwd ? x.gPos : a.ix.twoG - x.gPos - x.gLen) - 1, a2, x.aln_len); continue; }
		int64_t g = fwd ? x.gPos : a.ix.twoG - (x.gPos + x.gLen);
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
			else { int b = base_field(a1[j]); if (b >= 0) prof_add16(a, g, b); j++; g++; }
		}
	}
}

// MappingRecord_t image of one column (reference src/structure.h:152-163)
MC_HD void profpack_body(int64_t i, const DevProfile& p, int64_t beg, uint64_t* out)
{
	const int64_t g = beg + i;
	const uint32_t w0 = p.cnt16[g * 4], w1 = p.cnt16[g * 4 + 1], w2 = p.cnt16[g * 4 + 2], w3 = p.cnt16[g * 4 + 3];
	uint64_t A = w0 & 0xFFFF, C = w0 >> 16, G_ = w1 & 0xFFFF, T = w1 >> 16, M = p.multi[g], R = p.rcount[g];
	if (A > 4095) A = 4095; if (C > 4095) C = 4095; if (G_ > 4095) G_ = 4095; if (T > 4095) T = 4095; if (M > 4095) M = 4095;
	out[2 * i] = A | C << 12 | G_ << 24 | T << 36 | M << 48 | (R & 15) << 60;
	out[2 * i + 1] = (uint64_t)w2 | (uint64_t)w3 << 32;
}

#endif
