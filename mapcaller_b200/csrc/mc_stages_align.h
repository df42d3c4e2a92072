// Candidate -> base-level alignment (included by mc_stages.h).
#ifndef MC_STAGES_ALIGN_H
#define MC_STAGES_ALIGN_H

MC_HD bool frag_less_readpos(const mc_frag_out& x, const mc_frag_out& y)
{
	return x.rPos == y.rPos ? x.gPos < y.gPos : x.rPos < y.rPos; // CompByReadPos, reference src/ReadAlignment.cpp:23
}

// workspace of one fill: traceback bytes (row-major for the per-thread form, anti-diagonal-major strips of 32 columns
// for the warp form, whichever is larger and for either algorithm) followed by two int rows
MC_HOST_HD int64_t dp_tb_bytes(int m, int n)
{
	const int64_t full = (int64_t)(m + 1) * (n + 1);
	const int64_t w_nw = (int64_t)((n + 31) >> 5) * (m + 32) * 32, w_ksw = (int64_t)((m + 31) >> 5) * (n + 32) * 32;
	int64_t t = full > w_nw ? full : w_nw; if (w_ksw > t) t = w_ksw;
	return (t + 7) & ~7ll;
}
MC_HOST_HD int64_t dp_ws_bytes(int m, int n) { return dp_tb_bytes(m, n) + 8 * (int64_t)((m > n ? m : n) + 2); }

// ------------------------------------------------------------------------------------------------
// alnprep: one thread per read.  For every live candidate: order the seeds along the read, resolve
// overlaps, insert the "normal" (non-seed) pieces between seeds and at both read ends, check that the
// alignment stays inside one chromosome, lay out the piece strings and queue the pieces that need a
// gapped fill.  (reference src/ReadAlignment.cpp:306-342 with RemoveOverlaps :38, RemoveNullFragPairs :29,
// IdentifyNormalPairs :67, CheckAlignmentValidity src/tools.cpp:119, ProcessNormalPair :155)
// ------------------------------------------------------------------------------------------------
// The three arenas (fragments, alignment strings, piece queue) are claimed with block-wide cursor bumps: every thread of the
// block takes part in both (`live` = false for the threads past the end of the batch).
MC_HD void alnprep_body(int64_t r, bool live, const PipeArgs& a)
{
	live = live && a.read_redo[r];
	const int rlen = live ? (int)(a.roff[r + 1] - a.roff[r]) : 0;
	const int64_t co = live ? pa_cand_off(a, r) : 0;
	const int nc = live ? a.ncand[r] : 0;
	int cap_total = 0;
	for (int ci = 0; ci < nc; ci++)
	{
		a.cfrag[co + ci] = 0; a.cnfrag[co + ci] = 0; a.corient[co + ci] = -1;
		if (a.cscore[co + ci] == 0) continue;
		const Cand c = a.cands[co + ci];
		cap_total += 2 * (c.pend - c.pbeg) + 1;
	}
	int64_t fb = mc_block_bump(a.frag_bump, (uint32_t)cap_total);
	bool dead = false;
	if (cap_total && fb + cap_total > a.frag_cap) { mc_atomic_or(&a.st->overflow, (mc_u64)1 << 8); dead = true; }
	int need_total = 0, np_total = 0;
	for (int ci = 0; ci < nc && !dead; ci++)
	{
		if (a.cscore[co + ci] == 0) continue;
		const Cand c = a.cands[co + ci];
		const int ns = c.pend - c.pbeg;
		const int cap = 2 * ns + 1;
		mc_frag_out* f = a.frags + fb;
		const int64_t fbase = fb;
		fb += cap;
		// seeds sorted by (rPos, gPos), written into the upper half so the final list can be built in place below
		mc_frag_out* s = f + (cap - ns);
		for (int i = 0; i < ns; i++)
		{
			const SPair sp = a.pairs[c.pbeg + i];
			mc_frag_out x; x.gPos = sp.gpos; x.rPos = sp.rpos; x.rLen = sp.len; x.gLen = sp.len; x.bSimple = 1;
			x.aln_off = 0; x.aln_len = 0; x.aln_cap = 0; x.pad = 0;
			int j = i - 1;
			while (j >= 0 && frag_less_readpos(x, s[j])) { s[j + 1] = s[j]; j--; }
			s[j + 1] = x;
		}
		// RemoveOverlaps + RemoveNullFragPairs
		bool ov = false;
		for (int i = 0, j = 1; j < ns; i++, j++)
		{
			if (s[i].rPos == s[j].rPos) { ov = true; s[i].rLen = s[i].gLen = 0; }
			else if (s[i].gPos >= s[j].gPos || s[i].gPos + s[i].gLen > s[j].gPos)
			{
				ov = true;
				const int o = (int)(s[i].gPos + s[i].gLen - s[j].gPos);
				if ((s[i].rLen -= o) < 0) s[i].rLen = 0;
				if ((s[i].gLen -= o) < 0) s[i].gLen = 0;
			}
		}
		int m = ns;
		if (ov) { m = 0; for (int i = 0; i < ns; i++) if (s[i].rLen != 0) { if (m != i) s[m] = s[i]; m++; } }
		if (m == 0) { a.cscore[co + ci] = 0; continue; } // cannot happen for seeds of one read; the reference would index an empty vector
		// IdentifyNormalPairs: head piece, then seed / gap piece interleaved, then tail piece
		int nf = 0;
		if (s[0].rPos > 0)
		{
			mc_frag_out x = s[0]; x.bSimple = 0; x.rLen = x.gLen = s[0].rPos; x.gPos = s[0].gPos - s[0].rPos; x.rPos = 0;
			f[nf++] = x;
		}
		for (int i = 0; i < m; i++)
		{
			const mc_frag_out cur = s[i];
			f[nf++] = cur; // nf <= index of s[i] inside f, so nothing unread is overwritten
			if (i + 1 < m)
			{
				int rg = s[i + 1].rPos - (cur.rPos + cur.rLen); if (rg < 0) rg = 0;
				int64_t gg64 = s[i + 1].gPos - (cur.gPos + cur.gLen); int gg = gg64 < 0 ? 0 : (int)gg64;
				if (rg > 0 || gg > 0)
				{
					mc_frag_out x = cur; x.bSimple = 0; x.rPos = cur.rPos + cur.rLen; x.gPos = cur.gPos + cur.gLen; x.rLen = rg; x.gLen = gg;
					if (x.rPos > s[i + 1].rPos) mc_atomic_add(&a.st->odd_merge, (mc_u64)1);
					f[nf++] = x;
				}
			}
		}
		{
			const mc_frag_out last = f[nf - 1];
			if (last.rPos + last.rLen < rlen)
			{
				mc_frag_out x = last; x.bSimple = 0; x.rPos = last.rPos + last.rLen; x.gPos = last.gPos + last.gLen; x.rLen = x.gLen = rlen - x.rPos;
				f[nf++] = x;
			}
		}
		// CheckAlignmentValidity
		{
			const int64_t g0 = f[0].gPos, g1 = f[nf - 1].gPos + f[nf - 1].gLen;
			bool ok = !(g0 < 0 || g1 > a.ix.twoG);
			if (ok)
			{
				int i1 = mc_chrom_lower_bound(a.ix, g0), i2 = mc_chrom_lower_bound(a.ix, g1 - 1);
				ok = i1 < a.ix.n_end && i2 < a.ix.n_end && a.ix.chrom_end[i1] == a.ix.chrom_end[i2];
			}
			if (!ok) { a.cscore[co + ci] = 0; continue; }
		}
		// space for the strings of the normal pieces; piece_body lays them out
		for (int i = 0; i < nf; i++) if (!f[i].bSimple) { need_total += 2 * (f[i].rLen + f[i].gLen); np_total++; }
		a.cfrag[co + ci] = (int32_t)fbase; a.cnfrag[co + ci] = nf;
	}
	if (dead) for (int ci = 0; ci < nc; ci++) a.cscore[co + ci] = 0;
	int64_t ab, pt;                                       // pieces never outnumber fragments: the piece list has frag_cap entries
	mc_block_bump2(a.aln_bump, (uint32_t)need_total, a.ptask_bump, (uint32_t)np_total, &ab, &pt);
	if (!np_total) return;
	if (ab + need_total > a.aln_cap)
	{
		mc_atomic_or(&a.st->overflow, (mc_u64)1 << 16);
		for (int ci = 0; ci < nc; ci++) { a.cscore[co + ci] = 0; a.cfrag[co + ci] = 0; a.cnfrag[co + ci] = 0; }
		return;
	}
	for (int ci = 0; ci < nc; ci++)
	{
		const int nf = a.cnfrag[co + ci];
		if (!nf || a.cscore[co + ci] == 0) continue;
		const int64_t fbase = a.cfrag[co + ci];
		mc_frag_out* f = a.frags + fbase;
		for (int i = 0; i < nf; i++)
		{
			mc_frag_out& x = f[i];
			if (x.bSimple) continue;
			const int cp = x.rLen + x.gLen;
			x.aln_off = (int32_t)ab; x.aln_cap = cp; ab += 2 * cp;
			x.aln_len = x.rLen > x.gLen ? x.rLen : x.gLen;
			x.pad = (int32_t)r;                     // internal: the read this piece belongs to
			if (pt < a.frag_cap) a.ptask[pt] = (int32_t)(fbase + i); else mc_atomic_or(&a.st->overflow, (mc_u64)1 << 8);
			pt++;
		}
	}
}

// ------------------------------------------------------------------------------------------------
// piece: `nl` lanes (an 8-lane tile of a warp: the median piece is 13 bases) lay out the two strings of one normal piece - read bases against RefSequence, both
// reverse-complemented on the reverse strand - count the mismatches and queue the piece for a gapped fill when the
// reference would run one (ProcessNormalPair, reference src/ReadAlignment.cpp:155-191)
// ------------------------------------------------------------------------------------------------
MC_HD void piece_body(int64_t t, int lane, int nl, const PipeArgs& a)
{
	if (a.st->overflow) return;                      // an arena ran out in alnprep: queue slots may be unwritten, the attempt is repeated
	const int32_t fi = a.ptask[a.ptask_begin + t];   // the caller bounds t by the list's cursor
	const mc_frag_out x = a.frags[fi];
	const uint8_t* rs = a.seq + a.roff[x.pad];
	uint8_t* a1 = a.aln + x.aln_off; uint8_t* a2 = a1 + x.aln_cap;
	const bool rev = x.gPos >= a.ix.G;
	if (x.rLen > 0)
	{
		if (!rev) for (int k = lane; k < x.rLen; k += nl) a1[k] = rs[x.rPos + k];
		else for (int k = lane; k < x.rLen; k += nl) a1[k] = mc_complement(rs[x.rPos + x.rLen - 1 - k]);
	}
	else for (int k = lane; k < x.gLen; k += nl) a1[k] = '-';
	if (x.gLen > 0)
	{
		if (!rev) for (int k = lane; k < x.gLen; k += nl) a2[k] = mc_ref_char(a.ix, x.gPos + k);
		else for (int k = lane; k < x.gLen; k += nl) a2[k] = (uint8_t)("TGCA"[mc_ref_code(a.ix, x.gPos + x.gLen - 1 - k)]);
	}
	else for (int k = lane; k < x.rLen; k += nl) a2[k] = '-';
	if (x.rLen <= 0 || x.gLen <= 0) return;
	bool dp = x.rLen != x.gLen;
	if (!dp)
	{
		MC_TILE8_SYNC();
		int mis = 0;
		for (int k = lane; k < x.rLen; k += nl) if (a1[k] != a2[k]) mis++;
		mis = mc_tile8_sum(mis);
		dp = mis > 1 && mis >= (int)(x.rLen * 0.2);
	}
	if (dp && lane == 0)
	{
		const int64_t tt = (int64_t)mc_atomic_add(a.task_bump, (mc_u64)1);
		const int64_t wsn = dp_ws_bytes(x.rLen, x.gLen);
		const int64_t ws = (int64_t)mc_atomic_add(a.dpws_bump, (mc_u64)wsn);
		if (tt >= a.task_cap) { mc_atomic_or(&a.st->overflow, (mc_u64)1 << 24); return; }
		if (ws + wsn > a.dpws_cap) { mc_atomic_or(&a.st->overflow, (mc_u64)1 << 32); return; }
		DpTask tk; tk.frag = fi; tk.m = x.rLen; tk.n = x.gLen; tk.pad = 0; tk.ws_off = ws;
		a.tasks[tt] = tk;
	}
}

// ------------------------------------------------------------------------------------------------
// gapped fill of one piece: full-matrix global affine alignment with traceback.
//   nw   : reference src/nw_alignment.cpp:18-83 restated in exact x2 integers (match +2, mismatch -2,
//          gap of length k = -2-k), value-equality traceback with the horizontal gap tested first.
//   ksw2 : reference src/ksw2_alignment.cpp:70-272 (match +1, mismatch -1, N = 0, gap of length k =
//          -(2+k)), direction bits and backtrack state machine of ksw_extz2_sse / ksw_backtrack.
// One byte of traceback per cell in the workspace; the gapped strings overwrite the raw ones.
// ------------------------------------------------------------------------------------------------
#define MC_NEG_INF (-0x20000000)

// s1 / s2: where the gapped strings go (the piece's slots in the alignment arena); in1 / in2: where the raw strings are
// read from - the same place (rewritten in place, back to front) or a private copy
MC_HD int dp_core(bool ksw2, int m, int n, uint8_t* s1, uint8_t* s2, const uint8_t* in1, const uint8_t* in2, uint8_t* tb, int* row0, int* row1)
{
	int len = 0;
	if (!ksw2)
	{
		// s1 down the rows (i), s2 along the columns (j); row0 = s[i-1][*], row1 = t[i-1][*]
		const int W = n + 1;
		row0[0] = 0; row1[0] = 0; tb[0] = 0;
		for (int j = 1; j <= n; j++) { row0[j] = -2 - j; row1[j] = -131072; tb[j] = 1; }   // s[0][j] == r[0][j]
		for (int i = 1; i <= m; i++)
		{
			const int c1 = mc_nt4(in1[i - 1]);
			int diag = row0[0];               // s[i-1][0]
			int sl = -2 - i;                  // s[i][0] == t[i][0]
			int rl = -131072;                 // r[i][0]
			row0[0] = sl; row1[0] = sl;
			uint8_t* row = tb + (int64_t)i * W;
			row[0] = 2;
			for (int j = 1; j <= n; j++)
			{
				const int rr = (rl - 1 > sl - 3) ? rl - 1 : sl - 3;
				const int up_t = row1[j], up_s = row0[j];
				const int tt = (up_t - 1 > up_s - 3) ? up_t - 1 : up_s - 3;
				const int dd = diag + (c1 == mc_nt4(in2[j - 1]) ? 2 : -2);
				const int ss = mc_max3(dd, rr, tt);
				row[j] = (uint8_t)((ss == rr ? 1 : 0) | (ss == tt ? 2 : 0));
				diag = up_s; row0[j] = ss; row1[j] = tt; sl = ss; rl = rr;
			}
		}
		int i = m, j = n;
		while (i > 0 || j > 0) { const uint8_t d = tb[(int64_t)i * W + j]; if (d & 1) j--; else if (d & 2) i--; else { i--; j--; } len++; }
		i = m; j = n; int k = len;
		while (i > 0 || j > 0)   // back to front: the write cursor never passes an unread source byte
		{
			const uint8_t d = tb[(int64_t)i * W + j];
			k--;
			if (d & 1) { s2[k] = in2[j - 1]; s1[k] = '-'; j--; }
			else if (d & 2) { s1[k] = in1[i - 1]; s2[k] = '-'; i--; }
			else { s1[k] = in1[i - 1]; s2[k] = in2[j - 1]; i--; j--; }
		}
	}
	else
	{
		// i over the genome piece (target), j over the read piece (query); row0 = H(i-1, *), row1 = E(i, *) carried down
		const int Wq = m;
		for (int j = 0; j < m; j++) { row0[j] = -(2 + (j + 1)); row1[j] = MC_NEG_INF; }
		for (int i = 0; i < n; i++)
		{
			const int ct = mc_nt4(in2[i]);
			int hdiag = i == 0 ? 0 : -(2 + i);       // H(i-1, -1)
			int hleft = -(2 + (i + 1));              // H(i, -1)
			int f = MC_NEG_INF;
			uint8_t* row = tb + (int64_t)i * Wq;
			for (int j = 0; j < m; j++)
			{
				const int cq = mc_nt4(in1[j]);
				const int sc = (ct == 4 || cq == 4) ? 0 : (ct == cq ? 1 : -1);
				const int hup = row0[j];
				int e = MC_NEG_INF;
				if (i > 0) { const int ho = hup - 2; const int eo = row1[j]; e = (ho > eo ? ho : eo) - 1; }
				int ff = MC_NEG_INF;
				if (j > 0) { const int ho = hleft - 2; ff = (ho > f ? ho : f) - 1; }
				int z = hdiag + sc; uint8_t d = 0;
				if (e > z) { d = 1; z = e; }
				if (ff > z) { d = 2; z = ff; }
				if (e > z - 2) d |= 0x08;
				if (ff > z - 2) d |= 0x10;
				row[j] = d;
				hdiag = hup; row0[j] = z; row1[j] = e; hleft = z; f = ff;
			}
		}
		// ksw_backtrack (reference src/ksw2_alignment.cpp:25-68); 'D' consumes the genome piece, 'I' the read piece
		for (int pass = 0; pass < 2; pass++)
		{
			int i = n - 1, j = m - 1, state = 0, k = len, cnt = 0;
			while (i >= 0 && j >= 0)
			{
				const uint32_t tmp = tb[(int64_t)i * Wq + j];
				if (state == 0) state = tmp & 7;
				else if (!((tmp >> (state + 2)) & 1)) state = 0;
				if (state == 0) state = tmp & 7;
				if (state == 0) { if (pass) { k--; s1[k] = in1[j]; s2[k] = in2[i]; } i--; j--; }
				else if (state == 1 || state == 3) { if (pass) { k--; s2[k] = in2[i]; s1[k] = '-'; } i--; }
				else { if (pass) { k--; s1[k] = in1[j]; s2[k] = '-'; } j--; }
				cnt++;
			}
			while (i >= 0) { if (pass) { k--; s2[k] = in2[i]; s1[k] = '-'; } i--; cnt++; }
			while (j >= 0) { if (pass) { k--; s1[k] = in1[j]; s2[k] = '-'; } j--; cnt++; }
			if (!pass) len = cnt;
		}
	}
	return len;
}

MC_HD void dp_body(int64_t t, const PipeArgs& a)
{
	if (a.st->overflow) return;            // an arena ran out earlier in this attempt: a task slot may be unwritten, the attempt is repeated
	{
		int64_t end = (int64_t)*a.task_bump; if (end > a.task_cap) end = a.task_cap;
		if (a.task_begin + t >= end) return;
	}
	const DpTask tk = a.tasks[a.task_begin + t];
	mc_frag_out& x = a.frags[tk.frag];
	const int m = tk.m, n = tk.n;
	uint8_t* s1 = a.aln + x.aln_off; uint8_t* s2 = s1 + x.aln_cap;   // read piece (m), genome piece (n)
	uint8_t* tb = a.dpws + tk.ws_off;
	int* row0 = (int*)(tb + dp_tb_bytes(m, n));
	x.aln_len = dp_core(a.pr.alg_ksw2 != 0, m, n, s1, s2, s1, s2, tb, row0, row0 + ((m > n ? m : n) + 2));
	mc_atomic_add(&a.st->dp_cells, (mc_u64)((int64_t)m * n));
	mc_atomic_add(&a.st->dp_tasks, (mc_u64)1);
}

// Small fills (the bulk: the median piece is 13 x 13) by ONE thread each, everything it touches in a private slice of shared
// memory: the 31 idle steps that fill and drain the 32-lane wavefront of the warp form cost more than such a matrix itself.
#define MC_DP_SMALL_CELLS 289          // (m + 1) * (n + 1) <= 17 * 17
#define MC_DP_SMALL_MAXLEN 32          // and m, n <= 32 (a 1 x 143 strip is not "small")
#define MC_DP_SMALL_STRIDE 636         // bytes per thread: 292 traceback + 2 x 32 raw strings + 2 x 34 ints = 628, rounded up to an odd word count
MC_HOST_HD bool dp_is_small(int m, int n) { return (m + 1) * (n + 1) <= MC_DP_SMALL_CELLS && m <= MC_DP_SMALL_MAXLEN && n <= MC_DP_SMALL_MAXLEN; }
MC_HD void dp_small_body(int64_t t, const PipeArgs& a, uint8_t* ws, uint32_t* cells, uint32_t* tasks)
{
	const DpTask tk = a.tasks[t];
	const int m = tk.m, n = tk.n;
	if (!dp_is_small(m, n)) return;
	mc_frag_out& x = a.frags[tk.frag];
	uint8_t* s1 = a.aln + x.aln_off; uint8_t* s2 = s1 + x.aln_cap;
	uint8_t* tb = ws; uint8_t* in1 = ws + 292; uint8_t* in2 = in1 + 32; int* row0 = (int*)(in2 + 32); int* row1 = row0 + 34;
	for (int i = 0; i < m; i++) in1[i] = s1[i];
	for (int j = 0; j < n; j++) in2[j] = s2[j];
	x.aln_len = dp_core(a.pr.alg_ksw2 != 0, m, n, s1, s2, in1, in2, tb, row0, row1);
	*cells += (uint32_t)(m * n); *tasks += 1;
}



// ------------------------------------------------------------------------------------------------
// alnfin: one thread per read, candidates in order (best / sub-score bookkeeping is sequential)
// ------------------------------------------------------------------------------------------------
// RemoveHeadingGaps / RemoveTailingGaps (reference src/ReadAlignment.cpp:264-304)
MC_HD void trim_gaps(const PipeArgs& a, mc_frag_out& x, bool heading, bool first)
{
	const uint8_t* a1 = a.aln + x.aln_off; const uint8_t* a2 = a1 + x.aln_cap;
	int rs = 0, gs = 0, cut = 0;
	if (heading) { for (int j = 0; j < x.aln_len; j++) { if (a1[j] == '-') gs++; else if (a2[j] == '-') rs++; else break; cut++; } }
	else { for (int j = x.aln_len - 1; j >= 0; j--) { if (a1[j] == '-') gs++; else if (a2[j] == '-') rs++; else break; cut++; } }
	if (cut == 0) return;
	if (heading) x.aln_off += cut;     // both strings advance: aln2 still sits at aln_off + aln_cap
	x.aln_len -= cut;
	x.rLen -= rs; x.gLen -= gs;
	if (first) { x.rPos += rs; x.gPos += gs; }
}

// CheckLocalAlignmentQuality (reference src/ReadAlignment.cpp:193-232)
MC_HD bool local_quality_ok(const PipeArgs& a, const mc_frag_out& x)
{
	const uint8_t* a1 = a.aln + x.aln_off; const uint8_t* a2 = a1 + x.aln_cap;
	int type = -1, n = 0, mis = 0, status = 0;
	for (int i = 0; i < x.aln_len; i++)
	{
		if (a1[i] == '-') { if (type != 0) { type = 0; status++; } }
		else if (a2[i] == '-') { if (type != 1) { type = 1; status++; } }
		else { n++; if (a1[i] != a2[i]) mis++; if (type != 2) { type = 2; status++; } }
	}
	return !(status >= 4 || (mis >= 3 && mis >= (int)(n * 0.3)));
}

MC_HD void alnfin_body(int64_t r, const PipeArgs& a)
{
	if (a.st->overflow) return;          // an arena ran out earlier in this attempt (fills may be missing): the attempt is repeated
	if (!a.read_redo[r]) return;
	const int rlen = (int)(a.roff[r + 1] - a.roff[r]);
	const int64_t co = pa_cand_off(a, r);
	const int nc = a.ncand[r];
	// (int)(read.rlen*MaxMisMatchRate) and (int)(read.rlen*(1 - MaxMisMatchRate)) are float products (src/ReadAlignment.cpp:312,396)
#if MC_DEV_ONLY
	const int max_mm = (int)__fmul_rn((float)rlen, a.pr.max_mismatch_rate);
	const int min_score = (int)__fmul_rn((float)rlen, __fsub_rn(1.0f, a.pr.max_mismatch_rate));
#else
	volatile float one_minus = 1.0f - a.pr.max_mismatch_rate;
	volatile float prod1 = (float)rlen * a.pr.max_mismatch_rate, prod2 = (float)rlen * one_minus;
	const int max_mm = (int)prod1, min_score = (int)prod2;
#endif
	ReadSum sum; sum.score = 0; sum.sub_score = 0; sum.best_idx = -1; sum.n_live = 0;
	for (int ci = 0; ci < nc; ci++)
	{
		if (a.cscore[co + ci] == 0) continue;
		mc_frag_out* f = a.frags + a.cfrag[co + ci];
		const int nf = a.cnfrag[co + ci], tail = nf - 1;
		bool head_ok = true, tail_ok = true, dead = false;
		for (int i = 0; i < nf && !dead; i++)
		{
			mc_frag_out& x = f[i];
			if (x.bSimple) continue;
			if (i == 0)
			{
				trim_gaps(a, x, x.gPos < a.ix.G, true);
				if (x.aln_len >= 5 && !local_quality_ok(a, x))
				{
					head_ok = false; x.rLen = x.gLen = 0; x.aln_len = 0; x.rPos = f[1].rPos; x.gPos = f[1].gPos;
				}
			}
			else if (i == tail)
			{
				trim_gaps(a, x, !(x.gPos < a.ix.G), false);
				if (x.aln_len >= 5 && !local_quality_ok(a, x))
				{
					tail_ok = false; x.rLen = x.gLen = 0; x.aln_len = 0; x.rPos = f[i - 1].rPos + f[i - 1].rLen; x.gPos = f[i - 1].gPos + f[i - 1].gLen;
				}
			}
			else if (x.rLen >= 5 && x.gLen >= 5 && !local_quality_ok(a, x)) dead = true;
		}
		int score = 0;
		if (!dead && (head_ok || tail_ok))
		{
			int mism = 0;
			for (int i = 0; i < nf; i++)
			{
				const mc_frag_out& x = f[i];
				if (x.bSimple) score += x.rLen;
				else
				{
					const uint8_t* a1 = a.aln + x.aln_off; const uint8_t* a2 = a1 + x.aln_cap;
					for (int k = 0; k < x.aln_len; k++)
					{
						if (a1[k] == a2[k]) score++;
						else if (a1[k] != '-' && a2[k] != '-') mism++;
					}
				}
			}
			if (score != 0 && score < min_score && mism > max_mm) score = 0;
		}
		a.cscore[co + ci] = score;
		if (score == 0) continue;
		const bool fwd = f[0].gPos < a.ix.G;
		a.corient[co + ci] = fwd ? 1 : 0;
		if (!fwd) for (int i = 0, j = nf - 1; i < j; i++, j--) { mc_frag_out tmp = f[i]; f[i] = f[j]; f[j] = tmp; }
		if (score > sum.score) { sum.score = score; sum.best_idx = ci; }
		else if (score > sum.sub_score) sum.sub_score = score;
	}
	for (int ci = 0; ci < nc; ci++)
	{
		if (a.cscore[co + ci] < sum.score) a.cscore[co + ci] = 0;
		if (a.cscore[co + ci] > 0) sum.n_live++;
	}
	a.rsum[r] = sum;
}

// ------------------------------------------------------------------------------------------------
// pair statistics: GenCoordinatePair (reference src/ReadMapping.cpp:343-394) per pair; the proper-pair
// sums of one 200-read chunk (reference :529-531,538) are accumulated by one thread per chunk so the
// order of additions is fixed.
// ------------------------------------------------------------------------------------------------
MC_HD int64_t cand_first_gpos(const PipeArgs& a, int64_t co, int ci) { return a.frags[a.cfrag[co + ci]].gPos; }

MC_HD void pairstat_body(int64_t p, const PipeArgs& a)
{
	if (a.st->overflow) return;
	const int64_t r0 = 2 * p, r1 = r0 + 1;
	if (!a.read_redo[r0]) return;
	const int64_t c0 = pa_cand_off(a, r0), c1 = pa_cand_off(a, r1);
	const int n0 = a.ncand[r0], n1 = a.ncand[r1];
	mc_pair_out o; o.gPos1 = 0; o.gPos2 = 0; o.dist = 0;
	for (int i = 0; i < n0; i++)
	{
		const int j = a.cpaired[c0 + i];
		if (a.cscore[c0 + i] > 0 && j != -1 && a.cscore[c1 + j] > 0)
		{
			o.gPos1 = cand_first_gpos(a, c0, i); o.gPos2 = cand_first_gpos(a, c1, j);
			o.dist = o.gPos2 - o.gPos1; if (o.dist < 0) o.dist = -o.dist;
			break;
		}
	}
	if (o.dist == 0)
	{
		int k0 = 0, k1 = 0, f0 = -1, f1 = -1;
		for (int i = 0; i < n0; i++) if (a.cscore[c0 + i] > 0) { if (f0 < 0) f0 = i; k0++; }
		for (int j = 0; j < n1; j++) if (a.cscore[c1 + j] > 0) { if (f1 < 0) f1 = j; k1++; }
		if (k0 == 1 && k1 == 1)
		{
			o.gPos1 = cand_first_gpos(a, c0, f0); o.gPos2 = cand_first_gpos(a, c1, f1);
			o.dist = o.gPos2 - o.gPos1; if (o.dist < 0) o.dist = -o.dist;
		}
		else if (k0 == 0 && k1 >= 1) { o.gPos1 = -1; o.dist = o.gPos2 = cand_first_gpos(a, c1, f1); }
		else if (k0 >= 1 && k1 == 0) { o.dist = o.gPos1 = cand_first_gpos(a, c0, f0); o.gPos2 = -1; }
	}
	a.pair_out[p] = o;
}

// pairs the host has to look at in file order (inversion / translocation candidates, src/ReadMapping.cpp:486-522):
// everything that is neither unmapped / one-end-anchored nor a proper pair.  Run once, after the last attempt.
struct DiscRec { int64_t pair; mc_pair_out v; };
MC_HD void disclist_body(int64_t p, const PipeArgs& a, DiscRec* out, mc_u64* bump, int64_t cap)
{
	const mc_pair_out q = a.pair_out[p];
	if (q.dist == 0 || q.gPos1 == -1 || q.gPos2 == -1) return;
	const bool h1 = q.gPos1 < a.ix.G, h2 = q.gPos2 < a.ix.G;
	if (h1 == h2 && q.dist <= 1000) return;
	const int64_t k = (int64_t)mc_atomic_add(bump, (mc_u64)1);
	if (k < cap) { DiscRec d; d.pair = p; d.v = q; out[k] = d; }
}

MC_HD void chunkstat_body(int64_t c, int lane, int nl, const PipeArgs& a)
{
	if (!a.active[c] || a.st->overflow) return;   // overflow: the attempt is repeated, nothing of it is used
	const int64_t rb = c * MC_CHUNK_READS;
	int64_t re = rb + MC_CHUNK_READS; if (re > a.n_reads) re = a.n_reads;
	int mapped = 0, paired = 0, dsum = 0, lsum = 0;   // a chunk holds 100 pairs with dist <= 1000: the sums fit an int
	int lo = -2147483647, hi = 2147483647;
	for (int64_t r = rb + lane; r < re; r += nl) if (a.rsum[r].score > 0) mapped++;
	if (a.pr.paired)
	{
		for (int64_t p = rb / 2 + lane; p < re / 2; p += nl)
		{
			if (a.est_lo[p] > lo) lo = a.est_lo[p];
			if (a.est_hi[p] < hi) hi = a.est_hi[p];
			const mc_pair_out q = a.pair_out[p];
			if (q.dist == 0 || q.gPos1 == -1 || q.gPos2 == -1) continue;
			const bool h1 = q.gPos1 < a.ix.G, h2 = q.gPos2 < a.ix.G;
			if (h1 != h2) continue;                 // inversion candidates (handled on the host in file order)
			if (q.dist > 1000) continue;            // translocation candidates (MinTranslocationSize)
			paired++; dsum += (int)q.dist;
			lsum += (int)((a.roff[2 * p + 1] - a.roff[2 * p]) + (a.roff[2 * p + 2] - a.roff[2 * p + 1]));
		}
	}
	mapped = mc_warp_sum(mapped); paired = mc_warp_sum(paired); dsum = mc_warp_sum(dsum); lsum = mc_warp_sum(lsum);
	lo = mc_warp_max(lo); hi = -mc_warp_max(-hi);
	if (lane == 0)
	{
		mc_chunk_out o; o.n_reads = (int32_t)(re - rb); o.mapped = mapped; o.paired = paired; o.est_distance = a.pr.paired ? a.est[c] : 0; o.dist_sum = dsum; o.len_sum = lsum;
		a.chunk_out[c] = o; a.chunk_lo[c] = lo; a.chunk_hi[c] = hi;
	}
}

#endif
