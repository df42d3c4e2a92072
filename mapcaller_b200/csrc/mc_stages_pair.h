// Paired-end candidate pairing and mate rescue (included by mc_stages.h).
#ifndef MC_STAGES_PAIR_H
#define MC_STAGES_PAIR_H

MC_HD int64_t cand_posdiff(const PipeArgs& a, const Cand& c) { const SPair& p = a.pairs[c.pbeg]; return p.gpos - p.rpos; }

// RemoveRedundantAlnCan (reference src/ReadMapping.cpp:228-242) on the live scores
MC_HD void remove_redundant(int32_t* sc, int n)
{
	if (n <= 1) return;
	int mx = 0;
	for (int i = 0; i < n; i++) if (sc[i] > mx) mx = sc[i];
	for (int i = 0; i < n; i++) if (sc[i] < mx) sc[i] = 0;
}

// MaskUnPairedAlnCan (reference src/ReadMapping.cpp:305-322)
MC_HD void mask_unpaired(int32_t* s0, int32_t* p0, int n0, int32_t* s1, int32_t* p1, int n1)
{
	int mx = 0;
	for (int i = 0; i < n0; i++) if (p0[i] != -1 && mx < s0[i] + s1[p0[i]]) mx = s0[i] + s1[p0[i]];
	for (int i = 0; i < n0; i++) if (p0[i] == -1 || s0[i] + s1[p0[i]] < mx) s0[i] = 0;
	for (int j = 0; j < n1; j++) if (p1[j] == -1 || s1[j] + s0[p1[j]] < mx) s1[j] = 0;
}

// single-end reads: RemoveRedundantAlnCan only (reference src/ReadMapping.cpp:583)
MC_HD void single_body(int64_t r, const PipeArgs& a)
{
	if (!a.active[pa_chunk_of_read(r)]) return;
	const int64_t co = pa_cand_off(a, r);
	const int n = a.ncand0[r];
	a.ncand[r] = n;
	for (int i = 0; i < n; i++) { a.cscore[co + i] = a.cands[co + i].score; a.cpaired[co + i] = -1; }
	remove_redundant(a.cscore + co, n);
}

// one thread per read pair: CheckPairedAlignmentDistance (reference src/ReadMapping.cpp:244-303) and,
// when something paired, MaskUnPairedAlnCan.  Pairs that found nothing are queued for rescue_body.
// Also records the interval [est_lo, est_hi] of EstiDistance values for which every distance test of
// this pair has the same outcome (used to validate the avgDist speculation).
MC_HD void pair_body(int64_t p, const PipeArgs& a)
{
	const int64_t r0 = 2 * p, r1 = r0 + 1;
	const int chunk = pa_chunk_of_read(r0);
	if (!a.active[chunk]) return;
	const int est = a.est[chunk];
	const int64_t c0 = pa_cand_off(a, r0), c1 = pa_cand_off(a, r1);
	const int n0 = a.ncand0[r0], n1 = a.ncand0[r1];
	a.ncand[r0] = n0; a.ncand[r1] = n1;
	int32_t *s0 = a.cscore + c0, *s1 = a.cscore + c1, *p0 = a.cpaired + c0, *p1 = a.cpaired + c1, *t0 = a.ctmp + c0;
	for (int i = 0; i < n0; i++) { s0[i] = a.cands[c0 + i].score; p0[i] = -1; }
	for (int j = 0; j < n1; j++) { s1[j] = a.cands[c1 + j].score; p1[j] = -1; }
	if (n0 * n1 > 100) { remove_redundant(s0, n0); remove_redundant(s1, n1); }
	int lo = -2147483647, hi = 2147483647;
	int64_t max_score = 0;
	for (int i = 0; i < n0; i++)
	{
		t0[i] = -1;
		if (s0[i] == 0) continue;
		const int64_t d0 = cand_posdiff(a, a.cands[c0 + i]);
		int best = -1, ps = 0;
		for (int j = 0; j < n1; j++)
		{
			if (s1[j] == 0) continue;
			const int64_t d = cand_posdiff(a, a.cands[c1 + j]) - d0;
			if (d < 0) continue;
			if (d < est)
			{
				if (d + 1 > lo) lo = (int)(d + 1);
				if (s1[j] > ps) { best = j; ps = s1[j]; }
			}
			else if (d < hi) hi = (int)(d > 2147483646 ? 2147483646 : d);
		}
		t0[i] = best;
		if (best != -1 && s0[i] + s1[best] > max_score) max_score = s0[i] + s1[best];
	}
	int paired = 0;
	if (max_score > 0)
		for (int i = 0; i < n0; i++)
			if (t0[i] != -1 && s0[i] + s1[t0[i]] == max_score) { paired++; p0[i] = t0[i]; p1[t0[i]] = i; }
	if (paired == 0)
	{
		// AlignmentRescue looks at windows sized by EstDist: any other value may change its outcome
		// unless it returns before using it (both mates below the score floor, src/AlignmentRescue.cpp:43)
		int b0 = 0, b1 = 0;
		for (int i = 0; i < n0; i++) if (s0[i] > b0) b0 = s0[i];
		for (int j = 0; j < n1; j++) if (s1[j] > b1) b1 = s1[j];
		const int l0 = (int)(a.roff[r0 + 1] - a.roff[r0]), l1 = (int)(a.roff[r1 + 1] - a.roff[r1]);
		if (b0 < (l0 >> 2) && b1 < (l1 >> 2)) { remove_redundant(s0, n0); remove_redundant(s1, n1); a.pair_flag[p] = 0; }
		else
		{
			a.pair_flag[p] = 1; lo = est; hi = est;
			int64_t k = (int64_t)mc_atomic_add(a.rtask_bump, (mc_u64)1);
			a.rtask[k] = (int32_t)p;
		}
	}
	else { mask_unpaired(s0, p0, n0, s1, p1, n1); a.pair_flag[p] = 0; }
	a.est_lo[p] = lo; a.est_hi[p] = hi;
}

// ---- rescue ---------------------------------------------------------------------------------------
// AlignmentRescue (reference src/AlignmentRescue.cpp:28-111) matches every 8-mer of the unplaced mate
// against every 8-mer of a reference window, sorts the hits by diagonal and merges runs of consecutive
// hits into seeds of >= 10 bases (src/KmerAnalysis.cpp:57-163).  A run of k consecutive 8-mer hits on
// one diagonal is exactly a maximal stretch of k+7 matching bases inside the window, so the same seeds
// come out of a direct scan of each diagonal; the best diagonal is the one with the largest total seed
// length, smallest PosDiff first on ties (IdentifyBestAlnCan, src/AlignmentRescue.cpp:3-26).
// Read bases that are not ACGT/acgt never match (the reference skips windows with 'N').
struct RescueHit { int score; int64_t diag; };

// scans diagonal d (window offset minus read offset) and returns the sum of seed lengths; when out != 0
// also writes the seeds
MC_HD int rescue_scan_diag(const PipeArgs& a, const uint8_t* rs, int rlen, int64_t left, int slen, int d, SPair* out, int* nout)
{
	int lo = d < 0 ? -d : 0;                 // first read offset whose window base exists
	int hi = rlen; if (hi > slen - d) hi = slen - d;
	int total = 0, run = 0, n = 0;
	for (int q = lo; q <= hi; q++)
	{
		bool match = false;
		if (q < hi)
		{
			int c = mc_nt4(rs[q]);
			int64_t g = left + d + q;
			match = c <= 3 && g >= 0 && c == mc_ref_code(a.ix, g);
		}
		if (match) run++;
		else
		{
			if (run >= 10)
			{
				total += run;
				if (out) { SPair s; s.rpos = q - run; s.gpos = left + d + (q - run); s.len = run; out[n] = s; }
				n++;
			}
			run = 0;
		}
	}
	if (nout) *nout = n;
	return total;
}

// tries to place `rs` (the mate without a partner) inside [left, right); appends a candidate to read `rt`
MC_HD bool rescue_try(const PipeArgs& a, int64_t rt, const uint8_t* rs, int rlen, int64_t left, int64_t right, int floor_score,
                      int anchor_idx, int32_t* new_idx)
{
	if (right > a.ix.twoG) right = a.ix.twoG;
	int i1 = mc_chrom_lower_bound(a.ix, left), i2 = mc_chrom_lower_bound(a.ix, right);
	if (i1 >= a.ix.n_end || i2 >= a.ix.n_end) return false; // the reference dereferences end() here; treated as "different chromosome"
	if (a.ix.chrom_id[i1] != a.ix.chrom_id[i2]) return false;
	const int64_t sl = right - left;
	if (sl < rlen) return false;
	const int slen = (int)sl;
	int best = 0, bd = 0;
	for (int d = -(rlen - 8); d <= slen - 8; d++)
	{
		int sc = rescue_scan_diag(a, rs, rlen, left, slen, d, 0, 0);
		if (sc > best) { best = sc; bd = d; }
	}
	if (best == 0 || best <= floor_score) return false;
	int n = 0;
	rescue_scan_diag(a, rs, rlen, left, slen, bd, 0, &n);
	const int64_t pb = (int64_t)mc_atomic_add(a.pair_bump, (mc_u64)n);
	if (pb + n > a.pair_cap) { mc_atomic_add(&a.st->overflow, (mc_u64)1); return false; }
	rescue_scan_diag(a, rs, rlen, left, slen, bd, a.pairs + pb, &n);
	const int64_t co = pa_cand_off(a, rt);
	const int k = a.ncand[rt];
	if (k >= pa_cand_cap(a, rt)) { mc_atomic_add(&a.st->overflow, (mc_u64)1); return false; }
	Cand c; c.score = best; c.pbeg = (int32_t)pb; c.pend = (int32_t)(pb + n);
	a.cands[co + k] = c; a.cscore[co + k] = best; a.cpaired[co + k] = anchor_idx;
	a.ncand[rt] = k + 1;
	*new_idx = k;
	return true;
}

MC_HD void rescue_body(int64_t t, const PipeArgs& a)
{
	const int64_t p = a.rtask[t];
	const int64_t r0 = 2 * p, r1 = r0 + 1;
	const int est = a.est[pa_chunk_of_read(r0)];
	const int64_t c0 = pa_cand_off(a, r0), c1 = pa_cand_off(a, r1);
	const int l0 = (int)(a.roff[r0 + 1] - a.roff[r0]), l1 = (int)(a.roff[r1 + 1] - a.roff[r1]);
	int32_t *s0 = a.cscore + c0, *s1 = a.cscore + c1, *p0 = a.cpaired + c0, *p1 = a.cpaired + c1;
	int n0 = a.ncand[r0], n1 = a.ncand[r1];
	int b0 = 0, b1 = 0;
	for (int i = 0; i < n0; i++) if (s0[i] > b0) b0 = s0[i];
	for (int j = 0; j < n1; j++) if (s1[j] > b1) b1 = s1[j];
	int strat = (b0 - b1 > (l1 >> 2)) ? 1 : (b1 - b0 > (l0 >> 2)) ? 2 : 3;
	int rescued = 0;
	if (strat == 1 || strat == 3) // place mate 2 next to mate 1's candidates
	{
		const int thr = b0 >> 1;
		for (int i = 0; i < n0; i++)
		{
			if (s0[i] < thr || p0[i] != -1) continue;
			const int64_t d = cand_posdiff(a, a.cands[c0 + i]);
			int32_t k;
			if (rescue_try(a, r1, a.seq + a.roff[r1], l1, d, d + (int64_t)(uint32_t)est + l1, b1, i, &k)) { p0[i] = k; rescued++; }
		}
	}
	if (strat == 2 || strat == 3) // place mate 1 next to mate 2's candidates
	{
		const int thr = b1 >> 1;
		const int n1_now = a.ncand[r1];
		for (int j = 0; j < n1_now; j++)
		{
			if (s1[j] < thr || p1[j] != -1) continue;
			const int64_t d = cand_posdiff(a, a.cands[c1 + j]);
			int32_t k;
			if (rescue_try(a, r0, a.seq + a.roff[r0], l0, d - (int64_t)(uint32_t)est, d + l0, b0, j, &k)) { p1[j] = k; rescued++; }
		}
	}
	n0 = a.ncand[r0]; n1 = a.ncand[r1];
	if (rescued == 0) { remove_redundant(s0, n0); remove_redundant(s1, n1); }
	else mask_unpaired(s0, p0, n0, s1, p1, n1);
}

#endif
