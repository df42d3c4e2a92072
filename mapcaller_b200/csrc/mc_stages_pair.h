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
	a.read_redo[r] = a.active[pa_chunk_of_read(r)];
	if (!a.active[pa_chunk_of_read(r)]) return;
	const int64_t co = pa_cand_off(a, r);
	const int n = a.ncand0[r];
	a.ncand[r] = n;
	for (int i = 0; i < n; i++) { a.cscore[co + i] = a.cands[co + i].score; a.cpaired[co + i] = -1; }
	remove_redundant(a.cscore + co, n);
}

// one thread per read pair: CheckPairedAlignmentDistance (reference src/ReadMapping.cpp:244-303) and,
// when something paired, MaskUnPairedAlnCan.  Pairs that found nothing are queued for the rescue stage.
// Also records the interval [est_lo, est_hi] of EstiDistance values for which every distance test of
// this pair has the same outcome (used to validate the avgDist speculation).
MC_HD void pair_body(int64_t p, const PipeArgs& a)
{
	const int64_t r0 = 2 * p, r1 = r0 + 1;
	const int chunk = pa_chunk_of_read(r0);
	a.read_redo[r0] = 0; a.read_redo[r1] = 0;
	if (!a.active[chunk]) return;
	const int est = a.est[chunk];
	// a result computed earlier in this batch stays valid while the new EstiDistance is inside its interval
	if ((a.pair_flag[p] & 1) && a.est_lo[p] <= est && est <= a.est_hi[p]) return;
	a.pair_flag[p] = 1; a.read_redo[r0] = 1; a.read_redo[r1] = 1;
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
		if (b0 < (l0 >> 2) && b1 < (l1 >> 2)) { remove_redundant(s0, n0); remove_redundant(s1, n1); }
		else
		{
			int64_t k = mc_bump_alloc(a.rtask_bump, 1u);   // rcommit_body narrows [lo, hi] further
			a.rtask[k] = (int32_t)p;
		}
	}
	else mask_unpaired(s0, p0, n0, s1, p1, n1);
	a.est_lo[p] = lo; a.est_hi[p] = hi;
}

// ---- rescue ---------------------------------------------------------------------------------------
// AlignmentRescue (reference src/AlignmentRescue.cpp:28-111) joins the 8-mers of the unplaced mate with the
// 8-mers of a reference window on their 16-bit word id, sorts the hits by (diagonal, read offset) and merges
// runs of hits with consecutive read offsets into seeds of >= 10 bases (src/KmerAnalysis.cpp:57-163); the
// best diagonal is the one with the largest total seed length, smallest diagonal first on ties
// (IdentifyBestAlnCan, src/AlignmentRescue.cpp:3-26).  The sort + join is equivalent to walking each
// diagonal with the read's word list, which is what rescue_scan_diag does.
//
// The read's word list has to be produced exactly as CreateKmerVecFromReadSeq does (src/KmerAnalysis.cpp:57-103),
// including what happens after an 'N': the scan restarts with a fresh word at [t-8, t), but the loop
// increment then steps over base t, so every later word of the read is labelled one position early and the
// first seven of them straddle the skipped base.  Rescued seeds inherit those labels.
struct KmerEnt { int32_t label; uint32_t wid; };

MC_HD uint32_t kmer_fresh_id(const uint8_t* s, int pos)
{
	uint32_t id = 0;
	for (int i = 0; i < 8; i++) id = (id << 2) + (uint32_t)mc_nt4(s[pos + i]);
	return id;
}

// returns the number of words written to `out` (at most len)
MC_HD int kmer_list_of_read(const uint8_t* s, int len, KmerEnt* out)
{
	int n = 0, tail = 0, cnt = 0;
	while (cnt < 8 && tail < len) { if (s[tail++] != 'N') cnt++; else cnt = 0; }
	if (cnt != 8) return 0;
	int head = tail - 8;
	uint32_t wid = kmer_fresh_id(s, head);
	out[n].label = head; out[n].wid = wid; n++;
	for (head += 1; tail < len; head++, tail++)
	{
		if (s[tail] != 'N')
		{
			wid = ((wid & 0x3FFFu) << 2) + (uint32_t)mc_nt4(s[tail]);
			out[n].label = head; out[n].wid = wid; n++;
		}
		else
		{
			cnt = 0; tail++;
			while (cnt < 8 && tail < len) { if (s[tail++] != 'N') cnt++; else cnt = 0; }
			if (cnt != 8) break;
			head = tail - 8; wid = kmer_fresh_id(s, head);
			out[n].label = head; out[n].wid = wid; n++;
			// the for-increment now advances head AND tail: base `tail` is never rolled in (reference behaviour)
		}
	}
	return n;
}

// The same list by all lanes of the group when the read is plain ACGT/acgt (no 'N', nothing the sequential rolling
// formula would treat specially): word p is the fresh id of bases [p, p+8) and is labelled p.  Otherwise lane 0 runs the
// sequential routine.  Returns the number of words to every lane through lanebuf[0].
MC_HD int kmer_list_coop(const uint8_t* s, int len, KmerEnt* out, int lane, int nl, int32_t* lanebuf)
{
	int bad = 0;
	for (int i = lane; i < len; i += nl) if (mc_nt4(s[i]) > 3) bad = 1;
	bad = mc_group_any(bad);
	if (!bad)
	{
		const int n = len >= 8 ? len - 7 : 0;
		for (int p = lane; p < n; p += nl) { out[p].label = p; out[p].wid = kmer_fresh_id(s, p); }
		MC_GROUP_SYNC();
		return n;
	}
	if (lane == 0) lanebuf[0] = kmer_list_of_read(s, len, out);
	MC_GROUP_SYNC();
	const int n = lanebuf[0];
	MC_GROUP_SYNC();
	return n;
}

// word id of the 8-mer at offset g of a staged window (codes 0..3, 4 = outside the text); 0xFFFFFFFF if it touches the outside
MC_HD uint32_t win_kmer_id(const uint8_t* wref, int g)
{
	uint32_t w = 0, bad = 0;
	for (int i = 0; i < 8; i++) { const uint32_t c = wref[g + i]; bad |= c & 4u; w = (w << 2) + (c & 3u); }
	return bad ? 0xFFFFFFFFu : w;
}

// walks diagonal d (window offset minus word label); returns the sum of seed lengths, optionally writes the seeds
MC_HD int rescue_scan_diag(const uint32_t* wkid0, const KmerEnt* km, int nk, int64_t left, int slen, int d, SPair* out, int* nout)
{
	int total = 0, run = 0, first = 0, n = 0, prev_label = -2;
	for (int i = 0; i <= nk; i++)
	{
		bool hit = false;
		if (i < nk)
		{
			const int g = km[i].label + d;
			if (g >= 0 && g <= slen - 8) hit = wkid0[g] == km[i].wid;   // wkid0[g] = word id at window offset g
		}
		if (hit && run > 0 && km[i].label == prev_label + 1) run++;
		else
		{
			if (run >= 3)
			{
				const int l = 7 + run;
				total += l;
				if (out) { SPair s; s.rpos = km[first].label; s.gpos = left + d + km[first].label; s.len = l; out[n] = s; }
				n++;
			}
			run = hit ? 1 : 0; first = i;
		}
		prev_label = hit ? km[i].label : -2;
	}
	if (nout) *nout = n;
	return total;
}

#if MC_DEV_ONLY
// The same walk with the 32 lanes of a warp side by side: lane l tests word c0 + l of the mate's list against diagonal d, two
// ballots give the hit mask H and the mask C of hits that CONTINUE a run (previous word hit as well and labelled one lower);
// runs are then read off the masks with bit scans - uniform code, a few instructions per run instead of a dozen per word of
// the list on one lane while 31 wait.  Returns the score of the diagonal (every lane); Hs / Cs, if given, receive the masks.
static __device__ __forceinline__ int rescue_diag_warp(const uint32_t* wkid0, const KmerEnt* km, int nk, int slen, int d, int l, uint32_t* Hs, uint32_t* Cs)
{
	const unsigned full = 0xffffffffu;
	int total = 0, run = 0, carry_lab = -2, carry_hit = 0;
	for (int c0 = 0, wi = 0; c0 < nk; c0 += 32, wi++)
	{
		const int i = c0 + l;
		int lab = 0, hit = 0;
		if (i < nk) { const KmerEnt e = km[i]; lab = e.label; const int g = lab + d; if (g >= 0 && g <= slen - 8) hit = wkid0[g] == e.wid; }
		int plab = __shfl_up_sync(full, lab, 1), phit = __shfl_up_sync(full, hit, 1);
		if (l == 0) { plab = carry_lab; phit = carry_hit; }
		const int cont = hit && phit && lab == plab + 1;
		const uint32_t H = __ballot_sync(full, hit), C = __ballot_sync(full, cont);
		carry_lab = __shfl_sync(full, lab, 31); carry_hit = __shfl_sync(full, hit, 31);
		if (Hs && l == 0) { Hs[wi] = H; Cs[wi] = C; }
		int pos = 0;
		if (run > 0)                                          // a run reaches over from the previous 32 words
		{
			const int k = C == full ? 32 : __ffs((int)~C) - 1;
			run += k; pos = k;
			if (k == 32) continue;
			if (run >= 3) total += 7 + run;
			run = 0;
		}
		while (pos < 32)
		{
			const uint32_t rest = H & (full << pos);
			if (!rest) break;
			const int s0 = __ffs((int)rest) - 1;
			const int k = s0 < 31 ? __ffs((int)~(C >> (s0 + 1))) - 1 : 0;
			run = 1 + k; pos = s0 + 1 + k;
			if (pos >= 32) break;                             // may go on in the next 32 words
			if (run >= 3) total += 7 + run;
			run = 0;
		}
	}
	if (run >= 3) total += 7 + run;
	return total;
}
// the seeds of a diagonal from its masks (one lane): counts them, and writes them when `out` is given
static __device__ __forceinline__ int rescue_seeds_from_masks(const uint32_t* Hs, const uint32_t* Cs, int nk, const KmerEnt* km, int64_t left, int d, SPair* out)
{
	int n = 0, run = 0, first = 0;
	const int nw = (nk + 31) >> 5;
	for (int wi = 0; wi <= nw; wi++)
	{
		const uint32_t H = wi < nw ? Hs[wi] : 0u, C = wi < nw ? Cs[wi] : 0u;
		int pos = 0;
		if (run > 0)
		{
			const int k = C == 0xffffffffu ? 32 : __ffs((int)~C) - 1;
			run += k; pos = k;
			if (k == 32) continue;
			if (run >= 3) { if (out) { SPair sp; sp.rpos = km[first].label; sp.gpos = left + d + km[first].label; sp.len = 7 + run; out[n] = sp; } n++; }
			run = 0;
		}
		while (pos < 32)
		{
			const uint32_t rest = H & (0xffffffffu << pos);
			if (!rest) break;
			const int s0 = __ffs((int)rest) - 1;
			const int k = s0 < 31 ? __ffs((int)~(C >> (s0 + 1))) - 1 : 0;
			run = 1 + k; first = wi * 32 + s0; pos = s0 + 1 + k;
			if (pos >= 32) break;
			if (run >= 3) { if (out) { SPair sp; sp.rpos = km[first].label; sp.gpos = left + d + km[first].label; sp.len = 7 + run; out[n] = sp; } n++; }
			run = 0;
		}
	}
	return n;
}
#endif

#define MC_RESCUE_HASH 2048  // chain heads of the read's 8-mer table (low 11 bits of the 16-bit word id)
#define MC_RESCUE_MARGIN 32   // how far beyond the moving window edge hits are looked for (validity interval of EstiDistance)

// Tries to place a mate (word list km) inside the window an anchored candidate of the other mate implies and appends a
// candidate to read `rt` (one iteration of the loops of AlignmentRescue, reference src/AlignmentRescue.cpp:55-79 / :84-106).
//   dir 0: window [d, d + est + rlen)   - the right edge moves with EstiDistance
//   dir 1: window [d - est, d + rlen)   - the left edge moves
// `nl` lanes cooperate (a thread block on the GPU): (1) every reference 8-mer of the window (plus a margin beyond the moving
// edge) is looked up in a hash table of the read's words (2048 chain heads over the low 11 bits, built by rwin_body) - a lane
// per window offset, the matching words of its chain are histogrammed per diagonal, (2) only diagonals with >= 3 hits (a seed
// needs a run of 3) are walked exactly, (3) the best diagonal is reduced over the lanes, (4) lane 0 appends the candidate.
// *iv_lo / *iv_hi are narrowed to the EstiDistance values for which this call provably does the same: the moving edge stays
// between the same chromosome ends and no 8-mer hit enters or leaves the window.  All lanes return the same values.
MC_HD bool rescue_try(const PipeArgs& a, int lane, int nl, RWin* res, const KmerEnt* km, int nk, int rlen, int dir, int64_t d, int est,
                      int floor_score, uint32_t* hits, int64_t n_hits, int32_t* lanebuf, const int32_t* khead, const int32_t* knext, uint32_t* wkid, int* iv_lo, int* iv_hi)
{
	const int M = MC_RESCUE_MARGIN;
	int64_t left = dir == 0 ? d : d - (int64_t)(uint32_t)est;
	int64_t right = dir == 0 ? d + (int64_t)(uint32_t)est + rlen : d + rlen;
	const bool clipped = right > a.ix.twoG;
	if (clipped) right = a.ix.twoG;
	int lo = 0, hi = est + M;
	const int i1 = mc_chrom_lower_bound(a.ix, left), i2 = mc_chrom_lower_bound(a.ix, right);
	{
		// the moving edge has to stay between the same two chromosome ends (and on the same side of the 2G clip)
		bool pin = false;
		if (dir == 0)
		{
			if (clipped) { const int64_t v = a.ix.twoG - d - rlen; if (v > lo) lo = (int)(v > est ? est : v); }
			else if (i2 >= a.ix.n_end) pin = true;
			else
			{
				const int64_t up = a.ix.chrom_end[i2] - d - rlen; if (up < hi) hi = (int)up;
				if (i2 > 0) { const int64_t dn = a.ix.chrom_end[i2 - 1] - d - rlen + 1; if (dn > lo) lo = (int)dn; }
			}
		}
		else
		{
			if (clipped || i1 >= a.ix.n_end) pin = true;
			else
			{
				const int64_t dn = d - a.ix.chrom_end[i1]; if (dn > lo) lo = (int)dn;
				if (i1 > 0) { const int64_t up = d - a.ix.chrom_end[i1 - 1] - 1; if (up < hi) hi = (int)up; }
			}
		}
		if (pin) { lo = est; hi = est; }
		if (lo > est) lo = est;
		if (hi < est) hi = est;
	}
	*iv_lo = *iv_lo > lo ? *iv_lo : lo; *iv_hi = *iv_hi < hi ? *iv_hi : hi;
	// A window that reaches 2G makes the reference read PosChrIdMap.end()->second (src/AlignmentRescue.cpp:62-63).  What it finds
	// there, with GCC's layout of the globals of src/main.cpp, is the zeroed first word of the map defined next: chromosome id
	// 0 - which is also the id of the chromosome that ends at 2G.  Reproduced as such (checked against oracle/_ref).
	if (i1 >= a.ix.n_end) return false;
	if (a.ix.chrom_id[i1] != (i2 >= a.ix.n_end ? 0 : a.ix.chrom_id[i2])) return false;
	const int64_t sl = right - left;
	if (sl < rlen) return false;
	const int slen = (int)sl;
	const int dmin = -(rlen - 8), ndiag = slen - 8 - dmin + 1;
	for (int i = lane; i < ndiag; i += nl) hits[i] = 0;
	// window offsets scanned: the window itself plus the margin beyond the moving edge
	const int gs = dir == 1 ? -M : 0, ge = dir == 0 ? slen - 8 + M : slen - 8;
	// stage the 2-bit codes of the scanned stretch once (lanes read consecutive bases); 4 marks positions outside the text
	const uint32_t* wkid0 = wkid - gs;                                          // word id of every scanned offset
	uint8_t* wref = (uint8_t*)(wkid + n_hits); const uint8_t* wref0 = wref - gs;
	for (int i = lane; i < ge - gs + 8; i += nl) { const int64_t pos = left + gs + i; wref[i] = (pos < 0 || pos >= a.ix.twoG) ? 4 : (uint8_t)mc_ref_code(a.ix, pos); }
	MC_GROUP_SYNC();
	int in_far = -1, out_near = 1 << 30;   // inside hit closest to the moving edge (distance from it), margin hit closest to it
	for (int g = gs + lane; g <= ge; g += nl)
	{
		const uint32_t w = win_kmer_id(wref0, g);
		wkid[g - gs] = w;
		if (w == 0xFFFFFFFFu) continue;
		const bool inside = g >= 0 && g <= slen - 8;
		for (int i = khead[w & (MC_RESCUE_HASH - 1)]; i >= 0; i = knext[i])      // the read's words that share the low bits
		{
			if (km[i].wid != w) continue;
			if (inside)
			{
				mc_atomic_add(&hits[g - km[i].label - dmin], 1u);
				const int dist = dir == 0 ? slen - 8 - g : g;          // how far the edge may retreat before this hit drops out
				if (in_far < 0 || dist < in_far) in_far = dist;
			}
			else { const int dist = dir == 0 ? g - (slen - 8) : -g; if (dist < out_near) out_near = dist; }   // >= 1: advance that lets it in
		}
	}
	lanebuf[4 * nl + lane] = in_far; lanebuf[5 * nl + lane] = out_near;
	MC_GROUP_SYNC();
	int best = 0, bd = 0;
#if MC_DEV_ONLY
	{
		// a WARP per diagonal with >= 3 hits: the warps take 32 diagonals at a time, the qualifying ones of a turn one after the
		// other (ascending, so the first maximum wins inside a warp); every lane of a warp ends up with the warp's best
		const int l = lane & 31, nwarp = nl >> 5;
		for (int i0 = (lane >> 5) << 5; i0 < ndiag; i0 += nl)
		{
			const int i = i0 + l;
			unsigned q = __ballot_sync(0xffffffffu, i < ndiag && hits[i] >= 3);
			while (q)
			{
				const int dd = i0 + (__ffs((int)q) - 1) + dmin; q &= q - 1;
				const int sc = rescue_diag_warp(wkid0, km, nk, slen, dd, l, nullptr, nullptr);
				if (sc > best) { best = sc; bd = dd; }
			}
		}
		// closest hits to the moving edge: reduced inside the warp, then over the warps
		const unsigned fw = __reduce_min_sync(0xffffffffu, in_far < 0 ? 0xffffffffu : (unsigned)in_far), ow = __reduce_min_sync(0xffffffffu, (unsigned)out_near);
		MC_GROUP_SYNC();                                    // lanebuf[4 nl ..] (per-lane in_far / out_near of the scan above) is no longer needed
		if (l == 0) { int32_t* wb = lanebuf + 4 * (lane >> 5); wb[0] = best; wb[1] = bd; wb[2] = (int32_t)fw; wb[3] = (int32_t)ow; }
		MC_GROUP_SYNC();
		best = 0; bd = 0; in_far = -1; out_near = 1 << 30;
		for (int w = 0; w < nwarp; w++)
		{
			const int sc = lanebuf[4 * w], dd = lanebuf[4 * w + 1];
			if (sc > best || (sc == best && sc > 0 && dd < bd)) { best = sc; bd = dd; }
			const unsigned f = (unsigned)lanebuf[4 * w + 2]; const int o = lanebuf[4 * w + 3];
			if (f != 0xffffffffu && (in_far < 0 || (int)f < in_far)) in_far = (int)f;
			if (o < out_near) out_near = o;
		}
		MC_GROUP_SYNC();
	}
#else
	for (int i = lane; i < ndiag; i += nl)
	{
		if (hits[i] < 3) continue;
		const int sc = rescue_scan_diag(wkid0, km, nk, left, slen, i + dmin, 0, 0);
		if (sc > best) { best = sc; bd = i + dmin; }   // ascending diagonals inside a lane: first maximum wins
	}
	lanebuf[2 * lane] = best; lanebuf[2 * lane + 1] = bd;
	MC_GROUP_SYNC();
	best = 0; bd = 0; in_far = -1; out_near = 1 << 30;
	for (int l = 0; l < nl; l++)
	{
		const int sc = lanebuf[2 * l], dd = lanebuf[2 * l + 1];
		if (sc > best || (sc == best && sc > 0 && dd < bd)) { best = sc; bd = dd; }
		const int f = lanebuf[4 * nl + l], o = lanebuf[5 * nl + l];
		if (f >= 0 && (in_far < 0 || f < in_far)) in_far = f;
		if (o < out_near) out_near = o;
	}
	MC_GROUP_SYNC();
#endif
	if (!clipped || dir == 1)
	{
		if (in_far >= 0 && est - in_far > *iv_lo) *iv_lo = est - in_far;
		if (out_near < (1 << 30) && est + out_near - 1 < *iv_hi) *iv_hi = est + out_near - 1;
	}
	if (best == 0 || best <= floor_score) return false;
#if MC_DEV_ONLY
	if (lane < 32)
	{
		// the seeds of the winning diagonal: its masks once more (warp 0), then lane 0 reads the runs off them twice - count, reserve, write
		uint32_t* Hs = (uint32_t*)lanebuf + 16; uint32_t* Cs = Hs + ((MC_MAX_RLEN + 31) >> 5);     // fits below lanebuf[4 nl]: 2 x 125 words
		rescue_diag_warp(wkid0, km, nk, slen, bd, lane, Hs, Cs);
		__syncwarp();
		if (lane == 0)
		{
			int ok = 1;
			const int n = rescue_seeds_from_masks(Hs, Cs, nk, km, left, bd, nullptr);
			const int64_t pb = (int64_t)mc_atomic_add(a.pair_bump, (mc_u64)n);
			if (pb + n > a.pair_cap) { mc_atomic_or(&a.st->overflow, (mc_u64)1 << 0); ok = 0; }
			if (ok) { rescue_seeds_from_masks(Hs, Cs, nk, km, left, bd, a.pairs + pb); res->score = best; res->pbeg = (int32_t)pb; res->n = n; }
			lanebuf[6 * nl] = ok;
		}
	}
#else
	if (lane == 0)
	{
		int n = 0, ok = 1;
		rescue_scan_diag(wkid0, km, nk, left, slen, bd, 0, &n);
		const int64_t pb = (int64_t)mc_atomic_add(a.pair_bump, (mc_u64)n);
		if (pb + n > a.pair_cap) { mc_atomic_or(&a.st->overflow, (mc_u64)1 << 0); ok = 0; }
		if (ok) { rescue_scan_diag(wkid0, km, nk, left, slen, bd, a.pairs + pb, &n); res->score = best; res->pbeg = (int32_t)pb; res->n = n; }
		lanebuf[6 * nl] = ok;
	}
#endif
	MC_GROUP_SYNC();
	const int ok = lanebuf[6 * nl];
	MC_GROUP_SYNC();
	return ok != 0;
}

// AlignmentRescue (reference src/AlignmentRescue.cpp:28-111) in three steps, because one pair may imply dozens of windows
// (a mate anchored in a repeat family) and the windows are independent of each other: the candidates they test, the score
// floors (the mates' best scores on entry) and the skip tests do not change while the reference loops over them.
//   rwenum_body   one thread per rescue pair: strategy, thresholds, the list of windows
//   rwin_body     one thread block per window: the search itself
//   rcommit_body  one thread per rescue pair: appends the found candidates in the reference's order, links the partners,
//                 masks, and intersects the EstiDistance validity intervals
MC_HD void rwenum_body(int64_t t, const PipeArgs& a)
{
	if (a.rtask_begin + t >= (int64_t)*a.rtask_bump) return;
	const int64_t p = a.rtask[a.rtask_begin + t];
	const int64_t r0 = 2 * p, r1 = r0 + 1;
	const int64_t c0 = pa_cand_off(a, r0), c1 = pa_cand_off(a, r1);
	const int l0 = (int)(a.roff[r0 + 1] - a.roff[r0]), l1 = (int)(a.roff[r1 + 1] - a.roff[r1]);
	const int32_t *s0 = a.cscore + c0, *s1 = a.cscore + c1;
	const int n0 = a.ncand[r0], n1 = a.ncand[r1];
	int b0 = 0, b1 = 0;
	for (int i = 0; i < n0; i++) if (s0[i] > b0) b0 = s0[i];
	for (int j = 0; j < n1; j++) if (s1[j] > b1) b1 = s1[j];
	const int strat = (b0 - b1 > (l1 >> 2)) ? 1 : (b1 - b0 > (l0 >> 2)) ? 2 : 3;
	int nw = 0;
	if (strat != 2) for (int i = 0; i < n0; i++) if (s0[i] >= (b0 >> 1)) nw++;   // nothing is paired yet: PairedAlnCanIdx == -1 everywhere
	if (strat != 1) for (int j = 0; j < n1; j++) if (s1[j] >= (b1 >> 1)) nw++;
	const int64_t wb = mc_bump_alloc(a.rwin_bump, (uint32_t)nw);
	a.rw_beg[p] = (int32_t)wb;
	if (wb + nw > a.rwin_cap) { mc_atomic_or(&a.st->overflow, (mc_u64)1 << 0); return; }
	RWin w; w.pair = (int32_t)p; w.pad = 0; w.ok = 0; w.score = 0; w.pbeg = 0; w.n = 0; w.lo = -2147483647; w.hi = 2147483647;
	int k = 0;
	if (strat != 2) for (int i = 0; i < n0; i++) if (s0[i] >= (b0 >> 1)) { w.dir = 0; w.cand = i; w.floor = b1; a.rwin[wb + k++] = w; }
	if (strat != 1) for (int j = 0; j < n1; j++) if (s1[j] >= (b1 >> 1)) { w.dir = 1; w.cand = j; w.floor = b0; a.rwin[wb + k++] = w; }
}

// `fast` / `fast_bytes`: optional on-chip scratch of the calling warp (shared memory on the GPU); windows that need more fall
// back to the global gapped-fill workspace.
MC_HD void rwin_body(int64_t t, int lane, int nl, const PipeArgs& a, uint8_t* fast, int64_t fast_bytes)
{
	RWin* w = a.rwin + a.rwin_begin + t;
	const int64_t p = w->pair;
	const int dir = w->dir;
	const int64_t ra = 2 * p + dir, rm = 2 * p + (1 - dir);       // anchored read, mate to place
	const int est = a.est[pa_chunk_of_read(ra)];
	const int lm = (int)(a.roff[rm + 1] - a.roff[rm]);
	const int64_t d = cand_posdiff(a, a.cands[pa_cand_off(a, ra) + w->cand]);
	const int64_t n_hits = (int64_t)(uint32_t)est + 2 * (int64_t)lm + 2 * MC_RESCUE_MARGIN + 32;
	const int64_t wsn = ((int64_t)(lm + 2) * (int64_t)sizeof(KmerEnt) + n_hits * 4 + (6 * nl + 4) * 4 + MC_RESCUE_HASH * 4 + (int64_t)(lm + 2) * 4 + n_hits * 4 + n_hits + 15) & ~15ll;
	uint8_t* scratch = fast;
	if (!fast || wsn > fast_bytes)
	{
		int64_t ws = 0;
		if (lane == 0) ws = (int64_t)mc_atomic_add(a.dpws_bump, (mc_u64)wsn);
		ws = mc_group_bcast64(ws, lane);
		if (ws + wsn > a.dpws_cap) { if (lane == 0) mc_atomic_or(&a.st->overflow, (mc_u64)1 << 32); return; }
		scratch = a.dpws + ws;
	}
	KmerEnt* km = (KmerEnt*)scratch;
	uint32_t* hits = (uint32_t*)(km + lm + 2); int32_t* lanebuf = (int32_t*)(hits + n_hits);
	// hash table over the low bits of the mate's word ids: khead[h] = last word with that hash, knext[] chains the others
	int32_t* khead = lanebuf + 6 * nl + 4; int32_t* knext = khead + MC_RESCUE_HASH; uint32_t* wkid = (uint32_t*)(knext + lm + 2);
	for (int i = lane; i < MC_RESCUE_HASH; i += nl) khead[i] = -1;
	MC_GROUP_SYNC();
	const int nk = kmer_list_coop(a.seq + a.roff[rm], lm, km, lane, nl, lanebuf);
	for (int i = lane; i < nk; i += nl) knext[i] = mc_atomic_exch(&khead[km[i].wid & (MC_RESCUE_HASH - 1)], i);
	MC_GROUP_SYNC();
	int lo = -2147483647, hi = 2147483647;
	const bool ok = rescue_try(a, lane, nl, w, km, nk, lm, dir, d, est, w->floor, hits, n_hits, lanebuf, khead, knext, wkid, &lo, &hi);
	if (lane == 0) { w->ok = ok ? 1 : 0; w->lo = lo; w->hi = hi; }
}

MC_HD void rcommit_body(int64_t t, const PipeArgs& a)
{
	if (a.rtask_begin + t >= (int64_t)*a.rtask_bump) return;
	const int64_t p = a.rtask[a.rtask_begin + t];
	const int64_t r0 = 2 * p, r1 = r0 + 1;
	const int64_t c0 = pa_cand_off(a, r0), c1 = pa_cand_off(a, r1);
	int32_t *s0 = a.cscore + c0, *s1 = a.cscore + c1, *p0 = a.cpaired + c0, *p1 = a.cpaired + c1;
	int rescued = 0, iv_lo = -2147483647, iv_hi = 2147483647;
	const RWin* w = a.rwin + a.rw_beg[p];
	// the windows of this pair are contiguous and in the reference's loop order; the list ends where another pair's begins
	for (int k = 0; a.rw_beg[p] + k < (int64_t)*a.rwin_bump && w[k].pair == (int32_t)p; k++)
	{
		if (w[k].lo > iv_lo) iv_lo = w[k].lo;
		if (w[k].hi < iv_hi) iv_hi = w[k].hi;
		if (!w[k].ok) continue;
		const int64_t rt = w[k].dir == 0 ? r1 : r0, ct = w[k].dir == 0 ? c1 : c0;
		const int slot = a.ncand[rt];
		if (slot >= pa_cand_cap(a, rt)) { mc_atomic_or(&a.st->overflow, (mc_u64)1 << 40); continue; }
		Cand c; c.score = w[k].score; c.pbeg = w[k].pbeg; c.pend = w[k].pbeg + w[k].n;
		a.cands[ct + slot] = c; a.cscore[ct + slot] = w[k].score; a.cpaired[ct + slot] = w[k].cand;
		a.ncand[rt] = slot + 1;
		if (w[k].dir == 0) p0[w[k].cand] = slot; else p1[w[k].cand] = slot;
		rescued++;
	}
	const int m0 = a.ncand[r0], m1 = a.ncand[r1];
	if (rescued == 0) { remove_redundant(s0, m0); remove_redundant(s1, m1); }
	else mask_unpaired(s0, p0, m0, s1, p1, m1);
	// the pair's outcome holds for every EstiDistance inside both the interval of its distance tests (pair_body) and the
	// intervals of its rescue windows
	if (iv_lo > a.est_lo[p]) a.est_lo[p] = iv_lo;
	if (iv_hi < a.est_hi[p]) a.est_hi[p] = iv_hi;
}

#endif
