// SAM record fields on the device: what GeneratePairedSamStream / GenerateSingleSamStream
// (reference src/SamReport.cpp:324-488) compute per read before they print - SetPairedAlignmentFlag /
// SetSingledAlignmentFlag (:7-84), EvaluateMAPQ (:86-101), GetAlnCoordinate (:119-148) with DetermineCoordinate
// (src/tools.cpp:132-164), GenerateCIGARstring (:171-316), mate position and template length (:428-431, :472-475) - from the
// batch's candidate / fragment / alignment-string arenas where they lie in HBM.  One thread per read, two passes (CIGAR
// text length, then the text at its scanned offset).  Names, bases and qualities never leave the host, which prints
//   QNAME flag RNAME pos mapq CIGAR RNEXT PNEXT TLEN SEQ QUAL [NM:i:] AS:i: XS:i:
// from these fields (-m / bUnique = false, several lines per read, is not covered: one record per read).
#ifndef MC_STAGES_SAM_H
#define MC_STAGES_SAM_H

struct SamArgs {
	DevIndex ix; int32_t paired; int64_t n_reads;
	const int64_t* roff; const int32_t *cand_off, *ncand, *cscore, *cpaired, *corient, *cfrag, *cnfrag; const ReadSum* rsum;
	const mc_frag_out* frags; const uint8_t* aln;
	const uint8_t* mapq_tab;   // [5][MC_MAX_RLEN + 1]: EvaluateMAPQ for score - sub_score = 1..5, evaluated by the host's libm
	mc_sam_rec* out; uint32_t* clen; const int64_t* coff; uint8_t* cigar;
};

struct SamCoor { int32_t chrom; int64_t pos; };

// DetermineCoordinate, src/tools.cpp:132-164 (keys of PosChrIdMap: forward ends of chromosome 0..n-1, then the reverse ends)
MC_HD SamCoor sam_coordinate(const DevIndex& ix, int64_t g)
{
	SamCoor c; c.chrom = 0;
	const int n_chrom = ix.n_end >> 1;
	if (g < ix.G)
	{
		if (n_chrom == 1) { c.pos = g + 1; return c; }
		const int k = mc_chrom_lower_bound(ix, g);
		c.chrom = mc_ldg(ix.chrom_id + k);
		c.pos = g + 1 - (c.chrom == 0 ? 0 : mc_ldg(ix.chrom_end + c.chrom - 1) + 1);
		return c;
	}
	if (n_chrom == 1) { c.pos = ix.twoG - g; return c; }
	const int k = mc_chrom_lower_bound(ix, g);
	c.chrom = mc_ldg(ix.chrom_id + k); c.pos = mc_ldg(ix.chrom_end + k) - g + 1;
	return c;
}

// GetAlnCoordinate, src/SamReport.cpp:119-148: the first fragment with genome bases decides
MC_HD SamCoor sam_aln_coordinate(const SamArgs& a, int64_t cand, bool fwd)
{
	const int64_t f0 = a.cfrag[cand]; const int nf = a.cnfrag[cand];
	for (int f = 0; f < nf; f++)
	{
		const mc_frag_out& x = a.frags[f0 + f];
		if (x.gLen > 0) return sam_coordinate(a.ix, fwd ? x.gPos : x.gPos + x.gLen - 1);
	}
	SamCoor c; c.chrom = 0; c.pos = 0;   // the reference returns an uninitialised Coordinate_t here
	return c;
}

// "%d%c" of GenerateCIGARstring; p == nullptr only counts
MC_HD int sam_put(uint8_t* p, int at, int c, char op)
{
	char d[12]; int nd = 0;
	do { d[nd++] = (char)('0' + c % 10); c /= 10; } while (c > 0);
	if (p) { for (int k = 0; k < nd; k++) p[at + k] = (uint8_t)d[nd - 1 - k]; p[at + nd] = (uint8_t)op; }
	return at + nd + 1;
}

// GenerateCIGARstring, src/SamReport.cpp:171-316
MC_HD int sam_cigar(const SamArgs& a, int64_t cand, int rlen, bool fwd, uint8_t* p)
{
	const int64_t f0 = a.cfrag[cand]; const int nf = a.cnfrag[cand];
	int at = 0, c = 0; char state = ' ';
	if (nf <= 0) return 0;
	{
		const mc_frag_out& x = a.frags[f0];
		if (!x.bSimple)
		{
			if (fwd) { if (x.rPos != 0) at = sam_put(p, at, x.rPos, 'S'); }
			else { const int k = rlen - (x.rPos + x.rLen); if (k > 0) at = sam_put(p, at, k, 'S'); }
		}
	}
#define SAM_STATE(s) do { if (state != (s)) { if (c > 0) at = sam_put(p, at, c, state); state = (s); c = 0; } } while (0)
	for (int f = 0; f < nf; f++)
	{
		const mc_frag_out& x = a.frags[f0 + f];
		if (x.bSimple) { SAM_STATE('M'); c += x.rLen; }
		else if (x.aln_len > 0)
		{
			const uint8_t* a1 = a.aln + x.aln_off; const uint8_t* a2 = a1 + x.aln_cap;
			for (int j = 0; j < x.aln_len; j++)
			{
				if (a1[j] == '-') SAM_STATE('D'); else if (a2[j] == '-') SAM_STATE('I'); else SAM_STATE('M');
				c++;
			}
		}
		else if (x.rLen > 0) { SAM_STATE('I'); c += x.rLen; }
		else if (x.gLen > 0) { SAM_STATE('D'); c += x.gLen; }
	}
#undef SAM_STATE
	if (c > 0) at = sam_put(p, at, c, state);
	if (nf - 1 > 0 && !a.frags[f0 + nf - 1].bSimple)
	{
		const mc_frag_out& x = a.frags[f0 + nf - 1];
		if (fwd) { const int k = rlen - (x.rPos + x.rLen); if (k > 0) at = sam_put(p, at, k, 'S'); }
		else if (x.rPos != 0) at = sam_put(p, at, x.rPos, 'S');
	}
	return at;
}

MC_HD void samrec_body(int64_t r, const SamArgs& a, bool emit)
{
	const ReadSum rs = a.rsum[r];
	const int rlen = (int)(a.roff[r + 1] - a.roff[r]);
	const bool paired = a.paired != 0, first = !(r & 1);
	const int64_t m = paired ? (r ^ 1) : r;
	const int64_t co = a.cand_off[r];
	mc_sam_rec o; memset(&o, 0, sizeof(o));
	o.chrom = -1; o.nm = -1;
	if (rs.score == 0)   // :330-334, :385-399, :442-456
	{
		if (!paired) o.flag = 0x4;
		else
		{
			o.flag = 0x1 | 0x4 | (first ? 0x40 : 0x80);
			if (a.rsum[m].score == 0) o.flag |= 0x8; else if (a.ncand[m] > 0) o.flag |= 0x30;
		}
		o.reverse = (paired && !first) ? 1 : 0;   // an unmapped mate 2 is printed as it was mapped: reverse-complemented
		if (!emit) a.clen[r] = 0; else a.out[r] = o;
		return;
	}
	int i = rs.best_idx; const int nc = a.ncand[r];
	while (i < nc && a.cscore[co + i] != rs.score) i++;
	if (i >= nc) { o.flag = -1; if (!emit) a.clen[r] = 0; else a.out[r] = o; return; }   // the reference prints no line then
	const bool fwd = a.corient[co + i] == 1;
	if (!emit) { a.clen[r] = (uint32_t)sam_cigar(a, co + i, rlen, fwd, nullptr); return; }
	const int j = paired ? a.cpaired[co + i] : -1;
	const int64_t co2 = paired ? a.cand_off[m] : 0;
	if (!paired) o.flag = fwd ? 0 : 0x10;   // :7-24
	else                                       // :26-84
	{
		const bool unique = rs.score > rs.sub_score, ok = j != -1 && a.cscore[co2 + j] > 0;
		const int same = first ? (fwd ? 0x20 : 0x10) : (fwd ? 0x10 : 0x20), other = same ^ 0x30;
		o.flag = (first ? 0x41 : 0x81) | same;
		if (ok) o.flag |= 0x2; else { if (unique) o.flag |= other; o.flag |= 0x8; }
	}
	if (rs.score == rs.sub_score) o.mapq = 0;   // :86-101
	else if (rs.sub_score == 0 || rs.score - rs.sub_score > 5) o.mapq = 60;
	else o.mapq = a.mapq_tab[(rs.score - rs.sub_score - 1) * (MC_MAX_RLEN + 1) + rs.score];
	const SamCoor c1 = sam_aln_coordinate(a, co + i, fwd);
	o.chrom = c1.chrom; o.pos = c1.pos;
	// mate 2 is mapped (and printed) as the reverse complement of its FASTQ record (ReverseOrientation, src/ReadMapping.cpp:455)
	o.reverse = (paired && !first) ? (fwd ? 1 : 0) : (fwd ? 0 : 1);
	o.nm = rlen - a.cscore[co + i]; o.as = rs.score; o.xs = rs.sub_score;
	if (paired && j != -1 && a.rsum[m].score > 0 && a.cscore[co2 + j] == a.rsum[m].score)   // :428-431, :472-475
	{
		const bool mfwd = a.corient[co2 + j] == 1;
		const SamCoor c2 = sam_aln_coordinate(a, co2 + j, mfwd);
		const int mlen = (int)(a.roff[m + 1] - a.roff[m]);
		o.has_mate = 1; o.mate_pos = c2.pos;
		// both formulas are written from mate 1's side: coor2 - coor1 + (mate 1 forward ? rlen2 : -rlen1), negated for mate 2
		if (first) o.tlen = (int)(c2.pos - c1.pos + (fwd ? mlen : 0 - rlen));
		else o.tlen = 0 - (int)(c1.pos - c2.pos + (mfwd ? rlen : 0 - mlen));
	}
	o.cigar_off = (int32_t)a.coff[r]; o.cigar_len = (int32_t)a.clen[r];
	sam_cigar(a, co + i, rlen, fwd, a.cigar + a.coff[r]);
	a.out[r] = o;
}

#endif
