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

// ------------------------------------------------------------------------------------------------------------------
// SAM text on the device: the lines GenerateSingleSamStream / GeneratePairedSamStream sprintf (src/SamReport.cpp:324-488),
// assembled from the fields above and the batch's FASTQ text where mc_ingest_fastq left it in HBM (QNAME as
// IdentifyHeaderBegPos / IdentifyHeaderEndPos cut it, src/GetData.cpp:3-21; SEQ / QUAL reverse-complemented / reversed for a
// reverse-strand alignment, GetComplementarySeq src/tools.cpp:19).  One thread per read, two passes (line length, then the
// text at its scanned offset).  `unique` = bUnique: one line per read; otherwise (-m) one per candidate that reaches the
// read's best score.
struct SamTextArgs {
	SamArgs s;
	const uint8_t* text[2]; const int64_t* line_end[2];   // FASTQ text of the batch and its newline table (line_end[-1] = -1)
	int32_t two_files, unique, lpr;              // lpr: lines per record of the text (4 FASTQ, 2 FASTA: no qualities, QUAL is "*")
	const uint8_t* chrom_names; const int32_t* chrom_name_off;   // names back to back, n_chrom + 1 offsets
	uint32_t* tlen; const int64_t* toff; uint8_t* out;
};
struct SamWriter {
	uint8_t* p; int64_t at;
	MC_HD void ch(char c) { if (p) p[at] = (uint8_t)c; at++; }
	MC_HD void bytes(const uint8_t* s, int n) { if (p) for (int i = 0; i < n; i++) p[at + i] = s[i]; at += n; }
	MC_HD void lit(const char* s) { for (; *s; s++) ch(*s); }
	MC_HD void num(long long v)
	{
		char d[24]; int nd = 0; unsigned long long u = v < 0 ? 0ull - (unsigned long long)v : (unsigned long long)v;
		if (v < 0) ch('-');
		do { d[nd++] = (char)('0' + u % 10); u /= 10; } while (u > 0);
		if (p) for (int k = 0; k < nd; k++) p[at + k] = (uint8_t)d[nd - 1 - k];
		at += nd;
	}
};
// bases / qualities of a read as the line shows them
// `twice`: mate 2 on the reverse strand - it was reverse-complemented before mapping (ReverseOrientation) and is complemented
// back for the line, which folds lower case to upper and everything that is not ACGT to N on the way
MC_HD void sam_put_seq(SamWriter& w, const uint8_t* seq, int n, bool reverse, bool twice)
{
	if (w.p)
	{
		if (reverse) for (int i = 0; i < n; i++) w.p[w.at + i] = mc_complement(seq[n - 1 - i]);
		else if (twice) for (int i = 0; i < n; i++) w.p[w.at + i] = mc_complement(mc_complement(seq[i]));
		else for (int i = 0; i < n; i++) w.p[w.at + i] = seq[i];
	}
	w.at += n;
}
MC_HD void sam_put_qual(SamWriter& w, const uint8_t* q, int n, bool reverse)
{
	if (!q) { w.ch('*'); return; }            // FASTA reads: no qualities (src/SamReport.cpp:332,351)
	if (!reverse) { w.bytes(q, n); return; }
	if (w.p) for (int i = 0; i < n; i++) w.p[w.at + i] = q[n - 1 - i];
	w.at += n;
}

MC_HD void samtext_body(int64_t r, const SamTextArgs& t, bool emit)
{
	const SamArgs& a = t.s;
	SamWriter w; w.p = emit ? t.out + t.toff[r] : nullptr; w.at = 0;
	// the read's record in the FASTQ text: header line, bases, '+' line, qualities
	const int f = t.two_files ? (int)(r & 1) : 0;
	const int64_t rec = t.two_files ? (r >> 1) : r;
	const int64_t* le = t.line_end[f] + t.lpr * rec;
	const uint8_t* hdr = t.text[f] + le[-1] + 1; const int hlen = (int)(le[0] - le[-1]);   // including the newline, as getline() counts
	const uint8_t* seq = t.text[f] + le[0] + 1; const int rlen = (int)(le[1] - le[0] - 1);
	const uint8_t* qual = t.lpr == 4 ? t.text[f] + le[2] + 1 : nullptr;
	int p1 = hlen - 1, p2 = (hlen > 100 ? 100 : hlen) - 1;
	for (int i = 1; i < hlen; i++) if (hdr[i] != '>' && hdr[i] != '@') { p1 = i; break; }
	for (int i = 1; i < (hlen > 100 ? 100 : hlen); i++) if (hdr[i] == ' ' || hdr[i] == '/' || hdr[i] < 0x20 || hdr[i] > 0x7E) { p2 = i; break; }
	const int nlen = p2 > p1 ? p2 - p1 : 0;

	const ReadSum rs = a.rsum[r];
	const bool paired = a.paired != 0, first = !(r & 1);
	const int64_t m = paired ? (r ^ 1) : r;
	const int64_t co = a.cand_off[r];
	if (rs.score == 0)
	{
		int flag = 0x4;
		if (paired)
		{
			flag = 0x1 | 0x4 | (first ? 0x40 : 0x80);
			if (a.rsum[m].score == 0) flag |= 0x8; else if (a.ncand[m] > 0) flag |= 0x30;
		}
		const bool rev = paired && !first;      // an unmapped mate 2 is printed as it was mapped: reverse-complemented
		w.bytes(hdr + p1, nlen); w.ch('\t'); w.num(flag); w.lit("\t*\t0\t0\t*\t*\t0\t0\t");
		sam_put_seq(w, seq, rlen, rev, false); w.ch('\t'); sam_put_qual(w, qual, rlen, rev); w.lit("\tAS:i:0\tXS:i:0\n");
		if (!emit) t.tlen[r] = (uint32_t)w.at;
		return;
	}
	int mapq;
	if (rs.score == rs.sub_score) mapq = 0;
	else if (rs.sub_score == 0 || rs.score - rs.sub_score > 5) mapq = 60;
	else mapq = a.mapq_tab[(rs.score - rs.sub_score - 1) * (MC_MAX_RLEN + 1) + rs.score];
	const int nc = a.ncand[r];
	const int64_t co2 = paired ? a.cand_off[m] : 0;
	for (int i = rs.best_idx; i < nc; i++)
	{
		if (a.cscore[co + i] != rs.score) continue;
		const bool fwd = a.corient[co + i] == 1;
		const int j = paired ? a.cpaired[co + i] : -1;
		int flag;
		if (!paired) flag = fwd ? 0 : 0x10;
		else
		{
			const bool uniq = rs.score > rs.sub_score, ok = j != -1 && a.cscore[co2 + j] > 0;
			const int same = first ? (fwd ? 0x20 : 0x10) : (fwd ? 0x10 : 0x20), other = same ^ 0x30;
			flag = (first ? 0x41 : 0x81) | same;
			if (ok) flag |= 0x2; else { if (uniq) flag |= other; flag |= 0x8; }
		}
		const SamCoor c1 = sam_aln_coordinate(a, co + i, fwd);
		const bool rev = (paired && !first) ? fwd : !fwd;
		w.bytes(hdr + p1, nlen); w.ch('\t'); w.num(flag); w.ch('\t');
		w.bytes(t.chrom_names + t.chrom_name_off[c1.chrom], t.chrom_name_off[c1.chrom + 1] - t.chrom_name_off[c1.chrom]); w.ch('\t');
		w.num(c1.pos); w.ch('\t'); w.num(mapq); w.ch('\t');
		w.at += sam_cigar(a, co + i, rlen, fwd, w.p ? w.p + w.at : nullptr);
		if (paired && j != -1 && a.rsum[m].score > 0 && a.cscore[co2 + j] == a.rsum[m].score)
		{
			const bool mfwd = a.corient[co2 + j] == 1;
			const SamCoor c2 = sam_aln_coordinate(a, co2 + j, mfwd);
			const int mlen = (int)(a.roff[m + 1] - a.roff[m]);
			const int tl = first ? (int)(c2.pos - c1.pos + (fwd ? mlen : 0 - rlen)) : 0 - (int)(c1.pos - c2.pos + (mfwd ? rlen : 0 - mlen));
			w.lit("\t=\t"); w.num(c2.pos); w.ch('\t'); w.num(tl); w.ch('\t');
		}
		else w.lit("\t*\t0\t0\t");
		sam_put_seq(w, seq, rlen, rev, paired && !first && !rev); w.ch('\t'); sam_put_qual(w, qual, rlen, rev);
		w.lit("\tNM:i:"); w.num(rlen - a.cscore[co + i]); w.lit("\tAS:i:"); w.num(rs.score); w.lit("\tXS:i:"); w.num(rs.sub_score); w.ch('\n');
		if (t.unique) break;
	}
	if (!emit) t.tlen[r] = (uint32_t)w.at;
}

#endif
